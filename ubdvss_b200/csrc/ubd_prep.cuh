// Input side of the path on the GPU: what SegmapManager._rescale_image_and_markup and
// BatchGenerator._prepare_image do to a decoded image before the network (segmap_manager.py:136-167:
// ``image.resize((new_w, new_h), Image.BICUBIC)``; data_generators.py:176-177: ``image.convert('L')``).
//
// Pillow resamples 8-bit images in fixed point (Resample.c): per output coordinate a window [xmin, xmin + n) of
// the input and n coefficients of the bicubic kernel (a = -0.5), stretched by the scale factor when shrinking
// (antialiasing), normalised in double and rounded to 22 fractional bits; one pass per axis (horizontal first),
// each output = clip8((2^21 + sum(pixel * coeff)) >> 22), the horizontal result stored as uint8.  The tables are
// built on the host with Pillow's expressions (ubd_api.cu, prep_coeffs) so that the kernels below - plain integer
// dot products - reproduce Pillow bit for bit.  ``convert('L')`` is (19595 R + 38470 G + 7471 B + 0x8000) >> 16.
#pragma once
#include "ubd_common.cuh"

struct PrepAxis { const int* bounds; const int* coeffs; int ksize; };    // bounds[2*i] = xmin, [2*i+1] = n

__device__ __forceinline__ uint8_t prep_clip8(int v) {
  v >>= 22;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in (n, H, W, C) -> out (n, H, Wo, C)
__global__ void __launch_bounds__(256)
prep_resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, PrepAxis ax, int n, int H, int W, int Wo, int C) {
  const size_t total = (size_t)n * H * Wo * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int xo = (int)((i / C) % Wo);
    const size_t row = i / ((size_t)C * Wo);
    const int x0 = ax.bounds[2 * xo], cnt = ax.bounds[2 * xo + 1];
    const int* k = ax.coeffs + (size_t)xo * ax.ksize;
    const uint8_t* p = in + (row * W + x0) * C + c;
    int ss = 1 << 21;
    for (int j = 0; j < cnt; ++j) ss += (int)p[(size_t)j * C] * k[j];
    out[i] = prep_clip8(ss);
  }
}

// in (n, H, W, C) -> out (n, Ho, W, C)
__global__ void __launch_bounds__(256)
prep_resize_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, PrepAxis ax, int n, int H, int Ho, int W, int C) {
  const size_t rowlen = (size_t)W * C;
  const size_t total = (size_t)n * Ho * rowlen;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t xc = i % rowlen;
    const int yo = (int)((i / rowlen) % Ho);
    const size_t img = i / (rowlen * Ho);
    const int y0 = ax.bounds[2 * yo], cnt = ax.bounds[2 * yo + 1];
    const int* k = ax.coeffs + (size_t)yo * ax.ksize;
    const uint8_t* p = in + (img * H + y0) * rowlen + xc;
    int ss = 1 << 21;
    for (int j = 0; j < cnt; ++j) ss += (int)p[(size_t)j * rowlen] * k[j];
    out[i] = prep_clip8(ss);
  }
}

// RGB (n, H, W, 3) -> L (n, H, W): ImagingConvert rgb2l
__global__ void __launch_bounds__(256)
prep_rgb2l_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t npx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
    const uint8_t* p = in + 3 * i;
    out[i] = (uint8_t)(((unsigned)p[0] * 19595u + (unsigned)p[1] * 38470u + (unsigned)p[2] * 7471u + 0x8000u) >> 16);
  }
}
