// FP32 CUDA-core kernels of the forward graph (exact mode and the validation reference of the
// tcgen05 path).  Reference semantics: net.py:225-252 (conv_bn), net.py:286-313 (graph).
#pragma once
#include "ubd_common.cuh"

// ----------------------------------------------------------------------------------------------
// Separable layer (net.py:234-246 with separable=True): depthwise 3x3 (no bias / activation)
// -> pointwise 1x1 + bias -> ReLU.  Thread = one output pixel.
//   RAW:  input is the caller's image (n,H,W,CIN) of u8 / f32 with preprocessing folded in
//         (net.py:163-169,217-218); otherwise planar-by-4 feature map with CIN = 24.
//   pad_t / pad_l: zero rows/cols before the image (1,1 for the FML stride-2 layers net.py:229-232
//         and for 'same' stride 1; 0,0 for TF 'same' stride 2 on even sizes).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

template <int CIN, int STRIDE, bool RAW, typename TIn>
__global__ void __launch_bounds__(128)
sep_layer_kernel(const TIn* __restrict__ in, float4* __restrict__ out,
                 const float* __restrict__ dwk,   // [9][CIN]
                 const float* __restrict__ pwk,   // [CIN][24]
                 const float* __restrict__ bias,  // [24]
                 const float* __restrict__ lut,   // RAW u8: 256-entry preprocessing table (or null)
                 float pre_scale, float pre_shift,
                 int N, int H, int W, int Ho, int Wo, int pad_t, int pad_l, int in_mpad, int out_mpad,
                 float4* __restrict__ dwout = nullptr) {   // training: keep the depthwise output (pad 0) for the pointwise wgrad
  __shared__ float s_dw[9 * CIN];
  __shared__ __align__(16) float s_pw[CIN * UBD_NF];
  __shared__ float s_b[UBD_NF];
  __shared__ float s_lut[256];
  for (int i = threadIdx.x; i < 9 * CIN; i += blockDim.x) s_dw[i] = dwk[i];
  for (int i = threadIdx.x; i < CIN * UBD_NF; i += blockDim.x) s_pw[i] = pwk[i];
  for (int i = threadIdx.x; i < UBD_NF; i += blockDim.x) s_b[i] = bias[i];
  if (RAW && lut != nullptr)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();

  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (x >= Wo || y >= Ho) return;

  float d[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) d[c] = 0.f;

#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int iy = y * STRIDE + i - pad_t;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ix = x * STRIDE + j - pad_l;
      if (ix < 0 || ix >= W) continue;
      const float* wk = &s_dw[(i * 3 + j) * CIN];
      if constexpr (RAW) {
        const TIn* p = in + (((size_t)n * H + iy) * W + ix) * CIN;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          float v;
          if constexpr (sizeof(TIn) == 1) {
            v = (lut != nullptr) ? s_lut[(int)p[c]] : (float)p[c];
          } else {
            v = (float)p[c];
            if (pre_scale != 0.f) v = (v - pre_shift) / pre_scale;   // (x - 127.5) / 127.5
          }
          d[c] = fmaf(v, wk[c], d[c]);
        }
      } else {
        const float4* p = reinterpret_cast<const float4*>(in);
#pragma unroll
        for (int g = 0; g < CIN / 4; ++g) {
          const float4 v = ldg4(&p[act_index(n, g, iy, ix, H, W, in_mpad)]);
          d[4 * g + 0] = fmaf(v.x, wk[4 * g + 0], d[4 * g + 0]);
          d[4 * g + 1] = fmaf(v.y, wk[4 * g + 1], d[4 * g + 1]);
          d[4 * g + 2] = fmaf(v.z, wk[4 * g + 2], d[4 * g + 2]);
          d[4 * g + 3] = fmaf(v.w, wk[4 * g + 3], d[4 * g + 3]);
        }
      }
    }
  }

  if constexpr (!RAW && CIN == UBD_NF) {
    if (dwout != nullptr) {
#pragma unroll
      for (int g = 0; g < UBD_NG; ++g)
        dwout[act_index(n, g, y, x, Ho, Wo, 0)] = make_float4(d[4 * g], d[4 * g + 1], d[4 * g + 2], d[4 * g + 3]);
    }
  }
  const float4* pw4 = reinterpret_cast<const float4*>(s_pw);
#pragma unroll
  for (int og = 0; og < UBD_NG; ++og) {
    float4 a = make_float4(s_b[4 * og], s_b[4 * og + 1], s_b[4 * og + 2], s_b[4 * og + 3]);
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float4 w = pw4[c * UBD_NG + og];
      a.x = fmaf(d[c], w.x, a.x); a.y = fmaf(d[c], w.y, a.y);
      a.z = fmaf(d[c], w.z, a.z); a.w = fmaf(d[c], w.w, a.w);
    }
    a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    if (!RAW && pre_scale < 0.f) {     // producer-side rounding to the tf32 grid for the tensor-core layers
      a.x = rna_tf32(a.x); a.y = rna_tf32(a.y); a.z = rna_tf32(a.z); a.w = rna_tf32(a.w);
    }
    out[act_index(n, og, y, x, Ho, Wo, out_mpad)] = a;
  }
}

// ----------------------------------------------------------------------------------------------
// Dilated 3x3 conv 24->24 (net.py:298-304): out = act(bias + sum_taps in[y+(i-1)d, x+(j-1)d,:] k[i,j])
// cross-correlation, zero padding d.  Thread = 4 pixels (rows y, y+4, y+8, y+12) x 24 outputs;
// weights [tap][ic][oc] staged in shared memory and read as broadcast float4.
// Also used for the backward-data pass (flipped/transposed weights, no bias, gate by relu_src>0).
// ----------------------------------------------------------------------------------------------
#define DIL_TW 32
#define DIL_TH 16
#define DIL_P 4
template <bool RELU, bool BIAS, bool GATE>
__global__ void __launch_bounds__(128)
dilconv_fp32_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                    const float* __restrict__ wts,    // [9][24][24]
                    const float* __restrict__ bias,   // [24]
                    const float4* __restrict__ gate,  // same layout as out (GATE): out *= (gate > 0)
                    int N, int H, int W, int d, int mpad) {
  extern __shared__ __align__(16) float s_w[];        // 9*24*24 floats
  for (int i = threadIdx.x; i < 9 * UBD_NF * UBD_NF / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wts) + i);
  __syncthreads();

  const int x = blockIdx.x * DIL_TW + (threadIdx.x & 31);
  const int y0 = blockIdx.y * DIL_TH + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  const bool xin = x < W;

  float acc[DIL_P][UBD_NF];
#pragma unroll
  for (int p = 0; p < DIL_P; ++p)
#pragma unroll
    for (int o = 0; o < UBD_NF; ++o) acc[p][o] = BIAS ? __ldg(&bias[o]) : 0.f;

  const float4* w4 = reinterpret_cast<const float4*>(s_w);
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = (tap / 3 - 1) * d, dx = (tap % 3 - 1) * d;
    const int xx = x + dx;
    const bool xok = xin && xx >= 0 && xx < W;
#pragma unroll
    for (int g = 0; g < UBD_NG; ++g) {
      float4 v[DIL_P];
#pragma unroll
      for (int p = 0; p < DIL_P; ++p) {
        const int yy = y0 + 4 * p + dy;
        v[p] = (xok && yy >= 0 && yy < H) ? ldg4(&in[act_index(n, g, yy, xx, H, W, mpad)])
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int icc = 0; icc < 4; ++icc) {
        const float4* wrow = &w4[(tap * UBD_NF + g * 4 + icc) * UBD_NG];
#pragma unroll
        for (int og = 0; og < UBD_NG; ++og) {
          const float4 w = wrow[og];
#pragma unroll
          for (int p = 0; p < DIL_P; ++p) {
            const float a = icc == 0 ? v[p].x : icc == 1 ? v[p].y : icc == 2 ? v[p].z : v[p].w;
            acc[p][4 * og + 0] = fmaf(a, w.x, acc[p][4 * og + 0]);
            acc[p][4 * og + 1] = fmaf(a, w.y, acc[p][4 * og + 1]);
            acc[p][4 * og + 2] = fmaf(a, w.z, acc[p][4 * og + 2]);
            acc[p][4 * og + 3] = fmaf(a, w.w, acc[p][4 * og + 3]);
          }
        }
      }
    }
  }

#pragma unroll
  for (int p = 0; p < DIL_P; ++p) {
    const int y = y0 + 4 * p;
    if (!xin || y >= H) continue;
#pragma unroll
    for (int og = 0; og < UBD_NG; ++og) {
      float4 a = make_float4(acc[p][4 * og], acc[p][4 * og + 1], acc[p][4 * og + 2], acc[p][4 * og + 3]);
      if (RELU) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
      const size_t idx = act_index(n, og, y, x, H, W, mpad);
      if (GATE) {
        const float4 gt = ldg4(&gate[idx]);
        a.x = gt.x > 0.f ? a.x : 0.f; a.y = gt.y > 0.f ? a.y : 0.f;
        a.z = gt.z > 0.f ? a.z : 0.f; a.w = gt.w > 0.f ? a.w : 0.f;
      }
      out[idx] = a;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Head (net.py:307-311: Conv2D(1+C,(1,1)), linear) + logit threshold (model_runner.py:121-124:
// strict '>', float32 compare).  logits NHWC (n,h,w,1+C) as model.predict returns them.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
head_threshold_kernel(const float4* __restrict__ in, float* __restrict__ logits,
                      uint8_t* __restrict__ mask, const float* __restrict__ hk,   // [24][n_out]
                      const float* __restrict__ hb, int n_out, float thr, int N, int H, int W, int mpad) {
  __shared__ float s_k[UBD_NF * (1 + UBD_MAX_CLASSES)];
  __shared__ float s_b[1 + UBD_MAX_CLASSES];
  for (int i = threadIdx.x; i < UBD_NF * n_out; i += blockDim.x) s_k[i] = hk[i];
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) s_b[i] = hb[i];
  __syncthreads();
  const size_t npx = (size_t)H * W;
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= npx) return;
  const int y = (int)(p / W), x = (int)(p % W);
  float v[UBD_NF];
#pragma unroll
  for (int g = 0; g < UBD_NG; ++g) {
    const float4 a = ldg4(&in[act_index(n, g, y, x, H, W, mpad)]);
    v[4 * g] = a.x; v[4 * g + 1] = a.y; v[4 * g + 2] = a.z; v[4 * g + 3] = a.w;
  }
  float* lo = logits ? logits + ((size_t)n * npx + p) * n_out : nullptr;
  for (int o = 0; o < n_out; ++o) {
    float a = s_b[o];
#pragma unroll
    for (int c = 0; c < UBD_NF; ++c) a = fmaf(v[c], s_k[c * n_out + o], a);
    if (lo) lo[o] = a;
    if (o == 0 && mask) mask[(size_t)n * npx + p] = a > thr ? 1 : 0;
  }
}
