// Training kernels (FP32 CUDA cores): pixelwise loss with gradient (losses.py:27-126), backward of
// every layer of net.py:286-313, Keras-2 Adam (train.py:110).  All cross-block reductions go through
// a per-block partials buffer that is summed in a fixed order, so a step is bit-reproducible.
#pragma once
#include "ubd_handle.cuh"

namespace tr {

// ------------------------------------------------------------------------------------------------
// loss
// ------------------------------------------------------------------------------------------------
constexpr int LOSS_RANGE = 4096;      // consecutive pixels per block (index-ordered tie handling)
constexpr float KERAS_EPS = 1e-7f;    // K.epsilon()

struct LossState {                    // device-resident scalars of one loss evaluation
  double n_pos, n_neg, sum_pos, sum_neg, sum_cls, sum_gt;
  double npos_c, nneg_c;              // clamped to >= 1 (losses.py:99,105)
  double pos, neg, hard, cls, loss;
  int k, k_rem;                       // losses.py:110; ties still to take at the threshold value
  unsigned prefix;                    // radix-select state: bits of the k-th largest value
  unsigned hist[256];
};

// K.binary_crossentropy(target, sigmoid(z)) of TF1-Keras in float32: clip p to [eps, 1-eps], re-logit,
// sigmoid_cross_entropy_with_logits.  `open` = the clip passes the gradient.
__device__ __forceinline__ float keras_bce(float z, float t, float* x_out, bool* open) {
  const float p = 1.f / (1.f + expf(-z));
  const float lo = KERAS_EPS, hi = 1.f - KERAS_EPS;
  const float pc = fminf(fmaxf(p, lo), hi);
  const float x = logf(pc / (1.f - pc));
  *x_out = x;
  *open = p >= lo && p <= hi;
  return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red /*[32]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T s = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;
}

// per pixel: ce, v = ce*(1-t) (the top-k input, losses.py:104); per block partial sums
__global__ void __launch_bounds__(256)
loss_pixel_kernel(const float* __restrict__ logits, const int* __restrict__ y_true, int n_out, size_t P,
                  float* __restrict__ v, double* __restrict__ partials /*[grid][4]*/) {
  __shared__ double red[32];
  const size_t base = (size_t)blockIdx.x * LOSS_RANGE;
  double npos = 0, spos = 0, sneg = 0, scls = 0;
  for (size_t p = base + threadIdx.x; p < min(base + (size_t)LOSS_RANGE, P); p += blockDim.x) {
    const float* lg = logits + p * n_out;
    const int yt = y_true[p];
    const float t = yt > 0 ? 1.f : 0.f;
    float x; bool open;
    const float ce = keras_bce(lg[0], t, &x, &open);
    v[p] = ce * (1.f - t);
    npos += t; spos += (double)(ce * t); sneg += (double)(ce * (1.f - t));
    if (n_out > 1 && yt > 0) {                       // losses.py:65-83, masked sparse softmax CE
      float mx = lg[1];
      for (int c = 2; c < n_out; ++c) mx = fmaxf(mx, lg[c]);
      float s = 0.f;
      for (int c = 1; c < n_out; ++c) s += expf(lg[c] - mx);
      // label = yt - 1 -> channel yt; a label outside the class head gives NaN, as tf.nn.sparse_softmax_cross_entropy_with_logits
      // does on the GPU (losses.py:78), instead of reading past the pixel's logits
      scls += yt < n_out ? (double)(logf(s) + mx - lg[yt]) : (double)__int_as_float(0x7fc00000);
    }
  }
  npos = block_sum(npos, red); spos = block_sum(spos, red); sneg = block_sum(sneg, red); scls = block_sum(scls, red);
  if (threadIdx.x == 0) {
    double* o = partials + (size_t)blockIdx.x * 4;
    o[0] = npos; o[1] = spos; o[2] = sneg; o[3] = scls;
  }
}

// Pixel statistics behind the training metrics (keras_metrics.py:116-191): confusion counts of the detection
// channel (prediction = logit > 0, truth = y_true > 0) and, over the pixels of objects, how often the arg-max
// class (first maximum, as tf.argmax) equals the label y_true - 1.  counts: tp, tn, fp, fn, cls_correct, cls_total.
__global__ void __launch_bounds__(256)
metric_counts_kernel(const float* __restrict__ logits, const int* __restrict__ y_true, int n_out, size_t P,
                     unsigned long long* __restrict__ counts) {
  unsigned c[6] = {0u, 0u, 0u, 0u, 0u, 0u};
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float* lg = logits + p * n_out;
    const int yt = y_true[p];
    const bool t = yt > 0, d = lg[0] > 0.f;
    c[t ? (d ? 0 : 3) : (d ? 2 : 1)] += 1u;
    if (n_out > 1 && t) {
      int best = 1;
      for (int k = 2; k < n_out; ++k) if (lg[k] > lg[best]) best = k;
      c[4] += best == yt ? 1u : 0u;
      c[5] += 1u;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const unsigned w = __reduce_add_sync(0xffffffffu, c[i]);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(counts + i, (unsigned long long)w);
  }
}

__global__ void loss_reduce_kernel(const double* __restrict__ partials, int nblocks, size_t P, LossState* st) {
  if (threadIdx.x != 0) { if (threadIdx.x < 256) st->hist[threadIdx.x] = 0; return; }
  st->hist[0] = 0;
  double a = 0, b = 0, c = 0, d = 0;
  for (int i = 0; i < nblocks; ++i) { a += partials[4 * i]; b += partials[4 * i + 1]; c += partials[4 * i + 2]; d += partials[4 * i + 3]; }
  st->n_pos = a; st->n_neg = (double)P - a; st->sum_pos = b; st->sum_neg = c; st->sum_cls = d;
  st->npos_c = fmax(a, 1.0); st->nneg_c = fmax((double)P - a, 1.0);
  st->pos = b / st->npos_c; st->neg = c / st->nneg_c; st->cls = d / st->npos_c;
  st->k = (int)fmin(st->npos_c, st->nneg_c);
  st->k_rem = st->k; st->prefix = 0u; st->sum_gt = 0;
}

// radix select of the k-th largest v (all v >= 0, so the uint32 order of the bit patterns is the
// float order): 4 passes of 8 bits, most significant first.
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ v, size_t P, int pass, LossState* st) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const unsigned mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
  const unsigned prefix = st->prefix;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const unsigned b = __float_as_uint(v[p]);
    if ((b & mask) == prefix) atomicAdd(&h[(b >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void select_pick_kernel(int pass, LossState* st) {
  if (threadIdx.x == 0) {
    const int shift = 24 - 8 * pass;
    int k = st->k_rem;
    int d = 255;
    for (; d > 0; --d) {
      const int c = (int)st->hist[d];
      if (k <= c) break;
      k -= c;
    }
    st->k_rem = k;                 // rank inside digit d (>= 1)
    st->prefix |= (unsigned)d << shift;
  }
  __syncthreads();
  st->hist[threadIdx.x] = 0;
}

// per block (a contiguous index range): number of ties at the threshold, sum of values above it
__global__ void __launch_bounds__(256)
tie_count_kernel(const float* __restrict__ v, size_t P, const LossState* st, int* __restrict__ block_ties,
                 double* __restrict__ partials /*[grid]*/) {
  __shared__ double red[32];
  __shared__ int redi[32];
  const unsigned tau = st->prefix;
  const size_t base = (size_t)blockIdx.x * LOSS_RANGE;
  int ties = 0;
  double sgt = 0;
  for (size_t p = base + threadIdx.x; p < min(base + (size_t)LOSS_RANGE, P); p += blockDim.x) {
    const unsigned b = __float_as_uint(v[p]);
    ties += b == tau;
    if (b > tau) sgt += (double)v[p];
  }
  sgt = block_sum(sgt, red);
  ties = block_sum(ties, redi);
  if (threadIdx.x == 0) { block_ties[blockIdx.x] = ties; partials[blockIdx.x] = sgt; }
}

__global__ void tie_scan_kernel(int* __restrict__ block_ties, const double* __restrict__ partials, int nblocks,
                                LossState* st, int classification, float* __restrict__ parts_out /*[6]*/) {
  if (threadIdx.x != 0) return;
  int run = 0;
  double sgt = 0;
  for (int i = 0; i < nblocks; ++i) { const int c = block_ties[i]; block_ties[i] = run; run += c; sgt += partials[i]; }
  const double tau = (double)__uint_as_float(st->prefix);
  double hard = (sgt + (double)st->k_rem * tau) / (double)st->k;         // mean of the top k (losses.py:116)
  if (hard != hard) hard = 0.0;                                           // losses.py:117-121
  st->hard = hard;
  const double det = 15.0 * st->pos + 1.0 * st->neg + 5.0 * hard;          // losses.py:13-15,123-125
  st->loss = classification ? 1.0 * det + 1.0 * st->cls : det;           // losses.py:16-17,60
  parts_out[0] = (float)st->loss; parts_out[1] = (float)st->pos; parts_out[2] = (float)st->neg;
  parts_out[3] = (float)hard; parts_out[4] = classification ? (float)st->cls : 0.f; parts_out[5] = (float)st->k;
}

// dL/dlogits.  Ties at the threshold are taken in flat index order (tf.nn.top_k).
__global__ void __launch_bounds__(256)
loss_grad_kernel(const float* __restrict__ logits, const int* __restrict__ y_true, const float* __restrict__ v,
                 int n_out, size_t P, const LossState* st, const int* __restrict__ block_tie_offset,
                 int classification, float* __restrict__ dlogits) {
  __shared__ int warp_cnt[8];
  __shared__ int running;
  const unsigned tau = st->prefix;
  const int r = st->k_rem;
  const float wpos = (float)(15.0 / st->npos_c), wneg = (float)(1.0 / st->nneg_c), whard = (float)(5.0 / (double)st->k);
  const float wcls = (float)(1.0 / st->npos_c);
  if (threadIdx.x == 0) running = block_tie_offset[blockIdx.x];
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * LOSS_RANGE;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (size_t p0 = base; p0 < min(base + (size_t)LOSS_RANGE, P); p0 += blockDim.x) {
    const size_t p = p0 + threadIdx.x;
    const bool in = p < P;
    const unsigned b = in ? __float_as_uint(v[p]) : 0u;
    const bool tie = in && b == tau;
    const unsigned bits = __ballot_sync(0xffffffffu, tie);
    if (lane == 0) warp_cnt[wid] = __popc(bits);
    __syncthreads();
    int off = running;
    for (int i = 0; i < wid; ++i) off += warp_cnt[i];
    const int rank = off + __popc(bits & ((1u << lane) - 1u));
    if (in) {
      const float* lg = logits + p * n_out;
      const int yt = y_true[p];
      const float t = yt > 0 ? 1.f : 0.f;
      float x; bool open;
      keras_bce(lg[0], t, &x, &open);
      const bool sel = b > tau || (tie && rank < r);
      const float dce = open ? (1.f / (1.f + expf(-x)) - t) : 0.f;
      const float wgt = wpos * t + (1.f - t) * (wneg + (sel ? whard : 0.f));
      float* dl = dlogits + p * n_out;
      dl[0] = wgt * dce;
      if (n_out > 1) {
        if (classification && yt > 0) {
          float mx = lg[1];
          for (int c = 2; c < n_out; ++c) mx = fmaxf(mx, lg[c]);
          float s = 0.f;
          for (int c = 1; c < n_out; ++c) s += expf(lg[c] - mx);
          for (int c = 1; c < n_out; ++c) dl[c] = (expf(lg[c] - mx) / s - (c == yt ? 1.f : 0.f)) * wcls;
        } else {
          for (int c = 1; c < n_out; ++c) dl[c] = 0.f;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int i = 0; i < 8; ++i) tot += warp_cnt[i]; running += tot; }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// X^T Y reductions: out[kx][o] = sum_px X[px][kx] * Y[px][o], colsum[o] = sum_px Y[px][o]
// (all weight gradients of the dense layers).  Block = 256 threads, tiles of 32 pixels in smem.
// ------------------------------------------------------------------------------------------------
constexpr int XTY_TP = 32;

struct PixelGeom { int N, H, W, pad; };     // map the X / Y rows live on

struct XDil {          // X[px][tap*24 + c] = a_in[y + dy*d, x + dx*d, c]   (dilated conv wgrad, KX = 216)
  const float4* a; int d;
  __device__ float4 get4(const PixelGeom& g, int n, int y, int x, int k4) const {
    const int tap = k4 / UBD_NG, pl = k4 % UBD_NG;
    const int yy = y + (tap / 3 - 1) * d, xx = x + (tap % 3 - 1) * d;      // x halo lands in the zero pad
    if (yy < 0 || yy >= g.H) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(&a[act_index(n, pl, yy, xx, g.H, g.W, g.pad)]);
  }
};
struct XMap {          // X[px][c] = map[y, x, c]                          (head / plain 1x1, KX = 24)
  const float4* a;
  __device__ float4 get4(const PixelGeom& g, int n, int y, int x, int k4) const {
    return __ldg(&a[act_index(n, k4, y, x, g.H, g.W, g.pad)]);
  }
};
struct XDw {           // X[px][c] = depthwise3x3(x_in)[px][c]               (pointwise wgrad of a separable layer)
  const float4* xin; const float* dwk; int Hi, Wi, ipad, stride, pad_t, pad_l;
  __device__ float4 get4(const PixelGeom& g, int n, int y, int x, int k4) const {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int iy = y * stride + i - pad_t;
      if (iy < 0 || iy >= Hi) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int ix = x * stride + j - pad_l;
        if (ix < 0 || ix >= Wi) continue;
        const float4 v = __ldg(&xin[act_index(n, k4, iy, ix, Hi, Wi, ipad)]);
        const float* wk = dwk + (i * 3 + j) * UBD_NF + 4 * k4;
        s.x = fmaf(v.x, wk[0], s.x); s.y = fmaf(v.y, wk[1], s.y); s.z = fmaf(v.z, wk[2], s.z); s.w = fmaf(v.w, wk[3], s.w);
      }
    }
    return s;
  }
};
struct YMap {          // Y = a 24-channel gradient map
  const float4* gmap;
  __device__ float get(const PixelGeom& g, int n, int y, int x, size_t p, int o) const {
    const float4 v = __ldg(&gmap[act_index(n, o >> 2, y, x, g.H, g.W, g.pad)]);
    return (o & 3) == 0 ? v.x : (o & 3) == 1 ? v.y : (o & 3) == 2 ? v.z : v.w;
  }
};
struct YNhwc {         // Y = dL/dlogits (N,h,w,n_out)
  const float* dl; int n_out;
  __device__ float get(const PixelGeom&, int, int, int, size_t p, int o) const { return __ldg(&dl[p * n_out + o]); }
};

template <int KX, class XL, class YL>
__global__ void __launch_bounds__(256)
xty_kernel(XL xl, YL yl, PixelGeom g, int KY, float* __restrict__ partials /*[grid][KX*KY + KY]*/) {
  constexpr int MAXACC = (KX * 32 + 255) / 256;
  extern __shared__ float sm[];
  float* sX = sm;                                 // [TP][KX]
  float* sY = sm + XTY_TP * KX;                   // [TP][KY]
  const size_t P = (size_t)g.N * g.H * g.W;
  const int nout = KX * KY;
  float acc[MAXACC];
#pragma unroll
  for (int a = 0; a < MAXACC; ++a) acc[a] = 0.f;
  float colsum = 0.f;
  const size_t ntiles = (P + XTY_TP - 1) / XTY_TP;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t p0 = tile * XTY_TP;
    for (int i = threadIdx.x; i < XTY_TP * (KX / 4); i += blockDim.x) {
      const int px = i / (KX / 4), k4 = i % (KX / 4);
      const size_t p = p0 + px;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < P) {
        const int x = (int)(p % g.W), y = (int)((p / g.W) % g.H), n = (int)(p / ((size_t)g.W * g.H));
        v = xl.get4(g, n, y, x, k4);
      }
      float* d = sX + px * KX + 4 * k4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    for (int i = threadIdx.x; i < XTY_TP * KY; i += blockDim.x) {
      const int px = i / KY, o = i % KY;
      const size_t p = p0 + px;
      float v = 0.f;
      if (p < P) {
        const int x = (int)(p % g.W), y = (int)((p / g.W) % g.H), n = (int)(p / ((size_t)g.W * g.H));
        v = yl.get(g, n, y, x, p, o);
      }
      sY[px * KY + o] = v;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < MAXACC; ++a) {
      const int idx = threadIdx.x + a * 256;
      if (idx < nout) {
        const int kx = idx / KY, o = idx % KY;
        float s = acc[a];
#pragma unroll 8
        for (int px = 0; px < XTY_TP; ++px) s = fmaf(sX[px * KX + kx], sY[px * KY + o], s);
        acc[a] = s;
      }
    }
    if ((int)threadIdx.x < KY)
      for (int px = 0; px < XTY_TP; ++px) colsum += sY[px * KY + threadIdx.x];
    __syncthreads();
  }
  float* o = partials + (size_t)blockIdx.x * (nout + KY);
#pragma unroll
  for (int a = 0; a < MAXACC; ++a) {
    const int idx = threadIdx.x + a * 256;
    if (idx < nout) o[idx] = acc[a];
  }
  if ((int)threadIdx.x < KY) o[nout + threadIdx.x] = colsum;
}

// Weight + bias gradient of a one-output head (detection only: net.py:307-311 with C = 0): dHK[c] = sum_px a[px][c] *
// dl[px], dHB = sum_px dl[px].  One pass over the last map; per-block partial sums in a fixed order.
__global__ void __launch_bounds__(256)
head_wgrad1_kernel(const float4* __restrict__ a, const float* __restrict__ dl, int N, int H, int W, int mpad,
                   float* __restrict__ partials /*[grid][25]*/) {
  __shared__ float red[8][UBD_NF + 1];
  float acc[UBD_NF + 1];
#pragma unroll
  for (int c = 0; c <= UBD_NF; ++c) acc[c] = 0.f;
  const size_t P = (size_t)N * H * W;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(p % W), y = (int)((p / W) % H), n = (int)(p / ((size_t)W * H));
    const float d = __ldg(&dl[p]);
#pragma unroll
    for (int g = 0; g < UBD_NG; ++g) {
      const float4 v = __ldg(&a[act_index(n, g, y, x, H, W, mpad)]);
      acc[4 * g] = fmaf(v.x, d, acc[4 * g]); acc[4 * g + 1] = fmaf(v.y, d, acc[4 * g + 1]);
      acc[4 * g + 2] = fmaf(v.z, d, acc[4 * g + 2]); acc[4 * g + 3] = fmaf(v.w, d, acc[4 * g + 3]);
    }
    acc[UBD_NF] += d;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c <= UBD_NF; ++c) {
    float v = acc[c];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[wid][c] = v;
  }
  __syncthreads();
  if (threadIdx.x <= UBD_NF) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    partials[(size_t)blockIdx.x * (UBD_NF + 1) + threadIdx.x] = v;
  }
}

// out[i] = sum_b partials[b][i] in a fixed order (deterministic); two destinations (kernel, bias).  Eight lanes share an
// output: lane l adds the blocks b = l, l + 8, ... in order, then the eight sums are combined as a fixed tree.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int nblocks, int stride,
                                       float* __restrict__ dst0, int n0, float* __restrict__ dst1, int n1) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 3, l = t & 7;
  float s = 0.f;
  if (i < n0 + n1)
    for (int b = l; b < nblocks; b += 8) s += partials[(size_t)b * stride + i];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (i < n0 + n1 && l == 0) { if (i < n0) dst0[i] = s; else dst1[i - n0] = s; }
}

// ------------------------------------------------------------------------------------------------
// head backward (data): g9[c] = (a9[c] > 0) * sum_o dlogits[o] * hk[c][o]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
head_bwd_data_kernel(const float* __restrict__ dlogits, const float4* __restrict__ a9, float4* __restrict__ g9,
                     const float* __restrict__ hk, int n_out, int N, int H, int W, int mpad, int rnd) {
  __shared__ float s_k[UBD_NF * (1 + UBD_MAX_CLASSES)];
  for (int i = threadIdx.x; i < UBD_NF * n_out; i += blockDim.x) s_k[i] = hk[i];
  __syncthreads();
  const size_t npx = (size_t)H * W;
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= npx) return;
  const int y = (int)(p / W), x = (int)(p % W);
  const float* dl = dlogits + ((size_t)n * npx + p) * n_out;
  float g[UBD_NF];
#pragma unroll
  for (int c = 0; c < UBD_NF; ++c) g[c] = 0.f;
  for (int o = 0; o < n_out; ++o) {
    const float d = dl[o];
#pragma unroll
    for (int c = 0; c < UBD_NF; ++c) g[c] = fmaf(d, s_k[c * n_out + o], g[c]);
  }
#pragma unroll
  for (int pl = 0; pl < UBD_NG; ++pl) {
    const size_t idx = act_index(n, pl, y, x, H, W, mpad);
    const float4 a = __ldg(&a9[idx]);
    float4 o = make_float4(a.x > 0.f ? g[4 * pl] : 0.f, a.y > 0.f ? g[4 * pl + 1] : 0.f,
                           a.z > 0.f ? g[4 * pl + 2] : 0.f, a.w > 0.f ? g[4 * pl + 3] : 0.f);
    if (rnd) {        // the tensor-core backward reads this map: round (magnitude, half up) to the tf32 grid instead of truncating
      o.x = __uint_as_float((__float_as_uint(o.x) + 0x1000u) & 0xFFFFE000u); o.y = __uint_as_float((__float_as_uint(o.y) + 0x1000u) & 0xFFFFE000u);
      o.z = __uint_as_float((__float_as_uint(o.z) + 0x1000u) & 0xFFFFE000u); o.w = __uint_as_float((__float_as_uint(o.w) + 0x1000u) & 0xFFFFE000u);
    }
    g9[idx] = o;
  }
}

// K'[l][tap'][oc][ic] = W[l][8 - tap'][ic][oc]: the backward-data pass of a dilated layer is the same
// conv with the kernel flipped and transposed.
__global__ void build_wflip_kernel(const float* __restrict__ params, const int64_t* __restrict__ koff,
                                   float* __restrict__ wflip /*[6][9*24*24]*/) {
  const int l = blockIdx.x;
  const float* K = params + koff[l];
  float* D = wflip + (size_t)l * 9 * UBD_NF * UBD_NF;
  for (int i = threadIdx.x; i < 9 * UBD_NF * UBD_NF; i += blockDim.x) {
    const int tap = i / (UBD_NF * UBD_NF), oc = (i / UBD_NF) % UBD_NF, ic = i % UBD_NF;
    D[i] = K[((8 - tap) * UBD_NF + ic) * UBD_NF + oc];
  }
}

// ------------------------------------------------------------------------------------------------
// separable layers, backward
// ------------------------------------------------------------------------------------------------
// g_d[px][c] = sum_o g_y[px][o] * pw[c][o]     (through the pointwise conv)
__global__ void __launch_bounds__(128)
pw_bwd_data_kernel(const float4* __restrict__ gy, float4* __restrict__ gd, const float* __restrict__ pwk,
                   int N, int H, int W, int pad_in, int pad_out) {
  __shared__ float s_pw[UBD_NF * UBD_NF];
  for (int i = threadIdx.x; i < UBD_NF * UBD_NF; i += blockDim.x) s_pw[i] = pwk[i];
  __syncthreads();
  const size_t npx = (size_t)H * W;
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= npx) return;
  const int y = (int)(p / W), x = (int)(p % W);
  float g[UBD_NF];
#pragma unroll
  for (int pl = 0; pl < UBD_NG; ++pl) {
    const float4 v = __ldg(&gy[act_index(n, pl, y, x, H, W, pad_in)]);
    g[4 * pl] = v.x; g[4 * pl + 1] = v.y; g[4 * pl + 2] = v.z; g[4 * pl + 3] = v.w;
  }
#pragma unroll
  for (int pl = 0; pl < UBD_NG; ++pl) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < UBD_NF; ++k) {
      o.x = fmaf(g[k], s_pw[(4 * pl + 0) * UBD_NF + k], o.x);
      o.y = fmaf(g[k], s_pw[(4 * pl + 1) * UBD_NF + k], o.y);
      o.z = fmaf(g[k], s_pw[(4 * pl + 2) * UBD_NF + k], o.z);
      o.w = fmaf(g[k], s_pw[(4 * pl + 3) * UBD_NF + k], o.w);
    }
    gd[act_index(n, pl, y, x, H, W, pad_out)] = o;
  }
}

// dDW[tap][c] = sum_px x[px*s + tap - pad][c] * g_d[px][c].  Warp = one plane of 32 consecutive pixels.
__global__ void __launch_bounds__(192)
dw_bwd_weight_kernel(const float4* __restrict__ xin, const float4* __restrict__ gd, int N, int Hi, int Wi, int ipad,
                     int Ho, int Wo, int opad, int stride, int pad_t, int pad_l,
                     float* __restrict__ partials /*[grid][216]*/) {
  const int pl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int wtiles = (Wo + 31) / 32;
  const size_t ntiles = (size_t)N * Ho * wtiles;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int x = (int)(tile % wtiles) * 32 + lane;
    const int y = (int)((tile / wtiles) % Ho), n = (int)(tile / ((size_t)wtiles * Ho));
    if (x >= Wo) continue;
    const float4 g = __ldg(&gd[act_index(n, pl, y, x, Ho, Wo, opad)]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int iy = y * stride + i - pad_t;
      if (iy < 0 || iy >= Hi) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int ix = x * stride + j - pad_l;
        if (ix < 0 || ix >= Wi) continue;
        const float4 v = __ldg(&xin[act_index(n, pl, iy, ix, Hi, Wi, ipad)]);
        float4& a = acc[i * 3 + j];
        a.x = fmaf(v.x, g.x, a.x); a.y = fmaf(v.y, g.y, a.y); a.z = fmaf(v.z, g.z, a.z); a.w = fmaf(v.w, g.w, a.w);
      }
    }
  }
  float* o = partials + (size_t)blockIdx.x * 9 * UBD_NF;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float4 a = acc[t];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, s); a.y += __shfl_xor_sync(0xffffffffu, a.y, s);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, s); a.w += __shfl_xor_sync(0xffffffffu, a.w, s);
    }
    if (lane == 0) { float* d = o + t * UBD_NF + 4 * pl; d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; }
  }
}

// g_x[q][c] = (x[q][c] > 0) * sum_taps g_d[(q + pad - tap) / s][c] * dw[tap][c]   (transposed depthwise)
// Block = 32 pixels x 6 planes, DW_ROWS rows per block: the thread's 9 x 4 weights are loaded once, the stride is a
// template parameter (the run-time % and / per tap dominated the first version: 225 us per half-resolution map).
constexpr int DW_ROWS = 8;
template <int STRIDE>
__global__ void __launch_bounds__(192)
dw_bwd_data_kernel(const float4* __restrict__ gd, const float4* __restrict__ xin, float4* __restrict__ gx,
                   const float* __restrict__ dwk, int N, int Hi, int Wi, int ipad, int Ho, int Wo, int opad,
                   int pad_t, int pad_l) {
  const int pl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x = blockIdx.x * 32 + lane, n = blockIdx.z;
  if (x >= Wi) return;
  float4 wk[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wk[t] = __ldg(reinterpret_cast<const float4*>(dwk + t * UBD_NF + 4 * pl));
  const int y0 = blockIdx.y * DW_ROWS;
  for (int y = y0; y < min(y0 + DW_ROWS, Hi); ++y) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int yy = y + pad_t - i;
      if (yy < 0 || (yy % STRIDE)) continue;
      const int oy = yy / STRIDE;
      if (oy >= Ho) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int xx = x + pad_l - j;
        if (xx < 0 || (xx % STRIDE)) continue;
        const int ox = xx / STRIDE;
        if (ox >= Wo) continue;
        const float4 g = __ldg(&gd[act_index(n, pl, oy, ox, Ho, Wo, opad)]);
        const float4 w4 = wk[i * 3 + j];
        s.x = fmaf(g.x, w4.x, s.x); s.y = fmaf(g.y, w4.y, s.y); s.z = fmaf(g.z, w4.z, s.z); s.w = fmaf(g.w, w4.w, s.w);
      }
    }
    const size_t idx = act_index(n, pl, y, x, Hi, Wi, ipad);
    const float4 a = __ldg(&xin[idx]);
    gx[idx] = make_float4(a.x > 0.f ? s.x : 0.f, a.y > 0.f ? s.y : 0.f, a.z > 0.f ? s.z : 0.f, a.w > 0.f ? s.w : 0.f);
  }
}

// First layer (raw image input, CIN = 1 or 3): all three weight gradients in one pass.
//   d[c] = depthwise(image)[c];  g_d[c] = sum_o g_y[o] pw[c][o]
//   dPW[c][o] += d[c] g_y[o];  db[o] += g_y[o];  dDW[tap][c] += image[tap][c] g_d[c]
template <int CIN, typename TIn>
__global__ void __launch_bounds__(128)
l1_bwd_kernel(const TIn* __restrict__ img, const float4* __restrict__ gy, const float* __restrict__ dwk,
              const float* __restrict__ pwk, const float* __restrict__ lut, float pre_scale, float pre_shift,
              int N, int H, int W, int Ho, int Wo, int pad_t, int pad_l,
              float* __restrict__ partials /*[grid][9*CIN + CIN*24 + 24]*/) {
  constexpr int NOUT = 9 * CIN + CIN * UBD_NF + UBD_NF;
  __shared__ float s_dw[9 * CIN], s_pw[CIN * UBD_NF], s_lut[256];
  __shared__ float s_red[4][NOUT];
  for (int i = threadIdx.x; i < 9 * CIN; i += blockDim.x) s_dw[i] = dwk[i];
  for (int i = threadIdx.x; i < CIN * UBD_NF; i += blockDim.x) s_pw[i] = pwk[i];
  if (lut) for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  float acc[NOUT];
#pragma unroll
  for (int i = 0; i < NOUT; ++i) acc[i] = 0.f;
  const size_t P = (size_t)N * Ho * Wo;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(p % Wo), y = (int)((p / Wo) % Ho), n = (int)(p / ((size_t)Wo * Ho));
    float g[UBD_NF];
#pragma unroll
    for (int pl = 0; pl < UBD_NG; ++pl) {
      const float4 v = __ldg(&gy[act_index(n, pl, y, x, Ho, Wo, 0)]);
      g[4 * pl] = v.x; g[4 * pl + 1] = v.y; g[4 * pl + 2] = v.z; g[4 * pl + 3] = v.w;
    }
    float xin[9][CIN], d[CIN], gdc[CIN];
#pragma unroll
    for (int c = 0; c < CIN; ++c) { d[c] = 0.f; gdc[c] = 0.f; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int iy = y * 2 + i - pad_t, ix = x * 2 + j - pad_l;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          float v = 0.f;
          if (ok) {
            const TIn raw = img[(((size_t)n * H + iy) * W + ix) * CIN + c];
            if constexpr (sizeof(TIn) == 1) v = lut ? s_lut[(int)raw] : (float)raw;
            else { v = (float)raw; if (pre_scale != 0.f) v = (v - pre_shift) / pre_scale; }
          }
          xin[i * 3 + j][c] = v;
          d[c] = fmaf(v, s_dw[(i * 3 + j) * CIN + c], d[c]);
        }
      }
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int o = 0; o < UBD_NF; ++o) {
        gdc[c] = fmaf(g[o], s_pw[c * UBD_NF + o], gdc[c]);
        acc[9 * CIN + c * UBD_NF + o] = fmaf(d[c], g[o], acc[9 * CIN + c * UBD_NF + o]);
      }
#pragma unroll
    for (int o = 0; o < UBD_NF; ++o) acc[9 * CIN + CIN * UBD_NF + o] += g[o];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < CIN; ++c) acc[t * CIN + c] = fmaf(xin[t][c], gdc[c], acc[t * CIN + c]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    float v = acc[i];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) s_red[wid][i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NOUT; i += blockDim.x)
    partials[(size_t)blockIdx.x * NOUT + i] = s_red[0][i] + s_red[1][i] + s_red[2][i] + s_red[3][i];
}

// ------------------------------------------------------------------------------------------------
// Keras-2 Adam over the flat parameter buffer (train.py:110)
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_t, float b1, float b2, float eps, float gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace tr
