// Fused separable stem for the tensor-core path (net.py:292-296), sm_100a.
//
// A separable layer is depthwise 3x3 (216 MAC/px, no reuse across channels -> FP32 pipes) followed by a
// pointwise 24->24 (576 MAC/px, a dense contraction -> tcgen05).  Both kernels below compute the
// depthwise output on the FP32 pipes and write it, rounded to tf32, STRAIGHT INTO a shared-memory UMMA
// A tile ([plane][128 px][16 B], K-major SWIZZLE_NONE, same convention as ubd_tc.cuh); three
// kind::tf32 MMAs per 128-pixel segment then do the pointwise conv into TMEM, and the epilogue adds
// bias + ReLU.  The pointwise result never exists as scalar code and the depthwise result never
// leaves the SM.
//
//   stem12_tc_kernel : image -> L1 (separable s2, Cin = 1|3, all FP32, exact) -> L2 depthwise -> L2
//                      pointwise (tcgen05) -> act2 (half resolution).  The 24-channel L1 map (25 MB per
//                      1024x1024 image) lives only in shared memory.
//   stem3_tc_kernel  : act2 -> L3 depthwise s2 -> pointwise (tcgen05) -> act3 (quarter resolution,
//                      x-padded layout, rounded to the tf32 grid for the dilated tensor-core layers).
//
// Tile = 4 rows x 128 px of the layer's output = 4 MMA segments = 4 TMEM accumulators of 32 columns.
// 256 threads, persistent CTAs; phases of one tile are separated by __syncthreads (generic-proxy
// writes of the A tile are published to the async proxy with fence.proxy.async).
#pragma once
#include "ubd_tc.cuh"

namespace stem {

constexpr int SEGPX = 128;
constexpr int THREADS = 256;
constexpr int A_PLANE = SEGPX * 16;           // 2048 B
constexpr int A_SEG = UBD_NG * A_PLANE;       // 12288 B
constexpr int PW_IMG_BYTES = 3 * tc::B_TILE_BYTES;      // 3 K-pairs x 1 KB
constexpr int PW_WB_BYTES = PW_IMG_BYTES + 128;         // + bias[32]

// pointwise weights (1,1,24,24) [c][o] -> UMMA B image, tf32-rounded; then bias
__global__ void build_pw_img_kernel(const float* __restrict__ params, int64_t pw_off, int64_t b_off, uint8_t* __restrict__ dst_, int f16) {
  float* dst = reinterpret_cast<float*>(dst_);
  for (int i = threadIdx.x; i < PW_IMG_BYTES / 4; i += blockDim.x) {
    const int kp = i / 256, rem = i % 256;
    const int kcore = rem / 128, ngroup = (rem % 128) / 32, row = (rem % 32) / 4, col = rem % 4;
    const int ic = kp * 8 + kcore * 4 + col, oc = ngroup * 8 + row;
    dst[i] = oc < UBD_NF ? tc::round_tf32(params[pw_off + ic * UBD_NF + oc]) : 0.f;
  }
  for (int i = threadIdx.x; i < 32; i += blockDim.x)
    dst[PW_IMG_BYTES / 4 + i] = i < UBD_NF ? params[b_off + i] : (i == tc::F16_FLAG_SLOT && f16 ? 1.f : 0.f);     // + 16-bit container flag
}

template <int ROWS>
struct PwSmem {                                // common head of both kernels' shared memory
  uint8_t A[ROWS * A_SEG];                     // depthwise output, UMMA A tiles (12 KB per segment)
  uint8_t wimg[PW_IMG_BYTES];
  float bias[32];
  uint64_t mma_bar;
  uint32_t tmem_base;
};

// Pointwise conv of the ROWS staged segments: 3 MMAs each, one commit.  Call from warp 0 only.
template <int ROWS>
__device__ __forceinline__ void pw_issue(PwSmem<ROWS>& S, uint32_t tmem_base) {
  if (tc::elect_one()) {
    const uint32_t a0 = ((tc::smem_u32(S.A) >> 4) & 0x3FFFu) | (((uint32_t)A_PLANE >> 4) << 16);
    const uint32_t b0 = ((tc::smem_u32(S.wimg) >> 4) & 0x3FFFu) | ((512u >> 4) << 16);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int kp = 0; kp < 3; ++kp)
        tc::umma_tf32(tmem_base + r * tc::UMMA_N, tc::make_desc(a0 + ((r * A_SEG + kp * 2 * A_PLANE) >> 4), tc::DESC_HI),
                      tc::make_desc(b0 + ((kp * tc::B_TILE_BYTES) >> 4), tc::DESC_HI), kp != 0);
    tc::umma_commit(tc::smem_u32(&S.mma_bar));
  }
  __syncwarp();
}

// Epilogue of one tile: warp w reads TMEM quadrant w&3 of segments (w>>2), (w>>2)+2, ...
// MODE 0: fp32 as is; 1: fp32 rounded to the tf32 grid; 2: bf16, 3 planes of 8 channels.
template <int MODE, int ROWS>
__device__ __forceinline__ void pw_epilogue(PwSmem<ROWS>& S, uint32_t tmem_base, float4* __restrict__ out, int n, int y0, int x0,
                                            int Ho, int Wo, int opad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, quad = warp & 3;
#pragma unroll
  for (int h = 0; h < ROWS / 2; ++h) {
    const int r = (warp >> 2) + 2 * h;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + r * tc::UMMA_N;
    uint32_t v[24];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23])
                 : "r"(taddr + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int y = y0 + r, x = x0 + quad * 32 + lane;
    const bool f16 = MODE == 2 && S.bias[tc::F16_FLAG_SLOT] != 0.f;
    (void)f16;
    if (y < Ho && x < Wo) {
      float o[UBD_NF];
#pragma unroll
      for (int c = 0; c < UBD_NF; ++c) o[c] = fmaxf(__uint_as_float(v[c]) + S.bias[c], 0.f);
      if constexpr (MODE == 2) {
        uint4* dst = reinterpret_cast<uint4*>(out) + (((size_t)n * Ho + y) * tc::NG_BF16) * (size_t)(Wo + 2 * opad) + opad + x;
#pragma unroll
        for (int g = 0; g < tc::NG_BF16; ++g)
          dst[(size_t)g * (Wo + 2 * opad)] =
              make_uint4(tc::pack16(o[8 * g], o[8 * g + 1], f16), tc::pack16(o[8 * g + 2], o[8 * g + 3], f16),
                         tc::pack16(o[8 * g + 4], o[8 * g + 5], f16), tc::pack16(o[8 * g + 6], o[8 * g + 7], f16));
      } else {
#pragma unroll
        for (int g = 0; g < UBD_NG; ++g) {
          float4 q = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
          if (MODE == 1) { q.x = tc::round_tf32(q.x); q.y = tc::round_tf32(q.y); q.z = tc::round_tf32(q.z); q.w = tc::round_tf32(q.w); }
          out[act_index(n, g, y, x, Ho, Wo, opad)] = q;
        }
      }
    }
  }
}

template <int ROWS>
__device__ __forceinline__ void pw_setup(PwSmem<ROWS>& S, const uint8_t* __restrict__ wb, int warp) {
  constexpr int TMEM_COLS = ROWS * tc::UMMA_N;
  for (int i = threadIdx.x; i < PW_WB_BYTES / 4; i += blockDim.x)
    reinterpret_cast<float*>(S.wimg)[i] = __ldg(reinterpret_cast<const float*>(wb) + i);     // wimg then bias
  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&S.mma_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&S.tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // wimg was written through the generic proxy
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
}

// Waits for the tile's MMAs.  Plain (unbounded-by-design but HW-suspending) parity wait; the producer of
// this barrier is the tcgen05.commit issued a few instructions earlier by this same CTA.
template <int ROWS>
__device__ __forceinline__ void wait_mma(PwSmem<ROWS>& S, uint32_t parity, int* gerr) {
  const long long t0 = clock64();
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(tc::smem_u32(&S.mma_bar)), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 1500000000LL) { atomicCAS(gerr, 0, 7); break; }
  }
  tc::tc_fence_after();
}

// ------------------------------------------------------------------------------------------------
// image -> L1 -> L2 (act2), streaming.  CIN = 1 or 3; TIn = uint8_t (lut or plain cast) or float.
//
// A CTA takes work items (image, 128-px column strip, RS act2 rows) from an atomic counter and walks
// the strip top to bottom two act2 rows at a time.  Every image row is loaded and preprocessed once
// (even / odd column arrays so the stride-2 taps are conflict-free), every L1 row is computed once
// into a 4-row ring in shared memory, the L2 depthwise of a row pair is written as two UMMA A
// segments, 6 MMAs do the pointwise conv, the epilogue of the pair overlaps the next pair's L1 work.
// Two __syncthreads per row pair; two CTAs per SM hide each other's barrier and MMA latency.
// ------------------------------------------------------------------------------------------------
constexpr int R12 = 2;                         // act2 rows per MMA batch (two segments)
constexpr int RS = 16;                         // act2 rows per work item
constexpr int A1_COLS = SEGPX + 2;             // 130 L1 columns feed 128 L2 columns
constexpr int A1_PITCH = 132;                  // float4 per (row, plane)
constexpr int A1_RING = 4;                     // L1 rows resident (power of two)
constexpr int IMG_COLS = 2 * A1_COLS + 1;      // 261 image columns feed 130 L1 columns
constexpr int IMG_HALF = 132;                  // floats per (row, parity)
constexpr int IMG_RING = 16;                   // image rows resident (power of two)

template <int CIN, typename TIn>
struct Smem12 {
  PwSmem<R12> pw;
  float4 act1[A1_RING * UBD_NG * A1_PITCH];    // 50688 B
  float imgE[CIN][IMG_RING * IMG_HALF];        // preprocessed image rows, even patch columns
  float imgO[CIN][IMG_RING * IMG_HALF];        // odd patch columns
  __align__(16) float pw1[CIN * UBD_NF];
  __align__(16) float b1[UBD_NF];
  __align__(16) float dw2[9 * UBD_NF];
  float dw1[9 * CIN], lut[256];
  int item;
};

template <int CIN, typename TIn>
__global__ void __launch_bounds__(THREADS, 2)
stem12_tc_kernel(const TIn* __restrict__ img, float4* __restrict__ act2, const float* __restrict__ params,
                 int64_t off_dw1, int64_t off_pw1, int64_t off_b1, int64_t off_dw2, const uint8_t* __restrict__ wb2,
                 const float* __restrict__ lut, float pre_scale, float pre_shift,
                 int N, int H, int W, int pad_t, int pad_l, int* __restrict__ work_counter, int* gerr) {
  using SM = Smem12<CIN, TIn>;
  constexpr int ELT = (int)sizeof(TIn);
  constexpr int EPQ = 4 / ELT;                                    // elements per 4-byte load (4 for u8, 1 for float)
  constexpr int QUADS = (IMG_COLS * CIN * ELT + 3) / 4 + 1;       // 4-byte words covering one patch row
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // keep the shared address space (no integer casts)
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int H2 = H / 2, W2 = W / 2;
  for (int i = tid; i < 9 * CIN; i += THREADS) S.dw1[i] = params[off_dw1 + i];
  for (int i = tid; i < CIN * UBD_NF; i += THREADS) S.pw1[i] = params[off_pw1 + i];
  for (int i = tid; i < UBD_NF; i += THREADS) S.b1[i] = params[off_b1 + i];
  for (int i = tid; i < 9 * UBD_NF; i += THREADS) S.dw2[i] = params[off_dw2 + i];
  if (lut) for (int i = tid; i < 256; i += THREADS) S.lut[i] = lut[i];
  pw_setup(S.pw, wb2, warp);
  const uint32_t tmem_base = S.pw.tmem_base;
  const int row_bytes_img = W * CIN * ELT;                        // multiple of 16
  float dw1r[9 * CIN];
#pragma unroll
  for (int i = 0; i < 9 * CIN; ++i) dw1r[i] = S.dw1[i];

  const int xt = (W2 + SEGPX - 1) / SEGPX, yt = (H2 + RS - 1) / RS;
  const int nitems = N * yt * xt;
  uint32_t mma_count = 0;                                        // MMA batches issued by this CTA (barrier parity)

  // image row `iy` (may be outside the image) -> ring slot; words [q0, q0 + nq) of the patch row
  auto load_word = [&](const uint8_t* img_n, int iy, int a0, int q) -> uint32_t {
    const int off = a0 + 4 * q;
    if (iy < 0 || iy >= H || off < 0 || off >= row_bytes_img) return 0u;
    return __ldg(reinterpret_cast<const uint32_t*>(img_n + (size_t)iy * row_bytes_img + off));
  };
  auto store_word = [&](uint32_t v, int iy, int a0, int e0, int q) {
    const int off = a0 + 4 * q;
    const bool inside = iy >= 0 && iy < H && off >= 0 && off < row_bytes_img;
    const int slot = iy & (IMG_RING - 1);
    const TIn* e = reinterpret_cast<const TIn*>(&v);
#pragma unroll
    for (int k = 0; k < EPQ; ++k) {
      const int idx = e0 + q * EPQ + k;
      if ((unsigned)idx >= (unsigned)(IMG_COLS * CIN)) continue;
      const int col = CIN == 1 ? idx : idx / CIN, ch = CIN == 1 ? 0 : idx - col * CIN;
      float f = 0.f;
      if (inside) {
        if constexpr (sizeof(TIn) == 1) f = lut ? S.lut[(int)e[k]] : (float)e[k];
        else { f = (float)e[k]; if (pre_scale != 0.f) f = (f - pre_shift) / pre_scale; }
      }
      ((col & 1) ? S.imgO[ch] : S.imgE[ch])[slot * IMG_HALF + (col >> 1)] = f;
    }
  };
  // one L1 pixel (row yy, tile column c) -> ring.  Outside the L1 map: zero (L2's 'same' padding).
  auto l1_pixel = [&](int yy, int c, int x0) {
    const int xx = x0 - 1 + c;
    float4 o[UBD_NG];
    if (yy >= 0 && yy < H2 && xx >= 0 && xx < W2) {
      float d[CIN];
#pragma unroll
      for (int ch = 0; ch < CIN; ++ch) {
        float a = 0.f;
#pragma unroll
        for (int ti = 0; ti < 3; ++ti) {
          const int slot = (2 * yy - pad_t + ti) & (IMG_RING - 1);
          const float* E = &S.imgE[ch][slot * IMG_HALF + c];
          const float* O = &S.imgO[ch][slot * IMG_HALF + c];
          a = fmaf(E[0], dw1r[(ti * 3 + 0) * CIN + ch], a);
          a = fmaf(O[0], dw1r[(ti * 3 + 1) * CIN + ch], a);
          a = fmaf(E[1], dw1r[(ti * 3 + 2) * CIN + ch], a);
        }
        d[ch] = a;
      }
      const float4* b4 = reinterpret_cast<const float4*>(S.b1);
      const float4* p4 = reinterpret_cast<const float4*>(S.pw1);
#pragma unroll
      for (int g = 0; g < UBD_NG; ++g) {
        float4 a = b4[g];
#pragma unroll
        for (int ch = 0; ch < CIN; ++ch) {
          const float4 w = p4[ch * UBD_NG + g];
          a.x = fmaf(d[ch], w.x, a.x); a.y = fmaf(d[ch], w.y, a.y); a.z = fmaf(d[ch], w.z, a.z); a.w = fmaf(d[ch], w.w, a.w);
        }
        o[g] = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
      }
    } else {
#pragma unroll
      for (int g = 0; g < UBD_NG; ++g) o[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int rs = yy & (A1_RING - 1);
#pragma unroll
    for (int g = 0; g < UBD_NG; ++g) S.act1[(rs * UBD_NG + g) * A1_PITCH + c] = o[g];
  };

  while (true) {
    if (tid == 0) S.item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int item = S.item;
    if (item >= nitems) break;
    const int x0 = (item % xt) * SEGPX, y0 = ((item / xt) % yt) * RS, n = item / (xt * yt);
    const int npairs = min(RS, H2 - y0) / 2;
    const uint8_t* img_n = reinterpret_cast<const uint8_t*>(img) + (size_t)n * H * row_bytes_img;
    const int ix0 = 2 * (x0 - 1) - pad_l;
    const int b0 = ix0 * CIN * ELT;
    const int a0 = b0 >= 0 ? (b0 & ~3) : -(((-b0) + 3) & ~3);       // floored to 4 bytes
    const int e0 = (a0 - b0) / ELT;                                 // patch element index of word 0 (<= 0)
    // ---- prologue: image rows for L1 rows y0-1 .. y0+4, then L1 rows y0-1 .. y0+2
    const int iyA = 2 * (y0 - 1) - pad_t;                           // first image row needed
    for (int i = tid; i < 13 * QUADS; i += THREADS) {
      const int r = i / QUADS, q = i - r * QUADS;
      store_word(load_word(img_n, iyA + r, a0, q), iyA + r, a0, e0, q);
    }
    __syncthreads();
    for (int i = tid; i < 4 * A1_COLS; i += THREADS) l1_pixel(y0 - 1 + i / A1_COLS, i % A1_COLS, x0);
    __syncthreads();

    for (int p = 0; p < npairs; ++p) {
      const int y = y0 + 2 * p;                                     // act2 rows y, y+1 this iteration
      // (a) issue the loads of the 4 image rows that L1 rows y+5, y+6 add (used next iteration)
      const int iyN = 2 * (y + 5) - pad_t + 1;
      constexpr bool PREFETCH = 4 * QUADS <= 2 * THREADS;          // grey uint8: two words per thread in flight
      uint32_t w0 = 0u, w1 = 0u;
      const int i0 = tid, i1 = tid + THREADS;
      if (PREFETCH && p + 1 < npairs) {
        w0 = load_word(img_n, iyN + i0 / QUADS, a0, i0 % QUADS);
        if (i1 < 4 * QUADS) w1 = load_word(img_n, iyN + i1 / QUADS, a0, i1 % QUADS);
      }
      // (c) previous pair: its MMAs are done by now -> epilogue; also frees the A tiles
      if (p > 0) {
        wait_mma(S.pw, (mma_count - 1) & 1, gerr);
        pw_epilogue<0>(S.pw, tmem_base, act2, n, y - 2, x0, H2, W2, 0);
        tc::tc_fence_before();
      }
      // (d) L2 depthwise of rows y, y+1 -> A segments 0, 1.  Task = (plane, pixel column).
      {
        const int px = tid & (SEGPX - 1), gb = (tid >> 7) * 3;
        int rs[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) rs[rr] = (y - 1 + rr) & (A1_RING - 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int g = gb + k;
          const float4* w4 = reinterpret_cast<const float4*>(S.dw2) + g;       // [tap][6 planes]
          float4 wt[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) wt[t] = w4[t * UBD_NG];
          float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const float4* row = &S.act1[(rs[rr] * UBD_NG + g) * A1_PITCH + px];
            const float4 in0 = row[0], in1 = row[1], in2 = row[2];
            if (rr < 3) {
              const float4 wa = wt[rr * 3], wb = wt[rr * 3 + 1], wc = wt[rr * 3 + 2];
              acc0.x = fmaf(in0.x, wa.x, fmaf(in1.x, wb.x, fmaf(in2.x, wc.x, acc0.x)));
              acc0.y = fmaf(in0.y, wa.y, fmaf(in1.y, wb.y, fmaf(in2.y, wc.y, acc0.y)));
              acc0.z = fmaf(in0.z, wa.z, fmaf(in1.z, wb.z, fmaf(in2.z, wc.z, acc0.z)));
              acc0.w = fmaf(in0.w, wa.w, fmaf(in1.w, wb.w, fmaf(in2.w, wc.w, acc0.w)));
            }
            if (rr >= 1) {
              const float4 wa = wt[(rr - 1) * 3], wb = wt[(rr - 1) * 3 + 1], wc = wt[(rr - 1) * 3 + 2];
              acc1.x = fmaf(in0.x, wa.x, fmaf(in1.x, wb.x, fmaf(in2.x, wc.x, acc1.x)));
              acc1.y = fmaf(in0.y, wa.y, fmaf(in1.y, wb.y, fmaf(in2.y, wc.y, acc1.y)));
              acc1.z = fmaf(in0.z, wa.z, fmaf(in1.z, wb.z, fmaf(in2.z, wc.z, acc1.z)));
              acc1.w = fmaf(in0.w, wa.w, fmaf(in1.w, wb.w, fmaf(in2.w, wc.w, acc1.w)));
            }
          }
          reinterpret_cast<float4*>(S.pw.A + 0 * A_SEG + g * A_PLANE)[px] =
              make_float4(tc::round_tf32(acc0.x), tc::round_tf32(acc0.y), tc::round_tf32(acc0.z), tc::round_tf32(acc0.w));
          reinterpret_cast<float4*>(S.pw.A + 1 * A_SEG + g * A_PLANE)[px] =
              make_float4(tc::round_tf32(acc1.x), tc::round_tf32(acc1.y), tc::round_tf32(acc1.z), tc::round_tf32(acc1.w));
        }
      }
      // (e) the image words loaded in (a) have arrived: preprocess into the ring
      if (p + 1 < npairs) {
        if (PREFETCH) {
          store_word(w0, iyN + i0 / QUADS, a0, e0, i0 % QUADS);
          if (i1 < 4 * QUADS) store_word(w1, iyN + i1 / QUADS, a0, e0, i1 % QUADS);
        } else {
          for (int i = tid; i < 4 * QUADS; i += THREADS)
            store_word(load_word(img_n, iyN + i / QUADS, a0, i % QUADS), iyN + i / QUADS, a0, e0, i % QUADS);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      tc::tc_fence_after();
      if (warp == 0) pw_issue(S.pw, tmem_base);
      ++mma_count;
      // (b) L1 rows y+3, y+4 for the next iteration, into the ring slots of rows y-1, y (free since the
      //     barrier above); runs while the tensor core does this pair's pointwise conv
      if (p + 1 < npairs) {
        l1_pixel(y + 3 + tid / A1_COLS, tid % A1_COLS, x0);
        if (tid < 2 * A1_COLS - THREADS) l1_pixel(y + 3 + (tid + THREADS) / A1_COLS, (tid + THREADS) % A1_COLS, x0);
        __syncthreads();
      }
    }
    // ---- last pair of the item
    wait_mma(S.pw, (mma_count - 1) & 1, gerr);
    pw_epilogue<0>(S.pw, tmem_base, act2, n, y0 + 2 * (npairs - 1), x0, H2, W2, 0);
    tc::tc_fence_before();
    __syncthreads();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(R12 * tc::UMMA_N) : "memory");
}

// ------------------------------------------------------------------------------------------------
// act2 -> L3 (act3, quarter resolution, padded layout, tf32 grid)
// ------------------------------------------------------------------------------------------------
constexpr int R3 = 4;                          // act3 rows per tile
struct Smem3 {
  PwSmem<R3> pw;
  float dw3[9 * UBD_NF];
};

template <int OUT_MODE>
__global__ void __launch_bounds__(THREADS, 3)
stem3_tc_kernel(const float4* __restrict__ act2, float4* __restrict__ act3, const float* __restrict__ params,
                int64_t off_dw3, const uint8_t* __restrict__ wb3, int N, int H2, int W2, int pad_t, int pad_l, int* gerr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // keep the shared address space (no integer casts)
  Smem3& S = *reinterpret_cast<Smem3*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int H4 = H2 / 2, W4 = W2 / 2;
  for (int i = threadIdx.x; i < 9 * UBD_NF; i += THREADS) S.dw3[i] = params[off_dw3 + i];
  pw_setup(S.pw, wb3, warp);
  const uint32_t tmem_base = S.pw.tmem_base;
  const int xt = (W4 + SEGPX - 1) / SEGPX, yt = (H4 + R3 - 1) / R3;
  const int ntiles = N * yt * xt;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int x0 = (tile % xt) * SEGPX, y0 = ((tile / xt) % yt) * R3, n = tile / (xt * yt);
    // depthwise stride 2 straight from the act2 map (L1/L2-cached) into the A tiles
    for (int task = threadIdx.x; task < R3 * UBD_NG * SEGPX; task += THREADS) {
      const int p = task % SEGPX, g = (task / SEGPX) % UBD_NG, r = task / (SEGPX * UBD_NG);
      const int y = y0 + r, x = x0 + p;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y < H4 && x < W4) {
#pragma unroll
        for (int ti = 0; ti < 3; ++ti) {
          const int iy = 2 * y + ti - pad_t;
          if (iy < 0 || iy >= H2) continue;
#pragma unroll
          for (int tj = 0; tj < 3; ++tj) {
            const int ix = 2 * x + tj - pad_l;
            if (ix < 0 || ix >= W2) continue;
            const float4 v = __ldg(&act2[act_index(n, g, iy, ix, H2, W2, 0)]);
            const float* wk = &S.dw3[(ti * 3 + tj) * UBD_NF + 4 * g];
            a.x = fmaf(v.x, wk[0], a.x); a.y = fmaf(v.y, wk[1], a.y); a.z = fmaf(v.z, wk[2], a.z); a.w = fmaf(v.w, wk[3], a.w);
          }
        }
      }
      reinterpret_cast<float4*>(S.pw.A + r * A_SEG + g * A_PLANE)[p] =
          make_float4(tc::round_tf32(a.x), tc::round_tf32(a.y), tc::round_tf32(a.z), tc::round_tf32(a.w));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) pw_issue(S.pw, tmem_base);
    wait_mma(S.pw, it & 1, gerr);
    pw_epilogue<OUT_MODE>(S.pw, tmem_base, act3, n, y0, x0, H4, W4, UBD_MAP_PAD);
    tc::tc_fence_before();
    __syncthreads();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(R3 * tc::UMMA_N) : "memory");
}

}  // namespace stem

// Function attributes are per device: called from every ubd_create (a second handle on another GPU of the
// same process must not inherit a "done" flag).
template <int CIN, typename TIn>
static void stem12_set_attr() {
  const int smem = (int)sizeof(stem::Smem12<CIN, TIn>) + 128;
  cudaFuncSetAttribute(stem::stem12_tc_kernel<CIN, TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(stem::stem12_tc_kernel<CIN, TIn>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   // 2 CTAs / SM
}
static void stem_setup_attributes() {
  stem12_set_attr<1, uint8_t>(); stem12_set_attr<1, float>(); stem12_set_attr<3, uint8_t>(); stem12_set_attr<3, float>();
  const int smem3 = (int)sizeof(stem::Smem3) + 128;
  cudaFuncSetAttribute(stem::stem3_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
  cudaFuncSetAttribute(stem::stem3_tc_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(stem::stem3_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
  cudaFuncSetAttribute(stem::stem3_tc_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <int CIN, typename TIn>
static int stem12_launch(ubd_handle h, const TIn* img, float4* act2, const float* lut, float ps, float psh,
                         int n, int H, int W, int p2) {
  auto kern = stem::stem12_tc_kernel<CIN, TIn>;
  const size_t smem = sizeof(stem::Smem12<CIN, TIn>) + 128;
  const int nitems = n * ((H / 2 + stem::RS - 1) / stem::RS) * ((W / 2 + stem::SEGPX - 1) / stem::SEGPX);
  const int grid = std::min(nitems, 2 * h->n_sm);
  const uint8_t* wb2 = (const uint8_t*)h->stem_wimg.p;
  int* counter = reinterpret_cast<int*>((uint8_t*)h->stem_wimg.p + 2 * stem::PW_WB_BYTES);
  UBD_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), h->stream));
  kern<<<grid, stem::THREADS, smem, h->stream>>>(img, act2, h->d_params, h->spec.off[0], h->spec.off[1], h->spec.off[2],
                                                 h->spec.off[3], wb2, lut, ps, psh, n, H, W, p2, p2, counter, tc_err_flag(h));
  ++h->launches;
  UBD_CUDA(cudaGetLastError());
  return UBD_OK;
}

static int stemf_launch(ubd_handle h, const void* d_img, int in_dtype, int preproc, int n, int H, int W, float4* act3);

// Does the fused separable kernel (ubd_stemf.cuh) take this input?  Grey images only (one L1 scalar per pixel).
// opt_stem_variant: 2 = always; 1 = never; 0 (auto) = for float input - uint8 input takes the dense tcgen05 pair
// (L1+L2 dense, L3 dense), measured 1.035 ms vs 1.07 ms per 64 x 1024^2 on B200 although it round-trips the
// half-resolution map through HBM (profiles/r02_stem_summary.md).
static inline bool stem_is_fused(ubd_handle h, int in_dtype) {
  if (h->spec.cin != 1 || h->opt_stem_variant == 1) return false;
  return h->opt_stem_variant == 2 || in_dtype != UBD_U8;
}

// tensor-core stem: image -> act3.  Needs tc_prepare (error flag) and the pointwise B images.
static int run_stem_tc(ubd_handle h, const void* d_img, int in_dtype, int preproc, int n, int H, int W,
                       float4* act2, float4* act3) {
  int rc = tc_prepare(h);
  if (rc) return rc;
  if (!h->stem_wimg.p) {
    UBD_CUDA(cudaMalloc(&h->stem_wimg.p, 2 * stem::PW_WB_BYTES + 64));       // + work counter
    h->stem_wimg.cap = 2 * stem::PW_WB_BYTES + 64;
    h->stem_weights_dirty = true;
  }
  if (ubd_is16(h) && h->stem_w16 != h->precision) h->stem_weights_dirty = true;
  if (h->stem_weights_dirty) {
    for (int l = 1; l <= 2; ++l) {
      stem::build_pw_img_kernel<<<1, 256, 0, h->stream>>>(h->d_params, h->spec.off[3 * l + 1], h->spec.off[3 * l + 2],
                                                          (uint8_t*)h->stem_wimg.p + (l - 1) * stem::PW_WB_BYTES, h->precision == UBD_F16);
      ++h->launches;
    }
    UBD_CUDA(cudaGetLastError());
    h->stem_weights_dirty = false;
    h->stem_w16 = ubd_is16(h) ? h->precision : h->stem_w16;
  }
  if (stem_is_fused(h, in_dtype)) return stemf_launch(h, d_img, in_dtype, preproc, n, H, W, act3);
  const int p2 = stride2_pad(h);
  const bool mob = preproc == UBD_PREPROC_MOBILENET;
  if (h->spec.cin == 1 && in_dtype == UBD_U8 && h->opt_dense_l2) {
    // Whole stem on the tensor cores: L2 and L3 as dense 3x3 convs (merged separable kernels), L1 computed
    // by producer warps of the L2 kernel.  act2 travels split by column parity so that L3's stride-2 taps
    // are unit-stride descriptor offsets.
    const int H2 = H / 2, W2 = W / 2;
    const size_t split_bytes = (size_t)n * H2 * 2 * UBD_NG * (size_t)(W2 / 2 + 2 * UBD_MAP_PAD) * sizeof(float4);
    const long long tag = ((long long)h->precision << 60) ^ ((long long)n << 40) ^ ((long long)H << 20) ^ (long long)W ^ (1LL << 59);
    if (h->act2_tag != tag) {          // zero pads of the split layout: clear when the geometry / layout changes
      UBD_CUDA(cudaMemsetAsync(act2, 0, std::min(split_bytes, h->act2.cap), h->stream));
      h->act2_tag = tag;
    }
    tc::L1Args la{mob ? h->d_lut : nullptr, h->d_params + h->spec.off[0], h->d_params + h->spec.off[1], h->d_params + h->spec.off[2], H, W, p2, p2};
    rc = tc4_launch_dilconv(h, d_img, act2, UBD_NLAYERS_DIL, n, H2, W2, 1, /*out_mode=*/3, UBD_MAP_PAD, nullptr, &la);
    if (rc) return rc;
    return tc_launch_dilconv(h, act2, act3, UBD_NLAYERS_DIL + 1, n, H2 / 2, W2 / 2, 1, /*out_mode=*/0, UBD_MAP_PAD, nullptr,
                             /*s2=*/p2 ? 1 : 2);
  } else if (h->spec.cin == 1) {
    if (in_dtype == UBD_U8) rc = stem12_launch<1, uint8_t>(h, (const uint8_t*)d_img, act2, mob ? h->d_lut : nullptr, 0.f, 0.f, n, H, W, p2);
    else rc = stem12_launch<1, float>(h, (const float*)d_img, act2, nullptr, mob ? 127.5f : 0.f, 127.5f, n, H, W, p2);
  } else {
    if (in_dtype == UBD_U8) rc = stem12_launch<3, uint8_t>(h, (const uint8_t*)d_img, act2, mob ? h->d_lut : nullptr, 0.f, 0.f, n, H, W, p2);
    else rc = stem12_launch<3, float>(h, (const float*)d_img, act2, nullptr, mob ? 127.5f : 0.f, 127.5f, n, H, W, p2);
  }
  if (rc) return rc;
  h->act2_tag = 0;                     // plain-layout act2 from here on
  {
    const size_t smem = sizeof(stem::Smem3) + 128;
    const int H2 = H / 2, W2 = W / 2;
    const int ntiles = n * ((H2 / 2 + stem::R3 - 1) / stem::R3) * ((W2 / 2 + stem::SEGPX - 1) / stem::SEGPX);
    const int grid = std::min(ntiles, 3 * h->n_sm);
    const uint8_t* wb3 = (const uint8_t*)h->stem_wimg.p + stem::PW_WB_BYTES;
    if (ubd_is16(h))
      stem::stem3_tc_kernel<2><<<grid, stem::THREADS, smem, h->stream>>>(act2, act3, h->d_params, h->spec.off[6], wb3, n, H2, W2, p2, p2, tc_err_flag(h));
    else
      stem::stem3_tc_kernel<1><<<grid, stem::THREADS, smem, h->stream>>>(act2, act3, h->d_params, h->spec.off[6], wb3, n, H2, W2, p2, p2, tc_err_flag(h));
    ++h->launches;
    UBD_CUDA(cudaGetLastError());
  }
  return UBD_OK;
}
