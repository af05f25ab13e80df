// Fused separable stem (net.py:292-296): image -> L1 -> L2 -> L3 -> act3 in ONE kernel, sm_100a.
//
// Nothing between the uint8 / float image and the quarter-resolution 24-channel map touches HBM:
//   L1 (separable s2, Cin = 1): its 24-channel output is relu(b1[c] + pw1[c] * s) with ONE scalar s per
//       pixel (the depthwise sum of the preprocessed image patch), so only s lives in shared memory (a
//       ring of 130-float rows) and the L2 depthwise rebuilds the channels it needs in registers;
//   L2 (separable s1): depthwise 3x3 on the FP32 pipes written (tf32 grid) straight into UMMA A tiles,
//       pointwise 24->24 as three kind::tf32 tcgen05 MMAs per 128-px row, accumulators in TMEM; the
//       epilogue (bias + ReLU) writes a 3-row ring of L2 rows in shared memory, split by column parity so
//       that the stride-2 taps of L3 are unit-stride;
//   L3 (separable s2): depthwise from that ring into an A tile, three MMAs, epilogue -> act3 in the padded
//       planar layout the dilated layers stage with cp.async.bulk (tf32 grid, or bf16 planes).
//
// Work item = (image, band of BH act3 rows); a CTA walks the band strip by strip (64 act3 px = 128 L2 px,
// aligned), carrying the one L2 column the next strip's stride-2 taps need (FML padding: column 2x-1;
// TF 'same': column 2x+2, strips then run right to left) in shared memory, so no column is computed
// twice and no halo is exchanged.  One step = two new L2 rows + one act3 row; the L3 part of a step lags
// the L2 part by one step, so both MMA groups of a step are issued together and their latency is covered
// by the next rows' image loads and L1 sums.  Two __syncthreads per step, 2 CTAs per SM.
#pragma once
#include "ubd_stem_tc.cuh"

namespace stemf {

constexpr int THREADS = 256;
constexpr int SW2 = 128;                      // L2 pixels per strip = UMMA M
constexpr int SW3 = 64;                       // act3 pixels per strip
constexpr int BH = 16;                        // act3 rows per work item
constexpr int A_PLANE = SW2 * 16;             // 2048 B: one channel plane of an A tile
constexpr int A_TILE = UBD_NG * A_PLANE;      // 12288 B
constexpr int SC_RING = 8, SC_PITCH = 132;    // L1 scalar rows (130 columns feed 128 L2 columns)
constexpr int IMG_RING = 8, IMG_PITCH = 272;  // preprocessed image rows (floats), element 0 = the word-aligned patch start
constexpr int IMG_COLS = 2 * (SW2 + 2) + 1;   // 261 image columns feed 130 L1 columns
constexpr int R2_PITCH = 134;                 // float4 per (slot, plane): A[0..64] then B[0..63] at offset 68 / 69
constexpr int R2_SLOT = UBD_NG * R2_PITCH;    // float4 per ring row
constexpr int CARRY_ROWS = 2 * BH + 1;
constexpr int TMEM_COLS = 128;                // D2[0], D2[1], D3 (32 columns each)

struct Smem {
  uint8_t A2[2 * A_TILE];                     // L2 depthwise output of the step's two rows
  uint8_t A3[A_TILE];                         // L3 depthwise output (64 of 128 rows used)
  float4 r2[3 * R2_SLOT];                     // L2 rows (post-ReLU), ring of 3, parity split
  float4 carry[2][CARRY_ROWS * UBD_NG];       // L2 column handed to the next strip (double-buffered by strip)
  uint8_t wimg2[stem::PW_IMG_BYTES], wimg3[stem::PW_IMG_BYTES];
  __align__(16) float imgp[IMG_RING * IMG_PITCH];
  __align__(16) float sc[2 * SC_RING * SC_PITCH];   // L1 sums, each stored twice (fp32x2 operand)
  __align__(16) float dw2[9 * UBD_NF];
  __align__(16) float dw3[9 * UBD_NF];
  __align__(16) float b2[32], b3[32];
  __align__(16) float pw1[UBD_NF], b1[UBD_NF];
  float dw1[12], lut[256];
  uint64_t mma_bar;
  uint32_t tmem_base;
  int item;
};

__device__ __forceinline__ float rna_bits(float v) { return __uint_as_float(__float_as_uint(v) + 0x1000u); }

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, int* gerr) {
  uint32_t polls = 0;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    if (++polls > 20000000u) { atomicCAS(gerr, 0, 8); break; }
  }
  tc::tc_fence_after();
}

// packed fp32x2 arithmetic (sm_100: one FFMA2 / FADD2 / FMUL2 issue slot for two lanes of work)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 relu2(float2 a) { return make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)); }
__device__ __forceinline__ float2 rna2(float2 a) { return make_float2(rna_bits(a.x), rna_bits(a.y)); }

// OUT_MODE 1: fp32 planes rounded to the tf32 grid; 2: bf16, 3 planes of 8 channels.
//
// Thread roles inside a step (256 threads):
//   L2 depthwise : thread = (pixel pair pp = tid & 63 -> strip pixels 2pp, 2pp+1; channel group tid >> 6 -> 6 channels
//                  as 3 fp32x2 pairs).  The two pixels share their four L1 columns, so a rebuilt L1 value feeds
//                  both; M row of pixel px in the A tile / TMEM lane: m = (px & 1) * 64 + (px >> 1), which makes
//                  both the A-tile stores here and the parity-split ring stores of the epilogue unit-stride.
//   L3 depthwise : three (pixel j = tid & 63, channel pair) outputs per thread.
//   epilogues    : warp w -> L2 row (w >> 2), TMEM quadrant (w & 3); the act3 row goes to warps 0,1,4,5 (12 channels each).
template <typename TIn, int OUT_MODE>
__global__ void __launch_bounds__(THREADS, 2)
stem_fused_kernel(const TIn* __restrict__ img, float4* __restrict__ act3, const float* __restrict__ params,
                  int64_t off_dw1, int64_t off_pw1, int64_t off_b1, int64_t off_dw2, int64_t off_dw3,
                  const uint8_t* __restrict__ wb2, const uint8_t* __restrict__ wb3,
                  const float* __restrict__ lut, float pre_scale, float pre_shift,
                  int N, int H, int W, int p2, int* __restrict__ work_counter, int* gerr) {
  constexpr int ELT = (int)sizeof(TIn);
  constexpr int EPQ = 4 / ELT;                                   // elements per 4-byte word
  constexpr int QUADS = ELT == 1 ? (IMG_COLS + 3 + 3) / 4 : IMG_COLS;    // words covering one patch row (u8: word-floored start)
  constexpr int NW = (4 * QUADS + THREADS - 1) / THREADS;        // words per thread for the 4 new image rows of a step
  static_assert(QUADS * EPQ <= IMG_PITCH, "image ring pitch");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;

  for (int i = tid; i < 9; i += THREADS) S.dw1[i] = params[off_dw1 + i];
  for (int i = tid; i < UBD_NF; i += THREADS) { S.pw1[i] = params[off_pw1 + i]; S.b1[i] = params[off_b1 + i]; }
  for (int i = tid; i < 9 * UBD_NF; i += THREADS) { S.dw2[i] = params[off_dw2 + i]; S.dw3[i] = params[off_dw3 + i]; }
  for (int i = tid; i < 256; i += THREADS) S.lut[i] = lut ? lut[i] : (float)i;
  for (int i = tid; i < stem::PW_IMG_BYTES / 4; i += THREADS) {
    reinterpret_cast<float*>(S.wimg2)[i] = __ldg(reinterpret_cast<const float*>(wb2) + i);
    reinterpret_cast<float*>(S.wimg3)[i] = __ldg(reinterpret_cast<const float*>(wb3) + i);
  }
  for (int i = tid; i < 32; i += THREADS) {
    S.b2[i] = __ldg(reinterpret_cast<const float*>(wb2 + stem::PW_IMG_BYTES) + i);
    S.b3[i] = __ldg(reinterpret_cast<const float*>(wb3 + stem::PW_IMG_BYTES) + i);
  }
  // rows 64..127 of the L3 A tile are never written: keep them finite
  for (int i = tid; i < A_TILE / 16; i += THREADS) reinterpret_cast<float4*>(S.A3)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    tc::mbar_init(tc::smem_u32(&S.mma_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&S.tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;

  // ---- per-thread constants
  const int pp = tid & 63, qg = tid >> 6;                        // L2 depthwise: pixel pair, channel group (pairs 3qg .. 3qg+2)
  // A2 byte offset of channel pair 3qg (even pixel -> M row pp, odd pixel -> M row 64 + pp); pair c2 lives at
  // plane c2 >> 1, half c2 & 1: consecutive pairs are 8 bytes apart inside a plane, else A_PLANE - 8
  const int row_bytes_img = W * ELT;
  const int boff = p2 ? 69 : 68;                                 // bank-conflict-free parity split for this padding mode
  const int ns = (W4 + SW3 - 1) / SW3;
  const int nbands = (H4 + BH - 1) / BH;
  const int nitems = N * nbands;
  uint32_t mma_count = 0;
  // image words of a step: word slot k of this thread = (row r_k of the 4 new rows, word q_k of the patch row)
  int wr_[NW], wq_[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) { const int idx = tid + k * THREADS; wr_[k] = idx < 4 * QUADS ? idx / QUADS : -1; wq_[k] = idx % QUADS; }
  // L1 sums of a step: (row, column) of the two new rows
  const int sc_r0 = tid / (SW2 + 2), sc_c0 = tid % (SW2 + 2);    // second task (tid < 4): row 1, column 126 + tid

  const uint32_t a2_0 = ((tc::smem_u32(S.A2) >> 4) & 0x3FFFu) | (((uint32_t)A_PLANE >> 4) << 16);
  const uint32_t a3_0 = ((tc::smem_u32(S.A3) >> 4) & 0x3FFFu) | (((uint32_t)A_PLANE >> 4) << 16);
  const uint32_t b2_0 = ((tc::smem_u32(S.wimg2) >> 4) & 0x3FFFu) | ((512u >> 4) << 16);
  const uint32_t b3_0 = ((tc::smem_u32(S.wimg3) >> 4) & 0x3FFFu) | ((512u >> 4) << 16);

  while (true) {
    if (tid == 0) S.item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int item = S.item;
    if (item >= nitems) break;
    const int n = item / nbands, Y0 = (item % nbands) * BH;
    const int nrows3 = min(BH, H4 - Y0);
    const int r0 = 2 * Y0 - p2;                                  // first L2 row the band needs
    const uint8_t* img_n = reinterpret_cast<const uint8_t*>(img) + (size_t)n * H * row_bytes_img;

    for (int si = 0; si < ns; ++si) {
      const int s = p2 ? si : ns - 1 - si;
      const int X2 = s * SW2;                                    // first L2 column of the strip
      const int w3s = min(SW3, W4 - s * SW3);
      const float4* carry_in = S.carry[si & 1];
      float4* carry_out = S.carry[(si & 1) ^ 1];
      const int ix0 = 2 * (X2 - 1) - p2;                         // image column of patch column 0
      const int b0 = ix0 * ELT;
      const int a0 = b0 >= 0 ? (b0 & ~3) : -(((-b0) + 3) & ~3);  // floored to a word
      const int e0 = (a0 - b0) / ELT;                            // patch element index of word 0 (<= 0)
      const int eoff = -e0;                                      // ring element of patch column 0
      // per-strip constants of this thread's image words: is the word inside the row, its address in image row 0
      bool wok[NW];
      const uint8_t* wp[NW];
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        const int off = a0 + 4 * wq_[k];
        wok[k] = wr_[k] >= 0 && off >= 0 && off < row_bytes_img;
        wp[k] = img_n + off;
      }
      auto load_word = [&](int iy, int q) -> uint32_t {
        const int off = a0 + 4 * q;
        if (iy < 0 || iy >= H || off < 0 || off >= row_bytes_img) return 0u;
        return __ldg(reinterpret_cast<const uint32_t*>(img_n + (size_t)iy * row_bytes_img + off));
      };
      // word q of image row iy -> ring (zeros outside the image): uint8: four table look-ups, one 16-byte store
      auto store_word = [&](uint32_t v, int iy, bool inside, int q) {
        float* dst = &S.imgp[(iy & (IMG_RING - 1)) * IMG_PITCH + q * EPQ];
        if constexpr (sizeof(TIn) == 1) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (inside) f = make_float4(S.lut[v & 0xFFu], S.lut[(v >> 8) & 0xFFu], S.lut[(v >> 16) & 0xFFu], S.lut[v >> 24]);
          *reinterpret_cast<float4*>(dst) = f;
        } else {
          float f = 0.f;
          if (inside) { f = __uint_as_float(v); if (pre_scale != 0.f) f = (f - pre_shift) / pre_scale; }
          *dst = f;
        }
      };
      // L1 depthwise sum of map pixel (yy, local column c), stored twice (fp32x2 operand); rows outside the map are never read
      auto l1_scalar = [&](int yy, int c) {
        float a = 0.f;
#pragma unroll
        for (int ti = 0; ti < 3; ++ti) {
          const float* r = &S.imgp[((2 * yy - p2 + ti) & (IMG_RING - 1)) * IMG_PITCH + eoff + 2 * c];
          a = fmaf(r[0], S.dw1[ti * 3 + 0], a);
          a = fmaf(r[1], S.dw1[ti * 3 + 1], a);
          a = fmaf(r[2], S.dw1[ti * 3 + 2], a);
        }
        reinterpret_cast<float2*>(S.sc)[(yy & (SC_RING - 1)) * SC_PITCH + c] = make_float2(a, a);
      };
      // zero padding of L2's input in x: only the column left of pixel 2pp and the one right of 2pp+1 can be outside
      const float m0f = (X2 + 2 * pp - 1 >= 0) ? 1.f : 0.f, m3f = (X2 + 2 * pp + 2 < W2) ? 1.f : 0.f;
      const float2 m0 = make_float2(m0f, m0f), m3 = make_float2(m3f, m3f);

      // ---- prologue: image rows of L1 rows r0-1 .. r0+1, then those L1 sums
      {
        const int iyA = 2 * (r0 - 1) - p2;
        for (int i = tid; i < 7 * QUADS; i += THREADS) {
          const int r = i / QUADS, q = i - r * QUADS;
          const int off = a0 + 4 * q;
          const bool inside = iyA + r >= 0 && iyA + r < H && off >= 0 && off < row_bytes_img;
          store_word(load_word(iyA + r, q), iyA + r, inside, q);
        }
        __syncthreads();
        for (int i = tid; i < 3 * (SW2 + 2); i += THREADS) {
          const int yy = r0 - 1 + i / (SW2 + 2);
          if (yy >= 0 && yy < H2) l1_scalar(yy, i % (SW2 + 2));
        }
        __syncthreads();
      }

      for (int i = 0; i <= nrows3 + 1; ++i) {
        const int nL2 = i == 0 ? 1 : (i <= nrows3 ? 2 : 0);
        const int relA = i == 0 ? 0 : 2 * i - 1;                 // L2 rows r0 + relA (, + 1) this step
        const bool hasL3 = i >= 2;                               // act3 row Y0 + i - 2
        const bool next_l2 = i + 1 <= nrows3;
        const int q = r0 + 2 * i + 2;                            // L1 rows q, q+1 are summed in this step for the next one
        // (1) issue the loads of the four image rows those sums add
        uint32_t wv[NW];
        const int iyN = 2 * q - p2 + 1;
        if (next_l2) {
#pragma unroll
          for (int k = 0; k < NW; ++k) {
            const int iy = iyN + wr_[k];
            wv[k] = (wok[k] && (unsigned)iy < (unsigned)H) ? __ldg(reinterpret_cast<const uint32_t*>(wp[k] + (size_t)iy * row_bytes_img)) : 0u;
          }
        }
        // (2) L2 depthwise of the step's rows -> A2 tiles
        if (nL2 > 0) {
          const int rowA = r0 + relA;
          float4 sv[4][2];                                       // L1 sums of columns 2pp-1 .. 2pp+2, each duplicated
          bool rv[4];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int y = rowA - 1 + rr;
            rv[rr] = rr < nL2 + 2 && y >= 0 && y < H2;
            if (rv[rr]) {
              const float4* sp = reinterpret_cast<const float4*>(S.sc) + ((y & (SC_RING - 1)) * SC_PITCH >> 1) + pp;
              sv[rr][0] = sp[0]; sv[rr][1] = sp[1];
            } else {
              sv[rr][0] = sv[rr][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float2* w2 = reinterpret_cast<const float2*>(S.dw2) + (3 * qg + k);     // [tap][12 pairs]
            const float2 pwk = reinterpret_cast<const float2*>(S.pw1)[3 * qg + k], b1k = reinterpret_cast<const float2*>(S.b1)[3 * qg + k];
            float2 w[9], wl[3], wrr[3];
#pragma unroll
            for (int t = 0; t < 9; ++t) w[t] = w2[t * 12];
#pragma unroll
            for (int ti = 0; ti < 3; ++ti) { wl[ti] = fmul2(w[ti * 3], m0); wrr[ti] = fmul2(w[ti * 3 + 2], m3); }
            float2 acc[2][2];
            acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = make_float2(0.f, 0.f);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              if (!rv[rr]) continue;
              float2 a[4];
              a[0] = relu2(ffma2(make_float2(sv[rr][0].x, sv[rr][0].y), pwk, b1k));
              a[1] = relu2(ffma2(make_float2(sv[rr][0].z, sv[rr][0].w), pwk, b1k));
              a[2] = relu2(ffma2(make_float2(sv[rr][1].x, sv[rr][1].y), pwk, b1k));
              a[3] = relu2(ffma2(make_float2(sv[rr][1].z, sv[rr][1].w), pwk, b1k));
              if (rr < 3) {
                acc[0][0] = ffma2(a[0], wl[rr], ffma2(a[1], w[rr * 3 + 1], ffma2(a[2], w[rr * 3 + 2], acc[0][0])));
                acc[0][1] = ffma2(a[1], w[rr * 3], ffma2(a[2], w[rr * 3 + 1], ffma2(a[3], wrr[rr], acc[0][1])));
              }
              if (rr >= 1) {
                acc[1][0] = ffma2(a[0], wl[rr - 1], ffma2(a[1], w[(rr - 1) * 3 + 1], ffma2(a[2], w[(rr - 1) * 3 + 2], acc[1][0])));
                acc[1][1] = ffma2(a[1], w[(rr - 1) * 3], ffma2(a[2], w[(rr - 1) * 3 + 1], ffma2(a[3], wrr[rr - 1], acc[1][1])));
              }
            }
            // channel pair 3qg + k of both pixels -> A2 (8-byte stores, unit-stride in M across the warp)
            const int c2 = 3 * qg + k;
            uint8_t* dst = S.A2 + (c2 >> 1) * A_PLANE + (c2 & 1) * 8 + pp * 16;
            *reinterpret_cast<float2*>(dst) = rna2(acc[0][0]);
            *reinterpret_cast<float2*>(dst + 64 * 16) = rna2(acc[0][1]);
            if (nL2 == 2) {
              *reinterpret_cast<float2*>(dst + A_TILE) = rna2(acc[1][0]);
              *reinterpret_cast<float2*>(dst + A_TILE + 64 * 16) = rna2(acc[1][1]);
            }
          }
        }
        // (3) L3 depthwise (stride 2) of act3 row Y0 + i - 2 from the ring -> A3 tile
        if (hasL3) {
          const int j = pp;
          const int rel0 = 2 * (i - 2);                          // ring rows rel0 .. rel0 + 2
          const float2* ring = reinterpret_cast<const float2*>(S.r2);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const int c2 = qg + 4 * r, g = c2 >> 1, hf = c2 & 1;
            const float2* w2 = reinterpret_cast<const float2*>(S.dw3) + c2;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int ti = 0; ti < 3; ++ti) {
              const float2* row = ring + ((((rel0 + ti) % 3) * R2_SLOT + g * R2_PITCH) << 1) + hf;
              const float2 v0 = row[2 * j], v1 = row[2 * (boff + j)], v2 = row[2 * (j + 1)];
              acc = ffma2(v0, w2[(ti * 3 + 0) * 12], ffma2(v1, w2[(ti * 3 + 1) * 12], ffma2(v2, w2[(ti * 3 + 2) * 12], acc)));
            }
            *reinterpret_cast<float2*>(S.A3 + g * A_PLANE + j * 16 + hf * 8) = rna2(acc);
          }
        }
        // (4) the image words of (1) have arrived: preprocess into the ring
        if (next_l2) {
#pragma unroll
          for (int k = 0; k < NW; ++k) {
            if (wr_[k] < 0) continue;
            const int iy = iyN + wr_[k];
            store_word(wv[k], iy, wok[k] && (unsigned)iy < (unsigned)H, wq_[k]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                         // S1
        tc::tc_fence_after();
        // (5) pointwise convs of both layers on the tensor cores
        if (warp == 0) {
          if (tc::elect_one()) {
            for (int r = 0; r < nL2; ++r)
#pragma unroll
              for (int kp = 0; kp < 3; ++kp)
                tc::umma_tf32(tmem_base + r * tc::UMMA_N, tc::make_desc(a2_0 + ((r * A_TILE + kp * 2 * A_PLANE) >> 4), tc::DESC_HI),
                              tc::make_desc(b2_0 + ((kp * tc::B_TILE_BYTES) >> 4), tc::DESC_HI), kp != 0);
            if (hasL3)
#pragma unroll
              for (int kp = 0; kp < 3; ++kp)
                tc::umma_tf32(tmem_base + 2 * tc::UMMA_N, tc::make_desc(a3_0 + ((kp * 2 * A_PLANE) >> 4), tc::DESC_HI),
                              tc::make_desc(b3_0 + ((kp * tc::B_TILE_BYTES) >> 4), tc::DESC_HI), kp != 0);
            tc::umma_commit(tc::smem_u32(&S.mma_bar));
          }
          __syncwarp();
        }
        // (6) L1 sums of rows q, q+1 (next step's new input rows) while the MMAs run
        if (next_l2) {
          if (q + sc_r0 >= 0 && q + sc_r0 < H2) l1_scalar(q + sc_r0, sc_c0);
          if (tid < 2 * (SW2 + 2) - THREADS && q + 1 >= 0 && q + 1 < H2) l1_scalar(q + 1, THREADS - (SW2 + 2) + tid);
        }
        // (7) epilogues
        wait_bar(&S.mma_bar, mma_count & 1, gerr);
        ++mma_count;
        const int quad = warp & 3;
        if ((warp >> 2) < nL2) {
          const int t = warp >> 2;                               // tile = row relA + t
          const int rel = relA + t, y2 = r0 + rel;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + t * tc::UMMA_N;
          uint32_t v[24];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(taddr));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23])
                       : "r"(taddr + 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int m = quad * 32 + lane;                        // M row -> strip pixel
          const int p = 2 * (m & 63) + (m >> 6);
          const bool inside = y2 >= 0 && y2 < H2 && X2 + p < W2;
          float4 o[UBD_NG];
          const float2* b2p = reinterpret_cast<const float2*>(S.b2);
#pragma unroll
          for (int g = 0; g < UBD_NG; ++g) {
            const float2 lo = relu2(fadd2(make_float2(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1])), b2p[2 * g]));
            const float2 hi = relu2(fadd2(make_float2(__uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3])), b2p[2 * g + 1]));
            o[g] = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
          if (!inside) {
#pragma unroll
            for (int g = 0; g < UBD_NG; ++g) o[g] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const int idx = p + p2;                                // parity is uniform per warp (quadrants 0,1: even pixels)
          float4* dst = &S.r2[(rel % 3) * R2_SLOT + ((idx & 1) ? boff + (idx >> 1) : (idx >> 1))];
#pragma unroll
          for (int g = 0; g < UBD_NG; ++g) dst[g * R2_PITCH] = o[g];
          // the column the next strip needs / the one the previous strip left for this row
          if (p == (p2 ? SW2 - 1 : 0)) {
#pragma unroll
            for (int g = 0; g < UBD_NG; ++g) carry_out[rel * UBD_NG + g] = o[g];
          }
          if (p == (p2 ? 0 : SW2 - 1)) {
            float4* cd = &S.r2[(rel % 3) * R2_SLOT + (p2 ? 0 : SW3)];
#pragma unroll
            for (int g = 0; g < UBD_NG; ++g) cd[g * R2_PITCH] = si == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : carry_in[rel * UBD_NG + g];
          }
        }
        if (hasL3 && quad < 2) {
          // act3 row: quadrant = 32 pixels, warp >> 2 = channels 0..11 / 12..23
          const int ch = warp >> 2;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + 2 * tc::UMMA_N + 12 * ch;
          uint32_t v[12];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(taddr));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11])
                       : "r"(taddr + 8));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int j = quad * 32 + lane;
          if (j < w3s) {
            const int y3 = Y0 + i - 2, x3 = s * SW3 + j;
            float o[12];
#pragma unroll
            for (int c = 0; c < 12; ++c) o[c] = fmaxf(__uint_as_float(v[c]) + S.b3[12 * ch + c], 0.f);
            if constexpr (OUT_MODE == 2) {
              const bool f16 = S.b3[tc::F16_FLAG_SLOT] != 0.f;
              // 16-bit planes hold 8 channels: channels 0..11 = plane 0 + low half of plane 1, 12..23 = high half + plane 2
              uint4* dst = reinterpret_cast<uint4*>(act3) + (((size_t)n * H4 + y3) * tc::NG_BF16) * (size_t)(W4 + 2 * UBD_MAP_PAD) + UBD_MAP_PAD + x3;
              const size_t ps = (size_t)(W4 + 2 * UBD_MAP_PAD);
              if (ch == 0) {
                dst[0] = make_uint4(tc::pack16(o[0], o[1], f16), tc::pack16(o[2], o[3], f16), tc::pack16(o[4], o[5], f16), tc::pack16(o[6], o[7], f16));
                *reinterpret_cast<uint2*>(dst + ps) = make_uint2(tc::pack16(o[8], o[9], f16), tc::pack16(o[10], o[11], f16));
              } else {
                *(reinterpret_cast<uint2*>(dst + ps) + 1) = make_uint2(tc::pack16(o[0], o[1], f16), tc::pack16(o[2], o[3], f16));
                dst[2 * ps] = make_uint4(tc::pack16(o[4], o[5], f16), tc::pack16(o[6], o[7], f16), tc::pack16(o[8], o[9], f16), tc::pack16(o[10], o[11], f16));
              }
            } else {
#pragma unroll
              for (int g = 0; g < 3; ++g)
                act3[act_index(n, 3 * ch + g, y3, x3, H4, W4, UBD_MAP_PAD)] =
                    make_float4(rna_bits(o[4 * g]), rna_bits(o[4 * g + 1]), rna_bits(o[4 * g + 2]), rna_bits(o[4 * g + 3]));
            }
          }
        }
        tc::tc_fence_before();
        __syncthreads();                                         // S2
        tc::tc_fence_after();
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

}  // namespace stemf

static void stemf_setup_attributes() {
  const int smem = (int)sizeof(stemf::Smem) + 128;
#define UBD_STEMF_ATTR(T, M)                                                                                      \
  cudaFuncSetAttribute(stemf::stem_fused_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);        \
  cudaFuncSetAttribute(stemf::stem_fused_kernel<T, M>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
  UBD_STEMF_ATTR(uint8_t, 1); UBD_STEMF_ATTR(uint8_t, 2); UBD_STEMF_ATTR(float, 1); UBD_STEMF_ATTR(float, 2);
#undef UBD_STEMF_ATTR
}

// image -> act3 (grey input).  Needs the pointwise B images of L2 / L3 (h->stem_wimg, run_stem_tc builds them).
static int stemf_launch(ubd_handle h, const void* d_img, int in_dtype, int preproc, int n, int H, int W, float4* act3) {
  const bool mob = preproc == UBD_PREPROC_MOBILENET;
  const int p2 = stride2_pad(h);
  const size_t smem = sizeof(stemf::Smem) + 128;
  const int nitems = n * ((H / 4 + stemf::BH - 1) / stemf::BH);
  const int grid = std::min(nitems, 2 * h->n_sm);
  const uint8_t* wb2 = (const uint8_t*)h->stem_wimg.p;
  const uint8_t* wb3 = wb2 + stem::PW_WB_BYTES;
  int* counter = reinterpret_cast<int*>((uint8_t*)h->stem_wimg.p + 2 * stem::PW_WB_BYTES);
  UBD_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), h->stream));
  const bool bf16 = ubd_is16(h);
#define UBD_STEMF_LAUNCH(T, M, LUT, PS, PSH)                                                                                   \
  stemf::stem_fused_kernel<T, M><<<grid, stemf::THREADS, smem, h->stream>>>(                                                  \
      (const T*)d_img, act3, h->d_params, h->spec.off[0], h->spec.off[1], h->spec.off[2], h->spec.off[3], h->spec.off[6], wb2, \
      wb3, LUT, PS, PSH, n, H, W, p2, counter, tc_err_flag(h))
  if (in_dtype == UBD_U8) {
    const float* lut = mob ? h->d_lut : nullptr;
    if (bf16) UBD_STEMF_LAUNCH(uint8_t, 2, lut, 0.f, 0.f); else UBD_STEMF_LAUNCH(uint8_t, 1, lut, 0.f, 0.f);
  } else {
    const float ps = mob ? 127.5f : 0.f;
    if (bf16) UBD_STEMF_LAUNCH(float, 2, nullptr, ps, 127.5f); else UBD_STEMF_LAUNCH(float, 1, nullptr, ps, 127.5f);
  }
#undef UBD_STEMF_LAUNCH
  ++h->launches;
  UBD_CUDA(cudaGetLastError());
  return UBD_OK;
}
