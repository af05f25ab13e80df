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
constexpr int IMG_RING = 8, IMG_HALF = 132;   // preprocessed image rows, even / odd patch columns
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
  float imgE[IMG_RING * IMG_HALF], imgO[IMG_RING * IMG_HALF];
  float sc[SC_RING * SC_PITCH];
  __align__(16) float dw2[9 * UBD_NF];
  __align__(16) float dw3[9 * UBD_NF];
  __align__(16) float zeros[9 * UBD_NF];
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

// OUT_MODE 1: fp32 planes rounded to the tf32 grid; 2: bf16, 3 planes of 8 channels.
template <typename TIn, int OUT_MODE>
__global__ void __launch_bounds__(THREADS, 2)
stem_fused_kernel(const TIn* __restrict__ img, float4* __restrict__ act3, const float* __restrict__ params,
                  int64_t off_dw1, int64_t off_pw1, int64_t off_b1, int64_t off_dw2, int64_t off_dw3,
                  const uint8_t* __restrict__ wb2, const uint8_t* __restrict__ wb3,
                  const float* __restrict__ lut, float pre_scale, float pre_shift,
                  int N, int H, int W, int p2, int* __restrict__ work_counter, int* gerr) {
  constexpr int ELT = (int)sizeof(TIn);
  constexpr int EPQ = 4 / ELT;                                   // elements per 4-byte word
  constexpr int QUADS = (IMG_COLS * ELT + 3) / 4 + 1;            // words covering one patch row
  constexpr int NW = (4 * QUADS + THREADS - 1) / THREADS;        // words per thread for the 4 new image rows of a step
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;

  for (int i = tid; i < 9; i += THREADS) S.dw1[i] = params[off_dw1 + i];
  for (int i = tid; i < UBD_NF; i += THREADS) { S.pw1[i] = params[off_pw1 + i]; S.b1[i] = params[off_b1 + i]; }
  for (int i = tid; i < 9 * UBD_NF; i += THREADS) { S.dw2[i] = params[off_dw2 + i]; S.dw3[i] = params[off_dw3 + i]; S.zeros[i] = 0.f; }
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
  const int px = tid & (SW2 - 1);                                // L2 depthwise: pixel column of the strip
  const int gb = (tid >> 7) * 3;                                 // ... and its three channel planes
  float pw1r[12], b1r[12];
#pragma unroll
  for (int c = 0; c < 12; ++c) { pw1r[c] = S.pw1[4 * gb + c]; b1r[c] = S.b1[4 * gb + c]; }
  float dw1r[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) dw1r[i] = S.dw1[i];
  const int row_bytes_img = W * ELT;
  const int boff = p2 ? 69 : 68;                                 // bank-conflict-free parity split for this padding mode
  const int ns = (W4 + SW3 - 1) / SW3;
  const int nbands = (H4 + BH - 1) / BH;
  const int nitems = N * nbands;
  uint32_t mma_count = 0;

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
      auto load_word = [&](int iy, int q) -> uint32_t {
        const int off = a0 + 4 * q;
        if (iy < 0 || iy >= H || off < 0 || off >= row_bytes_img) return 0u;
        return __ldg(reinterpret_cast<const uint32_t*>(img_n + (size_t)iy * row_bytes_img + off));
      };
      auto store_word = [&](uint32_t v, int iy, int q) {
        const int off = a0 + 4 * q;
        const bool inside = iy >= 0 && iy < H && off >= 0 && off < row_bytes_img;
        const int slot = iy & (IMG_RING - 1);
        const TIn* e = reinterpret_cast<const TIn*>(&v);
#pragma unroll
        for (int k = 0; k < EPQ; ++k) {
          const int col = e0 + q * EPQ + k;
          if ((unsigned)col >= (unsigned)IMG_COLS) continue;
          float f = 0.f;
          if (inside) {
            if constexpr (sizeof(TIn) == 1) f = S.lut[(int)e[k]];
            else { f = (float)e[k]; if (pre_scale != 0.f) f = (f - pre_shift) / pre_scale; }
          }
          ((col & 1) ? S.imgO : S.imgE)[slot * IMG_HALF + (col >> 1)] = f;
        }
      };
      // L1 depthwise sum of map pixel (yy, local column c); rows outside the map are never read
      auto l1_scalar = [&](int yy, int c) {
        float a = 0.f;
#pragma unroll
        for (int ti = 0; ti < 3; ++ti) {
          const int slot = (2 * yy - p2 + ti) & (IMG_RING - 1);
          const float* E = &S.imgE[slot * IMG_HALF + c];
          const float* O = &S.imgO[slot * IMG_HALF + c];
          a = fmaf(E[0], dw1r[ti * 3 + 0], a);
          a = fmaf(O[0], dw1r[ti * 3 + 1], a);
          a = fmaf(E[1], dw1r[ti * 3 + 2], a);
        }
        S.sc[(yy & (SC_RING - 1)) * SC_PITCH + c] = a;
      };
      // column validity of this thread's three L1 columns (zero padding of L2's input in x)
      const int xg = X2 + px;
      const float4* wl = reinterpret_cast<const float4*>((xg - 1 >= 0 && xg - 1 < W2) ? S.dw2 : S.zeros);
      const float4* wc = reinterpret_cast<const float4*>((xg < W2) ? S.dw2 : S.zeros);
      const float4* wr = reinterpret_cast<const float4*>((xg + 1 < W2) ? S.dw2 : S.zeros);

      // ---- prologue: image rows of L1 rows r0-1 .. r0+1, then those L1 sums
      {
        const int iyA = 2 * (r0 - 1) - p2;
        for (int i = tid; i < 7 * QUADS; i += THREADS) {
          const int r = i / QUADS, q = i - r * QUADS;
          store_word(load_word(iyA + r, q), iyA + r, q);
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
        uint32_t wq[NW];
        const int iyN = 2 * q - p2 + 1;
        if (next_l2) {
#pragma unroll
          for (int k = 0; k < NW; ++k) {
            const int idx = tid + k * THREADS;
            wq[k] = idx < 4 * QUADS ? load_word(iyN + idx / QUADS, idx % QUADS) : 0u;
          }
        }
        // (2) L2 depthwise of the step's rows -> A2 tiles
        if (nL2 > 0) {
          const int rowA = r0 + relA;
          float sv[4][3];
          bool rv[4];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int y = rowA - 1 + rr;
            rv[rr] = rr < nL2 + 2 && y >= 0 && y < H2;
            if (rv[rr]) {
              const float* sp = &S.sc[(y & (SC_RING - 1)) * SC_PITCH + px];
              sv[rr][0] = sp[0]; sv[rr][1] = sp[1]; sv[rr][2] = sp[2];
            } else {
              sv[rr][0] = sv[rr][1] = sv[rr][2] = 0.f;
            }
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int g = gb + k;
            float4 wt[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) wt[t] = (t % 3 == 0 ? wl : (t % 3 == 1 ? wc : wr))[t * UBD_NG + g];
            float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              if (!rv[rr]) continue;
              float4 a[3];
#pragma unroll
              for (int tj = 0; tj < 3; ++tj) {
                a[tj].x = fmaxf(fmaf(sv[rr][tj], pw1r[4 * k + 0], b1r[4 * k + 0]), 0.f);
                a[tj].y = fmaxf(fmaf(sv[rr][tj], pw1r[4 * k + 1], b1r[4 * k + 1]), 0.f);
                a[tj].z = fmaxf(fmaf(sv[rr][tj], pw1r[4 * k + 2], b1r[4 * k + 2]), 0.f);
                a[tj].w = fmaxf(fmaf(sv[rr][tj], pw1r[4 * k + 3], b1r[4 * k + 3]), 0.f);
              }
              if (rr < 3) {
#pragma unroll
                for (int tj = 0; tj < 3; ++tj) {
                  const float4 w = wt[rr * 3 + tj];
                  acc0.x = fmaf(a[tj].x, w.x, acc0.x); acc0.y = fmaf(a[tj].y, w.y, acc0.y);
                  acc0.z = fmaf(a[tj].z, w.z, acc0.z); acc0.w = fmaf(a[tj].w, w.w, acc0.w);
                }
              }
              if (rr >= 1) {
#pragma unroll
                for (int tj = 0; tj < 3; ++tj) {
                  const float4 w = wt[(rr - 1) * 3 + tj];
                  acc1.x = fmaf(a[tj].x, w.x, acc1.x); acc1.y = fmaf(a[tj].y, w.y, acc1.y);
                  acc1.z = fmaf(a[tj].z, w.z, acc1.z); acc1.w = fmaf(a[tj].w, w.w, acc1.w);
                }
              }
            }
            reinterpret_cast<float4*>(S.A2 + g * A_PLANE)[px] =
                make_float4(rna_bits(acc0.x), rna_bits(acc0.y), rna_bits(acc0.z), rna_bits(acc0.w));
            if (nL2 == 2)
              reinterpret_cast<float4*>(S.A2 + A_TILE + g * A_PLANE)[px] =
                  make_float4(rna_bits(acc1.x), rna_bits(acc1.y), rna_bits(acc1.z), rna_bits(acc1.w));
          }
        }
        // (3) L3 depthwise (stride 2) of act3 row Y0 + i - 2 from the ring -> A3 tile
        if (hasL3 && tid < 3 * SW3) {
          const int j = tid & (SW3 - 1), pp = tid >> 6;
          const int rel0 = 2 * (i - 2);                          // ring rows rel0 .. rel0 + 2
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int g = 2 * pp + k;
            const float4* w4 = reinterpret_cast<const float4*>(S.dw3) + g;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int ti = 0; ti < 3; ++ti) {
              const float4* row = &S.r2[((rel0 + ti) % 3) * R2_SLOT + g * R2_PITCH];
              const float4 v0 = row[j], v1 = row[boff + j], v2 = row[j + 1];
              const float4 w0 = w4[(ti * 3 + 0) * UBD_NG], w1 = w4[(ti * 3 + 1) * UBD_NG], w2 = w4[(ti * 3 + 2) * UBD_NG];
              acc.x = fmaf(v0.x, w0.x, fmaf(v1.x, w1.x, fmaf(v2.x, w2.x, acc.x)));
              acc.y = fmaf(v0.y, w0.y, fmaf(v1.y, w1.y, fmaf(v2.y, w2.y, acc.y)));
              acc.z = fmaf(v0.z, w0.z, fmaf(v1.z, w1.z, fmaf(v2.z, w2.z, acc.z)));
              acc.w = fmaf(v0.w, w0.w, fmaf(v1.w, w1.w, fmaf(v2.w, w2.w, acc.w)));
            }
            reinterpret_cast<float4*>(S.A3 + g * A_PLANE)[j] =
                make_float4(rna_bits(acc.x), rna_bits(acc.y), rna_bits(acc.z), rna_bits(acc.w));
          }
        }
        // (4) the image words of (1) have arrived: preprocess into the ring
        if (next_l2) {
#pragma unroll
          for (int k = 0; k < NW; ++k) {
            const int idx = tid + k * THREADS;
            if (idx < 4 * QUADS) store_word(wq[k], iyN + idx / QUADS, idx % QUADS);
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
          for (int idx = tid; idx < 2 * (SW2 + 2); idx += THREADS) {
            const int yy = q + idx / (SW2 + 2);
            if (yy >= 0 && yy < H2) l1_scalar(yy, idx % (SW2 + 2));
          }
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
          const int p = quad * 32 + lane;                        // L2 pixel of the strip
          const bool inside = y2 >= 0 && y2 < H2 && X2 + p < W2;
          float4 o[UBD_NG];
#pragma unroll
          for (int g = 0; g < UBD_NG; ++g) {
            o[g].x = inside ? fmaxf(__uint_as_float(v[4 * g + 0]) + S.b2[4 * g + 0], 0.f) : 0.f;
            o[g].y = inside ? fmaxf(__uint_as_float(v[4 * g + 1]) + S.b2[4 * g + 1], 0.f) : 0.f;
            o[g].z = inside ? fmaxf(__uint_as_float(v[4 * g + 2]) + S.b2[4 * g + 2], 0.f) : 0.f;
            o[g].w = inside ? fmaxf(__uint_as_float(v[4 * g + 3]) + S.b2[4 * g + 3], 0.f) : 0.f;
          }
          const int idx = p + p2;
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
        if (hasL3 && (warp == 4 || warp == 5)) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + 2 * tc::UMMA_N;
          uint32_t v[24];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(taddr));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23])
                       : "r"(taddr + 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int j = quad * 32 + lane;
          if (j < w3s) {
            const int y3 = Y0 + i - 2, x3 = s * SW3 + j;
            float o[UBD_NF];
#pragma unroll
            for (int c = 0; c < UBD_NF; ++c) o[c] = fmaxf(__uint_as_float(v[c]) + S.b3[c], 0.f);
            if constexpr (OUT_MODE == 2) {
              uint4* dst = reinterpret_cast<uint4*>(act3) + (((size_t)n * H4 + y3) * tc::NG_BF16) * (size_t)(W4 + 2 * UBD_MAP_PAD) + UBD_MAP_PAD + x3;
#pragma unroll
              for (int g = 0; g < tc::NG_BF16; ++g)
                dst[(size_t)g * (W4 + 2 * UBD_MAP_PAD)] =
                    make_uint4(tc::pack_bf16x2(o[8 * g], o[8 * g + 1]), tc::pack_bf16x2(o[8 * g + 2], o[8 * g + 3]),
                               tc::pack_bf16x2(o[8 * g + 4], o[8 * g + 5]), tc::pack_bf16x2(o[8 * g + 6], o[8 * g + 7]));
            } else {
#pragma unroll
              for (int g = 0; g < UBD_NG; ++g)
                act3[act_index(n, g, y3, x3, H4, W4, UBD_MAP_PAD)] =
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
  const bool bf16 = h->precision == UBD_BF16;
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
