// Dilated 3x3 24->24 layers (net.py:298-304) on tcgen05, "row-rotating" formulation, sm_100a only.
//
// ubd_tc.cuh puts the pixels on the M side of the MMA (D[128 px, 32 oc]); with only 24 output
// channels that shape reads 5 KB of shared memory per 32 K-MACs and is operand-feed bound at 1/3 of
// the tensor peak.  Here the roles are swapped and the three kernel rows share one instruction:
//
//   D[128 = 4 lane groups x 32 oc, N <= 256 px]  +=  A[128, K] (weights)  *  B[K, N] (one staged map row)
//
// One input row r of a y-phase (rows c, c+d, c+2d, ... form a 1-dilated problem) contributes to the
// output rows r+1, r, r-1 through the kernel rows ky = 0, 1, 2.  Output row o owns TMEM lane group
// o mod 4; the weight image of an MMA holds the blocks [ky0, 0, ky2, ky1] twice in a row, so starting
// the A descriptor at group (-r mod 4) rotates them onto the right lane groups: the three partial
// products of an input row are accumulated by ONE set of 9 MMAs (3 dx x 3 K-blocks of 8 tf32; bf16: 5
// MMAs of K = 16) over N = 256 pixels, every staged row is read from shared memory once instead of three
// times, and an MMA moves 12 KB for 262 K-MACs (tensor-bound).  The fourth lane group is the one whose
// output row was completed by the previous input row: its lanes are masked out of the MMA
// (disable-output-lane) while the drain warp of that quadrant copies it out and re-arms it, so
// draining overlaps the next row's MMAs without a second accumulator.
//
// Warp roles (320 threads, 1 CTA/SM, each CTA owns a contiguous range of the global row sequence):
//   warp 0      producer : cp.async.bulk of map rows into the slot ring          (empty[] -> full[])
//   warp 1      MMA issue: per input row 9 (5) tcgen05.mma + commits to empty[slot] and gfull[group]
//   warps 4-7   drainers : quadrant q = lane group q: tcgen05.ld of the finished row -> raw fp32 tile in
//               shared memory, tcgen05.st of the bias (the accumulator's initial value) -> gempty[group]
//   warps 2,3,8,9 finishers: ReLU, rounding / packing (or the fused 1x1 head + threshold), coalesced
//               16-byte stores.  TMEM quadrant q is only reachable from SM sub-partition q, so the drain
//               itself is one store per element and everything else runs on all four sub-partitions.
#pragma once
#include "ubd_tc.cuh"

namespace tc3 {

using tc::smem_u32; using tc::elect_one; using tc::mbar_init; using tc::mbar_arrive; using tc::mbar_expect_tx;
using tc::mbar_wait; using tc::bulk_g2s; using tc::umma_commit; using tc::tc_fence_before; using tc::tc_fence_after;
using tc::make_desc; using tc::round_tf32; using tc::pack_bf16x2; using tc::HeadArgs;

// Bounded wait like tc::mbar_wait; on a stall every warp leaves (code << 24 | info) in gerr[1 + warp].
__device__ __forceinline__ bool mbar_wait3(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* gerr, int code, uint32_t info) {
  const long long t0 = clock64();
  bool ok = true;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (*abort_flag || clock64() - t0 > 1500000000LL) {
      atomicCAS(gerr, 0, code);
      *abort_flag = 1;
      gerr[1 + (threadIdx.x >> 5)] = (code << 24) | (int)(info & 0xFFFFFFu);
      ok = false;
      break;
    }
  }
  return __all_sync(0xffffffffu, ok);
}

// Waits until all four finisher warps have written out `need` rows (dump ring release).
__device__ __forceinline__ bool wait_released(const uint32_t* cnt, uint32_t need, volatile int* abort_flag, int* gerr, int code, uint32_t info) {
  const long long t0 = clock64();
  bool ok = true;
  while (true) {
    const volatile uint32_t* vc = cnt;
    const uint32_t v0 = vc[0], v1 = vc[1], v2 = vc[2], v3 = vc[3];
    if (min(min(v0, v1), min(v2, v3)) >= need) break;
    if (*abort_flag || clock64() - t0 > 1500000000LL) {
      atomicCAS(gerr, 0, code);
      *abort_flag = 1;
      gerr[1 + (threadIdx.x >> 5)] = (code << 24) | (int)(info & 0xFFFFFFu);
      ok = false;
      break;
    }
  }
  return __all_sync(0xffffffffu, ok);
}

constexpr int PAD = UBD_MAP_PAD;
constexpr int NW_MAX = 256;                               // widest strip = UMMA N
constexpr int GROUP_BYTES = 32 * 16;                      // one lane group of an A K-core: 32 rows x 16 B
constexpr int A_KCORE_BYTES = 7 * GROUP_BYTES;            // blocks [ky0,0,ky2,ky1,ky0,0,ky2] (rotation by start address)
constexpr int A_IMG_BYTES = 2 * A_KCORE_BYTES;            // 7168 per MMA
constexpr int N_MMA_TF32 = 9, N_MMA_BF16 = 5;
constexpr int W_BYTES_TF32 = N_MMA_TF32 * A_IMG_BYTES;    // 64512
constexpr int W_BYTES_BF16 = N_MMA_BF16 * A_IMG_BYTES;    // 35840
constexpr int WB_BYTES_TF32 = W_BYTES_TF32 + 128;         // + bias[32]
constexpr int WB_BYTES_BF16 = W_BYTES_BF16 + 128;
constexpr int SLOT_BYTES_TF32 = UBD_NG * (NW_MAX + 2 * PAD) * 16;      // 27648
constexpr int SLOT_BYTES_BF16 = 3 * (NW_MAX + 2 * PAD) * 16;           // 13824
constexpr int NS_TF32 = 4, NS_BF16 = 6;
constexpr int DUMP_PLANE = NW_MAX + 1;                    // units of 16 B; +1 keeps the drain's 4-byte stores conflict-free
constexpr int DUMP_BYTES = UBD_NG * DUMP_PLANE * 16;      // 24672
constexpr int ND = 2;
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 512;                            // two accumulator tiles of 256 columns (alternating pieces)

template <bool BF16> struct Smem {
  static constexpr int NS = BF16 ? NS_BF16 : NS_TF32;
  static constexpr int SLOT = BF16 ? SLOT_BYTES_BF16 : SLOT_BYTES_TF32;
  static constexpr int WB = BF16 ? WB_BYTES_BF16 : WB_BYTES_TF32;
  uint8_t slots[NS * SLOT];
  uint8_t wimg[WB];                                       // A images, then bias[32]
  uint8_t dump[ND * DUMP_BYTES];
  float headw[UBD_NF * (1 + UBD_MAX_CLASSES) + 1 + UBD_MAX_CLASSES];
  uint64_t full[NS], empty[NS], gfull[8], gempty[8], dfull[ND], wbar;
  // rows completely written out by each finisher warp.  A counter, not an mbarrier: the four drainers take
  // turns on the dump ring, so a drainer can be several phases behind and a parity wait would be ambiguous.
  alignas(16) uint32_t fin_done[4];
  uint32_t tmem_base;
  int abort_flag;
};

// A contiguous run of output rows inside one (image, strip, y-phase).
struct Piece { int n, x0, nw, c, j0, rows, R; };

struct Walk {
  long long t, t1;
  int h, w, d, sw, n_strips, q, rem;
  __device__ Walk(int n_imgs, int h_, int w_, int d_, int sw_, int cta, int n_cta) : h(h_), w(w_), d(d_), sw(sw_) {
    n_strips = (w + sw - 1) / sw;
    q = h / d; rem = h % d;
    const long long total = (long long)n_imgs * n_strips * h;
    t = total * cta / n_cta;
    t1 = total * (cta + 1) / n_cta;
  }
  __device__ bool next(Piece& p) {
    if (t >= t1) return false;
    const long long is = t / h;
    const int pos = (int)(t - is * h);
    p.n = (int)(is / n_strips);
    p.x0 = (int)(is % n_strips) * sw;
    p.nw = min(sw, w - p.x0);
    const int big = rem * (q + 1);
    if (pos < big) { p.c = pos / (q + 1); p.j0 = pos % (q + 1); p.R = q + 1; }
    else { const int p2 = pos - big; p.c = rem + p2 / q; p.j0 = p2 % q; p.R = q; }
    p.rows = (int)min((long long)(p.R - p.j0), t1 - t);
    t += p.rows;
    return true;
  }
};

__device__ __forceinline__ void umma_masked(bool bf16, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
  if (bf16)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %5, %6, %7}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(m0), "r"(m1), "r"(m2), "r"(m3) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%4, %5, %6, %7}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(m0), "r"(m1), "r"(m2), "r"(m3) : "memory");
}

#define UBD_LDTM32(v, taddr)                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),      \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
               : "r"(taddr))
#define UBD_LDTM16(v, taddr)                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),      \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
               : "r"(taddr))
// 16 columns of one value (the bias of this lane's output channel)
#define UBD_STTM16(taddr, b)                                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"     \
               ::"r"(taddr), "r"(b) : "memory")

// in / out: padded row-interleaved maps (pad = PAD) in 16-byte units (tf32: 6 planes of float4, bf16: 3
// planes of 8 x bf16).  wb: this layer's A images followed by bias[32].  sw: strip width (w if w <= 256).
// out_mode 0: same format as the input (tf32 rna / bf16), 1: fp32 6-plane unrounded, 2: fused 1x1 head +
// logit threshold (net.py:307-311, model_runner.py:124): the last map never reaches HBM.
template <bool BF16>
__global__ void __launch_bounds__(THREADS, 1)
dilconv_rot_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, const uint8_t* __restrict__ wb,
                   int n_imgs, int h, int w, int d, int sw, int out_mode, int out_pad, int* gerr, HeadArgs head, long long* trace) {
  using S_t = Smem<BF16>;
  constexpr int NS = S_t::NS;
  constexpr int NGI = BF16 ? 3 : UBD_NG;
  constexpr uint32_t WBB = S_t::WB;
  constexpr uint32_t WBYTES = BF16 ? W_BYTES_BF16 : W_BYTES_TF32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  S_t& S = *reinterpret_cast<S_t*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  volatile int* abort_flag = &S.abort_flag;
  // optional event trace of CTA 0 (tuning): trace[role][event][4] cycle stamps
  const bool tr = trace != nullptr && blockIdx.x == 0 && lane == 0;
  int tr_n = 0;
#define TC3_TRACE(role, slot) do { if (tr && tr_n < 1024) trace[((role) * 1024 + tr_n) * 4 + (slot)] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&S.full[i]), 1); mbar_init(smem_u32(&S.empty[i]), 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&S.gfull[i]), 1); mbar_init(smem_u32(&S.gempty[i]), 1); }
    for (int i = 0; i < ND; ++i) mbar_init(smem_u32(&S.dfull[i]), 1);
    for (int i = 0; i < 4; ++i) S.fin_done[i] = 0u;
    mbar_init(smem_u32(&S.wbar), 1);
    S.abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, S.tmem_base, 0);
  const bool drainer = warp >= 4 && warp < 8;
  const int quad = warp & 3;
  // The accumulators start from the bias: every lane of a group is one output channel.
  uint32_t bias_bits = 0u;
  if (drainer) {
    bias_bits = lane < UBD_NF ? __float_as_uint(__ldg(reinterpret_cast<const float*>(wb + WBYTES) + lane)) : 0u;
    const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c0 = 0; c0 < TMEM_COLS; c0 += 16) UBD_STTM16(t0 + c0, bias_bits);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t slots0 = smem_u32(S.slots);
  const int wp = w + 2 * PAD;
  const int wpo = w + 2 * out_pad;
  const uint32_t plane_bytes = (uint32_t)(sw + 2 * PAD) * 16;
  const uint32_t slot_bytes = (uint32_t)NGI * plane_bytes;
  const bool one_copy = (w == sw);
  Walk walk(n_imgs, h, w, d, sw, (int)blockIdx.x, (int)gridDim.x);
  Piece pc;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&S.wbar), WBB);
      bulk_g2s(smem_u32(S.wimg), wb, WBB, smem_u32(&S.wbar));
    }
    uint32_t lseq = 0;
    bool ok = true;
    while (ok && walk.next(pc)) {
      const uint32_t copy_bytes = (uint32_t)(pc.nw + 2 * PAD) * 16;
      for (int i = 0; i < pc.rows + 2 && ok; ++i) {
        const int jj = pc.j0 - 1 + i;
        if (jj < 0 || jj >= pc.R) continue;                 // zero row above / below the image: no MMAs at all
        const uint32_t slot = lseq % NS;
        TC3_TRACE(0, 0);
        ok = mbar_wait3(smem_u32(&S.empty[slot]), ((lseq / NS) & 1) ^ 1, abort_flag, gerr, 11, lseq);
        if (!ok) break;
        TC3_TRACE(0, 1);
        const uint32_t bar = smem_u32(&S.full[slot]);
        const uint32_t dst = slots0 + slot * slot_bytes;
        const int y = pc.c + jj * d;
        const uint4* src = in + (((size_t)pc.n * h + y) * NGI) * wp + pc.x0;
        if (elect_one()) {
          if (one_copy) {
            mbar_expect_tx(bar, slot_bytes);
            bulk_g2s(dst, src, slot_bytes, bar);
          } else {
            mbar_expect_tx(bar, (uint32_t)NGI * copy_bytes);
            for (int g = 0; g < NGI; ++g) bulk_g2s(dst + g * plane_bytes, src + (size_t)g * wp, copy_bytes, bar);
          }
        }
        __syncwarp();
        TC3_TRACE(0, 2);
        ++tr_n;
        ++lseq;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    bool ok = mbar_wait(smem_u32(&S.wbar), 0, abort_flag, gerr, 12);
    const uint32_t a_lo0 = ((smem_u32(S.wimg) >> 4) & 0x3FFFu) | ((uint32_t)(A_KCORE_BYTES >> 4) << 16);   // LBO = K-core stride
    const uint32_t plane_units = plane_bytes >> 4;
    const uint32_t b_lbo = (plane_units & 0x3FFFu) << 16;
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);                                                 // SBO = 128 B
    const uint32_t idesc0 = BF16 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24))
                                 : ((1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24));
    uint32_t lseq = 0, par_e = 0, npiece = 0;
    while (ok && walk.next(pc)) {
      const uint32_t tile = npiece & 1u;
      ++npiece;
      const uint32_t tmem_d = tmem_base + tile * 256u;
      const uint32_t nmma = (uint32_t)((pc.nw + 15) & ~15);
      const uint32_t idesc = idesc0 | ((nmma >> 3) << 17);
      for (int i = 0; i < pc.rows + 2 && ok; ++i) {
        TC3_TRACE(1, 0);
        if (i < pc.rows) {
          // output row i gets its first contribution now: its lane group must have been drained and re-armed
          const uint32_t bit = tile * 4u + (uint32_t)(i & 3);
          ok = mbar_wait3(smem_u32(&S.gempty[bit]), ((par_e >> bit) & 1u) ^ 1u, abort_flag, gerr, 13, (npiece << 12) | (uint32_t)i);
          par_e ^= 1u << bit;
          if (!ok) break;
          tc_fence_after();
        }
        TC3_TRACE(1, 1);
        const int jj = pc.j0 - 1 + i;
        const bool valid = jj >= 0 && jj < pc.R;
        uint32_t slot = 0;
        if (valid) {
          slot = lseq % NS;
          ok = mbar_wait3(smem_u32(&S.full[slot]), (lseq / NS) & 1, abort_flag, gerr, 14, (npiece << 12) | (uint32_t)i);
          if (!ok) break;
          ++lseq;
          tc_fence_after();
        }
        // lane group G is written iff it belongs to one of the (existing) output rows i, i-1, i-2
        uint32_t msk[4];
#pragma unroll
        for (int G = 0; G < 4; ++G) {
          const bool en = (i < pc.rows && (i & 3) == G) || (i >= 1 && i - 1 < pc.rows && ((i - 1) & 3) == G) ||
                          (i >= 2 && i - 2 < pc.rows && ((i - 2) & 3) == G);
          msk[G] = en ? 0u : 0xFFFFFFFFu;
        }
        TC3_TRACE(1, 2);
        const uint32_t a_rot = a_lo0 + (uint32_t)((4 - (i & 3)) & 3) * (GROUP_BYTES >> 4);
        const uint32_t b_row = (((slots0 + slot * slot_bytes) >> 4) & 0x3FFFu) + PAD;
        if (elect_one()) {
          if (valid) {
            if constexpr (!BF16) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int kp = 0; kp < 3; ++kp)
                  umma_masked(false, tmem_d, make_desc(a_rot + (uint32_t)(dx * 3 + kp) * (A_IMG_BYTES >> 4), DESC_HI),
                              make_desc((b_row + (uint32_t)((dx - 1) * d) + (uint32_t)kp * 2u * plane_units) | b_lbo, DESC_HI),
                              idesc, msk[0], msk[1], msk[2], msk[3]);
            } else {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)               // planes 0 + 1 (channels 0..15) of tap dx
                umma_masked(true, tmem_d, make_desc(a_rot + (uint32_t)dx * (A_IMG_BYTES >> 4), DESC_HI),
                            make_desc((b_row + (uint32_t)((dx - 1) * d)) | b_lbo, DESC_HI), idesc, msk[0], msk[1], msk[2], msk[3]);
              // plane 2 of taps dx = -1, 0: the two K cores are d pixels apart in the same plane (LBO = d px)
              umma_masked(true, tmem_d, make_desc(a_rot + 3u * (A_IMG_BYTES >> 4), DESC_HI),
                          make_desc((b_row + (uint32_t)(-d) + 2u * plane_units) | ((uint32_t)d << 16), DESC_HI), idesc,
                          msk[0], msk[1], msk[2], msk[3]);
              // plane 2 of tap dx = +1; the second K core of the A image is zero, LBO = 0 re-reads the same core
              umma_masked(true, tmem_d, make_desc(a_rot + 4u * (A_IMG_BYTES >> 4), DESC_HI),
                          make_desc(b_row + (uint32_t)d + 2u * plane_units, DESC_HI), idesc, msk[0], msk[1], msk[2], msk[3]);
            }
            umma_commit(smem_u32(&S.empty[slot]));
          }
          if (i >= 2) umma_commit(smem_u32(&S.gfull[tile * 4u + (uint32_t)((i - 2) & 3)]));
        }
        __syncwarp();
        TC3_TRACE(1, 3);
        ++tr_n;
      }
    }
  } else if (drainer) {
    // ------------------------------------------------------------------ drainers (TMEM quadrant = lane group)
    bool ok = true;
    uint32_t par_f = 0, dbase = 0, npiece = 0;
    const int g4 = lane >> 2, e4 = lane & 3;
    while (ok && walk.next(pc)) {
      const uint32_t tile = npiece & 1u;
      ++npiece;
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + tile * 256u;
      const int ncol = (pc.nw + 15) & ~15;
      for (int o = quad; o < pc.rows && ok; o += 4) {
        if (warp == 4) TC3_TRACE(2, 0);
        ok = mbar_wait3(smem_u32(&S.gfull[tile * 4u + quad]), (par_f >> tile) & 1u, abort_flag, gerr, 15, (npiece << 12) | (uint32_t)o);
        par_f ^= 1u << tile;
        if (!ok) break;
        tc_fence_after();
        const uint32_t dsq = dbase + (uint32_t)o;
        const uint32_t ds = dsq % ND;
        if (dsq >= ND) ok = wait_released(S.fin_done, dsq + 1u - ND, abort_flag, gerr, 16, dsq);
        if (!ok) break;
        if (warp == 4) TC3_TRACE(2, 1);
        float* dp = reinterpret_cast<float*>(S.dump + ds * DUMP_BYTES) + (size_t)g4 * DUMP_PLANE * 4 + e4;
        int c0 = 0;
        for (; c0 + 32 <= ncol; c0 += 32) {
          uint32_t v[32];
          UBD_LDTM32(v, t0 + c0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (lane < UBD_NF) {
#pragma unroll
            for (int k = 0; k < 32; ++k) dp[(c0 + k) * 4] = __uint_as_float(v[k]);
          }
        }
        if (c0 < ncol) {
          uint32_t v[16];
          UBD_LDTM16(v, t0 + c0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (lane < UBD_NF) {
#pragma unroll
            for (int k = 0; k < 16; ++k) dp[(c0 + k) * 4] = __uint_as_float(v[k]);
          }
        }
        if (warp == 4) TC3_TRACE(2, 2);
        for (c0 = 0; c0 < ncol; c0 += 16) UBD_STTM16(t0 + c0, bias_bits);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&S.gempty[tile * 4u + quad]));
          mbar_arrive(smem_u32(&S.dfull[ds]));
        }
        if (warp == 4) { TC3_TRACE(2, 3); ++tr_n; }
      }
      dbase += (uint32_t)pc.rows;
    }
  } else {
    // ------------------------------------------------------------------ finishers (warps 2, 3, 8, 9)
    const int fw = warp < 4 ? warp - 2 : warp - 6;          // 0..3
    const int ft = fw * 32 + lane;                          // 0..127
    if (out_mode == 2) {
      for (int i = ft; i < UBD_NF * head.n_out; i += 128) S.headw[i] = head.hk[i];
      for (int i = ft; i < head.n_out; i += 128) S.headw[UBD_NF * head.n_out + i] = head.hb[i];
      asm volatile("bar.sync 2, 128;" ::: "memory");         // the finisher warps only
    }
    bool ok = true;
    uint32_t dbase = 0;
    while (ok && walk.next(pc)) {
      for (int o = 0; o < pc.rows && ok; ++o) {
        const uint32_t dsq = dbase + (uint32_t)o;
        const uint32_t ds = dsq % ND;
        if (warp == 2) TC3_TRACE(3, 0);
        ok = mbar_wait3(smem_u32(&S.dfull[ds]), (dsq / ND) & 1u, abort_flag, gerr, 17, dsq);
        if (!ok) break;
        if (warp == 2) TC3_TRACE(3, 1);
        const float4* dp = reinterpret_cast<const float4*>(S.dump + ds * DUMP_BYTES);
        const int y = pc.c + (pc.j0 + o) * d;
        if (out_mode == 2) {
          for (int px = ft; px < pc.nw; px += 128) {
            float a[UBD_NF];
#pragma unroll
            for (int g = 0; g < UBD_NG; ++g) {
              const float4 q = dp[g * DUMP_PLANE + px];
              a[4 * g] = fmaxf(q.x, 0.f); a[4 * g + 1] = fmaxf(q.y, 0.f); a[4 * g + 2] = fmaxf(q.z, 0.f); a[4 * g + 3] = fmaxf(q.w, 0.f);
            }
            const size_t p = ((size_t)pc.n * h + y) * w + pc.x0 + px;
            const float* hw = S.headw;
            float* lo = head.logits ? head.logits + p * head.n_out : nullptr;
            for (int oc = 0; oc < head.n_out; ++oc) {
              float acc = hw[UBD_NF * head.n_out + oc];
#pragma unroll
              for (int c = 0; c < UBD_NF; ++c) acc = fmaf(a[c], hw[c * head.n_out + oc], acc);
              if (lo) lo[oc] = acc;
              if (oc == 0 && head.mask) head.mask[p] = acc > head.thr ? 1 : 0;
            }
          }
        } else if (BF16 && out_mode == 0) {
          uint4* orow = out + (((size_t)pc.n * h + y) * 3) * wpo + out_pad + pc.x0;
#pragma unroll
          for (int g8 = 0; g8 < 3; ++g8)
            for (int px = ft; px < pc.nw; px += 128) {
              const float4 q0 = dp[(2 * g8) * DUMP_PLANE + px], q1 = dp[(2 * g8 + 1) * DUMP_PLANE + px];
              orow[(size_t)g8 * wpo + px] =
                  make_uint4(pack_bf16x2(fmaxf(q0.x, 0.f), fmaxf(q0.y, 0.f)), pack_bf16x2(fmaxf(q0.z, 0.f), fmaxf(q0.w, 0.f)),
                             pack_bf16x2(fmaxf(q1.x, 0.f), fmaxf(q1.y, 0.f)), pack_bf16x2(fmaxf(q1.z, 0.f), fmaxf(q1.w, 0.f)));
            }
        } else {
          uint4* orow = out + (((size_t)pc.n * h + y) * UBD_NG) * wpo + out_pad + pc.x0;
          const bool rnd = !BF16 && out_mode == 0;
#pragma unroll
          for (int g = 0; g < UBD_NG; ++g)
            for (int px = ft; px < pc.nw; px += 128) {
              float4 q = dp[g * DUMP_PLANE + px];
              q.x = fmaxf(q.x, 0.f); q.y = fmaxf(q.y, 0.f); q.z = fmaxf(q.z, 0.f); q.w = fmaxf(q.w, 0.f);
              if (rnd) { q.x = round_tf32(q.x); q.y = round_tf32(q.y); q.z = round_tf32(q.z); q.w = round_tf32(q.w); }
              orow[(size_t)g * wpo + px] = *reinterpret_cast<uint4*>(&q);
            }
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();                              // the row's shared-memory reads precede its release
          *reinterpret_cast<volatile uint32_t*>(&S.fin_done[fw]) = dsq + 1u;
        }
        if (warp == 2) { TC3_TRACE(3, 2); ++tr_n; }
      }
      dbase += (uint32_t)pc.rows;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// A images of one layer from its Keras HWIO kernel (3,3,24,24) in the flat parameter buffer.
// tf32: image m = dx*3 + kp; [2 K cores][7 groups][32 oc rows][4 ic], group gi holds kernel row
// {ky0, -, ky2, ky1}[gi & 3]; values rounded to tf32 (rna); then bias[32].
__global__ void build_aimg_tf32_kernel(const float* __restrict__ params, const int64_t* __restrict__ koff,
                                       const int64_t* __restrict__ boff, uint8_t* __restrict__ dst_all) {
  const int layer = blockIdx.x;
  const float* K = params + koff[layer];
  const float* B = params + boff[layer];
  float* dst = reinterpret_cast<float*>(dst_all + (size_t)layer * WB_BYTES_TF32);
  constexpr int PER_IMG = A_IMG_BYTES / 4, PER_CORE = A_KCORE_BYTES / 4;
  for (int i = threadIdx.x; i < W_BYTES_TF32 / 4; i += blockDim.x) {
    const int m = i / PER_IMG, rem = i % PER_IMG;
    const int kcore = rem / PER_CORE, r2 = rem % PER_CORE;
    const int gi = r2 / 128, oc = (r2 % 128) / 4, col = r2 % 4;
    const int blk = gi & 3;
    const int ky = blk == 0 ? 0 : (blk == 2 ? 2 : (blk == 3 ? 1 : -1));
    const int dx = m / 3, kp = m % 3;
    const int ic = kp * 8 + kcore * 4 + col;
    dst[i] = (ky >= 0 && oc < UBD_NF) ? round_tf32(K[((ky * 3 + dx) * UBD_NF + ic) * UBD_NF + oc]) : 0.f;
  }
  for (int i = threadIdx.x; i < 32; i += blockDim.x) dst[W_BYTES_TF32 / 4 + i] = i < UBD_NF ? B[i] : 0.f;
}

// bf16: images 0..2 = tap dx, ic 0..15; image 3 = K core 0: tap dx=-1, K core 1: tap dx=0, ic 16..23;
// image 4 = K core 0: tap dx=+1, ic 16..23, K core 1 zero.  [2 K cores][7 groups][32 oc rows][8 ic].
__global__ void build_aimg_bf16_kernel(const float* __restrict__ params, const int64_t* __restrict__ koff,
                                       const int64_t* __restrict__ boff, uint8_t* __restrict__ dst_all) {
  const int layer = blockIdx.x;
  const float* K = params + koff[layer];
  const float* B = params + boff[layer];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(dst_all + (size_t)layer * WB_BYTES_BF16);
  constexpr int PER_IMG = A_IMG_BYTES / 2, PER_CORE = A_KCORE_BYTES / 2;
  for (int i = threadIdx.x; i < W_BYTES_BF16 / 2; i += blockDim.x) {
    const int m = i / PER_IMG, rem = i % PER_IMG;
    const int kcore = rem / PER_CORE, r2 = rem % PER_CORE;
    const int gi = r2 / 256, oc = (r2 % 256) / 8, col = r2 % 8;
    const int blk = gi & 3;
    const int ky = blk == 0 ? 0 : (blk == 2 ? 2 : (blk == 3 ? 1 : -1));
    int dx = -1, ic = 0;
    if (m < 3) { dx = m; ic = kcore * 8 + col; }
    else if (m == 3) { dx = kcore; ic = 16 + col; }
    else if (kcore == 0) { dx = 2; ic = 16 + col; }
    const float v = (ky >= 0 && dx >= 0 && oc < UBD_NF) ? K[((ky * 3 + dx) * UBD_NF + ic) * UBD_NF + oc] : 0.f;
    dst[i] = __float2bfloat16_rn(v);
  }
  float* bias = reinterpret_cast<float*>(dst_all + (size_t)layer * WB_BYTES_BF16 + W_BYTES_BF16);
  for (int i = threadIdx.x; i < 32; i += blockDim.x) bias[i] = i < UBD_NF ? B[i] : 0.f;
}

}  // namespace tc3

static constexpr size_t kTc3Tf32 = (size_t)UBD_NLAYERS_DIL * tc3::WB_BYTES_TF32;
static constexpr size_t kTc3Bf16 = (size_t)UBD_NLAYERS_DIL * tc3::WB_BYTES_BF16;

static void tc3_setup_attributes() {
  cudaFuncSetAttribute(tc3::dilconv_rot_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc3::Smem<false>));
  cudaFuncSetAttribute(tc3::dilconv_rot_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc3::Smem<true>));
}

// Rebuilt together with the ubd_tc.cuh images (tc_prepare clears tc_weights_dirty; tc3 keeps its own flag).
static int tc3_prepare(ubd_handle h) {
  int rc = tc_prepare(h);                                  // error flag, offsets table
  if (rc) return rc;
  if (!h->tc3_weights.p) {
    UBD_CUDA(cudaMalloc(&h->tc3_weights.p, kTc3Tf32 + kTc3Bf16));
    h->tc3_weights.cap = kTc3Tf32 + kTc3Bf16;
    h->tc3_weights_dirty = true;
  }
  if (h->tc3_weights_dirty) {
    const int64_t* d_offs = reinterpret_cast<const int64_t*>((uint8_t*)h->tc_weights.p + kTcZeroOff + tc::ZERO_BYTES + 64);
    tc3::build_aimg_tf32_kernel<<<UBD_NLAYERS_DIL, 256, 0, h->stream>>>(h->d_params, d_offs, d_offs + 6, (uint8_t*)h->tc3_weights.p);
    tc3::build_aimg_bf16_kernel<<<UBD_NLAYERS_DIL, 256, 0, h->stream>>>(h->d_params, d_offs, d_offs + 6, (uint8_t*)h->tc3_weights.p + kTc3Tf32);
    h->launches += 2;
    UBD_CUDA(cudaGetLastError());
    h->tc3_weights_dirty = false;
  }
  return UBD_OK;
}

static int tc3_launch_dilconv(ubd_handle h, const void* in, void* out, int layer, int n, int hh, int ww, int d,
                              int out_mode, int out_pad = UBD_MAP_PAD, const tc::HeadArgs* head = nullptr) {
  if (h->precision != UBD_TF32 && h->precision != UBD_BF16) UBD_FAIL(UBD_ERR_UNSUPPORTED, "tensor-core path needs tf32 or bf16");
  int rc = tc3_prepare(h);
  if (rc) return rc;
  const bool bf16 = h->precision == UBD_BF16;
  const uint8_t* base = (const uint8_t*)h->tc3_weights.p;
  const uint8_t* wb = bf16 ? base + kTc3Tf32 + (size_t)layer * tc3::WB_BYTES_BF16 : base + (size_t)layer * tc3::WB_BYTES_TF32;
  const int sw = ww <= tc3::NW_MAX ? ww : tc3::NW_MAX;
  const int n_strips = (ww + sw - 1) / sw;
  const long long rows = (long long)n * n_strips * hh;
  const int grid = (int)std::min<long long>(rows, h->n_sm);
  tc::HeadArgs ha{};
  if (head) ha = *head;
  if (bf16)
    tc3::dilconv_rot_kernel<true><<<grid, tc3::THREADS, sizeof(tc3::Smem<true>), h->stream>>>(
        (const uint4*)in, (uint4*)out, wb, n, hh, ww, d, sw, out_mode, out_pad, tc_err_flag(h), ha, (long long*)h->tc_trace.p);
  else
    tc3::dilconv_rot_kernel<false><<<grid, tc3::THREADS, sizeof(tc3::Smem<false>), h->stream>>>(
        (const uint4*)in, (uint4*)out, wb, n, hh, ww, d, sw, out_mode, out_pad, tc_err_flag(h), ha, (long long*)h->tc_trace.p);
  ++h->launches;
  UBD_CUDA(cudaGetLastError());
  return UBD_OK;
}
