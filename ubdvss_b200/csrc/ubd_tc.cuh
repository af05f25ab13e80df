// First-generation tcgen05 implicit-GEMM kernel (one output row segment per accumulator), sm_100a only.
// It runs the stem's L3 (separable 24->24, stride 2, as a dense stride-2 conv, s2 != 0) and - behind option
// "tc_variant" 0 - the dilated 3x3 24->24 layers (net.py:298-304), whose default path is ubd_tc4.cuh.  This
// header also owns the shared tcgen05 / mbarrier helpers, the error flag and the merged stem kernels.
//
// GEMM view of one output row segment:  D[128 px, 32 oc] += A[128 px, K] * B[K, 32 oc],
//   K = 9 taps x 24 channels; kind::tf32 consumes K = 8 per instruction -> 27 MMAs per segment.
//
// Operand A is the input feature map itself.  Maps are row-interleaved planes of float4 with zero x
// padding (ubd_common.cuh), so one pixel of one plane is 16 bytes = one row of a UMMA core matrix
// (8 rows x 16 B) of the K-major SWIZZLE_NONE canonical layout ((8,m),2):((16 B, SBO),LBO).  A row
// slot in shared memory is an image of one map row: [6 planes][SW + 32 px][16 B]; 8 consecutive
// pixels form a core matrix (SBO = 128 B), the two K cores of one MMA are two adjacent channel
// planes (LBO = plane stride).  A (dy,dx) tap is then nothing but a different descriptor start
// address: slot(row + dy) + (16 + dx*d) * 16 B.  No im2col copy exists anywhere.
//
// Dilation is handled by phase decomposition in y: the rows r, r+d, r+2d, ... of an image form an
// independent 1-dilated problem, so a CTA walks a phase top to bottom with a ring of row slots:
// each input row is fetched once (+2 halo rows per item) and used by three output rows.  A row is
// staged by the TMA unit with cp.async.bulk: ONE copy when the strip spans the padded row (maps up
// to 256 px wide), else one per plane; rows above / below the image come from a zero page;
// completion is signalled on mbarriers.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over a static item list).  Role code is
// warp-uniform; only the asynchronous instruction itself sits under elect.sync, so descriptors live
// in uniform registers and an MMA costs a handful of issue slots:
//   warp 0   : producer  - bulk copies into the slot ring      (empty[] -> full[])
//   warps 1,6: MMA issue - even / odd segments: 27 tcgen05.mma per segment into a TMEM accumulator,
//              tcgen05.commit to tmem_full[].  Two issuers because the issue of one segment blocks on
//              the operand feed (~47 cycles / MMA) and each barrier wait costs ~230 cycles: while one
//              warp waits, the other keeps the tensor pipe fed (measured with the in-kernel trace).
//   warps 2-5: epilogue  - tcgen05.ld (lane = pixel, 24 columns = channels), bias + ReLU,
//              six coalesced float4 stores (one per plane); warp 2 also releases the staged rows: once
//              it has seen tmem_full of a row's last segment every MMA that read the top row of that
//              window has completed, whichever warp issued it.
#pragma once
#include <cuda_bf16.h>
#include "ubd_handle.cuh"

namespace tc {

constexpr int SEG = 128;                      // pixels per MMA segment = UMMA M
constexpr int MAX_SW = 256;                   // widest strip: two segments per staged row
constexpr int PAD = UBD_MAP_PAD;              // 16 = largest dilation
constexpr int MAX_PLANE_BYTES = (MAX_SW + 2 * PAD) * 16;        // 4608
constexpr int MAX_SLOT_BYTES = 30720;         // max(6 planes x 288 px, stride-2: 2 parities x 6 planes x 160 px) x 16 B
constexpr int NS = 6;                         // row slots in the ring
constexpr int NACC = 4;                       // TMEM accumulator stages
constexpr int UMMA_N = 32;                    // 24 output channels padded to a legal N for M=128
constexpr int TMEM_COLS = NACC * UMMA_N;      // 128
constexpr int N_MMA = 27;                     // 9 taps x 3 K-pairs (8 tf32 each)
constexpr int B_TILE_BYTES = UMMA_N * 8 * 4;  // 1024: [2 K cores][4 oc groups][8 rows][16 B]
constexpr int W_BYTES = N_MMA * B_TILE_BYTES; // 27648
constexpr int WB_BYTES = W_BYTES + 128;       // + bias[24] padded to 32 floats
// bf16 variant: maps are 3 planes of 8 channels per 16 B; kind::f16 consumes K = 16 (two planes) per
// MMA.  27 (tap, plane) K-cores pair up as: planes 0+1 of every tap (9 MMAs, LBO = plane stride),
// plane 2 of taps dx=-1 and dx=0 of one row (3 MMAs, LBO = d pixels), plane 2 of tap dx=+1 with a
// zero B core (3 MMAs): 15 MMAs per segment, each the same 5 KB operand read as a tf32 MMA.
constexpr int N_MMA_BF16 = 15;
constexpr int NG_BF16 = 3;
constexpr int W_BYTES_BF16 = N_MMA_BF16 * B_TILE_BYTES;   // 15360 (32 oc x 16 ic x 2 B per MMA)
constexpr int WB_BYTES_BF16 = W_BYTES_BF16 + 128;
constexpr int RQ = 8;                         // output rows per work item
constexpr int THREADS = 224;                  // producer, MMA issuer A, 4 epilogue warps, MMA issuer B
constexpr int ZERO_BYTES = MAX_SLOT_BYTES;

struct Smem {
  uint8_t slots[NS * MAX_SLOT_BYTES];
  uint8_t wimg[W_BYTES];
  float bias[32];
  uint64_t full[NS], empty[NS], tfull[NACC], tempty[NACC], wbar;
  uint32_t tmem_base;
  int abort_flag;
  float headw[UBD_NF * (1 + UBD_MAX_CLASSES) + 1 + UBD_MAX_CLASSES];   // out_mode 2: head kernel [24][n_out], bias
};
constexpr size_t SMEM_BYTES = sizeof(Smem) + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait, executed by every lane of the role's warp; the result is made warp-uniform with a
// vote so that everything downstream stays in the uniform datapath.  A wrong descriptor / byte count
// ends in an error code (host: tc_check_error), never in a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* gerr, int code) {
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
    if (*abort_flag) { ok = false; break; }
    if (clock64() - t0 > 1500000000LL) { atomicCAS(gerr, 0, code); *abort_flag = 1; ok = false; break; }
  }
  return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// lo word = start address >> 4 | (LBO >> 4) << 16, hi word = SBO >> 4 | version bit 14.
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);          // SBO = 128 B, version 1
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 32 (cute::UMMA::InstrDescriptor).
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((UMMA_N >> 3) << 17) | ((SEG >> 4) << 24);

// kind::f16 with bf16 operands, fp32 accumulate, K-major A and B, M = 128, N = 32.
constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((UMMA_N >> 3) << 17) | ((SEG >> 4) << 24);
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc = IDESC_BF16) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC_TF32), "r"(accumulate), "r"(0u) : "memory");
}

struct Item { int n, x0, r, q0, rows, nq, nseg, copy_px; };

struct Sched {
  int n_imgs, h, w, d, sw, n_strips, n_chunks, total;
  __device__ Sched(int n_imgs_, int h_, int w_, int d_, int sw_) : n_imgs(n_imgs_), h(h_), w(w_), d(d_), sw(sw_) {
    n_strips = (w + sw - 1) / sw;
    n_chunks = ((h + d - 1) / d + RQ - 1) / RQ;
    total = n_imgs * n_strips * d * n_chunks;
  }
  __device__ bool get(int idx, Item& it) const {
    const int c = idx % n_chunks; int t = idx / n_chunks;
    it.r = t % d; t /= d;
    it.x0 = (t % n_strips) * sw;
    it.n = t / n_strips;
    it.nq = it.r < h ? (h - it.r + d - 1) / d : 0;
    it.q0 = c * RQ;
    it.rows = min(RQ, it.nq - it.q0);
    const int valid = min(sw, w - it.x0);                 // output pixels of this strip
    it.nseg = (valid + SEG - 1) / SEG;
    it.copy_px = min(sw, w - it.x0) + 2 * PAD;            // padded pixels staged per plane
    return it.rows > 0;
  }
};

__device__ __forceinline__ float round_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// The 16-bit paths serve two containers: bf16 (precision UBD_BF16) and IEEE half (UBD_F16: the 10-bit significand of
// tf32 in 16 bits, so maps cost half the traffic and an MMA consumes K = 16).  Which one a launch uses travels in the
// weight blob: the last float of its 32-float bias block (channels 24..31 are padding) is non-zero for half.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));     // saturates at 65504: no inf * 0 in an MMA
  return r;
}
__device__ __forceinline__ uint32_t pack16(float lo, float hi, bool f16) { return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
// ReLU inside the conversion (cvt ... .relu): the 16-bit epilogues and L1 producers spend no FMNMX on it
__device__ __forceinline__ uint32_t pack16_relu(float lo, float hi, bool f16) {
  uint32_t r;
  if (f16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
constexpr int F16_FLAG_SLOT = 31;

// in/out: padded row-interleaved maps (pad = PAD) of n_imgs images, in 16-byte units: tf32 maps have
// 6 planes of float4, bf16 maps 3 planes of 8 x bf16.  wb: this layer's B image followed by bias[32];
// sw: strip width (128 or 256).  out_mode: 0 = same format as the input, tf32 output rounded (rna);
// 1 = fp32 6-plane output without rounding (last layer, feeds the fp32 head).
//
// out_mode 2: the 1x1 head (net.py:307-311) and the logit threshold (model_runner.py:124) run in the
// epilogue: the thread that owns a pixel holds its 24 channels, so the last map never reaches HBM.
struct HeadArgs { const float* hk; const float* hb; int n_out; float thr; float* logits; uint8_t* mask;
                  const uint4* gate; };   // out_mode 4 (ubd_tc4.cuh): map whose sign gates the output (ReLU derivative)
// Arguments of the L1-producer variant of the column-rotating kernel (ubd_tc4.cuh): grey uint8 image -> L1 rows.
struct L1Args { const float* lut; const float* dw1; const float* pw1; const float* b1; int H, W, pad_t, pad_l; };

template <bool BF16>
__global__ void __launch_bounds__(THREADS, 1)
dilconv_tc_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, const uint8_t* __restrict__ wb,
                  const uint8_t* __restrict__ zeros, int n_imgs, int h, int w, int d, int sw, int out_mode, int out_pad,
                  int* gerr, long long* trace, HeadArgs head, int s2) {
  // s2 != 0: stride-2 conv (the stem's L3 as a dense conv).  The input map (2h rows) is stored split by
  // column parity, [n][y][parity][plane][k + PAD] with E[k] = column 2k, O[k] = column 2k+1, so the tap
  // (ti,tj) of output pixel x reads row 2y - pad + ti of array (tj - pad) & 1 at offset floor((tj - pad)/2):
  // unit stride again, i.e. just another descriptor start address.  s2 = 1: pad 1 (FML), 2: pad 0.
  const int s2pad = s2 == 1 ? 1 : 0;
  const int rstep = s2 ? 2 : 1;                              // staged rows consumed per output row
  constexpr int NGI = BF16 ? NG_BF16 : UBD_NG;                  // planes of the input map
  const int ngs = s2 ? 2 * NGI : NGI;                        // planes per staged row
  constexpr uint32_t WBB = BF16 ? WB_BYTES_BF16 : WB_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // keep the shared address space (no integer casts)
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // provably warp-uniform
  const int lane = threadIdx.x & 31;
  volatile int* abort_flag = &S.abort_flag;
  // optional event trace of CTA 0 (bring-up / tuning): trace[role][event][4] cycle stamps
  const bool tr = trace != nullptr && blockIdx.x == 0 && lane == 0;
  int tr_n = 0;
#define TC_TRACE(role, slot, val) do { if (tr && tr_n < 1024) trace[((role) * 1024 + tr_n) * 4 + (slot)] = (val); } while (0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&S.full[i]), 1); mbar_init(smem_u32(&S.empty[i]), 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(smem_u32(&S.tfull[i]), 1); mbar_init(smem_u32(&S.tempty[i]), 4); }
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
  const Sched sched(n_imgs, h, w, d, sw);
  const uint32_t slots0 = smem_u32(S.slots);
  const int wp = w + 2 * PAD;                              // global row pitch of the input map in pixels
  const int wpo = w + 2 * out_pad;                         // ... of the output map
  const uint32_t plane_bytes = (uint32_t)(sw + 2 * PAD) * 16;   // slot plane stride = LBO
  const uint32_t slot_bytes = (uint32_t)ngs * plane_bytes;
  const bool one_copy = (w == sw) && !s2;                         // slot is an exact image of the global row block

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&S.wbar), WBB);
      bulk_g2s(smem_u32(S.wimg), wb, WBB, smem_u32(&S.wbar));
    }
    uint32_t lseq = 0;
    bool ok = true;
    for (int idx = blockIdx.x; idx < sched.total && ok; idx += gridDim.x) {
      Item it;
      if (!sched.get(idx, it)) continue;
      // stride-2 taps reach one pixel to either side only: stage 1 halo pixel instead of PAD (19 % less read traffic)
      const int halo = s2 ? 1 : PAD;
      const uint32_t copy_bytes = (uint32_t)(it.copy_px - 2 * (PAD - halo)) * 16;
      const uint32_t row_bytes = one_copy ? slot_bytes : (uint32_t)ngs * copy_bytes;
      const int nloads = s2 ? 2 * it.rows + 1 : it.rows + 2;
      for (int li = 0; li < nloads && ok; ++li, ++lseq) {
        const uint32_t slot = lseq % NS;
        TC_TRACE(0, 0, clock64());
        ok = mbar_wait(smem_u32(&S.empty[slot]), ((lseq / NS) & 1) ^ 1, abort_flag, gerr, 1);
        if (!ok) break;
        TC_TRACE(0, 1, clock64());
        const uint32_t bar = smem_u32(&S.full[slot]);
        const uint32_t dst = slots0 + slot * slot_bytes;
        // input row: phase row q of the dilated walk, or row 2*q0 - pad + li of the stride-2 input
        const int q = it.q0 - 1 + li;
        const int yin = s2 ? 2 * it.q0 - s2pad + li : it.r + q * d;
        const int hin = s2 ? 2 * h : h;
        const bool valid = s2 ? (yin >= 0 && yin < hin) : (q >= 0 && q < it.nq);
        // source of plane 0: padded pixel x0 of the row (the strip's left halo starts there)
        const uint4* src = valid ? in + (((size_t)it.n * hin + yin) * ngs) * wp + it.x0
                                 : reinterpret_cast<const uint4*>(zeros);
        if (elect_one()) {
          mbar_expect_tx(bar, row_bytes);
          if (one_copy) {
            bulk_g2s(dst, src, slot_bytes, bar);
          } else {
            for (int g = 0; g < ngs; ++g)
              bulk_g2s(dst + g * plane_bytes + (PAD - halo) * 16, valid ? src + (size_t)g * wp + (PAD - halo) : src, copy_bytes, bar);
          }
        }
        __syncwarp();
        TC_TRACE(0, 2, clock64());
        ++tr_n;
      }
    }
  } else if (warp == 1 || warp == 6) {
    // ------------------------------------------------------------------ MMA issuers (even / odd segments)
    const uint32_t parity = warp == 6 ? 1u : 0u;
    bool ok = mbar_wait(smem_u32(&S.wbar), 0, abort_flag, gerr, 2);
    // 16-bit operands: bf16 or IEEE half (flag in the weight blob), fp32 accumulate
    const bool f16 = BF16 && ok && reinterpret_cast<const float*>(S.wimg + W_BYTES_BF16)[F16_FLAG_SLOT] != 0.f;
    const uint32_t idesc16 = f16 ? (IDESC_BF16 & ~((7u << 7) | (7u << 10))) : IDESC_BF16;
    (void)idesc16;
    const uint32_t b_lo0 = ((smem_u32(S.wimg) >> 4) & 0x3FFFu) | ((512u >> 4) << 16);      // LBO = 512 B
    const uint32_t a_lbo = ((plane_bytes >> 4) & 0x3FFFu) << 16;
    // per-tap column offsets in 16-byte units: (PAD + dx*d) pixels, + K-pair (two planes)
    const uint32_t kp_units = (2 * plane_bytes) >> 4;
    // per-tap column offset inside a staged row, in pixels (= 16-byte units), relative to PAD:
    //   dilated:  tap dx -> dx * d
    //   stride-2: tap tj -> parity array (tj - pad) & 1 (plane offset) and pixel offset floor((tj - pad) / 2)
    uint32_t tap_off[3], pair_off, pair_lbo, single_off;
    {
      const uint32_t par_units = (uint32_t)NGI * (plane_bytes >> 4);
#pragma unroll
      for (int tj = 0; tj < 3; ++tj) {
        if (s2) {
          const int v = tj - s2pad;                                   // -1, 0, 1, 2
          tap_off[tj] = (uint32_t)((v & 1) ? (int)par_units : 0) + (uint32_t)((v - (v & 1)) / 2);
        } else {
          tap_off[tj] = (uint32_t)((tj - 1) * d);
        }
      }
      if (s2) { pair_off = tap_off[0]; pair_lbo = 1u; single_off = tap_off[1]; }      // taps tj = 0, 2 | tj = 1
      else { pair_off = (uint32_t)(-d); pair_lbo = (uint32_t)d; single_off = (uint32_t)d; }   // dx = -1, 0 | dx = +1
    }
    uint32_t lbase = 0, oseq = 0, waited = 0;      // waited = number of row loads known to have landed
    for (int idx = blockIdx.x; idx < sched.total && ok; idx += gridDim.x) {
      Item it;
      if (!sched.get(idx, it)) continue;
      for (int j = 0; j < it.rows && ok; ++j) {
        for (int s = 0; s < it.nseg && ok; ++s, ++oseq) {
          if ((oseq & 1u) != parity) continue;
          if (parity == 0) TC_TRACE(1, 0, clock64());
          while (waited < lbase + rstep * j + 3 && ok) {
            ok = mbar_wait(smem_u32(&S.full[waited % NS]), (waited / NS) & 1, abort_flag, gerr, 3);
            ++waited;
          }
          if (!ok) break;
          if (parity == 0) TC_TRACE(1, 1, clock64());
          const uint32_t acc = oseq % NACC;
          ok = mbar_wait(smem_u32(&S.tempty[acc]), ((oseq / NACC) & 1) ^ 1, abort_flag, gerr, 4);
          if (!ok) break;
          if (parity == 0) TC_TRACE(1, 2, clock64());
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * UMMA_N;
          uint32_t a_row[3];
#pragma unroll
          for (int t = 0; t < 3; ++t)
            a_row[t] = ((((slots0 + ((lbase + rstep * j + t) % NS) * slot_bytes) >> 4) & 0x3FFFu) | a_lbo) + PAD;
          if (elect_one()) {
            if constexpr (!BF16) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap % 3 - 1;
                const uint32_t a_lo = a_row[dy] + (uint32_t)(s * SEG) + tap_off[tap % 3];
#pragma unroll
                for (int kp = 0; kp < 3; ++kp) {
                  umma_tf32(tmem_d, make_desc(a_lo + kp * kp_units, DESC_HI),
                            make_desc(b_lo0 + (tap * 3 + kp) * (B_TILE_BYTES >> 4), DESC_HI), (tap | kp) != 0);
                }
              }
            } else {
              const uint32_t plane_units = plane_bytes >> 4;
              const uint32_t lbo_mask = ~(0x3FFFu << 16);
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {                 // planes 0 + 1 of every tap
                const int dy = tap / 3;
                umma_bf16(tmem_d, make_desc(a_row[dy] + (uint32_t)(s * SEG) + tap_off[tap % 3], DESC_HI),
                          make_desc(b_lo0 + tap * (B_TILE_BYTES >> 4), DESC_HI), tap != 0, idesc16);
              }
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                // plane 2 of the two taps that are `pair_lbo` pixels apart in the same array (K cores 0, 1):
                // dilated: dx = -1, 0 (LBO = d pixels); stride-2: tj = 0, 2 (same parity array, LBO = 1 pixel)
                const uint32_t a2 = (a_row[dy] & lbo_mask) + 2 * plane_units + (uint32_t)(s * SEG);
                umma_bf16(tmem_d, make_desc((a2 + pair_off) | (pair_lbo << 16), DESC_HI),
                          make_desc(b_lo0 + (9 + dy) * (B_TILE_BYTES >> 4), DESC_HI), 1u, idesc16);
                // plane 2 of the remaining tap; the B image's second K core is zero and LBO = 0 makes the A
                // side re-read the same (finite) core instead of whatever lies beyond the slot
                umma_bf16(tmem_d, make_desc(a2 + single_off, DESC_HI),
                          make_desc(b_lo0 + (12 + dy) * (B_TILE_BYTES >> 4), DESC_HI), 1u, idesc16);
              }
            }
            umma_commit(smem_u32(&S.tfull[acc]));
          }
          __syncwarp();
          if (parity == 0) { TC_TRACE(1, 3, clock64()); ++tr_n; }
        }
      }
      lbase += s2 ? 2 * it.rows + 1 : it.rows + 2;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int quad = warp & 3;                              // TMEM lane quadrant this warp may read
    bool ok = mbar_wait(smem_u32(&S.wbar), 0, abort_flag, gerr, 5);
    float bias[UBD_NF];
#pragma unroll
    for (int c = 0; c < UBD_NF; ++c)
      bias[c] = ok ? reinterpret_cast<const float*>(S.wimg + (BF16 ? W_BYTES_BF16 : W_BYTES))[c] : 0.f;
    const bool f16 = BF16 && ok && reinterpret_cast<const float*>(S.wimg + W_BYTES_BF16)[F16_FLAG_SLOT] != 0.f;
    (void)f16;
    uint32_t it_rows_plus2 = 0;
    if (out_mode == 2) {
      const int et = (int)threadIdx.x - 64;                  // 0..127 over the four epilogue warps
      for (int i = et; i < UBD_NF * head.n_out; i += 128) S.headw[i] = head.hk[i];
      for (int i = et; i < head.n_out; i += 128) S.headw[UBD_NF * head.n_out + i] = head.hb[i];
      asm volatile("bar.sync 2, 128;" ::: "memory");         // the epilogue warps only
    }
    uint32_t oseq = 0, lbase = 0;
    for (int idx = blockIdx.x; idx < sched.total && ok; idx += gridDim.x, lbase += it_rows_plus2) {
      Item it;
      it_rows_plus2 = 0;
      if (!sched.get(idx, it)) continue;
      it_rows_plus2 = s2 ? 2 * it.rows + 1 : it.rows + 2;
      for (int j = 0; j < it.rows && ok; ++j) {
        const int y = it.r + (it.q0 + j) * d;
        for (int s = 0; s < it.nseg && ok; ++s, ++oseq) {
          const uint32_t acc = oseq % NACC;
          if (warp == 2) TC_TRACE(2, 0, clock64());
          ok = mbar_wait(smem_u32(&S.tfull[acc]), (oseq / NACC) & 1, abort_flag, gerr, 6);
          if (!ok) break;
          if (warp == 2) TC_TRACE(2, 1, clock64());
          if (warp == 2 && lane == 0 && s == it.nseg - 1) {
            // every MMA of output rows <= j of this item has completed: the top row of the window
            // (and, after the item's last row, the two rows below it) can be overwritten
            for (int rr = 0; rr < rstep; ++rr) mbar_arrive(smem_u32(&S.empty[(lbase + rstep * j + rr) % NS]));
            if (j == it.rows - 1)
              for (int rr = rstep; rr < 3; ++rr) mbar_arrive(smem_u32(&S.empty[(lbase + rstep * j + rr) % NS]));
          }
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * UMMA_N;
          uint32_t v[24];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(taddr));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23])
                       : "r"(taddr + 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&S.tempty[acc]));
          const int x = it.x0 + s * SEG + quad * 32 + lane;
          if (x < w) {
            float o[UBD_NF];
#pragma unroll
            for (int c = 0; c < UBD_NF; ++c) o[c] = fmaxf(__uint_as_float(v[c]) + bias[c], 0.f);
            if (out_mode == 2) {
              const size_t p = ((size_t)it.n * h + y) * w + x;
              const float* hw = S.headw;
              float* lo = head.logits ? head.logits + p * head.n_out : nullptr;
              for (int oc = 0; oc < head.n_out; ++oc) {
                float a = hw[UBD_NF * head.n_out + oc];
#pragma unroll
                for (int c = 0; c < UBD_NF; ++c) a = fmaf(o[c], hw[c * head.n_out + oc], a);
                if (lo) lo[oc] = a;
                if (oc == 0 && head.mask) head.mask[p] = a > head.thr ? 1 : 0;
              }
            } else if (out_mode == 3) {
              // parity-split output in the precision's format (input of the stride-2 layer):
              // [n][y][x & 1][plane][PAD + (x >> 1)], pitch w/2 + 2*PAD
              constexpr int NGO = BF16 ? NG_BF16 : UBD_NG;
              const size_t wps = (size_t)(w / 2 + 2 * PAD);
              uint4* o_px = out + ((((size_t)it.n * h + y) * 2 + (x & 1)) * NGO) * wps + PAD + (x >> 1);
              if constexpr (BF16) {
#pragma unroll
                for (int g = 0; g < NG_BF16; ++g)
                  o_px[(size_t)g * wps] = make_uint4(pack16(o[8 * g], o[8 * g + 1], f16), pack16(o[8 * g + 2], o[8 * g + 3], f16),
                                                     pack16(o[8 * g + 4], o[8 * g + 5], f16), pack16(o[8 * g + 6], o[8 * g + 7], f16));
              } else {
#pragma unroll
                for (int g = 0; g < UBD_NG; ++g) {
                  float4 q = make_float4(round_tf32(o[4 * g]), round_tf32(o[4 * g + 1]), round_tf32(o[4 * g + 2]), round_tf32(o[4 * g + 3]));
                  o_px[(size_t)g * wps] = *reinterpret_cast<uint4*>(&q);
                }
              }
            } else if (BF16 && out_mode == 0) {
              uint4* o_px = out + (((size_t)it.n * h + y) * NG_BF16) * wpo + out_pad + x;
#pragma unroll
              for (int g = 0; g < NG_BF16; ++g)
                o_px[(size_t)g * wpo] = make_uint4(pack16(o[8 * g], o[8 * g + 1], f16), pack16(o[8 * g + 2], o[8 * g + 3], f16),
                                                  pack16(o[8 * g + 4], o[8 * g + 5], f16), pack16(o[8 * g + 6], o[8 * g + 7], f16));
            } else {
              uint4* o_px = out + (((size_t)it.n * h + y) * UBD_NG) * wpo + out_pad + x;
              const bool rnd = !BF16 && out_mode == 0;
#pragma unroll
              for (int g = 0; g < UBD_NG; ++g) {
                float4 q = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                if (rnd) { q.x = round_tf32(q.x); q.y = round_tf32(q.y); q.z = round_tf32(q.z); q.w = round_tf32(q.w); }
                o_px[(size_t)g * wpo] = *reinterpret_cast<uint4*>(&q);
              }
            }
          }
          if (warp == 2) { TC_TRACE(2, 2, clock64()); ++tr_n; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// Builds, for the 6 dilated layers, the UMMA B images from the Keras HWIO kernels in the flat
// parameter buffer: per (tap, K-pair) a [32 oc x 8 ic] K-major tile, values rounded to tf32 (rna),
// output channels 24..31 zero; followed by the bias.
__global__ void build_wimg_kernel(const float* __restrict__ params, const int64_t* __restrict__ koff,
                                  const int64_t* __restrict__ boff, uint8_t* __restrict__ wimg_all) {
  const int layer = blockIdx.x;
  const float* K = params + koff[layer];
  const float* B = params + boff[layer];
  float* dst = reinterpret_cast<float*>(wimg_all + (size_t)layer * WB_BYTES);
  for (int i = threadIdx.x; i < W_BYTES / 4; i += blockDim.x) {
    const int t = i / 256, rem = i % 256;              // t = tap*3 + kp
    const int kcore = rem / 128, ngroup = (rem % 128) / 32, row = (rem % 32) / 4, col = rem % 4;
    const int tap = t / 3, kp = t % 3;
    const int ic = kp * 8 + kcore * 4 + col, oc = ngroup * 8 + row;
    dst[i] = oc < UBD_NF ? round_tf32(K[(tap * UBD_NF + ic) * UBD_NF + oc]) : 0.f;
  }
  for (int i = threadIdx.x; i < 32; i += blockDim.x) dst[W_BYTES / 4 + i] = i < UBD_NF ? B[i] : 0.f;
}

// K[tap][c][o] = dw[tap][c] * pw[c][o] (+ bias copy): a separable layer as one dense 3x3 kernel.
__global__ void merge_sep_kernel(const float* __restrict__ dw, const float* __restrict__ pw, const float* __restrict__ b,
                                 float* __restrict__ dst) {
  for (int i = threadIdx.x; i < 9 * UBD_NF * UBD_NF; i += blockDim.x) {
    const int tap = i / (UBD_NF * UBD_NF), c = (i / UBD_NF) % UBD_NF, o = i % UBD_NF;
    dst[i] = dw[tap * UBD_NF + c] * pw[c * UBD_NF + o];
  }
  for (int i = threadIdx.x; i < 32; i += blockDim.x) dst[9 * UBD_NF * UBD_NF + i] = i < UBD_NF ? b[i] : 0.f;
}

// bf16 B images: per MMA [2 K cores][4 oc groups][8 oc rows][8 ic x bf16]; MMA list as in the kernel:
// 0..8 = tap t, ic 0..15; 9..11 = row dy: core0 = tap (dy,-1) ic 16..23, core1 = tap (dy,0) ic 16..23;
// 12..14 = row dy: core0 = tap (dy,+1) ic 16..23, core1 = 0.  Then the fp32 bias.
__global__ void build_wimg_bf16_kernel(const float* __restrict__ params, const int64_t* __restrict__ koff,
                                       const int64_t* __restrict__ boff, uint8_t* __restrict__ wimg_all, int s2_pairing, int f16) {
  const int layer = blockIdx.x;
  const float* K = params + koff[layer];
  const float* B = params + boff[layer];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(wimg_all + (size_t)layer * WB_BYTES_BF16);
  for (int i = threadIdx.x; i < W_BYTES_BF16 / 2; i += blockDim.x) {
    const int m = i / 512, rem = i % 512;                       // 512 bf16 per MMA image
    const int kcore = rem / 256, ngroup = (rem % 256) / 64, row = (rem % 64) / 8, col = rem % 8;
    const int oc = ngroup * 8 + row;
    int tap = -1, ic = 0;
    if (m < 9) { tap = m; ic = kcore * 8 + col; }
    else if (m < 12) { tap = (m - 9) * 3 + (s2_pairing ? 2 * kcore : kcore); ic = 16 + col; }   // cores 0,1 = dx -1,0 | tj 0,2
    else if (kcore == 0) { tap = (m - 12) * 3 + (s2_pairing ? 1 : 2); ic = 16 + col; }          // dx = +1 | tj = 1
    const float v = (tap >= 0 && oc < UBD_NF) ? K[(tap * UBD_NF + ic) * UBD_NF + oc] : 0.f;
    if (f16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v); else dst[i] = __float2bfloat16_rn(v);
  }
  float* bias = reinterpret_cast<float*>(wimg_all + (size_t)layer * WB_BYTES_BF16 + W_BYTES_BF16);
  for (int i = threadIdx.x; i < 32; i += blockDim.x) bias[i] = i < UBD_NF ? B[i] : (i == F16_FLAG_SLOT && f16 ? 1.f : 0.f);
}

}  // namespace tc

static void tc_setup_attributes() {
  cudaFuncSetAttribute(tc::dilconv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
  cudaFuncSetAttribute(tc::dilconv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
}

// tc_weights buffer: [6 x WB_BYTES tf32 images][6 x WB_BYTES_BF16 bf16 images][zero page][err flag][offsets]
static constexpr int kTcNumImg = UBD_NLAYERS_DIL + 2;            // + the stem's L2 and L3 as merged dense 3x3 kernels
static constexpr size_t kTcImgTf32 = (size_t)kTcNumImg * tc::WB_BYTES;
static constexpr size_t kTcImgBf16 = (size_t)kTcNumImg * tc::WB_BYTES_BF16;
static constexpr size_t kTcZeroOff = (kTcImgTf32 + kTcImgBf16 + 127) & ~(size_t)127;

#define ENSURE_RAW(buf, bytes)                                                     \
  do {                                                                             \
    if ((buf).cap < (bytes)) {                                                     \
      if ((buf).p) cudaFree((buf).p);                                              \
      UBD_CUDA(cudaMalloc(&(buf).p, (bytes)));                                     \
      (buf).cap = (bytes);                                                         \
    }                                                                              \
  } while (0)

static int tc_prepare(ubd_handle h) {
  const size_t total = kTcZeroOff + tc::ZERO_BYTES + 256;
  if (!h->tc_weights.p) {
    UBD_CUDA(cudaMalloc(&h->tc_weights.p, total));
    h->tc_weights.cap = total;
    UBD_CUDA(cudaMemsetAsync(h->tc_weights.p, 0, total, h->stream));
    h->tc_weights_dirty = true;
  }
  if (ubd_is16(h) && h->tc_w16 != h->precision) h->tc_weights_dirty = true;
  if (h->tc_weights_dirty) {
    int64_t offs[14];
    for (int l = 0; l < 6; ++l) { offs[l] = h->spec.off[9 + 2 * l]; offs[6 + l] = h->spec.off[10 + 2 * l]; }
    offs[12] = 0; offs[13] = 9 * UBD_NF * UBD_NF;           // merged L2 kernel / bias inside h->l2dense
    int64_t* d_offs = reinterpret_cast<int64_t*>((uint8_t*)h->tc_weights.p + kTcZeroOff + tc::ZERO_BYTES + 64);
    UBD_CUDA(cudaMemcpyAsync(d_offs, offs, sizeof(offs), cudaMemcpyHostToDevice, h->stream));
    UBD_CUDA(cudaStreamSynchronize(h->stream));      // offs is a stack array
    tc::build_wimg_kernel<<<6, 256, 0, h->stream>>>(h->d_params, d_offs, d_offs + 6, (uint8_t*)h->tc_weights.p);
    const int f16 = h->precision == UBD_F16;
    tc::build_wimg_bf16_kernel<<<6, 256, 0, h->stream>>>(h->d_params, d_offs, d_offs + 6, (uint8_t*)h->tc_weights.p + kTcImgTf32, 0, f16);
    // L2 (separable 24->24) as one dense 3x3 kernel: K[tap][c][o] = dw2[tap][c] * pw2[c][o], bias b2
    ENSURE_RAW(h->l2dense, 2 * (9 * UBD_NF * UBD_NF + 32) * sizeof(float));
    tc::merge_sep_kernel<<<1, 256, 0, h->stream>>>(h->d_params + h->spec.off[3], h->d_params + h->spec.off[4], h->d_params + h->spec.off[5],
                                                   (float*)h->l2dense.p);
    tc::build_wimg_kernel<<<1, 256, 0, h->stream>>>((const float*)h->l2dense.p, d_offs + 12, d_offs + 13,
                                                    (uint8_t*)h->tc_weights.p + (size_t)UBD_NLAYERS_DIL * tc::WB_BYTES);
    tc::build_wimg_bf16_kernel<<<1, 256, 0, h->stream>>>((const float*)h->l2dense.p, d_offs + 12, d_offs + 13,
                                                         (uint8_t*)h->tc_weights.p + kTcImgTf32 + (size_t)UBD_NLAYERS_DIL * tc::WB_BYTES_BF16, 0, f16);
    // L3 (separable 24->24, stride 2) likewise, image index 7
    float* l3 = (float*)h->l2dense.p + (9 * UBD_NF * UBD_NF + 32);
    tc::merge_sep_kernel<<<1, 256, 0, h->stream>>>(h->d_params + h->spec.off[6], h->d_params + h->spec.off[7], h->d_params + h->spec.off[8], l3);
    tc::build_wimg_kernel<<<1, 256, 0, h->stream>>>(l3, d_offs + 12, d_offs + 13,
                                                    (uint8_t*)h->tc_weights.p + (size_t)(UBD_NLAYERS_DIL + 1) * tc::WB_BYTES);
    tc::build_wimg_bf16_kernel<<<1, 256, 0, h->stream>>>(l3, d_offs + 12, d_offs + 13,
                                                         (uint8_t*)h->tc_weights.p + kTcImgTf32 + (size_t)(UBD_NLAYERS_DIL + 1) * tc::WB_BYTES_BF16, 1, f16);
    h->launches += 8;
    UBD_CUDA(cudaGetLastError());
    h->tc_weights_dirty = false;
    h->tc_w16 = ubd_is16(h) ? h->precision : h->tc_w16;
  }
  return UBD_OK;
}

static inline int* tc_err_flag(ubd_handle h) {
  return reinterpret_cast<int*>((uint8_t*)h->tc_weights.p + kTcZeroOff + tc::ZERO_BYTES);
}

// in / out: padded maps in the precision's layout (tf32: 6 float4 planes, bf16: 3 planes of 8 bf16).
// out_mode 1 writes fp32 6-plane output.  layer 0..5 = conv2d_1..6 (option "tc_variant" 0); layer 7 = the stem's
// L3 as a merged dense stride-2 kernel (s2 = 1: FML padding, 2: none).
static int tc_launch_dilconv(ubd_handle h, const void* in, void* out, int layer, int n, int hh, int ww, int d,
                             int out_mode, int out_pad = UBD_MAP_PAD, const tc::HeadArgs* head = nullptr, int s2 = 0) {
  if (h->precision == UBD_FP32) UBD_FAIL(UBD_ERR_UNSUPPORTED, "tensor-core path needs tf32, bf16 or f16");
  int rc = tc_prepare(h);
  if (rc) return rc;
  const bool bf16 = ubd_is16(h);
  const uint8_t* base = (const uint8_t*)h->tc_weights.p;
  const uint8_t* wb = bf16 ? base + kTcImgTf32 + (size_t)layer * tc::WB_BYTES_BF16 : base + (size_t)layer * tc::WB_BYTES;
  const uint8_t* zeros = base + kTcZeroOff;
  const int sw = (ww <= tc::SEG || s2) ? tc::SEG : tc::MAX_SW;      // stride-2 rows stage both parities: 128-px strips
  const int n_strips = (ww + sw - 1) / sw;
  const int n_chunks = ((hh + d - 1) / d + tc::RQ - 1) / tc::RQ;
  const long long items = (long long)n * n_strips * d * n_chunks;
  const int grid = (int)std::min<long long>(items, h->n_sm);
  tc::HeadArgs ha{};
  if (head) ha = *head;
  if (bf16)
    tc::dilconv_tc_kernel<true><<<grid, tc::THREADS, tc::SMEM_BYTES, h->stream>>>((const uint4*)in, (uint4*)out, wb, zeros, n, hh, ww, d, sw,
                                                                                 out_mode, out_pad, tc_err_flag(h), (long long*)h->tc_trace.p, ha, s2);
  else
    tc::dilconv_tc_kernel<false><<<grid, tc::THREADS, tc::SMEM_BYTES, h->stream>>>((const uint4*)in, (uint4*)out, wb, zeros, n, hh, ww, d, sw,
                                                                                  out_mode, out_pad, tc_err_flag(h), (long long*)h->tc_trace.p, ha, s2);
  ++h->launches;
  UBD_CUDA(cudaGetLastError());
  return UBD_OK;
}

// Reads (and clears) the device-side barrier-timeout flag; call after a stream synchronize.  `on` = the stream the
// read is queued on: the handle's own, or the read-back stream when the next batch is already queued on the former.
static int tc_check_error_on(ubd_handle h, cudaStream_t on) {
  if (!h->tc_weights.p) return UBD_OK;
  int codes[16] = {0};
  UBD_CUDA(cudaMemcpyAsync(codes, tc_err_flag(h), sizeof(codes), cudaMemcpyDeviceToHost, on));
  UBD_CUDA(cudaStreamSynchronize(on));
  if (codes[0]) {
    cudaMemsetAsync(tc_err_flag(h), 0, sizeof(codes), on);
    std::string where;
    for (int i = 1; i < 16; ++i)
      if (codes[i]) where += " w" + std::to_string(i - 1) + ":" + std::to_string(codes[i] >> 24) + "/" + std::to_string(codes[i] & 0xFFFFFF);
    UBD_FAIL(UBD_ERR_CUDA, "tcgen05 pipeline barrier timed out (role code " + std::to_string(codes[0]) + ";" + where + ")");
  }
  return UBD_OK;
}
static int tc_check_error(ubd_handle h) { return tc_check_error_on(h, h->stream); }
