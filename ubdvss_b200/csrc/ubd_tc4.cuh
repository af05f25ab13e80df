// Dilated 3x3 24->24 layers (net.py:298-304) on tcgen05, "column-rotating" formulation, sm_100a only.
//
// Pixels stay on the M side as in ubd_tc.cuh (D[128 px, N], lane = pixel, so the epilogue owns all 24
// channels of its pixel and all four TMEM quadrants drain in parallel), but the three kernel rows share
// one instruction through the N side:
//
//   D[128 px, 96 = (ky2 | ky1 | ky0) x 32 oc]  +=  A[128 px, K] (one staged map row)  *  B[K, 96] (weights)
//
// An input row r of a y-phase (rows c, c+d, c+2d, ...) contributes to the output rows r-1, r, r+1 through
// the kernel rows ky = 2, 1, 0.  A segment's accumulator tile has four 32-column groups; output row o owns
// group o mod 4, so input row i writes the groups (i-2, i-1, i) mod 4 while the fourth group - the output
// row completed by input row i-1 - is being read out.  Every staged row is read from shared memory by 10
// MMAs (bf16: 6) instead of 27 (15): 8 KB of operands per 128x128xK MMA instead of 5 KB per 128x32xK.
//
// The accumulators are never cleared or pre-loaded: the first MMA that contributes to an output row (kernel
// row ky0, first K step) is issued on that row's group alone with accumulate = 0; everything else of an
// interior input row is ONE instruction per K step over all four groups (N = 128) with a weight image
// rotated by its start address, [ky2 | ky1 | ky0 | 0] landing on groups (i-2, i-1, i, i+1) mod 4 - the group
// being read out only sees "+= 0".  Rows at the ends of a CTA's range use per-group N = 32 instructions.
//
// Warp roles (384 threads, 1 CTA/SM, each CTA owns a contiguous range of the global row sequence):
//   warp 0      producer : cp.async.bulk of map rows into the slot ring           (empty[] -> full[])
//   warps 1, 2  MMA issue: segment 0 / 1 (128 px each) of every staged row; each owns its accumulator tile
//   warps 4-7   epilogue of segment 0, warps 8-11 of segment 1: tcgen05.ld of the finished group
//               (lane = pixel), bias + ReLU, rounding / packing or the fused 1x1 head + threshold,
//               coalesced 16-byte stores
//   warps 12-20 (L1SRC variant, the stem's L2 layer as a dense conv): compute layer L1 (separable s2 1->24,
//               FP32, exact) of every staged row straight into the slot ring instead of the TMA producer
#pragma once
#include "ubd_tc.cuh"
#ifndef UBD_TC_TRACE
#define UBD_TC_TRACE 0
#endif
#ifndef UBD_L1_WARPS
#define UBD_L1_WARPS 9
#endif

namespace tc4 {

using tc::smem_u32; using tc::elect_one; using tc::mbar_init; using tc::mbar_arrive; using tc::mbar_expect_tx;
using tc::mbar_wait; using tc::bulk_g2s; using tc::umma_commit; using tc::tc_fence_before; using tc::tc_fence_after;
using tc::make_desc; using tc::round_tf32; using tc::pack_bf16x2; using tc::pack16; using tc::pack16_relu; using tc::HeadArgs;

// Bounded wait like tc::mbar_wait; on a stall every warp leaves (code << 24 | info) in gerr[1 + warp].
__device__ __forceinline__ bool mbar_wait3(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* gerr, int code, uint32_t info) {
  uint32_t polls = 0;                                        // bounded by poll count: no clock reads on the fast path
  bool ok = true;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (*abort_flag || ++polls > 20000000u) {                // a failed try_wait takes >= ~50 cycles: ~1 s
      atomicCAS(gerr, 0, code);
      *abort_flag = 1;
      gerr[1 + ((threadIdx.x >> 5) & 15)] = (code << 24) | (int)(info & 0xFFFFFFu);
      ok = false;
      break;
    }
  }
  return __all_sync(0xffffffffu, ok);
}

// Rounding (rna) to the tf32 grid of a finite, non-negative value (everything after a ReLU) that is only ever
// read by a kind::tf32 MMA: the tensor core ignores the low 13 mantissa bits
// (tests/test_gpu_tc.py::test_tf32_operands_are_truncated), so adding half a tf32 ulp is the whole rounding -
// one integer instruction where cvt.rna.tf32.f32 is emulated with four on sm_100a.
__device__ __forceinline__ float rna_mma(float v) { return __uint_as_float(__float_as_uint(v) + 0x1000u); }
// ReLU and that rounding in ONE instruction (VIADDMNMX): max(bits + half ulp, 0) as signed integers - a negative float is a
// negative integer and becomes +0.0, a positive one gets its half ulp.  One issue slot instead of FMNMX + IADD for every
// channel of every pixel in the L1 producers and the epilogues (the busiest scheduler of the stem kernel paces its ring).
__device__ __forceinline__ uint32_t relu_rna_bits(float pre) { return (uint32_t)__viaddmax_s32(__float_as_int(pre), 0x1000, 0); }

// One non-blocking probe of a barrier phase (the fast path of the waits on the MMA warps' critical path).
__device__ __forceinline__ uint32_t mbar_probe(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}

// packed fp32x2 FMA (one issue slot for two lanes of work on sm_100)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

constexpr int PAD = UBD_MAP_PAD;
constexpr int SEG = 128;
constexpr int SW_MAX = 256;                               // strip = two segments
constexpr int GROUP_BYTES = 32 * 16;                      // one group of a weight K-core: 32 oc rows x 16 B
constexpr int KCORE_BYTES = 7 * GROUP_BYTES;              // groups hold kernel rows {ky0, -, ky2, ky1, ky0, -, ky2}
constexpr int IMG_BYTES = 2 * KCORE_BYTES;                // 7168 per MMA: [2 K cores][7 groups][32 oc][16 B]
constexpr int N_IMG_TF32 = 10, N_IMG_BF16 = 6;            // 9 (5) K steps + the first K step without ky0
constexpr int W_BYTES_TF32 = N_IMG_TF32 * IMG_BYTES;      // 71680
constexpr int W_BYTES_BF16 = N_IMG_BF16 * IMG_BYTES;      // 43008
constexpr int WB_BYTES_TF32 = W_BYTES_TF32 + 128;         // + bias[32]
constexpr int WB_BYTES_BF16 = W_BYTES_BF16 + 128;
constexpr int SLOT_BYTES_TF32 = UBD_NG * (SW_MAX + 2 * PAD) * 16;      // 27648
constexpr int SLOT_BYTES_BF16 = 3 * (SW_MAX + 2 * PAD) * 16;           // 13824
constexpr int NS_TF32 = 5, NS_BF16 = 8;
constexpr int THREADS = 384;
constexpr int L1_WARPS = UBD_L1_WARPS;                    // L1-producer warps; each owns L1_PXW consecutive staged pixels
constexpr int L1_PXW = (SW_MAX + 2 + L1_WARPS - 1) / L1_WARPS;   // (sw + 2 = 258 staged pixels per row)
constexpr int L1_THREADS = 32 * L1_WARPS;
constexpr int THREADS_L1 = THREADS + L1_THREADS;
constexpr int TMEM_COLS = 256;                            // per segment one tile of 4 groups x 32 columns

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
constexpr int HEAD_STRIDE = 36;                           // 1 + UBD_MAX_CLASSES = 33 outputs, padded to whole float4s
// Head kernel of the launch in the constant bank: [24][HEAD_STRIDE] (columns >= n_out zero), then the bias row.  Every
// lane of a warp needs the same weight, and from shared memory a uniform 16-byte load still costs four LSU passes (the
// class head was bound by them: ncu, 207 LDS.128 per warp and row); as constant operands the weights ride inside the FMAs.
// One bank per device: the launcher copies the handle's padded head (3.6 KB, device to device) in front of every head
// launch on its stream; two handles running class-head launches concurrently on one device are not supported.
__constant__ float c_headw[(UBD_NF + 1) * HEAD_STRIDE];
// L1 (separable 1 -> 24, grey input) of the stem launch: pw1[24] then b1[24].  Uniform across a warp, so they ride in the
// FMAs as constant operands instead of twelve 16-byte shared-memory loads per pixel and row (the stem kernel keeps the
// shared-memory pipe ~2/3 busy with MMA operand reads, table look-ups and the staging stores: ncu mio_throttle).
__constant__ float c_l1w[2 * UBD_NF];

// Class head of one pixel: acc[4 g4 .. 4 g4 + 3] = bias + sum_c a[c] * hk[c][4 g4 ..] for N4 groups of four outputs.  N4 is a
// template parameter so that the loops unroll into FMAs with immediate constant-bank operands.
template <int N4>
__device__ __forceinline__ void head_fma(const float (&a)[UBD_NF], float (&acc)[HEAD_STRIDE]) {
#pragma unroll
  for (int o = 0; o < 4 * N4; ++o) acc[o] = c_headw[UBD_NF * HEAD_STRIDE + o];
#pragma unroll
  for (int c = 0; c < UBD_NF; ++c)
#pragma unroll
    for (int o = 0; o < 4 * N4; ++o) acc[o] = fmaf(a[c], c_headw[c * HEAD_STRIDE + o], acc[o]);
}

__global__ void build_headw_kernel(const float* __restrict__ hk, const float* __restrict__ hb, int n_out, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < (UBD_NF + 1) * HEAD_STRIDE; i += blockDim.x) {
    const int c = i / HEAD_STRIDE, oc = i % HEAD_STRIDE;
    dst[i] = oc < n_out ? (c < UBD_NF ? hk[c * n_out + oc] : hb[oc]) : 0.f;
  }
}

template <bool BF16> struct Smem {
  static constexpr int NS = BF16 ? NS_BF16 : NS_TF32;
  static constexpr int SLOT = BF16 ? SLOT_BYTES_BF16 : SLOT_BYTES_TF32;
  static constexpr int WB = BF16 ? WB_BYTES_BF16 : WB_BYTES_TF32;
  uint8_t slots[NS * SLOT];
  uint8_t wimg[WB];                                       // weight images, then bias[32]
  __align__(16) float hstage[BF16 ? 8 * 32 * (1 + UBD_MAX_CLASSES) : 4];   // class head: per epilogue warp 32 px x n_out logits (coalesced write-out)
  uint64_t full[NS], empty[NS], gfull[8], gempty[8], wbar;
  uint32_t tmem_base;
  int abort_flag;
  float lut[260];                // L1-producer variant: uint8 -> preprocessed float; entry 256 = 0 (a tap outside the image)
  __align__(16) float l1w[12 + UBD_NF + UBD_NF]; // dw1[9] (+3 pad), pw1[24], b1[24] (grey input)
};

// Layer-pipelined launch (PIPE): the six dilated layers run CONCURRENTLY on disjoint groups of CTAs (CTA b works on layer
// b % 6; its group splits every image's rows among its members), chained image by image through ring buffers of
// `ring_imgs` maps per layer boundary that are meant to stay in L2, and every CTA keeps ONE layer's weight images for the
// whole launch.  done[layer * n_imgs + image] counts the epilogue warps of that layer's CTAs that have finished the image
// (release: __threadfence + atomicAdd; acquire: ld.acquire.gpu by the producer warp, then a proxy fence before the
// bulk copies read the map).  A layer starts image m when the layer before it has finished m and the layer after it
// has finished m - ring_imgs (its ring slot is free).
struct PipeArgs { uint4* ring; int* done; int ring_imgs; long long img_units; };

__device__ __forceinline__ bool pipe_wait(const int* flag, int target, volatile int* abort_flag, int* gerr, int code) {
  uint32_t polls = 0;
  while (true) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= target) return true;
    if (*abort_flag || ++polls > 8000000u) { atomicCAS(gerr, 0, code); *abort_flag = 1; return false; }
    __nanosleep(32);
  }
}

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

__device__ __forceinline__ void umma(bool bf16, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (bf16)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// in / out: padded row-interleaved maps (pad = PAD) in 16-byte units (tf32: 6 planes of float4, bf16: 3
// planes of 8 x bf16).  wb: this layer's weight images followed by bias[32].  sw: strip width.
// out_mode 0: same format as the input (tf32 rna / bf16), 1: fp32 6-plane unrounded, 2: fused 1x1 head +
// logit threshold (net.py:307-311, model_runner.py:124): the last map never reaches HBM.
// out_mode 3: parity-split output [n][y][x & 1][plane][PAD + (x >> 1)] (input of the stride-2 layer).
// out_mode 5: as 0 (tf32 only), with the rounded-off mantissa bits cleared (training forward).
// out_mode 4: no bias / ReLU; output gated by the sign of head.gate (backward-data pass of the training step, tf32).
// L1SRC: `in` is the uint8 grey image, h / w the half-resolution map size, d = 1 (see tc::L1Args).
template <bool BF16, bool L1SRC, bool PIPE = false>
__global__ void __launch_bounds__(L1SRC ? THREADS_L1 : THREADS, 1)
dilconv_col_kernel(const uint4* __restrict__ in_, uint4* __restrict__ out_, const uint8_t* __restrict__ wb_,
                   int n_imgs, int h, int w, int d_, int sw, int out_mode_, int out_pad, int* gerr, HeadArgs head, long long* trace,
                   tc::L1Args l1, PipeArgs pipe) {
  using S_t = Smem<BF16>;
  // PIPE: this CTA's layer, its index inside the layer's group and the group size; else one layer for the whole grid
  int layer = 0, kidx = (int)blockIdx.x, K = (int)gridDim.x;
  const uint4* in = in_;
  uint4* out = out_;
  const uint8_t* wb = wb_;
  int d = d_, out_mode = L1SRC ? 3 : out_mode_;            // the L1-producer variant only ever writes the parity-split map (dead epilogues compile out)
  if constexpr (PIPE) {
    layer = (int)blockIdx.x % UBD_NLAYERS_DIL;
    kidx = (int)blockIdx.x / UBD_NLAYERS_DIL;
    K = ((int)gridDim.x - layer + UBD_NLAYERS_DIL - 1) / UBD_NLAYERS_DIL;
    d = layer == 1 ? 2 : (layer == 2 ? 4 : (layer == 3 ? 8 : (layer == 4 ? 16 : 1)));
    wb = wb_ + (size_t)layer * S_t::WB;
    out_mode = layer == UBD_NLAYERS_DIL - 1 ? 2 : 0;
    if (layer > 0) in = pipe.ring + (size_t)(layer - 1) * pipe.ring_imgs * pipe.img_units;
    out = pipe.ring + (size_t)layer * pipe.ring_imgs * pipe.img_units;
  }
  // 16-bit maps: bf16 or IEEE half (flag in the weight blob, tc::F16_FLAG_SLOT)
  const bool f16 = BF16 && __ldg(reinterpret_cast<const float*>(wb + (BF16 ? W_BYTES_BF16 : W_BYTES_TF32)) + tc::F16_FLAG_SLOT) != 0.f;
  (void)f16;
  const int n_outer = PIPE ? n_imgs : 1;                     // PIPE: the roles walk image by image
  auto make_walk = [&](int m) { return PIPE ? Walk(1, h, w, d, sw, (kidx + m) % K, K) : Walk(n_imgs, h, w, d, sw, kidx, K); };
  constexpr int NS = S_t::NS;
  constexpr int NGI = BF16 ? 3 : UBD_NG;
  constexpr uint32_t WBB = S_t::WB;
  constexpr uint32_t WBYTES = BF16 ? W_BYTES_BF16 : W_BYTES_TF32;
  constexpr int N_STEP = BF16 ? 5 : 9;                      // K steps per input row; image N_STEP = step 0 without ky0
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  S_t& S = *reinterpret_cast<S_t*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  volatile int* abort_flag = &S.abort_flag;
  // optional event trace of CTA 0 (tuning): trace[role][event][4] cycle stamps
  // (compiled in with -DUBD_TC_TRACE=1 only: the stamps cost clock reads and issue slots in every role)
#if UBD_TC_TRACE
  const bool tr = trace != nullptr && blockIdx.x == 0 && lane == 0;
  int tr_n = 0;
#define TC4_TRACE(role, slot) do { if (tr && tr_n < 1024) trace[(((L1SRC ? 4 : 0) + (role)) * 1024 + tr_n) * 4 + (slot)] = clock64(); } while (0)
#define TC4_TRACE_NEXT() (++tr_n)
#else
  (void)trace;
#define TC4_TRACE(role, slot) do { } while (0)
#define TC4_TRACE_NEXT() do { } while (0)
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&S.full[i]), L1SRC ? L1_THREADS / 32 : 1); mbar_init(smem_u32(&S.empty[i]), 2); }
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&S.gfull[i]), 1); mbar_init(smem_u32(&S.gempty[i]), 4); }
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
  const bool epi = warp >= 4 && warp < 12;
  const int quad = warp & 3;
  const int seg = epi ? (warp - 4) >> 2 : (warp == 2 ? 1 : 0);       // segment column this warp works on
  float bias[UBD_NF];
  if (epi) {
#pragma unroll
    for (int c = 0; c < UBD_NF; ++c) bias[c] = __ldg(reinterpret_cast<const float*>(wb + WBYTES) + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t slots0 = smem_u32(S.slots);
  if (!L1SRC && w == sw) {
    // single-strip maps: the pad columns of the staged rows are never copied (12 % of the read traffic), so
    // their pad columns are cleared once; the MMAs read them through the async proxy
    for (int i = threadIdx.x; i < NS * NGI * 2 * PAD; i += blockDim.x) {      // the 2 x 16 pad pixels of every plane of every slot
      const int pp = i % (2 * PAD), pl = (i / (2 * PAD)) % NGI, sl = i / (2 * PAD * NGI);
      const int px = pp < PAD ? pp : w + pp;
      *reinterpret_cast<uint4*>(S.slots + (size_t)sl * S_t::SLOT + ((size_t)pl * (sw + 2 * PAD) + px) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  const int wp = w + 2 * PAD;
  const int wpo = w + 2 * out_pad;
  const uint32_t plane_bytes = (uint32_t)(sw + 2 * PAD) * 16;
  const uint32_t slot_bytes = (uint32_t)NGI * plane_bytes;
  const bool one_copy = (w == sw);
  Piece pc;
  uint64_t* gfull = S.gfull + seg * 4;
  uint64_t* gempty = S.gempty + seg * 4;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&S.wbar), WBB);
      bulk_g2s(smem_u32(S.wimg), wb, WBB, smem_u32(&S.wbar));
    }
    uint32_t lseq = 0;
    bool ok = !L1SRC;                                      // L1SRC: the rows come from the L1 warps below
    for (int m = 0; m < n_outer && ok; ++m) {
    Walk walk = make_walk(m);
    if constexpr (PIPE) {
      // the layer before has finished this image, and the layer after has finished the image whose ring slot we reuse
      bool okf = true;
      if (lane == 0) {
        const int g = (int)gridDim.x;
        if (layer > 0) okf = pipe_wait(pipe.done + (size_t)(layer - 1) * n_imgs + m, 8 * ((g - (layer - 1) + UBD_NLAYERS_DIL - 1) / UBD_NLAYERS_DIL), abort_flag, gerr, 27);
        if (okf && layer < UBD_NLAYERS_DIL - 1 && m >= pipe.ring_imgs)
          okf = pipe_wait(pipe.done + (size_t)(layer + 1) * n_imgs + m - pipe.ring_imgs, 8 * ((g - (layer + 1) + UBD_NLAYERS_DIL - 1) / UBD_NLAYERS_DIL), abort_flag, gerr, 28);
        asm volatile("fence.proxy.async;" ::: "memory");   // other SMs' generic-proxy stores -> this SM's bulk copies
      }
      ok = __all_sync(0xffffffffu, okf);
    }
    while (ok && walk.next(pc)) {
      if constexpr (PIPE) pc.n = m;
      const int in_n = PIPE ? (layer == 0 ? pc.n : pc.n % pipe.ring_imgs) : pc.n;
      const uint32_t copy_bytes = (uint32_t)(pc.nw + 2 * PAD) * 16;
      for (int i = 0; i < pc.rows + 2 && ok; ++i) {
        const int jj = pc.j0 - 1 + i;
        if (jj < 0 || jj >= pc.R) continue;                 // zero row above / below the image: no MMAs at all
        const uint32_t slot = lseq % NS;
        TC4_TRACE(0, 0);
        ok = mbar_wait3(smem_u32(&S.empty[slot]), ((lseq / NS) & 1) ^ 1, abort_flag, gerr, 21, lseq);
        if (!ok) break;
        TC4_TRACE(0, 1);
        const uint32_t bar = smem_u32(&S.full[slot]);
        const uint32_t dst = slots0 + slot * S_t::SLOT;
        const int y = pc.c + jj * d;
        const uint4* src = in + (((size_t)in_n * h + y) * NGI) * wp + pc.x0;
        if (elect_one()) {
          if (one_copy) {
            // the strip spans the image: its x padding is zero (cleared once above), copy the interior only
            mbar_expect_tx(bar, (uint32_t)NGI * (uint32_t)w * 16u);
            for (int g = 0; g < NGI; ++g) bulk_g2s(dst + g * plane_bytes + PAD * 16, src + (size_t)g * wp + PAD, (uint32_t)w * 16u, bar);
          } else {
            mbar_expect_tx(bar, (uint32_t)NGI * copy_bytes);
            for (int g = 0; g < NGI; ++g) bulk_g2s(dst + g * plane_bytes, src + (size_t)g * wp, copy_bytes, bar);
          }
        }
        __syncwarp();
        TC4_TRACE(0, 2);
        TC4_TRACE_NEXT();
        ++lseq;
      }
    }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers (segment 0 / 1)
    bool ok = mbar_wait(smem_u32(&S.wbar), 0, abort_flag, gerr, 22);
    // weight images: K-major, LBO = K-core stride, SBO = 128 B; the run [ky2 | ky1 | ky0] starts at image group 2
    const uint32_t b_lo0 = ((smem_u32(S.wimg) >> 4) & 0x3FFFu) | ((uint32_t)(KCORE_BYTES >> 4) << 16);
    const uint32_t plane_units = plane_bytes >> 4;
    const uint32_t a_lbo = (plane_units & 0x3FFFu) << 16;
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
    const uint32_t idesc0 = BF16 ? ((1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((128u >> 4) << 24))
                                 : ((1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24));
    const uint32_t tmem_t = tmem_base + (uint32_t)seg * 128u;
    uint32_t lseq = 0, par_e = 0;
    for (int m = 0; m < n_outer && ok; ++m) {
    Walk walk = make_walk(m);
    while (ok && walk.next(pc)) {
      const bool active = seg * SEG < pc.nw;                 // a narrow strip has no second segment
      for (int i = 0; i < pc.rows + 2 && ok; ++i) {
        if (warp == 1) TC4_TRACE(1, 0);
        // (a) output row i gets its first contribution now: its column group must have been read out;
        // (b) the staged input row must have landed.  Both barriers are probed together first: in steady
        // state they have completed long ago and the bounded (clocked) wait is never entered.
        const bool need_g = active && i < pc.rows;
        const uint32_t bit = (uint32_t)(i & 3);
        const int jj = pc.j0 - 1 + i;
        const bool valid = jj >= 0 && jj < pc.R;
        const uint32_t slot = lseq % NS;
        const uint32_t g_par = ((par_e >> bit) & 1u) ^ 1u, f_par = (lseq / NS) & 1u;
        uint32_t d1 = 1u, d2 = 1u;
        if (need_g) d1 = mbar_probe(smem_u32(&gempty[bit]), g_par);
        if (valid) d2 = mbar_probe(smem_u32(&S.full[slot]), f_par);
        if (!__all_sync(0xffffffffu, (d1 & d2) != 0u)) {
          if (need_g) ok = mbar_wait3(smem_u32(&gempty[bit]), g_par, abort_flag, gerr, 23, (uint32_t)i);
          if (ok && valid) ok = mbar_wait3(smem_u32(&S.full[slot]), f_par, abort_flag, gerr, 24, (uint32_t)i);
          if (!ok) break;
        }
        if (need_g) par_e ^= 1u << bit;
        if (valid) ++lseq;
        tc_fence_after();
        if (warp == 1) TC4_TRACE(1, 1);
        if (warp == 1) TC4_TRACE(1, 2);
        // kernel rows t = 0, 1, 2 <-> (ky2, ky1, ky0) <-> output rows (i-2, i-1, i) <-> image groups (2, 3, 4) <->
        // column groups (i-2, i-1, i) mod 4.  The first valid input row of an output row overwrites its group.
        const uint32_t a_row = ((((slots0 + slot * S_t::SLOT) >> 4) & 0x3FFFu) | a_lbo) + PAD + (uint32_t)(seg * SEG);
        const bool interior = i >= 2 && i < pc.rows;        // all three output rows exist (and were started earlier)
        auto a_desc = [&](int m) -> uint64_t {                // A operand (pixels) of K step m
          if constexpr (!BF16) {
            const int dx = m / 3, kp = m % 3;
            return make_desc(a_row + (uint32_t)((dx - 1) * d) + (uint32_t)kp * 2u * plane_units, DESC_HI);
          } else {
            const uint32_t a2 = (a_row & ~(0x3FFFu << 16)) + 2u * plane_units;
            if (m < 3) return make_desc(a_row + (uint32_t)((m - 1) * d), DESC_HI);             // planes 0 + 1 of tap dx = m
            if (m == 3) return make_desc((a2 + (uint32_t)(-d)) | ((uint32_t)d << 16), DESC_HI);  // plane 2 of dx = -1, 0 (LBO = d px)
            return make_desc(a2 + (uint32_t)d, DESC_HI);                                         // plane 2 of dx = +1 (LBO = 0)
          }
        };
        if (elect_one()) {
          if (valid && active) {
            if (interior) {
              // image groups s..s+3 -> column groups 0..3 with ky2 on group (i-2) mod 4
              const uint32_t b_rot = b_lo0 + (uint32_t)((4 - (i & 3)) & 3) * (GROUP_BYTES >> 4);
              umma(BF16, tmem_t + (uint32_t)(i & 3) * 32u, a_desc(0), make_desc(b_lo0, DESC_HI), idesc0 | ((32u >> 3) << 17), 0u);
              umma(BF16, tmem_t, a_desc(0), make_desc(b_rot + (uint32_t)N_STEP * (IMG_BYTES >> 4), DESC_HI), idesc0 | ((128u >> 3) << 17), 1u);
#pragma unroll
              for (int m = 1; m < N_STEP; ++m)
                umma(BF16, tmem_t, a_desc(m), make_desc(b_rot + (uint32_t)m * (IMG_BYTES >> 4), DESC_HI), idesc0 | ((128u >> 3) << 17), 1u);
            } else {
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                const int o = i - 2 + t;
                if (o >= 0 && o < pc.rows) {
                  const int first = (pc.j0 == 0 && o == 0) ? 1 : o;      // first input row of output o that exists
                  const uint32_t tmem_d = tmem_t + (uint32_t)(o & 3) * 32u;
                  const uint32_t b_g = b_lo0 + (uint32_t)(2 + t) * (GROUP_BYTES >> 4);
#pragma unroll
                  for (int m = 0; m < N_STEP; ++m)
                    umma(BF16, tmem_d, a_desc(m), make_desc(b_g + (uint32_t)m * (IMG_BYTES >> 4), DESC_HI), idesc0 | ((32u >> 3) << 17),
                         (m == 0 && i == first) ? 0u : 1u);
                }
              }
            }
          }
          if (valid) {
            if (active) umma_commit(smem_u32(&S.empty[slot])); else mbar_arrive(smem_u32(&S.empty[slot]));
          }
          if (active && i >= 2) umma_commit(smem_u32(&gfull[(i - 2) & 3]));
        }
        __syncwarp();
        if (warp == 1) { TC4_TRACE(1, 3); TC4_TRACE_NEXT(); }
      }
    }
    }
  } else if (epi) {
    // ------------------------------------------------------------------ epilogue (lane = pixel, 24 columns = channels)
    const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)seg * 128u;
    bool ok = true;
    uint32_t par_f = 0;
    for (int m = 0; m < n_outer && ok; ++m) {
    Walk walk = make_walk(m);
    while (ok && walk.next(pc)) {
      if constexpr (PIPE) pc.n = m;
      const int out_n = PIPE ? pc.n % pipe.ring_imgs : pc.n;   // ring slot of the output map (the head writes by image index)
      if (seg * SEG >= pc.nw) continue;
      for (int o = 0; o < pc.rows && ok; ++o) {
        const uint32_t G = (uint32_t)(o & 3);
        if (warp == 4) TC4_TRACE(2, 0);
        ok = mbar_wait3(smem_u32(&gfull[G]), (par_f >> G) & 1u, abort_flag, gerr, 25, (uint32_t)o);
        par_f ^= 1u << G;
        if (!ok) break;
        tc_fence_after();
        if (warp == 4) TC4_TRACE(2, 1);
        const uint32_t taddr = tq + G * 32u;
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
        if (lane == 0) mbar_arrive(smem_u32(&gempty[G]));
        if (warp == 4) TC4_TRACE(2, 2);
        const int y = pc.c + (pc.j0 + o) * d;
        const int xs = seg * SEG + quad * 32 + lane;           // pixel inside the strip
        if (xs < pc.nw) {
          const int x = pc.x0 + xs;
          float a[UBD_NF];
#pragma unroll
          for (int c = 0; c < UBD_NF; ++c) a[c] = out_mode == 4 ? __uint_as_float(v[c]) : __uint_as_float(v[c]) + bias[c];     // pre-activation
          if (out_mode == 4) {
            // backward-data pass (training): the conv ran with the flipped kernel; gate by the ReLU of the layer below
            // (its saved activation, same layout as the output) and leave the gradient on the tf32 grid for the next MMAs
            if constexpr (!BF16) {
              const size_t o_idx = (((size_t)out_n * h + y) * UBD_NG) * wpo + out_pad + x;
#pragma unroll
              for (int g = 0; g < UBD_NG; ++g) {
                const uint4 gt = __ldg(head.gate + o_idx + (size_t)g * wpo);
                float4 q = make_float4(__uint_as_float(gt.x) > 0.f ? rna_mma(a[4 * g]) : 0.f, __uint_as_float(gt.y) > 0.f ? rna_mma(a[4 * g + 1]) : 0.f,
                                       __uint_as_float(gt.z) > 0.f ? rna_mma(a[4 * g + 2]) : 0.f, __uint_as_float(gt.w) > 0.f ? rna_mma(a[4 * g + 3]) : 0.f);
                uint4 u = *reinterpret_cast<uint4*>(&q);
                u.x &= 0xFFFFE000u; u.y &= 0xFFFFE000u; u.z &= 0xFFFFE000u; u.w &= 0xFFFFE000u;
                out[o_idx + (size_t)g * wpo] = u;
              }
            }
          } else if (out_mode == 2) {
#pragma unroll
            for (int c = 0; c < UBD_NF; ++c) a[c] = fmaxf(a[c], 0.f);
            const size_t p = ((size_t)pc.n * h + y) * w + x;
            if (head.n_out == 1) {
              float acc = c_headw[UBD_NF * HEAD_STRIDE];
#pragma unroll
              for (int c = 0; c < UBD_NF; ++c) acc = fmaf(a[c], c_headw[c * HEAD_STRIDE], acc);
              if (head.logits) head.logits[p] = acc;
              if (head.mask) head.mask[p] = acc > head.thr ? 1 : 0;
            } else {
              // class head (net.py:307-311, up to 33 outputs)
              const int n4 = (head.n_out + 3) >> 2;
              float acc[HEAD_STRIDE];
#pragma unroll
              for (int i = 0; i < HEAD_STRIDE; ++i) acc[i] = 0.f;
              switch (n4) {
                case 1: head_fma<1>(a, acc); break;
                case 2: head_fma<2>(a, acc); break;
                case 3: head_fma<3>(a, acc); break;
                case 4: head_fma<4>(a, acc); break;
                case 5: head_fma<5>(a, acc); break;
                case 6: head_fma<6>(a, acc); break;
                case 7: head_fma<7>(a, acc); break;
                case 8: head_fma<8>(a, acc); break;
                default: head_fma<9>(a, acc); break;
              }
              if (head.mask) head.mask[p] = acc[0] > head.thr ? 1 : 0;
              if (head.logits) {
                float* lo = BF16 ? S.hstage + ((size_t)(warp - 4) * 32 + lane) * head.n_out : head.logits + p * head.n_out;
#pragma unroll
                for (int oc = 0; oc < HEAD_STRIDE; ++oc)
                  if (oc < head.n_out) lo[oc] = acc[oc];
              }
            }
          } else if (out_mode == 3) {
            constexpr int NGO = BF16 ? 3 : UBD_NG;
            const size_t wps = (size_t)(w / 2 + 2 * PAD);
            uint4* o_px = out + ((((size_t)pc.n * h + y) * 2 + (x & 1)) * NGO) * wps + PAD + (x >> 1);
            if constexpr (BF16) {
#pragma unroll
              for (int g = 0; g < 3; ++g)
                o_px[(size_t)g * wps] = make_uint4(pack16_relu(a[8 * g], a[8 * g + 1], f16), pack16_relu(a[8 * g + 2], a[8 * g + 3], f16),
                                                   pack16_relu(a[8 * g + 4], a[8 * g + 5], f16), pack16_relu(a[8 * g + 6], a[8 * g + 7], f16));
            } else {
#pragma unroll
              for (int g = 0; g < UBD_NG; ++g)
                o_px[(size_t)g * wps] = make_uint4(relu_rna_bits(a[4 * g]), relu_rna_bits(a[4 * g + 1]), relu_rna_bits(a[4 * g + 2]), relu_rna_bits(a[4 * g + 3]));
            }
          } else if (BF16 && out_mode == 0) {
            // (streaming stores: the next layer reads this map after the whole sweep, long after L2 has turned over)
            uint4* o_px = out + (((size_t)out_n * h + y) * 3) * wpo + out_pad + x;
#pragma unroll
            for (int g = 0; g < 3; ++g)
              __stcs(o_px + (size_t)g * wpo, make_uint4(pack16_relu(a[8 * g], a[8 * g + 1], f16), pack16_relu(a[8 * g + 2], a[8 * g + 3], f16),
                                                        pack16_relu(a[8 * g + 4], a[8 * g + 5], f16), pack16_relu(a[8 * g + 6], a[8 * g + 7], f16)));
          } else {
            uint4* o_px = out + (((size_t)out_n * h + y) * UBD_NG) * wpo + out_pad + x;
            const bool rnd = !BF16 && (out_mode == 0 || out_mode == 5);
#pragma unroll
            for (int g = 0; g < UBD_NG; ++g) {
              uint4 u;
              if (rnd) u = make_uint4(relu_rna_bits(a[4 * g]), relu_rna_bits(a[4 * g + 1]), relu_rna_bits(a[4 * g + 2]), relu_rna_bits(a[4 * g + 3]));
              else u = make_uint4(__float_as_uint(fmaxf(a[4 * g], 0.f)), __float_as_uint(fmaxf(a[4 * g + 1], 0.f)),
                                  __float_as_uint(fmaxf(a[4 * g + 2], 0.f)), __float_as_uint(fmaxf(a[4 * g + 3], 0.f)));
              if (out_mode == 5) {
                // training forward: the map is also read by FP32 code (ReLU gates test a > 0, the head, the weight gradient),
                // so the low bits are really cleared - a clamped 0 must stay 0, not half a tf32 ulp
                u.x &= 0xFFFFE000u; u.y &= 0xFFFFE000u; u.z &= 0xFFFFE000u; u.w &= 0xFFFFE000u;
              }
              __stcs(o_px + (size_t)g * wpo, u);
            }
          }
        }
        if constexpr (BF16) {
          if (out_mode == 2 && head.n_out > 1 && head.logits) {
            // the warp's 32 (or fewer, multiple of 4) consecutive pixels x n_out logits are contiguous in the NHWC output:
            // write them as coalesced float4s instead of n_out strided 4-byte stores per pixel
            __syncwarp();
            const int xs0 = seg * SEG + quad * 32;
            const int cnt = min(32, pc.nw - xs0);
            if (cnt > 0) {
              const float4* src4 = reinterpret_cast<const float4*>(S.hstage + (size_t)(warp - 4) * 32 * head.n_out);
              float4* dst4 = reinterpret_cast<float4*>(head.logits + (((size_t)pc.n * h + y) * w + pc.x0 + xs0) * head.n_out);
              const int n16 = cnt * head.n_out / 4;
              for (int i = lane; i < n16; i += 32) dst4[i] = src4[i];
            }
            __syncwarp();
          }
        }
        if (warp == 4) { TC4_TRACE(2, 3); TC4_TRACE_NEXT(); }
      }
    }
    if constexpr (PIPE) {
      // this warp's part of image m is stored: publish it (release) to the next layer's producers and the previous layer's
      __threadfence();
      __syncwarp();
      if (ok && lane == 0) atomicAdd(pipe.done + (size_t)layer * n_imgs + m, 1);
    }
    }
  }

  if (L1SRC && warp >= 12) {
    // ------------------------------------------------------------------ L1 producers (9 warps)
    // One thread = one staged pixel j of every row (map column x0 - 1 + j; only x0-1 .. x0+nw can be read
    // by the d = 1 taps).  The image bytes of the NEXT row are loaded before waiting for its slot, so the
    // global latency overlaps the wait and the previous row's arithmetic.  The nine warps run independently
    // (each waits for the slot and publishes its 32 pixels with one arrival), so a slow warp does not hold
    // the others back; they are only bounded by the slot ring.
    const int tl = (int)threadIdx.x - THREADS;              // 0 .. L1_THREADS-1 (table loads)
    const int t = (warp - 12) * L1_PXW + lane;              // staged pixel of this thread
    const uint8_t* img = reinterpret_cast<const uint8_t*>(in);
    for (int i = tl; i < 260; i += L1_THREADS) S.lut[i] = i < 256 ? (l1.lut ? l1.lut[i] : (float)i) : 0.f;
    for (int i = tl; i < 12 + 2 * UBD_NF; i += L1_THREADS)
      S.l1w[i] = i < 9 ? l1.dw1[i] : (i < 12 ? 0.f : (i < 12 + UBD_NF ? l1.pw1[i - 12] : l1.b1[i - 12 - UBD_NF]));
    asm volatile("bar.sync 1, %0;" ::"n"(L1_THREADS) : "memory");   // the L1 warps only
    float dwr[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) dwr[i] = S.l1w[i];
    uint32_t lseq = 0, ring_slot = 0, ring_phase = 0;        // ring position = (lseq % NS, (lseq / NS) & 1)
    bool ok = true;
    Walk walk = make_walk(0);
    while (ok && walk.next(pc)) {
      const int x = pc.x0 - 1 + t;                           // this thread's map column
      const bool use = lane < L1_PXW && t < pc.nw + 2;
      const bool okx = use && x >= 0 && x < w;
      // image columns 2x - pad_l + {0,1,2}: validity mask and CLAMPED column offsets.  Every lane executes the same nine
      // loads (a tap outside the image reads a clamped address and is replaced by table index 256, whose entry is 0), so
      // the warps that own an image edge or idle lanes no longer serialise an edge path next to the interior one - the
      // in-kernel trace showed the first warp of a strip 4 rows behind the others, pacing the whole ring.
      uint32_t cm = 0u;
      int coff[3];
#pragma unroll
      for (int tj = 0; tj < 3; ++tj) {
        const int ix = 2 * x - l1.pad_l + tj;
        if (okx && ix >= 0 && ix < l1.W) cm |= 1u << tj;
        coff[tj] = min(max(ix, 0), l1.W - 1);
      }
      const uint8_t* pimg = img + ((size_t)pc.n * l1.H) * l1.W;
      const size_t W1 = (size_t)l1.W;
      // the 9 table indices of map pixel (yy, x), one register each: nothing consumes them before the next row's
      // arithmetic, so the loads really stay in flight
      auto load9 = [&](int yy, uint32_t (&r)[9]) {
        const int iy0 = 2 * yy - l1.pad_t;
        if (iy0 >= 0 && iy0 + 2 < l1.H) {                    // (warp-uniform) all three image rows exist
          const uint8_t* p = pimg + (size_t)iy0 * W1;
#pragma unroll
          for (int ti = 0; ti < 3; ++ti)
#pragma unroll
            for (int tj = 0; tj < 3; ++tj) {
              const uint32_t v = __ldg(p + ti * W1 + coff[tj]);
              r[ti * 3 + tj] = (cm & (1u << tj)) ? v : 256u;
            }
        } else {
#pragma unroll
          for (int ti = 0; ti < 3; ++ti) {
            const int iy = iy0 + ti;
            const bool rv = iy >= 0 && iy < l1.H;
            const uint8_t* p = pimg + (size_t)min(max(iy, 0), l1.H - 1) * W1;
#pragma unroll
            for (int tj = 0; tj < 3; ++tj) {
              const uint32_t v = __ldg(p + coff[tj]);
              r[ti * 3 + tj] = (rv && (cm & (1u << tj))) ? v : 256u;
            }
          }
        }
      };
      int i = pc.j0 == 0 ? 1 : 0;                             // input rows that exist: jj = j0 - 1 + i in [0, R)
      const int i_end = min(pc.rows + 1, pc.R - pc.j0);
      uint32_t b[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) b[q] = 256u;
      if (i <= i_end) load9(pc.c + (pc.j0 - 1 + i) * d, b);
      for (; i <= i_end && ok; ++i, ++lseq) {
        const uint32_t slot = ring_slot, ring_par = ring_phase ^ 1u;
        if (++ring_slot == NS) { ring_slot = 0; ring_phase ^= 1u; }
        if (warp == 12) TC4_TRACE(3, 0);
        uint32_t nb[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) nb[q] = 256u;
        if (i + 1 <= i_end) load9(pc.c + (pc.j0 + i) * d, nb);
        ok = mbar_wait3(smem_u32(&S.empty[slot]), ring_par, abort_flag, gerr, 26, lseq);
        if (warp == 12) TC4_TRACE(3, 1);
        if (ok && use) {
          uint8_t* px = S.slots + (size_t)slot * S_t::SLOT + (size_t)(PAD - 1 + t) * 16;
          if (okx) {
            // depthwise 3x3 (stride 2) on the preprocessed bytes, pointwise 1 -> 24, bias, ReLU
            float a = 0.f;
#pragma unroll
            for (int q = 0; q < 9; ++q) a = fmaf(S.lut[b[q]], dwr[q], a);
            // pointwise 1 -> 24, bias, ReLU: packed fp32x2 FMAs on 16-byte weight loads (pw1 at l1w[12..36), b1 at l1w[36..60))
            float o[UBD_NF];                                   // pre-activation
#pragma unroll
            for (int c = 0; c < UBD_NF; ++c) o[c] = fmaf(a, c_l1w[c], c_l1w[UBD_NF + c]);
            if constexpr (BF16) {
#pragma unroll
              for (int g = 0; g < 3; ++g)
                *reinterpret_cast<uint4*>(px + g * plane_bytes) =
                    make_uint4(pack16_relu(o[8 * g], o[8 * g + 1], f16), pack16_relu(o[8 * g + 2], o[8 * g + 3], f16),
                               pack16_relu(o[8 * g + 4], o[8 * g + 5], f16), pack16_relu(o[8 * g + 6], o[8 * g + 7], f16));
            } else {
#pragma unroll
              for (int g = 0; g < UBD_NG; ++g)
                *reinterpret_cast<uint4*>(px + g * plane_bytes) =
                    make_uint4(relu_rna_bits(o[4 * g]), relu_rna_bits(o[4 * g + 1]), relu_rna_bits(o[4 * g + 2]), relu_rna_bits(o[4 * g + 3]));
            }
          } else {
            // column outside the map: L2's zero padding
#pragma unroll
            for (int g = 0; g < (BF16 ? 3 : UBD_NG); ++g) *reinterpret_cast<uint4*>(px + g * plane_bytes) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        if (warp == 12) TC4_TRACE(3, 2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (ok && lane == 0) mbar_arrive(smem_u32(&S.full[slot]));
        if (warp == 12) { TC4_TRACE(3, 3); TC4_TRACE_NEXT(); }
#pragma unroll
        for (int q = 0; q < 9; ++q) b[q] = nb[q];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}


// Weight images of one layer from its Keras HWIO kernel (3,3,24,24) in the flat parameter buffer.
// tf32: image m = dx*3 + kp (K step); [2 K cores][7 groups][32 oc rows][4 ic], group gi holds kernel row
// {ky0, -, ky2, ky1}[gi & 3]; values rounded to tf32 (rna).  Image 9 = image 0 without ky0.  Then bias[32].
__global__ void build_img_tf32_kernel(const float* __restrict__ params, const int64_t* __restrict__ koff,
                                      const int64_t* __restrict__ boff, uint8_t* __restrict__ dst_all) {
  const int layer = blockIdx.x;
  const float* K = params + koff[layer];
  const float* B = params + boff[layer];
  float* dst = reinterpret_cast<float*>(dst_all + (size_t)layer * WB_BYTES_TF32);
  constexpr int PER_IMG = IMG_BYTES / 4, PER_CORE = KCORE_BYTES / 4;
  for (int i = threadIdx.x; i < W_BYTES_TF32 / 4; i += blockDim.x) {
    const int mi = i / PER_IMG, rem = i % PER_IMG;
    const int m = mi == 9 ? 0 : mi;
    const int kcore = rem / PER_CORE, r2 = rem % PER_CORE;
    const int gi = r2 / 128, oc = (r2 % 128) / 4, col = r2 % 4;
    const int blk = gi & 3;
    int ky = blk == 0 ? 0 : (blk == 2 ? 2 : (blk == 3 ? 1 : -1));
    if (mi == 9 && ky == 0) ky = -1;
    const int dx = m / 3, kp = m % 3;
    const int ic = kp * 8 + kcore * 4 + col;
    dst[i] = (ky >= 0 && oc < UBD_NF) ? tc::round_tf32(K[((ky * 3 + dx) * UBD_NF + ic) * UBD_NF + oc]) : 0.f;
  }
  for (int i = threadIdx.x; i < 32; i += blockDim.x) dst[W_BYTES_TF32 / 4 + i] = i < UBD_NF ? B[i] : 0.f;
}

// bf16: images 0..2 = tap dx, ic 0..15; image 3 = K core 0: tap dx=-1, K core 1: tap dx=0, ic 16..23;
// image 4 = K core 0: tap dx=+1, ic 16..23, K core 1 zero; image 5 = image 0 without ky0.
__global__ void build_img_bf16_kernel(const float* __restrict__ params, const int64_t* __restrict__ koff,
                                      const int64_t* __restrict__ boff, uint8_t* __restrict__ dst_all, int f16) {
  const int layer = blockIdx.x;
  const float* K = params + koff[layer];
  const float* B = params + boff[layer];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(dst_all + (size_t)layer * WB_BYTES_BF16);
  constexpr int PER_IMG = IMG_BYTES / 2, PER_CORE = KCORE_BYTES / 2;
  for (int i = threadIdx.x; i < W_BYTES_BF16 / 2; i += blockDim.x) {
    const int mi = i / PER_IMG, rem = i % PER_IMG;
    const int m = mi == 5 ? 0 : mi;
    const int kcore = rem / PER_CORE, r2 = rem % PER_CORE;
    const int gi = r2 / 256, oc = (r2 % 256) / 8, col = r2 % 8;
    const int blk = gi & 3;
    int ky = blk == 0 ? 0 : (blk == 2 ? 2 : (blk == 3 ? 1 : -1));
    if (mi == 5 && ky == 0) ky = -1;
    int dx = -1, ic = 0;
    if (m < 3) { dx = m; ic = kcore * 8 + col; }
    else if (m == 3) { dx = kcore; ic = 16 + col; }
    else if (kcore == 0) { dx = 2; ic = 16 + col; }
    const float v = (ky >= 0 && dx >= 0 && oc < UBD_NF) ? K[((ky * 3 + dx) * UBD_NF + ic) * UBD_NF + oc] : 0.f;
    if (f16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v); else dst[i] = __float2bfloat16_rn(v);
  }
  float* bias = reinterpret_cast<float*>(dst_all + (size_t)layer * WB_BYTES_BF16 + W_BYTES_BF16);
  for (int i = threadIdx.x; i < 32; i += blockDim.x) bias[i] = i < UBD_NF ? B[i] : (i == tc::F16_FLAG_SLOT && f16 ? 1.f : 0.f);
}

}  // namespace tc4

static constexpr int kTc4NumImg = UBD_NLAYERS_DIL + 1;          // + the stem's L2 as a merged dense 3x3 kernel
static constexpr size_t kTc4Tf32 = (size_t)kTc4NumImg * tc4::WB_BYTES_TF32;
static constexpr size_t kTc4Bf16 = (size_t)kTc4NumImg * tc4::WB_BYTES_BF16;

static void tc4_setup_attributes() {
  cudaFuncSetAttribute(tc4::dilconv_col_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc4::Smem<false>));
  cudaFuncSetAttribute(tc4::dilconv_col_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc4::Smem<true>));
  cudaFuncSetAttribute(tc4::dilconv_col_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc4::Smem<false>));
  cudaFuncSetAttribute(tc4::dilconv_col_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc4::Smem<true>));
  cudaFuncSetAttribute(tc4::dilconv_col_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc4::Smem<false>));
  cudaFuncSetAttribute(tc4::dilconv_col_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc4::Smem<true>));
}

// Weight images, rebuilt whenever the parameters change (tc_prepare owns the error flag and the offsets table).
static int tc4_prepare(ubd_handle h) {
  int rc = tc_prepare(h);
  if (rc) return rc;
  if (!h->tc4_weights.p) {
    UBD_CUDA(cudaMalloc(&h->tc4_weights.p, kTc4Tf32 + kTc4Bf16));
    h->tc4_weights.cap = kTc4Tf32 + kTc4Bf16;
    h->tc4_weights_dirty = true;
  }
  if (ubd_is16(h) && h->tc4_w16 != h->precision) h->tc4_weights_dirty = true;
  if (h->tc4_weights_dirty) {
    const int64_t* d_offs = reinterpret_cast<const int64_t*>((uint8_t*)h->tc_weights.p + kTcZeroOff + tc::ZERO_BYTES + 64);
    tc4::build_img_tf32_kernel<<<UBD_NLAYERS_DIL, 256, 0, h->stream>>>(h->d_params, d_offs, d_offs + 6, (uint8_t*)h->tc4_weights.p);
    tc4::build_img_bf16_kernel<<<UBD_NLAYERS_DIL, 256, 0, h->stream>>>(h->d_params, d_offs, d_offs + 6, (uint8_t*)h->tc4_weights.p + kTc4Tf32, h->precision == UBD_F16);
    // image 6: the stem's L2 (separable 24->24) as one dense 3x3 kernel, merged by tc_prepare into h->l2dense
    tc4::build_img_tf32_kernel<<<1, 256, 0, h->stream>>>((const float*)h->l2dense.p, d_offs + 12, d_offs + 13,
                                                         (uint8_t*)h->tc4_weights.p + (size_t)UBD_NLAYERS_DIL * tc4::WB_BYTES_TF32);
    tc4::build_img_bf16_kernel<<<1, 256, 0, h->stream>>>((const float*)h->l2dense.p, d_offs + 12, d_offs + 13,
                                                         (uint8_t*)h->tc4_weights.p + kTc4Tf32 + (size_t)UBD_NLAYERS_DIL * tc4::WB_BYTES_BF16, h->precision == UBD_F16);
    h->launches += 4;
    UBD_CUDA(cudaGetLastError());
    h->tc4_weights_dirty = false;
    h->tc4_w16 = ubd_is16(h) ? h->precision : h->tc4_w16;
  }
  return UBD_OK;
}

// The head of an out_mode 2 launch goes into the constant bank (tc4::c_headw), stream-ordered in front of the launch.
static int tc4_stage_head(ubd_handle h, const tc::HeadArgs* head) {
  constexpr size_t kBytes = (size_t)(UBD_NF + 1) * tc4::HEAD_STRIDE * sizeof(float);
  if (!h->headw_dev.p) {
    UBD_CUDA(cudaMalloc(&h->headw_dev.p, kBytes));
    h->headw_dev.cap = kBytes;
    h->headw_dirty = true;
  }
  if (h->headw_dirty) {
    tc4::build_headw_kernel<<<1, 256, 0, h->stream>>>(head->hk, head->hb, head->n_out, (float*)h->headw_dev.p);
    ++h->launches;
    UBD_CUDA(cudaGetLastError());
    h->headw_dirty = false;
  }
  UBD_CUDA(cudaMemcpyToSymbolAsync(tc4::c_headw, h->headw_dev.p, kBytes, 0, cudaMemcpyDeviceToDevice, h->stream));
  return UBD_OK;
}

// Training step: between two Adam updates only the six dilated tf32 images are needed (forward and, with flipped kernels,
// backward-data launches).  Rebuild them alone and leave the full-rebuild flags set for the next inference call.
static int tc4_prepare_train(ubd_handle h) {
  if (!h->tc4_weights.p || !h->tc_weights.p) {          // first use of the handle: allocate and build everything once
    int rc = tc4_prepare(h);
    if (rc) return rc;
    h->tc4_train_dirty = false;
    return UBD_OK;
  }
  if (h->tc4_train_dirty) {
    const int64_t* d_offs = reinterpret_cast<const int64_t*>((uint8_t*)h->tc_weights.p + kTcZeroOff + tc::ZERO_BYTES + 64);
    tc4::build_img_tf32_kernel<<<UBD_NLAYERS_DIL, 256, 0, h->stream>>>(h->d_params, d_offs, d_offs + 6, (uint8_t*)h->tc4_weights.p);
    ++h->launches;
    UBD_CUDA(cudaGetLastError());
    h->tc4_train_dirty = false;
  }
  return UBD_OK;
}

// layer 0..5 = conv2d_1..6; layer 6 = the stem's L2 as a dense conv with `in` = uint8 grey image and L1
// computed by the producer warps (l1 != nullptr).
static int tc4_launch_dilconv(ubd_handle h, const void* in, void* out, int layer, int n, int hh, int ww, int d,
                              int out_mode, int out_pad = UBD_MAP_PAD, const tc::HeadArgs* head = nullptr,
                              const tc::L1Args* l1 = nullptr, const uint8_t* wb_override = nullptr, bool train = false) {
  if (h->precision == UBD_FP32) UBD_FAIL(UBD_ERR_UNSUPPORTED, "tensor-core path needs tf32, bf16 or f16");
  int rc = train ? tc4_prepare_train(h) : tc4_prepare(h);
  if (rc) return rc;
  const bool bf16 = ubd_is16(h);
  const uint8_t* base = (const uint8_t*)h->tc4_weights.p;
  const uint8_t* wb = bf16 ? base + kTc4Tf32 + (size_t)layer * tc4::WB_BYTES_BF16 : base + (size_t)layer * tc4::WB_BYTES_TF32;
  if (wb_override) wb = wb_override;         // e.g. the flipped-kernel images of the backward-data pass
  const int sw = ww <= tc4::SW_MAX ? ww : tc4::SW_MAX;
  const int n_strips = (ww + sw - 1) / sw;
  const long long rows = (long long)n * n_strips * hh;
  const int grid = (int)std::min<long long>(rows, h->n_sm);
  tc::HeadArgs ha{};
  if (head) ha = *head;
  if (out_mode == 2 && head) { rc = tc4_stage_head(h, head); if (rc) return rc; }
  if (l1) {       // pw1 and b1 into the constant bank (device-to-device, stream-ordered in front of the launch)
    UBD_CUDA(cudaMemcpyToSymbolAsync(tc4::c_l1w, l1->pw1, UBD_NF * sizeof(float), 0, cudaMemcpyDeviceToDevice, h->stream));
    UBD_CUDA(cudaMemcpyToSymbolAsync(tc4::c_l1w, l1->b1, UBD_NF * sizeof(float), UBD_NF * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  }
  tc::L1Args la{};
  if (l1) la = *l1;
#define UBD_TC4_LAUNCH(BF, L1S, THR)                                                                                   \
  tc4::dilconv_col_kernel<BF, L1S><<<grid, THR, sizeof(tc4::Smem<BF>), h->stream>>>(                                   \
      (const uint4*)in, (uint4*)out, wb, n, hh, ww, d, sw, out_mode, out_pad, tc_err_flag(h), ha, (long long*)h->tc_trace.p, la, tc4::PipeArgs{})
  if (l1) { if (bf16) UBD_TC4_LAUNCH(true, true, tc4::THREADS_L1); else UBD_TC4_LAUNCH(false, true, tc4::THREADS_L1); }
  else { if (bf16) UBD_TC4_LAUNCH(true, false, tc4::THREADS); else UBD_TC4_LAUNCH(false, false, tc4::THREADS); }
#undef UBD_TC4_LAUNCH
  ++h->launches;
  UBD_CUDA(cudaGetLastError());
  return UBD_OK;
}

// All six dilated layers + head of `n` images in ONE launch, layer-pipelined over CTA groups with ring buffers between the
// layers (see tc4::PipeArgs).  in: the stem's output maps of the n images; the head writes head->logits / mask.
// Measured on B200, 64 x 1024^2 tf32 (profiles/r02_summary.md): 0.84 ms (ring 3) against 0.83 ms for one launch per layer,
// DRAM traffic 2.6 GB against 4.2 GB: a group of 25 CTAs splits an image into 10-row pieces, whose two halo rows each cost
// a fifth of the staging and tensor work, and 106 MB of rings do not stay resident in the 126 MB L2 (hit rate 41 %).
static int tc4_launch_pipeline(ubd_handle h, const void* in, int n, int hh, int ww, const tc::HeadArgs* head) {
  if (h->precision == UBD_FP32) UBD_FAIL(UBD_ERR_UNSUPPORTED, "tensor-core path needs tf32, bf16 or f16");
  int rc = tc4_prepare(h);
  if (rc) return rc;
  if (head) { rc = tc4_stage_head(h, head); if (rc) return rc; }
  const bool bf16 = ubd_is16(h);
  const int ring = std::max(2, h->opt_pipe_ring);
  const size_t img_units = act_elems(1, hh, ww, UBD_MAP_PAD) / (bf16 ? 2 : 1);               // 16-byte units per map
  const size_t ring_bytes = (size_t)(UBD_NLAYERS_DIL - 1) * ring * img_units * 16;
  const long long tag = ((long long)h->precision << 56) ^ ((long long)ring << 48) ^ ((long long)hh << 24) ^ (long long)ww;
  if (h->pipe_ring.cap < ring_bytes) {
    if (h->pipe_ring.p) cudaFree(h->pipe_ring.p);
    h->pipe_ring.p = nullptr; h->pipe_ring.cap = 0;
    UBD_CUDA(cudaMalloc(&h->pipe_ring.p, ring_bytes));
    h->pipe_ring.cap = ring_bytes;
    h->pipe_tag = 0;
  }
  if (h->pipe_tag != tag) {                 // the x padding of the ring maps is the convolutions' zero padding
    UBD_CUDA(cudaMemsetAsync(h->pipe_ring.p, 0, h->pipe_ring.cap, h->stream));
    h->pipe_tag = tag;
  }
  const size_t flag_bytes = (size_t)UBD_NLAYERS_DIL * n * sizeof(int);
  if (h->pipe_flags.cap < flag_bytes) {
    if (h->pipe_flags.p) cudaFree(h->pipe_flags.p);
    h->pipe_flags.p = nullptr; h->pipe_flags.cap = 0;
    UBD_CUDA(cudaMalloc(&h->pipe_flags.p, flag_bytes + 1024));
    h->pipe_flags.cap = flag_bytes + 1024;
  }
  UBD_CUDA(cudaMemsetAsync(h->pipe_flags.p, 0, flag_bytes, h->stream));
  const uint8_t* base = (const uint8_t*)h->tc4_weights.p;
  const uint8_t* wb = bf16 ? base + kTc4Tf32 : base;
  const int sw = ww <= tc4::SW_MAX ? ww : tc4::SW_MAX;
  const int grid = h->n_sm;
  tc::HeadArgs ha{};
  if (head) ha = *head;
  tc4::PipeArgs pa{(uint4*)h->pipe_ring.p, (int*)h->pipe_flags.p, ring, (long long)img_units};
  if (bf16)
    tc4::dilconv_col_kernel<true, false, true><<<grid, tc4::THREADS, sizeof(tc4::Smem<true>), h->stream>>>(
        (const uint4*)in, nullptr, wb, n, hh, ww, 1, sw, 0, UBD_MAP_PAD, tc_err_flag(h), ha, (long long*)h->tc_trace.p, tc::L1Args{}, pa);
  else
    tc4::dilconv_col_kernel<false, false, true><<<grid, tc4::THREADS, sizeof(tc4::Smem<false>), h->stream>>>(
        (const uint4*)in, nullptr, wb, n, hh, ww, 1, sw, 0, UBD_MAP_PAD, tc_err_flag(h), ha, (long long*)h->tc_trace.p, tc::L1Args{}, pa);
  ++h->launches;
  UBD_CUDA(cudaGetLastError());
  return UBD_OK;
}
