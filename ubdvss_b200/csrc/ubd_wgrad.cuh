// Weight gradient of a dilated 3x3 24->24 layer on the tensor cores (training step, BASELINE configs[4]; the graph of
// net.py:298-304 differentiated through losses.py by train.py:110-112), sm_100a.
//
//   dK[ky][kx][ic][oc] = sum over pixels (n, y, x) of  X[n][y + (ky-1) d][x + (kx-1) d][ic] * G[n][y][x][oc]
//   dB[oc]             = sum over pixels of G[n][y][x][oc]
//
// is a GEMM whose contraction index is the PIXEL: D[216 (+1), 24] += X^T[216, px] * G[px, 24] (row 216 = ones: the bias).
// Why this one is warp-level mma.sync (m16n8k8, tf32) and not tcgen05: with pixels on K both operands would have to be
// read MN-major (channels contiguous, pixels strided), and for 32-bit operands tcgen05 accepts MN-major only in the
// SWIZZLE_128B_BASE32B layout - 128-byte lines of 32 consecutive channels per pixel - which the plane-interleaved maps
// (16 bytes = 4 channels per pixel and plane) cannot provide without a transposing producer; measured on B200, the
// SWIZZLE_NONE MN-major descriptors of a first version simply read zeros (tools/dbg_wgrad.py).  K-major would need an
// im2col transpose of every tap.  mma.sync reads its fragments with plain shared-memory loads at arbitrary addresses,
// so the three kernel rows are three ring slots, the tap kx is a pixel offset into the zero x-padding, and the whole
// thing is 672 MMAs + 1,568 conflict-free loads per 128-pixel row: ~40 us per layer of a 32 x 128 x 128 batch against
// 1,060 us for the FP32-pipe kernel it replaces (profiles/r02_summary.md).
//
// CTA = 7 consumer warps, each owning two of the fourteen 16-row M tiles (all three 8-column N tiles, 24 accumulators)
// for the CTA's whole contiguous range of rows, + 1 producer warp (cp.async.bulk of map rows into slot rings, rows
// above / below the image from a zero page).  Plane strides are padded by 64 bytes so that the eight rows of a fragment
// (two planes) and its four pixel columns fall on 32 different banks.  Per-CTA partial sums are added in block order by
// reduce_partials_kernel: the step stays bit-reproducible.
#pragma once
#include "ubd_tc4.cuh"

namespace wg {

using tc::smem_u32; using tc::elect_one; using tc::mbar_init; using tc::mbar_arrive; using tc::mbar_expect_tx; using tc::bulk_g2s;
using tc4::mbar_wait3;

constexpr int PAD = UBD_MAP_PAD;
constexpr int SW = 128;                               // strip width in pixels
constexpr int XPLANE = (SW + 2 * PAD) * 16 + 64;      // 2624 B: one plane of a staged X row with its x padding (+ bank skew)
constexpr int XSLOT = UBD_NG * XPLANE;                // 15744 B
constexpr int NSX = 8;
constexpr int GPLANE = SW * 16 + 64;                  // 2112 B
constexpr int GSLOT = UBD_NG * GPLANE;                // 12672 B
constexpr int NSG = 4;
constexpr int N_CONSUMERS = 7;                        // 14 M tiles of 16 rows: (ky, kx, ic) = 216 rows + the ones row
constexpr int THREADS = 32 * (N_CONSUMERS + 1);
constexpr int M_ROWS = 9 * UBD_NF;                    // 216
constexpr int N_PART = M_ROWS * UBD_NF + UBD_NF;      // 5184 kernel + 24 bias partial sums per CTA

struct Smem {
  uint8_t xs[NSX * XSLOT];                            // 125952
  uint8_t gs[NSG * GSLOT];                            // 50688
  uint64_t xfull[NSX], xempty[NSX], gfull[NSG], gempty[NSG];
  int abort_flag;
};

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// x: input map of the layer (a[l]), g: gradient map at the layer's output (already gated by its ReLU); both padded
// row-interleaved fp32 maps (pad = PAD) with values on the tf32 grid.
__global__ void __launch_bounds__(THREADS, 1)
wgrad_kernel(const uint4* __restrict__ x, const uint4* __restrict__ g, const uint8_t* __restrict__ zeros,
             int n_imgs, int h, int w, int d, float* __restrict__ partials, int* gerr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  volatile int* abort_flag = &S.abort_flag;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSX; ++i) { mbar_init(smem_u32(&S.xfull[i]), 1); mbar_init(smem_u32(&S.xempty[i]), N_CONSUMERS); }
    for (int i = 0; i < NSG; ++i) { mbar_init(smem_u32(&S.gfull[i]), 1); mbar_init(smem_u32(&S.gempty[i]), N_CONSUMERS); }
    S.abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // everything a fragment load can touch is finite: zero the rings once (later leftovers are real map data)
  for (int i = threadIdx.x; i < (int)((sizeof(S.xs) + sizeof(S.gs)) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(S.xs)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int wp = w + 2 * PAD;
  tc4::Piece pc;
  tc4::Walk walk(n_imgs, h, w, d, SW, (int)blockIdx.x, (int)gridDim.x);

  if (warp == N_CONSUMERS) {
    // ------------------------------------------------------------------ producer
    uint32_t xq = 0, gq = 0;
    bool ok = true;
    while (ok && walk.next(pc)) {
      const uint32_t xbytes = (uint32_t)(pc.nw + 2 * PAD) * 16u;
      const uint32_t gbytes = (uint32_t)((pc.nw + 7) & ~7) * 16u;
      for (int i = 0; i < pc.rows + 2 && ok; ++i) {
        {
          const uint32_t slot = xq % NSX;
          ok = mbar_wait3(smem_u32(&S.xempty[slot]), ((xq / NSX) & 1u) ^ 1u, abort_flag, gerr, 31, xq);
          if (!ok) break;
          const int jj = pc.j0 - 1 + i;
          const bool valid = jj >= 0 && jj < pc.R;
          const uint4* src = x + (((size_t)pc.n * h + (valid ? pc.c + jj * d : 0)) * UBD_NG) * wp + pc.x0;
          if (elect_one()) {
            const uint32_t bar = smem_u32(&S.xfull[slot]);
            mbar_expect_tx(bar, (uint32_t)UBD_NG * xbytes);
            for (int p = 0; p < UBD_NG; ++p)
              bulk_g2s(smem_u32(S.xs) + slot * XSLOT + p * XPLANE, valid ? (const void*)(src + (size_t)p * wp) : (const void*)zeros, xbytes, bar);
          }
          __syncwarp();
          ++xq;
        }
        if (i >= 2) {
          const uint32_t slot = gq % NSG;
          ok = mbar_wait3(smem_u32(&S.gempty[slot]), ((gq / NSG) & 1u) ^ 1u, abort_flag, gerr, 32, gq);
          if (!ok) break;
          const int y = pc.c + (pc.j0 + i - 2) * d;
          const uint4* src = g + (((size_t)pc.n * h + y) * UBD_NG) * wp + PAD + pc.x0;
          if (elect_one()) {
            const uint32_t bar = smem_u32(&S.gfull[slot]);
            mbar_expect_tx(bar, (uint32_t)UBD_NG * gbytes);
            for (int p = 0; p < UBD_NG; ++p) bulk_g2s(smem_u32(S.gs) + slot * GSLOT + p * GPLANE, src + (size_t)p * wp, gbytes, bar);
          }
          __syncwarp();
          ++gq;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers: M tiles 2 * warp, 2 * warp + 1
    const int gi = lane >> 2, t = lane & 3;
    // fragment rows of this thread: tile s (0, 1), half hf (rows gi, gi + 8): m = 16 (2 warp + s) + gi + 8 hf = (ky, kx, ic)
    int ky_[2][2];
    uint32_t off_[2][2];          // byte offset inside the slot of kernel row ky: plane, tap kx, channel, pixel column t
    uint32_t kind_[2][2];         // 0: data, 1: zero (m > 216), 2: ones (m == 216: bias row)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = 16 * (2 * warp + s) + gi + 8 * hf;
        kind_[s][hf] = m < M_ROWS ? 0u : (m == M_ROWS ? 2u : 1u);
        const int mm = m < M_ROWS ? m : 0;
        const int tap = mm / UBD_NF, ic = mm % UBD_NF;
        ky_[s][hf] = tap / 3;
        off_[s][hf] = (uint32_t)((ic >> 2) * XPLANE + (PAD + (tap % 3 - 1) * d + t) * 16 + (ic & 3) * 4);
      }
    // B fragment: k = pixel t (+4), n = oc = 8 nt + gi
    uint32_t boff[3];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) { const int oc = 8 * nt + gi; boff[nt] = (uint32_t)((oc >> 2) * GPLANE + t * 16 + (oc & 3) * 4); }
    float acc[2][3][4];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[s][nt][e] = 0.f;
    const uint32_t xs0 = smem_u32(S.xs), gs0 = smem_u32(S.gs);
    const uint32_t one = __float_as_uint(1.0f);
    uint32_t xbase = 0, gq = 0;
    bool ok = true;
    while (ok && walk.next(pc)) {
      const int nchunks = (pc.nw + 7) >> 3;
      for (int o = 0; o < pc.rows && ok; ++o) {
        for (int s = (o == 0 ? 0 : 2); s < 3 && ok; ++s) {               // rows of the window that are new
          const uint32_t q = xbase + (uint32_t)o + (uint32_t)s;
          ok = mbar_wait3(smem_u32(&S.xfull[q % NSX]), (q / NSX) & 1u, abort_flag, gerr, 33, q);
        }
        const uint32_t gslot = gq % NSG;
        if (ok) ok = mbar_wait3(smem_u32(&S.gfull[gslot]), (gq / NSG) & 1u, abort_flag, gerr, 34, gq);
        if (!ok) break;
        uint32_t abase[2][2];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
            abase[s][hf] = xs0 + ((xbase + (uint32_t)o + (uint32_t)ky_[s][hf]) % NSX) * XSLOT + off_[s][hf];
        const uint32_t bbase = gs0 + gslot * GSLOT;
        for (int k = 0; k < nchunks; ++k) {
          const uint32_t ko = (uint32_t)k * 128u;
          uint32_t b[3][2];
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) { b[nt][0] = lds32(bbase + boff[nt] + ko); b[nt][1] = lds32(bbase + boff[nt] + ko + 64u); }
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            uint32_t a[4];                                               // a0: (gi, t), a1: (gi + 8, t), a2: (gi, t + 4), a3: (gi + 8, t + 4)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              if (kind_[s][hf] == 0u) { a[hf] = lds32(abase[s][hf] + ko); a[2 + hf] = lds32(abase[s][hf] + ko + 64u); }
              else { a[hf] = a[2 + hf] = kind_[s][hf] == 2u ? one : 0u; }
            }
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) mma_tf32(acc[s][nt], a[0], a[1], a[2], a[3], b[nt][0], b[nt][1]);
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&S.xempty[(xbase + (uint32_t)o) % NSX]));                 // the oldest row of the window is done
          if (o == pc.rows - 1) {
            mbar_arrive(smem_u32(&S.xempty[(xbase + (uint32_t)o + 1u) % NSX]));
            mbar_arrive(smem_u32(&S.xempty[(xbase + (uint32_t)o + 2u) % NSX]));
          }
          mbar_arrive(smem_u32(&S.gempty[gslot]));
        }
        ++gq;
      }
      xbase += (uint32_t)pc.rows + 2u;
    }
    // partial sums of this CTA, Keras HWIO order: [m = (ky*3 + kx)*24 + ic][oc], then the bias row
    float* out = partials + (size_t)blockIdx.x * N_PART;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = 16 * (2 * warp + s) + gi + 8 * hf;
        if (m <= M_ROWS) {
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
            *reinterpret_cast<float2*>(out + (size_t)m * UBD_NF + 8 * nt + 2 * t) = make_float2(acc[s][nt][2 * hf], acc[s][nt][2 * hf + 1]);
        }
      }
  }
}


// ------------------------------------------------------------------------------------------------------------------------
// Pointwise (1x1, 24 -> 24) weight gradient of a separable stem layer: dPW[ic][oc] = sum_px X[px][ic] * G[px][oc] with
// X = the layer's depthwise output (kept by the training forward), dB[oc] = sum_px G[px][oc].  Same mma.sync scheme with
// M = 24 (+ the ones row) in two M tiles; the seven consumer warps split the 8-pixel K chunks of every staged row among
// themselves and add their accumulators in warp order at the end.  fp32 maps: the MMA reads their upper 19 bits.
constexpr int PW_SW = 256;                            // strip width
constexpr int PW_PLANE = PW_SW * 16 + 64;             // 4160 B (+ bank skew)
constexpr int PW_STAGE = 2 * UBD_NG * PW_PLANE;       // X row + G row: 49920 B
constexpr int PW_NST = 3;
constexpr int PW_ROWS = UBD_NF + 1;                   // 24 channels + the ones row
constexpr int PW_PART = PW_ROWS * UBD_NF;             // 576 + 24 partial sums per CTA

struct PwSmem {
  uint8_t st[PW_NST * PW_STAGE];                      // 149760; reused for the cross-warp sum at the end
  uint64_t full[PW_NST], empty[PW_NST];
  int abort_flag;
};

__global__ void __launch_bounds__(THREADS, 1)
pwgrad_kernel(const uint4* __restrict__ x, int xpad, const uint4* __restrict__ g, int gpad, int n_imgs, int h, int w,
              float* __restrict__ partials, int* gerr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  PwSmem& S = *reinterpret_cast<PwSmem*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  volatile int* abort_flag = &S.abort_flag;
  if (threadIdx.x == 0) {
    for (int i = 0; i < PW_NST; ++i) { mbar_init(smem_u32(&S.full[i]), 1); mbar_init(smem_u32(&S.empty[i]), N_CONSUMERS); }
    S.abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (int)(sizeof(S.st) / 16); i += blockDim.x) reinterpret_cast<uint4*>(S.st)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int n_strips = (w + PW_SW - 1) / PW_SW;
  const long long total = (long long)n_imgs * h * n_strips;
  const long long u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;
  const int xwp = w + 2 * xpad, gwp = w + 2 * gpad;

  if (warp == N_CONSUMERS) {
    bool ok = true;
    uint32_t q = 0;
    for (long long u = u0; u < u1 && ok; ++u, ++q) {
      const int strip = (int)(u % n_strips);
      const long long row = u / n_strips;                 // n * h + y
      const int x0 = strip * PW_SW, nw = min(PW_SW, w - x0);
      const uint32_t stg = q % PW_NST;
      ok = mbar_wait3(smem_u32(&S.empty[stg]), ((q / PW_NST) & 1u) ^ 1u, abort_flag, gerr, 36, q);
      if (!ok) break;
      if (elect_one()) {
        const uint32_t bar = smem_u32(&S.full[stg]);
        const uint32_t bytes = (uint32_t)nw * 16u;
        mbar_expect_tx(bar, 2u * UBD_NG * bytes);
        const uint4* xs = x + (size_t)row * UBD_NG * xwp + xpad + x0;
        const uint4* gs = g + (size_t)row * UBD_NG * gwp + gpad + x0;
        const uint32_t dst = smem_u32(S.st) + stg * PW_STAGE;
        for (int p = 0; p < UBD_NG; ++p) {
          bulk_g2s(dst + p * PW_PLANE, xs + (size_t)p * xwp, bytes, bar);
          bulk_g2s(dst + (UBD_NG + p) * PW_PLANE, gs + (size_t)p * gwp, bytes, bar);
        }
      }
      __syncwarp();
    }
  } else {
    const int gi = lane >> 2, t = lane & 3;
    uint32_t aoff[2][2], kind[2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = 16 * s + gi + 8 * hf;
        kind[s][hf] = m < UBD_NF ? 0u : (m == UBD_NF ? 2u : 1u);
        const int ic = m < UBD_NF ? m : 0;
        aoff[s][hf] = (uint32_t)((ic >> 2) * PW_PLANE + t * 16 + (ic & 3) * 4);
      }
    uint32_t boff[3];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) { const int oc = 8 * nt + gi; boff[nt] = (uint32_t)((UBD_NG + (oc >> 2)) * PW_PLANE + t * 16 + (oc & 3) * 4); }
    float acc[2][3][4];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[s][nt][e] = 0.f;
    const uint32_t one = __float_as_uint(1.0f);
    bool ok = true;
    uint32_t q = 0;
    for (long long u = u0; u < u1 && ok; ++u, ++q) {
      const int strip = (int)(u % n_strips);
      const int nw = min(PW_SW, w - strip * PW_SW);
      const uint32_t stg = q % PW_NST;
      ok = mbar_wait3(smem_u32(&S.full[stg]), (q / PW_NST) & 1u, abort_flag, gerr, 37, q);
      if (!ok) break;
      const uint32_t base = smem_u32(S.st) + stg * PW_STAGE;
      const int nchunks = (nw + 7) >> 3;
      for (int k = warp; k < nchunks; k += N_CONSUMERS) {
        const uint32_t ko = (uint32_t)k * 128u;
        const bool half = 8 * k + 4 >= nw;                 // ragged end: the pixels t + 4 of the last chunk do not exist
        uint32_t b[3][2];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) { b[nt][0] = lds32(base + boff[nt] + ko); b[nt][1] = half ? 0u : lds32(base + boff[nt] + ko + 64u); }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          uint32_t a[4];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (kind[s][hf] == 0u) { a[hf] = lds32(base + aoff[s][hf] + ko); a[2 + hf] = lds32(base + aoff[s][hf] + ko + 64u); }
            else { a[hf] = a[2 + hf] = kind[s][hf] == 2u ? one : 0u; }
          }
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) mma_tf32(acc[s][nt], a[0], a[1], a[2], a[3], b[nt][0], b[nt][1]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&S.empty[stg]));
    }
    // sum over the seven warps in warp order (the staging area is free: every copy has been waited for)
    asm volatile("bar.sync 1, %0;" ::"n"(32 * N_CONSUMERS) : "memory");
    float* red = reinterpret_cast<float*>(S.st) + (size_t)warp * 32 * UBD_NF;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int m = 16 * s + gi + 8 * hf;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          red[m * UBD_NF + 8 * nt + 2 * t] = acc[s][nt][2 * hf];
          red[m * UBD_NF + 8 * nt + 2 * t + 1] = acc[s][nt][2 * hf + 1];
        }
      }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * N_CONSUMERS) : "memory");
    const float* all = reinterpret_cast<const float*>(S.st);
    for (int i = (int)threadIdx.x; i < PW_PART; i += 32 * N_CONSUMERS) {
      float sum = 0.f;
#pragma unroll
      for (int wi = 0; wi < N_CONSUMERS; ++wi) sum += all[(size_t)wi * 32 * UBD_NF + i];
      partials[(size_t)blockIdx.x * PW_PART + i] = sum;
    }
  }
}

}  // namespace wg
