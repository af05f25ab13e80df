// GPU connected components with cv2.findContours(RETR_EXTERNAL) semantics (utils.py:51-60,
// segmap_manager.py:54-69; exact statement in SURVEY.md 8a/P2 and oracle/postproc.py::ccl_spec):
//   1. union-find over all pixels: foreground 8-connected, background 4-connected -- tile-local in
//      shared memory (warp-ballot run starts + atomicMin links inside 32x32 tiles), then a boundary
//      merge over the tile borders; the background sets that own an image-border pixel are flagged
//      "outer" (connected to the outside of the image);
//   2. "filled" = foreground or background whose set is not outer (holes); union 8-adjacent filled;
//   3. label = root = smallest raster index of the filled component (first pixel in raster order);
//   4. per-component reductions: bbox, foreground / filled pixel counts, 2x2 bit-quad counts
//      (2*contourArea = 2*#Q4 + #Q3), class-probability sums over the filled pixels.
// Unions always attach the larger root under the smaller one with atomicMin, so every parent
// chain is strictly decreasing and the final root is the minimum index of the set.
#pragma once
#include "ubd_common.cuh"

struct CompRec {            // device-side accumulator, one per component (slot order = raster order)
  int label;
  int xmin, ymin, xmax, ymax;
  int n_pixels, n_filled;
  int q3, q4;
};

__device__ __forceinline__ int uf_find(const int* parent, int i) {
  int p;
  while ((p = *((volatile const int*)(parent + i))) != i) i = p;
  return i;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }      // a > b: hang a under b
    const int old = atomicMin(parent + a, b);
    if (old == a) return;
    a = old;                                      // a was no longer a root; retry from its parent
  }
}

// ------------------------------------------------------------------------------------------------
// Phase 1 (foreground 8-connectivity, background 4-connectivity), tile-local then boundary merge.
// ------------------------------------------------------------------------------------------------
constexpr int CCL_T = 32;     // tile side; block = 32 x 8 threads, 4 rows per thread

__device__ __forceinline__ int suf_find(const int* par, int i) {      // shared-memory find
  int p;
  while ((p = *((volatile const int*)(par + i))) != i) i = p;
  return i;
}
__device__ __forceinline__ void suf_union(int* par, int a, int b) {   // shared-memory union (atomicMin)
  while (true) {
    a = suf_find(par, a);
    b = suf_find(par, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    const int old = atomicMin(par + a, b);
    if (old == a) return;
    a = old;
  }
}

// One 32x32 tile per block, fully labelled in shared memory: run starts by warp ballot (horizontal
// links cost nothing), vertical / diagonal links with the decision tree (one link per run overlap),
// then every pixel's parent is written as the GLOBAL raster index of its tile-local root.
__global__ void __launch_bounds__(256)
ccl_local_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, int h, int w, size_t pstride) {
  __shared__ int par[CCL_T * CCL_T];
  __shared__ uint8_t m[CCL_T * CCL_T];
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * CCL_T, y0 = blockIdx.y * CCL_T;
  const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
  const uint8_t* gm = mask + (size_t)n * h * w;
  const int gx = x0 + lx;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = ly0 + 8 * k, gy = y0 + ly;
    const bool in = gx < w && gy < h;
    const bool fg = in && gm[(size_t)gy * w + gx] != 0;
    // pixels outside the image get class 2: they never match a real pixel
    m[ly * CCL_T + lx] = in ? (fg ? 1 : 0) : 2;
    const unsigned bits = __ballot_sync(0xffffffffu, fg);
    const unsigned inb = __ballot_sync(0xffffffffu, in);
    const unsigned same = fg ? bits : (in ? (~bits & inb) : ~inb);
    const unsigned below = (~same) & ((1u << lx) - 1u);
    const int start = below ? 32 - __clz(below) : 0;
    par[ly * CCL_T + lx] = ly * CCL_T + start;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = ly0 + 8 * k;
    if (ly == 0) continue;
    const int p = ly * CCL_T + lx;
    const int c = m[p];
    if (c == 2) continue;
    const bool hasW = lx > 0, hasE = lx < CCL_T - 1;
    const bool N_ = m[p - CCL_T] == c;
    if (c == 1) {
      if (N_) {
        if (!hasW || m[p - 1] != 1 || m[p - CCL_T - 1] != 1) suf_union(par, p, p - CCL_T);
      } else {
        if (hasE && m[p - CCL_T + 1] == 1) suf_union(par, p, p - CCL_T + 1);
        if (hasW && m[p - CCL_T - 1] == 1) suf_union(par, p, p - CCL_T - 1);
      }
    } else if (N_ && (!hasW || m[p - 1] != 0 || m[p - CCL_T - 1] != 0)) {
      suf_union(par, p, p - CCL_T);
    }
  }
  __syncthreads();
  int* gp = parent + (size_t)n * pstride;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = ly0 + 8 * k, gy = y0 + ly;
    if (gx < w && gy < h) {
      const int r = suf_find(par, ly * CCL_T + lx);
      gp[(size_t)gy * w + gx] = (y0 + r / CCL_T) * w + x0 + (r % CCL_T);
    }
  }
}

// Boundary merge: global unions only for pixels on a tile border, towards their cross-tile neighbours
// among {W, NW, N, NE} (foreground) or {W, N} (background).  Thread index enumerates the
// border pixels: the first row of every tile row (y = 32k, k >= 1), then the first and last column
// of every tile column for the remaining rows.
__global__ void __launch_bounds__(256)
ccl_border_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, int h, int w, size_t pstride) {
  const int n = blockIdx.y;
  const int nrows = (h - 1) / CCL_T;                 // tile-row boundaries y = 32, 64, ...
  const int ncols = (w - 1) / CCL_T;                 // tile-column boundaries x = 32, 64, ...
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int x, y;
  bool row_border;
  if (i < nrows * w) { y = (i / w + 1) * CCL_T; x = i % w; row_border = true; }
  else {
    const int j = i - nrows * w;
    if (j >= ncols * 2 * h) return;
    y = j % h;
    const int cb = j / h;                            // 2 * boundary + side
    x = (cb / 2 + 1) * CCL_T - (cb & 1);             // side 0: first column of the right tile; 1: last column of the left
    row_border = false;
    if ((y % CCL_T) == 0 && y > 0) return;           // already handled as a row-border pixel
  }
  const uint8_t* m = mask + (size_t)n * h * w;
  int* par = parent + (size_t)n * pstride;
  const int p = y * w + x;
  const bool c = m[p] != 0;
  const bool hasW = x > 0, hasN = y > 0, hasE = x < w - 1;
  const bool tile_x0 = (x % CCL_T) == 0, tile_x31 = (x % CCL_T) == CCL_T - 1;
  if (row_border) {
    // N, NW, NE are all in the tile row above
    if (c) {
      if (m[p - w] != 0) {
        // p~W and NW~N are tile-local links (if W, NW are in my tile column), W~NW is W's link
        if (tile_x0 || m[p - 1] == 0 || m[p - w - 1] == 0) uf_union(par, p, p - w);
      } else {
        if (hasE && m[p - w + 1] != 0) uf_union(par, p, p - w + 1);
        if (hasW && m[p - w - 1] != 0) uf_union(par, p, p - w - 1);
      }
      if (tile_x0 && hasW && m[p - 1] != 0) uf_union(par, p, p - 1);
    } else {
      if (m[p - w] == 0 && (tile_x0 || m[p - 1] != 0 || m[p - w - 1] != 0)) uf_union(par, p, p - w);
      if (tile_x0 && hasW && m[p - 1] == 0) uf_union(par, p, p - 1);
    }
  } else if (tile_x0) {
    // first column of a tile, not on a row border: W and NW are in the tile to the left
    if (c) {
      if (hasW && m[p - 1] != 0) uf_union(par, p, p - 1);
      else if (hasW && hasN && m[p - w - 1] != 0) uf_union(par, p, p - w - 1);
    } else if (hasW && m[p - 1] == 0) {
      uf_union(par, p, p - 1);
    }
  } else if (tile_x31) {
    // last column of a tile: only the NE diagonal crosses (E's own W-link covers p~E)
    if (c && hasN && hasE && m[p - w + 1] != 0 && m[p - w] == 0 && m[p + 1] == 0) uf_union(par, p, p - w + 1);
  }
}

// Full path compression: parent[p] = root(p); clears the outer flags for the next kernel.
__global__ void __launch_bounds__(256)
ccl_flatten_kernel(int* __restrict__ parent, uint8_t* __restrict__ outer, int h, int w, size_t pstride) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w) return;
  int* par = parent + (size_t)n * pstride;
  par[i] = uf_find(par, i);
  outer[(size_t)n * pstride + i] = 0;
}

// outer[root] = 1 for every background set that owns a pixel on the image border (those are
// 4-connected to the outside); plain stores of the same value, no atomics.  Needs flattened parents.
__global__ void __launch_bounds__(256)
ccl_mark_outer_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ parent,
                      uint8_t* __restrict__ outer, int h, int w, size_t pstride) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // index along the border: 2w + 2h
  int y, x;
  if (i < w) { y = 0; x = i; }
  else if (i < 2 * w) { y = h - 1; x = i - w; }
  else if (i < 2 * w + h) { y = i - 2 * w; x = 0; }
  else if (i < 2 * w + 2 * h) { y = i - 2 * w - h; x = w - 1; }
  else return;
  const int p = y * w + x;
  if (mask[(size_t)n * h * w + p] == 0) outer[(size_t)n * pstride + parent[(size_t)n * pstride + p]] = 1;
}

// Phase 2: 8-connectivity over filled pixels.  Requires flattened parents from phase 1: then a
// non-root's parent never changes again, and a hole root is only ever re-parented inside its own
// filled set, so "background q is outer" == outer[parent[q]] at any time during this kernel.
// Pairs of two foreground pixels are already connected by phase 1 and are skipped.
__global__ void __launch_bounds__(256)
ccl_merge2_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, const uint8_t* __restrict__ outer,
                  int h, int w, size_t pstride) {
  const int n = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const uint8_t* m = mask + (size_t)n * h * w;
  int* par = parent + (size_t)n * pstride;
  const uint8_t* out = outer + (size_t)n * pstride;
  const int p = y * w + x;
  const bool fg = m[p] != 0;
  auto filled = [&](int q) -> bool { return m[q] != 0 || out[*((volatile int*)(par + q))] == 0; };
  if (!fg && !filled(p)) return;
  const bool hasW = x > 0, hasN = y > 0, hasE = x < w - 1;
  auto link = [&](int q) { if (!(fg && m[q] != 0)) uf_union(par, p, q); };
  if (hasN && filled(p - w)) {
    link(p - w);
  } else {
    if (hasN && hasE && filled(p - w + 1)) link(p - w + 1);
    if (hasN && hasW && filled(p - w - 1)) link(p - w - 1);
    else if (hasW && filled(p - 1)) link(p - 1);
  }
}

// labels[p] = filled ? root : -1.  An outer background pixel was never re-parented, so its parent is
// still its phase-1 root and carries the outer flag; everything else is a filled pixel.
__global__ void __launch_bounds__(256)
ccl_label_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, const uint8_t* __restrict__ outer,
                 int* __restrict__ labels, int h, int w, size_t pstride) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w) return;
  int* par = parent + (size_t)n * pstride;
  const bool filled = mask[(size_t)n * h * w + i] != 0 || outer[(size_t)n * pstride + par[i]] == 0;
  labels[(size_t)n * h * w + i] = filled ? uf_find(par, i) : -1;
}

// One CTA per image: rank the roots in raster order -> slot_of[root], n_comps[n]; initialise records.
__global__ void __launch_bounds__(1024)
ccl_slots_kernel(const int* __restrict__ labels, int* __restrict__ slot_of, CompRec* __restrict__ comps,
                 unsigned long long* __restrict__ cls_sums, int n_cls,
                 int* __restrict__ n_comps, int h, int w, int max_comps) {
  const int n = blockIdx.x;
  const int* lab = labels + (size_t)n * h * w;
  int* so = slot_of + (size_t)n * h * w;
  CompRec* cr = comps + (size_t)n * max_comps;
  __shared__ int warp_cnt[32];
  __shared__ int warp_off[32];
  __shared__ int chunk_total;
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int hw = h * w;
  const bool vec = (hw & 3) == 0;                            // rows of 4 labels are 16-byte aligned
  // a thread ranks 4 consecutive pixels per round: 16 rounds of 3 barriers for a 256x256 map
  for (int i0 = 0; i0 < hw; i0 += (int)blockDim.x * 4) {
    const int ib = i0 + (int)threadIdx.x * 4;
    int l4[4] = {-1, -1, -1, -1};
    if (vec) {
      if (ib < hw) { const int4 q = *reinterpret_cast<const int4*>(lab + ib); l4[0] = q.x; l4[1] = q.y; l4[2] = q.z; l4[3] = q.w; }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) if (ib + k < hw) l4[k] = lab[ib + k];
    }
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt += (l4[k] == ib + k) ? 1 : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_cnt[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int own = lane < nwarps ? warp_cnt[lane] : 0;
      int v = own;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
      warp_off[lane] = v - own;
      if (lane == 31) chunk_total = v;
    }
    __syncthreads();
    int slot = base + warp_off[wid] + incl - cnt;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (l4[k] == ib + k) {
        const int i = ib + k;
        so[i] = slot;
        if (slot < max_comps) {
          CompRec r; r.label = i; r.xmin = w; r.ymin = h; r.xmax = -1; r.ymax = -1;
          r.n_pixels = 0; r.n_filled = 0; r.q3 = 0; r.q4 = 0;
          cr[slot] = r;
          for (int c = 0; c < n_cls; ++c) cls_sums[((size_t)n * max_comps + slot) * n_cls + c] = 0ull;
        }
        ++slot;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) base += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_comps[n] = base;
}

__device__ __forceinline__ int group_sum(unsigned grp, int v) { return __reduce_add_sync(grp, v); }

// Thread = one 2x2 window whose bottom-right pixel is (y,x), y in [0,h], x in [0,w]: contributes the
// pixel (y,x) itself (bbox, counts, class probabilities) and the bit-quad of the window.
// Lanes of a warp that hit the same component are combined first (match_any + redux).
__global__ void __launch_bounds__(256)
ccl_stats_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ labels,
                 const int* __restrict__ slot_of, CompRec* __restrict__ comps,
                 const float* __restrict__ cls_logits, int cls_stride,
                 unsigned long long* __restrict__ cls_sums, int n_cls,
                 int h, int w, int max_comps) {
  const int n = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int* lab = labels + (size_t)n * h * w;
  const int* so = slot_of + (size_t)n * h * w;
  CompRec* cr = comps + (size_t)n * max_comps;
  auto L = [&](int yy, int xx) -> int {
    return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? lab[yy * w + xx] : -1;
  };
  int slot = -1, l00 = -1;
  int q3 = 0, q4 = 0;
  if (y <= h && x <= w) {
    const int a = L(y - 1, x - 1), b = L(y - 1, x), c = L(y, x - 1);
    l00 = L(y, x);
    const int cnt = (a >= 0) + (b >= 0) + (c >= 0) + (l00 >= 0);
    const int any = max(max(a, b), max(c, l00));          // all non-negative ones are equal
    if (any >= 0) slot = so[any];
    q3 = cnt == 3; q4 = cnt == 4;
  }
  if (slot >= max_comps) slot = -1;                       // overflow: reported via n_comps > max_comps
  const unsigned active = __ballot_sync(0xffffffffu, slot >= 0);
  if (slot < 0) return;
  const unsigned grp = __match_any_sync(active, slot);
  const int leader = __ffs(grp) - 1;
  const int lane = threadIdx.x & 31;
  const bool own = l00 >= 0;
  const int sq3 = group_sum(grp, q3), sq4 = group_sum(grp, q4);
  const int sfill = group_sum(grp, own ? 1 : 0);
  const int spix = group_sum(grp, (own && mask[(size_t)n * h * w + y * w + x] != 0) ? 1 : 0);
  const int xmn = __reduce_min_sync(grp, own ? x : 0x7fffffff);
  const int ymn = __reduce_min_sync(grp, own ? y : 0x7fffffff);
  const int xmx = __reduce_max_sync(grp, own ? x : -1);
  const int ymx = __reduce_max_sync(grp, own ? y : -1);
  if (lane == leader) {
    CompRec* r = cr + slot;
    if (sq3) atomicAdd(&r->q3, sq3);
    if (sq4) atomicAdd(&r->q4, sq4);
    if (sfill) {
      atomicAdd(&r->n_filled, sfill);
      if (spix) atomicAdd(&r->n_pixels, spix);
      atomicMin(&r->xmin, xmn); atomicMin(&r->ymin, ymn);
      atomicMax(&r->xmax, xmx); atomicMax(&r->ymax, ymx);
    }
  }
  if (n_cls > 0) {
    // np_softmax (utils.py:135-138) in float32, accumulated in 2^-24 fixed point so that the sum
    // does not depend on the order of the atomics (segmap_manager.py:65 takes the mean, argmax).
    float e[UBD_MAX_CLASSES];
    float s = 1.f;
    if (own) {
      const float* lg = cls_logits + ((size_t)n * h * w + y * w + x) * cls_stride;
      float mx = lg[0];
      for (int c = 1; c < n_cls; ++c) mx = fmaxf(mx, lg[c]);
      s = 0.f;
      for (int c = 0; c < n_cls; ++c) { e[c] = expf(lg[c] - mx); s += e[c]; }
    }
    for (int c = 0; c < n_cls; ++c) {
      const unsigned fx = own ? (unsigned)(e[c] / s * 16777216.0f) : 0u;
      const unsigned tot = __reduce_add_sync(grp, fx);
      if (lane == leader && tot)
        atomicAdd(&cls_sums[((size_t)n * max_comps + slot) * n_cls + c], (unsigned long long)tot);
    }
  }
}

// Kept components (2*contourArea > min_area_x2, utils.py:55), compacted image-major and, inside an
// image, in DESCENDING label order (cv2 lists external contours bottom-up).

__global__ void __launch_bounds__(256)
ccl_count_kept_kernel(const CompRec* __restrict__ comps, const int* __restrict__ n_comps,
                      int* __restrict__ kept_count, CclTotals* __restrict__ totals,
                      int max_comps, int min_area_x2) {
  const int n = blockIdx.x;
  const int cnt = min(n_comps[n], max_comps);
  const CompRec* cr = comps + (size_t)n * max_comps;
  int k = 0;
  for (int s = threadIdx.x; s < cnt; s += blockDim.x) k += (2 * cr[s].q4 + cr[s].q3) > min_area_x2;
  __shared__ int red[8];
  k = __reduce_add_sync(0xffffffffu, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    kept_count[n] = t;
    atomicAdd(&totals->total_kept, t);
    atomicMax(&totals->max_ncomp, n_comps[n]);
  }
}

__global__ void __launch_bounds__(256)
ccl_compact_kernel(const CompRec* __restrict__ comps, const unsigned long long* __restrict__ cls_sums,
                   int n_cls, const int* __restrict__ n_comps, const int* __restrict__ kept_count,
                   OutRec* __restrict__ out, int* __restrict__ out_index_of_slot,
                   int max_comps, int max_out, int min_area_x2,
                   CclTotals* __restrict__ totals, int* __restrict__ ext, int max_rows) {
  const int n = blockIdx.x;
  const int cnt = min(n_comps[n], max_comps);
  const CompRec* cr = comps + (size_t)n * max_comps;
  __shared__ int s_base;
  __shared__ int warp_cnt[8], warp_off[8], chunk_total;
  if (threadIdx.x == 0) {
    int o = 0;
    for (int i = 0; i < n; ++i) o += kept_count[i];
    s_base = o;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i0 = 0; i0 < cnt; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const int s = cnt - 1 - i;                       // descending slot = descending label
    bool keep = false;
    CompRec r;
    if (i < cnt) { r = cr[s]; keep = (2 * r.q4 + r.q3) > min_area_x2; }
    const unsigned bits = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(bits);
    __syncthreads();
    if (wid == 0) {
      const int own = lane < nwarps ? warp_cnt[lane] : 0;
      int v = own;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
      if (lane < nwarps) warp_off[lane] = v - own;
      if (lane == 31) chunk_total = v;
    }
    __syncthreads();
    if (i < cnt) {
      int oi = -1;
      if (keep) {
        oi = s_base + warp_off[wid] + __popc(bits & ((1u << lane) - 1u));
        if (oi < max_out) {
          OutRec o;
          o.image = n; o.label = r.label; o.xmin = r.xmin; o.ymin = r.ymin; o.xmax = r.xmax; o.ymax = r.ymax;
          o.n_pixels = r.n_pixels; o.n_filled = r.n_filled; o.area_x2 = 2 * r.q4 + r.q3; o.slot = s;
          int best = -1;
          unsigned long long bv = 0ull;
          for (int c = 0; c < n_cls; ++c) {            // np.argmax: first maximum wins
            const unsigned long long v = cls_sums[((size_t)n * max_comps + s) * n_cls + c];
            if (best < 0 || v > bv) { best = c; bv = v; }
          }
          o.class_id = best;
          // rows of the component in the row-extent arrays (ccl_extents_kernel fills them, ccl_boxes_kernel reads them):
          // ext[2r] = max(w-1-x) over the row's pixels (left end), ext[2r+1] = max(x) (right end), both start at -1
          o.row_base = -1;
          if (ext != nullptr) {
            const int rows = r.ymax - r.ymin + 1;
            const int rb = atomicAdd(&totals->total_rows, rows);
            if (rb + rows <= max_rows) {
              o.row_base = rb;
              for (int k = 0; k < 2 * rows; ++k) ext[2 * (size_t)rb + k] = -1;
            }
          }
          out[oi] = o;
        } else {
          oi = -1;
        }
      }
      out_index_of_slot[(size_t)n * max_comps + s] = oi;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += chunk_total;
    __syncthreads();
  }
}

// Hull candidates of kept components.  A convex-hull vertex of a component uniquely maximises u.q
// over its pixels for some direction u, so its neighbour in the x-direction of u and its neighbour
// in the y-direction of u are both outside the component: only such corner pixels are emitted
// (a superset of the hull vertices, all that cv2.minAreaRect needs, utils.py:56).
__global__ void __launch_bounds__(256)
ccl_points_kernel(const int* __restrict__ labels, const int* __restrict__ slot_of,
                  const int* __restrict__ out_index_of_slot, HullPt* __restrict__ pts,
                  CclTotals* __restrict__ totals, int h, int w, int max_comps, int max_pts) {
  const int n = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  int comp = -1;
  if (x < w && y < h) {
    const int* lab = labels + (size_t)n * h * w;
    const int l = lab[y * w + x];
    if (l >= 0) {
      const bool wo = x == 0 || lab[y * w + x - 1] != l;
      const bool eo = x == w - 1 || lab[y * w + x + 1] != l;
      const bool no = y == 0 || lab[(y - 1) * w + x] != l;
      const bool so = y == h - 1 || lab[(y + 1) * w + x] != l;
      if ((wo || eo) && (no || so)) {
        const int slot = slot_of[(size_t)n * h * w + l];
        if (slot < max_comps) comp = out_index_of_slot[(size_t)n * max_comps + slot];
      }
    }
  }
  const bool emit = comp >= 0;
  const unsigned bits = __ballot_sync(0xffffffffu, emit);
  if (!bits) return;
  const int lane = threadIdx.x & 31;
  const int lead = __ffs(bits) - 1;
  int base = 0;
  if (lane == lead) base = atomicAdd(&totals->total_pts, __popc(bits));
  base = __shfl_sync(0xffffffffu, base, lead);
  if (emit) {
    const int idx = base + __popc(bits & ((1u << lane) - 1u));
    if (idx < max_pts) { HullPt p; p.comp = comp; p.xy = (y << 16) | x; pts[idx] = p; }
  }
}

// ------------------------------------------------------------------------------------------------
// Min-area rectangle of every kept component on the GPU (utils.py:56-57, cv2.minAreaRect): one warp per
// component.  (1) the warp scans the component's bounding box in the filled label image for the left- and
// right-most pixel of every row - only those can be hull vertices; (2) lane 0 builds the strict convex hull
// from the two y-monotone chains (integer cross products) in the order ubd_rect.cpp feeds the calipers:
// top edge left to right, right side down, bottom edge right to left, left side up, ENDING at the top-most,
// then left-most pixel (the contour's first point); (3) lane 0 runs the float32 rotating calipers with exactly
// the host's operation order (explicit round-to-nearest intrinsics: no FMA contraction, double sqrt / divide
// are IEEE on both sides), so the six numbers it leaves are bit-identical to ubd_min_area_box's.  The host only
// adds the angle / corner trigonometry of cv2.boxPoints (libm), a few hundred nanoseconds per component.
// ------------------------------------------------------------------------------------------------

// Left / right end of every row of every kept component: one thread per pixel, only run ends touch memory.
__global__ void __launch_bounds__(256)
ccl_extents_kernel(const int* __restrict__ labels, const int* __restrict__ slot_of, const int* __restrict__ out_index_of_slot,
                   const OutRec* __restrict__ recs, int* __restrict__ ext, int h, int w, int max_comps) {
  const int n = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const int* lab = labels + (size_t)n * h * w;
  const int l = lab[y * w + x];
  if (l < 0) return;
  const bool le = x == 0 || lab[y * w + x - 1] != l;
  const bool re = x == w - 1 || lab[y * w + x + 1] != l;
  if (!le && !re) return;
  const int slot = slot_of[(size_t)n * h * w + l];
  if (slot >= max_comps) return;
  const int comp = out_index_of_slot[(size_t)n * max_comps + slot];
  if (comp < 0) return;
  const OutRec& r = recs[comp];
  if (r.row_base < 0) return;
  int* e = ext + 2 * ((size_t)r.row_base + (y - r.ymin));
  if (le) atomicMax(e, w - 1 - x);
  if (re) atomicMax(e + 1, x);
}

__device__ __forceinline__ long long ccl_cross(int ox, int oy, int ax, int ay, int bx, int by) {
  return (long long)(ax - ox) * (by - oy) - (long long)(ay - oy) * (bx - ox);
}

// cv2.boxPoints(cv2.minAreaRect(contour)) of one kept component per warp: row extents -> integer strict hull ->
// float32 rotating calipers in the operation order and with the tie rule of OpenCV's rotatingCalipers()
// (modules/imgproc/src/rotcalipers.cpp, Copyright (C) 2000 Intel Corporation / OpenCV contributors, Apache-2.0 since
// OpenCV 4.5; see the note in ubd_rect.cpp), so that the boxes are bit-identical to the host path and to cv2.
__global__ void __launch_bounds__(32)
ccl_boxes_kernel(const int* __restrict__ ext, const OutRec* __restrict__ recs, const CclTotals* __restrict__ totals,
                 BoxRec* __restrict__ boxes, int h, int w, int max_out) {
  extern __shared__ int box_smem[];
  const int comp = blockIdx.x, lane = threadIdx.x;
  if (comp >= min(totals->total_kept, max_out)) return;
  const OutRec r = recs[comp];
  const int rows = r.ymax - r.ymin + 1;
  int* L = box_smem;                       // [rows]
  int* R = L + h;                          // [rows]
  int* hx = R + h;                         // [2 * rows + 2]
  int* hy = hx + 2 * h + 2;
  float* vx = reinterpret_cast<float*>(hy + 2 * h + 2);
  float* vy = vx + 2 * h + 2;
  float* inv = vy + 2 * h + 2;
  if (r.row_base < 0) {                    // the row-extent arrays overflowed (reported by the host): leave a marker
    if (lane == 0) { BoxRec o; o.cx = o.cy = o.w = o.h = o.ax = o.ay = 0.f; o.n_hull = -1; o.x0 = o.y0 = o.x1 = o.y1 = 0; boxes[comp] = o; }
    return;
  }
  // (1) row extents
  const int* e = ext + 2 * (size_t)r.row_base;
  for (int i = lane; i < rows; i += 32) { L[i] = w - 1 - e[2 * i]; R[i] = e[2 * i + 1]; }
  __syncwarp();
  if (lane != 0) return;
  // (2) hull.  Every row of an 8-connected filled component between ymin and ymax holds at least one pixel.
  int k = 0;
  for (int i = 0; i < rows; ++i) {                       // right side, top to bottom
    const int px = R[i], py = r.ymin + i;
    while (k >= 2 && ccl_cross(hx[k - 2], hy[k - 2], hx[k - 1], hy[k - 1], px, py) <= 0) --k;
    hx[k] = px; hy[k] = py; ++k;
  }
  const int t = k + 1;
  for (int i = rows - 1; i >= 0; --i) {                  // left side, bottom to top
    const int px = L[i], py = r.ymin + i;
    if (hx[k - 1] == px && hy[k - 1] == py) continue;
    while (k >= t && ccl_cross(hx[k - 2], hy[k - 2], hx[k - 1], hy[k - 1], px, py) <= 0) --k;
    hx[k] = px; hy[k] = py; ++k;
  }
  if (k > 1 && hx[k - 1] == hx[0] && hy[k - 1] == hy[0]) {
    // single-pixel top row: the walk started at the contour's first point; it has to come last
    --k;
    const int fx = hx[0], fy = hy[0];
    for (int i = 0; i + 1 < k; ++i) { hx[i] = hx[i + 1]; hy[i] = hy[i + 1]; }
    hx[k - 1] = fx; hy[k - 1] = fy;
  }
  const int n = k;
  BoxRec o;
  o.n_hull = n; o.x0 = hx[0]; o.y0 = hy[0]; o.x1 = n > 1 ? hx[1] : hx[0]; o.y1 = n > 1 ? hy[1] : hy[0];
  o.cx = o.cy = o.w = o.h = o.ax = o.ay = 0.f;
  if (n > 2) {
    // (3) rotating calipers (cv2 rotatingCalipers, CALIPERS_MINAREARECT), float32, host operation order
    float minarea = 3.402823466e+38f;
    int left = 0, bottom = 0, right = 0, top = 0;
    float orientation = 0.f, base_a, base_b = 0.f;
    float p0x = (float)hx[0], p0y = (float)hy[0];
    float left_x = p0x, right_x = p0x, top_y = p0y, bottom_y = p0y;
    for (int i = 0; i < n; ++i) {
      if (p0x < left_x) { left_x = p0x; left = i; }
      if (p0x > right_x) { right_x = p0x; right = i; }
      if (p0y > top_y) { top_y = p0y; top = i; }
      if (p0y < bottom_y) { bottom_y = p0y; bottom = i; }
      const int j = (i + 1 < n) ? i + 1 : 0;
      const float qx = (float)hx[j], qy = (float)hy[j];
      const double dx = (double)qx - (double)p0x, dy = (double)qy - (double)p0y;
      vx[i] = (float)dx; vy[i] = (float)dy;
      inv[i] = __double2float_rn(__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)))));
      p0x = qx; p0y = qy;
    }
    {
      double ax = vx[n - 1], ay = vy[n - 1];
      for (int i = 0; i < n; ++i) {
        const double bx = vx[i], by = vy[i];
        const double convexity = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
        if (convexity != 0) { orientation = convexity > 0 ? 1.f : -1.f; break; }
        ax = bx; ay = by;
      }
    }
    base_a = orientation;
    int seq[4] = {bottom, right, top, left};
    int best_left = 0, best_bottom = 0;
    float best_a = 1.f, best_b = 0.f, best_w = 0.f, best_h = 0.f;
    for (int it = 0; it < n; ++it) {
      float dp[4];
      dp[0] = __fadd_rn(__fmul_rn(base_a, vx[seq[0]]), __fmul_rn(base_b, vy[seq[0]]));
      dp[1] = __fadd_rn(__fmul_rn(-base_b, vx[seq[1]]), __fmul_rn(base_a, vy[seq[1]]));
      dp[2] = __fsub_rn(__fmul_rn(-base_a, vx[seq[2]]), __fmul_rn(base_b, vy[seq[2]]));
      dp[3] = __fsub_rn(__fmul_rn(base_b, vx[seq[3]]), __fmul_rn(base_a, vy[seq[3]]));
      float maxcos = __fmul_rn(dp[0], inv[seq[0]]);
      int main_element = 0;
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        const float cosalpha = __fmul_rn(dp[i], inv[seq[i]]);
        if (cosalpha > maxcos) { main_element = i; maxcos = cosalpha; }
      }
      {
        const int pindex = seq[main_element];
        const float lead_x = __fmul_rn(vx[pindex], inv[pindex]);
        const float lead_y = __fmul_rn(vy[pindex], inv[pindex]);
        switch (main_element) {
          case 0: base_a = lead_x; base_b = lead_y; break;
          case 1: base_a = lead_y; base_b = -lead_x; break;
          case 2: base_a = -lead_x; base_b = -lead_y; break;
          default: base_a = -lead_y; base_b = lead_x; break;
        }
      }
      seq[main_element] += 1;
      if (seq[main_element] == n) seq[main_element] = 0;
      float dx = __fsub_rn((float)hx[seq[1]], (float)hx[seq[3]]);
      float dy = __fsub_rn((float)hy[seq[1]], (float)hy[seq[3]]);
      const float width = __fadd_rn(__fmul_rn(dx, base_a), __fmul_rn(dy, base_b));
      dx = __fsub_rn((float)hx[seq[2]], (float)hx[seq[0]]);
      dy = __fsub_rn((float)hy[seq[2]], (float)hy[seq[0]]);
      const float height = __fadd_rn(__fmul_rn(-dx, base_b), __fmul_rn(dy, base_a));
      const float area = __fmul_rn(width, height);
      if (area <= minarea) {
        minarea = area;
        best_left = seq[3]; best_a = base_a; best_w = width; best_b = base_b; best_h = height;
        best_bottom = seq[0];
      }
    }
    const float A1 = best_a, B1 = best_b, A2 = -best_b, B2 = best_a;
    const float C1 = __fadd_rn(__fmul_rn(A1, (float)hx[best_left]), __fmul_rn((float)hy[best_left], B1));
    const float C2 = __fadd_rn(__fmul_rn(A2, (float)hx[best_bottom]), __fmul_rn((float)hy[best_bottom], B2));
    const float idet = __fdiv_rn(1.f, __fsub_rn(__fmul_rn(A1, B2), __fmul_rn(A2, B1)));
    const float o0x = __fmul_rn(__fsub_rn(__fmul_rn(C1, B2), __fmul_rn(C2, B1)), idet);
    const float o0y = __fmul_rn(__fsub_rn(__fmul_rn(A1, C2), __fmul_rn(A2, C1)), idet);
    const float o1x = __fmul_rn(A1, best_w), o1y = __fmul_rn(B1, best_w);
    const float o2x = __fmul_rn(A2, best_h), o2y = __fmul_rn(B2, best_h);
    o.cx = __fadd_rn(o0x, __fmul_rn(__fadd_rn(o1x, o2x), 0.5f));
    o.cy = __fadd_rn(o0y, __fmul_rn(__fadd_rn(o1y, o2y), 0.5f));
    o.w = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn((double)o1x, (double)o1x), __dmul_rn((double)o1y, (double)o1y))));
    o.h = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn((double)o2x, (double)o2x), __dmul_rn((double)o2y, (double)o2y))));
    o.ax = o1x; o.ay = o1y;
  }
  boxes[comp] = o;
}

// ------------------------------------------------------------------------------------------------
// Whole-image variant: maps of up to 65,536 pixels (a 1024x1024 input's 256x256 map) are labelled by ONE
// CTA with everything in shared memory - 16-bit parents (the raster index fits), the mask bytes, one bit per
// pixel for "outer background root" and "filled" - so the eight passes of the algorithm above are separated
// by __syncthreads instead of kernel launches and no tile-border merge exists.  The kernel also ranks the
// roots, accumulates the component records and counts the kept ones; compaction and the rectangles follow
// (3 launches per batch instead of 12).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf16_find(const uint16_t* par, int i) {
  int p;
  while ((p = *((volatile const uint16_t*)(par + i))) != i) i = p;
  return i;
}
// find with path halving: a non-root is re-pointed at its grandparent with a plain store.  Only non-roots are written
// and only with one of their ancestors, so the race with the 32-bit CAS of atomic_min16 on the neighbouring half-word
// is benign (a lost or repeated store leaves a valid, possibly longer, chain).
__device__ __forceinline__ int uf16_find_halve(uint16_t* par, int i) {
  while (true) {
    const int p = *((volatile uint16_t*)(par + i));
    if (p == i) return i;
    const int g = *((volatile uint16_t*)(par + p));
    if (g == p) return p;
    *((volatile uint16_t*)(par + i)) = (uint16_t)g;
    i = g;
  }
}
__device__ __forceinline__ int atomic_min16(uint16_t* addr, int val) {
  unsigned* wp = reinterpret_cast<unsigned*>(reinterpret_cast<uintptr_t>(addr) & ~(uintptr_t)3);
  const unsigned shift = (reinterpret_cast<uintptr_t>(addr) & 2) ? 16u : 0u;
  unsigned cur = *((volatile unsigned*)wp);
  while (true) {
    const int old = (int)((cur >> shift) & 0xFFFFu);
    if (old <= val) return old;
    const unsigned nw = (cur & ~(0xFFFFu << shift)) | ((unsigned)val << shift);
    const unsigned prev = atomicCAS(wp, cur, nw);
    if (prev == cur) return old;
    cur = prev;
  }
}
__device__ __forceinline__ void uf16_union(uint16_t* par, int a, int b) {
  while (true) {
    a = uf16_find_halve(par, a);
    b = uf16_find_halve(par, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    const int old = atomic_min16(par + a, b);
    if (old == a) return;
    a = old;
  }
}

constexpr int CCL_IMG_THREADS = 1024;
constexpr int CCL_IMG_MAX_PX = 65536;
static inline size_t ccl_image_smem(int hw) {
  const size_t words = ((size_t)hw + 31) / 32;
  return (((size_t)hw * 2 + 15) & ~(size_t)15) + (((size_t)hw + 15) & ~(size_t)15) + 2 * words * 4 + 512;
}

__global__ void __launch_bounds__(CCL_IMG_THREADS, 1)
ccl_image_kernel(const uint8_t* __restrict__ mask, int* __restrict__ labels, int* __restrict__ slot_of,
                 CompRec* __restrict__ comps, const float* __restrict__ cls_logits, int cls_stride,
                 unsigned long long* __restrict__ cls_sums, int n_cls, int* __restrict__ n_comps,
                 int* __restrict__ kept_count, CclTotals* __restrict__ totals, int h, int w, int max_comps, int min_area_x2) {
  extern __shared__ __align__(16) uint8_t ccl_smem[];
  const int hw = h * w;
  const int words = (hw + 31) / 32;
  uint16_t* par = reinterpret_cast<uint16_t*>(ccl_smem);
  uint8_t* m = ccl_smem + (((size_t)hw * 2 + 15) & ~(size_t)15);
  unsigned* outer = reinterpret_cast<unsigned*>(m + (((size_t)hw + 15) & ~(size_t)15));
  unsigned* filled = outer + words;
  int* scr = reinterpret_cast<int*>(filled + words);            // 128 ints of scan scratch
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint8_t* gm = mask + (size_t)n * hw;
  // ---- A: mask -> shared, parent = start of the pixel's row run (warp per row, ballot per 32 columns)
  for (int y = wid; y < h; y += CCL_IMG_THREADS / 32) {
    int carry_cls = 3, carry_start = 0;
    for (int x0 = 0; x0 < w; x0 += 32) {
      const int x = x0 + lane;
      const bool in = x < w;
      const int c = in ? (gm[(size_t)y * w + x] != 0 ? 1 : 0) : 2;
      const unsigned fgb = __ballot_sync(0xffffffffu, c == 1);
      const unsigned inb = __ballot_sync(0xffffffffu, in);
      const unsigned same = c == 1 ? fgb : (c == 0 ? (~fgb & inb) : ~inb);
      const unsigned below = (~same) & ((1u << lane) - 1u);
      const int start = below ? x0 + 32 - __clz(below) : (c == carry_cls ? carry_start : x0);
      if (in) { m[y * w + x] = (uint8_t)c; par[y * w + x] = (uint16_t)(y * w + start); }
      carry_cls = __shfl_sync(0xffffffffu, c, 31);
      carry_start = __shfl_sync(0xffffffffu, start, 31);
    }
  }
  for (int i = tid; i < words; i += CCL_IMG_THREADS) { outer[i] = 0u; filled[i] = 0u; }
  __syncthreads();
  // ---- B: vertical / diagonal links (foreground 8-connected, background 4-connected), one link per run overlap
  for (int y = 1 + wid; y < h; y += CCL_IMG_THREADS / 32) {
    for (int x = lane; x < w; x += 32) {
      const int p = y * w + x;
      const int c = m[p];
      const bool hasW = x > 0, hasE = x < w - 1;
      const bool N_ = m[p - w] == c;
      if (c == 1) {
        if (N_) {
          if (!hasW || m[p - 1] != 1 || m[p - w - 1] != 1) uf16_union(par, p, p - w);
        } else {
          if (hasE && m[p - w + 1] == 1) uf16_union(par, p, p - w + 1);
          if (hasW && m[p - w - 1] == 1) uf16_union(par, p, p - w - 1);
        }
      } else if (N_ && (!hasW || m[p - 1] != 0 || m[p - w - 1] != 0)) {
        uf16_union(par, p, p - w);
      }
    }
  }
  __syncthreads();
  // ---- C: full path compression
  for (int i = tid; i < hw; i += CCL_IMG_THREADS) par[i] = (uint16_t)uf16_find_halve(par, i);
  __syncthreads();
  // ---- D: background sets that own an image-border pixel are connected to the outside
  for (int i = tid; i < 2 * (w + h); i += CCL_IMG_THREADS) {
    int y, x;
    if (i < w) { y = 0; x = i; }
    else if (i < 2 * w) { y = h - 1; x = i - w; }
    else if (i < 2 * w + h) { y = i - 2 * w; x = 0; }
    else { y = i - 2 * w - h; x = w - 1; }
    const int p = y * w + x;
    if (m[p] == 0) { const int r = par[p]; atomicOr(&outer[r >> 5], 1u << (r & 31)); }
  }
  __syncthreads();
  // ---- E: 8-connectivity over filled pixels (foreground + holes); see ccl_merge2_kernel for the invariants
  {
    auto is_filled = [&](int q) -> bool {
      if (m[q] != 0) return true;
      const int r = *((volatile const uint16_t*)(par + q));
      return ((outer[r >> 5] >> (r & 31)) & 1u) == 0u;
    };
    for (int y = wid; y < h; y += CCL_IMG_THREADS / 32) {
      for (int x = lane; x < w; x += 32) {
        const int p = y * w + x;
        const bool fg = m[p] != 0;
        if (!fg && !is_filled(p)) continue;
        const bool hasW = x > 0, hasN = y > 0, hasE = x < w - 1;
        auto link = [&](int q) { if (!(fg && m[q] != 0)) uf16_union(par, p, q); };
        if (hasN && is_filled(p - w)) {
          link(p - w);
        } else {
          if (hasN && hasE && is_filled(p - w + 1)) link(p - w + 1);
          if (hasN && hasW && is_filled(p - w - 1)) link(p - w - 1);
          else if (hasW && is_filled(p - 1)) link(p - 1);
        }
      }
    }
  }
  __syncthreads();
  // ---- F: labels (root = first raster pixel of the filled component) -> global; filled bits
  int* lab = labels + (size_t)n * hw;
  for (int i0 = 0; i0 < hw; i0 += CCL_IMG_THREADS) {
    const int i = i0 + tid;
    bool f = false;
    int root = -1;
    if (i < hw) {
      const int pr = par[i];
      f = m[i] != 0 || ((outer[pr >> 5] >> (pr & 31)) & 1u) == 0u;      // an outer background pixel still points at its phase-1 root
      if (f) root = uf16_find_halve(par, i);
      lab[i] = root;
    }
    const unsigned fb = __ballot_sync(0xffffffffu, f);
    if (lane == 0 && i0 + wid * 32 < hw) filled[(i0 >> 5) + wid] = fb;
    if (f) par[i] = (uint16_t)root;       // compression while others still walk through i is safe: parents only move rootwards
  }
  __syncthreads();
  // ---- G: rank the roots in raster order -> slot_of[root], records; thread = one run of `chunk` consecutive pixels
  int* so = slot_of + (size_t)n * hw;
  CompRec* cr = comps + (size_t)n * max_comps;
  const int chunk = (hw + CCL_IMG_THREADS - 1) / CCL_IMG_THREADS;       // <= 64
  const int c0 = tid * chunk;
  unsigned long long rootbits = 0ull;
  for (int k = 0; k < chunk; ++k) {
    const int j = (k + lane) % chunk;                                   // rotated start: conflict-free 16-bit reads
    const int i = c0 + j;
    if (i < hw && par[i] == i && ((filled[i >> 5] >> (i & 31)) & 1u)) rootbits |= 1ull << j;
  }
  const int cnt = __popcll(rootbits);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) scr[1 + wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int own = scr[1 + lane];
    int v = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    scr[33 + lane] = v - own;
    if (lane == 31) scr[0] = v;
  }
  __syncthreads();
  const int n_roots = scr[0];
  {
    int slot = scr[33 + wid] + incl - cnt;
    unsigned long long b = rootbits;
    while (b) {
      const int j = __ffsll((long long)b) - 1;
      b &= b - 1;
      const int i = c0 + j;
      so[i] = slot;
      if (slot < max_comps) {
        CompRec r; r.label = i; r.xmin = w; r.ymin = h; r.xmax = -1; r.ymax = -1;
        r.n_pixels = 0; r.n_filled = 0; r.q3 = 0; r.q4 = 0;
        cr[slot] = r;
        for (int c = 0; c < n_cls; ++c) cls_sums[((size_t)n * max_comps + slot) * n_cls + c] = 0ull;
      }
      ++slot;
    }
  }
  __threadfence_block();
  __syncthreads();
  // ---- H: per-component reductions, one 2x2 window per thread (bottom-right pixel (y,x), y in [0,h], x in [0,w])
  auto Lab = [&](int yy, int xx) -> int {
    if (yy < 0 || yy >= h || xx < 0 || xx >= w) return -1;
    const int q = yy * w + xx;
    return ((filled[q >> 5] >> (q & 31)) & 1u) ? (int)par[q] : -1;
  };
  for (int y = wid; y <= h; y += CCL_IMG_THREADS / 32) {
    for (int x0 = 0; x0 <= w; x0 += 32) {
      const int x = x0 + lane;
      int slot = -1, l00 = -1, q3 = 0, q4 = 0;
      if (x <= w) {
        const int a = Lab(y - 1, x - 1), b = Lab(y - 1, x), c = Lab(y, x - 1);
        l00 = Lab(y, x);
        const int cntw = (a >= 0) + (b >= 0) + (c >= 0) + (l00 >= 0);
        const int any = max(max(a, b), max(c, l00));          // all non-negative ones are equal
        if (any >= 0) slot = so[any];
        q3 = cntw == 3; q4 = cntw == 4;
      }
      if (slot >= max_comps) slot = -1;                       // overflow: reported via n_comps > max_comps
      const unsigned active = __ballot_sync(0xffffffffu, slot >= 0);
      if (slot < 0) continue;
      const unsigned grp = __match_any_sync(active, slot);
      const int leader = __ffs(grp) - 1;
      const bool own = l00 >= 0;
      const int sq3 = __reduce_add_sync(grp, q3), sq4 = __reduce_add_sync(grp, q4);
      const int sfill = __reduce_add_sync(grp, own ? 1 : 0);
      const int spix = __reduce_add_sync(grp, (own && m[y * w + x] != 0) ? 1 : 0);
      const int xmn = __reduce_min_sync(grp, own ? x : 0x7fffffff);
      const int xmx = __reduce_max_sync(grp, own ? x : -1);
      if (lane == leader) {
        CompRec* r = cr + slot;
        if (sq3) atomicAdd(&r->q3, sq3);
        if (sq4) atomicAdd(&r->q4, sq4);
        if (sfill) {
          atomicAdd(&r->n_filled, sfill);
          if (spix) atomicAdd(&r->n_pixels, spix);
          atomicMin(&r->xmin, xmn); atomicMin(&r->ymin, y);
          atomicMax(&r->xmax, xmx); atomicMax(&r->ymax, y);
        }
      }
      if (n_cls > 0) {
        float e[UBD_MAX_CLASSES];
        float s = 1.f;
        if (own) {
          const float* lg = cls_logits + ((size_t)n * hw + y * w + x) * cls_stride;
          float mx = lg[0];
          for (int c = 1; c < n_cls; ++c) mx = fmaxf(mx, lg[c]);
          s = 0.f;
          for (int c = 0; c < n_cls; ++c) { e[c] = expf(lg[c] - mx); s += e[c]; }
        }
        for (int c = 0; c < n_cls; ++c) {
          const unsigned fx = own ? (unsigned)(e[c] / s * 16777216.0f) : 0u;
          const unsigned tot = __reduce_add_sync(grp, fx);
          if (lane == leader && tot)
            atomicAdd(&cls_sums[((size_t)n * max_comps + slot) * n_cls + c], (unsigned long long)tot);
        }
      }
    }
  }
  __threadfence();
  __syncthreads();
  // ---- I: kept components of this image (2*contourArea > min_area_x2, utils.py:55)
  {
    const int cntc = min(n_roots, max_comps);
    int k = 0;
    for (int s = tid; s < cntc; s += CCL_IMG_THREADS) {
      const volatile CompRec* r = cr + s;
      k += (2 * r->q4 + r->q3) > min_area_x2;
    }
    k = __reduce_add_sync(0xffffffffu, k);
    if (lane == 0) scr[1 + wid] = k;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int i = 0; i < CCL_IMG_THREADS / 32; ++i) t += scr[1 + i];
      n_comps[n] = n_roots;
      kept_count[n] = t;
      atomicAdd(&totals->total_kept, t);
      atomicMax(&totals->max_ncomp, n_roots);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Run-length variant of the whole-image kernel ("rle"): the same algorithm on ROW RUNS instead of pixels.
// A row is a few 32-bit words of foreground bits; a run = maximal interval of one class (foreground or
// background) in a row, identified by its rank in raster order: run id of pixel (y,x) = (number of run starts before
// its word) + popc(start bits of the word up to x) - 1, so runs need no table of their own - only a 16-bit parent
// per run (at most h*w <= 65,536 runs).  Links between rows are found with bit operations on the words (a run
// touches the runs above it whose bits intersect its window, widened by one pixel for 8-connectivity); the
// filled mask, the bit-quad counts of 2*contourArea and the per-run pixel counts are popcounts.  One CTA per
// image, every pass separated by __syncthreads; typical masks have a few thousand runs, so a pass is a handful
// of iterations per thread.
// ------------------------------------------------------------------------------------------------
static inline size_t ccl_rle_smem(int h, int w) {
  const size_t nw = (size_t)h * ((w + 31) / 32), hw = (size_t)h * w;
  return 3 * nw * 4 + ((nw + 1) * 2 + 15 & ~(size_t)15) + ((hw + 31) / 32) * 4 + ((hw * 2 + 15) & ~(size_t)15) + 1024;
}

__global__ void __launch_bounds__(CCL_IMG_THREADS, 1)
ccl_rle_kernel(const uint8_t* __restrict__ mask, int* __restrict__ labels, int* __restrict__ slot_of, int* __restrict__ run_label,
               CompRec* __restrict__ comps, const float* __restrict__ cls_logits, int cls_stride,
               unsigned long long* __restrict__ cls_sums, int n_cls, int* __restrict__ n_comps,
               int* __restrict__ kept_count, CclTotals* __restrict__ totals, int h, int w, int max_comps, int min_area_x2) {
  extern __shared__ __align__(16) uint8_t ccl_smem[];
  const int hw = h * w, WPR = (w + 31) >> 5, NW = h * WPR;
  unsigned* F = reinterpret_cast<unsigned*>(ccl_smem);            // foreground bits
  unsigned* X = F + NW;                                           // filled bits (foreground + holes)
  unsigned* S = X + NW;                                           // run-start bits
  uint16_t* WB = reinterpret_cast<uint16_t*>(S + NW);             // runs that start before the word (raster order)
  unsigned* outer = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(WB) + (((size_t)(NW + 1) * 2 + 15) & ~(size_t)15));
  uint16_t* par = reinterpret_cast<uint16_t*>(outer + ((hw + 31) >> 5));
  int* scr = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(par) + (((size_t)hw * 2 + 15) & ~(size_t)15));   // 256 ints
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NWARP = CCL_IMG_THREADS / 32;
  const uint8_t* gm = mask + (size_t)n * hw;
  const unsigned vlast = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xFFFFFFFFu;      // valid bits of a row's last word
  auto valid = [&](int wd) -> unsigned { return wd == WPR - 1 ? vlast : 0xFFFFFFFFu; };
  // run id of the pixel at bit `bit` of word `wi`
  auto runid = [&](int wi, int bit) -> int { return (int)WB[wi] + __popc(S[wi] & (0xFFFFFFFFu >> (31 - bit))) - 1; };
  // block-wide exclusive scan of one int per thread; returns the exclusive prefix, total in scr[0]
  auto block_scan = [&](int v) -> int {
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();                                              // scr is reused between scans
    if (lane == 31) scr[1 + wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int own = scr[1 + lane];
      int s = own;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
      scr[33 + lane] = s - own;
      if (lane == 31) scr[0] = s;
    }
    __syncthreads();
    return scr[33 + wid] + incl - v;
  };

  // ---- A: foreground words and run-start bits, one warp per row
  for (int y = wid; y < h; y += NWARP) {
    unsigned carry = 0u;
    for (int wd = 0; wd < WPR; ++wd) {
      const int x = wd * 32 + lane;
      const bool fg = x < w && gm[(size_t)y * w + x] != 0;
      const unsigned bits = __ballot_sync(0xffffffffu, fg);
      if (lane == 0) {
        const unsigned diff = (bits ^ ((bits << 1) | carry)) | (wd == 0 ? 1u : 0u);
        F[y * WPR + wd] = bits;
        S[y * WPR + wd] = diff & valid(wd);
      }
      carry = bits >> 31;
    }
  }
  for (int i = tid; i < ((hw + 31) >> 5); i += CCL_IMG_THREADS) outer[i] = 0u;
  __syncthreads();
  // run-id bases: exclusive scan of the per-word start counts in raster order (thread = `chunk` consecutive words)
  const int chunk = (NW + CCL_IMG_THREADS - 1) / CCL_IMG_THREADS;
  int n_runs;
  {
    int c = 0;
    for (int k = 0; k < chunk; ++k) { const int wi = tid * chunk + k; if (wi < NW) c += __popc(S[wi]); }
    int base = block_scan(c);
    n_runs = scr[0];
    for (int k = 0; k < chunk; ++k) { const int wi = tid * chunk + k; if (wi < NW) { WB[wi] = (uint16_t)base; base += __popc(S[wi]); } }
  }
  for (int r = tid; r < n_runs; r += CCL_IMG_THREADS) par[r] = (uint16_t)r;
  __syncthreads();
  // contiguous run of set bits of m that starts at its lowest set bit
  auto low_run = [](unsigned m) -> unsigned { return m & ~(m + (m & (0u - m))); };

  // ---- B: links to the row above: foreground 8-connected, background 4-connected
  for (int wi = WPR + tid; wi < NW; wi += CCL_IMG_THREADS) {
    const int wd = wi % WPR, up = wi - WPR;
    const unsigned v = valid(wd);
    const unsigned cur = F[wi], above = F[up];
    unsigned m = cur & v;
    while (m) {                                                   // foreground pieces of this word
      const unsigned rm = low_run(m);
      m &= ~rm;
      const int a = __ffs(rm) - 1, b = 31 - __clz(rm);
      const int r = runid(wi, a);
      unsigned t = above & (rm | (rm << 1) | (rm >> 1));
      while (t) { const unsigned seg = low_run(t); t &= ~seg; uf16_union(par, r, runid(up, __ffs(seg) - 1)); }
      if (a == 0 && wd > 0 && (F[up - 1] >> 31)) uf16_union(par, r, runid(up - 1, 31));
      if (b == 31 && wd + 1 < WPR && (F[up + 1] & 1u)) uf16_union(par, r, runid(up + 1, 0));
    }
    m = ~cur & v;
    while (m) {                                                   // background pieces
      const unsigned rm = low_run(m);
      m &= ~rm;
      const int r = runid(wi, __ffs(rm) - 1);
      unsigned t = ~above & rm;
      while (t) { const unsigned seg = low_run(t); t &= ~seg; uf16_union(par, r, runid(up, __ffs(seg) - 1)); }
    }
  }
  __syncthreads();
  // ---- C: flatten
  for (int r = tid; r < n_runs; r += CCL_IMG_THREADS) par[r] = (uint16_t)uf16_find_halve(par, r);
  __syncthreads();
  // ---- D: background sets that own an image-border pixel are connected to the outside
  {
    auto mark = [&](int r) { const int root = par[r]; atomicOr(&outer[root >> 5], 1u << (root & 31)); };
    for (int i = tid; i < 2 * WPR; i += CCL_IMG_THREADS) {        // first and last row: every background piece
      const int wd = i % WPR, wi = (i < WPR ? 0 : (h - 1) * WPR) + wd;
      unsigned m = ~F[wi] & valid(wd);
      while (m) { const unsigned rm = low_run(m); m &= ~rm; mark(runid(wi, __ffs(rm) - 1)); }
    }
    for (int y = tid; y < h; y += CCL_IMG_THREADS) {              // first and last pixel of every row
      if (!(F[y * WPR] & 1u)) mark(runid(y * WPR, 0));
      const int lw = y * WPR + WPR - 1, lb = (w - 1) & 31;
      if (!((F[lw] >> lb) & 1u)) mark(runid(lw, lb));
    }
  }
  __syncthreads();
  // ---- E0: filled bits = foreground + background runs whose set is not outer
  for (int wi = tid; wi < NW; wi += CCL_IMG_THREADS) {
    const unsigned cur = F[wi];
    unsigned x = cur, m = ~cur & valid(wi % WPR);
    while (m) {
      const unsigned rm = low_run(m);
      m &= ~rm;
      const int root = par[runid(wi, __ffs(rm) - 1)];
      if (!((outer[root >> 5] >> (root & 31)) & 1u)) x |= rm;
    }
    X[wi] = x;
  }
  __syncthreads();
  // ---- E1: 8-connectivity over filled pixels: neighbouring runs of a row, then the row above
  for (int wi = tid; wi < NW; wi += CCL_IMG_THREADS) {
    const int wd = wi % WPR;
    const unsigned xc = X[wi];
    {   // a run that starts at column > 0 and its left neighbour, both filled
      unsigned st = S[wi] & xc;
      if (wd == 0) st &= ~1u;
      const unsigned left = (xc << 1) | (wd > 0 ? (X[wi - 1] >> 31) : 0u);
      st &= left;
      while (st) { const int b = __ffs(st) - 1; st &= st - 1; const int r = runid(wi, b); uf16_union(par, r, r - 1); }
    }
    if (wi >= WPR) {
      const int up = wi - WPR;
      const unsigned xa = X[up];
      unsigned m = xc;
      while (m) {
        const unsigned rm = low_run(m);
        m &= ~rm;
        const int a = __ffs(rm) - 1, b = 31 - __clz(rm);
        const int r = runid(wi, a);
        unsigned t = xa & (rm | (rm << 1) | (rm >> 1));
        while (t) { const unsigned seg = low_run(t); t &= ~seg; uf16_union(par, r, runid(up, __ffs(seg) - 1)); }
        if (a == 0 && wd > 0 && (X[up - 1] >> 31)) uf16_union(par, r, runid(up - 1, 31));
        if (b == 31 && wd + 1 < WPR && (X[up + 1] & 1u)) uf16_union(par, r, runid(up + 1, 0));
      }
    }
  }
  __syncthreads();
  // ---- F: flatten
  for (int r = tid; r < n_runs; r += CCL_IMG_THREADS) par[r] = (uint16_t)uf16_find_halve(par, r);
  __syncthreads();
  // ---- G: rank the root runs in raster order -> slot, label (= raster index of the run's first pixel), records
  int* so = slot_of + (size_t)n * hw;                 // slot by LABEL (raster index of the first pixel): what the later kernels read
  int* rs = run_label + (size_t)n * 2 * hw;           // slot by run id of the root run (this kernel's own look-ups)
  int* rl = rs + hw;                                  // label by run id of the root run
  CompRec* cr = comps + (size_t)n * max_comps;
  int n_roots;
  {
    auto roots_of = [&](int wi) -> unsigned {                     // start bits of this word whose run is a filled root
      unsigned st = S[wi] & X[wi], out = 0u;
      while (st) { const int b = __ffs(st) - 1; st &= st - 1; const int r = runid(wi, b); if (par[r] == r) out |= 1u << b; }
      return out;
    };
    int c = 0;
    for (int k = 0; k < chunk; ++k) { const int wi = tid * chunk + k; if (wi < NW) c += __popc(roots_of(wi)); }
    int slot = block_scan(c);
    n_roots = scr[0];
    for (int k = 0; k < chunk; ++k) {
      const int wi = tid * chunk + k;
      if (wi >= NW) break;
      unsigned rb = roots_of(wi);
      while (rb) {
        const int b = __ffs(rb) - 1; rb &= rb - 1;
        const int r = runid(wi, b);
        const int label = (wi / WPR) * w + (wi % WPR) * 32 + b;
        so[label] = slot; rs[r] = slot; rl[r] = label;
        if (slot < max_comps) {
          CompRec rec; rec.label = label; rec.xmin = w; rec.ymin = h; rec.xmax = -1; rec.ymax = -1;
          rec.n_pixels = 0; rec.n_filled = 0; rec.q3 = 0; rec.q4 = 0;
          cr[slot] = rec;
          for (int q = 0; q < n_cls; ++q) cls_sums[((size_t)n * max_comps + slot) * n_cls + q] = 0ull;
        }
        ++slot;
      }
    }
  }
  __threadfence_block();
  __syncthreads();
  // ---- H: per-run reductions.  Thread = word; every filled run that STARTS in the word is followed to its end.
  //   2x2 windows are attributed by their bottom row: a window whose bottom-right pixel x is filled belongs to x's run
  //   (BL = filled bit left of x, TL / TR the two bits above); a window whose bottom-right pixel is empty but whose
  //   bottom-left pixel ends a run can only reach 3 (both pixels above filled).
  for (int wi = tid; wi < NW; wi += CCL_IMG_THREADS) {
    const int y = wi / WPR, wd0 = wi - y * WPR;
    unsigned st = S[wi] & X[wi];
    while (st) {
      const int b0 = __ffs(st) - 1; st &= st - 1;
      const int r = runid(wi, b0);
      const int slot = rs[par[r]];
      if (slot >= max_comps) continue;
      const bool fgrun = (F[wi] >> b0) & 1u;
      int q3 = 0, q4 = 0, len = 0, xend = 0;
      // walk the words of the run
      int wd = wd0, lo = b0;
      while (true) {
        const int cw = y * WPR + wd;
        const unsigned above_lo = lo == 31 ? 0u : (0xFFFFFFFFu << (lo + 1));
        const unsigned nxt = S[cw] & above_lo & (wd == wd0 ? 0xFFFFFFFFu : 0xFFFFFFFFu);
        unsigned stops = wd == wd0 ? nxt : S[cw];                 // the run ends before the next start bit
        const unsigned v = valid(wd);
        int hi;
        bool ended;
        if (stops) { hi = __ffs(stops) - 2; ended = true; }
        else { hi = 31 - __clz(v); ended = wd == WPR - 1; }
        if (hi >= lo) {
          const unsigned M = (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
          const unsigned xs = (X[cw] << 1) | (wd > 0 ? (X[cw - 1] >> 31) : 0u);
          const unsigned T = y > 0 ? X[cw - WPR] : 0u;
          const unsigned ts = y > 0 ? ((T << 1) | (wd > 0 ? (X[cw - WPR - 1] >> 31) : 0u)) : 0u;
          q4 += __popc(M & xs & ts & T);
          q3 += __popc(M & ((xs & ts & ~T) | (xs & ~ts & T) | (~xs & ts & T)));
          len += hi - lo + 1;
          xend = wd * 32 + hi;
        }
        if (ended) break;
        ++wd; lo = 0;
      }
      // window right of the run's last pixel (bottom-right empty, bottom-left = last pixel): 3 iff both pixels above are filled
      if (y > 0) {
        const int xr = xend + 1;
        const bool br = xr < w && ((X[y * WPR + (xr >> 5)] >> (xr & 31)) & 1u);
        if (!br) {
          const bool tl = (X[(y - 1) * WPR + (xend >> 5)] >> (xend & 31)) & 1u;
          const bool tr = xr < w && ((X[(y - 1) * WPR + (xr >> 5)] >> (xr & 31)) & 1u);
          q3 += tl && tr;
        }
      }
      CompRec* rec = cr + slot;
      if (q3) atomicAdd(&rec->q3, q3);
      if (q4) atomicAdd(&rec->q4, q4);
      atomicAdd(&rec->n_filled, len);
      if (fgrun) atomicAdd(&rec->n_pixels, len);
      atomicMin(&rec->xmin, wd0 * 32 + b0); atomicMin(&rec->ymin, y);
      atomicMax(&rec->xmax, xend); atomicMax(&rec->ymax, y);
    }
  }
  // ---- I: labels to global memory (and the class vote sums), one warp per row, lane = pixel
  int* lab = labels + (size_t)n * hw;
  for (int y = wid; y < h; y += NWARP) {
    for (int wd = 0; wd < WPR; ++wd) {
      const int x = wd * 32 + lane, wi = y * WPR + wd;
      const bool own = x < w && ((X[wi] >> lane) & 1u);
      int slot = -1;
      if (x < w) {
        int l = -1;
        if (own) { const int root = par[runid(wi, lane)]; l = rl[root]; slot = rs[root]; }
        lab[(size_t)y * w + x] = l;
      }
      if (n_cls > 0) {
        if (slot >= max_comps) slot = -1;
        const unsigned active = __ballot_sync(0xffffffffu, slot >= 0);
        if (slot >= 0) {
          const unsigned grp = __match_any_sync(active, slot);
          const int leader = __ffs(grp) - 1;
          const float* lg = cls_logits + ((size_t)n * hw + (size_t)y * w + x) * cls_stride;
          float e[UBD_MAX_CLASSES];
          float mx = lg[0];
          for (int c = 1; c < n_cls; ++c) mx = fmaxf(mx, lg[c]);
          float s = 0.f;
          for (int c = 0; c < n_cls; ++c) { e[c] = expf(lg[c] - mx); s += e[c]; }
          for (int c = 0; c < n_cls; ++c) {
            const unsigned fx = (unsigned)(e[c] / s * 16777216.0f);
            const unsigned tot = __reduce_add_sync(grp, fx);
            if (lane == leader && tot) atomicAdd(&cls_sums[((size_t)n * max_comps + slot) * n_cls + c], (unsigned long long)tot);
          }
        }
      }
    }
  }
  __threadfence();
  __syncthreads();
  // ---- J: kept components of this image (2*contourArea > min_area_x2, utils.py:55)
  {
    const int cntc = min(n_roots, max_comps);
    int k = 0;
    for (int s = tid; s < cntc; s += CCL_IMG_THREADS) {
      const volatile CompRec* r = cr + s;
      k += (2 * r->q4 + r->q3) > min_area_x2;
    }
    k = __reduce_add_sync(0xffffffffu, k);
    __syncthreads();
    if (lane == 0) scr[1 + wid] = k;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int i = 0; i < NWARP; ++i) t += scr[1 + i];
      n_comps[n] = n_roots;
      kept_count[n] = t;
      atomicAdd(&totals->total_kept, t);
      atomicMax(&totals->max_ncomp, n_roots);
    }
  }
}
