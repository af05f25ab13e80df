// Shared definitions for libubd (sm_100a).  See include/ubd.h for the ABI and DESIGN.md for layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/ubd.h"

#define UBD_NF 24          // n_filters, net.py:289
#define UBD_NG 6           // 24 channels = 6 planes of 4 fp32 (one 16-byte unit per pixel and plane)
#define UBD_NLAYERS_DIL 6  // net.py:298-304

// Feature maps live in HBM as row-interleaved planes of float4 with x padding:
//   act[n][y][g][xp], g = 0..5 (channels 4g..4g+3), xp = x + pad, row pitch Wp = W + 2*pad.
// One pixel of one plane is 16 bytes = one row of a UMMA core matrix of the K-major no-swizzle
// layout, so a (dy,dx) tap is a pure address offset.  The pad columns are zero (buffers are zeroed
// when (re)shaped and every kernel writes the interior only): they are the conv zero padding in x,
// and they make the six planes of one image row ONE contiguous block that the TMA unit stages with a
// single cp.async.bulk.  Quarter-resolution maps use pad = UBD_MAP_PAD (largest dilation);
// the stem's half-resolution maps use pad = 0.
#define UBD_MAP_PAD 16
__host__ __device__ __forceinline__ size_t act_index(int n, int g, int y, int x, int H, int W, int pad) {
  return (((size_t)n * H + y) * UBD_NG + g) * (size_t)(W + 2 * pad) + pad + x;
}
__host__ __device__ __forceinline__ size_t act_elems(int n, int H, int W, int pad) {
  return (size_t)n * H * UBD_NG * (size_t)(W + 2 * pad);
}

struct WeightSpec {
  // offsets (in floats) of the 23 Keras arrays inside the flat parameter buffer (W1 order)
  int64_t off[UBD_N_WEIGHT_ARRAYS];
  int64_t size[UBD_N_WEIGHT_ARRAYS];
  int64_t total;
  int cin, n_out;
};

static inline WeightSpec make_weight_spec(int grey, int n_classes) {
  WeightSpec s;
  s.cin = grey ? 1 : 3;
  s.n_out = 1 + n_classes;
  int k = 0;
  int64_t o = 0;
  // every array starts on a 128-byte boundary (vector loads / bulk copies); the gaps stay zero
  auto add = [&](int64_t n) { s.off[k] = o; s.size[k] = n; o += (n + 31) & ~(int64_t)31; ++k; };
  int cins[3] = {s.cin, UBD_NF, UBD_NF};
  for (int i = 0; i < 3; ++i) { add(9 * cins[i]); add((int64_t)cins[i] * UBD_NF); add(UBD_NF); }
  for (int i = 0; i < 6; ++i) { add(9 * UBD_NF * UBD_NF); add(UBD_NF); }
  add((int64_t)UBD_NF * s.n_out); add(s.n_out);
  s.total = o;
  return s;
}

// Results of the connected-component stage as the host reads them back (ubd_ccl.cuh writes them).
struct OutRec {
  int image, label, xmin, ymin, xmax, ymax, n_pixels, n_filled, area_x2, class_id, slot;
  int row_base;                            // first entry of the component's rows in the row-extent arrays (GPU rectangles)
};
struct CclTotals { int total_kept, total_pts, max_ncomp, total_rows; };
// min-area rectangle of a kept component as ccl_boxes_kernel leaves it (centre, size, first edge vector, hull size
// and its first two points for the degenerate cases)
struct BoxRec { float cx, cy, w, h, ax, ay; int n_hull, x0, y0, x1, y1; };
struct HullPt { int comp; int xy; };     // comp = index into the compacted output; xy = (y << 16) | x

#define UBD_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
      return UBD_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define UBD_FAIL(code, msg) do { h->err = (msg); return (code); } while (0)

static const int kDilations[UBD_NLAYERS_DIL] = {1, 2, 4, 8, 16, 1};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
