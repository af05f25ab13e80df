// The handle behind ubd_handle: one device, one stream, lazily sized workspaces.
#pragma once
#include <utility>
#include "ubd_common.cuh"

struct DevBuf { void* p = nullptr; size_t cap = 0; };

// Host-visible results of one batch (see ccl_enqueue / ccl_finish in ubd_api.cu): kSlots slots so that a submitted batch
// can be finished on the host while the next ones run and are being copied in.
constexpr int kSlots = 3;
struct ResultSlot {
  DevBuf hdr, out_recs, hull_pts, box_recs;          // device: raw / kept counts + totals, kept records, hull candidates | rectangles
  bool gpu_boxes = true;
  DevBuf d_images, d_mask, d_logits;       // device: staging of a submitted batch and its outputs
  int* h_hdr = nullptr; size_t h_hdr_cap = 0;     // pinned copy of hdr (kept counts + totals)
  cudaEvent_t ev_cc = nullptr, ev_fwd = nullptr, ev_in = nullptr, ev_d2h = nullptr;
  int n = 0, mh = 0, mw = 0, max_pts = 0, max_comps_img = 0, max_out = 0;
  const uint8_t* d_mask_used = nullptr; const float* d_cls_used = nullptr;
  int cls_stride = 0, n_cls = 0, min_area_x2 = 0;
  bool busy = false, d2h_pending = false;
  long long ticket = 0;
  std::vector<OutRec> recs; std::vector<HullPt> pts; std::vector<BoxRec> boxes;      // host scratch, reused
  std::vector<int> row0, ext; std::vector<int32_t> xy;
};

struct ubd_handle_s {
  int device = 0;
  bool grey = true, fml = true;
  int n_classes = 0;
  int precision = UBD_FP32;
  int n_sm = 148;
  WeightSpec spec;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;      // H2D of chunk k+1 overlaps the compute of chunk k
  cudaStream_t d2h_stream = nullptr;       // read-back of results while the next batch computes
  cudaStream_t cc_stream = nullptr;        // connected components of batch k under the network of batch k+1
  bool opt_cc_stream = true;
  cudaStream_t rec_stream = nullptr;       // read-back of a finished batch's records (idle except inside ubd_segment_wait)
  ResultSlot rs[kSlots];
  long long next_ticket = 0;
  std::vector<cudaEvent_t> copy_events;
  // optional CUDA-event profiling of kernel groups (option "profile")
  bool profile = false;
  struct Prof { double ms = 0; int64_t launches = 0; };
  Prof prof_dil, prof_stem, prof_ccl, prof_head;
  double host_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // wall-clock of host phases (option "profile")
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<std::pair<Prof*, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
  std::string err;
  int64_t launches = 0;

  float* d_params = nullptr;      // flat parameter buffer, Keras get_weights() order and layout
  float* d_lut = nullptr;         // uint8 -> mobilenet_like preprocessing table
  bool have_weights = false;
  bool tc_weights_dirty = true;   // tensor-core weight images must be rebuilt from d_params
  bool tc4_weights_dirty = true;  // ... the column-rotating kernel's weight images (ubd_tc4.cuh)
  bool tc4_train_dirty = true;    // ... the six dilated tf32 images the training step needs (rebuilt alone between Adam steps)
  int tc_w16 = -1, tc4_w16 = -1, stem_w16 = -1;   // which 16-bit container (UBD_BF16 / UBD_F16) the built images hold
  int opt_pipeline = 0;           // dilated stack: 1 = one layer-pipelined launch with L2 ring buffers for chunks of >= 24 images
                                  // (measured 0.84 ms vs 0.83 ms per 64 x 1024^2 for one launch per layer, profiles/r02_summary.md)
  int opt_pipe_ring = 3;          // maps per layer boundary in that launch
  DevBuf pipe_ring, pipe_flags;
  long long pipe_tag = 0;
  int opt_tc_variant = 1;         // dilated layers: 1 = ubd_tc4.cuh, 0 = ubd_tc.cuh (see launch_dil_tc)

  int opt_stem_variant = 0;       // grey input: 2 = fused separable kernel (ubd_stemf.cuh), 1 = two-kernel paths, 0 = auto (stem_is_fused)
  int opt_dense_l2 = 1;           // stem L2 as dense tensor-core conv for grey uint8 input (0: FP32-pipe depthwise path)
  int opt_stem_chunk = 0;         // images per stem launch (0 = auto)
  int opt_chunk = 0;              // images per L2-resident chunk (0 = auto)
  int opt_max_comps = 4096;       // component slots per image
  int opt_max_points = 0;         // hull candidate capacity (0 = auto)
  int opt_fused_ccl = 0;          // 1: maps of <= 65,536 px are labelled by one CTA per image in shared memory (measured slower at batch 64:
                                  // 0.79 vs 0.36 ms, 64 CTAs on 148 SMs and divergent pixel-level finds; profiles/r02_summary.md)
  int opt_gpu_boxes = 1;          // min-area rectangles on the GPU (0: hull candidates to the host, ubd_min_area_box)

  // inference workspaces
  DevBuf d_images, d_logits, d_mask;
  DevBuf act1, act2, mapA, mapB, mapC;     // mapC: fp32 output of the last layer in bf16 mode
  int map_h = 0, map_w = 0, map_n = 0, map_prec = -1;
  long long act2_tag = 0;                  // geometry the parity-split act2 buffer was last zeroed for     // shape the padded maps were last zeroed for
  DevBuf outer;
  DevBuf parent, labels, slot_of, comps, cls_sums, out_index, row_ext, run_label;
  DevBuf prep_tab, prep_a, prep_b, prep_in, prep_out;     // input-side resize tables and intermediates (ubd_prep.cuh)
  DevBuf l2dense;                 // merged dense 3x3 kernel of the stem's L2 (+ bias)
  DevBuf stem_wimg;               // pointwise B images of L2 / L3 for the tensor-core stem
  bool stem_weights_dirty = true;
  DevBuf tc_trace;                // optional event trace of CTA 0 (option "tc_trace")
  DevBuf tc_weights;              // per-layer UMMA B-operand images (+ bias)
  DevBuf tc4_weights;             // per-layer weight images of the column-rotating kernel (+ bias)
  DevBuf headw_dev;               // padded head kernel + bias, staged into the constant bank before a head launch
  bool headw_dirty = true;
  DevBuf tc4_bwd;                 // training: flipped kernels (HWIO) + zero bias + offsets, then their tf32 weight images
  bool bwd_offs_ready = false;
  int opt_train_tc_bits = 7;
  bool opt_train_tc = true;       // training step of a tf32 handle runs its dilated layers (fwd, dgrad, wgrad) on tcgen05
  // training workspaces
  DevBuf t_acts, t_grads_act, t_scratch, t_partials, d_grads, d_adam_m, d_adam_v, d_ytrue, d_dlogits, t_loss, d_metric;
  int64_t adam_t = 0;
  int train_n = 0, train_h = 0, train_w = 0;   // geometry the training maps were last zeroed for
  bool have_grads = false;
  void* nccl_comm = nullptr;      // ncclComm_t of the data-parallel group (ubd_comm_init)
  int nccl_world = 0;
  const float* last_logits = nullptr;   // logits / targets of that evaluation (may be the caller's device buffers)
  const int* last_ytrue = nullptr;
  size_t loss_pixels = 0;         // pixels of the logits / targets of the last loss evaluation (ubd_metric_counts)
  void* h_stage = nullptr;        // pinned staging (unused unless requested)
  float* h_parts = nullptr;       // pinned loss parts of a deferred training step (ubd_train_update)

  std::vector<DevBuf*> all_bufs() {
    std::vector<DevBuf*> v = {&d_images, &d_logits, &d_mask, &act1, &act2, &mapA, &mapB, &mapC, &outer, &parent, &labels, &slot_of, &comps,
            &cls_sums, &out_index, &row_ext, &run_label, &pipe_ring, &pipe_flags, &prep_tab, &prep_a, &prep_b, &prep_in, &prep_out,
            &tc_weights, &tc4_weights, &tc4_bwd, &headw_dev, &tc_trace, &stem_wimg, &l2dense, &t_acts, &t_grads_act,
            &t_scratch, &t_partials, &d_grads, &d_adam_m, &d_adam_v, &d_ytrue, &d_dlogits, &t_loss, &d_metric};
    for (ResultSlot& R : rs)
      for (DevBuf* b : {&R.hdr, &R.out_recs, &R.hull_pts, &R.box_recs, &R.d_images, &R.d_mask, &R.d_logits}) v.push_back(b);
    return v;
  }
};

// 16-bit map containers: bf16 or IEEE half (same layouts and kernels, see tc::pack16)
static inline bool ubd_is16(ubd_handle h) { return h->precision == UBD_BF16 || h->precision == UBD_F16; }

// TF 'same' stride-2 padding before the image for even sizes is 0; FML pads 1 (net.py:229-232).
static inline int stride2_pad(ubd_handle h) { return h->fml ? 1 : 0; }
