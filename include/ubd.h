/*
 * ubd.h -- C ABI of libubd.so: the B200 (sm_100a) implementation of the ubdvss segment + CC hot path.
 *
 * The reference (asmekal/ubdvss) is pure Python and has no FFI of its own: its boundary for this
 * path is a Keras ``Model`` plus three callables.  Each entry point below replaces
 * one of those call sites; the Python shim in ``ubdvss_b200/`` binds them with ctypes and mirrors
 * the reference's names, argument meaning and error behaviour (INTEGRATION.md shows the stub a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 on success or a negative ubd_status; no exception crosses the ABI;
 *     ubd_last_error(h) returns the message of the last failing call on that handle;
 *   - plain pointers and sizes only; the caller owns every buffer passed in; the library owns its
 *     device workspaces, sized lazily from (n, h, w);
 *   - one handle = one device + one stream; a handle is NOT thread-safe, distinct handles are
 *     independent (one per GPU / process);
 *   - image sides must be positive multiples of 16 (the reference feeds multiples of 64,
 *     segmap_manager.py:153-165), the segmentation map is (h/4, w/4) (net.py:314);
 *   - "host" entry points take host pointers (pageable or pinned) and include the copies;
 *     "_dev" entry points take device pointers on the handle's device and run on its stream.
 */
#ifndef UBD_H_
#define UBD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ubd_handle_s* ubd_handle;

typedef enum {
  UBD_OK = 0,
  UBD_ERR_ARG = -1,           /* bad argument (shape, dtype, NULL)                       */
  UBD_ERR_CUDA = -2,          /* CUDA runtime error; message in ubd_last_error           */
  UBD_ERR_NO_WEIGHTS = -3,    /* forward/train called before ubd_set_weights             */
  UBD_ERR_OVERFLOW = -4,      /* more components / hull points than the caller's buffers */
  UBD_ERR_UNSUPPORTED = -5,   /* configuration not built (e.g. precision on this path)   */
  UBD_ERR_STATE = -6          /* call order (e.g. adam step before a train step)         */
} ubd_status;

typedef enum { UBD_U8 = 0, UBD_F32 = 1 } ubd_dtype;                /* image element type */
typedef enum { UBD_PREPROC_NONE = 0,                               /* net.py:164-165     */
               UBD_PREPROC_MOBILENET = 1 } ubd_preproc;            /* net.py:217-218     */
typedef enum { UBD_FP32 = 0,      /* FP32 CUDA-core path (exact mode)                    */
               UBD_TF32 = 1,      /* tcgen05 kind::tf32 implicit GEMM for the dilated layers */
               UBD_BF16 = 2,      /* bf16 maps and weights, fp32 accumulate                                  */
               UBD_F16 = 3 }      /* IEEE-half maps and weights (the 10-bit significand of tf32 in 16 bits), fp32
                                     accumulate: tf32's accuracy at bf16's traffic; activations must stay below 65504 */
               ubd_precision;

#define UBD_N_WEIGHT_ARRAYS 23    /* model.get_weights() of net.py:286-313    */
#define UBD_MAX_CLASSES 32

/* One external connected component of a thresholded map (utils.py:51-60, segmap_manager.py:54-69). */
typedef struct {
  int32_t image;        /* index in the batch                                                   */
  int32_t label;        /* raster index y*w+x of the component's first pixel (canonical id)     */
  int32_t xmin, ymin, xmax, ymax;   /* = cv2.boundingRect of the contour, inclusive             */
  int32_t n_pixels;     /* foreground pixels                                                    */
  int32_t n_filled;     /* pixels of the filled contour (holes and nested blobs included)       */
  int32_t area_x2;      /* 2 * cv2.contourArea(contour) (utils.py:55), exact integer            */
  int32_t class_id;     /* argmax of the mean softmax over the filled contour, -1 if no classes */
  float box[8];         /* cv2.boxPoints(cv2.minAreaRect(contour)) in map pixels (utils.py:56-57) */
} ubd_component;

/* ---- lifetime ------------------------------------------------------------------------------- */

/* Replaces NetManager.build_model / load_model (net.py:273-314, 443-466): fixes the architecture
 * options that change the graph.  grey: 1 channel else 3 (net.py:286); fml_compatible: top/left
 * zero pad before the stride-2 layers (net.py:229-232); n_classes: class-head width (net.py:307-311). */
int ubd_create(int device, int grey, int fml_compatible, int n_classes, int precision, ubd_handle* out);
int ubd_destroy(ubd_handle h);
const char* ubd_last_error(ubd_handle h);          /* h may be NULL: error of a failed ubd_create */
int ubd_version(void);
int ubd_device_count(void);                         /* 0 when no CUDA device is visible            */

/* Replaces model.set_weights / load_weights and get_weights / save_weights (net.py:418-427):
 * 23 host float32 arrays in Keras order and layout (depthwise (3,3,Cin,1), pointwise (1,1,Cin,24),
 * bias; conv (3,3,24,24) HWIO, bias; head (1,1,24,1+C), bias).  n_elems[i] is checked. */
int ubd_set_weights(ubd_handle h, const float* const* arrays, const int64_t* n_elems, int n_arrays);
int ubd_get_weights(ubd_handle h, float* const* arrays, const int64_t* n_elems, int n_arrays);
/* Tuning / diagnostic switches: "chunk" (images per sweep of the dilated layers, 0 = auto), "stem_chunk"
 * (images per stem launch, 0 = auto), "max_comps" (component slots per image), "max_points" (hull
 * candidate capacity), "precision" (UBD_FP32/TF32/BF16), "profile" (CUDA-event stage timers, see
 * ubd_get_stat), "stem_variant" (grey input: 2 = fused image->L1->L2->L3 kernel, 1 = two-kernel stem, 0 = auto),
 * "dense_l2" (0: depthwise stem on the FP32 pipes), "tc_variant" (0: first-generation tensor-core kernel for
 * the dilated layers and the stem's L2), "tc_trace" (in-kernel cycle trace).  "max_comps" grows by itself when
 * an image has more raw components than slots (the reference's cv2 path has no limit, utils.py:52). */
int ubd_set_option(ubd_handle h, const char* name, int64_t value);

/* ---- inference ------------------------------------------------------------------------------ */

/* Replaces model.predict(images) (model_runner.py:119, predict.py:74-76).  images: host (n,h,w,Cin)
 * of in_dtype; preproc folds NetConfig.get_preprocessing_fn (net.py:163-169) into the first layer.
 * logits_out: host float32 (n,h/4,w/4,1+C). */
int ubd_forward(ubd_handle h, const void* images, int in_dtype, int n, int height, int width,
                int preproc, float* logits_out);

/* Replaces ModelRunner.predict (model_runner.py:105-138): forward, ``logit > logit_thr`` (strict,
 * float32), external components, ``2*contourArea > min_area_x2`` filter, min-area boxes, class vote.
 * mask_out: host uint8 (n,h/4,w/4); logits_out (nullable): host float32 (n,h/4,w/4,1+C);
 * labels_out (nullable): host int32 (n,h/4,w/4), -1 = no component; comps_out: capacity max_comps
 * (kept components of all images, image-major, bottom-up within an image as cv2 lists them);
 * n_comps_per_image: host int32[n].  */
int ubd_segment(ubd_handle h, const void* images, int in_dtype, int n, int height, int width,
                int preproc, float logit_thr, int min_area_x2,
                uint8_t* mask_out, float* logits_out, int32_t* labels_out,
                ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image);

/* Replaces SegmapManager.postprocess / utils.get_contours_and_boxes (segmap_manager.py:42-69,
 * utils.py:51-60) on a given map.  mask: host uint8 (n,mh,mw), non-zero = foreground;
 * cls_logits (nullable): host float32 (n,mh,mw,n_cls).  Outputs as in ubd_segment. */
int ubd_postprocess(ubd_handle h, const uint8_t* mask, const float* cls_logits, int n, int mh, int mw,
                    int n_cls, int min_area_x2, int32_t* labels_out,
                    ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image);

/* Device-resident variants (inputs already in HBM; used for the kernel-throughput number).
 * d_images: device (n,h,w,Cin); d_mask / d_logits (nullable): device outputs.  The kept components
 * are still returned to the host (a few dozen bytes per component). */
int ubd_segment_dev(ubd_handle h, const void* d_images, int in_dtype, int n, int height, int width,
                    int preproc, float logit_thr, int min_area_x2,
                    uint8_t* d_mask, float* d_logits,
                    ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image);
int ubd_forward_dev(ubd_handle h, const void* d_images, int in_dtype, int n, int height, int width,
                    int preproc, float* d_logits);

/* Pipelined form of ubd_segment (the loop of ModelRunner.run, model_runner.py:40-103, calls predict once per
 * batch): submit queues one batch - host-to-device copy of the images, network, threshold, CC kernels, the
 * copies of mask / logits back to mask_out / logits_out (nullable; must stay valid until the wait) - and returns a
 * ticket; wait blocks until that batch is finished, computes its boxes on the host and fills comps_out
 * (capacity max_comps >= the value given to submit) and n_comps_per_image.  At most three batches may be in flight
 * and tickets are collected in submission order: the copies of batches k+1, k+2 and the host part of batch k then run
 * under the kernels of the others, and the CC stage of batch k overlaps the network of batch k+1.  Pass pinned host memory for the copies to be asynchronous.  The _dev form
 * takes images already resident on the handle's device.  The synchronous entry points return UBD_ERR_STATE while
 * a submitted batch is in flight. */
int ubd_segment_submit(ubd_handle h, const void* images, int in_dtype, int n, int height, int width,
                       int preproc, float logit_thr, int min_area_x2,
                       uint8_t* mask_out, float* logits_out, int max_comps, int* ticket);
int ubd_segment_submit_dev(ubd_handle h, const void* d_images, int in_dtype, int n, int height, int width,
                           int preproc, float logit_thr, int min_area_x2, int max_comps, int* ticket);
int ubd_segment_wait(ubd_handle h, int ticket, ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image);

/* Input side (segmap_manager.py:136-167, data_generators.py:176-177): what the reference does to a decoded 8-bit image
 * before the network - ``image.resize((out_w, out_h), Image.BICUBIC)`` then ``image.convert('L')`` when the net is grey -
 * on the GPU, bit-identical to Pillow (fixed-point separable resampling with antialiasing, luma weights 19595/38470/7471).
 * images: (n, h, w, c) uint8, c = 1 or 3; out: (n, out_h, out_w, 1 if to_grey and c == 3 else c).  The _dev form takes
 * and leaves device buffers, so its output can be handed to ubd_segment_dev / ubd_segment_submit_dev directly. */
int ubd_prepare_images(ubd_handle h, const uint8_t* images, int n, int height, int width, int channels,
                       int out_height, int out_width, int to_grey, uint8_t* out);
int ubd_prepare_images_dev(ubd_handle h, const uint8_t* d_images, int n, int height, int width, int channels,
                           int out_height, int out_width, int to_grey, uint8_t* d_out);

/* Pure host helper, usable without a GPU: cv2.boxPoints(cv2.minAreaRect(pts)) for one point set
 * (utils.py:56-57).  pts: n_pts (x,y) int32 pairs (any superset of the hull); box: 8 floats. */
int ubd_min_area_box(const int32_t* pts_xy, int n_pts, float* box);

/* ---- training (train.py:110-112, 176-188; losses.py) ---------------------------------------- */

/* One Keras train_on_batch minus the optimizer: forward, detection[_and_classification]_loss
 * (losses.py:33-126), backward into the handle's flat gradient buffer (W1 order).
 * y_true: host int32 (n,h/4,w/4), 0 = background, i>0 = class i-1.
 * loss_parts[6] = {loss, positive, negative, hard_negative, classification, k}. */
int ubd_train_step(ubd_handle h, const void* images, int in_dtype, int n, int height, int width,
                   int preproc, const int32_t* y_true, float* loss_parts);
int ubd_train_step_dev(ubd_handle h, const void* d_images, int in_dtype, int n, int height, int width,
                       int preproc, const int32_t* d_y_true, float* loss_parts);
/* Loss and dL/dlogits only (losses.py on given logits): host (n,mh,mw,1+C) in, same shape out. */
int ubd_loss(ubd_handle h, const float* logits, const int32_t* y_true, int n, int mh, int mw,
             float* loss_parts, float* dlogits_out);
int ubd_get_grads(ubd_handle h, float* const* arrays, const int64_t* n_elems, int n_arrays);
/* The flat device gradient buffer, for the NCCL all-reduce (torch.distributed) between
 * ubd_train_step and ubd_adam_step in data-parallel training. */
int ubd_grad_buffer(ubd_handle h, void** d_ptr, int64_t* n_floats);
/* Keras-2 Adam (train.py:110): grads are multiplied by grad_scale first (1/world after a sum
 * all-reduce); step count is kept in the handle. */
int ubd_adam_step(ubd_handle h, float lr, float beta_1, float beta_2, float epsilon, float grad_scale);
/* One whole optimizer step with a single host synchronisation (Keras train_on_batch as train.py:110-112 drives it):
 * ubd_train_step + ubd_allreduce_grads (when ubd_comm_init has been called; Adam then scales by 1 / world) +
 * ubd_adam_step, queued back to back; loss_parts as ubd_train_step. */
int ubd_train_update(ubd_handle h, const void* images, int in_dtype, int n, int height, int width, int preproc,
                     const int32_t* y_true, float lr, float beta_1, float beta_2, float epsilon, float* loss_parts);

/* Data-parallel training exchange (one process per GPU; the reference is single-GPU, train.py): rank 0 creates a
 * 128-byte NCCL unique id and hands it to every rank out of band; each rank joins with ubd_comm_init; between
 * ubd_train_step and ubd_adam_step(grad_scale = 1/world) ubd_allreduce_grads sums the flat gradient buffer over
 * the ranks in place on the handle's stream.  NCCL is bound at run time (dlopen "libnccl.so.2"). */
int ubd_comm_unique_id(void* id128);
int ubd_comm_init(ubd_handle h, const void* id128, int rank, int world);
int ubd_allreduce_grads(ubd_handle h);
int ubd_comm_destroy(ubd_handle h);

/* Pixel statistics of the last ubd_train_step / ubd_loss batch for the training metrics the reference logs
 * (keras_metrics.py:116-191): counts[6] = tp, tn, fp, fn of the detection channel (prediction: logit > 0,
 * truth: y_true > 0), then correct / total class predictions over object pixels (arg-max class vs y_true - 1). */
int ubd_metric_counts(ubd_handle h, int64_t* counts);
int ubd_synchronize(ubd_handle h);
/* Run on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)
 * instead of the handle's own, so that the caller's events bracket the work.  NULL restores it. */
int ubd_set_stream(ubd_handle h, void* cuda_stream);
/* Named counters: "launches"; with option "profile" = 1 also "dilconv_ms" / "dilconv_launches"
 * (CUDA-event time of the dilated-conv kernels), "stem_ms", "ccl_ms", "head_ms". */
int ubd_get_stat(ubd_handle h, const char* name, double* value);
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t ubd_launch_count(ubd_handle h);

/* Test hook (layer-level parity, bring-up): dilated layer `layer` (0..5 = conv2d_1..6 with its own
 * weights and dilation) on a host NHWC (n,mh,mw,24) map through the FP32 or the tcgen05 kernel. */
int ubd_debug_dilated_layer(ubd_handle h, const float* in_nhwc, float* out_nhwc, int layer,
                            int n, int mh, int mw, int precision);
/* Debug: weight and bias gradient of one dilated 3x3 layer on the tensor cores (the K = pixels GEMM of the training
 * step, ubd_wgrad.cuh): x = the layer's input map, g = the gradient at its output, both (n,h,w,24) fp32 NHWC;
 * dK = (3,3,24,24) HWIO, dB = (24). */
int ubd_debug_wgrad(ubd_handle h, const float* x_nhwc, const float* g_nhwc, int n, int height, int width,
                    int dilation, float* dK, float* dB);

/* Tuning hook: with option "tc_trace" = 1 the tcgen05 kernel records cycle stamps of CTA 0
 * ([3 roles][1024 events][4 stamps], int64); this reads and clears them. */
int ubd_debug_read_trace(ubd_handle h, long long* out, int n_values);

#ifdef __cplusplus
}
#endif
#endif  /* UBD_H_ */
