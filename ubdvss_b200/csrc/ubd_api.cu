// libubd.so: C ABI (include/ubd.h) over the sm_100a kernels.  Host orchestration only: workspace
// management, chunking of the batch so that inter-layer feature maps stay L2-resident, launches.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include <dlfcn.h>
#include "ubd_ccl.cuh"
#include "ubd_prep.cuh"
#include "ubd_common.cuh"
#include "ubd_fp32.cuh"
#include "ubd_handle.cuh"
#include "ubd_tc.cuh"
#include "ubd_tc4.cuh"
#include "ubd_stem_tc.cuh"
#include "ubd_stemf.cuh"
#include "ubd_train.cuh"
#include "ubd_wgrad.cuh"

static std::string g_create_error;

void ubd_box_from_device(float cx, float cy, float w, float hgt, float ax, float ay, int n_hull,
                         int x0, int y0, int x1, int y1, float* box);      // ubd_rect.cpp

#define LAUNCH_CHECK()                                                     \
  do {                                                                     \
    ++h->launches;                                                         \
    cudaError_t e_ = cudaGetLastError();                                   \
    if (e_ != cudaSuccess) {                                               \
      h->err = std::string("kernel launch: ") + cudaGetErrorString(e_);    \
      return UBD_ERR_CUDA;                                                 \
    }                                                                      \
  } while (0)

// CUDA-event bracket of a kernel group on the handle's stream; resolved lazily at the next sync.
struct ProfScope {
  ubd_handle h; ubd_handle_s::Prof* p; cudaEvent_t a = nullptr, b = nullptr; int n0; cudaStream_t st;
  ProfScope(ubd_handle h_, ubd_handle_s::Prof* p_, cudaStream_t st_ = nullptr) : h(h_), p(p_), n0((int)h_->launches), st(st_ ? st_ : h_->stream) {
    if (!h->profile) return;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (!h->profile) return;
    cudaEventRecord(b, st);
    p->launches += h->launches - n0;
    h->pending.push_back({p, {a, b}});
  }
};
static void resolve_profile(ubd_handle h) {
  for (auto& e : h->pending) {
    float ms = 0.f;
    cudaEventSynchronize(e.second.second);
    cudaEventElapsedTime(&ms, e.second.first, e.second.second);
    e.first->ms += ms;
    cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second);
  }
  h->pending.clear();
}

struct HostTimer {
  ubd_handle h; int slot; std::chrono::steady_clock::time_point t0;
  HostTimer(ubd_handle h_, int slot_) : h(h_), slot(slot_), t0(std::chrono::steady_clock::now()) {}
  ~HostTimer() {
    if (h->profile) h->host_ms[slot] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

static int ensure(ubd_handle h, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return UBD_OK;
  if (b.p) { UBD_CUDA(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
  size_t want = bytes + (bytes >> 3);
  UBD_CUDA(cudaMalloc(&b.p, want));
  b.cap = want;
  return UBD_OK;
}
#define ENSURE(buf, bytes) do { int rc_ = ensure(h, (buf), (bytes)); if (rc_) return rc_; } while (0)

extern "C" int ubd_version(void) { return 100; }

extern "C" int ubd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const char* ubd_last_error(ubd_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int ubd_create(int device, int grey, int fml_compatible, int n_classes, int precision, ubd_handle* out) {
  if (!out) { g_create_error = "ubd_create: out is NULL"; return UBD_ERR_ARG; }
  *out = nullptr;
  if (n_classes < 0 || n_classes > UBD_MAX_CLASSES) { g_create_error = "ubd_create: n_classes out of range"; return UBD_ERR_ARG; }
  if (precision < UBD_FP32 || precision > UBD_F16) { g_create_error = "ubd_create: unknown precision"; return UBD_ERR_ARG; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_create_error = "ubd_create: no CUDA device visible (libubd has no CPU fallback)";
    return UBD_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_error = "ubd_create: device index out of range"; return UBD_ERR_ARG; }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return UBD_ERR_CUDA; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = "ubd_create: libubd is built for sm_100a (B200) only; device is sm_" +
                     std::to_string(prop.major) + std::to_string(prop.minor);
    return UBD_ERR_UNSUPPORTED;
  }
  ubd_handle h = new ubd_handle_s();
  h->device = device; h->grey = grey != 0; h->fml = fml_compatible != 0; h->n_classes = n_classes;
  h->precision = precision; h->spec = make_weight_spec(grey, n_classes);
  h->n_sm = prop.multiProcessorCount;
  if (const char* ev = getenv("UBD_STEM_VARIANT")) h->opt_stem_variant = atoi(ev);      // A/B runs of bench.py
  e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  h->stream = h->own_stream;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->rec_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    // the CC stage runs on its own (higher-priority) stream so that it overlaps the next batch's network kernels
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    e = cudaStreamCreateWithPriority(&h->cc_stream, cudaStreamNonBlocking, hi);
  }
  if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete h; return UBD_ERR_CUDA; }
  // preprocessing table for uint8 input: exactly what numpy computes, (v - 127.5) / 127.5 in
  // float64 (net.py:217-218 on a uint8 image) then cast to float32 at the Keras boundary
  float lut[256];
  for (int v = 0; v < 256; ++v) lut[v] = (float)(((double)v - 127.5) / 127.5);
  if (cudaMalloc(&h->d_lut, sizeof(lut)) != cudaSuccess ||
      cudaMemcpy(h->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice) != cudaSuccess) {
    g_create_error = "ubd_create: cudaMalloc failed"; delete h; return UBD_ERR_CUDA;
  }
  size_t wbytes = (size_t)h->spec.total * sizeof(float);
  if (cudaMalloc(&h->d_params, wbytes) != cudaSuccess) { g_create_error = "ubd_create: cudaMalloc failed"; delete h; return UBD_ERR_CUDA; }
  cudaFuncSetAttribute(dilconv_fp32_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 24 * 24 * 4);
  cudaFuncSetAttribute(dilconv_fp32_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 24 * 24 * 4);
  cudaFuncSetAttribute(dilconv_fp32_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 24 * 24 * 4);
  tc_setup_attributes();
  tc4_setup_attributes();
  stem_setup_attributes();
  stemf_setup_attributes();
  cudaFuncSetAttribute(ccl_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ccl_image_smem(CCL_IMG_MAX_PX));
  cudaFuncSetAttribute(ccl_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  *out = h;
  return UBD_OK;
}

extern "C" int ubd_destroy(ubd_handle h) {
  if (!h) return UBD_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->nccl_comm) ubd_comm_destroy(h);
  for (DevBuf* b : h->all_bufs()) if (b->p) cudaFree(b->p);
  if (h->d_params) cudaFree(h->d_params);
  if (h->d_lut) cudaFree(h->d_lut);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  resolve_profile(h);
  for (cudaEvent_t ev : h->copy_events) cudaEventDestroy(ev);
  for (ResultSlot& R : h->rs) {
    if (R.h_hdr) cudaFreeHost(R.h_hdr);
    if (R.ev_cc) cudaEventDestroy(R.ev_cc);
    if (R.ev_in) cudaEventDestroy(R.ev_in);
    if (R.ev_d2h) cudaEventDestroy(R.ev_d2h);
    if (R.ev_fwd) cudaEventDestroy(R.ev_fwd);
  }
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  if (h->h_parts) cudaFreeHost(h->h_parts);
  if (h->rec_stream) cudaStreamDestroy(h->rec_stream);
  if (h->cc_stream) cudaStreamDestroy(h->cc_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  cudaStreamDestroy(h->own_stream);
  delete h;
  return UBD_OK;
}

extern "C" int ubd_synchronize(ubd_handle h) {
  if (!h) return UBD_ERR_ARG;
  UBD_CUDA(cudaSetDevice(h->device));
  UBD_CUDA(cudaStreamSynchronize(h->stream));
  return UBD_OK;
}

extern "C" int64_t ubd_launch_count(ubd_handle h) { return h ? h->launches : 0; }

extern "C" int ubd_set_stream(ubd_handle h, void* cuda_stream) {
  if (!h) return UBD_ERR_ARG;
  UBD_CUDA(cudaSetDevice(h->device));
  UBD_CUDA(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return UBD_OK;
}

extern "C" int ubd_get_stat(ubd_handle h, const char* name, double* value) {
  if (!h || !name || !value) return UBD_ERR_ARG;
  resolve_profile(h);
  if (!strcmp(name, "launches")) *value = (double)h->launches;
  else if (!strcmp(name, "dilconv_ms")) *value = h->prof_dil.ms;
  else if (!strcmp(name, "dilconv_launches")) *value = (double)h->prof_dil.launches;
  else if (!strcmp(name, "stem_ms")) *value = h->prof_stem.ms;
  else if (!strcmp(name, "stem_launches")) *value = (double)h->prof_stem.launches;
  else if (!strcmp(name, "head_ms")) *value = h->prof_head.ms;
  else if (!strcmp(name, "ccl_ms")) *value = h->prof_ccl.ms;
  else if (!strcmp(name, "ccl_launches")) *value = (double)h->prof_ccl.launches;
  else if (!strncmp(name, "host_ms", 7) && name[7] >= '0' && name[7] <= '7') *value = h->host_ms[name[7] - '0'];
  else UBD_FAIL(UBD_ERR_ARG, std::string("unknown stat ") + name);
  return UBD_OK;
}

extern "C" int ubd_set_option(ubd_handle h, const char* name, int64_t value) {
  if (!h || !name) return UBD_ERR_ARG;
  if (!strcmp(name, "chunk")) { if (value < 0) UBD_FAIL(UBD_ERR_ARG, "chunk must be >= 0"); h->opt_chunk = (int)value; }
  else if (!strcmp(name, "max_comps")) { if (value < 1) UBD_FAIL(UBD_ERR_ARG, "max_comps must be >= 1"); h->opt_max_comps = (int)value; }
  else if (!strcmp(name, "max_points")) { if (value < 1) UBD_FAIL(UBD_ERR_ARG, "max_points must be >= 1"); h->opt_max_points = (int)value; }
  else if (!strcmp(name, "profile")) {
    resolve_profile(h);
    h->profile = value != 0;
    h->prof_dil = h->prof_stem = h->prof_ccl = h->prof_head = ubd_handle_s::Prof();
    for (double& v : h->host_ms) v = 0;
  }
  else if (!strcmp(name, "dense_l2")) h->opt_dense_l2 = value != 0;
  else if (!strcmp(name, "stem_variant")) h->opt_stem_variant = (int)value;
  else if (!strcmp(name, "gpu_boxes")) h->opt_gpu_boxes = value != 0;
  else if (!strcmp(name, "fused_ccl")) h->opt_fused_ccl = (int)value;
  else if (!strcmp(name, "tc_variant")) h->opt_tc_variant = (int)value;
  else if (!strcmp(name, "pipeline")) h->opt_pipeline = value != 0;
  else if (!strcmp(name, "cc_stream")) h->opt_cc_stream = value != 0;
  else if (!strcmp(name, "train_tc")) h->opt_train_tc = value != 0;
  else if (!strcmp(name, "train_tc_bits")) h->opt_train_tc_bits = (int)value;
  else if (!strcmp(name, "pipe_ring")) { if (value < 2) UBD_FAIL(UBD_ERR_ARG, "pipe_ring must be >= 2"); h->opt_pipe_ring = (int)value; }
  else if (!strcmp(name, "stem_chunk")) h->opt_stem_chunk = (int)value;
  else if (!strcmp(name, "tc_trace")) {
    if (value) { ENSURE(h->tc_trace, 8 * 1024 * 4 * sizeof(long long)); UBD_CUDA(cudaMemset(h->tc_trace.p, 0, h->tc_trace.cap)); }
    else if (h->tc_trace.p) { cudaFree(h->tc_trace.p); h->tc_trace.p = nullptr; h->tc_trace.cap = 0; }
  }
  else if (!strcmp(name, "precision")) {
    if (value < UBD_FP32 || value > UBD_F16) UBD_FAIL(UBD_ERR_ARG, "bad precision");
    // the 16-bit weight images hold bf16 or half: rebuild them when the container changes
    if ((int)value != h->precision) { h->tc_weights_dirty = h->tc4_weights_dirty = h->tc4_train_dirty = h->stem_weights_dirty = true; h->act2_tag = 0; }
    h->precision = (int)value;
  }
  else UBD_FAIL(UBD_ERR_ARG, std::string("unknown option ") + name);
  return UBD_OK;
}

static int check_weight_args(ubd_handle h, const void* arrays, const int64_t* n_elems, int n_arrays) {
  if (!arrays || !n_elems) UBD_FAIL(UBD_ERR_ARG, "weights: NULL argument");
  if (n_arrays != UBD_N_WEIGHT_ARRAYS) UBD_FAIL(UBD_ERR_ARG, "weights: expected 23 arrays (Keras get_weights() order)");
  for (int i = 0; i < n_arrays; ++i)
    if (n_elems[i] != h->spec.size[i])
      UBD_FAIL(UBD_ERR_ARG, "weights: array " + std::to_string(i) + " has " + std::to_string(n_elems[i]) +
                                " elements, expected " + std::to_string(h->spec.size[i]));
  return UBD_OK;
}

extern "C" int ubd_set_weights(ubd_handle h, const float* const* arrays, const int64_t* n_elems, int n_arrays) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_weight_args(h, arrays, n_elems, n_arrays);
  if (rc) return rc;
  UBD_CUDA(cudaSetDevice(h->device));
  std::vector<float> flat(h->spec.total, 0.f);
  for (int i = 0; i < n_arrays; ++i) {
    if (!arrays[i]) UBD_FAIL(UBD_ERR_ARG, "weights: NULL array");
    memcpy(flat.data() + h->spec.off[i], arrays[i], h->spec.size[i] * sizeof(float));
  }
  UBD_CUDA(cudaMemcpyAsync(h->d_params, flat.data(), flat.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  UBD_CUDA(cudaStreamSynchronize(h->stream));
  h->have_weights = true;
  h->tc_weights_dirty = true;
  h->tc4_weights_dirty = true;
  h->tc4_train_dirty = true;
  h->headw_dirty = true;
  h->stem_weights_dirty = true;
  return UBD_OK;
}

extern "C" int ubd_get_weights(ubd_handle h, float* const* arrays, const int64_t* n_elems, int n_arrays) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_weight_args(h, arrays, n_elems, n_arrays);
  if (rc) return rc;
  if (!h->have_weights) UBD_FAIL(UBD_ERR_NO_WEIGHTS, "get_weights before set_weights");
  UBD_CUDA(cudaSetDevice(h->device));
  std::vector<float> flat(h->spec.total);
  UBD_CUDA(cudaMemcpyAsync(flat.data(), h->d_params, flat.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  UBD_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n_arrays; ++i) memcpy(arrays[i], flat.data() + h->spec.off[i], h->spec.size[i] * sizeof(float));
  return UBD_OK;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------

static int check_image_args(ubd_handle h, const void* images, int in_dtype, int n, int H, int W, int preproc) {
  if (!images) UBD_FAIL(UBD_ERR_ARG, "images is NULL");
  if (in_dtype != UBD_U8 && in_dtype != UBD_F32) UBD_FAIL(UBD_ERR_ARG, "in_dtype must be UBD_U8 or UBD_F32");
  if (preproc != UBD_PREPROC_NONE && preproc != UBD_PREPROC_MOBILENET) UBD_FAIL(UBD_ERR_ARG, "unknown preprocessing");
  if (n < 1 || H < 16 || W < 16 || (H % 16) || (W % 16))
    UBD_FAIL(UBD_ERR_ARG, "image sides must be positive multiples of 16 (reference feeds multiples of 64, segmap_manager.py:153-165)");
  if (!h->have_weights) UBD_FAIL(UBD_ERR_NO_WEIGHTS, "no weights loaded (ubd_set_weights)");
  return UBD_OK;
}

static size_t image_bytes(ubd_handle h, int in_dtype, int n, int H, int W) {
  return (size_t)n * H * W * h->spec.cin * (in_dtype == UBD_U8 ? 1 : 4);
}

// Separable layer launcher (L1: raw image; L2/L3: planar maps).
template <int CIN, int STRIDE, bool RAW, typename TIn>
static int launch_sep(ubd_handle h, const TIn* in, float4* out, int layer, int n, int H, int W, int Ho, int Wo,
                      int pad_t, int pad_l, const float* lut, float pre_scale, float pre_shift,
                      int in_mpad = 0, int out_mpad = 0, float4* dwout = nullptr) {
  const float* base = h->d_params;
  const float* dwk = base + h->spec.off[3 * layer];
  const float* pwk = base + h->spec.off[3 * layer + 1];
  const float* b = base + h->spec.off[3 * layer + 2];
  dim3 grid((Wo + 31) / 32, (Ho + 3) / 4, n), block(128);
  sep_layer_kernel<CIN, STRIDE, RAW, TIn><<<grid, block, 0, h->stream>>>(in, out, dwk, pwk, b, lut, pre_scale, pre_shift,
                                                                        n, H, W, Ho, Wo, pad_t, pad_l, in_mpad, out_mpad, dwout);
  LAUNCH_CHECK();
  return UBD_OK;
}


static int run_stem(ubd_handle h, const void* d_img, int in_dtype, int preproc, int n, int H, int W,
                    float4* act1, float4* act2, float4* act3, float4* dw2 = nullptr, float4* dw3 = nullptr) {
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  const int p2 = stride2_pad(h);
  const bool mob = preproc == UBD_PREPROC_MOBILENET;
  int rc;
  if (h->spec.cin == 1) {
    if (in_dtype == UBD_U8)
      rc = launch_sep<1, 2, true, uint8_t>(h, (const uint8_t*)d_img, act1, 0, n, H, W, H2, W2, p2, p2, mob ? h->d_lut : nullptr, 0.f, 0.f);
    else
      rc = launch_sep<1, 2, true, float>(h, (const float*)d_img, act1, 0, n, H, W, H2, W2, p2, p2, nullptr, mob ? 127.5f : 0.f, 127.5f);
  } else {
    if (in_dtype == UBD_U8)
      rc = launch_sep<3, 2, true, uint8_t>(h, (const uint8_t*)d_img, act1, 0, n, H, W, H2, W2, p2, p2, mob ? h->d_lut : nullptr, 0.f, 0.f);
    else
      rc = launch_sep<3, 2, true, float>(h, (const float*)d_img, act1, 0, n, H, W, H2, W2, p2, p2, nullptr, mob ? 127.5f : 0.f, 127.5f);
  }
  if (rc) return rc;
  rc = launch_sep<24, 1, false, float>(h, (const float*)act1, act2, 1, n, H2, W2, H2, W2, 1, 1, nullptr, 0.f, 0.f, 0, 0, dw2);
  if (rc) return rc;
  // the tensor-core layers read tf32: round (rna) where the map is produced instead of letting the
  // MMA truncate it
  rc = launch_sep<24, 2, false, float>(h, (const float*)act2, act3, 2, n, H2, W2, H4, W4, p2, p2, nullptr,
                                       h->precision == UBD_TF32 ? -1.f : 0.f, 0.f, 0, UBD_MAP_PAD, dw3);
  return rc;
}

static int launch_dil_fp32(ubd_handle h, const float4* in, float4* out, const float* w, const float* b,
                           const float4* gate, int n, int hh, int ww, int d, int mode) {
  dim3 grid((ww + DIL_TW - 1) / DIL_TW, (hh + DIL_TH - 1) / DIL_TH, n), block(128);
  const size_t smem = 9 * 24 * 24 * 4;
  if (mode == 0) dilconv_fp32_kernel<true, true, false><<<grid, block, smem, h->stream>>>(in, out, w, b, nullptr, n, hh, ww, d, UBD_MAP_PAD);
  else if (mode == 1) dilconv_fp32_kernel<false, false, true><<<grid, block, smem, h->stream>>>(in, out, w, nullptr, gate, n, hh, ww, d, UBD_MAP_PAD);
  else dilconv_fp32_kernel<false, false, false><<<grid, block, smem, h->stream>>>(in, out, w, nullptr, nullptr, n, hh, ww, d, UBD_MAP_PAD);
  LAUNCH_CHECK();
  return UBD_OK;
}

static int launch_head(ubd_handle h, const float4* in, float* logits, uint8_t* mask, float thr, int n, int hh, int ww) {
  const float* hk = h->d_params + h->spec.off[21];
  const float* hb = h->d_params + h->spec.off[22];
  dim3 grid((unsigned)(((size_t)hh * ww + 127) / 128), n), block(128);
  head_threshold_kernel<<<grid, block, 0, h->stream>>>(in, logits, mask, hk, hb, h->spec.n_out, thr, n, hh, ww, UBD_MAP_PAD);
  LAUNCH_CHECK();
  return UBD_OK;
}

// Dilated layer on the tensor cores: 1 = column-rotating kernel (ubd_tc4.cuh, default), 0 = one output row
// per accumulator (ubd_tc.cuh); option "tc_variant".
static int launch_dil_tc(ubd_handle h, const void* in, void* out, int layer, int n, int hh, int ww, int d, int out_mode,
                         const tc::HeadArgs* head = nullptr) {
  if (h->opt_tc_variant) return tc4_launch_dilconv(h, in, out, layer, n, hh, ww, d, out_mode, UBD_MAP_PAD, head);
  return tc_launch_dilconv(h, in, out, layer, n, hh, ww, d, out_mode, UBD_MAP_PAD, head);
}

// Images per sweep of the dilated layers / head ("chunk") and per stem launch ("stem chunk").
// Measured on B200 (64 x 1024x1024, tf32), six dilated layers: chunk 16 -> 1.49 ms, 32 -> 1.20 ms,
// 64 -> 1.11 ms: per-launch fixed cost and the tail of the static item schedule outweigh L2 residency
// of the ping-pong maps, so the dilated layers take up to 64 images at once (bounded by ~1 GB of maps).
// The stem runs in sub-chunks of 16 so that the H2D copy of sub-chunk k+1 overlaps the stem of k.
static int pick_chunk(ubd_handle h, int n, int H, int W) {
  if (h->opt_chunk > 0) return std::min(n, h->opt_chunk);
  const double per_img = (double)(H / 4) * (W / 4 + 2 * UBD_MAP_PAD) * 2.0 * 96.0;
  const int c = (int)(1.0e9 / per_img);
  return std::max(1, std::min(n, std::min(std::max(c, 16), 64)));
}
// Host input: 16-image stem launches follow the H2D copies closely (measured end to end 16: 22.7 k, 32: 20.9 k,
// 64: 17.5 k img/s); device-resident input: one launch per chunk is fastest (stem 1.25 -> 1.17 ms per 64 images).
static int pick_stem_chunk(ubd_handle h, int chunk, int H, int W, bool host_input) {
  const double per_img = (double)(H / 2) * (W / 2) * 96.0;        // one half-resolution map
  const int c = (int)((host_input ? 0.5e9 : 2.0e9) / per_img);
  if (h->opt_stem_chunk > 0) return std::max(1, std::min(chunk, h->opt_stem_chunk));
  return std::max(1, std::min(chunk, std::min(std::max(c, 1), host_input ? 16 : 64)));
}

// The two ping-pong quarter-resolution maps.  Their zero x-padding is the convolution's zero padding,
// so the buffers are cleared whenever they are (re)allocated or the map shape changes; kernels only
// ever write the interior.
static int ensure_maps(ubd_handle h, int n, int mh, int mw) {
  const size_t bytes = act_elems(n, mh, mw, UBD_MAP_PAD) * sizeof(float4);
  const bool grow = bytes > h->mapA.cap || bytes > h->mapB.cap;
  ENSURE(h->mapA, bytes);
  ENSURE(h->mapB, bytes);
  // the interior of one layout (fp32: 6 planes, bf16: 3 planes per row) overlaps the pads of the other
  if (grow || h->map_h != mh || h->map_w != mw || h->map_n < n || h->map_prec != h->precision) {
    UBD_CUDA(cudaMemsetAsync(h->mapA.p, 0, h->mapA.cap, h->stream));
    UBD_CUDA(cudaMemsetAsync(h->mapB.p, 0, h->mapB.cap, h->stream));
    h->map_h = mh; h->map_w = mw; h->map_n = n; h->map_prec = h->precision;
  }
  return UBD_OK;
}

// d_img: device images.  d_logits (nullable) / d_mask (nullable): device outputs for the whole batch.
// h_img (nullable): host images; then d_img is the device staging buffer and every chunk is copied on
// the copy stream just ahead of its compute, so the PCIe transfer of chunk k+1 overlaps chunk k.
static int forward_device(ubd_handle h, const void* d_img, int in_dtype, int n, int H, int W, int preproc,
                          float* d_logits, uint8_t* d_mask, float thr, const void* h_img = nullptr, bool copy_hidden = false) {
  h->loss_pixels = 0;            // the handle's logits are about to be overwritten: ubd_metric_counts needs a new loss batch
  h->last_logits = nullptr; h->last_ytrue = nullptr;
  HostTimer ht_fwd(h, 0);
  const int h4 = H / 4, w4 = W / 4;
  const int chunk = pick_chunk(h, n, H, W);
  // copy_hidden: other submitted batches are still running, so this batch's H2D copies finish under their kernels and
  // the stem need not follow them in small launches
  const int schunk = pick_stem_chunk(h, chunk, H, W, h_img != nullptr && !copy_hidden);
  const size_t half_px = (size_t)(H / 2) * (W / 2), q_px = (size_t)h4 * w4;
  if (h->precision == UBD_FP32) ENSURE(h->act1, (size_t)schunk * UBD_NG * half_px * sizeof(float4));
  if (h->precision == UBD_FP32 || !stem_is_fused(h, in_dtype)) {
    // plain layout, or split by column parity with x padding (tensor-core stem); the fused stem has no act2
    const size_t plain = (size_t)schunk * UBD_NG * half_px * sizeof(float4);
    const size_t split = (size_t)schunk * (H / 2) * 2 * UBD_NG * (size_t)(W / 4 + 2 * UBD_MAP_PAD) * sizeof(float4);
    const size_t cap0 = h->act2.cap;
    ENSURE(h->act2, std::max(plain, split));
    if (h->act2.cap != cap0) h->act2_tag = 0;
  }
  { int rc_ = ensure_maps(h, chunk, h4, w4); if (rc_) return rc_; }
  const size_t img_stride = (size_t)H * W * h->spec.cin * (in_dtype == UBD_U8 ? 1 : 4);
  // 16-byte units per image of the L3 output map: 6 planes (fp32 / tf32) or 3 planes (bf16)
  const size_t map_img = act_elems(1, h4, w4, UBD_MAP_PAD) / (ubd_is16(h) ? 2 : 1);
  size_t copy_k = 0;
  for (int c0 = 0; c0 < n; c0 += chunk) {
    const int cn = std::min(chunk, n - c0);
    float4* A = (float4*)h->mapA.p;
    float4* B = (float4*)h->mapB.p;
    int rc;
    // ---- stem, sub-chunk by sub-chunk, each right behind its H2D copy
    for (int s0 = 0; s0 < cn; s0 += schunk) {
      const int sn = std::min(schunk, cn - s0);
      const char* img = (const char*)d_img + (size_t)(c0 + s0) * img_stride;
      if (h_img) {
        while (h->copy_events.size() <= copy_k) {
          cudaEvent_t ev;
          UBD_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
          h->copy_events.push_back(ev);
        }
        UBD_CUDA(cudaMemcpyAsync((void*)img, (const char*)h_img + (size_t)(c0 + s0) * img_stride, (size_t)sn * img_stride,
                                 cudaMemcpyHostToDevice, h->copy_stream));
        UBD_CUDA(cudaEventRecord(h->copy_events[copy_k], h->copy_stream));
        UBD_CUDA(cudaStreamWaitEvent(h->stream, h->copy_events[copy_k], 0));
        ++copy_k;
      }
      ProfScope ps(h, &h->prof_stem);
      float4* a3 = A + (size_t)s0 * map_img;
      if (h->precision != UBD_FP32) rc = run_stem_tc(h, img, in_dtype, preproc, sn, H, W, (float4*)h->act2.p, a3);
      else rc = run_stem(h, img, in_dtype, preproc, sn, H, W, (float4*)h->act1.p, (float4*)h->act2.p, a3);
      if (rc) return rc;
    }
    // ---- dilated layers and head over the whole chunk: one launch per layer, or (option "pipeline") one layer-pipelined
    //      launch whose inter-layer maps stay in L2 rings
    if (h->precision != UBD_FP32 && h->opt_pipeline && h->opt_tc_variant && cn >= 4 * UBD_NLAYERS_DIL && h->n_sm >= 4 * UBD_NLAYERS_DIL) {
      ProfScope ps(h, &h->prof_dil);
      tc::HeadArgs ha{h->d_params + h->spec.off[21], h->d_params + h->spec.off[22], h->spec.n_out, thr,
                      d_logits ? d_logits + (size_t)c0 * q_px * h->spec.n_out : nullptr,
                      d_mask ? d_mask + (size_t)c0 * q_px : nullptr};
      rc = tc4_launch_pipeline(h, A, cn, h4, w4, &ha);
      if (rc) return rc;
      continue;
    }
    for (int l = 0; l < UBD_NLAYERS_DIL; ++l) {
      ProfScope ps(h, &h->prof_dil);
      const float* w = h->d_params + h->spec.off[9 + 2 * l];
      const float* b = h->d_params + h->spec.off[10 + 2 * l];
      const bool last = l == UBD_NLAYERS_DIL - 1;
      if (h->precision == UBD_FP32) {
        rc = launch_dil_fp32(h, A, B, w, b, nullptr, cn, h4, w4, kDilations[l], 0);
      } else if (last) {
        // head + threshold fused into the last layer's epilogue (logits / mask straight from registers)
        tc::HeadArgs ha{h->d_params + h->spec.off[21], h->d_params + h->spec.off[22], h->spec.n_out, thr,
                        d_logits ? d_logits + (size_t)c0 * q_px * h->spec.n_out : nullptr,
                        d_mask ? d_mask + (size_t)c0 * q_px : nullptr};
        rc = launch_dil_tc(h, A, B, l, cn, h4, w4, kDilations[l], /*out_mode=*/2, &ha);
      } else {
        rc = launch_dil_tc(h, A, B, l, cn, h4, w4, kDilations[l], /*out_mode=*/0);
      }
      if (rc) return rc;
      std::swap(A, B);
    }
    if (h->precision == UBD_FP32) {
      ProfScope ps(h, &h->prof_head);
      rc = launch_head(h, A, d_logits ? d_logits + (size_t)c0 * q_px * h->spec.n_out : nullptr,
                       d_mask ? d_mask + (size_t)c0 * q_px : nullptr, thr, cn, h4, w4);
    }
    if (rc) return rc;
  }
  return UBD_OK;
}

// ------------------------------------------------------------------------------------------------
// connected components on device mask -> host component list
// ------------------------------------------------------------------------------------------------

// Result slot s: everything the host reads back from one batch (kept counts, records, hull points) lives in the
// slot's own buffers, so that the kernels of the next batch can run while the host finishes this one
// (ubd_segment_submit / ubd_segment_wait); the per-pixel workspaces (parent, labels, ...) are shared.
static int ccl_enqueue(ubd_handle h, int s, const uint8_t* d_mask, const float* d_cls, int cls_stride, int n_cls,
                       int n, int mh, int mw, int min_area_x2, int max_out, int32_t* labels_out_host) {
  ResultSlot& R = h->rs[s];
  const size_t npx = (size_t)mh * mw;
  const size_t pstride = (npx + 1 + 31) & ~(size_t)31;
  const int max_comps = h->opt_max_comps;
  const int max_pts = h->opt_max_points > 0 ? h->opt_max_points : (int)std::min<size_t>((size_t)n * npx / 2 + 1024, (size_t)1 << 26);
  // whole-image kernels (one CTA per image, shared memory): 2 = run-length variant, 1 = pixel variant; else the tiled path
  const bool rle = h->opt_fused_ccl == 2 && npx <= (size_t)CCL_IMG_MAX_PX && ccl_rle_smem(mh, mw) <= (size_t)227 * 1024;
  const bool fused = rle || (h->opt_fused_ccl == 1 && npx <= (size_t)CCL_IMG_MAX_PX);
  if (rle) ENSURE(h->run_label, (size_t)n * npx * 2 * sizeof(int));
  if (!fused) {
    ENSURE(h->parent, (size_t)n * pstride * sizeof(int));
    ENSURE(h->outer, (size_t)n * pstride);
  }
  ENSURE(h->labels, (size_t)n * npx * sizeof(int));
  ENSURE(h->slot_of, (size_t)n * npx * sizeof(int));
  ENSURE(h->comps, (size_t)n * max_comps * sizeof(CompRec));
  ENSURE(h->cls_sums, (size_t)n * max_comps * std::max(n_cls, 1) * sizeof(unsigned long long));
  ENSURE(h->out_index, (size_t)n * max_comps * sizeof(int));
  ENSURE(R.hdr, (size_t)(2 * n) * sizeof(int) + sizeof(CclTotals));
  ENSURE(R.out_recs, (size_t)std::max(max_out, 1) * sizeof(OutRec));
  // boxes on the GPU unless the map is too tall for the per-warp shared-memory arrays (then: hull candidates + host)
  const size_t box_smem = ((size_t)12 * mh + 10) * sizeof(int);
  R.gpu_boxes = h->opt_gpu_boxes && box_smem <= 48 * 1024;
  // row-extent arrays of the kept components: two ints per (component, row); every row holds at least one pixel of its
  // component and components do not share pixels, so the total cannot exceed the pixel count
  const int max_rows = (int)std::min<size_t>((size_t)n * npx, (size_t)1 << 28);
  if (R.gpu_boxes) {
    ENSURE(R.box_recs, (size_t)std::max(max_out, 1) * sizeof(BoxRec));
    ENSURE(h->row_ext, (size_t)max_rows * 2 * sizeof(int));
  }
  else ENSURE(R.hull_pts, (size_t)max_pts * sizeof(HullPt));
  const size_t hdr_ints = n + sizeof(CclTotals) / sizeof(int);
  if (R.h_hdr_cap < hdr_ints) {
    if (R.h_hdr) cudaFreeHost(R.h_hdr);
    R.h_hdr = nullptr; R.h_hdr_cap = 0;
    UBD_CUDA(cudaMallocHost(&R.h_hdr, (hdr_ints + 64) * sizeof(int)));
    R.h_hdr_cap = hdr_ints + 64;
  }
  if (!R.ev_cc) UBD_CUDA(cudaEventCreateWithFlags(&R.ev_cc, cudaEventDisableTiming));
  if (!R.ev_in) UBD_CUDA(cudaEventCreateWithFlags(&R.ev_in, cudaEventDisableTiming));
  // The CC stage is a chain of small latency-bound kernels: it runs on its own stream behind the producer of the mask
  // (the handle's stream), so that with two batches in flight it overlaps the next batch's stem (whose CTAs leave room
  // for small blocks on every SM).  The per-pixel workspaces are shared by both result slots: CC stages serialise on cs.
  cudaStream_t cs = h->opt_cc_stream ? h->cc_stream : h->stream;
  if (cs != h->stream) {
    UBD_CUDA(cudaEventRecord(R.ev_in, h->stream));
    UBD_CUDA(cudaStreamWaitEvent(cs, R.ev_in, 0));
  }
  R.n = n; R.mh = mh; R.mw = mw; R.max_pts = max_pts; R.max_comps_img = max_comps; R.max_out = max_out;
  R.d_mask_used = d_mask; R.d_cls_used = d_cls; R.cls_stride = cls_stride; R.n_cls = n_cls; R.min_area_x2 = min_area_x2;
  int* d_ncomps = (int*)R.hdr.p;
  int* d_kept = d_ncomps + n;
  CclTotals* d_tot = (CclTotals*)(d_kept + n);
  int* parent = (int*)h->parent.p;
  int* labels = (int*)h->labels.p;
  int* slot_of = (int*)h->slot_of.p;
  CompRec* comps = (CompRec*)h->comps.p;
  unsigned long long* cls_sums = (unsigned long long*)h->cls_sums.p;

  dim3 tgrid((mw + 31) / 32, (mh + 7) / 8, n), tblock(256);
  dim3 lgrid((unsigned)((npx + 1 + 255) / 256), n);
  {
    HostTimer ht_enq(h, 1);
    ProfScope ps_ccl(h, &h->prof_ccl, cs);
    UBD_CUDA(cudaMemsetAsync(d_tot, 0, sizeof(CclTotals), cs));
    if (rle) {
      ccl_rle_kernel<<<n, CCL_IMG_THREADS, ccl_rle_smem(mh, mw), cs>>>(d_mask, labels, slot_of, (int*)h->run_label.p, comps, d_cls, cls_stride,
                                                                             cls_sums, n_cls, d_ncomps, d_kept, d_tot, mh, mw, max_comps, min_area_x2);
      LAUNCH_CHECK();
    } else if (fused) {
      // one CTA per image, everything in shared memory (ubd_ccl.cuh, "whole-image variant")
      ccl_image_kernel<<<n, CCL_IMG_THREADS, ccl_image_smem((int)npx), cs>>>(d_mask, labels, slot_of, comps, d_cls, cls_stride, cls_sums, n_cls,
                                                                                   d_ncomps, d_kept, d_tot, mh, mw, max_comps, min_area_x2);
      LAUNCH_CHECK();
    } else {
      dim3 lgrid32((mw + CCL_T - 1) / CCL_T, (mh + CCL_T - 1) / CCL_T, n);
      ccl_local_kernel<<<lgrid32, 256, 0, cs>>>(d_mask, parent, mh, mw, pstride); LAUNCH_CHECK();
      const int n_border = ((mh - 1) / CCL_T) * mw + ((mw - 1) / CCL_T) * 2 * mh;
      if (n_border > 0) {
        dim3 bgrid2((unsigned)((n_border + 255) / 256), n);
        ccl_border_kernel<<<bgrid2, 256, 0, cs>>>(d_mask, parent, mh, mw, pstride); LAUNCH_CHECK();
      }
      uint8_t* outer = (uint8_t*)h->outer.p;
      ccl_flatten_kernel<<<lgrid, 256, 0, cs>>>(parent, outer, mh, mw, pstride); LAUNCH_CHECK();
      dim3 bgrid((unsigned)((2 * (mh + mw) + 255) / 256), n);
      ccl_mark_outer_kernel<<<bgrid, 256, 0, cs>>>(d_mask, parent, outer, mh, mw, pstride); LAUNCH_CHECK();
      ccl_merge2_kernel<<<tgrid, tblock, 0, cs>>>(d_mask, parent, outer, mh, mw, pstride); LAUNCH_CHECK();
      ccl_label_kernel<<<lgrid, 256, 0, cs>>>(d_mask, parent, outer, labels, mh, mw, pstride); LAUNCH_CHECK();
      ccl_slots_kernel<<<n, 1024, 0, cs>>>(labels, slot_of, comps, cls_sums, n_cls, d_ncomps, mh, mw, max_comps); LAUNCH_CHECK();
      dim3 sgrid((mw + 1 + 31) / 32, (mh + 1 + 7) / 8, n);
      ccl_stats_kernel<<<sgrid, tblock, 0, cs>>>(d_mask, labels, slot_of, comps, d_cls, cls_stride, cls_sums, n_cls, mh, mw, max_comps); LAUNCH_CHECK();
      ccl_count_kept_kernel<<<n, 256, 0, cs>>>(comps, d_ncomps, d_kept, d_tot, max_comps, min_area_x2); LAUNCH_CHECK();
    }
    ccl_compact_kernel<<<n, 256, 0, cs>>>(comps, cls_sums, n_cls, d_ncomps, d_kept, (OutRec*)R.out_recs.p,
                                                 (int*)h->out_index.p, max_comps, max_out, min_area_x2,
                                                 d_tot, R.gpu_boxes ? (int*)h->row_ext.p : nullptr, max_rows); LAUNCH_CHECK();
    if (R.gpu_boxes) {
      // hull + rotating calipers of every kept component on the GPU, one warp each
      ccl_extents_kernel<<<tgrid, tblock, 0, cs>>>(labels, slot_of, (int*)h->out_index.p, (const OutRec*)R.out_recs.p,
                                                          (int*)h->row_ext.p, mh, mw, max_comps); LAUNCH_CHECK();
      ccl_boxes_kernel<<<std::max(max_out, 1), 32, box_smem, cs>>>((const int*)h->row_ext.p, (const OutRec*)R.out_recs.p, d_tot,
                                                                          (BoxRec*)R.box_recs.p, mh, mw, max_out); LAUNCH_CHECK();
    } else {
      ccl_points_kernel<<<tgrid, tblock, 0, cs>>>(labels, slot_of, (int*)h->out_index.p, (HullPt*)R.hull_pts.p,
                                                         d_tot, mh, mw, max_comps, max_pts); LAUNCH_CHECK();
    }
  }
  // header: kept counts per image + totals (one small D2H into pinned memory); the records and hull points follow
  // in ccl_finish once their sizes are known
  UBD_CUDA(cudaMemcpyAsync(R.h_hdr, d_kept, hdr_ints * sizeof(int), cudaMemcpyDeviceToHost, cs));
  if (labels_out_host)
    UBD_CUDA(cudaMemcpyAsync(labels_out_host, labels, (size_t)n * npx * sizeof(int), cudaMemcpyDeviceToHost, cs));
  UBD_CUDA(cudaEventRecord(R.ev_cc, cs));
  return UBD_OK;
}

static constexpr int kCclRetry = 1;     // ccl_finish: an image had more raw components than slots; opt_max_comps was raised

static int ccl_finish(ubd_handle h, int s, ubd_component* comps_out, int max_out, int32_t* n_comps_per_image) {
  ResultSlot& R = h->rs[s];
  const int n = R.n;
  {
    HostTimer ht_s1(h, 2);
    UBD_CUDA(cudaEventSynchronize(R.ev_cc));
  }
  CclTotals tot;
  memcpy(&tot, R.h_hdr + n, sizeof(tot));
  if (tot.max_ncomp > R.max_comps_img) {
    // the reference's cv2 path takes any number of contours (utils.py:52): grow the slot table and redo the CC stage only
    int want = R.max_comps_img;
    while (want < tot.max_ncomp) want *= 2;
    h->opt_max_comps = want;
    return kCclRetry;
  }
  if (tot.total_kept > max_out)
    UBD_FAIL(UBD_ERR_OVERFLOW, std::to_string(tot.total_kept) + " kept components exceed the caller's capacity " + std::to_string(max_out));
  if (!R.gpu_boxes && tot.total_pts > R.max_pts)
    UBD_FAIL(UBD_ERR_OVERFLOW, std::to_string(tot.total_pts) + " hull candidate points exceed max_points " + std::to_string(R.max_pts));
  for (int i = 0; i < n; ++i) n_comps_per_image[i] = R.h_hdr[i];
  if (tot.total_kept == 0) return UBD_OK;
  std::vector<OutRec>& recs = R.recs;
  std::vector<HullPt>& pts = R.pts;
  std::vector<BoxRec>& boxes = R.boxes;
  {
    HostTimer ht_s2(h, 3);
    recs.resize(tot.total_kept);
    pts.resize(R.gpu_boxes ? 0 : tot.total_pts);
    boxes.resize(R.gpu_boxes ? tot.total_kept : 0);
    // on a stream of their own: the handle's stream, the CC stream and the mask / logits read-back stream may already hold
    // work of the next batches (which would make this wait for THEIR network to finish)
    UBD_CUDA(cudaMemcpyAsync(recs.data(), R.out_recs.p, recs.size() * sizeof(OutRec), cudaMemcpyDeviceToHost, h->rec_stream));
    if (!pts.empty())
      UBD_CUDA(cudaMemcpyAsync(pts.data(), R.hull_pts.p, pts.size() * sizeof(HullPt), cudaMemcpyDeviceToHost, h->rec_stream));
    if (!boxes.empty())
      UBD_CUDA(cudaMemcpyAsync(boxes.data(), R.box_recs.p, boxes.size() * sizeof(BoxRec), cudaMemcpyDeviceToHost, h->rec_stream));
    UBD_CUDA(cudaStreamSynchronize(h->rec_stream));
  }
  HostTimer ht_host(h, 4);
  auto fill = [&](int i) -> ubd_component& {
    const OutRec& r = recs[i];
    ubd_component& c = comps_out[i];
    c.image = r.image; c.label = r.label; c.xmin = r.xmin; c.ymin = r.ymin; c.xmax = r.xmax; c.ymax = r.ymax;
    c.n_pixels = r.n_pixels; c.n_filled = r.n_filled; c.area_x2 = r.area_x2; c.class_id = r.class_id;
    return c;
  };
  if (R.gpu_boxes) {
    // the GPU left centre, size and first edge vector of every rectangle: only cv2.boxPoints' trigonometry remains
    for (int i = 0; i < tot.total_kept; ++i) {
      const BoxRec& b = boxes[i];
      ubd_box_from_device(b.cx, b.cy, b.w, b.h, b.ax, b.ay, b.n_hull, b.x0, b.y0, b.x1, b.y1, fill(i).box);
    }
    return UBD_OK;
  }
  // Reduce the hull candidates to the leftmost / rightmost one per (component, row) -- only those can
  // be hull vertices -- then the min-area box of each component on the host.
  std::vector<int>& row0 = R.row0;
  std::vector<int>& ext = R.ext;
  row0.assign(tot.total_kept + 1, 0);
  for (int i = 0; i < tot.total_kept; ++i) row0[i + 1] = row0[i] + (recs[i].ymax - recs[i].ymin + 1);
  ext.resize(2 * (size_t)row0[tot.total_kept]);
  for (size_t k = 0; k < ext.size(); k += 2) { ext[k] = 0x7fffffff; ext[k + 1] = -1; }
  for (const HullPt& p : pts) {
    const int x = p.xy & 0xffff, y = p.xy >> 16;
    int* e = &ext[2 * (size_t)(row0[p.comp] + y - recs[p.comp].ymin)];
    if (x < e[0]) e[0] = x;
    if (x > e[1]) e[1] = x;
  }
  std::vector<int32_t>& xy = R.xy;
  for (int i = 0; i < tot.total_kept; ++i) {
    const OutRec& r = recs[i];
    ubd_component& c = fill(i);
    xy.clear();
    for (int y = r.ymin; y <= r.ymax; ++y) {
      const int* e = &ext[2 * (size_t)(row0[i] + y - r.ymin)];
      if (e[1] < 0) continue;
      xy.push_back(e[0]); xy.push_back(y);
      if (e[1] != e[0]) { xy.push_back(e[1]); xy.push_back(y); }
    }
    ubd_min_area_box(xy.data(), (int)(xy.size() / 2), c.box);
  }
  return UBD_OK;
}

// CC stage of a batch, synchronously (slot 0); retried with a larger slot table when an image overflows it.
static int ccl_device(ubd_handle h, const uint8_t* d_mask, const float* d_cls, int cls_stride, int n_cls,
                      int n, int mh, int mw, int min_area_x2, int32_t* labels_out_host,
                      ubd_component* comps_out, int max_out, int32_t* n_comps_per_image) {
  for (const ResultSlot& R : h->rs)
    if (R.busy) UBD_FAIL(UBD_ERR_STATE, "a submitted batch is still in flight (ubd_segment_wait)");
  for (;;) {
    int rc = ccl_enqueue(h, 0, d_mask, d_cls, cls_stride, n_cls, n, mh, mw, min_area_x2, max_out, labels_out_host);
    if (rc) return rc;
    rc = ccl_finish(h, 0, comps_out, max_out, n_comps_per_image);
    if (rc != kCclRetry) return rc;
  }
}

// ------------------------------------------------------------------------------------------------
// public inference entry points
// ------------------------------------------------------------------------------------------------

extern "C" int ubd_forward_dev(ubd_handle h, const void* d_images, int in_dtype, int n, int H, int W, int preproc, float* d_logits) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_image_args(h, d_images, in_dtype, n, H, W, preproc);
  if (rc) return rc;
  if (!d_logits) UBD_FAIL(UBD_ERR_ARG, "d_logits is NULL");
  UBD_CUDA(cudaSetDevice(h->device));
  return forward_device(h, d_images, in_dtype, n, H, W, preproc, d_logits, nullptr, 0.f);
}

extern "C" int ubd_forward(ubd_handle h, const void* images, int in_dtype, int n, int H, int W, int preproc, float* logits_out) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_image_args(h, images, in_dtype, n, H, W, preproc);
  if (rc) return rc;
  if (!logits_out) UBD_FAIL(UBD_ERR_ARG, "logits_out is NULL");
  UBD_CUDA(cudaSetDevice(h->device));
  const size_t ib = image_bytes(h, in_dtype, n, H, W);
  const size_t lb = (size_t)n * (H / 4) * (W / 4) * h->spec.n_out * sizeof(float);
  ENSURE(h->d_images, ib);
  ENSURE(h->d_logits, lb);
  rc = forward_device(h, h->d_images.p, in_dtype, n, H, W, preproc, (float*)h->d_logits.p, nullptr, 0.f, images);
  if (rc) return rc;
  UBD_CUDA(cudaMemcpyAsync(logits_out, h->d_logits.p, lb, cudaMemcpyDeviceToHost, h->stream));
  UBD_CUDA(cudaStreamSynchronize(h->stream));
  return h->precision == UBD_FP32 ? UBD_OK : tc_check_error(h);
}

static int segment_common(ubd_handle h, const void* d_img, int in_dtype, int n, int H, int W, int preproc,
                          float logit_thr, int min_area_x2, uint8_t* d_mask, float* d_logits,
                          int32_t* labels_out_host, ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image) {
  const int n_cls = h->n_classes;
  int rc = forward_device(h, d_img, in_dtype, n, H, W, preproc, d_logits, d_mask, logit_thr);
  if (rc) return rc;
  rc = ccl_device(h, d_mask, n_cls ? d_logits + 1 : nullptr, h->spec.n_out, n_cls, n, H / 4, W / 4, min_area_x2,
                    labels_out_host, comps_out, max_comps, n_comps_per_image);
  if (rc) return rc;
  return h->precision == UBD_FP32 ? UBD_OK : tc_check_error(h);
}

extern "C" int ubd_segment_dev(ubd_handle h, const void* d_images, int in_dtype, int n, int H, int W, int preproc,
                               float logit_thr, int min_area_x2, uint8_t* d_mask, float* d_logits,
                               ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_image_args(h, d_images, in_dtype, n, H, W, preproc);
  if (rc) return rc;
  if (!comps_out || !n_comps_per_image || max_comps < 0) UBD_FAIL(UBD_ERR_ARG, "component outputs are NULL");
  UBD_CUDA(cudaSetDevice(h->device));
  const size_t q = (size_t)n * (H / 4) * (W / 4);
  if (!d_mask) { ENSURE(h->d_mask, q); d_mask = (uint8_t*)h->d_mask.p; }
  if (!d_logits) { ENSURE(h->d_logits, q * h->spec.n_out * sizeof(float)); d_logits = (float*)h->d_logits.p; }
  return segment_common(h, d_images, in_dtype, n, H, W, preproc, logit_thr, min_area_x2, d_mask, d_logits,
                        nullptr, comps_out, max_comps, n_comps_per_image);
}

extern "C" int ubd_segment(ubd_handle h, const void* images, int in_dtype, int n, int H, int W, int preproc,
                           float logit_thr, int min_area_x2, uint8_t* mask_out, float* logits_out, int32_t* labels_out,
                           ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_image_args(h, images, in_dtype, n, H, W, preproc);
  if (rc) return rc;
  if (!comps_out || !n_comps_per_image || max_comps < 0) UBD_FAIL(UBD_ERR_ARG, "component outputs are NULL");
  UBD_CUDA(cudaSetDevice(h->device));
  const size_t ib = image_bytes(h, in_dtype, n, H, W);
  const size_t q = (size_t)n * (H / 4) * (W / 4);
  ENSURE(h->d_images, ib);
  ENSURE(h->d_mask, q);
  ENSURE(h->d_logits, q * h->spec.n_out * sizeof(float));
  // the mask / logits copies are queued before the component read-back so they overlap it
  rc = forward_device(h, h->d_images.p, in_dtype, n, H, W, preproc, (float*)h->d_logits.p, (uint8_t*)h->d_mask.p, logit_thr, images);
  if (rc) return rc;
  if (mask_out || logits_out) {
    // the result copies run on the copy stream so that they overlap the CC kernels
    if (h->copy_events.empty()) {
      cudaEvent_t ev;
      UBD_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      h->copy_events.push_back(ev);
    }
    UBD_CUDA(cudaEventRecord(h->copy_events[0], h->stream));
    UBD_CUDA(cudaStreamWaitEvent(h->copy_stream, h->copy_events[0], 0));
    if (mask_out) UBD_CUDA(cudaMemcpyAsync(mask_out, h->d_mask.p, q, cudaMemcpyDeviceToHost, h->copy_stream));
    if (logits_out) UBD_CUDA(cudaMemcpyAsync(logits_out, h->d_logits.p, q * h->spec.n_out * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
  }
  const int n_cls = h->n_classes;
  rc = ccl_device(h, (uint8_t*)h->d_mask.p, n_cls ? (float*)h->d_logits.p + 1 : nullptr, h->spec.n_out, n_cls, n, H / 4, W / 4,
                  min_area_x2, labels_out, comps_out, max_comps, n_comps_per_image);
  if (rc) return rc;
  if (mask_out || logits_out) UBD_CUDA(cudaStreamSynchronize(h->copy_stream));
  return h->precision == UBD_FP32 ? UBD_OK : tc_check_error(h);
}

// ---- pipelined form of ubd_segment / ubd_segment_dev -------------------------------------------------------------
// submit enqueues everything of one batch (H2D of the images on the copy stream, network, threshold, CC kernels, the
// D2H of mask / logits on the read-back stream) and returns; wait blocks until that batch is done and finishes its
// boxes on the host.  With two batches in flight the H2D copy of batch k+1 and the host tail of batch k run under
// the kernels of the other batch.
static int segment_submit_common(ubd_handle h, const void* images, bool on_device, int in_dtype, int n, int H, int W, int preproc,
                                 float logit_thr, int min_area_x2, uint8_t* mask_out, float* logits_out, int max_comps, int* ticket) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_image_args(h, images, in_dtype, n, H, W, preproc);
  if (rc) return rc;
  if (!ticket || max_comps < 0) UBD_FAIL(UBD_ERR_ARG, "ticket is NULL or max_comps < 0");
  UBD_CUDA(cudaSetDevice(h->device));
  const int s = (int)(h->next_ticket % kSlots);
  ResultSlot& R = h->rs[s];
  if (R.busy) UBD_FAIL(UBD_ERR_STATE, std::to_string(kSlots) + " batches are already in flight: call ubd_segment_wait first");
  const size_t ib = image_bytes(h, in_dtype, n, H, W);
  const size_t q = (size_t)n * (H / 4) * (W / 4);
  const void* d_img = images;
  if (!on_device) { ENSURE(R.d_images, ib); d_img = R.d_images.p; }
  ENSURE(R.d_mask, q);
  ENSURE(R.d_logits, q * h->spec.n_out * sizeof(float));
  // with two batches already queued this batch's H2D copies finish under their kernels (measured on B200, 64 x 1024^2
  // tf32: 29.4 k img/s end to end at depth 3; at depth 2 the copy is exposed and the stem follows it in 16-image launches)
  int others = 0;
  for (const ResultSlot& O : h->rs) others += O.busy ? 1 : 0;
  rc = forward_device(h, d_img, in_dtype, n, H, W, preproc, (float*)R.d_logits.p, (uint8_t*)R.d_mask.p, logit_thr,
                      on_device ? nullptr : images, others >= 2);
  if (rc) return rc;
  if (mask_out || logits_out) {
    if (!R.ev_fwd) UBD_CUDA(cudaEventCreateWithFlags(&R.ev_fwd, cudaEventDisableTiming));
    UBD_CUDA(cudaEventRecord(R.ev_fwd, h->stream));
    UBD_CUDA(cudaStreamWaitEvent(h->d2h_stream, R.ev_fwd, 0));
    if (mask_out) UBD_CUDA(cudaMemcpyAsync(mask_out, R.d_mask.p, q, cudaMemcpyDeviceToHost, h->d2h_stream));
    if (logits_out) UBD_CUDA(cudaMemcpyAsync(logits_out, R.d_logits.p, q * h->spec.n_out * sizeof(float), cudaMemcpyDeviceToHost, h->d2h_stream));
    if (!R.ev_d2h) UBD_CUDA(cudaEventCreateWithFlags(&R.ev_d2h, cudaEventDisableTiming));
    UBD_CUDA(cudaEventRecord(R.ev_d2h, h->d2h_stream));
  }
  R.d2h_pending = mask_out || logits_out;
  const int n_cls = h->n_classes;
  rc = ccl_enqueue(h, s, (uint8_t*)R.d_mask.p, n_cls ? (float*)R.d_logits.p + 1 : nullptr, h->spec.n_out, n_cls, n, H / 4, W / 4,
                   min_area_x2, max_comps, nullptr);
  if (rc) return rc;
  R.busy = true;
  R.ticket = h->next_ticket++;
  *ticket = (int)(R.ticket & 0x7fffffff);
  return UBD_OK;
}

extern "C" int ubd_segment_submit(ubd_handle h, const void* images, int in_dtype, int n, int H, int W, int preproc,
                                  float logit_thr, int min_area_x2, uint8_t* mask_out, float* logits_out, int max_comps, int* ticket) {
  return segment_submit_common(h, images, false, in_dtype, n, H, W, preproc, logit_thr, min_area_x2, mask_out, logits_out, max_comps, ticket);
}

extern "C" int ubd_segment_submit_dev(ubd_handle h, const void* d_images, int in_dtype, int n, int H, int W, int preproc,
                                      float logit_thr, int min_area_x2, int max_comps, int* ticket) {
  return segment_submit_common(h, d_images, true, in_dtype, n, H, W, preproc, logit_thr, min_area_x2, nullptr, nullptr, max_comps, ticket);
}

extern "C" int ubd_segment_wait(ubd_handle h, int ticket, ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image) {
  if (!h) return UBD_ERR_ARG;
  if (!comps_out || !n_comps_per_image) UBD_FAIL(UBD_ERR_ARG, "component outputs are NULL");
  UBD_CUDA(cudaSetDevice(h->device));
  int s = -1;
  for (int i = 0; i < kSlots; ++i)
    if (h->rs[i].busy && (int)(h->rs[i].ticket & 0x7fffffff) == ticket) s = i;
  if (s < 0) UBD_FAIL(UBD_ERR_STATE, "unknown ticket");
  ResultSlot& R = h->rs[s];
  // tickets complete in submission order: the older batch must be collected first
  for (const ResultSlot& O : h->rs)
    if (O.busy && O.ticket < R.ticket) UBD_FAIL(UBD_ERR_STATE, "an older batch has not been collected yet");
  if (max_comps < R.max_out) UBD_FAIL(UBD_ERR_ARG, "comps_out is smaller than the capacity given to ubd_segment_submit");
  int rc;
  for (;;) {
    rc = ccl_finish(h, s, comps_out, R.max_out, n_comps_per_image);
    if (rc != kCclRetry) break;
    // rare: more raw components in an image than slots; redo the CC stage of this batch behind whatever is queued
    rc = ccl_enqueue(h, s, R.d_mask_used, R.d_cls_used, R.cls_stride, R.n_cls, R.n, R.mh, R.mw, R.min_area_x2, R.max_out, nullptr);
    if (rc) break;
  }
  R.busy = false;
  if (rc) return rc;
  // mask / logits of THIS batch have landed (the stream itself may already hold the next batches' copies, which wait
  // for their networks)
  if (R.d2h_pending) UBD_CUDA(cudaEventSynchronize(R.ev_d2h));
  return h->precision == UBD_FP32 ? UBD_OK : tc_check_error_on(h, h->rec_stream);
}

extern "C" int ubd_postprocess(ubd_handle h, const uint8_t* mask, const float* cls_logits, int n, int mh, int mw,
                               int n_cls, int min_area_x2, int32_t* labels_out,
                               ubd_component* comps_out, int max_comps, int32_t* n_comps_per_image) {
  if (!h) return UBD_ERR_ARG;
  if (!mask || !comps_out || !n_comps_per_image) UBD_FAIL(UBD_ERR_ARG, "NULL argument");
  if (n < 1 || mh < 1 || mw < 1 || mh > 65535 || mw > 65535) UBD_FAIL(UBD_ERR_ARG, "bad map shape");
  if (n_cls < 0 || n_cls > UBD_MAX_CLASSES || (n_cls > 0 && !cls_logits)) UBD_FAIL(UBD_ERR_ARG, "bad class logits");
  UBD_CUDA(cudaSetDevice(h->device));
  const size_t q = (size_t)n * mh * mw;
  h->loss_pixels = 0; h->last_logits = nullptr; h->last_ytrue = nullptr;      // d_logits is reused below
  ENSURE(h->d_mask, q);
  UBD_CUDA(cudaMemcpyAsync(h->d_mask.p, mask, q, cudaMemcpyHostToDevice, h->stream));
  if (n_cls > 0) {
    ENSURE(h->d_logits, q * n_cls * sizeof(float));
    UBD_CUDA(cudaMemcpyAsync(h->d_logits.p, cls_logits, q * n_cls * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  return ccl_device(h, (uint8_t*)h->d_mask.p, n_cls ? (float*)h->d_logits.p : nullptr, n_cls, n_cls, n, mh, mw,
                    min_area_x2, labels_out, comps_out, max_comps, n_comps_per_image);
}

// ------------------------------------------------------------------------------------------------
// input side: Pillow-exact bicubic resize + convert('L') on the GPU
// ------------------------------------------------------------------------------------------------

// Pillow's precompute_coeffs + normalize_coeffs_8bpc (src/libImaging/Resample.c) for the bicubic filter.
static int prep_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
  auto bicubic = [](double x) -> double {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
  };
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  bounds.assign(2 * (size_t)out_size, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < xmax; ++x) {
      const double v = k[x] * (1 << 22);
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v) : (int)(0.5 + v);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

static int prepare_common(ubd_handle h, const uint8_t* d_src, int n, int H, int W, int C, int out_h, int out_w, int to_grey,
                          uint8_t* d_dst) {
  const bool need_h = out_w != W, need_v = out_h != H, grey = to_grey && C == 3;
  const int Co = grey ? 1 : C;
  std::vector<int> bx, kx, by, ky;
  int ksx = 0, ksy = 0;
  if (need_h) ksx = prep_coeffs(W, out_w, bx, kx);
  if (need_v) ksy = prep_coeffs(H, out_h, by, ky);
  const size_t tab = (bx.size() + kx.size() + by.size() + ky.size()) * sizeof(int);
  ENSURE(h->prep_tab, std::max<size_t>(tab, 16));
  int* t = (int*)h->prep_tab.p;
  std::vector<int> all;
  all.insert(all.end(), bx.begin(), bx.end()); all.insert(all.end(), kx.begin(), kx.end());
  all.insert(all.end(), by.begin(), by.end()); all.insert(all.end(), ky.begin(), ky.end());
  if (!all.empty()) {
    UBD_CUDA(cudaMemcpyAsync(t, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    UBD_CUDA(cudaStreamSynchronize(h->stream));                    // `all` is a local
  }
  PrepAxis ax{t, t + bx.size(), ksx};
  PrepAxis ay{t + bx.size() + kx.size(), t + bx.size() + kx.size() + by.size(), ksy};
  // passes: horizontal (if the width changes), vertical (if the height changes), luma; intermediates in prep_a / prep_b
  const size_t sz_h = (size_t)n * H * out_w * C, sz_v = (size_t)n * out_h * out_w * C;
  const uint8_t* cur = d_src;
  auto blocks = [&](size_t total) { return (unsigned)std::min<size_t>((total + 255) / 256, (size_t)h->n_sm * 16); };
  if (need_h) {
    uint8_t* dst = d_dst;
    if (need_v || grey) { ENSURE(h->prep_a, sz_h); dst = (uint8_t*)h->prep_a.p; }
    prep_resize_h_kernel<<<blocks(sz_h), 256, 0, h->stream>>>(cur, dst, ax, n, H, W, out_w, C); LAUNCH_CHECK();
    cur = dst;
  }
  if (need_v) {
    uint8_t* dst = d_dst;
    if (grey) { ENSURE(h->prep_b, sz_v); dst = (uint8_t*)h->prep_b.p; }
    prep_resize_v_kernel<<<blocks(sz_v), 256, 0, h->stream>>>(cur, dst, ay, n, H, out_h, out_w, C); LAUNCH_CHECK();
    cur = dst;
  }
  if (grey) {
    const size_t npx = (size_t)n * out_h * out_w;
    prep_rgb2l_kernel<<<blocks(npx), 256, 0, h->stream>>>(cur, d_dst, npx); LAUNCH_CHECK();
  } else if (cur == d_src) {
    UBD_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t)n * out_h * out_w * Co, cudaMemcpyDeviceToDevice, h->stream));
  }
  return UBD_OK;
}

static int check_prepare_args(ubd_handle h, const void* images, void* out, int n, int H, int W, int C, int out_h, int out_w) {
  if (!images || !out) UBD_FAIL(UBD_ERR_ARG, "NULL argument");
  if (n < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1 || (C != 1 && C != 3)) UBD_FAIL(UBD_ERR_ARG, "bad image shape (channels must be 1 or 3)");
  return UBD_OK;
}

extern "C" int ubd_prepare_images_dev(ubd_handle h, const uint8_t* d_images, int n, int H, int W, int C, int out_h, int out_w,
                                      int to_grey, uint8_t* d_out) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_prepare_args(h, d_images, d_out, n, H, W, C, out_h, out_w);
  if (rc) return rc;
  UBD_CUDA(cudaSetDevice(h->device));
  return prepare_common(h, d_images, n, H, W, C, out_h, out_w, to_grey, d_out);
}

extern "C" int ubd_prepare_images(ubd_handle h, const uint8_t* images, int n, int H, int W, int C, int out_h, int out_w,
                                  int to_grey, uint8_t* out) {
  if (!h) return UBD_ERR_ARG;
  int rc = check_prepare_args(h, images, out, n, H, W, C, out_h, out_w);
  if (rc) return rc;
  UBD_CUDA(cudaSetDevice(h->device));
  const size_t ib = (size_t)n * H * W * C, ob = (size_t)n * out_h * out_w * ((to_grey && C == 3) ? 1 : C);
  ENSURE(h->prep_in, ib);
  ENSURE(h->prep_out, ob);
  UBD_CUDA(cudaMemcpyAsync(h->prep_in.p, images, ib, cudaMemcpyHostToDevice, h->stream));
  rc = prepare_common(h, (const uint8_t*)h->prep_in.p, n, H, W, C, out_h, out_w, to_grey, (uint8_t*)h->prep_out.p);
  if (rc) return rc;
  UBD_CUDA(cudaMemcpyAsync(out, h->prep_out.p, ob, cudaMemcpyDeviceToHost, h->stream));
  UBD_CUDA(cudaStreamSynchronize(h->stream));
  return UBD_OK;
}

// Test hook: one dilated layer (0..5, its own weights and dilation) on an NHWC 24-channel host map,
// through the FP32 kernel or the tcgen05 kernel, for layer-level parity tests and bring-up.
__global__ void nhwc_to_planar_kernel(const float* __restrict__ src, float4* __restrict__ dst, int n, int hh, int ww) {
  const size_t npx = (size_t)hh * ww;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * npx) return;
  const size_t img = i / npx, p = i % npx;
  for (int g = 0; g < UBD_NG; ++g) {
    const float* s = src + i * UBD_NF + 4 * g;
    dst[act_index((int)img, g, (int)(p / ww), (int)(p % ww), hh, ww, UBD_MAP_PAD)] = make_float4(s[0], s[1], s[2], s[3]);
  }
}
__global__ void planar_to_nhwc_kernel(const float4* __restrict__ src, float* __restrict__ dst, int n, int hh, int ww) {
  const size_t npx = (size_t)hh * ww;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * npx) return;
  const size_t img = i / npx, p = i % npx;
  for (int g = 0; g < UBD_NG; ++g) {
    const float4 v = src[act_index((int)img, g, (int)(p / ww), (int)(p % ww), hh, ww, UBD_MAP_PAD)];
    float* d = dst + i * UBD_NF + 4 * g;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

__global__ void nhwc_to_planar_bf16_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int n, int hh, int ww, bool f16) {
  const size_t npx = (size_t)hh * ww;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * npx) return;
  const size_t img = i / npx, p = i % npx;
  const int y = (int)(p / ww), x = (int)(p % ww);
  const size_t wp = ww + 2 * UBD_MAP_PAD;
  for (int g = 0; g < 3; ++g) {
    const float* s = src + i * UBD_NF + 8 * g;
    dst[((img * hh + y) * 3 + g) * wp + UBD_MAP_PAD + x] =
        make_uint4(tc::pack16(s[0], s[1], f16), tc::pack16(s[2], s[3], f16), tc::pack16(s[4], s[5], f16), tc::pack16(s[6], s[7], f16));
  }
}

extern "C" int ubd_debug_dilated_layer(ubd_handle h, const float* in_nhwc, float* out_nhwc, int layer,
                                       int n, int mh, int mw, int precision) {
  if (!h) return UBD_ERR_ARG;
  if (!in_nhwc || !out_nhwc || layer < 0 || layer >= UBD_NLAYERS_DIL || n < 1 || mh < 1 || mw < 1)
    UBD_FAIL(UBD_ERR_ARG, "bad argument");
  if (!h->have_weights) UBD_FAIL(UBD_ERR_NO_WEIGHTS, "no weights loaded");
  UBD_CUDA(cudaSetDevice(h->device));
  const size_t elems = (size_t)n * mh * mw * UBD_NF;
  ENSURE(h->t_scratch, elems * sizeof(float));
  { int rc_ = ensure_maps(h, n, mh, mw); if (rc_) return rc_; }
  UBD_CUDA(cudaMemcpyAsync(h->t_scratch.p, in_nhwc, elems * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  const unsigned blocks = (unsigned)(((size_t)n * mh * mw + 255) / 256);
  if (precision == UBD_BF16 || precision == UBD_F16) {
    // bf16 input layout shares the buffer with fp32 layouts of earlier calls: clear the pads
    UBD_CUDA(cudaMemsetAsync(h->mapA.p, 0, h->mapA.cap, h->stream));
    UBD_CUDA(cudaMemsetAsync(h->mapB.p, 0, h->mapB.cap, h->stream));
    h->map_prec = -1;
    nhwc_to_planar_bf16_kernel<<<blocks, 256, 0, h->stream>>>((const float*)h->t_scratch.p, (uint4*)h->mapA.p, n, mh, mw, precision == UBD_F16); LAUNCH_CHECK();
  } else {
    if (h->map_prec == UBD_BF16 || h->map_prec == UBD_F16 || h->map_prec == -1) {
      UBD_CUDA(cudaMemsetAsync(h->mapA.p, 0, h->mapA.cap, h->stream));
      UBD_CUDA(cudaMemsetAsync(h->mapB.p, 0, h->mapB.cap, h->stream));
      h->map_prec = UBD_FP32;
    }
    nhwc_to_planar_kernel<<<blocks, 256, 0, h->stream>>>((const float*)h->t_scratch.p, (float4*)h->mapA.p, n, mh, mw); LAUNCH_CHECK();
  }
  const float* w = h->d_params + h->spec.off[9 + 2 * layer];
  const float* b = h->d_params + h->spec.off[10 + 2 * layer];
  int rc;
  if (precision == UBD_FP32) {
    rc = launch_dil_fp32(h, (float4*)h->mapA.p, (float4*)h->mapB.p, w, b, nullptr, n, mh, mw, kDilations[layer], 0);
  } else {
    const int saved = h->precision;
    h->precision = precision;
    rc = launch_dil_tc(h, h->mapA.p, h->mapB.p, layer, n, mh, mw, kDilations[layer], /*out_mode=*/1);
    h->precision = saved;
  }
  if (rc) return rc;
  planar_to_nhwc_kernel<<<blocks, 256, 0, h->stream>>>((const float4*)h->mapB.p, (float*)h->t_scratch.p, n, mh, mw); LAUNCH_CHECK();
  UBD_CUDA(cudaMemcpyAsync(out_nhwc, h->t_scratch.p, elems * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  UBD_CUDA(cudaStreamSynchronize(h->stream));
  return precision == UBD_FP32 ? UBD_OK : tc_check_error(h);
}

extern "C" int ubd_debug_read_trace(ubd_handle h, long long* out, int n_values) {
  if (!h || !out || !h->tc_trace.p) return UBD_ERR_ARG;
  UBD_CUDA(cudaSetDevice(h->device));
  UBD_CUDA(cudaStreamSynchronize(h->stream));
  UBD_CUDA(cudaMemcpy(out, h->tc_trace.p, std::min<size_t>((size_t)n_values * 8, 8 * 1024 * 4 * 8), cudaMemcpyDeviceToHost));
  UBD_CUDA(cudaMemset(h->tc_trace.p, 0, h->tc_trace.cap));
  return UBD_OK;
}

#include "ubd_train_api.inc"
