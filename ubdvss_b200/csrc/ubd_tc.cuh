// tcgen05 implicit-GEMM path for the dilated layers (filled in by ubd_tc_impl).
#pragma once
#include "ubd_handle.cuh"
static void tc_setup_attributes() {}
static int tc_launch_dilconv(ubd_handle h, const float4* in, float4* out, int layer, int n, int hh, int ww, int d) {
  (void)in; (void)out; (void)layer; (void)n; (void)hh; (void)ww; (void)d;
  h->err = "tensor-core path not built in this revision";
  return UBD_ERR_UNSUPPORTED;
}
