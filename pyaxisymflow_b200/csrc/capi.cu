// capi.cu -- library-level bookkeeping of libaxisym_b200.so
#include "axb_common.cuh"

int64_t g_axb_launches = 0;
int g_axb_legacy_stencils = 0;
int g_axb_solid_march = 0;
int g_axb_tri_one_warp = -1;   // -1: read AXB_TRI_ONE_WARP on first use

extern "C" {
int axb_version(void) { return 100; }
int64_t axb_launch_count(void) { return g_axb_launches; }
int axb_set_stencil_path(int legacy_tiled) {
  g_axb_legacy_stencils = legacy_tiled;
  return AXB_OK;
}
int axb_set_solid_march(int on) {
  g_axb_solid_march = on;
  return AXB_OK;
}
int axb_set_tridiag_sweep(int one_warp) {
  g_axb_tri_one_warp = one_warp ? 1 : 0;
  return AXB_OK;
}
}
