// capi.cu -- library-level bookkeeping of libaxisym_b200.so
#include "axb_common.cuh"

int64_t g_axb_launches = 0;

extern "C" {
int axb_version(void) { return 100; }
int64_t axb_launch_count(void) { return g_axb_launches; }
}
