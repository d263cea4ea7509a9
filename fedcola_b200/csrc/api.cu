// Library-wide entry points of libfedcola_b200.so.
#include "common.cuh"
#include "../../include/fedcola_b200.h"

thread_local char fc_last_error_buf[512] = {0};
unsigned long long fc_launch_counter = 0;

extern "C" unsigned long long fc_launch_count(void) { return __atomic_load_n(&fc_launch_counter, __ATOMIC_RELAXED); }

int fc_grid_cap_value = 0;
extern "C" int fc_set_grid_cap(int max_ctas) {
  const int prev = __atomic_exchange_n(&fc_grid_cap_value, max_ctas > 0 ? max_ctas : 0, __ATOMIC_RELAXED);
  return prev;
}

extern "C" const char* fc_last_error(void) { return fc_last_error_buf; }
extern "C" int fc_abi_version(void) { return FC_ABI_VERSION; }

// Struct-layout handshake for FFI bindings (ctypes / cgo stubs assert these at load time).
extern "C" int fc_sizeof_mat_desc(void) { return (int)sizeof(fc_mat_desc); }
extern "C" int fc_sizeof_step_args(void) { return (int)sizeof(fc_step_args); }
