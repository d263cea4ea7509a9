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

// Peer (NVLink) access from `device` to memory allocated on `peer_device`: lets fc_aggregate read client arenas that
// were trained on another GPU of the same process (the reference's thread-per-client `cuda:(i % ngpu)` placement,
// /root/reference/src/server/fedavgserver.py:310-311) in place, keeping the sequential lerp bit-exact across GPUs.
// Returns FC_OK when access is (already) enabled, FC_ERR_UNSUPPORTED when the hardware path does not exist.
extern "C" int fc_enable_peer_access(int device, int peer_device) {
  if (device == peer_device) return FC_OK;
  int can = 0;
  FC_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, device, peer_device));
  if (!can) FC_FAIL(FC_ERR_UNSUPPORTED, "fc_enable_peer_access: device %d cannot access device %d", device, peer_device);
  FcDeviceGuard guard(device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();      // clear the sticky-free error state
    return FC_OK;
  }
  FC_CUDA_CHECK(e);
  return FC_OK;
}

// Host -> device gather of rows for one training batch: dst[i, :] = src_host[idx[i], :], issued as cudaMemcpyAsync on
// `stream` with runs of consecutive indices merged into one copy.  src_host should be pinned (the copies are then
// truly asynchronous); idx is a HOST array.  Replaces the DataLoader's per-item collation + pin + .to(device) of
// /root/reference/src/client/fedavgclient.py:79-84 when the client's set is held as one host tensor.
extern "C" int fc_h2d_rows(void* dst, const void* src_host, const long long* idx, int n, long long row_bytes,
                           int device, void* stream) {
  FC_REQUIRE(n >= 0 && row_bytes > 0, "fc_h2d_rows: bad sizes");
  if (n == 0) return FC_OK;
  FC_REQUIRE(dst != nullptr && src_host != nullptr && idx != nullptr, "fc_h2d_rows: null pointer");
  FcDeviceGuard guard(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* d = reinterpret_cast<char*>(dst);
  const char* s = reinterpret_cast<const char*>(src_host);
  int i = 0;
  while (i < n) {
    int j = i + 1;
    while (j < n && idx[j] == idx[j - 1] + 1) ++j;
    FC_REQUIRE(idx[i] >= 0, "fc_h2d_rows: negative row index");
    FC_CUDA_CHECK(cudaMemcpyAsync(d + (size_t)i * row_bytes, s + (size_t)idx[i] * row_bytes, (size_t)(j - i) * row_bytes,
                                  cudaMemcpyHostToDevice, st));
    i = j;
  }
  return FC_OK;
}
