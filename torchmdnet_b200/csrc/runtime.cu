// Error reporting + version for libtmdnet_b200.so (C ABI, no torch types anywhere).
#include "common.cuh"
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

void tmd_set_error(const char* msg) { strncpy(g_err, msg, sizeof(g_err) - 1); g_err[sizeof(g_err) - 1] = 0; }

int tmd_check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return TMD_ERR_CUDA;
  }
  return TMD_OK;
}

extern "C" {
const char* tmd_last_error(void) { return g_err; }
int tmd_abi_version(void) { return 1; }
// number of kernel launches issued through this library since load (bench.py's gpu_launches)
}
