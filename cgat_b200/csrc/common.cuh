// Shared helpers for the cgat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cgat_b200.h"

namespace cgat {

extern thread_local char g_err[512];
extern long long g_launches;

inline int fail(int code, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s (code %d)", what, code);
  return code ? code : -1;
}

inline int check_launch(const char* name) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", name, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define CGAT_CUDA(expr)                                                     \
  do {                                                                      \
    cudaError_t _e = (expr);                                                \
    if (_e != cudaSuccess) {                                                \
      snprintf(cgat::g_err, sizeof(cgat::g_err), "%s: %s", #expr,           \
               cudaGetErrorString(_e));                                     \
      return (int)_e;                                                       \
    }                                                                       \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// hyper-linear kernels: output channels per work item — smaller chunks when there are few atom tiles, so that all
// SMs get work (shared by the tf32 and f16 forms: cgat_hyper_rowscale_parts sizes the partial buffer of both)
inline int hyper_chunk(int64_t n_atoms, int f) {
  const int n_tiles = (int)((n_atoms + 127) / 128);
  int oc = 16;
  while (oc > 4 && (int64_t)n_tiles * (f / oc) < 3 * kNumSMs) oc >>= 1;
  return oc;
}

}  // namespace cgat
