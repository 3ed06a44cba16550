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

// Library-wide device status word (sticky flags raised by kernels, read by cgat_status_flags — the one entry point
// that synchronises).  Kernels get the pointer as an argument: device symbols do not link across translation units.
enum StatusBits : unsigned int {
  kStatusBadDestination = 1u,   // cgat_csr_build: edge_index[1] outside [0, n_nodes)  (the edge is dropped)
  kStatusBadSource = 2u,        // cgat_csr_build: edge_index[0] outside [0, n_nodes)  (clamped to 0)
  kStatusBadRank = 4u,          // cgat_csr_build: edge_attr outside [0, n_ranks)       (clamped to 0)
  kStatusNonFinite = 8u,        // cgat_edge_attn_fwd*: a non-finite aggregate (fp16 operand overflow, or NaN/Inf input)
};
unsigned int* status_word();   // device address of the word (csr_build.cu)

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// hyper-linear kernels: output channels per work item — smaller chunks when there are few atom tiles, so that all
// SMs get work (shared by the tf32 and f16 forms: cgat_hyper_rowscale_parts sizes the partial buffer of both)
inline int hyper_chunk(int64_t n_atoms, int f) {
  const int n_tiles = (int)((n_atoms + 127) / 128);
  int oc = 16;
  while (oc > 4 && (int64_t)n_tiles * (f / oc) < 3 * kNumSMs) oc >>= 1;
  return oc;
}

// Partial-result slots of the hyper activation-gradient kernels (cgat_hyper_rowscale[_f16]).  Work items are
// (atom tile, chunk of output channels), chunk fastest, dealt to the CTAs as contiguous ranges; a CTA accumulates the
// consecutive chunks of one tile in registers and writes ONE partial row block per (CTA, tile) into slot
// (this CTA) - (first CTA that touches the tile).  hyper_slots() bounds that slot count: with >= per = floor(items /
// grid) items per CTA a tile's n_chunks items meet at most ceil(n_chunks / per) + 1 CTAs.  (Round 1 wrote one partial
// per chunk: 16 x N x F floats per launch at the bench size instead of 5, all re-read by the partial sum.)
#ifndef CGAT_HYPER_CLUSTER
#define CGAT_HYPER_CLUSTER 2
#endif
constexpr int kHyperCluster = CGAT_HYPER_CLUSTER;   // CTAs per cluster of the f16 hyper kernels: they share every weight stage by multicast

inline int64_t hyper_items(int64_t n_atoms, int f, int mode, int cs = 1) {
  // forward (mode 0): (atom tile group, output chunk); backward (mode 1): ((tile group, 128-column half), output chunk);
  // a tile group = the `cs` consecutive atom tiles that the CTAs of one cluster work on side by side
  const int64_t n_tg = ((n_atoms + 127) / 128 + cs - 1) / cs;
  return n_tg * (mode ? f / 128 : 1) * (f / hyper_chunk(n_atoms, f));
}
// number of CTAs (cs = 1) or clusters (cs > 1) of the launch
inline int hyper_grid(int64_t n_atoms, int f, int mode = 0, int cs = 1) {
  const int64_t n_items = hyper_items(n_atoms, f, mode, cs);
  const int cap = kNumSMs / cs;
  return (int)(n_items < cap ? (n_items > 0 ? n_items : 1) : cap);
}
inline int hyper_slots(int64_t n_atoms, int f, int cs = 1) {
  const int n_chunks = f / hyper_chunk(n_atoms, f);
  const int64_t per = hyper_items(n_atoms, f, 1, cs) / hyper_grid(n_atoms, f, 1, cs);
  const int64_t s = per > 0 ? (n_chunks + per - 1) / per + 1 : n_chunks;
  return (int)(s < n_chunks ? s : n_chunks);
}
// partial-result slots the caller allocates: enough for the tf32 kernel (no clusters) and the f16 kernel (clusters)
inline int hyper_parts(int64_t n_atoms, int f) {
  const int a = hyper_slots(n_atoms, f, 1), b = hyper_slots(n_atoms, f, kHyperCluster);
  return a > b ? a : b;
}
// first CTA whose item range [n_items*b/G, n_items*(b+1)/G) contains item i
__host__ __device__ inline int hyper_cta_of_item(int64_t i, int64_t n_items, int G) {
  int b = (int)((i * G) / n_items);
  while (b + 1 < G && (n_items * (b + 1)) / G <= i) ++b;
  while (b > 0 && (n_items * b) / G > i) --b;
  return b;
}

}  // namespace cgat
