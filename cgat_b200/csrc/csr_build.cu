// CSR-by-destination construction and segment pointers.
//
// The reference never builds this structure explicitly: torch_geometric's propagate gathers by
// edge_index and torch_scatter reduces with atomics on every call (reference CGAT/CGAT.py:313-326).
// Here the edge list is grouped by destination ONCE per batch so that every later kernel sees each
// softmax segment as a contiguous, deterministic range.  Integer work, bit-exact against
// torch.sort(stable=True) / bincount / cumsum: a stable counting sort whose within-segment order is
// the ascending original edge id.
#include "common.cuh"

namespace cgat {

thread_local char g_err[512] = "";
long long g_launches = 0;

__device__ unsigned int g_status_word = 0;

unsigned int* status_word() {
  static unsigned int* p = nullptr;
  if (p == nullptr) cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_status_word);
  return p;
}

namespace {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;  // items per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void histogram_kernel(const int64_t* __restrict__ dst, int64_t n_edges, int64_t n_nodes,
                                 int32_t* __restrict__ counts, int32_t* __restrict__ bad,
                                 unsigned int* __restrict__ status) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int64_t d = dst[e];
  if (d < 0 || d >= n_nodes) {
    atomicAdd(bad, 1);
    atomicOr(status, (unsigned int)kStatusBadDestination);
    return;
  }
  atomicAdd(&counts[d], 1);  // integer atomics: result independent of order
}

__device__ __forceinline__ int32_t block_exclusive_scan(int32_t v, int32_t* total, int32_t* smem) {
  // inclusive warp scan
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int32_t w = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    smem[lane] = w;
  }
  __syncthreads();
  int32_t warp_off = warp ? smem[warp - 1] : 0;
  *total = smem[(blockDim.x >> 5) - 1];
  __syncthreads();
  return warp_off + x - v;
}

// phase 1: per-tile sums
__global__ void scan_tile_sums(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ tile_sums) {
  __shared__ int32_t smem[32];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  int32_t total;
  block_exclusive_scan(s, &total, smem);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// phase 2: exclusive scan of tile sums by one block (loops if there are more tiles than threads)
__global__ void scan_tile_offsets(int32_t* __restrict__ tile_sums, int64_t n_tiles) {
  __shared__ int32_t smem[32];
  int32_t carry = 0;
  for (int64_t base = 0; base < n_tiles; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    int32_t v = i < n_tiles ? tile_sums[i] : 0;
    int32_t total;
    int32_t ex = block_exclusive_scan(v, &total, smem);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    carry += total;
  }
}

// phase 3: out[i] = exclusive prefix (the caller scans N+1 items with in[N] = 0, so out[N] = total)
__global__ void scan_apply(const int32_t* __restrict__ in, int64_t n, const int32_t* __restrict__ tile_offsets,
                           int32_t* __restrict__ out) {
  __shared__ int32_t smem[32];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int32_t total;
  int32_t ex = block_exclusive_scan(s, &total, smem) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

__global__ void fill_kernel(const int64_t* __restrict__ dst, int64_t n_edges, int64_t n_nodes,
                            const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                            int32_t* __restrict__ slots) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int64_t d = dst[e];
  if (d < 0 || d >= n_nodes) return;
  int32_t p = rowptr[d] + atomicAdd(&cursor[d], 1);
  slots[p] = (int32_t)e;  // arbitrary order inside the segment; fixed by order_kernel
}

// One warp per destination: rank the segment's edge ids ascending (= stable order), then emit the
// permuted views.  O(deg^2/32) per segment; in-degrees are ~max_nbr.
__global__ void order_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ attr,
                             const int32_t* __restrict__ rowptr, const int32_t* __restrict__ slots,
                             int64_t n_nodes, int32_t* __restrict__ perm, int32_t* __restrict__ src_sorted,
                             int32_t* __restrict__ dst_sorted, int32_t* __restrict__ rank_sorted, int64_t n_ranks,
                             unsigned int* __restrict__ status) {
  int64_t d = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (d >= n_nodes) return;
  int32_t b = rowptr[d], e = rowptr[d + 1];
  for (int32_t p = b + lane; p < e; p += 32) {
    int32_t id = slots[p];
    int32_t r = 0;
    for (int32_t q = b; q < e; ++q) r += (slots[q] < id);
    int32_t o = b + r;
    perm[o] = id;
    // what nn.Embedding / index_select would refuse in the reference becomes a sticky status flag here (no sync on
    // the hot path) and the index is clamped, so no kernel downstream gathers out of bounds
    int64_t sv = src[id], rv = attr[id];
    if (sv < 0 || sv >= n_nodes) atomicOr(status, (unsigned int)kStatusBadSource), sv = 0;
    if (rv < 0 || rv >= n_ranks) atomicOr(status, (unsigned int)kStatusBadRank), rv = 0;
    src_sorted[o] = (int32_t)sv;
    dst_sorted[o] = (int32_t)d;
    rank_sorted[o] = (int32_t)rv;
  }
}

__global__ void segment_ptr_kernel(const int64_t* __restrict__ index, int64_t n, int64_t n_seg,
                                   int32_t* __restrict__ ptr, int32_t* __restrict__ index32) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  // ptr[s] = first i with index[i] >= s  (index sorted, values in [0,n_seg))
  int64_t prev = (i == 0) ? -1 : index[i - 1];
  int64_t cur = (i == n) ? n_seg : index[i];
  if (i < n && index32) index32[i] = (int32_t)cur;
  if (cur > n_seg) cur = n_seg;
  for (int64_t s = prev + 1; s <= cur; ++s) ptr[s] = (int32_t)i;
}

}  // namespace
}  // namespace cgat

using namespace cgat;

extern "C" int cgat_abi_version(void) { return CGAT_B200_ABI_VERSION; }
extern "C" const char* cgat_last_error(void) { return g_err; }
extern "C" int64_t cgat_launch_count(void) { return (int64_t)g_launches; }

extern "C" size_t cgat_csr_workspace_bytes(int64_t n_edges, int64_t n_nodes) {
  int64_t n_tiles = ceil_div(n_nodes + 1, kScanTile);
  // [bad counter (4 ints)] [counts N+1] [cursor N+1] [slots E] [tile sums]
  return sizeof(int32_t) * (size_t)(4 + 2 * (n_nodes + 1) + n_edges + n_tiles + 8);
}

extern "C" int cgat_csr_build(const int64_t* edge_index, const int64_t* edge_attr, int64_t n_edges,
                              int64_t n_nodes, int32_t* perm, int32_t* rowptr, int32_t* src_sorted,
                              int32_t* dst_sorted, int32_t* rank_sorted, int32_t n_ranks, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_ranks <= 0 || n_ranks > 32) n_ranks = 32;   // the edge kernels index per-rank tables of at most 32 rows
  if (n_edges < 0 || n_nodes < 0 || n_edges >= (1ll << 31) || n_nodes >= (1ll << 31) - 1)
    return fail(-2, "cgat_csr_build: sizes out of int32 range");
  if (workspace_bytes < cgat_csr_workspace_bytes(n_edges, n_nodes))
    return fail(-3, "cgat_csr_build: workspace too small");
  int32_t* ws = (int32_t*)workspace;
  int32_t* bad = ws;
  int32_t* counts = ws + 4;
  int32_t* cursor = counts + (n_nodes + 1);
  int32_t* slots = cursor + (n_nodes + 1);
  int32_t* tile_sums = slots + n_edges;
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + n_edges;
  int64_t n_tiles = ceil_div(n_nodes + 1, kScanTile);

  CGAT_CUDA(cudaMemsetAsync(ws, 0, sizeof(int32_t) * (size_t)(4 + 2 * (n_nodes + 1)), stream));
  if (n_edges > 0) {
    histogram_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(dst, n_edges, n_nodes, counts, bad,
                                                                           status_word());
    if (int e = check_launch("histogram_kernel")) return e;
  }
  // exclusive scan of counts[0..N] (counts[N] = 0) -> rowptr[0..N]; rowptr[N] = E
  scan_tile_sums<<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(counts, n_nodes + 1, tile_sums);
  if (int e = check_launch("scan_tile_sums")) return e;
  scan_tile_offsets<<<1, kScanThreads, 0, stream>>>(tile_sums, n_tiles);
  if (int e = check_launch("scan_tile_offsets")) return e;
  scan_apply<<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(counts, n_nodes + 1, tile_sums, rowptr);
  if (int e = check_launch("scan_apply")) return e;
  if (n_edges > 0) {
    fill_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(dst, n_edges, n_nodes, rowptr, cursor, slots);
    if (int e = check_launch("fill_kernel")) return e;
    order_kernel<<<(unsigned)ceil_div(n_nodes * 32, 256), 256, 0, stream>>>(src, edge_attr, rowptr, slots, n_nodes,
                                                                           perm, src_sorted, dst_sorted, rank_sorted,
                                                                           n_ranks, status_word());
    if (int e = check_launch("order_kernel")) return e;
  }
  return 0;
}

// Sticky validation flags raised by kernels since the last reset (bit 0: destination, bit 1: source, bit 2: shell rank
// out of range in cgat_csr_build; bit 3: non-finite aggregate in cgat_edge_attn_fwd*).  SYNCHRONISES the device: call
// it where a sync is acceptable (after a step, in tests), not on the hot path.
extern "C" int cgat_status_flags(uint32_t* flags_out, int32_t reset) {
  unsigned int v = 0;
  CGAT_CUDA(cudaMemcpyFromSymbol(&v, g_status_word, sizeof(v)));
  if (flags_out) *flags_out = v;
  if (reset && v) {
    const unsigned int zero = 0;
    CGAT_CUDA(cudaMemcpyToSymbol(g_status_word, &zero, sizeof(zero)));
  }
  return 0;
}

extern "C" int cgat_segment_ptr(const int64_t* index, int64_t n, int64_t n_seg, int32_t* ptr,
                                int32_t* index32, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n_seg < 0 || n >= (1ll << 31) || n_seg >= (1ll << 31) - 1)
    return fail(-2, "cgat_segment_ptr: sizes out of int32 range");
  segment_ptr_kernel<<<(unsigned)ceil_div(n + 1, 256), 256, 0, stream>>>(index, n, n_seg, ptr, index32);
  return check_launch("segment_ptr_kernel");
}
