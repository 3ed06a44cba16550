// Deterministic, atomic-free segmented softmax + weighted sum.
//
// Replaces torch_geometric.utils.softmax + torch_scatter.scatter_add (reference CGAT/CGAT.py:59-61,
// :323-326) and the scatter_max / scatter_add pair of WeightedAttention (reference
// CGAT/roost_message.py:307-315).  Segments are contiguous row ranges (edges grouped by destination
// atom, atoms grouped by crystal, Roost pairs grouped by element), so each (segment, head, channel)
// is reduced by exactly one thread in row order: no atomics, bit-reproducible.
//
// HBM-bound: forward reads gate+value once (online softmax: running max, rescaled sum and
// accumulator), backward is a per-row elementwise pass given the saved per-segment statistics.
#include "common.cuh"

namespace cgat {
namespace {

// One CTA per segment; thread per (head, channel) column with a strided loop.
template <bool kVector>
__global__ void seg_softmax_fwd_kernel(const float* __restrict__ gate, const float* __restrict__ value,
                                       const float* __restrict__ u, const int32_t* __restrict__ ptr,
                                       int heads, int f, float eps, float* __restrict__ out,
                                       float* __restrict__ seg_max, float* __restrict__ seg_den) {
  const int64_t s = blockIdx.x;
  const int32_t b = ptr[s], e = ptr[s + 1];
  const int hf = heads * f;
  const int gstride = kVector ? hf : heads;
  for (int ch = threadIdx.x; ch < hf; ch += blockDim.x) {
    const int h = ch / f;
    const int gcol = kVector ? ch : h;
    float m = -INFINITY, den = 0.f, acc = 0.f;
    for (int32_t t = b; t < e; ++t) {
      float a = __ldg(gate + (int64_t)t * gstride + gcol);
      float v = __ldg(value + (int64_t)t * hf + ch);
      float w = u ? __ldg(u + t) : 1.f;
      float mn = fmaxf(m, a);
      float r = expf(m - mn);  // 0 on the first row (m = -inf)
      float p = w * expf(a - mn);
      den = den * r + p;
      acc = acc * r + p * v;
      m = mn;
    }
    const bool empty = (e <= b);
    out[s * hf + ch] = empty ? 0.f : acc / (den + eps);
    if (kVector || (ch % f) == 0) {
      seg_max[s * gstride + gcol] = empty ? 0.f : m;  // torch_scatter: empty segments read 0
      seg_den[s * gstride + gcol] = den;
    }
  }
}

// vector attention: thread per (row, head, channel)
__global__ void seg_softmax_bwd_vec_kernel(const float* __restrict__ gate, const float* __restrict__ value,
                                           const float* __restrict__ u, const int32_t* __restrict__ seg_of_row,
                                           const float* __restrict__ out, const float* __restrict__ seg_max,
                                           const float* __restrict__ seg_den, const float* __restrict__ d_out,
                                           int64_t n_rows, int hf, float eps, float* __restrict__ d_gate,
                                           float* __restrict__ d_value) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * hf) return;
  int64_t t = i / hf;
  int ch = (int)(i - t * hf);
  int64_t s = seg_of_row[t];
  float w = u ? __ldg(u + t) : 1.f;
  float alpha = w * expf(gate[i] - __ldg(seg_max + s * hf + ch)) / (__ldg(seg_den + s * hf + ch) + eps);
  float g = __ldg(d_out + s * hf + ch);
  d_value[i] = alpha * g;
  d_gate[i] = alpha * (value[i] - __ldg(out + s * hf + ch)) * g;
}

// scalar attention: one warp per (row, head); reduction over the F channels
__global__ void seg_softmax_bwd_scalar_kernel(const float* __restrict__ gate, const float* __restrict__ value,
                                              const float* __restrict__ u, const int32_t* __restrict__ seg_of_row,
                                              const float* __restrict__ out, const float* __restrict__ seg_max,
                                              const float* __restrict__ seg_den, const float* __restrict__ d_out,
                                              int64_t n_rows, int heads, int f, float eps,
                                              float* __restrict__ d_gate, float* __restrict__ d_value) {
  int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (wid >= n_rows * heads) return;
  int64_t t = wid / heads;
  int h = (int)(wid - t * heads);
  int64_t s = seg_of_row[t];
  float w = u ? __ldg(u + t) : 1.f;
  float alpha = w * expf(gate[wid] - __ldg(seg_max + s * heads + h)) / (__ldg(seg_den + s * heads + h) + eps);
  float dot = 0.f;
  const int64_t vb = wid * f, ob = (s * heads + h) * f;
  for (int c = lane; c < f; c += 32) {
    float g = __ldg(d_out + ob + c);
    d_value[vb + c] = alpha * g;
    dot += (value[vb + c] - __ldg(out + ob + c)) * g;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (lane == 0) d_gate[wid] = alpha * dot;
}

}  // namespace
}  // namespace cgat

using namespace cgat;

extern "C" int cgat_seg_softmax_fwd(const float* gate, const float* value, const float* u, const int32_t* ptr,
                                    int64_t n_seg, int32_t heads, int32_t f, int32_t fa, float eps, float* out,
                                    float* seg_max, float* seg_den, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (fa != f && fa != 1) return fail(-2, "cgat_seg_softmax_fwd: fa must be f or 1");
  if (n_seg == 0) return 0;
  int hf = heads * f;
  int threads = hf >= 256 ? 256 : ((hf + 31) / 32) * 32;
  if (fa == f)
    seg_softmax_fwd_kernel<true><<<(unsigned)n_seg, threads, 0, stream>>>(gate, value, u, ptr, heads, f, eps, out,
                                                                          seg_max, seg_den);
  else
    seg_softmax_fwd_kernel<false><<<(unsigned)n_seg, threads, 0, stream>>>(gate, value, u, ptr, heads, f, eps, out,
                                                                           seg_max, seg_den);
  return check_launch("seg_softmax_fwd_kernel");
}

extern "C" int cgat_seg_softmax_bwd(const float* gate, const float* value, const float* u,
                                    const int32_t* seg_of_row, const float* out, const float* seg_max,
                                    const float* seg_den, const float* d_out, int64_t n_rows, int32_t heads,
                                    int32_t f, int32_t fa, float eps, float* d_gate, float* d_value,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (fa != f && fa != 1) return fail(-2, "cgat_seg_softmax_bwd: fa must be f or 1");
  if (n_rows == 0) return 0;
  if (fa == f) {
    int64_t total = n_rows * heads * f;
    seg_softmax_bwd_vec_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(
        gate, value, u, seg_of_row, out, seg_max, seg_den, d_out, n_rows, heads * f, eps, d_gate, d_value);
  } else {
    int64_t warps = n_rows * heads;
    seg_softmax_bwd_scalar_kernel<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, stream>>>(
        gate, value, u, seg_of_row, out, seg_max, seg_den, d_out, n_rows, heads, f, eps, d_gate, d_value);
  }
  return check_launch("seg_softmax_bwd_kernel");
}
