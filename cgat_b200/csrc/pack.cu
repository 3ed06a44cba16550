// Weight packing for the tensor-core kernels: fp32 row-major [rows x k] -> tiles of 128 rows x 32
// floats, pre-split into (tf32 hi, tf32 lo) and pre-swizzled (canonical K-major SWIZZLE_128B), so the
// consuming kernel fetches a ready UMMA operand stage with ONE bulk async copy and spends no
// instructions on conversion.  Runs once per optimizer step (training) or once per model (inference).
#include "common.cuh"
#include "tc_common.cuh"

namespace cgat {
namespace {
using namespace tc;

// grid: (k_chunks, row_tiles); 256 threads: thread -> (row r = idx/8, 16-byte chunk c = idx%8)
__global__ void pack_kmajor_kernel(const float* __restrict__ w, int64_t ld, int rows, int k, int transpose,
                                   float* __restrict__ out) {
  const int kc = blockIdx.x, rt = blockIdx.y;
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + ((int64_t)rt * gridDim.x + kc) * kPackStageBytes;
  for (int idx = threadIdx.x; idx < kPackRows * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    const int gr = rt * kPackRows + r, gk = kc * kPackChunk + c * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = 0.f;
      if (gr < rows && gk + j < k) {
        if (transpose == 0) x = w[(int64_t)gr * ld + gk + j];
        else if (transpose == 1) x = w[(int64_t)(gk + j) * ld + gr];
        else {  // 2: transpose inside each k x k block: packed row (o, c) holds column c of block o
          const int64_t o = ((int64_t)rt * kPackRows) / k, c0 = ((int64_t)rt * kPackRows) % k;
          x = w[(o * k + gk + j) * ld + c0 + r];
        }
      }
      v[j] = x;
    }
    float4 hi, lo;
    split_tf32(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
    const uint32_t off = sw128_offset(r, c);
    *reinterpret_cast<float4*>(dst + off) = hi;
    *reinterpret_cast<float4*>(dst + kPackImageBytes + off) = lo;
  }
}


// fp16 form (f16x3, tc_common.cuh): tiles of 128 rows x 64 halves; thread -> (row r, 16-byte chunk c = 8 halves)
__global__ void pack_kmajor_f16_kernel(const float* __restrict__ w, int64_t ld, int rows, int k, int transpose,
                                       float pre_scale, float lo_scale, float* __restrict__ out) {
  const int kc = blockIdx.x, rt = blockIdx.y;
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + ((int64_t)rt * gridDim.x + kc) * kPackStageBytes;
  for (int idx = threadIdx.x; idx < kPackRows * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    const int gr = rt * kPackRows + r, gk = kc * kPackChunk16 + c * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = 0.f;
      if (gr < rows && gk + j < k) {
        if (transpose == 0) x = w[(int64_t)gr * ld + gk + j];
        else if (transpose == 1) x = w[(int64_t)(gk + j) * ld + gr];
        else {  // 2: transpose inside each k x k block: packed row (o, c) holds column c of block o
          const int64_t o = ((int64_t)rt * kPackRows) / k, c0 = ((int64_t)rt * kPackRows) % k;
          x = w[(o * k + gk + j) * ld + c0 + r];
        }
      }
      v[j] = x * pre_scale;
    }
    uint4 hi, lo;
    split_f16x8s(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), lo_scale, hi, lo);
    const uint32_t off = sw128_offset(r, c);
    *reinterpret_cast<uint4*>(dst + off) = hi;
    *reinterpret_cast<uint4*>(dst + kPackImageBytes + off) = lo;
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

extern "C" int64_t cgat_packed_floats(int64_t rows, int64_t k) { return tc::packed_floats(rows, k); }

// w: [rows x k] (transpose=0, leading dimension ld), its transpose stored as [k x rows] (transpose=1), or
// a stack of k x k blocks each of which is transposed (transpose=2; needs k % 128 == 0, rows % k == 0).
extern "C" int cgat_pack_kmajor(const float* w, int64_t ld, int64_t rows, int64_t k, int32_t transpose, float* out,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows <= 0 || k <= 0) return 0;
  if (transpose == 2 && (k % tc::kPackRows || rows % k)) return fail(-2, "cgat_pack_kmajor: block transpose needs k x k blocks, k a multiple of 128");
  dim3 grid((unsigned)ceil_div(k, tc::kPackChunk), (unsigned)ceil_div(rows, tc::kPackRows));
  pack_kmajor_kernel<<<grid, 256, 0, stream>>>(w, ld, (int)rows, (int)k, transpose, out);
  return check_launch("pack_kmajor_kernel");
}

extern "C" int64_t cgat_packed_floats_f16(int64_t rows, int64_t k) { return tc::packed_floats_f16(rows, k); }

// fp16 hi/lo form of cgat_pack_kmajor (operands of the kind::f16 kernels): hi = f16(w * pre_scale),
// lo = f16((w * pre_scale - hi) * lo_scale); `out` holds cgat_packed_floats_f16(rows, k) floats.
extern "C" int cgat_pack_kmajor_f16s(const float* w, int64_t ld, int64_t rows, int64_t k, int32_t transpose,
                                     float pre_scale, float lo_scale, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows <= 0 || k <= 0) return 0;
  if (transpose == 2 && (k % tc::kPackRows || rows % k)) return fail(-2, "cgat_pack_kmajor_f16: block transpose needs k x k blocks, k a multiple of 128");
  dim3 grid((unsigned)ceil_div(k, tc::kPackChunk16), (unsigned)ceil_div(rows, tc::kPackRows));
  pack_kmajor_f16_kernel<<<grid, 256, 0, stream>>>(w, ld, (int)rows, (int)k, transpose, pre_scale, lo_scale, out);
  return check_launch("pack_kmajor_f16_kernel");
}

// the form cgat_hyper_*_f16 read: unscaled hi, lo scaled by 2^11 (separate correction accumulator)
extern "C" int cgat_pack_kmajor_f16(const float* w, int64_t ld, int64_t rows, int64_t k, int32_t transpose, float* out,
                                    void* stream_) {
  return cgat_pack_kmajor_f16s(w, ld, rows, k, transpose, 1.f, tc::kF16LoScale, out, stream_);
}
