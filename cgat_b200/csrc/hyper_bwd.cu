// Hypernetwork linear layer, weight gradient (SURVEY.md §8a row A12 for row A5).
//
//   dL/dW[o*F+i, k] = sum_n g[n,o] * y[n,i] * z[n,k]           (g = dL/dy_out)
//
// i.e. for every output channel o a (F x F) product  (diag(g[:,o]) Y)^T Z  contracted over ATOMS.
// The reference reaches the same numbers through autograd of Linear(F -> F*F+F): it materialises the
// (N, F*F) gradient of the predicted weights (66 KB per atom) and runs one huge GEMM.  Here the scaled
// rows g[n,o]*y[n,:] are formed in registers while they are staged, so nothing of that size exists.
// Both operands are contracted over their ROWS, hence staged MN-major (SWIZZLE_128B_BASE32B): no
// transposed copies of y, z or g are needed.
//
// One CTA = (pair of output channels, atom range).  416 threads: 4 epilogue warps, 8 producer warps,
// 1 MMA warp.  TMEM: {main, correction} accumulators for the two channels (512 columns).
#include "common.cuh"
#include "tc_common.cuh"

namespace cgat {
namespace {
using namespace tc;

constexpr int kHF = 128;
constexpr int kHThreads = 128 + 256 + 32;
constexpr int kHImage = 32 * 128;       // 32 atoms (K rows) x 128 B
constexpr int kHPart = 4 * kHImage;     // 128 columns = 4 images = 16 KB
constexpr int kHStageBytes = 6 * kHPart;  // Z hi/lo, A0 hi/lo, A1 hi/lo = 96 KB
constexpr int kHStages = 2;
constexpr int kHSmemBytes = kHStages * kHStageBytes + 256 + 1024;

__global__ void __launch_bounds__(kHThreads, 1)
hyper_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ z,
                   float* __restrict__ out, int n_atoms, int n_split) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHStages * kHStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kHStages;
  uint64_t* accum = bars + 2 * kHStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.x / n_split, split = blockIdx.x % n_split;
  const int o0 = 2 * pair;
  const int n_lo = (int)((int64_t)n_atoms * split / n_split), n_hi = (int)((int64_t)n_atoms * (split + 1) / n_split);
  const int n_chunks = (n_hi - n_lo + 31) / 32;

  if (tid == 0) {
    for (int s = 0; s < kHStages; ++s) {
      mbar_init(&full[s], 256);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    mbar_init_fence();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    const int i = warp * 32 + lane;  // row of the (F x F) block = input channel i
    mbar_wait(accum, 0);
    tc_fence_after();
#pragma unroll 1
    for (int oo = 0; oo < 2; ++oo) {
      float* dst = out + ((int64_t)split * kHF * kHF + (int64_t)(o0 + oo) * kHF + i) * kHF;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float v[32], w[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + oo * 256 + cc * 32, v);
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + oo * 256 + 128 + cc * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 r;
          r.x = n_chunks ? v[4 * j] + w[4 * j] : 0.f;
          r.y = n_chunks ? v[4 * j + 1] + w[4 * j + 1] : 0.f;
          r.z = n_chunks ? v[4 * j + 2] + w[4 * j + 2] : 0.f;
          r.w = n_chunks ? v[4 * j + 3] + w[4 * j + 3] : 0.f;
          reinterpret_cast<float4*>(dst + cc * 32)[j] = r;
        }
      }
    }
    tc_fence_before();
  } else if (warp < 12) {
    const int pt = tid - 128;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int s = ch % kHStages, u = ch / kHStages;
      const int n0 = n_lo + ch * 32;
      float4 zv[4], yv[4];
      float g0[4], g1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = pt + 256 * j, r = idx >> 5, q = idx & 31;
        const int n = n0 + r;
        zv[j] = yv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        g0[j] = g1[j] = 0.f;
        if (n < n_hi) {
          zv[j] = __ldg(reinterpret_cast<const float4*>(z + (int64_t)n * kHF) + q);
          yv[j] = __ldg(reinterpret_cast<const float4*>(y + (int64_t)n * kHF) + q);
          const float2 gg = __ldg(reinterpret_cast<const float2*>(g + (int64_t)n * kHF + o0));
          g0[j] = gg.x, g1[j] = gg.y;
        }
      }
      mbar_wait(&empty[s], (u + 1) & 1u);
      uint8_t* st = smem + s * kHStageBytes;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = pt + 256 * j, r = idx >> 5, q = idx & 31;
        const uint32_t off = (q >> 3) * kHImage + mn_sw128_offset(r, q & 7);
        float4 hi, lo;
        split_tf32(zv[j], hi, lo);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + kHPart + off) = lo;
        float4 a = make_float4(yv[j].x * g0[j], yv[j].y * g0[j], yv[j].z * g0[j], yv[j].w * g0[j]);
        split_tf32(a, hi, lo);
        *reinterpret_cast<float4*>(st + 2 * kHPart + off) = hi;
        *reinterpret_cast<float4*>(st + 3 * kHPart + off) = lo;
        a = make_float4(yv[j].x * g1[j], yv[j].y * g1[j], yv[j].z * g1[j], yv[j].w * g1[j]);
        split_tf32(a, hi, lo);
        *reinterpret_cast<float4*>(st + 4 * kHPart + off) = hi;
        *reinterpret_cast<float4*>(st + 5 * kHPart + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(&full[s]);
    }
  } else {
    constexpr uint32_t idesc = umma_idesc_tf32(128, 128, 1, 1);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int s = ch % kHStages, u = ch / kHStages;
      mbar_wait(&full[s], u & 1u);
      tc_fence_after();
      {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
        const uint32_t z_hi = smem_u32(smem + s * kHStageBytes), z_lo = z_hi + kHPart;
#pragma unroll
        for (int oo = 0; oo < 2; ++oo) {
          const uint32_t a_hi = z_hi + (2 + 2 * oo) * kHPart, a_lo = a_hi + kHPart;
          const uint32_t d = tmem + oo * 256;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t o = ks * 1024;
            umma_tf32_e(d + 128, umma_desc_mn_sw128(a_lo + o, kHImage), umma_desc_mn_sw128(z_hi + o, kHImage), idesc,
                      (ch | ks) != 0);
            umma_tf32_e(d + 128, umma_desc_mn_sw128(a_hi + o, kHImage), umma_desc_mn_sw128(z_lo + o, kHImage), idesc, 1);
            umma_tf32_e(d, umma_desc_mn_sw128(a_hi + o, kHImage), umma_desc_mn_sw128(z_hi + o, kHImage), idesc,
                      (ch | ks) != 0);
          }
        }
        umma_commit_e(&empty[s]);
        if (ch == n_chunks - 1) umma_commit_e(accum);
      }
      __syncwarp();
    }
    if (n_chunks == 0 && lane == 0) mbar_arrive(accum);
  }
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

extern "C" int32_t cgat_hyper_wgrad_splits(int64_t n_atoms) {
  // grid = 64 channel pairs x splits on 148 SMs: 2 splits = 128 CTAs (one wave, 86 % of the SMs); 3 would be
  // 192 CTAs = two waves with the second one a third full
  return n_atoms > 1024 ? 2 : 1;
}

// out: (cgat_hyper_wgrad_splits(N), F*F, F) partial dL/dW[:F*F]; sum over dim 0.
extern "C" int cgat_hyper_wgrad(const float* g, const float* y, const float* z, float* out, int64_t n_atoms,
                                int32_t f, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (f != kHF) return fail(-2, "cgat_hyper_wgrad: only F = 128 is instantiated");
  if (n_atoms >= (1ll << 31) - 64) return fail(-2, "cgat_hyper_wgrad: too many atoms");
  if (n_atoms <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(hyper_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHSmemBytes));
    configured = true;
  }
  const int n_split = cgat_hyper_wgrad_splits(n_atoms);
  hyper_wgrad_kernel<<<(f / 2) * n_split, kHThreads, kHSmemBytes, stream>>>(g, y, z, out, (int)n_atoms, n_split);
  return check_launch("hyper_wgrad_kernel");
}
