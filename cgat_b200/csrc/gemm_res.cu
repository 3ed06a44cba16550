// cgat_gemm3x_nt_res: C[M,N] = act(A[M,K] * W[N,K]^T + bias) for a SHORT contraction (K <= 128) against a WIDE,
// pre-packed weight (N up to tens of 128-row tiles): the per-atom first-layer projections
//     P = x [W1A_i; W1M_i; W1A_j; W1M_j]^T        (N_atoms x 128) x (128 x 4*H*Hd)
// that replace the per-edge contraction of the reference's MultiHeadNetwork.fc_in (reference CGAT/CGAT.py:103-109,
// :320-322; see cgat_b200/CGAT.py for the split).  cgat_gemm3x_nt launches one CTA per output tile; with K = 128
// a tile is 48 MMAs, so set-up (TMEM allocation, barrier init, unpipelined staging and epilogue) dominates.
// Here one persistent CTA per SM owns a contiguous run of (row tile, column tile) items, stages the A tile once
// per row tile (fp32 -> tf32 hi/lo, SWIZZLE_128B K-major), streams the packed weight tiles with cp.async.bulk
// through a 3-stage ring and double-buffers the accumulators in TMEM, so the epilogue of tile i (TMEM -> bias /
// activation -> global) runs under the MMAs of tile i+1.
//
// Roles (288 threads): warps 0-3 epilogue, warps 4-7 A stagers (+ the weight TMA thread), warp 8 MMA issue.
#include "common.cuh"
#include "tc_common.cuh"

namespace cgat {
namespace {
using namespace tc;

constexpr int kRMaxKC = 4;                                        // K <= 128
constexpr int kRABytes = kRMaxKC * (int)kPackStageBytes;          // 128 KB: A tile, hi+lo
constexpr int kRStages = 3;
constexpr int kRSmemBytes = kRABytes + kRStages * (int)kPackStageBytes + 1024 + 512;
constexpr int kRThreads = 288;

__device__ __forceinline__ float res_act(float v, int act) {
  switch (act) {
    case 1: return v > 0.f ? v : 0.01f * v;
    case 2: return tanhf(v);
    case 3: return fmaxf(v, 0.f);
    default: return v;
  }
}

__global__ void __launch_bounds__(kRThreads, 1)
gemm3x_nt_res_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ w_packed,
                     const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N, int K, int act) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + kRABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + kRStages * kPackStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kRStages;
  uint64_t* tmem_full = bars + 2 * kRStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint64_t* a_full = tmem_empty + 2;
  uint64_t* a_free = a_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kcn = (K + kPackChunk - 1) / kPackChunk;
  const int m_tiles = (M + 127) / 128, n_tiles = (N + 127) / 128;
  const int n_items = m_tiles * n_tiles;  // item = (row tile, column tile), column tile fastest
  const int item_lo = (int)((int64_t)n_items * blockIdx.x / gridDim.x);
  const int item_hi = (int)((int64_t)n_items * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int s = 0; s < kRStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    mbar_init(a_full, 128);
    mbar_init(a_free, 1);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ epilogue (thread = output row)
    uint32_t tc = 0;
    const bool vec_ok = ((ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(C) & 31) == 0);
    for (int item = item_lo; item < item_hi; ++item, ++tc) {
      const int mt = item / n_tiles, nt = item - mt * n_tiles;
      const int m = mt * 128 + warp * 32 + lane;
      const uint32_t b = tc & 1u;
      mbar_wait(&tmem_full[b], (tc >> 1) & 1u);
      tc_fence_after();
      const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16) + b * 256;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float v[32], w[32];
        tmem_ld32(tb + cc * 32, v);
        tmem_ld32(tb + 128 + cc * 32, w);
        tmem_ld_wait();
        const int nb = nt * 128 + cc * 32;
        if (m < M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
          if (bias != nullptr || act != 0) {  // kept out of the common path: the inlined activations are large
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float bj = (bias != nullptr && nb + j < N) ? __ldg(bias + nb + j) : 0.f;
              v[j] = res_act(v[j] + bj, act);
            }
          }
          float* crow = C + (int64_t)m * ldc + nb;
          if (vec_ok && nb + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 4; ++j) st_global_v8(crow + 8 * j, v + 8 * j);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < N) crow[j] = v[j];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[b]);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ A-tile stagers + weight TMA
    const int st = tid - 128;
    uint32_t it = 0, cnt = 0;
    int staged = -1;
    for (int item = item_lo; item < item_hi; ++item) {
      const int mt = item / n_tiles, nt = item - mt * n_tiles;
      if (mt != staged) {
        staged = mt;
        mbar_wait(a_free, (it + 1) & 1u);
#pragma unroll 1
        for (int kc = 0; kc < kcn; ++kc) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int idx = st + 128 * j, r = idx >> 3, c = idx & 7;
            const int gr = mt * 128 + r, gk = kc * 32 + c * 4;
            v[j] = (gr < M && gk < K) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)gr * lda + gk))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          uint8_t* hi = a_smem + kc * kPackStageBytes;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int idx = st + 128 * j;
            const uint32_t off = sw128_offset(idx >> 3, idx & 7);
            float4 h, l;
            split_tf32(v[j], h, l);
            *reinterpret_cast<float4*>(hi + off) = h;
            *reinterpret_cast<float4*>(hi + kPackImageBytes + off) = l;
          }
        }
        fence_async_smem();
        mbar_arrive(a_full);
        ++it;
      }
      if (st == 0) {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(w_packed) + (int64_t)nt * kcn * kPackStageBytes;
        for (int kc = 0; kc < kcn; ++kc, ++cnt) {
          const uint32_t s = cnt % kRStages, u = cnt / kRStages;
          mbar_wait(&empty[s], (u + 1) & 1u);
          mbar_arrive_expect_tx(&full[s], kPackStageBytes);
          bulk_g2s(b_smem + s * kPackStageBytes, wsrc + (int64_t)kc * kPackStageBytes, kPackStageBytes, &full[s]);
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(128, 128), idesc2 = umma_idesc_tf32(128, 256);
    uint32_t it = 0, cnt = 0, tc = 0;
    int staged = -1;
    for (int item = item_lo; item < item_hi; ++item, ++tc) {
      const int mt = item / n_tiles;
      if (mt != staged) {
        mbar_wait(a_full, it & 1u);
        ++it;
        staged = mt;
      }
      const uint32_t b = tc & 1u;
      mbar_wait(&tmem_empty[b], ((tc >> 1) + 1) & 1u);
      tc_fence_after();
      for (int kc = 0; kc < kcn; ++kc, ++cnt) {
        const uint32_t s = cnt % kRStages, u = cnt / kRStages;
        mbar_wait(&full[s], u & 1u);
        tc_fence_after();
        {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
          const uint32_t a_hi = smem_u32(a_smem + kc * kPackStageBytes), a_lo = a_hi + kPackImageBytes;
          const uint32_t b_hi = smem_u32(b_smem + s * kPackStageBytes), b_lo = b_hi + kPackImageBytes;
          const uint32_t d = tmem + b * 256, dc = d + 128;
          (void)b_lo;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // one N = 256 MMA against the adjacent [hi | lo] weight images: a_hi*b_hi -> [d, d+128), a_hi*b_lo ->
            // [d+128, d+256); then a_lo*b_hi into the correction columns (see hyper_fwd.cu)
            const uint32_t off = ks * 32;
            umma_tf32_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc2, (kc | ks) != 0);
            umma_tf32_e(dc, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, 1);
          }
          umma_commit_e(&empty[s]);
          if (kc == kcn - 1) umma_commit_e(&tmem_full[b]);
        }
        __syncwarp();
      }
      if ((item + 1 == item_hi || (item + 1) / n_tiles != mt)) umma_commit_e(a_free);
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

// C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) with W given as cgat_pack_kmajor(W, ld, N, K, 0) (so N is padded to 128-row
// tiles inside the image) and K <= 128, K % 4 == 0.  Persistent, A tile resident, accumulators double-buffered.
extern "C" int cgat_gemm3x_nt_res(const float* A, int64_t lda, const float* w_packed, const float* bias, float* C,
                                  int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t act, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0 || K > 128 || (K & 3) || (lda & 3) || (reinterpret_cast<uintptr_t>(A) & 15))
    return fail(-2, "cgat_gemm3x_nt_res: 0 < K <= 128, K and lda multiples of 4, A 16-byte aligned");
  if (M >= (1ll << 31) - 128 || N >= (1ll << 31) - 128) return fail(-2, "cgat_gemm3x_nt_res: size overflow");
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(gemm3x_nt_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRSmemBytes));
    configured = true;
  }
  const int64_t n_items = ceil_div(M, 128) * ceil_div(N, 128);
  const int grid = (int)(n_items < kNumSMs ? n_items : kNumSMs);
  gemm3x_nt_res_kernel<<<grid, kRThreads, kRSmemBytes, stream>>>(A, lda, w_packed, bias, C, ldc, (int)M, (int)N, (int)K,
                                                                 act);
  return check_launch("gemm3x_nt_res_kernel");
}
