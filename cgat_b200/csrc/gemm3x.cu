// cgat_gemm3x_nt: C[M,N] = act(A[M,K] · B[N,K]^T + bias), fp32 in / fp32 out, on the 5th-gen tensor
// cores (tcgen05.mma kind::tf32, accumulators in TMEM) with error compensation: every operand is
// split into tf32 hi/lo parts while it is staged into shared memory and three MMA passes
// (lo·hi + hi·lo + hi·hi) reproduce the fp32 product to ~2^-21 relative.
//
// This is the building block of the dense pieces of the CGAT hot path that are plain matrix
// products: the per-atom first-layer projections x·W1_i^T / x·W1_j^T (reference CGAT/CGAT.py:320-322
// after the split described in cgat_b200/CGAT.py), the hypernetwork trunk layers (reference
// CGAT/Hypernetworksmp.py:36-83) and their transposes in backward.
//
// Structure (one 128 x BN output tile per CTA):
//   warps 0-7  producers in two groups of 128 that take alternate K chunks: global -> registers (float4,
//              coalesced) -> hi/lo split -> shared memory in the canonical K-major SWIZZLE_128B UMMA layout.
//              Every group keeps the loads of its NEXT chunk in flight while it converts and stores the
//              current one (round 1 loaded A, waited, loaded B, waited, once per chunk with 128 threads:
//              two exposed HBM round trips per chunk, 3 us per 32-float chunk in the in-graph profile
//              profiles/r03m — 40 TFLOP/s).  Then the epilogue (tcgen05.ld -> bias/activation -> global),
//              the column blocks dealt to the two groups.
//   warp  8    allocates TMEM and issues the MMAs; tcgen05.commit releases stages
//   kStages-deep ring of (A_hi, A_lo, B_hi, B_lo) K-chunks of 32 floats, full/empty mbarriers.
#include "common.cuh"
#include "tc_common.cuh"

namespace cgat {
namespace {

using namespace tc;

constexpr int kBM = 128;       // rows of the output tile = TMEM lanes
constexpr int kKC = 32;        // K floats per chunk = one 128-byte swizzle row
constexpr int kGroup = 128;    // threads of one producer group (= the 128 TMEM lanes in the epilogue)
constexpr int kWorkers = 2 * kGroup;  // producer / epilogue threads
constexpr int kMmaWarp = kWorkers / 32;
constexpr int kThreads = kWorkers + 32;

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN <= 128) ? 3 : 2;
  static constexpr int kABytes = kBM * kKC * 4;  // 16 KB
  static constexpr int kBBytes = BN * kKC * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case 1: return v > 0.f ? v : 0.01f * v;
    case 2: return tanhf(v);
    case 3: return fmaxf(v, 0.f);
    default: return v;
  }
}

// a [rows x 32] fp32 chunk (row-major source with leading dimension ld): loads into registers ...
template <int ROWS>
__device__ __forceinline__ void load_chunk(const float* __restrict__ src, int64_t ld, int row0, int n_rows, int k0,
                                           int k_total, float4 (&v)[ROWS * 8 / kGroup], int t) {
#pragma unroll
  for (int j = 0; j < ROWS * 8 / kGroup; ++j) {
    const int idx = t + kGroup * j, r = idx >> 3, c = idx & 7;
    const int gr = row0 + r, gk = k0 + c * 4;
    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < n_rows && gk < k_total) v[j] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)gr * ld + gk));
  }
}
// ... and from there as hi/lo SW128 tiles into a stage
template <int ROWS>
__device__ __forceinline__ void store_chunk(const float4 (&v)[ROWS * 8 / kGroup], uint8_t* hi, uint8_t* lo, int t) {
#pragma unroll
  for (int j = 0; j < ROWS * 8 / kGroup; ++j) {
    const int idx = t + kGroup * j;
    const uint32_t off = sw128_offset(idx >> 3, idx & 7);
    float4 h, l;
    split_tf32(v[j], h, l);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm3x_nt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                 const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N, int K, int act,
                 int k_per_split, int64_t split_stride) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base_u32 & 1023u)) & 1023u);  // SWIZZLE_128B needs 1024-B alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* full = bars;                  // [kStages]  producers -> MMA
  uint64_t* empty = bars + S::kStages;    // [kStages]  MMA -> producers
  uint64_t* accum = bars + 2 * S::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::kStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  // split-K over blockIdx.z: partial products `split_stride` floats apart (bias / act only with one split)
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  // a split that starts at or beyond K is EMPTY (nk = 0): its partial tile is all zeros (+ bias / act) and the MMA
  // warp arrives on `accum` by hand.  (Round 1 computed nk with a truncating division that went NEGATIVE for
  // k_begin >= K + 64 and never arrived: the epilogue warps then waited forever.)
  const int nk = k_end > k_begin ? (k_end - k_begin + kKC - 1) / kKC : 0;
  C += (int64_t)blockIdx.z * split_stride;

  if (tid == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full[s], kGroup);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    mbar_init_fence();
  }
  // columns [0,BN): hi*hi sum; [BN,2BN): the 2^-11-sized correction terms, kept apart so the large
  // accumulator is truncated K/8 times instead of 3K/8 (the tensor core truncates on every MMA)
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < kMmaWarp) {
    // ---------------- producers ----------------
    const int grp = warp >> 2, t = tid & (kGroup - 1);
    float4 a0[kBM * 8 / kGroup], b0[BN * 8 / kGroup], a1[kBM * 8 / kGroup], b1[BN * 8 / kGroup];
    auto fetch = [&](float4 (&a)[kBM * 8 / kGroup], float4 (&b)[BN * 8 / kGroup], int kc) {
      load_chunk<kBM>(A, lda, m0, M, k_begin + kc * kKC, k_end, a, t);
      load_chunk<BN>(B, ldb, n0, N, k_begin + kc * kKC, k_end, b, t);
    };
    auto put = [&](const float4 (&a)[kBM * 8 / kGroup], const float4 (&b)[BN * 8 / kGroup], int kc) {
      const int s = kc % S::kStages, u = kc / S::kStages;
      mbar_wait(&empty[s], (u + 1) & 1);  // passes immediately on the first use of a stage
      uint8_t* st = smem + s * S::kStageBytes;
      store_chunk<kBM>(a, st, st + S::kABytes, t);
      store_chunk<BN>(b, st + 2 * S::kABytes, st + 2 * S::kABytes + S::kBBytes, t);
      fence_async_smem();
      mbar_arrive(&full[s]);
    };
    int kc = grp;
    if (kc < nk) fetch(a0, b0, kc);
    while (kc < nk) {
      if (kc + 2 < nk) fetch(a1, b1, kc + 2);
      put(a0, b0, kc);
      kc += 2;
      if (kc >= nk) break;
      if (kc + 2 < nk) fetch(a0, b0, kc + 2);
      put(a1, b1, kc);
      kc += 2;
    }
    // ---------------- epilogue ----------------
    mbar_wait(accum, 0);
    tc_fence_after();
    const int m = m0 + (warp & 3) * 32 + lane;
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll 1
    for (int cc = grp; cc < BN / 32; cc += 2) {
      float v[32], w[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32, v);
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + BN + cc * 32, w);
      tmem_ld_wait();
      const int nb = n0 + cc * 32;
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          int n = nb + j;
          float b = (bias != nullptr && n < N) ? __ldg(bias + n) : 0.f;
          v[j] = apply_act((nk > 0 ? v[j] + w[j] : 0.f) + b, act);
        }
        float* crow = C + (int64_t)m * ldc + nb;
        if (vec_ok && nb + 32 <= N) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(crow)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < N) crow[j] = v[j];
        }
      }
    }
    tc_fence_before();
  } else {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % S::kStages, u = kc / S::kStages;
      mbar_wait(&full[s], u & 1);
      tc_fence_after();
      {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
        const uint32_t a_hi = smem_u32(smem + s * S::kStageBytes);
        const uint32_t a_lo = a_hi + S::kABytes;
        const uint32_t b_hi = a_hi + 2 * S::kABytes;
        const uint32_t b_lo = b_hi + S::kBBytes;
#pragma unroll
        for (int ks = 0; ks < kKC / 8; ++ks) {  // UMMA K = 8 tf32 = 32 bytes along the swizzled row
          const uint32_t o = ks * 32;
          umma_tf32_e(tmem + BN, umma_desc_k_sw128(a_lo + o), umma_desc_k_sw128(b_hi + o), idesc, (kc | ks) != 0);
          umma_tf32_e(tmem + BN, umma_desc_k_sw128(a_hi + o), umma_desc_k_sw128(b_lo + o), idesc, 1);
          umma_tf32_e(tmem, umma_desc_k_sw128(a_hi + o), umma_desc_k_sw128(b_hi + o), idesc, (kc | ks) != 0);
        }
        umma_commit_e(&empty[s]);  // stage reusable once these MMAs have read it
        if (kc == nk - 1) umma_commit_e(accum);
      }
      __syncwarp();
    }
    if (nk == 0 && lane == 0) mbar_arrive(accum);  // empty split: nothing to wait for
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 2 * BN);
  }
}

// ---------------------------------------------------------------------------------------------
// C[M,N] = sum_k A[k,m] * B[k,n]  (A: [K x M], B: [K x N], both row-major): the weight-gradient shape
// dW = dY^T X, where the contraction runs over the ROWS of both operands.  Both operands are staged
// MN-major (no transposes anywhere): a K-chunk of 32 rows becomes four 32-column images per operand.
// One 128 x 128 tile per CTA; split-K over blockIdx.z writes partial tiles that the caller sums.
constexpr int kTnImage = 32 * 128;  // 32 K-rows x 128 B

// a [32 K-rows x 128 columns] fp32 chunk of an MN-major operand: loads into registers ...
__device__ __forceinline__ void load_chunk_mn(const float* __restrict__ src, int64_t ld, int col0, int n_cols, int k0,
                                              int k_end, float4 (&v)[8], int t) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int idx = t + kGroup * j, r = idx >> 5, q = idx & 31;
    const int gk = k0 + r, gc = col0 + q * 4;
    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gk < k_end && gc < n_cols) v[j] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)gk * ld + gc));
  }
}
// ... and from there as hi/lo MN-major SW128 images into a stage (kSum: column sums on the way, q is fixed per thread)
template <bool kSum>
__device__ __forceinline__ void store_chunk_mn(const float4 (&v)[8], uint8_t* hi, uint8_t* lo, int t, float4& csum) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int idx = t + kGroup * j, r = idx >> 5, q = idx & 31;
    const uint32_t off = (q >> 3) * kTnImage + mn_sw128_offset(r, q & 7);
    if (kSum) csum.x += v[j].x, csum.y += v[j].y, csum.z += v[j].z, csum.w += v[j].w;
    float4 h, l;
    split_tf32(v[j], h, l);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

constexpr int kTnMaxBatch = 24;
struct TnArgs {
  const float* a[kTnMaxBatch];   // [K x M] row-major, one per batch entry
  const float* b[kTnMaxBatch];   // [K x N]
  float* c;                      // partial (split s, batch e) at c + s * split_stride + e * batch_stride
  float* colsum;                 // optional: column sums of A, (n_split, batch, M)
  int64_t lda, ldb, ldc, split_stride, batch_stride;
  int M, N, K, k_per_split, n_split, batch;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm3x_tn_kernel(const TnArgs g) {
  const int entry = blockIdx.z / g.n_split, split = blockIdx.z - entry * g.n_split;
  const float* __restrict__ A = g.a[entry];
  const float* __restrict__ B = g.b[entry];
  const int64_t lda = g.lda, ldb = g.ldb, ldc = g.ldc;
  const int M = g.M, N = g.N, K = g.K, k_per_split = g.k_per_split;
  constexpr int kStages = 3, kOp = 4 * kTnImage;  // 16 KB per operand part
  constexpr int kStageBytes = 4 * kOp;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* accum = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
  const int k_begin = split * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int nk = k_end > k_begin ? (k_end - k_begin + kKC - 1) / kKC : 0;  // empty split: zeros, see gemm3x_nt_kernel
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], kGroup);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    mbar_init_fence();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < kMmaWarp) {
    const bool do_sum = g.colsum != nullptr && blockIdx.x == 0;
    const int grp = warp >> 2, t = tid & (kGroup - 1);
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    // two groups on alternate chunks, the next chunk's loads in flight while the current one is converted (see the
    // file comment)
    float4 a0[8], b0[8], a1[8], b1[8];
    auto fetch = [&](float4 (&a)[8], float4 (&b)[8], int kc) {
      load_chunk_mn(A, lda, m0, M, k_begin + kc * kKC, k_end, a, t);
      load_chunk_mn(B, ldb, n0, N, k_begin + kc * kKC, k_end, b, t);
    };
    auto put = [&](const float4 (&a)[8], const float4 (&b)[8], int kc) {
      const int s = kc % kStages, u = kc / kStages;
      mbar_wait(&empty[s], (u + 1) & 1);
      uint8_t* st = smem + s * kStageBytes;
      if (do_sum) store_chunk_mn<true>(a, st, st + kOp, t, csum);
      else store_chunk_mn<false>(a, st, st + kOp, t, csum);
      store_chunk_mn<false>(b, st + 2 * kOp, st + 3 * kOp, t, csum);
      fence_async_smem();
      mbar_arrive(&full[s]);
    };
    int kc = grp;
    if (kc < nk) fetch(a0, b0, kc);
    while (kc < nk) {
      if (kc + 2 < nk) fetch(a1, b1, kc + 2);
      put(a0, b0, kc);
      kc += 2;
      if (kc >= nk) break;
      if (kc + 2 < nk) fetch(a0, b0, kc + 2);
      put(a1, b1, kc);
      kc += 2;
    }
    mbar_wait(accum, 0);
    tc_fence_after();
    if (do_sum) {
      // column sums of A over this split's rows (bias gradients): lane = column group, the 8 warps hold
      // disjoint row sets; the operand stages are free again once `accum` has fired
      float4* red = reinterpret_cast<float4*>(smem);
      red[tid] = csum;
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      const float* rf = reinterpret_cast<const float*>(smem);
      const int q = (tid & 127) >> 2, comp = tid & 3;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += rf[(w * 32 + q) * 4 + comp];
      if (tid < 128 && m0 + tid < M) g.colsum[((int64_t)split * g.batch + entry) * M + m0 + tid] = sum;
    }
    const int m = m0 + (warp & 3) * 32 + lane;
    float* cbase = g.c + (int64_t)split * g.split_stride + (int64_t)entry * g.batch_stride;
#pragma unroll 1
    for (int cc = grp; cc < 4; cc += 2) {
      float v[32], w[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32, v);
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128 + cc * 32, w);
      tmem_ld_wait();
      if (m < M) {
        float* crow = cbase + (int64_t)m * ldc + n0 + cc * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + cc * 32 + j < N) crow[j] = nk > 0 ? v[j] + w[j] : 0.f;
      }
    }
    tc_fence_before();
  } else {
    constexpr uint32_t idesc = umma_idesc_tf32(128, 128, 1, 1);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kStages, u = kc / kStages;
      mbar_wait(&full[s], u & 1);
      tc_fence_after();
      {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
        const uint32_t a_hi = smem_u32(smem + s * kStageBytes), a_lo = a_hi + kOp;
        const uint32_t b_hi = a_hi + 2 * kOp, b_lo = a_hi + 3 * kOp;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // one MMA consumes 8 K-rows = two 512-byte atoms of every image
          const uint32_t o = ks * 1024;
          umma_tf32_e(tmem + 128, umma_desc_mn_sw128(a_lo + o, kTnImage), umma_desc_mn_sw128(b_hi + o, kTnImage), idesc,
                    (kc | ks) != 0);
          umma_tf32_e(tmem + 128, umma_desc_mn_sw128(a_hi + o, kTnImage), umma_desc_mn_sw128(b_lo + o, kTnImage), idesc, 1);
          umma_tf32_e(tmem, umma_desc_mn_sw128(a_hi + o, kTnImage), umma_desc_mn_sw128(b_hi + o, kTnImage), idesc,
                    (kc | ks) != 0);
        }
        umma_commit_e(&empty[s]);
        if (kc == nk - 1) umma_commit_e(accum);
      }
      __syncwarp();
    }
    if (nk == 0 && lane == 0) mbar_arrive(accum);
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

template <int BN>
int launch_gemm(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                int M, int N, int K, int act, cudaStream_t stream, int n_split = 1, int64_t split_stride = 0) {
  using S = GemmSmem<BN>;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(gemm3x_nt_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes));
    configured = true;
  }
  int k_per_split = (int)ceil_div(ceil_div(K, n_split), kKC) * kKC;
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, kBM), (unsigned)n_split);
  gemm3x_nt_kernel<BN><<<grid, kThreads, S::kBytes, stream>>>(A, lda, B, ldb, bias, C, ldc, M, N, K, act, k_per_split,
                                                              split_stride);
  return check_launch("gemm3x_nt_kernel");
}

}  // namespace
}  // namespace cgat

using namespace cgat;

extern "C" int cgat_gemm3x_nt(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C,
                              int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t act, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0 || (K & 3) || (lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(B) & 15))
    return fail(-2, "cgat_gemm3x_nt: K, lda, ldb must be multiples of 4 and A, B 16-byte aligned");
  if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return fail(-2, "cgat_gemm3x_nt: size overflow");
  if (N <= 64) return launch_gemm<64>(A, lda, B, ldb, bias, C, ldc, (int)M, (int)N, (int)K, act, stream);
  return launch_gemm<128>(A, lda, B, ldb, bias, C, ldc, (int)M, (int)N, (int)K, act, stream);
}

namespace {
int launch_tn(const TnArgs& a, cudaStream_t stream) {
  constexpr int kSmem = 3 * 4 * 4 * kTnImage + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(gemm3x_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(a.N, 128), (unsigned)ceil_div(a.M, 128), (unsigned)(a.n_split * a.batch));
  gemm3x_tn_kernel<<<grid, kThreads, kSmem, stream>>>(a);
  return check_launch("gemm3x_tn_kernel");
}

int check_tn(const float* A, const float* B, int64_t lda, int64_t ldb, int64_t M, int64_t N, int64_t K) {
  if ((M & 3) || (N & 3) || (lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(B) & 15))
    return fail(-2, "cgat_gemm3x_tn: M, N, lda, ldb must be multiples of 4 and A, B 16-byte aligned");
  if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return fail(-2, "cgat_gemm3x_tn: size overflow");
  return 0;
}
}  // namespace

// Split-K form of cgat_gemm3x_nt for long contractions with few output tiles (dL/dx = dP W1: K = 4*H*Hd):
// C[s][M,N] = A[:, split s of K] * B[:, split s of K]^T, partial results `split_stride` floats apart.
extern "C" int cgat_gemm3x_nt_splitk(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                                     int64_t split_stride, int64_t M, int64_t N, int64_t K, int32_t n_split,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M <= 0 || N <= 0 || n_split <= 0) return 0;
  if (K <= 0 || (K & 3) || (lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(B) & 15))
    return fail(-2, "cgat_gemm3x_nt_splitk: K, lda, ldb must be multiples of 4 and A, B 16-byte aligned");
  if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return fail(-2, "cgat_gemm3x_nt_splitk: size overflow");
  if (N <= 64)
    return launch_gemm<64>(A, lda, B, ldb, nullptr, C, ldc, (int)M, (int)N, (int)K, 0, stream, n_split, split_stride);
  return launch_gemm<128>(A, lda, B, ldb, nullptr, C, ldc, (int)M, (int)N, (int)K, 0, stream, n_split, split_stride);
}

// C[split][M,N] = sum over this split's rows k of A[k,m] * B[k,n].  n_split partial results, `split_stride`
// floats apart (the caller sums them: keeps the K/8-step accumulation short and fills the SMs).
extern "C" int cgat_gemm3x_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                              int64_t split_stride, int64_t M, int64_t N, int64_t K, int32_t n_split, void* stream_) {
  if (M <= 0 || N <= 0 || n_split <= 0) return 0;
  if (int e = check_tn(A, B, lda, ldb, M, N, K)) return e;
  int k_per_split = (int)ceil_div(ceil_div(K, n_split), kKC) * kKC;
  if (k_per_split == 0) k_per_split = kKC;
  TnArgs a{};
  a.a[0] = A, a.b[0] = B, a.c = C, a.colsum = nullptr;
  a.lda = lda, a.ldb = ldb, a.ldc = ldc, a.split_stride = split_stride, a.batch_stride = 0;
  a.M = (int)M, a.N = (int)N, a.K = (int)K, a.k_per_split = k_per_split, a.n_split = n_split, a.batch = 1;
  return launch_tn(a, (cudaStream_t)stream_);
}

// Batched form: C[s][e] = sum over split s of A_e^T B_e for `batch` (<= 24) operand pairs of identical shape
// (the 20 weight gradients of a node layer's hypernetwork trunks in one launch).  A, B: HOST arrays of device
// pointers.  C: (n_split, batch, M, N) contiguous.  colsum (optional): (n_split, batch, M) column sums of A_e
// over the split's rows — the bias gradients, produced while the operand is staged.
extern "C" int cgat_gemm3x_tn_batched(const float* const* A, const float* const* B, int32_t batch, int64_t lda,
                                      int64_t ldb, float* C, float* colsum, int64_t M, int64_t N, int64_t K,
                                      int32_t n_split, void* stream_) {
  if (M <= 0 || N <= 0 || n_split <= 0 || batch <= 0) return 0;
  if (batch > kTnMaxBatch) return fail(-2, "cgat_gemm3x_tn_batched: at most 24 operand pairs per call");
  TnArgs a{};
  for (int e = 0; e < batch; ++e) {
    if (int err = check_tn(A[e], B[e], lda, ldb, M, N, K)) return err;
    a.a[e] = A[e], a.b[e] = B[e];
  }
  int k_per_split = (int)ceil_div(ceil_div(K, n_split), kKC) * kKC;
  if (k_per_split == 0) k_per_split = kKC;
  a.c = C, a.colsum = colsum;
  a.lda = lda, a.ldb = ldb, a.ldc = N, a.split_stride = (int64_t)batch * M * N, a.batch_stride = M * N;
  a.M = (int)M, a.N = (int)N, a.K = (int)K, a.k_per_split = k_per_split, a.n_split = n_split, a.batch = batch;
  return launch_tn(a, (cudaStream_t)stream_);
}
