// Hypernetwork linear layer, weight gradient on kind::f16 passes (SURVEY.md §8a row A12 for row A5; VERDICT r01 next #4:
// "move the gradient-operand kernels to f16x3 with a per-launch power-of-two scale").
//
//   dL/dW[o*F+i, k] = sum_n g[n,o] * y[n,i] * z[n,k]           (g = dL/dy_out)
//
// Same decomposition as hyper_bwd.cu (one CTA = pair of output channels x atom range; the scaled rows g[n,o]*y[n,:] are
// formed in registers while they are staged; both operands MN-major, contraction over atoms) with fp16 hi/lo operand
// pairs instead of tf32 hi/lo pairs: kind::f16 runs at twice the rate for the same operand bytes, a stage is 48 KB
// instead of 96 KB (four stages instead of two).  The gradient operand needs a range: the rows are multiplied by a
// power of two s = 2^(4 - ceil(log2 amax|g|)) read from device memory (amax is produced by cgat_hyper_rowscale_f16,
// which reads the same g just before), so max |g| s = 16 and fp16 has 2^12 of headroom left for |y|; the epilogue
// divides by s.  Small elements keep an ABSOLUTE error of 2^-25 / s, i.e. 2^-29 of the largest one (tc_common.cuh).
#include "common.cuh"
#include "tc_common.cuh"

#ifndef CGAT_HW_DBG
#define CGAT_HW_DBG 0   // timing experiments (F = 128 kernel): 1 no MMAs, 2 no operand conversion / stores, 4 no loads either, 8 no tail sums
#endif
namespace cgat {
namespace {
using namespace tc;

constexpr int kF = 128;
constexpr int kThreadsW = 128 + 512 + 32;
constexpr int kMmaWarpW = (128 + 512) / 32;
constexpr int kRows = 32;                     // atoms (K rows) per stage
constexpr int kImg = kRows * 128;             // one image: 64 columns x 32 rows of halves = 4 KB
constexpr int kPart = 2 * kImg;               // 128 columns = 2 images = 8 KB
constexpr int kStage = 6 * kPart;             // Z hi/lo, A0 hi/lo, A1 hi/lo = 48 KB
constexpr int kStagesW = 4;
constexpr int kTailRed = 16 * 32 * 16 * 4;    // [16 producer warps][32 lanes][4 sums x 4 columns] floats = 32 KB
constexpr int kSmemW = kStagesW * kStage + kTailRed + 256 + 1024;

__global__ void __launch_bounds__(kThreadsW, 1)
hyper_wgrad_f16_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ z,
                       const float* __restrict__ g_amax, float* __restrict__ out, float* __restrict__ tail,
                       int n_atoms, int n_split) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* tail_red = reinterpret_cast<float*>(smem + kStagesW * kStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStagesW * kStage + kTailRed);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStagesW;
  uint64_t* accum = bars + 2 * kStagesW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.x / n_split, split = blockIdx.x % n_split;
  const int o0 = 2 * pair;
  const int n_lo = (int)((int64_t)n_atoms * split / n_split), n_hi = (int)((int64_t)n_atoms * (split + 1) / n_split);
  const int n_chunks = (n_hi - n_lo + kRows - 1) / kRows;
  // power-of-two scale of the gradient operand: max |g| * s in (8, 16]
  const float amax = __ldg(g_amax);
  int ex;
  frexpf(amax, &ex);                                           // amax = m * 2^ex, m in [0.5, 1)
  const float s = (amax > 0.f && amax < INFINITY) ? ldexpf(1.f, 4 - ex) : 1.f;
  const float s_inv = (amax > 0.f && amax < INFINITY) ? ldexpf(1.f, ex - 4) : 1.f;

  if (tid == 0) {
    for (int st = 0; st < kStagesW; ++st) {
      mbar_init(&full[st], 256);
      mbar_init(&empty[st], 1);
    }
    mbar_init(accum, 1);
    mbar_init_fence();
  }
  if (warp == kMmaWarpW) tmem_alloc(tmem_slot, 512);
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
      float* dst = out + ((int64_t)split * kF * kF + (int64_t)(o0 + oo) * kF + i) * kF;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float v[32], w[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + oo * 256 + cc * 32, v);
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + oo * 256 + 128 + cc * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 r;
          r.x = n_chunks ? fmaf(w[4 * j], kF16LoInv, v[4 * j]) * s_inv : 0.f;
          r.y = n_chunks ? fmaf(w[4 * j + 1], kF16LoInv, v[4 * j + 1]) * s_inv : 0.f;
          r.z = n_chunks ? fmaf(w[4 * j + 2], kF16LoInv, v[4 * j + 2]) * s_inv : 0.f;
          r.w = n_chunks ? fmaf(w[4 * j + 3], kF16LoInv, v[4 * j + 3]) * s_inv : 0.f;
          reinterpret_cast<float4*>(dst + cc * 32)[j] = r;
        }
      }
    }
    tc_fence_before();
  } else if (warp < kMmaWarpW) {
    // 16 producer warps in two groups that take alternate chunks: with the MMAs twice as fast as the tf32 form the
    // conversion work of 8 warps (two per scheduler, dependent ALU chains) was pacing the kernel, and a software
    // prefetch of the next chunk's rows changed nothing (measured) — it is issue slots / latency hiding, not loads.
    const int pt = tid - 128, grp = pt >> 8, pl = pt & 255;
    // Bias-shaped gradients of the same Linear come for free from the rows staged here: the column sums of the scaled
    // rows, sum_n g[n,o] y[n,:], are dL/db[o*F : (o+1)*F], and sum_n g[n,o] z[n,:] is dL/dW[F*F + o, :] — a thread
    // always holds the same 4 columns (q = pl % 32), so it keeps four float4 running sums; they replace a separate
    // split-K GEMM g^T [y | z] per hyper-layer (20 launches of ~45 us per cfg2 train step).
    float4 sy0 = make_float4(0.f, 0.f, 0.f, 0.f), sy1 = sy0, sz0 = sy0, sz1 = sy0;
    for (int ch = grp; ch < n_chunks; ch += 2) {
      const int st = ch % kStagesW, u = ch / kStagesW;
      const int n0 = n_lo + ch * kRows;
      float4 zv[4], yv[4];
      float g0[4], g1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = pl + 256 * j, r = idx >> 5, q = idx & 31;
        const int n = n0 + r;
        zv[j] = yv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        g0[j] = g1[j] = 0.f;
        if (n < n_hi && !(CGAT_HW_DBG & 4)) {
          zv[j] = __ldg(reinterpret_cast<const float4*>(z + (int64_t)n * kF) + q);
          yv[j] = __ldg(reinterpret_cast<const float4*>(y + (int64_t)n * kF) + q);
          const float2 gg = __ldg(reinterpret_cast<const float2*>(g + (int64_t)n * kF + o0));
          g0[j] = gg.x * s, g1[j] = gg.y * s;
        }
      }
      mbar_wait(&empty[st], (u + 1) & 1u);
      uint8_t* sb = smem + st * kStage;
#pragma unroll
      for (int j = 0; j < ((CGAT_HW_DBG & 2) ? 0 : 4); ++j) {
        const int idx = pl + 256 * j, r = idx >> 5, q = idx & 31;
        const uint32_t off = (q >> 4) * kImg + mn16_offset(r, q & 15);
        uint2 hi, lo;
        split_f16x4s(zv[j], kF16LoScale, hi, lo);
        *reinterpret_cast<uint2*>(sb + off) = hi;
        *reinterpret_cast<uint2*>(sb + kPart + off) = lo;
        float4 a = make_float4(yv[j].x * g0[j], yv[j].y * g0[j], yv[j].z * g0[j], yv[j].w * g0[j]);
        sy0.x += a.x, sy0.y += a.y, sy0.z += a.z, sy0.w += a.w;
        sz0.x = fmaf(zv[j].x, g0[j], sz0.x), sz0.y = fmaf(zv[j].y, g0[j], sz0.y);
        sz0.z = fmaf(zv[j].z, g0[j], sz0.z), sz0.w = fmaf(zv[j].w, g0[j], sz0.w);
        split_f16x4s(a, kF16LoScale, hi, lo);
        *reinterpret_cast<uint2*>(sb + 2 * kPart + off) = hi;
        *reinterpret_cast<uint2*>(sb + 3 * kPart + off) = lo;
        a = make_float4(yv[j].x * g1[j], yv[j].y * g1[j], yv[j].z * g1[j], yv[j].w * g1[j]);
        sy1.x += a.x, sy1.y += a.y, sy1.z += a.z, sy1.w += a.w;
        sz1.x = fmaf(zv[j].x, g1[j], sz1.x), sz1.y = fmaf(zv[j].y, g1[j], sz1.y);
        sz1.z = fmaf(zv[j].z, g1[j], sz1.z), sz1.w = fmaf(zv[j].w, g1[j], sz1.w);
        split_f16x4s(a, kF16LoScale, hi, lo);
        *reinterpret_cast<uint2*>(sb + 4 * kPart + off) = hi;
        *reinterpret_cast<uint2*>(sb + 5 * kPart + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(&full[st]);
    }
    if (tail != nullptr) {
      // fixed-order reduction over the 16 producer warps (deterministic): [warp][lane][sum][component]
      float4* mine = reinterpret_cast<float4*>(tail_red) + ((pt >> 5) * 32 + lane) * 4;
      mine[0] = sy0, mine[1] = sy1, mine[2] = sz0, mine[3] = sz1;
      asm volatile("bar.sync 1, %0;" ::"n"(512) : "memory");
      // thread -> (sum kind = pt / 128: y of o0, y of o0+1, z of o0, z of o0+1; column = pt % 128)
      const int kind = pt >> 7, col = pt & 127;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) acc += tail_red[((w * 32 + (col >> 2)) * 4 + kind) * 4 + (col & 3)];
      // tail (n_split, F, 2F): row o = [sum_n g y | sum_n g z]; the rows were scaled by s on the way in
      tail[((int64_t)split * kF + o0 + (kind & 1)) * 2 * kF + (kind >> 1) * kF + col] = acc * s_inv;
    }
  } else {
    constexpr uint32_t idesc = umma_idesc_f16_mn(128, 128), idesc2 = umma_idesc_f16_mn(128, 256);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int st = ch % kStagesW, u = ch / kStagesW;
      mbar_wait(&full[st], u & 1u);
      tc_fence_after();
      {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
        const uint32_t z_hi = smem_u32(smem + st * kStage), z_lo = z_hi + kPart;
#pragma unroll
        for (int oo = 0; oo < 2; ++oo) {
          const uint32_t a_hi = z_hi + (2 + 2 * oo) * kPart, a_lo = a_hi + kPart;
          const uint32_t d = tmem + oo * 256;
#pragma unroll
          for (int ks = 0; ks < kRows / 16; ++ks) {   // one MMA consumes 16 K-rows = two 1024-byte groups of every image
            const uint32_t o = ks * 2048;
            if (CGAT_HW_DBG & 1) continue;
            // the hi and lo images of z are adjacent (4 images of 64 columns): ONE N = 256 MMA multiplies a_hi with both
            // and writes main | correction columns, a second N = 128 MMA adds a_lo * z_hi to the correction columns —
            // 20 KB of shared-memory operand reads per K step instead of 24 (the kernel runs at the shared-memory port:
            // operand reads + the producers' stores, profiles/r03z)
            (void)z_lo;
            umma_f16_e(d, umma_desc_mn_sw128_16b(a_hi + o, kImg), umma_desc_mn_sw128_16b(z_hi + o, kImg), idesc2,
                     (ch | ks) != 0);
            umma_f16_e(d + 128, umma_desc_mn_sw128_16b(a_lo + o, kImg), umma_desc_mn_sw128_16b(z_hi + o, kImg), idesc, 1);
          }
        }
        umma_commit_e(&empty[st]);
        if (ch == n_chunks - 1) umma_commit_e(accum);
      }
      __syncwarp();
    }
    if (n_chunks == 0 && lane == 0) mbar_arrive(accum);
  }
  __syncthreads();
  if (warp == kMmaWarpW) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---- F = 256 (BASELINE.json configs[3]) --------------------------------------------------------------------------
// One CTA = (output channel o, 128-row half of its F x F block, atom range): A = g[:,o] * y[:, half] (128 columns),
// B = z (256 columns) -> one N = 256 accumulator (main | correction = all 512 TMEM columns).  Same stage size (48 KB).
constexpr int kF2 = 256;
constexpr int kPartZ2 = 4 * kImg;              // 256 columns = 4 images = 16 KB
constexpr int kStage2 = 2 * kPartZ2 + 2 * kPart;   // Z hi/lo + A hi/lo = 48 KB

__global__ void __launch_bounds__(kThreadsW, 1)
hyper_wgrad_f16_256_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ z,
                           const float* __restrict__ g_amax, float* __restrict__ out, int n_atoms, int n_split) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStagesW * kStage2);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStagesW;
  uint64_t* accum = bars + 2 * kStagesW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = blockIdx.x / n_split, split = blockIdx.x % n_split;
  const int o = item >> 1, half = item & 1;
  const int n_lo = (int)((int64_t)n_atoms * split / n_split), n_hi = (int)((int64_t)n_atoms * (split + 1) / n_split);
  const int n_chunks = (n_hi - n_lo + kRows - 1) / kRows;
  const float amax = __ldg(g_amax);
  int ex;
  frexpf(amax, &ex);
  const float s = (amax > 0.f && amax < INFINITY) ? ldexpf(1.f, 4 - ex) : 1.f;
  const float s_inv = (amax > 0.f && amax < INFINITY) ? ldexpf(1.f, ex - 4) : 1.f;

  if (tid == 0) {
    for (int st = 0; st < kStagesW; ++st) {
      mbar_init(&full[st], 256);
      mbar_init(&empty[st], 1);
    }
    mbar_init(accum, 1);
    mbar_init_fence();
  }
  if (warp == kMmaWarpW) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    const int i = warp * 32 + lane;  // row of the half block = input channel half*128 + i
    mbar_wait(accum, 0);
    tc_fence_after();
    float* dst = out + ((int64_t)split * kF2 * kF2 + (int64_t)o * kF2 + half * 128 + i) * kF2;
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
      float v[32], w[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + cc * 32, w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 r;
        r.x = n_chunks ? fmaf(w[4 * j], kF16LoInv, v[4 * j]) * s_inv : 0.f;
        r.y = n_chunks ? fmaf(w[4 * j + 1], kF16LoInv, v[4 * j + 1]) * s_inv : 0.f;
        r.z = n_chunks ? fmaf(w[4 * j + 2], kF16LoInv, v[4 * j + 2]) * s_inv : 0.f;
        r.w = n_chunks ? fmaf(w[4 * j + 3], kF16LoInv, v[4 * j + 3]) * s_inv : 0.f;
        reinterpret_cast<float4*>(dst + cc * 32)[j] = r;
      }
    }
    tc_fence_before();
  } else if (warp < kMmaWarpW) {
    const int pt = tid - 128, grp = pt >> 8, pl = pt & 255;
    for (int ch = grp; ch < n_chunks; ch += 2) {
      const int st = ch % kStagesW, u = ch / kStagesW;
      const int n0 = n_lo + ch * kRows;
      float4 zv[8], yv[4];
      float gs[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // z: 32 rows x 64 float4
        const int idx = pl + 256 * j, r = idx >> 6, q = idx & 63;
        const int n = n0 + r;
        zv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < n_hi) zv[j] = __ldg(reinterpret_cast<const float4*>(z + (int64_t)n * kF2) + q);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // y half: 32 rows x 32 float4
        const int idx = pl + 256 * j, r = idx >> 5, q = idx & 31;
        const int n = n0 + r;
        yv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        gs[j] = 0.f;
        if (n < n_hi) {
          yv[j] = __ldg(reinterpret_cast<const float4*>(y + (int64_t)n * kF2 + half * 128) + q);
          gs[j] = __ldg(g + (int64_t)n * kF2 + o) * s;
        }
      }
      mbar_wait(&empty[st], (u + 1) & 1u);
      uint8_t* sb = smem + st * kStage2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = pl + 256 * j, r = idx >> 6, q = idx & 63;
        const uint32_t off = (q >> 4) * kImg + mn16_offset(r, q & 15);
        uint2 hi, lo;
        split_f16x4s(zv[j], kF16LoScale, hi, lo);
        *reinterpret_cast<uint2*>(sb + off) = hi;
        *reinterpret_cast<uint2*>(sb + kPartZ2 + off) = lo;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = pl + 256 * j, r = idx >> 5, q = idx & 31;
        const uint32_t off = (q >> 4) * kImg + mn16_offset(r, q & 15);
        const float4 a = make_float4(yv[j].x * gs[j], yv[j].y * gs[j], yv[j].z * gs[j], yv[j].w * gs[j]);
        uint2 hi, lo;
        split_f16x4s(a, kF16LoScale, hi, lo);
        *reinterpret_cast<uint2*>(sb + 2 * kPartZ2 + off) = hi;
        *reinterpret_cast<uint2*>(sb + 2 * kPartZ2 + kPart + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(&full[st]);
    }
  } else {
    constexpr uint32_t idesc = umma_idesc_f16_mn(128, 256);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int st = ch % kStagesW, u = ch / kStagesW;
      mbar_wait(&full[st], u & 1u);
      tc_fence_after();
      {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
        const uint32_t z_hi = smem_u32(smem + st * kStage2), z_lo = z_hi + kPartZ2;
        const uint32_t a_hi = z_hi + 2 * kPartZ2, a_lo = a_hi + kPart;
#pragma unroll
        for (int ks = 0; ks < kRows / 16; ++ks) {
          const uint32_t o2 = ks * 2048;
          umma_f16_e(tmem + 256, umma_desc_mn_sw128_16b(a_lo + o2, kImg), umma_desc_mn_sw128_16b(z_hi + o2, kImg), idesc,
                   (ch | ks) != 0);
          umma_f16_e(tmem + 256, umma_desc_mn_sw128_16b(a_hi + o2, kImg), umma_desc_mn_sw128_16b(z_lo + o2, kImg), idesc, 1);
          umma_f16_e(tmem, umma_desc_mn_sw128_16b(a_hi + o2, kImg), umma_desc_mn_sw128_16b(z_hi + o2, kImg), idesc,
                   (ch | ks) != 0);
        }
        umma_commit_e(&empty[st]);
        if (ch == n_chunks - 1) umma_commit_e(accum);
      }
      __syncwarp();
    }
    if (n_chunks == 0 && lane == 0) mbar_arrive(accum);
  }
  __syncthreads();
  if (warp == kMmaWarpW) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

// cgat_hyper_wgrad on kind::f16 passes: same result layout, (cgat_hyper_wgrad_splits(N), F*F, F) partial dL/dW[:F*F]
// for F = 128; for F = 256 a single (F*F, F) result (no split over atoms).
// g_amax: device float holding max |g| (written by cgat_hyper_rowscale_f16 with the same g as `scale`).
// tail (optional, F = 128 only): (cgat_hyper_wgrad_splits(N), F, 2F) partial [g^T y | g^T z] — the bias-shaped
// gradients of the same Linear (dL/db[:F*F] rows and dL/dW[F*F:]), accumulated from the rows the kernel stages anyway.
extern "C" int cgat_hyper_wgrad_f16(const float* g, const float* y, const float* z, const float* g_amax, float* out,
                                    float* tail, int64_t n_atoms, int32_t f, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (f != kF && f != kF2) return fail(-2, "cgat_hyper_wgrad_f16: instantiated for F = 128 and F = 256");
  if (n_atoms >= (1ll << 31) - 64) return fail(-2, "cgat_hyper_wgrad_f16: too many atoms");
  if (g_amax == nullptr) return fail(-2, "cgat_hyper_wgrad_f16: g_amax is required");
  if (n_atoms <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(hyper_wgrad_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemW));
    configured = true;
  }
  if (f == kF2 && tail != nullptr) return fail(-2, "cgat_hyper_wgrad_f16: the tail output exists for F = 128 only");
  if (f == kF2) {
    // (o, row half) x splits: 512 CTAs already fill the SMs three times over, so the atoms are not split
    static bool configured2 = false;
    if (!configured2) {
      CGAT_CUDA(cudaFuncSetAttribute(hyper_wgrad_f16_256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemW));
      configured2 = true;
    }
    hyper_wgrad_f16_256_kernel<<<2 * f, kThreadsW, kSmemW, stream>>>(g, y, z, g_amax, out, (int)n_atoms, 1);
    return check_launch("hyper_wgrad_f16_256_kernel");
  }
  const int n_split = cgat_hyper_wgrad_splits(n_atoms);
  hyper_wgrad_f16_kernel<<<(f / 2) * n_split, kThreadsW, kSmemW, stream>>>(g, y, z, g_amax, out, tail, (int)n_atoms,
                                                                         n_split);
  return check_launch("hyper_wgrad_f16_kernel");
}
