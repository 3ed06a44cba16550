// Fused edge-attention backward, step 3 on kind::f16 passes (SURVEY.md §8a row A12 for rows A2-A4).
//
//   dW2[net, h][c, k] = sum over edges t of dZ[t, h, c] * hid[t, h, k]
//
// Same decomposition as edge_wgrad_kernel (edge_attn_bwd.cu): one CTA = ((net, head), edge range), both operands staged
// MN-major (dZ rows as they are, hid rows re-gathered like the forward), contraction over EDGES, partial results summed
// by the caller.  Differences:
//   * fp16 hi/lo operand pairs (SWIZZLE_128B MN-major, tc_common.cuh): twice the tensor rate, 48 KB stages -> 4 stages;
//   * the gradient operand dZ is multiplied by a power of two s = 2^(4 - ceil(log2 amax|dZ|)) read from device memory
//     (cgat_edge_attn_bwd_prep leaves amax there), the epilogue divides by s;
//   * 16 producer warps in two groups that take alternate 32-edge chunks, so one group's gather -> wait -> convert ->
//     store chain overlaps the other's (the tf32 kernel's 8 producer warps were waiting on their gathers: ncu r01n,
//     tensor pipe 23 % active).
#include "common.cuh"
#include "tc_common.cuh"

#ifndef CGAT_EW_DBG
#define CGAT_EW_DBG 0   // timing experiments: 1 no MMAs, 2 no hidden-row gathers / conversion, 4 no dZ loads / conversion
#endif
namespace cgat {
namespace {
using namespace tc;

constexpr int kWF = 128;                 // channels per head
constexpr int kWGroup = 256;             // producer threads per group
constexpr int kWThreads = 128 + 2 * kWGroup + 32;
constexpr int kWMmaWarp = (128 + 2 * kWGroup) / 32;
constexpr int kWRows = 32;               // edges (K rows) per stage
constexpr int kWImg = kWRows * 128;      // image: 64 columns x 32 rows of halves = 4 KB
constexpr int kWA = 2 * kWImg;           // dZ: 128 channels = 2 images = 8 KB
constexpr int kWB = 4 * kWImg;           // hid: up to 256 hidden units = 4 images = 16 KB
constexpr int kWStage = 2 * kWA + 2 * kWB;   // 48 KB
constexpr int kWStagesN = 4;
constexpr int kWSmem = kWStagesN * kWStage + 2 * 2 * 3 * 32 * 4 + 256 + 1024;

struct WArgs {
  const float* P;
  const float* T;
  const int32_t* src;
  const int32_t* dst;
  const int32_t* rank;
  const float* d_gate;
  const float* d_msg;
  const float* dz_amax;   // device float: max |dZ| over both nets
  float* out;             // (n_split, 2, H, F, Hd) partial dW2
  int n_edges, heads, hd, n_split;
  int f, n_ch, n_kh;      // channels per head; 128-channel blocks per head; 256-wide blocks of the hidden units
};

__device__ __forceinline__ float lrelu_w(float x) { return fmaxf(x, 0.01f * x); }   // == (x > 0 ? x : 0.01 x), one instruction less

__global__ void __launch_bounds__(kWThreads, 1) edge_wgrad_f16_kernel(const WArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWStagesN * kWStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWStagesN;
  uint64_t* accum = bars + 2 * kWStagesN;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  int32_t* meta = reinterpret_cast<int32_t*>(tmem_slot + 2);  // [2 groups][2 parity][3][32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = g.heads, hd = g.hd, hhd = H * hd;
  // item = (net, head, 128-channel block, 256-hidden-unit block): one 128 x 256 accumulator tile of dW2[net, head]
  const int item = blockIdx.x / g.n_split, split = blockIdx.x % g.n_split;
  const int kh = item % g.n_kh, chb = (item / g.n_kh) % g.n_ch, h = (item / (g.n_kh * g.n_ch)) % H;
  const int net = item / (g.n_kh * g.n_ch * H);
  const int F = g.f, c0 = chb * 128, k0 = kh * 256;
  const int e_lo = (int)((int64_t)g.n_edges * split / g.n_split), e_hi = (int)((int64_t)g.n_edges * (split + 1) / g.n_split);
  const int n_chunks = (e_hi - e_lo + kWRows - 1) / kWRows;
  const int64_t ldp = 4 * (int64_t)hhd, ldt = 2 * (int64_t)hhd;
  const float amax = __ldg(g.dz_amax);
  int ex;
  frexpf(amax, &ex);
  const bool ok_amax = amax > 0.f && amax < INFINITY;
  const float s = ok_amax ? ldexpf(1.f, 4 - ex) : 1.f, s_inv = ok_amax ? ldexpf(1.f, ex - 4) : 1.f;

  if (tid == 0) {
    for (int st = 0; st < kWStagesN; ++st) {
      mbar_init(&full[st], kWGroup);
      mbar_init(&empty[st], 1);
    }
    mbar_init(accum, 1);
    mbar_init_fence();
  }
  if (warp == kWMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // epilogue: one accumulation over the whole edge range, written once
    const int c = warp * 32 + lane;
    mbar_wait(accum, 0);
    tc_fence_after();
    float* dst = g.out + ((((int64_t)split * 2 + net) * H + h) * F + c0 + c) * hd + k0;
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
      float v[32], w[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + cc * 32, w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (k0 + cc * 32 + j < hd) dst[cc * 32 + j] = n_chunks > 0 ? fmaf(w[j], kF16LoInv, v[j]) * s_inv : 0.f;
    }
    tc_fence_before();
  } else if (warp < kWMmaWarp) {
    const int pt = tid - 128, grp = pt >> 8, pl = pt & (kWGroup - 1);
    const float* dz = net ? g.d_msg : g.d_gate;
    int32_t* gmeta = meta + grp * 192;
    // edge ids of this group's NEXT chunk are fetched while the current one is staged (threads 0-31 of the group)
    int nd = -1, ns = 0, nr = 0;
    if (pl < 32 && grp < n_chunks) {
      const int e0 = e_lo + grp * kWRows;
      const bool ok = pl < min(kWRows, e_hi - e0);
      nd = ok ? g.dst[e0 + pl] : -1, ns = ok ? g.src[e0 + pl] : 0, nr = ok ? g.rank[e0 + pl] : 0;
    }
    int it = 0;
    for (int ch = grp; ch < n_chunks; ch += 2, ++it) {
      const int st = ch % kWStagesN, u = ch / kWStagesN;
      const int e0 = e_lo + ch * kWRows;
      const int nv = min(kWRows, e_hi - e0);
      int32_t* mt = gmeta + (it & 1) * 96;
      if (pl < 32) {
        mt[pl] = nd, mt[32 + pl] = ns, mt[64 + pl] = nr;
        const int e1 = e0 + 2 * kWRows;
        const bool ok = ch + 2 < n_chunks && pl < min(kWRows, e_hi - e1);
        nd = ok ? g.dst[e1 + pl] : -1, ns = ok ? g.src[e1 + pl] : 0, nr = ok ? g.rank[e1 + pl] : 0;
      }
      if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(kWGroup) : "memory");
      else asm volatile("bar.sync 2, %0;" ::"n"(kWGroup) : "memory");
      // issue every load of this chunk before waiting for the stage: dZ rows, then the three gathered rows per slot
      float4 a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = pl + kWGroup * j, r = idx >> 5, q = idx & 31;
        a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nv && !(CGAT_EW_DBG & 4)) a[j] = __ldg(reinterpret_cast<const float4*>(dz + ((int64_t)(e0 + r) * H + h) * F + c0 + q * 4));
      }
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = pl + kWGroup * j, r = idx >> 6, q = idx & 63;
        const int d = mt[r];
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d >= 0 && k0 + q * 4 < hd && !(CGAT_EW_DBG & 2)) {
          const int col = net * hhd + h * hd + k0 + q * 4;
          const float4 pd = __ldg(reinterpret_cast<const float4*>(g.P + (int64_t)d * ldp + col));
          const float4 ps = __ldg(reinterpret_cast<const float4*>(g.P + (int64_t)mt[32 + r] * ldp + 2 * hhd + col));
          const float4 te = __ldg(reinterpret_cast<const float4*>(g.T + (int64_t)mt[64 + r] * ldt + col));
          x[j].x = lrelu_w(pd.x + ps.x + te.x), x[j].y = lrelu_w(pd.y + ps.y + te.y);
          x[j].z = lrelu_w(pd.z + ps.z + te.z), x[j].w = lrelu_w(pd.w + ps.w + te.w);
        }
      }
      mbar_wait(&empty[st], (u + 1) & 1u);
      uint8_t* sb = stages + st * kWStage;
      // A operand: dZ * s (32 edges x 128 channels), images of 64 channels
#pragma unroll
      for (int j = 0; j < ((CGAT_EW_DBG & 4) ? 0 : 4); ++j) {
        const int idx = pl + kWGroup * j, r = idx >> 5, q = idx & 31;
        const float4 as = make_float4(a[j].x * s, a[j].y * s, a[j].z * s, a[j].w * s);
        uint2 hi, lo;
        split_f16x4s(as, kF16LoScale, hi, lo);
        const uint32_t off = (q >> 4) * kWImg + mn16_offset(r, q & 15);
        *reinterpret_cast<uint2*>(sb + off) = hi;
        *reinterpret_cast<uint2*>(sb + kWA + off) = lo;
      }
      // B operand: hidden activations (32 edges x Hd)
      uint8_t* bh = sb + 2 * kWA;
#pragma unroll
      for (int j = 0; j < ((CGAT_EW_DBG & 2) ? 0 : 8); ++j) {
        const int idx = pl + kWGroup * j, r = idx >> 6, q = idx & 63;
        uint2 hi, lo;
        split_f16x4s(x[j], kF16LoScale, hi, lo);
        const uint32_t off = (q >> 4) * kWImg + mn16_offset(r, q & 15);
        *reinterpret_cast<uint2*>(bh + off) = hi;
        *reinterpret_cast<uint2*>(bh + kWB + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(&full[st]);
    }
  } else {
    const uint32_t idesc = umma_idesc_f16_mn(128, 256);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int st = ch % kWStagesN, u = ch / kWStagesN;
      mbar_wait(&full[st], u & 1u);
      tc_fence_after();
      {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
        const uint32_t a_hi = smem_u32(stages + st * kWStage), a_lo = a_hi + kWA;
        const uint32_t b_hi = a_hi + 2 * kWA, b_lo = b_hi + kWB;
#pragma unroll
        for (int ks = 0; ks < kWRows / 16; ++ks) {
          const uint32_t o = ks * 2048;
          if (CGAT_EW_DBG & 1) continue;
          umma_f16_e(tmem + 256, umma_desc_mn_sw128_16b(a_lo + o, kWImg), umma_desc_mn_sw128_16b(b_hi + o, kWImg), idesc,
                   (ch | ks) != 0);
          umma_f16_e(tmem + 256, umma_desc_mn_sw128_16b(a_hi + o, kWImg), umma_desc_mn_sw128_16b(b_lo + o, kWImg), idesc, 1);
          umma_f16_e(tmem, umma_desc_mn_sw128_16b(a_hi + o, kWImg), umma_desc_mn_sw128_16b(b_hi + o, kWImg), idesc,
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
  if (warp == kWMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

// number of edge ranges (= partial results) of cgat_edge_attn_wgrad_f16: about one CTA per SM over all accumulator tiles
extern "C" int32_t cgat_edge_attn_wgrad_f16_splits(int32_t heads, int32_t f, int32_t hd) {
  const int items = 2 * heads * (f / kWF) * ((hd + 255) / 256);
  const int s = kNumSMs / (items > 0 ? items : 1);
  return s < 1 ? 1 : s;
}

// cgat_edge_attn_wgrad on kind::f16 passes: same arguments and result layout plus dz_amax, the device float that
// cgat_edge_attn_bwd_prep[_f16] leaves behind (max |d_gate|, |d_msg|).
extern "C" int cgat_edge_attn_wgrad_f16(const float* P, const float* T, const int32_t* src, const int32_t* dst,
                                        const int32_t* rank, const float* d_gate, const float* d_msg,
                                        const float* dz_amax, float* out, int64_t n_edges, int32_t heads, int32_t f,
                                        int32_t hd, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (f != kWF && f != 2 * kWF) return fail(-2, "cgat_edge_attn_wgrad_f16: F must be 128 or 256");
  if (heads < 1 || heads > 8 || hd <= 0 || (hd & 3) || hd > 512)
    return fail(-2, "cgat_edge_attn_wgrad_f16: hidden width must be a multiple of 4, at most 512");
  if (dz_amax == nullptr) return fail(-2, "cgat_edge_attn_wgrad_f16: dz_amax is required");
  if (n_edges <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(edge_wgrad_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem));
    configured = true;
  }
  const int n_split = cgat_edge_attn_wgrad_f16_splits(heads, f, hd);
  const int n_ch = f / kWF, n_kh = (hd + 255) / 256;
  WArgs a{P, T, src, dst, rank, d_gate, d_msg, dz_amax, out, (int)n_edges, heads, hd, n_split, f, n_ch, n_kh};
  edge_wgrad_f16_kernel<<<2 * heads * n_ch * n_kh * n_split, kWThreads, kWSmem, stream>>>(a);
  return check_launch("edge_wgrad_f16_kernel");
}
