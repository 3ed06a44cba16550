// Fused hypernetwork linear layer, forward (kernel K5).
//
// Reference arithmetic (CGAT/Hypernetworksmp.py:243-254 HyperLinear.forward + :205-209 BatchLinear):
//     p[n,:]   = Wl z[n,:] + bl                 (Linear F -> F*F+F, the last layer of FCBlock)
//     y_out[n,o] = sum_i p[n, o*F+i] * y_in[n,i] + p[n, F*F+o]
// As written the reference materialises p: (N, F*F+F) = 66 KB per atom per hyper-layer in HBM.
// Here p never leaves the SM: for each output channel o the tile D_o[n,i] = sum_k z[n,k] Wl[o*F+i,k]
// is produced by tcgen05 MMAs (3xTF32, accumulators in TMEM, atoms on the 128 TMEM lanes) and is
// contracted against y_in[n,:] by the thread that owns lane n while the next tile is being computed
// (two TMEM accumulator buffers).  The bias-shaped remainder
//     e[n,o] = sum_i bl[o*F+i] y_in[n,i] + sum_k Wl[F*F+o,k] z[n,k] + bl[F*F+o]
// is a plain (N x 2F) x (2F x F) product computed beforehand by cgat_gemm3x_nt and added here.
//
// Roles (416 threads, one persistent CTA per SM):
//   warps 0-7   epilogue, two groups of four warps, each owning half of the accumulator columns: half a y_in row in
//               registers, tcgen05.ld + FMA row-dot, store y_out
//   warps 8-11  stage the z tile (fp32 -> tf32 hi/lo, SWIZZLE_128B K-major); warp 8 lane 0 then streams the
//               pre-packed weight stages with cp.async.bulk (32 KB each) through a 3-deep mbarrier ring
//   warp  12    TMEM allocation + single-thread MMA issue
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cgat {
namespace {
using namespace tc;

// (A variant with the z tile in TENSOR memory — tcgen05.st by the stagers, TS-mode MMAs, 7-stage weight ring, single
// accumulator buffer — was measured slower in round 1, 140 vs 112 us backward, and has been removed: DESIGN.md §3.)
template <int F>
struct HyperCfg {
  static constexpr int kKC = F / kPackChunk;                    // K chunks of 32 floats
  static constexpr int kABytes = kKC * (int)kPackStageBytes;    // z tile in shared memory, hi+lo per chunk
  static constexpr int kStages = 3;                             // + 32 KB weight stages = 224 KB
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kABytes + kStages * (int)kPackStageBytes + 1024 + kBarBytes;
  static constexpr int kThreads = 416;
  // two buffers x (main, correction) accumulators.  The tensor core truncates the fp32 accumulator on
  // every MMA; keeping the 2^-11-sized lo*hi / hi*lo terms in their own accumulator means the large
  // hi*hi sum sees K/8 truncations instead of 3K/8, and the two are added once, rounded to nearest.
  static constexpr int kTmemCols = 4 * F;
};

// kMode 0 (forward):   y_out[n,o] = sum_j D_o[n,j] * y_in[n,j] + e_term[n,o]
// kMode 1 (backward):  partial[chunk][n,j] = sum_{o in chunk} y_in[n,o] * D_o[n,j]      (y_in carries dL/dy_out)
//   with D_o[n,j] = sum_m z[n,m] Wblk_o[j,m].  Backward uses it twice: (z, W blocks) -> dL/dy_in and
//   (y, transposed W blocks) -> dL/dz, i.e. both activation gradients of the hyper-linear layer without ever
//   forming the (N, F*F) predicted-weight tensor or its gradient.
template <int F, int kMode>
__global__ void __launch_bounds__(HyperCfg<F>::kThreads, 1)
hyper_rowdot_fwd_kernel(const float* __restrict__ z, const float* __restrict__ y_in, const float* __restrict__ e_term,
                        const float* __restrict__ e_term2, const float* __restrict__ w_bias,
                        const float* __restrict__ w_packed, float* __restrict__ y_out, int n_atoms, int oc, int n_slots) {
  using Cfg = HyperCfg<F>;
  static_assert(F == 128, "row-in-registers epilogue is instantiated for F = 128");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                   // [kKC][hi|lo][16 KB]
  uint8_t* b_smem = smem + Cfg::kABytes;                    // [kStages][hi|lo][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + Cfg::kStages * kPackStageBytes);
  uint64_t* full = bars;                                    // [kStages] TMA -> MMA
  uint64_t* empty = bars + Cfg::kStages;                    // [kStages] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;            // [2] MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;                     // [2] epilogue -> MMA
  uint64_t* a_full = tmem_empty + 2;                        // stagers -> MMA
  uint64_t* a_free = a_full + 1;                            // MMA -> stagers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (n_atoms + 127) / 128;
  const int n_chunks = F / oc;
  const int n_items = n_tiles * n_chunks;
  // work item = (atom tile, chunk of output channels), chunk fastest; every CTA owns a CONTIGUOUS range, so the
  // z tile (A operand, 64 KB) is restaged only when the atom tile changes (~once or twice per CTA)
  const int item_lo = (int)((int64_t)n_items * blockIdx.x / gridDim.x);
  const int item_hi = (int)((int64_t)n_items * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 256);
    }
    mbar_init(a_full, 128);
    mbar_init(a_free, 1);
    mbar_init_fence();
  }
  if (warp == 12) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue
    // Two groups of four warps (a warp reads the TMEM lanes 32*(warp%4)...): group g owns columns [64g, 64g+64) of
    // every accumulator, so a thread keeps half a row (64 registers) and the per-tile epilogue time is halved — it,
    // not the MMAs, was pacing the kernel.  kMode 0: the two half dot products of an output meet in y_out itself
    // (group 1 stores its half, a named barrier, group 0 adds its own + the e term); kMode 1 needs no exchange.
    constexpr int HF = F / 2;
    const int grp = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ocount = 0;
    float y[HF];  // kMode 0: this atom's half y_in row; kMode 1: the running partial sums over o
    for (int item = item_lo; item < item_hi; ++item) {
      const int tile = item / n_chunks, chunk = item - tile * n_chunks;
      const int n = tile * 128 + (warp & 3) * 32 + lane;
      const bool valid = n < n_atoms;
      // kMode 1: the sums are carried across the consecutive chunks of one tile (see hyper_slots, common.cuh)
      if (kMode == 0 || item == item_lo || chunk == 0) {
#pragma unroll
        for (int j = 0; j < HF / 4; ++j) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kMode == 0 && valid) t = __ldg(reinterpret_cast<const float4*>(y_in + (int64_t)n * F + grp * HF) + j);
          y[4 * j] = t.x, y[4 * j + 1] = t.y, y[4 * j + 2] = t.z, y[4 * j + 3] = t.w;
        }
      }
      float accs[16];  // kMode 0: half dot products of this item's <= 16 outputs
      for (int oi = 0; oi < oc; ++oi, ++ocount) {
        const int o = chunk * oc + oi;
        const uint32_t b = ocount & 1u;
        const float sc = (kMode == 1 && valid) ? __ldg(y_in + (int64_t)n * F + o) : 0.f;
        mbar_wait(&tmem_full[b], (ocount >> 1) & 1u);
        tc_fence_after();
        float acc = 0.f;
        const uint32_t tb = tmem + lane_base + b * 2 * F + grp * HF;
#pragma unroll
        for (int cc = 0; cc < HF / 16; ++cc) {
          float v[16], w[16];
          tmem_ld16(tb + cc * 16, v);
          tmem_ld16(tb + F + cc * 16, w);
          tmem_ld_wait();
          if (w_bias != nullptr) {
            // bias of the predicted weight row: p[n, o*F + j] = D_o[n, j] + bl[o*F + j] (same address for every
            // thread: broadcast loads), so neither forward nor dL/dy needs a separate (N x F) x (F x F) product
            const float4* bp = reinterpret_cast<const float4*>(w_bias + (int64_t)o * F + grp * HF + cc * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b4 = __ldg(bp + q);
              w[4 * q] += b4.x, w[4 * q + 1] += b4.y, w[4 * q + 2] += b4.z, w[4 * q + 3] += b4.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (kMode == 0) acc = fmaf(v[j] + w[j], y[cc * 16 + j], acc);
            else y[cc * 16 + j] = fmaf(v[j] + w[j], sc, y[cc * 16 + j]);
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[b]);
        if (kMode == 0) {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q == oi) accs[q] = acc;  // static indexing keeps accs in registers
        }
      }
      if (kMode == 0) {
        float* yo = y_out + (int64_t)n * F + chunk * oc;
        if (grp == 1 && valid) {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q < oc) yo[q] = accs[q];
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");  // both epilogue groups; orders the global stores above
        if (grp == 0 && valid) {
          const float* e1 = e_term + (int64_t)n * F + chunk * oc;
          const float* e2 = e_term2 ? e_term2 + (int64_t)n * F + chunk * oc : nullptr;
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q < oc) yo[q] = (accs[q] + __ldcg(yo + q)) + (__ldg(e1 + q) + (e2 ? __ldg(e2 + q) : 0.f));
        }
      }
      if (kMode == 1 && valid && (item + 1 == item_hi || chunk == n_chunks - 1)) {
        const int slot = (int)blockIdx.x - hyper_cta_of_item((int64_t)tile * n_chunks, n_items, (int)gridDim.x);
        float4* dst = reinterpret_cast<float4*>(y_out + ((int64_t)slot * n_atoms + n) * F + grp * HF);
#pragma unroll
        for (int j = 0; j < HF / 4; ++j) dst[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        if (chunk == n_chunks - 1)
          for (int sl = slot + 1; sl < n_slots; ++sl) {
            float4* z4 = reinterpret_cast<float4*>(y_out + ((int64_t)sl * n_atoms + n) * F + grp * HF);
#pragma unroll
            for (int j = 0; j < HF / 4; ++j) z4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
      }
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ z-tile stagers + weight TMA
    const int st = tid - 256;  // 0..127
    uint32_t it = 0, cnt = 0;
    int staged_tile = -1;
    for (int item = item_lo; item < item_hi; ++item) {
      const int tile = item / n_chunks, chunk = item - tile * n_chunks;
      const bool restage = tile != staged_tile;
      staged_tile = tile;
      if (restage) mbar_wait(a_free, (it + 1) & 1u);  // the previous tile's MMAs have finished reading the z tile
#pragma unroll 1
      for (int kc = 0; restage && kc < Cfg::kKC; ++kc) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = st + 128 * j, r = idx >> 3, c = idx & 7;
          const int gr = tile * 128 + r;
          v[j] = gr < n_atoms ? __ldg(reinterpret_cast<const float4*>(z + (int64_t)gr * F + kc * 32 + c * 4))
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
      if (restage) {
        fence_async_smem();
        mbar_arrive(a_full);
        ++it;
      }
      if (st == 0) {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(w_packed);
        for (int oi = 0; oi < oc; ++oi) {
          const int o = chunk * oc + oi;
          for (int kc = 0; kc < Cfg::kKC; ++kc, ++cnt) {
            const uint32_t s = cnt % Cfg::kStages, u = cnt / Cfg::kStages;
            mbar_wait(&empty[s], (u + 1) & 1u);
            mbar_arrive_expect_tx(&full[s], kPackStageBytes);
            bulk_g2s(b_smem + s * kPackStageBytes, wsrc + ((int64_t)o * Cfg::kKC + kc) * kPackStageBytes,
                     kPackStageBytes, &full[s]);
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(128, F), idesc2 = umma_idesc_tf32(128, 2 * F);
    uint32_t it = 0, cnt = 0, ocount = 0;
    int staged_tile = -1;
    for (int item = item_lo; item < item_hi; ++item) {
      const int tile = item / n_chunks;
      if (tile != staged_tile) {
        mbar_wait(a_full, it & 1u);
        ++it;
        staged_tile = tile;
      }
      tc_fence_after();
      for (int oi = 0; oi < oc; ++oi, ++ocount) {
        const uint32_t b = ocount & 1u;
        mbar_wait(&tmem_empty[b], ((ocount >> 1) + 1) & 1u);
        tc_fence_after();
        for (int kc = 0; kc < Cfg::kKC; ++kc, ++cnt) {
          const uint32_t s = cnt % Cfg::kStages, u = cnt / Cfg::kStages;
          mbar_wait(&full[s], u & 1u);
          tc_fence_after();
          {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
            const uint32_t b_hi = smem_u32(b_smem + s * kPackStageBytes);
            const uint32_t d = tmem + b * 2 * F, dc = d + F;
            // The hi and lo images of a weight stage are adjacent, so ONE N = 2F MMA multiplies a_hi with both:
            // columns [d, d+F) receive a_hi*b_hi (main), [d+F, d+2F) a_hi*b_lo (correction); a second N = F MMA adds
            // a_lo*b_hi to the correction columns.  Two instructions and 20 KB of shared-memory operand reads per
            // K step instead of three and 24 KB — with M = N = 128 operand tiles the SS-mode MMAs run at the
            // shared-memory bandwidth limit, not at the tensor pipe's.
            const uint32_t a_hi = smem_u32(a_smem + kc * kPackStageBytes), a_lo = a_hi + kPackImageBytes;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t off = ks * 32;
              umma_tf32_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc2, (kc | ks) != 0);
              umma_tf32_e(dc, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, 1);
            }
            umma_commit_e(&empty[s]);
            if (kc == Cfg::kKC - 1) umma_commit_e(&tmem_full[b]);
          }
          __syncwarp();
        }
      }
      // last item of this atom tile: the z tile may be overwritten once these MMAs are done
      if ((item + 1 == item_hi || (item + 1) / n_chunks != tile)) umma_commit_e(a_free);
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::kTmemCols);
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

namespace {
template <int kMode>
int launch_hyper_impl(const float* z, const float* y_in, const float* e_term, const float* e_term2, const float* w_bias,
                      const float* w_packed, float* y_out, int64_t n_atoms, int32_t f, cudaStream_t stream) {
  using Cfg = HyperCfg<128>;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(hyper_rowdot_fwd_kernel<128, kMode>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int oc = hyper_chunk(n_atoms, f);
  const int grid = hyper_grid(n_atoms, f, kMode);
  hyper_rowdot_fwd_kernel<128, kMode><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(
      z, y_in, e_term, e_term2, w_bias, w_packed, y_out, (int)n_atoms, oc, hyper_parts(n_atoms, f));
  return check_launch(kMode == 0 ? "hyper_rowdot_fwd_kernel" : "hyper_rowscale_kernel");
}

template <int kMode>
int launch_hyper(const float* z, const float* y_in, const float* e_term, const float* e_term2, const float* w_bias,
                 const float* w_packed, float* y_out, int64_t n_atoms, int32_t f, cudaStream_t stream) {
  if (n_atoms <= 0) return 0;
  if (f != 128) return fail(-2, "cgat_hyper_*: only F = 128 is instantiated");
  if (n_atoms >= (1ll << 31) - 128) return fail(-2, "cgat_hyper_*: too many atoms");
  return launch_hyper_impl<kMode>(z, y_in, e_term, e_term2, w_bias, w_packed, y_out, n_atoms, f, stream);
}
}  // namespace

// y_out[n,o] = sum_i (sum_k z[n,k] W[o*F+i,k] + w_bias[o*F+i]) y_in[n,i] + e_term[n,o]   (w_bias optional)
//   z, y_in, e_term, e_term2 (optional, added to e_term), y_out: (n_atoms, F) fp32 contiguous;
//   w_packed: cgat_pack_kmajor of W[:F*F, :F]
extern "C" int cgat_hyper_rowdot_fwd(const float* z, const float* y_in, const float* e_term, const float* e_term2,
                                     const float* w_bias, const float* w_packed, float* y_out, int64_t n_atoms,
                                     int32_t f, void* stream_) {
  if (w_bias && (reinterpret_cast<uintptr_t>(w_bias) & 15)) return fail(-2, "cgat_hyper_rowdot_fwd: w_bias must be 16-byte aligned");
  return launch_hyper<0>(z, y_in, e_term, e_term2, w_bias, w_packed, y_out, n_atoms, f, (cudaStream_t)stream_);
}

// number of partial results cgat_hyper_rowscale writes for this problem size
extern "C" int32_t cgat_hyper_rowscale_parts(int64_t n_atoms, int32_t f) { return hyper_parts(n_atoms, f); }

// partial[c][n,j] = sum_{o in chunk c} scale[n,o] * (sum_m a[n,m] Wblk_o[j,m] + w_bias[o*F+j]);  sum over c = the
// result (w_bias optional: the bias of the predicted weights when a = z, NULL when a = y).
//   a, scale: (n_atoms, F); w_packed: cgat_pack_kmajor of the F blocks Wblk_o (F x F each, stacked);
//   partial: (cgat_hyper_rowscale_parts, n_atoms, F)
extern "C" int cgat_hyper_rowscale(const float* a, const float* scale, const float* w_bias, const float* w_packed,
                                   float* partial, int64_t n_atoms, int32_t f, void* stream_) {
  if (w_bias && (reinterpret_cast<uintptr_t>(w_bias) & 15)) return fail(-2, "cgat_hyper_rowscale: w_bias must be 16-byte aligned");
  return launch_hyper<1>(a, scale, nullptr, nullptr, w_bias, w_packed, partial, n_atoms, f, (cudaStream_t)stream_);
}
