// Fused edge-attention backward, steps 2 and 3 (SURVEY.md §8a row A12 for rows A2-A4).
// Step 1 (cgat_edge_attn_bwd_prep, edge_attn_fwd.cu) leaves per-edge dL/da, dL/dv (E,H,F) and the
// LeakyReLU sign masks.  Here:
//
//  cgat_edge_attn_dgrad   d_hid = dZ W2 (second-layer dgrad) on the tensor cores with HIDDEN UNITS on the
//      128 TMEM lanes and a tile of 128 edges on the columns, then in the epilogue
//          d_pre = d_hid * leaky_relu'(pre)         (sign masks: nothing is re-gathered)
//          G[seg, col] = sum over the segment's edges of d_pre      (thread-sequential, no atomics)
//      The kernel is run twice: once over edges grouped by DESTINATION (-> dL/dP dst-block) and once over
//      edges grouped by SOURCE (-> dL/dP src-block and, run-length accumulated, dL/dT per shell rank).
//      Two passes cost one extra dgrad GEMM but keep every reduction deterministic and conflict-free.
//
//  cgat_edge_attn_wgrad   dW2[net,h] = dZ^T hid: the contraction runs over EDGES, so both operands are
//      staged MN-major (dZ rows as they are, hid rows re-gathered exactly like the forward); split over
//      (net, head) x edge ranges, partial results summed by the caller.
#include "common.cuh"
#include "tc_common.cuh"

namespace cgat {
namespace {
using namespace tc;

constexpr int kBT = 128;  // edges per tile
constexpr int kBF = 128;  // channels per head (instantiated for F = 128)
constexpr int kBProducers = 256;
constexpr int kBThreads = 128 + kBProducers + 32;            // wgrad (tf32): 4 epilogue + 8 producer + 1 MMA warps
constexpr int kDEpilogue = 256;                                // dgrad: two epilogue groups of 4 warps, one per TMEM buffer
constexpr int kDThreads = kDEpilogue + kBProducers + 32;
constexpr int kDMmaWarp = (kDEpilogue + kBProducers) / 32;
constexpr int kBStages = 3;
constexpr int kBStageBytes = 2 * (int)kPackStageBytes;
constexpr int kBMetaBufs = 4;
constexpr int kBMetaStride = 3 * kBT + 8;  // seg | row | rank | 4 words of segment-start flags | 4 words of rank-run flags
constexpr int kBMetaBytes = kBMetaBufs * kBMetaStride * 4;
constexpr int kBMaxRanks = 32;
constexpr int kBRankBytes = kBMaxRanks * 128 * 4;
constexpr int kBSmemBytes = kBStages * kBStageBytes + kBMetaBytes + kBRankBytes + 256 + 1024;

struct DgradArgs {
  const float* d_gate;     // (E, H, F) rows in destination-sorted order
  const float* d_msg;
  const uint32_t* signs;   // [2][H][kcn][E]
  const int32_t* segptr;   // (N+1) CSR pointer of THIS edge order
  const int32_t* seg;      // (E) segment id of each edge of this order
  const int32_t* row;      // (E) destination-sorted row of each edge of this order, or null (identity)
  const int32_t* rnk;      // (E) shell rank of each edge of this order (only when d_rank != null)
  const float* wt_a;       // packed (H*Hd, F): row h*Hd+k, col c = W2A[h*F+c, k]
  const float* wt_m;
  float* G;                // (N, ldg); this pass writes columns [col_off, col_off + 2*H*Hd)
  float* d_rank;           // (grid, n_ranks, 2*H*Hd) partial dL/dT per CTA, or null
  float* d_pre;            // (E, 2*H*Hd) per-edge pre-activation gradients in THIS edge order, or null
  const float* dz_amax;    // kF16: device float max |d_gate|, |d_msg| (power-of-two range of the gradient operand)
  int64_t ldg;
  int col_off, n_atoms, n_edges, heads, hd, n_ranks;
  int f;                   // channels per head (128; the f16 d_pre form also 256)
};

__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int n, int64_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// kF16: both MMA operands as fp16 hi/lo pairs on kind::f16 (twice the tensor rate, half the W2^T stream from L2): W2^T
// packed by cgat_pack_kmajor_f16 (lo scaled by 2^11, separate correction accumulator as in hyper_f16.cu), the gradient
// operand dZ multiplied by a power of two s = 2^(4 - ceil(log2 amax)) while it is staged; the epilogue multiplies by 1/s.
template <bool kF16, bool kLean>
__global__ void __launch_bounds__(kDThreads, 1) edge_dgrad_kernel(const DgradArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  int32_t* meta = reinterpret_cast<int32_t*>(smem + kBStages * kBStageBytes);  // [4][seg|row|rank|flags]
  float* rank_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(meta) + kBMetaBytes);  // [n_ranks][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(rank_acc) + kBRankBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kBStages;
  uint64_t* tmem_full = bars + 2 * kBStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int32_t* range = reinterpret_cast<int32_t*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = g.heads, hd = g.hd, hhd = H * hd;
  const int kcn = (hd + 31) / 32;       // sign words per (net, head, edge)
  const int nhalf = (hd + 127) / 128;   // M tiles of 128 hidden units per head
  const int n_items = 2 * H * nhalf;
  constexpr int kChunkF = kF16 ? kPackChunk16 : kPackChunk;   // channels per pipeline stage
  const int kcf = g.f / kChunkF;        // K chunks of the dgrad contraction (over channels)
  float s_scale = 1.f, s_inv = 1.f;
  if (kF16) {
    const float amax = __ldg(g.dz_amax);
    int ex;
    frexpf(amax, &ex);
    if (amax > 0.f && amax < INFINITY) s_scale = ldexpf(1.f, 4 - ex), s_inv = ldexpf(1.f, ex - 4);
  }
  const float corr = kF16 ? kF16LoInv : 1.f;   // the correction accumulator of the f16 form carries lo * 2^11

  if (tid == 0) {
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(&full[s], kBProducers);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    mbar_init_fence();
    const int G_ = gridDim.x;
    const int64_t t0 = (int64_t)g.n_edges * blockIdx.x / G_, t1 = (int64_t)g.n_edges * (blockIdx.x + 1) / G_;
    const int a_lo = lower_bound_i32(g.segptr, g.n_atoms + 1, t0);
    const int a_hi = (blockIdx.x == G_ - 1) ? g.n_atoms : lower_bound_i32(g.segptr, g.n_atoms + 1, t1);
    range[0] = g.segptr[a_lo];
    range[1] = g.segptr[a_hi];
  }
  if (warp == kDMmaWarp) tmem_alloc(tmem_slot, 512);
  if (tid < 128)
    for (int r = 0; r < kBMaxRanks; ++r) rank_acc[r * 128 + tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int e_lo = range[0], e_hi = range[1];
  const int n_tiles = (e_hi - e_lo + kBT - 1) / kBT;

  if (warp < kDEpilogue / 32) {
    // ---------------------------------------------------------------- epilogue
    // Two groups of four warps (a warp reads the TMEM lanes 32*(warp%4)...): in the d_pre form group g owns TMEM
    // buffer g, i.e. every other (item, tile) — with the MMAs on kind::f16 four warps (one per scheduler, no latency
    // hiding) were what the kernel waited for.  The segment-sum form carries state across tiles: group 0 runs it alone.
    const int grp = warp >> 2;
    const int warp = (threadIdx.x >> 5) & 3;   // lane quadrant from here on
    // Thread = hidden unit (TMEM lane).  Per edge column: one FADD for hi+lo, a select for the LeakyReLU slope and
    // one FADD into the running segment sum; segment / rank-run boundaries come as precomputed bit masks so the
    // common path is branch-free (boundaries are ~1 in max_nbr columns).
    const int c = warp * 32 + lane;
    const uint32_t bitpos = (uint32_t)((lane & 3) * 8 + (lane >> 2));
    uint32_t icount = 0;
    if constexpr (kLean) {
      // Lean epilogue: d_pre[e, col] = d_hid * leaky_relu'(pre), one 128-byte line per warp and edge; every
      // segment / rank sum is taken from this copy by the HBM-bound cgat_edge_attn_reduce, so the per-column
      // work here is an add, a select, a multiply and a store.
      const int64_t ldd = 2 * (int64_t)hhd;
      for (int item = 0; item < n_items; ++item) {
        const int net = item / (H * nhalf), h = (item / nhalf) % H, half = item % nhalf;
        const int kk = half * 128 + c;
        const bool kvalid = kk < hd;
        const int col = net * hhd + h * hd + kk;
        const uint32_t* sg = g.signs + ((int64_t)(net * H + h) * kcn + (half * 4 + warp)) * g.n_edges;
        for (int tile = 0; tile < n_tiles; ++tile, ++icount) {
          const int e0 = e_lo + tile * kBT;
          const int nv = min(kBT, e_hi - e0);
          const uint32_t b = icount & 1u;
          if ((int)b != grp) continue;
          mbar_wait(&tmem_full[b], (icount >> 1) & 1u);
          tc_fence_after();
          const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + b * 256;
          float* prow = g.d_pre + (int64_t)e0 * ldd + col;
#pragma unroll 1
          for (int cc = 0; cc < kBT / 16; ++cc) {   // 16 columns per batch: 48 live values under the 120-register cap
            uint32_t wd[16];
            const uint4* sp = reinterpret_cast<const uint4*>(sg + e0 + cc * 16);
            const bool aligned = ((reinterpret_cast<uintptr_t>(sp) & 15) == 0) && (cc * 16 + 16 <= nv);
            if (aligned) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 q = kvalid ? __ldg(sp + j) : make_uint4(0, 0, 0, 0);
                wd[4 * j] = q.x, wd[4 * j + 1] = q.y, wd[4 * j + 2] = q.z, wd[4 * j + 3] = q.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) wd[j] = (kvalid && cc * 16 + j < nv) ? __ldg(sg + e0 + cc * 16 + j) : 0u;
            }
            float v[16], w[16];
            tmem_ld16(tbase + cc * 16, v);
            tmem_ld16(tbase + 128 + cc * 16, w);
            tmem_ld_wait();
            const int left = kvalid ? nv - cc * 16 : 0;  // columns of this group that are real edges
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float dp = fmaf(w[j], corr, v[j]) * (((wd[j] >> bitpos) & 1u) ? s_inv : 0.01f * s_inv);
              if (j < left) *prow = dp;
              prow += ldd;
            }
          }
          tc_fence_before();
          mbar_arrive(&tmem_empty[b]);
        }
      }
    } else if (grp == 0)   // kLean == false: the segment-sum form (compiled separately: it needs far more registers)
    for (int item = 0; item < n_items; ++item) {
      const int net = item / (H * nhalf), h = (item / nhalf) % H, half = item % nhalf;
      const int kk = half * 128 + c;
      const bool kvalid = kk < hd;
      const int col = net * hhd + h * hd + kk;
      const uint32_t* sg = g.signs + ((int64_t)(net * H + h) * kcn + (half * 4 + warp)) * g.n_edges;
      float* gcol = g.G + g.col_off + col;
      float acc = 0.f, racc = 0.f;
      int cur = -1, currk = -1;
      for (int tile = 0; tile < n_tiles; ++tile, ++icount) {
        const int e0 = e_lo + tile * kBT;
        const int nv = min(kBT, e_hi - e0);
        const int32_t* mt = meta + (icount & (kBMetaBufs - 1)) * kBMetaStride;
        const uint32_t b = icount & 1u;
        mbar_wait(&tmem_full[b], (icount >> 1) & 1u);
        tc_fence_after();
        // tile start: the only place where "same segment as before?" needs a compare (once per tile)
        {
          const int sid0 = mt[0];
          if (sid0 != cur) {
            if (cur >= 0 && kvalid) gcol[(int64_t)cur * g.ldg] = acc;
            acc = 0.f, cur = sid0;
          }
          const int rk0 = mt[2 * kBT];
          if (g.d_rank && rk0 != currk) {
            if (currk >= 0) rank_acc[currk * 128 + c] += racc;
            racc = 0.f, currk = rk0;
          }
        }
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + b * 256;
#pragma unroll 1
        for (int cc = 0; cc < kBT / 32; ++cc) {
          // LeakyReLU sides of these 32 edges: all loads issued before anything else
          uint32_t wd[32];
          if (g.row == nullptr) {
            const uint4* sp = reinterpret_cast<const uint4*>(sg + e0 + cc * 32);
            const bool aligned = ((reinterpret_cast<uintptr_t>(sp) & 15) == 0) && (cc * 32 + 32 <= nv);
            if (aligned) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 q = kvalid ? __ldg(sp + j) : make_uint4(0, 0, 0, 0);
                wd[4 * j] = q.x, wd[4 * j + 1] = q.y, wd[4 * j + 2] = q.z, wd[4 * j + 3] = q.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) wd[j] = (kvalid && cc * 32 + j < nv) ? __ldg(sg + e0 + cc * 32 + j) : 0u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) wd[j] = (kvalid && cc * 32 + j < nv) ? __ldg(sg + mt[kBT + cc * 32 + j]) : 0u;
          }
          // bit j: edge cc*32+j starts a new segment / rank run (never set for the first edge of the tile and for
          // the padding columns, whose accumulator entries are exactly 0)
          const uint32_t segflags = (uint32_t)mt[3 * kBT + cc];
          const uint32_t rnkflags = g.d_rank ? (uint32_t)mt[3 * kBT + 4 + cc] : 0u;
          float v[32], w[32];
          tmem_ld32(tbase + cc * 32, v);
          tmem_ld32(tbase + 128 + cc * 32, w);
          tmem_ld_wait();
          // branch-free common path: the flush is one predicated store to a pointer kept ready, the rest selects
          float* gp = gcol + (int64_t)cur * g.ldg;
          float* rp = rank_acc + max(currk, 0) * 128 + c;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const bool sf = ((segflags >> j) & 1u) && kvalid;
            if (sf) *gp = acc;
            acc = ((segflags >> j) & 1u) ? 0.f : acc;
            gp = gcol + (int64_t)mt[cc * 32 + j] * g.ldg;
            if (g.d_rank) {
              const bool rf = (rnkflags >> j) & 1u;
              if (rf) *rp += racc;
              racc = rf ? 0.f : racc;
              rp = rank_acc + mt[2 * kBT + cc * 32 + j] * 128 + c;
            }
            const float dp = (v[j] + w[j]) * (((wd[j] >> bitpos) & 1u) ? 1.f : 0.01f);
            acc += dp;
            racc += dp;

          }
          cur = mt[cc * 32 + 31];
          currk = mt[2 * kBT + cc * 32 + 31];
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[b]);
      }
      if (cur >= 0 && kvalid) gcol[(int64_t)cur * g.ldg] = acc;
      if (g.d_rank) {
        if (currk >= 0) rank_acc[currk * 128 + c] += racc;
        float* dst = g.d_rank + (int64_t)blockIdx.x * g.n_ranks * 2 * hhd + col;
        for (int r = 0; r < g.n_ranks; ++r) {
          if (kvalid) dst[(int64_t)r * 2 * hhd] = rank_acc[r * 128 + c];
          rank_acc[r * 128 + c] = 0.f;
        }
      }
    }
  } else if (warp < kDMmaWarp) {
    // ---------------------------------------------------------------- producers
    const int pt = tid - kDEpilogue;
    uint32_t cnt = 0, icount = 0;
    const uint8_t* wt[2] = {reinterpret_cast<const uint8_t*>(g.wt_a), reinterpret_cast<const uint8_t*>(g.wt_m)};
    const float* dz[2] = {g.d_gate, g.d_msg};
    for (int item = 0; item < n_items; ++item) {
      const int net = item / (H * nhalf), h = (item / nhalf) % H, half = item % nhalf;
      for (int tile = 0; tile < n_tiles; ++tile, ++icount) {
        const int e0 = e_lo + tile * kBT;
        const int nv = min(kBT, e_hi - e0);
        int32_t* mt = meta + (icount & (kBMetaBufs - 1)) * kBMetaStride;
        if (pt < kBT) {
          const bool ok = pt < nv;
          const int e = e0 + (ok ? pt : nv - 1);  // padding columns repeat the last valid edge's ids (their data is 0)
          const int sid = g.seg[e];
          const int rk = g.rnk ? g.rnk[e] : 0;
          mt[pt] = sid;
          mt[kBT + pt] = ok ? (g.row ? g.row[e] : e) : -1;
          mt[2 * kBT + pt] = rk;
          // boundary flags (bit t of word t/32): segment / rank differs from the edge before (never for t = 0)
          const bool sflag = ok && pt > 0 && g.seg[e - 1] != sid;
          const bool rflag = ok && pt > 0 && g.rnk && g.rnk[e - 1] != rk;
          const uint32_t sw = __ballot_sync(0xffffffffu, sflag), rw = __ballot_sync(0xffffffffu, rflag);
          if (lane == 0) mt[3 * kBT + (pt >> 5)] = (int32_t)sw, mt[3 * kBT + 4 + (pt >> 5)] = (int32_t)rw;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kBProducers) : "memory");
        int64_t rowoff[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = (pt + kBProducers * j) >> 3;
          rowoff[j] = (r < nv) ? ((int64_t)mt[kBT + r] * H + h) * g.f : -1;  // padding rows are staged as zeros
        }
        // the dZ rows of K chunk kc+1 are requested before this thread waits for the stage of chunk kc, so the
        // L2 / HBM latency of the loads overlaps the wait and the conversion (was: load -> wait for data -> convert)
        if constexpr (kF16) {
          // slot = (edge row, 16-byte chunk) = 8 consecutive channels: two float4 per slot, 4 slots per thread
          float4 xa[4], xb[4], na[4], nb[4];
          auto load = [&](int kc, float4* aa, float4* bb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int cch = (pt + kBProducers * j) & 7;
              aa[j] = bb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (kc < kcf && rowoff[j] >= 0) {
                const float4* p = reinterpret_cast<const float4*>(dz[net] + rowoff[j] + kc * kChunkF + cch * 8);
                aa[j] = __ldg(p), bb[j] = __ldg(p + 1);
              }
            }
          };
          load(0, xa, xb);
          for (int kc = 0; kc < kcf; ++kc, ++cnt) {
            const uint32_t s = cnt % kBStages, u = cnt / kBStages;
            load(kc + 1, na, nb);
            mbar_wait(&empty[s], (u + 1) & 1u);
            uint8_t* st = stages + s * kBStageBytes;
            if (pt == 0) {
              mbar_expect_tx(&full[s], kPackStageBytes);
              bulk_g2s(st, wt[net] + ((int64_t)(h * nhalf + half) * kcf + kc) * kPackStageBytes, kPackStageBytes, &full[s]);
            }
            uint8_t* bh = st + kPackStageBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int idx = pt + kBProducers * j;
              const float4 a = make_float4(xa[j].x * s_scale, xa[j].y * s_scale, xa[j].z * s_scale, xa[j].w * s_scale);
              const float4 b = make_float4(xb[j].x * s_scale, xb[j].y * s_scale, xb[j].z * s_scale, xb[j].w * s_scale);
              uint4 hi, lo;
              split_f16x8(a, b, hi, lo);
              const uint32_t off = sw128_offset(idx >> 3, idx & 7);
              *reinterpret_cast<uint4*>(bh + off) = hi;
              *reinterpret_cast<uint4*>(bh + kPackImageBytes + off) = lo;
            }
            fence_async_smem();
            mbar_arrive(&full[s]);
#pragma unroll
            for (int j = 0; j < 4; ++j) xa[j] = na[j], xb[j] = nb[j];
          }
        } else {
        float4 x[4], nx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cch = (pt + kBProducers * j) & 7;
          x[j] = rowoff[j] >= 0 ? __ldg(reinterpret_cast<const float4*>(dz[net] + rowoff[j] + cch * 4))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int kc = 0; kc < kcf; ++kc, ++cnt) {
          const uint32_t s = cnt % kBStages, u = cnt / kBStages;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cch = (pt + kBProducers * j) & 7;
            nx[j] = (kc + 1 < kcf && rowoff[j] >= 0)
                        ? __ldg(reinterpret_cast<const float4*>(dz[net] + rowoff[j] + (kc + 1) * 32 + cch * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          mbar_wait(&empty[s], (u + 1) & 1u);
          uint8_t* st = stages + s * kBStageBytes;
          if (pt == 0) {
            mbar_expect_tx(&full[s], kPackStageBytes);
            bulk_g2s(st, wt[net] + ((int64_t)(h * nhalf + half) * kcf + kc) * kPackStageBytes, kPackStageBytes, &full[s]);
          }
          uint8_t* bh = st + kPackStageBytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int idx = pt + kBProducers * j;
            float4 hi, lo;
            split_tf32(x[j], hi, lo);
            const uint32_t off = sw128_offset(idx >> 3, idx & 7);
            *reinterpret_cast<float4*>(bh + off) = hi;
            *reinterpret_cast<float4*>(bh + kPackImageBytes + off) = lo;
          }
          fence_async_smem();
          mbar_arrive(&full[s]);
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = nx[j];
        }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = kF16 ? umma_idesc_f16(128, kBT) : umma_idesc_tf32(128, kBT);
    uint32_t cnt = 0, icount = 0;
    for (int item = 0; item < n_items; ++item) {
      for (int tile = 0; tile < n_tiles; ++tile, ++icount) {
        const uint32_t b = icount & 1u;
        mbar_wait(&tmem_empty[b], ((icount >> 1) + 1) & 1u);
        tc_fence_after();
        const uint32_t d = tmem + b * 256, dc = d + 128;
        for (int kc = 0; kc < kcf; ++kc, ++cnt) {
          const uint32_t s = cnt % kBStages, u = cnt / kBStages;
          mbar_wait(&full[s], u & 1u);
          tc_fence_after();
          {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
            const uint32_t a_hi = smem_u32(stages + s * kBStageBytes), a_lo = a_hi + kPackImageBytes;
            const uint32_t b_hi = a_hi + kPackStageBytes, b_lo = b_hi + kPackImageBytes;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t off = ks * 32;
              if constexpr (kF16) {
                umma_f16_e(dc, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, (kc | ks) != 0);
                umma_f16_e(dc, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc, 1);
                umma_f16_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc, (kc | ks) != 0);
              } else {
                umma_tf32_e(dc, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, (kc | ks) != 0);
                umma_tf32_e(dc, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc, 1);
                umma_tf32_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc, (kc | ks) != 0);
              }
            }
            umma_commit_e(&empty[s]);
            if (kc == kcf - 1) umma_commit_e(&tmem_full[b]);
          }
          __syncwarp();
        }
      }
    }
  }
  __syncthreads();
  if (warp == kDMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* P;
  const float* T;
  const int32_t* src;
  const int32_t* dst;
  const int32_t* rank;
  const float* d_gate;
  const float* d_msg;
  float* out;  // (n_split, 2, H, F, Hd) partial dW2
  int n_edges, heads, hd, n_split;
};

constexpr int kWImage = 32 * 128;          // 32 K-rows (edges) x 128 B
constexpr int kWAPart = 4 * kWImage;       // dZ: 128 channels = 4 images
constexpr int kWBPart = 8 * kWImage;       // hid: up to 256 hidden units = 8 images
constexpr int kWStageBytes = 2 * kWAPart + 2 * kWBPart;  // 96 KB
constexpr int kWStages = 2;
constexpr int kWSmemBytes = kWStages * kWStageBytes + 2 * 3 * 32 * 4 + 256 + 1024;

// max(x, 0.01 x): the same value as (x > 0 ? x : 0.01 x) for every x (signed zeros, infinities and NaN included), one
// instruction less per hidden element in producers that are bound by their instruction stream
__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, 0.01f * x); }

__global__ void __launch_bounds__(kBThreads, 1) edge_wgrad_kernel(const WgradArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWStages * kWStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWStages;
  uint64_t* accum = bars + 2 * kWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  int32_t* meta = reinterpret_cast<int32_t*>(tmem_slot + 2);  // [2 parity][3][32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = g.heads, hd = g.hd, hhd = H * hd;
  const int item = blockIdx.x / g.n_split, split = blockIdx.x % g.n_split;
  const int net = item / H, h = item % H;
  const int e_lo = (int)((int64_t)g.n_edges * split / g.n_split), e_hi = (int)((int64_t)g.n_edges * (split + 1) / g.n_split);
  const int n_chunks = (e_hi - e_lo + 31) / 32;
  const int64_t ldp = 4 * (int64_t)hhd, ldt = 2 * (int64_t)hhd;

  if (tid == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(&full[s], kBProducers);
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
    // epilogue: one accumulation over the whole edge range, written once
    const int c = warp * 32 + lane;
    mbar_wait(accum, 0);
    tc_fence_after();
    float* dst = g.out + ((((int64_t)split * 2 + net) * H + h) * kBF + c) * hd;
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
      float v[32], w[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + cc * 32, w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cc * 32 + j < hd) dst[cc * 32 + j] = n_chunks > 0 ? v[j] + w[j] : 0.f;
    }
    tc_fence_before();
  } else if (warp < 12) {
    const int pt = tid - 128;
    const float* dz = net ? g.d_msg : g.d_gate;
    // The edge ids of chunk ch+1 are fetched while chunk ch is being staged (they used to cost one exposed L2 round
    // trip per chunk, in front of the gathers that depend on them); threads 0-31 carry them in registers.
    int nd = -1, ns = 0, nr = 0;
    if (pt < 32 && n_chunks > 0) {
      const bool ok = pt < min(32, e_hi - e_lo);
      nd = ok ? g.dst[e_lo + pt] : -1, ns = ok ? g.src[e_lo + pt] : 0, nr = ok ? g.rank[e_lo + pt] : 0;
    }
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int s = ch % kWStages, u = ch / kWStages;
      const int e0 = e_lo + ch * 32;
      const int nv = min(32, e_hi - e0);
      int32_t* mt = meta + (ch & 1) * 96;
      if (pt < 32) {
        mt[pt] = nd, mt[32 + pt] = ns, mt[64 + pt] = nr;
        const int e1 = e0 + 32;
        const bool ok = ch + 1 < n_chunks && pt < min(32, e_hi - e1);
        nd = ok ? g.dst[e1 + pt] : -1, ns = ok ? g.src[e1 + pt] : 0, nr = ok ? g.rank[e1 + pt] : 0;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kBProducers) : "memory");
      mbar_wait(&empty[s], (u + 1) & 1u);
      uint8_t* st = stages + s * kWStageBytes;
      // A operand: dZ rows (32 edges x 128 channels), MN-major images of 32 channels
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = pt + kBProducers * j, r = idx >> 5, q = idx & 31;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nv) x = __ldg(reinterpret_cast<const float4*>(dz + ((int64_t)(e0 + r) * H + h) * kBF + q * 4));
        float4 hi, lo;
        split_tf32(x, hi, lo);
        const uint32_t off = (q >> 3) * kWImage + mn_sw128_offset(r, q & 7);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + kWAPart + off) = lo;
      }
      // B operand: hidden activations (32 edges x Hd), re-gathered like the forward
      uint8_t* bh = st + 2 * kWAPart;
      // all 8 slots unrolled: up to 24 gathered float4 in flight per thread (13 warps per CTA leave 128 registers)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = pt + kBProducers * j, r = idx >> 6, q = idx & 63;
        const int d = mt[r];
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d >= 0 && q * 4 < hd) {
          const int col = net * hhd + h * hd + q * 4;
          const float4 pd = __ldg(reinterpret_cast<const float4*>(g.P + (int64_t)d * ldp + col));
          const float4 ps = __ldg(reinterpret_cast<const float4*>(g.P + (int64_t)mt[32 + r] * ldp + 2 * hhd + col));
          const float4 te = __ldg(reinterpret_cast<const float4*>(g.T + (int64_t)mt[64 + r] * ldt + col));
          x.x = lrelu(pd.x + ps.x + te.x), x.y = lrelu(pd.y + ps.y + te.y);
          x.z = lrelu(pd.z + ps.z + te.z), x.w = lrelu(pd.w + ps.w + te.w);
        }
        float4 hi, lo;
        split_tf32(x, hi, lo);
        const uint32_t off = (q >> 3) * kWImage + mn_sw128_offset(r, q & 7);
        *reinterpret_cast<float4*>(bh + off) = hi;
        *reinterpret_cast<float4*>(bh + kWBPart + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(&full[s]);
    }
  } else {
    const uint32_t idesc = umma_idesc_tf32(128, 256, 1, 1);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int s = ch % kWStages, u = ch / kWStages;
      mbar_wait(&full[s], u & 1u);
      tc_fence_after();
      {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
        const uint32_t a_hi = smem_u32(stages + s * kWStageBytes), a_lo = a_hi + kWAPart;
        const uint32_t b_hi = a_hi + 2 * kWAPart, b_lo = b_hi + kWBPart;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t o = ks * 1024;
          umma_tf32_e(tmem + 256, umma_desc_mn_sw128(a_lo + o, kWImage), umma_desc_mn_sw128(b_hi + o, kWImage), idesc,
                    (ch | ks) != 0);
          umma_tf32_e(tmem + 256, umma_desc_mn_sw128(a_hi + o, kWImage), umma_desc_mn_sw128(b_lo + o, kWImage), idesc, 1);
          umma_tf32_e(tmem, umma_desc_mn_sw128(a_hi + o, kWImage), umma_desc_mn_sw128(b_hi + o, kWImage), idesc,
                    (ch | ks) != 0);
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

// ------------------------------------------------------------------------------------------------
// Segment sums of the per-edge pre-activation gradients (HBM-bound; replaces the thread-sequential segment
// sums of the dgrad epilogue and a whole second dgrad pass over source-grouped edges):
//     pass 0  G[a, dst_off + c] = sum over the in-edges of atom a   (rows of d_pre are destination-sorted)
//     pass 1  G[s, src_off + c] = sum over the out-edges t of atom s of d_pre[row_t, c]
//             d_rank[chunk, r, c] = sum over the chunk's edges with shell rank r
// One CTA = (pass, tile of 512 columns, contiguous chunk of atoms); thread = one float4 column group, sequential
// over its atoms and edges, so every sum has a fixed order (deterministic, no atomics).  Ranks are non-decreasing
// along an atom's out-edges, so the rank sums are run-length accumulated in registers and flushed to shared memory.
constexpr int kSRThreads = 128;
constexpr int kSRCols = kSRThreads * 4;

struct ReducePass {
  const int32_t* rowptr;   // (N+1) CSR of this edge order
  const int32_t* row;      // (E) row of d_pre of each edge of this order, or null (identity)
  const int32_t* rnk;      // (E) shell rank, or null (no per-rank sums)
  float* d_rank;           // (n_chunks, n_ranks, cols) or null
  int col_off;
};
struct ReduceArgs {
  ReducePass pass[2];
  const float* d_pre;
  float* G;
  int64_t ldd, ldg;
  int n_atoms, n_ranks, cols, n_chunks;
  float* rank_out;         // (n_groups, n_ranks, cols): per-rank sums of kSRGroup consecutive chunks
  int32_t* counters;       // (n_groups, column blocks), zeroed before the launch
};
constexpr int kSRGroup = 8;   // chunks whose per-rank partial sums are combined by the last CTA of the group to finish

__global__ void __launch_bounds__(kSRThreads) edge_reduce_kernel(const ReduceArgs g) {
  extern __shared__ float4 sr_acc[];  // [n_ranks][kSRThreads]
  __shared__ int s_last;
  // pass = lowest bit of blockIdx.y: the by-destination and the by-source CTA of one atom chunk are scheduled side by
  // side, and the out-edges of an atom end in its own crystal, i.e. in rows the neighbour CTA reads at about the same
  // time — part of the second read of d_pre (683 MB at the bench size, five times the L2) is then served by L2
  // (365 -> 330 us).  Measured and dropped: one CTA per chunk that stages the chunk's rows in shared memory once and
  // takes both sums from there — with 64- or 32-column blocks (all that fits next to the rows of 20-40 atoms) the
  // 128-256-byte row pieces cost more in DRAM efficiency than the second read saves (540-615 us).
  const ReducePass& ps = g.pass[blockIdx.y & 1u];
  const int tid = threadIdx.x;
  const int col = blockIdx.x * kSRCols + tid * 4;
  const bool active = col < g.cols;
  const int chunk = blockIdx.y >> 1;
  const int a_lo = (int)((int64_t)g.n_atoms * chunk / g.n_chunks);
  const int a_hi = (int)((int64_t)g.n_atoms * (chunk + 1) / g.n_chunks);
  const bool ranks = ps.rnk != nullptr && ps.d_rank != nullptr;
  if (ranks)
    for (int r = 0; r < g.n_ranks; ++r) sr_acc[r * kSRThreads + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* src = g.d_pre + col;
  float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur_rank = 0;
  for (int a = a_lo; a < a_hi; ++a) {
    const int b = __ldg(ps.rowptr + a), e = __ldg(ps.rowptr + a + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t0 = b; t0 < e; t0 += 8) {   // (12 / 16 rows in flight per thread: 275 / 327 us against 267 — occupancy)
      float4 v[8];
      int rk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        rk[j] = -1;
        if (t0 + j < e) {
          rk[j] = ranks ? __ldg(ps.rnk + t0 + j) : 0;
          const int64_t r = ps.row ? __ldg(ps.row + t0 + j) : t0 + j;
          if (active) v[j] = __ldg(reinterpret_cast<const float4*>(src + r * g.ldd));
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (rk[j] < 0) continue;  // uniform across the CTA
        acc.x += v[j].x, acc.y += v[j].y, acc.z += v[j].z, acc.w += v[j].w;
        if (ranks) {
          if (rk[j] != cur_rank) {  // uniform: rank runs are a property of the edge list
            float4 s = sr_acc[cur_rank * kSRThreads + tid];
            s.x += run.x, s.y += run.y, s.z += run.z, s.w += run.w;
            sr_acc[cur_rank * kSRThreads + tid] = s;
            run = make_float4(0.f, 0.f, 0.f, 0.f);
            cur_rank = rk[j];
          }
          run.x += v[j].x, run.y += v[j].y, run.z += v[j].z, run.w += v[j].w;
        }
      }
    }
    if (active) *reinterpret_cast<float4*>(g.G + (int64_t)a * g.ldg + ps.col_off + col) = acc;
  }
  if (ranks) {
    float4 s = sr_acc[cur_rank * kSRThreads + tid];
    s.x += run.x, s.y += run.y, s.z += run.z, s.w += run.w;
    sr_acc[cur_rank * kSRThreads + tid] = s;
    if (active)
      for (int r = 0; r < g.n_ranks; ++r)
        *reinterpret_cast<float4*>(ps.d_rank + ((int64_t)chunk * g.n_ranks + r) * g.cols + col) =
            sr_acc[r * kSRThreads + tid];
    // The per-chunk partials (133 KB each at the default width) used to go to the caller as they were — 31 MB written
    // and read again by the partial sum, and the reason smaller (faster) chunks did not pay (profiles/r04y).  Now the
    // CTA that finishes LAST among the kSRGroup chunks of its group (same column block) adds the group's partials in
    // chunk order — a fixed order whoever comes last, so still deterministic — while they are still in L2.
    const int group = chunk / kSRGroup;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int first = group * kSRGroup, members = min(kSRGroup, g.n_chunks - first);
      s_last = atomicAdd(g.counters + (int64_t)group * gridDim.x + blockIdx.x, 1) == members - 1;
    }
    __syncthreads();
    if (s_last && active) {
      __threadfence();
      const int first = group * kSRGroup, members = min(kSRGroup, g.n_chunks - first);
      for (int r = 0; r < g.n_ranks; ++r) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int m = 0; m < members; ++m) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(ps.d_rank + ((int64_t)(first + m) * g.n_ranks + r) * g.cols + col));
          t.x += v.x, t.y += v.y, t.z += v.z, t.w += v.w;
        }
        *reinterpret_cast<float4*>(g.rank_out + ((int64_t)group * g.n_ranks + r) * g.cols + col) = t;
      }
    }
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

extern "C" int32_t cgat_edge_attn_reduce_chunks(int64_t n_atoms) {
  // ~24 atoms per CTA: every CTA walks its atoms one after the other (index load -> row loads -> store), so the
  // memory-level parallelism comes from the number of resident CTAs (8+ per SM at 27 KB of shared memory each)
  // atoms per chunk, swept (profiles/r04y): 12 / 24 / 48 / 96 -> 303 / 330 / 391 / 448 us for the reduce itself.  With
  // one rank partial per chunk going back to the caller, 12 lost what it gained in the partial sum that follows; since
  // the partials are combined per group of 8 chunks inside the kernel, the smaller chunk wins.
#ifndef CGAT_SR_ATOMS
#define CGAT_SR_ATOMS 12
#endif
#ifndef CGAT_SR_MAXCHUNKS
#define CGAT_SR_MAXCHUNKS 4096
#endif
  int64_t c = (n_atoms + CGAT_SR_ATOMS - 1) / CGAT_SR_ATOMS;
  return (int32_t)(c < 1 ? 1 : (c > CGAT_SR_MAXCHUNKS ? CGAT_SR_MAXCHUNKS : c));
}
// groups of chunks = parts of the d_rank result the caller sums; ints of the counter workspace
extern "C" int32_t cgat_edge_attn_reduce_groups(int64_t n_atoms) {
  return (cgat_edge_attn_reduce_chunks(n_atoms) + kSRGroup - 1) / kSRGroup;
}
extern "C" int32_t cgat_edge_attn_reduce_counters(int64_t n_atoms, int32_t cols) {
  return cgat_edge_attn_reduce_groups(n_atoms) * (int32_t)ceil_div(cols, kSRCols);
}

// Segment sums of the per-edge pre-activation gradients d_pre (E, cols; rows in destination order) that
// cgat_edge_attn_dgrad wrote:  G[a, dst_col_off : +cols] = sum over the in-edges of atom a (contiguous rows
// [dst_rowptr[a], dst_rowptr[a+1])),  G[a, src_col_off : +cols] = sum over its out-edges (rows src_row[t] for t in
// [src_rowptr[a], src_rowptr[a+1])),  d_rank (cgat_edge_attn_reduce_chunks(N), n_ranks, cols) = per-chunk sums per
// shell rank src_rank[t] (sum over dim 0 = dL/dT).  HBM-bound: reads d_pre twice, deterministic, no atomics.
extern "C" int cgat_edge_attn_reduce(const float* d_pre, int64_t ldd, const int32_t* dst_rowptr,
                                     const int32_t* src_rowptr, const int32_t* src_row, const int32_t* src_rank,
                                     float* G, int64_t ldg, int32_t dst_col_off, int32_t src_col_off, float* d_rank,
                                     float* rank_scratch, int32_t* counters, int32_t n_ranks, int64_t n_atoms,
                                     int32_t cols, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if ((cols & 3) || (ldd & 3) || (ldg & 3) || (dst_col_off & 3) || (src_col_off & 3))
    return fail(-2, "cgat_edge_attn_reduce: cols, ldd, ldg and the column offsets must be multiples of 4");
  if (n_ranks < 1 || n_ranks > kBMaxRanks) return fail(-2, "cgat_edge_attn_reduce: n_ranks must be in [1,32]");
  if (n_atoms <= 0 || cols <= 0) return 0;
  const int n_chunks = cgat_edge_attn_reduce_chunks(n_atoms);
  ReduceArgs a{};
  a.pass[0] = ReducePass{dst_rowptr, nullptr, nullptr, nullptr, dst_col_off};
  if (!d_rank || !rank_scratch || !counters || !src_rank)
    return fail(-2, "cgat_edge_attn_reduce: d_rank, rank_scratch, counters and src_rank are required");
  a.pass[1] = ReducePass{src_rowptr, src_row, src_rank, rank_scratch, src_col_off};
  a.rank_out = d_rank, a.counters = counters;
  CGAT_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * cgat_edge_attn_reduce_counters(n_atoms, cols), stream));
  a.d_pre = d_pre, a.G = G, a.ldd = ldd, a.ldg = ldg;
  a.n_atoms = (int)n_atoms, a.n_ranks = n_ranks, a.cols = cols, a.n_chunks = n_chunks;
  const size_t smem = (size_t)n_ranks * kSRThreads * sizeof(float4);
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(edge_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kBMaxRanks * kSRThreads * (int)sizeof(float4)));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(cols, kSRCols), (unsigned)(2 * n_chunks), 1);
  edge_reduce_kernel<<<grid, kSRThreads, smem, stream>>>(a);
  return check_launch("edge_reduce_kernel");
}

extern "C" int32_t cgat_edge_attn_dgrad_grid(int64_t n_edges) {
  const int64_t tiles = ceil_div(n_edges, kBT);
  return (int32_t)(tiles < kNumSMs ? (tiles > 0 ? tiles : 1) : kNumSMs);
}

// One pass of the dgrad + segment reduction.  `segptr/seg/row/rnk` describe the edge order of this pass
// (grouped by destination or by source).  Writes G[:, col_off : col_off + 2*H*Hd]; rows of atoms without
// edges in this order are left untouched (zero them beforehand).  d_rank (optional):
// (cgat_edge_attn_dgrad_grid(E), n_ranks, 2*H*Hd) per-CTA partial sums per shell rank.  d_pre (optional, identity
// order only): (E, 2*H*Hd) the per-edge pre-activation gradients themselves.
extern "C" int cgat_edge_attn_dgrad(const float* d_gate, const float* d_msg, const uint32_t* signs,
                                    const int32_t* segptr, const int32_t* seg, const int32_t* row, const int32_t* rnk,
                                    const float* wt_a_packed, const float* wt_m_packed, float* G, int64_t ldg,
                                    int32_t col_off, float* d_rank, int32_t n_ranks, float* d_pre, int64_t n_atoms,
                                    int64_t n_edges, int32_t heads, int32_t f, int32_t hd, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (f != kBF) return fail(-2, "cgat_edge_attn_dgrad: only F = 128 is instantiated");
  if (heads < 1 || heads > 8 || hd <= 0 || (hd & 127))
    return fail(-2, "cgat_edge_attn_dgrad: heads must be in [1,8] and the hidden width a multiple of 128");
  if (d_rank && (n_ranks < 1 || n_ranks > kBMaxRanks)) return fail(-2, "cgat_edge_attn_dgrad: n_ranks must be <= 32");
  if (n_atoms <= 0 || n_edges <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(edge_dgrad_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmemBytes));
    CGAT_CUDA(cudaFuncSetAttribute(edge_dgrad_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmemBytes));
    configured = true;
  }
  if (d_pre && row) return fail(-2, "cgat_edge_attn_dgrad: d_pre needs the identity edge order (row == NULL)");
  DgradArgs a{d_gate, d_msg, signs, segptr, seg, row, rnk, wt_a_packed, wt_m_packed, G, d_rank, d_pre, nullptr, ldg,
              col_off, (int)n_atoms, (int)n_edges, heads, hd, n_ranks, kBF};
  if (d_pre) edge_dgrad_kernel<false, true><<<cgat_edge_attn_dgrad_grid(n_edges), kDThreads, kBSmemBytes, stream>>>(a);
  else edge_dgrad_kernel<false, false><<<cgat_edge_attn_dgrad_grid(n_edges), kDThreads, kBSmemBytes, stream>>>(a);
  return check_launch("edge_dgrad_kernel");
}

// kind::f16 form of the d_pre variant (identity edge order): wt_*_packed = cgat_pack_kmajor_f16 of W2^T per head,
// dz_amax = the device float cgat_edge_attn_bwd_prep left (max |d_gate|, |d_msg|).  Writes d_pre (E, 2*H*Hd).
// F = 128: the rebuilt d_pre kernel of edge_dgrad_f16.cu (dZ tile converted once for all hidden halves)
int cgat_edge_dgrad_zr_launch(const float* d_gate, const float* d_msg, const uint32_t* signs, const int32_t* segptr,
                              const float* wt_a_packed, const float* wt_m_packed, const float* dz_amax, float* d_pre,
                              int64_t n_atoms, int64_t n_edges, int32_t heads, int32_t hd, int32_t grid,
                              cudaStream_t stream);

extern "C" int cgat_edge_attn_dgrad_f16(const float* d_gate, const float* d_msg, const uint32_t* signs,
                                        const int32_t* segptr, const int32_t* seg, const float* wt_a_packed,
                                        const float* wt_m_packed, const float* dz_amax, float* d_pre, int64_t n_atoms,
                                        int64_t n_edges, int32_t heads, int32_t f, int32_t hd, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (f != kBF && f != 2 * kBF) return fail(-2, "cgat_edge_attn_dgrad_f16: F must be 128 or 256");
  if (heads < 1 || heads > 8 || hd <= 0 || (hd & 31))
    return fail(-2, "cgat_edge_attn_dgrad_f16: heads must be in [1,8] and the hidden width a multiple of 32 (W2^T packed "
                    "with ceil(hd / 128) * 128 rows per head)");
  if (!d_pre || !dz_amax) return fail(-2, "cgat_edge_attn_dgrad_f16: d_pre and dz_amax are required");
  if (n_atoms <= 0 || n_edges <= 0) return 0;
  if (f == kBF)
    return cgat_edge_dgrad_zr_launch(d_gate, d_msg, signs, segptr, wt_a_packed, wt_m_packed, dz_amax, d_pre, n_atoms,
                                     n_edges, heads, hd, cgat_edge_attn_dgrad_grid(n_edges), stream);
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(edge_dgrad_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmemBytes));
    configured = true;
  }
  DgradArgs a{d_gate, d_msg, signs, segptr, seg, nullptr, nullptr, wt_a_packed, wt_m_packed, nullptr, nullptr, d_pre,
              dz_amax, 0, 0, (int)n_atoms, (int)n_edges, heads, hd, 1, f};
  edge_dgrad_kernel<true, true><<<cgat_edge_attn_dgrad_grid(n_edges), kDThreads, kBSmemBytes, stream>>>(a);
  return check_launch("edge_dgrad_f16_kernel");
}

extern "C" int32_t cgat_edge_attn_wgrad_splits(int32_t heads) {
  int s = kNumSMs / (2 * heads);
  return s < 1 ? 1 : s;
}

// out: (cgat_edge_attn_wgrad_splits(H), 2, H, F, Hd) partial dL/dW2 (gate net, message net); sum over dim 0.
// (Passing the hidden activations from the backward-prep kernel instead of re-gathering them was measured: slower,
// the re-gather hits P in L2 while a saved (E, 2*H*Hd) copy comes back from HBM.)
extern "C" int cgat_edge_attn_wgrad(const float* P, const float* T, const int32_t* src, const int32_t* dst,
                                    const int32_t* rank, const float* d_gate, const float* d_msg, float* out,
                                    int64_t n_edges, int32_t heads, int32_t f, int32_t hd, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (f != kBF) return fail(-2, "cgat_edge_attn_wgrad: only F = 128 is instantiated");
  if (heads < 1 || heads > 8 || hd <= 0 || (hd & 15) || hd > 256)
    return fail(-2, "cgat_edge_attn_wgrad: hidden width must be a multiple of 16, at most 256");
  if (n_edges <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(edge_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBytes));
    configured = true;
  }
  const int n_split = cgat_edge_attn_wgrad_splits(heads);
  WgradArgs a{P, T, src, dst, rank, d_gate, d_msg, out, (int)n_edges, heads, hd, n_split};
  edge_wgrad_kernel<<<2 * heads * n_split, kBThreads, kWSmemBytes, stream>>>(a);
  return check_launch("edge_wgrad_kernel");
}
