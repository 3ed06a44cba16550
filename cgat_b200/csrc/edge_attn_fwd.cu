// Fused edge-attention forward (kernels K1+K2+K3 of SURVEY.md §2.3).
//
// Reference arithmetic per edge t (source s, destination d, shell rank r), reference
// CGAT/CGAT.py:319-329 (GATConvNodes.message / aggregate / update head-mean) and :103-109
// (MultiHeadNetwork), Appendix A of SURVEY.md:
//     m_t    = [x[d]; e(r); x[s]]
//     hid    = leaky_relu(W1 m_t + b1, 0.01)                     per net (gate A, message M), per head
//     a_t,h  = W2A_h hidA_t,h + b2A_h ;  v_t,h = W2M_h hidM_t,h + b2M_h
//     alpha  = softmax over the in-edges of d, per (head, channel), eps 1e-16
//     out[d,h,:] = sum_t alpha_t,h * v_t,h
// W1 m_t is linear in its three blocks, so it arrives pre-evaluated per ATOM and per RANK:
//     pre_t = P[d, dst-block] + P[s, src-block] + T[r]           (P = x W1_{i|j}^T, T = e W1_e^T + b1)
// This kernel gathers those rows (K1), applies LeakyReLU and the tf32 hi/lo split while staging
// them as the B operand, runs the second MLP layer on the tensor cores with CHANNELS on the 128 TMEM
// lanes and the tile's 128 destination-sorted EDGES on the columns (K2), and lets the thread that
// owns a channel walk its row of the accumulator with an online segmented softmax (K3): no atomics,
// no per-edge intermediate ever reaches HBM, bit-reproducible.
//
// One persistent CTA per SM owns a contiguous range of destination atoms (hence whole softmax
// segments), balanced by edge count.  Roles (800 threads):
//   warps 0-7   epilogue, two groups of 4 warps (a warp reads the TMEM lanes 32*(warp%4)...): group g owns TMEM
//               buffer g, i.e. every other (tile, head) item — tcgen05.ld gate/message rows, segmented softmax,
//               write out[d,h,:].  The per-head state carried across tiles passes between the groups through
//               shared memory + one mbarrier per head.
//   warps 8-23  producers: gather + LeakyReLU + split -> shared memory (B operand); the first of them also
//               issues the cp.async.bulk of the pre-packed W2 chunk (A operand); 3-stage full/empty mbarrier ring
//   warp  24    TMEM allocation (512 columns: 2 buffers x {gate, message} x 128) + single-thread MMA issue
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

#ifndef CGAT_EFWD_DBG
#define CGAT_EFWD_DBG 0   // timing experiments: 1 no MMAs, 2 no gathers / conversion, 4 no W2 stream, 8 no epilogue math
#endif
namespace cgat {
namespace {
using namespace tc;

constexpr int kET = 128;             // edges per tile (MMA N)
constexpr int kEF = 128;             // output channels per head (MMA M) — instantiated for F = 128
constexpr int kEProducers = 512;     // producer threads; kEProducers / kEGroup groups alternate pipeline stages
constexpr int kEGroup = 256;
constexpr int kEEpilogue = 256;      // two groups of 128 epilogue threads
constexpr int kEMmaWarp = (kEEpilogue + kEProducers) / 32;
constexpr int kEThreads = kEEpilogue + kEProducers + 32;
constexpr int kEStages = 3;
constexpr int kEStageBytes = 2 * (int)kPackStageBytes;  // [W2 chunk hi|lo][hidden chunk hi|lo]
constexpr int kEMaxHeads = 16;         // VIRTUAL heads: (real head, 128-channel block of its F output channels)
constexpr int kEMetaBufs = 4;          // tile metadata ring (producers run ahead of the epilogue)
constexpr int kEMetaStride = 4 * kET + 8;  // dst | src | rank | dst*H*F | 4 words segment-start flags | 4 words valid mask
constexpr int kEMetaBytes = kEMetaBufs * kEMetaStride * 4;
constexpr int kECarryBytes = kEMaxHeads * 3 * kEF * 4 + kEMaxHeads * 4;  // (max, den, acc) per head and channel + open dst per head
constexpr int kESmemBytes = kEStages * kEStageBytes + kEMetaBytes + kECarryBytes + 256 + 1024;

struct EdgeArgs {
  const float* P;       // (N, 4*HHd): [gate dst | msg dst | gate src | msg src]
  const float* T;       // (K+1, 2*HHd): [gate | msg], first-layer bias included
  const int32_t* rowptr;
  const int32_t* src;
  const int32_t* dst;
  const int32_t* rank;
  const float* w2a;     // packed (H*F, Hd)
  const float* w2m;
  const float* b2a;     // (H*F)
  const float* b2m;
  float* out;           // (N, H, F)   forward: written; backward-prep: read
  float* smax;          // (N, H, F)   forward: written if non-null; backward-prep: read
  float* sden;
  const float* g_out;   // backward-prep: dL/d out (N, H, F)
  float* d_gate;        // backward-prep: dL/d a_t,h (E, H, F), destination-sorted edge order
  float* d_msg;         // backward-prep: dL/d v_t,h (E, H, F)
  uint32_t* signs;      // backward-prep: [2][H][kcn][E] LeakyReLU side of the 32 hidden units of a chunk
  float* bias_sums;     // backward-prep (optional): (grid, 2, H, F) per-CTA column sums of d_msg | d_gate = dL/d b2
  unsigned int* dz_amax;  // backward-prep (optional): max |d_gate|, |d_msg| as float bits (range of the f16 wgrad / dgrad)
  int n_atoms, n_edges, heads, hd;   // heads = VIRTUAL heads = real heads * vh
  float eps;
  int vh;               // 128-channel blocks per real head (F / 128): F = 256 runs every real head as two virtual heads
  unsigned int* status; // library status word: kStatusNonFinite is raised when an aggregate is not finite
};

__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int n, int64_t key) {
  int lo = 0, hi = n;  // first index with a[idx] >= key
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// max(x, 0.01 x): the same value as (x > 0 ? x : 0.01 x) for every x (signed zeros, infinities and NaN included), one
// instruction less per hidden element in producers that are bound by their instruction stream
__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, 0.01f * x); }

// exp(x) for x <= 0 on the SFU with fp32-level accuracy: ex2.approx (2^-22.5 relative) on the rounded product
// x*log2(e), times (1 + residual*ln2) where the residual carries the rounding error of that product and the low
// bits of log2(e) — without it the error grows like |x| * 6e-8.  Inputs below -87 (incl. -inf) are clamped: the
// result (1.6e-38) only ever multiplies zeros or is added to sums >= 1.
__device__ __forceinline__ float fast_exp(float x) {
  const float xc = fmaxf(x, -87.f);
  const float t = xc * 1.4426950216293335f;
  float lo = fmaf(xc, 1.4426950216293335f, -t);
  lo = fmaf(xc, 1.9259629911266175e-8f, lo);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return fmaf(e, lo * 0.6931471805599453f, e);
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// kMode 0: forward (segmented online softmax, writes out / max / den)
// kMode 1: backward-prep: recompute a, v and emit  d_msg = alpha * g,  d_gate = alpha * (v - out) * g  per edge
//          (alpha from the saved per-segment max / den) plus the sign masks of the hidden pre-activations.
// kF16: both MMA operands as fp16 hi/lo pairs on kind::f16 (twice the rate of kind::tf32 and — what matters here — half
// the bytes of the W2 stream: ncu r01c shows this kernel moving 27 KB per edge from L2, 20 KB of which is every CTA
// re-reading W2 for each 128-edge tile).  The three products share ONE accumulator (TMEM has no room for a separate
// correction accumulator next to the gate / message double buffer), so the lo parts cannot carry their own 2^11
// scale; instead the operands are PRE-scaled by powers of two (hidden x 2^4, W2 x 2^6: cgat_pack_kmajor_f16s) which
// lifts the lo parts of typical values into fp16's normal range and puts the absolute error floor of small values
// (2^-25 / scale) below the fp32 rounding error of the typical ones; the accumulator is multiplied by 2^-10 in the
// epilogue.  Valid while |hidden| < 4094 (fp16 overflow), far beyond what a non-diverged network produces.
constexpr float kHidScale = 16.f, kW2Scale = 64.f, kF16AccInv = 1.f / (16.f * 64.f);

template <int kMode, bool kF16>
__global__ void __launch_bounds__(kEThreads, 1) edge_attn_kernel(const EdgeArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  int32_t* meta = reinterpret_cast<int32_t*>(smem + kEStages * kEStageBytes);                // [4][dst|src|rank|off|flags|valid]
  float* carry = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(meta) + kEMetaBytes);  // [H][4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(carry) + kECarryBytes);
  uint64_t* full = bars;                        // [3]
  uint64_t* empty = bars + kEStages;            // [3]
  uint64_t* tmem_full = bars + 2 * kEStages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint64_t* carry_bar = tmem_empty + 2;         // [kEMaxHeads] carried softmax state of head h is in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(carry_bar + kEMaxHeads);
  int32_t* range = reinterpret_cast<int32_t*>(tmem_slot + 1);  // e_lo, e_hi

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // H counts VIRTUAL heads: output channels [h*128, h*128+128) of the (N, H_real, F) tensors, rows h*128.. of the
  // packed W2; virtual head h reads the hidden units of real head h / vh
  const int H = g.heads, hd = g.hd, vh = g.vh, HR = H / vh, hhd = HR * hd;
  constexpr int kChunk = kF16 ? kPackChunk16 : kPackChunk;  // hidden units per pipeline stage
  const int kcn = (hd + kChunk - 1) / kChunk;
  const float acc_scale = kF16 ? kF16AccInv : 1.f;

  if (tid == 0) {
    for (int s = 0; s < kEStages; ++s) {
      mbar_init(&full[s], kEGroup);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    for (int h = 0; h < kEMaxHeads; ++h) mbar_init(&carry_bar[h], 128);
    mbar_init_fence();
    // this CTA's destination-atom range: whole softmax segments, balanced by edge count
    const int G = gridDim.x;
    const int64_t t0 = (int64_t)g.n_edges * blockIdx.x / G, t1 = (int64_t)g.n_edges * (blockIdx.x + 1) / G;
    const int a_lo = lower_bound_i32(g.rowptr, g.n_atoms + 1, t0);
    const int a_hi = (blockIdx.x == G - 1) ? g.n_atoms : lower_bound_i32(g.rowptr, g.n_atoms + 1, t1);
    range[0] = g.rowptr[a_lo];
    range[1] = g.rowptr[a_hi];
  }
  if (warp == kEMmaWarp) tmem_alloc(tmem_slot, 512);
  if (kMode == 1 && tid < kEEpilogue) {
    // backward-prep does not carry softmax state across tiles; the buffer holds the per-CTA column sums of
    // d_msg / d_gate instead (the bias gradients of the second layer): [epilogue group][d_msg | d_gate][head][channel]
    for (int i = tid; i < 2 * 2 * 8 * kEF; i += kEEpilogue) carry[i] = 0.f;   // 16 KB: [2][2][8] or [2][16] x 128
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int e_lo = range[0], e_hi = range[1];
  const int n_tiles = (e_hi - e_lo + kET - 1) / kET;

  if (warp < kEEpilogue / 32) {
    // ---------------------------------------------------------------- epilogue
    // Thread = channel (TMEM lane); it walks the tile's 128 edge columns.  Segment boundaries arrive as bit masks
    // from the producers and are warp-uniform, so "a segment ended here" is a uniform branch taken about once in
    // max_nbr columns.  Softmax with a LAZY reference: m is the gate of the segment's first edge and is only moved
    // (with a rescale of den / acc) when a later gate exceeds it by more than 16 — one exp per element instead of
    // the two of the textbook online softmax; the result acc / den is the same ratio.
    const int grp = warp >> 2;
    const int c = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int hf = H * kEF;
    uint32_t hcount = 0;
    float dz_max = 0.f;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int e0 = e_lo + tile * kET;
      const int nv = min(kET, e_hi - e0);
      const int32_t* mt = meta + (tile & (kEMetaBufs - 1)) * kEMetaStride;
      const bool last_tile = (tile == n_tiles - 1);
      for (int h = 0; h < H; ++h, ++hcount) {
        const uint32_t hb = hcount & 1u;
        if ((int)hb != grp) continue;
        float* cs = carry + (h * 3) * kEF;
        int* cd = reinterpret_cast<int*>(carry + kEMaxHeads * 3 * kEF) + h;   // the open destination of head h
        const int hc = h * kEF + c;
        const float ba = __ldg(g.b2a + hc), bm = __ldg(g.b2m + hc);
        mbar_wait(&tmem_full[hb], (hcount >> 1) & 1u);
        tc_fence_after();
        const uint32_t tbase = tmem + lane_base + hb * 256;
        if (kMode == 0) {
          float m = -INFINITY, den = 0.f, acc = 0.f;
          int d = -1;
          if (tile > 0) {
            mbar_wait(&carry_bar[h], (uint32_t)(tile - 1) & 1u);  // written by the group that had (tile-1, h)
            m = cs[c], den = cs[kEF + c], acc = cs[2 * kEF + c], d = *cd;
          }
          {  // tile start: the one place that needs a compare against the carried segment
            const int d0 = mt[0];
            if (d0 != d) {
              if (d >= 0) {
                const int64_t o = (int64_t)d * hf + hc;
                const float r = acc / (den + g.eps);
                if (!(fabsf(r) <= 3.0e38f)) atomicOr(g.status, (unsigned int)kStatusNonFinite);
                g.out[o] = r;
                if (g.smax) g.smax[o] = m, g.sden[o] = den;
              }
              m = -INFINITY, den = 0.f, acc = 0.f;
            }
          }
#pragma unroll 1
          for (int cc = 0; cc < ((CGAT_EFWD_DBG & 8) ? 0 : (nv + 15) >> 4); ++cc) {   // (the CTA's last tile is partial)
            float av[16], vv[16];
            tmem_ld16(tbase + cc * 16, av);
            tmem_ld16(tbase + 128 + cc * 16, vv);
            tmem_ld_wait();
            const uint32_t sflags = ((uint32_t)mt[4 * kET + (cc >> 1)]) >> ((cc & 1) * 16);
            const uint32_t valid = ((uint32_t)mt[4 * kET + 4 + (cc >> 1)]) >> ((cc & 1) * 16);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int t = cc * 16 + j;
              if ((sflags >> j) & 1u) {  // edge t opens a new segment: flush the finished one (never at t = 0)
                const int po = mt[3 * kET + t - 1] + hc;
                const float r = acc * fast_rcp(den + g.eps);
                if (!(fabsf(r) <= 3.0e38f)) atomicOr(g.status, (unsigned int)kStatusNonFinite);
                g.out[po] = r;
                if (g.smax != nullptr) g.smax[po] = m, g.sden[po] = den;
                m = -INFINITY, den = 0.f, acc = 0.f;
              }
              const float a = ((valid >> j) & 1u) ? fmaf(av[j], acc_scale, ba) : -INFINITY;  // padding: exp(-inf) = 0
              const float v = fmaf(vv[j], acc_scale, bm);
              if (a - m > 16.f) {  // first edge of a segment (m = -inf), or a rare much larger gate
                const float r = fast_exp(m - a);
                den *= r, acc *= r, m = a;
              }
              const float p = fast_exp(a - m);
              den += p;
              acc = fmaf(p, v, acc);
            }
          }
          d = mt[nv - 1];
          tc_fence_before();
          mbar_arrive(&tmem_empty[hb]);
          if (last_tile) {
            const int64_t o = (int64_t)d * hf + hc;
            const float r = acc / (den + g.eps);
            if (!(fabsf(r) <= 3.0e38f)) atomicOr(g.status, (unsigned int)kStatusNonFinite);
            g.out[o] = r;
            if (g.smax) g.smax[o] = m, g.sden[o] = den;
          } else {
            cs[c] = m, cs[kEF + c] = den, cs[2 * kEF + c] = acc;
            if (c == 0) *cd = d;
            mbar_arrive(&carry_bar[h]);
          }
        } else {
          float* pg = g.d_gate + (int64_t)e0 * hf + hc;
          float* pm = g.d_msg + (int64_t)e0 * hf + hc;
          float sum_m = 0.f, sum_g = 0.f;  // this tile's column sums of d_msg / d_gate for (head h, channel c)
#pragma unroll 1
          for (int cc = 0; cc < ((CGAT_EFWD_DBG & 8) ? 0 : (nv + 7) >> 3); ++cc) {   // (the CTA's last tile is partial)
            // per-segment statistics re-read per column: same address for the ~max_nbr edges of a segment, so these
            // are L1 hits; it keeps the column code free of branches.  8 columns per batch: 32 loads in flight and
            // everything stays in registers under the 80-register cap of an 800-thread CTA
            float sm[8], sd[8], so[8], sg_[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int o = mt[3 * kET + cc * 8 + j] + hc;
              sm[j] = __ldg(g.smax + o), sd[j] = __ldg(g.sden + o), so[j] = __ldg(g.out + o), sg_[j] = __ldg(g.g_out + o);
            }
            float av[8], vv[8];
            tmem_ld8(tbase + cc * 8, av);
            tmem_ld8(tbase + 128 + cc * 8, vv);
            tmem_ld_wait();
            const uint32_t valid = ((uint32_t)mt[4 * kET + 4 + (cc >> 2)]) >> ((cc & 3) * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float a = fmaf(av[j], acc_scale, ba), v = fmaf(vv[j], acc_scale, bm);
              const float alpha = fast_exp(a - sm[j]) * fast_rcp(sd[j] + g.eps);
              const float ag = alpha * sg_[j];
              if ((valid >> j) & 1u) {
                const float dg = ag * (v - so[j]);
                pm[(int64_t)(cc * 8 + j) * hf] = ag;
                pg[(int64_t)(cc * 8 + j) * hf] = dg;
                sum_m += ag, sum_g += dg;
                dz_max = fmaxf(dz_max, fmaxf(fabsf(ag), fabsf(dg)));
              }
            }
          }
          tc_fence_before();
          mbar_arrive(&tmem_empty[hb]);
          // only this thread ever touches these two words: group grp handles its items one after the other.  With an
          // even head count a head always meets the same group (item parity = head parity): one array; else one per group
          const int gi = (H & 1) ? grp : 0, hs = (H & 1) ? 8 : 16;
          carry[((gi * 2 + 0) * hs + h) * kEF + c] += sum_m;
          carry[((gi * 2 + 1) * hs + h) * kEF + c] += sum_g;
        }
      }
    }
    if (kMode == 1 && g.dz_amax != nullptr) {
      // order-independent maximum (non-negative floats compare like their bit patterns): deterministic
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dz_max = fmaxf(dz_max, __shfl_xor_sync(0xffffffffu, dz_max, o));
      if (lane == 0) atomicMax(g.dz_amax, __float_as_uint(dz_max));
    }
  } else if (warp < kEMmaWarp) {
    // ---------------------------------------------------------------- producers
    const int pt = tid - kEEpilogue;
    const uint32_t grp = (uint32_t)pt >> 8;  // stage cnt is produced by group (cnt & 1)
    const int pl = pt & (kEGroup - 1);
    uint32_t cnt = 0;
    const uint8_t* w2[2] = {reinterpret_cast<const uint8_t*>(g.w2a), reinterpret_cast<const uint8_t*>(g.w2m)};
    const int64_t ldp = 4 * (int64_t)hhd, ldt = 2 * (int64_t)hhd;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int e0 = e_lo + tile * kET;
      const int nv = min(kET, e_hi - e0);
      const int n16 = (nv + 15) & ~15;
      int32_t* mt = meta + (tile & (kEMetaBufs - 1)) * kEMetaStride;
      if (pt < kET) {
        const bool ok = pt < nv;
        const int e = e0 + (ok ? pt : nv - 1);  // padding columns repeat the last valid edge (their B rows are zeros)
        const int dd = g.dst[e];
        mt[pt] = ok ? dd : -1;
        mt[kET + pt] = g.src[e];
        mt[2 * kET + pt] = g.rank[e];
        mt[3 * kET + pt] = dd * H * kEF;
        const bool sflag = ok && pt > 0 && g.dst[e - 1] != dd;  // a new segment starts here (never at t = 0)
        const uint32_t sw = __ballot_sync(0xffffffffu, sflag), vw = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) mt[4 * kET + (pt >> 5)] = (int32_t)sw, mt[4 * kET + 4 + (pt >> 5)] = (int32_t)vw;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEProducers) : "memory");
      // this thread's 4 (edge row, 16-byte chunk) slots of every stage
      int rd[4], rs[4], rr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = (pl + kEGroup * j) >> 3;
        rd[j] = mt[r], rs[j] = mt[kET + r], rr[j] = mt[2 * kET + r];  // rd < 0 marks a padding row
      }
      for (int h = 0; h < H; ++h) {
        for (int net = 0; net < 2; ++net) {
          for (int kc = 0; kc < kcn; ++kc, ++cnt) {
            if (kEProducers > kEGroup && (cnt & 1u) != grp) continue;
            const uint32_t s = cnt % kEStages, u = cnt / kEStages;
            mbar_wait(&empty[s], (u + 1) & 1u);
            uint8_t* st = stages + s * kEStageBytes;
            if (pl == 0 && !(CGAT_EFWD_DBG & 4)) {
              mbar_expect_tx(&full[s], kPackStageBytes);
              bulk_g2s(st, w2[net] + ((int64_t)h * kcn + kc) * kPackStageBytes, kPackStageBytes, &full[s]);
            }
            if constexpr (kF16) {
              // slot = (edge row r, 16-byte chunk c) = 8 consecutive hidden units: two float4 from each of the three
              // gathered rows; two slots at a time (12 float4 in flight, like the tf32 path)
              uint8_t* bh = st + kPackStageBytes;
              const int col0 = (h / vh) * hd + kc * kPackChunk16;
#pragma unroll 1
              for (int j0 = 0; j0 < ((CGAT_EFWD_DBG & 2) ? 0 : 4); j0 += 2) {
                // rows beyond the last 16-column group with edges are not read by this tile's MMAs (N = n16): a group
                // of 8 lanes shares a row, so the skip is uniform per quarter-warp
                if (((pl + kEGroup * j0) >> 3) >= n16) break;   // rows grow with j0
                float4 pd[2][2], ps[2][2], te[2][2];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                  const int idx = pl + kEGroup * (j0 + jj);
                  const int r = idx >> 3, cch = idx & 7;
                  const int col = col0 + cch * 8;
                  // row ids straight from the tile metadata in shared memory (a register array indexed by the
                  // loop counter would live in local memory: ncu r01h showed 16 % of the samples waiting on it)
                  const int d = mt[r];
                  const bool ok = d >= 0 && (kc * kPackChunk16 + cch * 8) < hd;
#pragma unroll
                  for (int q = 0; q < 2; ++q) pd[jj][q] = ps[jj][q] = te[jj][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (ok) {
                    ldg_v8(g.P + (int64_t)d * ldp + net * hhd + col, pd[jj][0], pd[jj][1]);
                    ldg_v8(g.P + (int64_t)mt[kET + r] * ldp + 2 * hhd + net * hhd + col, ps[jj][0], ps[jj][1]);
                    ldg_v8(g.T + (int64_t)mt[2 * kET + r] * ldt + net * hhd + col, te[jj][0], te[jj][1]);
                  }
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                  const int idx = pl + kEGroup * (j0 + jj);
                  float x[8];  // the sums replace the gathered dst rows register for register
#pragma unroll
                  for (int q = 0; q < 2; ++q) {
                    x[4 * q] = (pd[jj][q].x + ps[jj][q].x) + te[jj][q].x;
                    x[4 * q + 1] = (pd[jj][q].y + ps[jj][q].y) + te[jj][q].y;
                    x[4 * q + 2] = (pd[jj][q].z + ps[jj][q].z) + te[jj][q].z;
                    x[4 * q + 3] = (pd[jj][q].w + ps[jj][q].w) + te[jj][q].w;
                  }
                  if (kMode == 1) {
                    // sign word of (edge, 32 hidden units) in the layout the dgrad epilogue reads: hidden unit u of the
                    // chunk sits at bit (u & 3) * 8 + (u >> 2).  This slot holds u = (c & 3) * 8 + t, t = 0..7; the four
                    // slots of a word live in four adjacent lanes.
                    const int c = idx & 7;
                    uint32_t part = 0;
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                      part |= (x[t] > 0.f ? 1u : 0u) << ((t & 3) * 8 + (c & 3) * 2 + (t >> 2));
                    part |= __shfl_xor_sync(0xffffffffu, part, 1);
                    part |= __shfl_xor_sync(0xffffffffu, part, 2);
                    const int r = idx >> 3;
                    if ((lane & 3) == 0 && r < nv && h % vh == 0)   // one sign word per REAL head
                      g.signs[((int64_t)(net * HR + h / vh) * (2 * kcn) + 2 * kc + (c >> 2)) * g.n_edges + e0 + r] = part;
                  }
#pragma unroll
                  for (int t = 0; t < 8; ++t) x[t] = lrelu(x[t]) * kHidScale;
                  uint4 hi, lo;
                  split_f16x8s(make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), 1.f, hi, lo);
                  const uint32_t off = sw128_offset(idx >> 3, idx & 7);
                  *reinterpret_cast<uint4*>(bh + off) = hi;
                  *reinterpret_cast<uint4*>(bh + kPackImageBytes + off) = lo;
                }
              }
            } else {
              float4 pd[4], ps[4], te[4];
              const int col0 = (h / vh) * hd + kc * 32;
  #pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int cch = (pl + kEGroup * j) & 7;
                const int col = col0 + cch * 4;
                const bool ok = rd[j] >= 0 && (kc * 32 + cch * 4) < hd;
                pd[j] = ps[j] = te[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {
                  pd[j] = __ldg(reinterpret_cast<const float4*>(g.P + (int64_t)rd[j] * ldp + net * hhd + col));
                  ps[j] = __ldg(reinterpret_cast<const float4*>(g.P + (int64_t)rs[j] * ldp + 2 * hhd + net * hhd + col));
                  te[j] = __ldg(reinterpret_cast<const float4*>(g.T + (int64_t)rr[j] * ldt + net * hhd + col));
                }
              }
              uint8_t* bh = st + kPackStageBytes;
  #pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int idx = pl + kEGroup * j;
                float4 x;
                x.x = pd[j].x + ps[j].x + te[j].x;
                x.y = pd[j].y + ps[j].y + te[j].y;
                x.z = pd[j].z + ps[j].z + te[j].z;
                x.w = pd[j].w + ps[j].w + te[j].w;
                if (kMode == 1) {
                  // which side of the LeakyReLU each hidden pre-activation is on: byte j' = component j' of the
                  // float4, bit c = 16-byte chunk c of the 128-byte row  ->  hidden unit c*4 + j' of this K chunk
                  const uint32_t b0 = __ballot_sync(0xffffffffu, x.x > 0.f), b1 = __ballot_sync(0xffffffffu, x.y > 0.f);
                  const uint32_t b2 = __ballot_sync(0xffffffffu, x.z > 0.f), b3 = __ballot_sync(0xffffffffu, x.w > 0.f);
                  const int sh = lane & 24;  // the 8 lanes that share an edge row
                  const uint32_t word = ((b0 >> sh) & 0xFFu) | (((b1 >> sh) & 0xFFu) << 8) | (((b2 >> sh) & 0xFFu) << 16) |
                                        (((b3 >> sh) & 0xFFu) << 24);
                  const int r = idx >> 3;
                  if ((lane & 7) == 0 && r < nv && h % vh == 0)
                    g.signs[((int64_t)(net * HR + h / vh) * kcn + kc) * g.n_edges + e0 + r] = word;
                }
                x.x = lrelu(x.x), x.y = lrelu(x.y), x.z = lrelu(x.z), x.w = lrelu(x.w);
                float4 hi, lo;
                split_tf32(x, hi, lo);
                const uint32_t off = sw128_offset(idx >> 3, idx & 7);
                *reinterpret_cast<float4*>(bh + off) = hi;
                *reinterpret_cast<float4*>(bh + kPackImageBytes + off) = lo;
              }
            }
            fence_async_smem();
            mbar_arrive(&full[s]);
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- MMA issuer
    uint32_t cnt = 0, hcount = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
      // the CTA's last tile is usually partial (E / 148 edges = 3.5 tiles at the bench size): its MMAs cover only the
      // 16-column groups that hold edges, and the producers / epilogue skip the rest
      const uint32_t n16 = (uint32_t)((min(kET, e_hi - (e_lo + tile * kET)) + 15) & ~15);
      const uint32_t idesc = kF16 ? umma_idesc_f16(kEF, n16) : umma_idesc_tf32(kEF, n16);
      for (int h = 0; h < H; ++h, ++hcount) {
        const uint32_t hb = hcount & 1u;
        mbar_wait(&tmem_empty[hb], ((hcount >> 1) + 1) & 1u);
        tc_fence_after();
        for (int net = 0; net < 2; ++net) {
          const uint32_t d = tmem + hb * 256 + net * 128;
          for (int kc = 0; kc < kcn; ++kc, ++cnt) {
            const uint32_t s = cnt % kEStages, u = cnt / kEStages;
            mbar_wait(&full[s], u & 1u);
            tc_fence_after();
            {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
              const uint32_t a_hi = smem_u32(stages + s * kEStageBytes), a_lo = a_hi + kPackImageBytes;
              const uint32_t b_hi = a_hi + kPackStageBytes, b_lo = b_hi + kPackImageBytes;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t off = ks * 32;
                if constexpr (kF16 && (CGAT_EFWD_DBG & 1)) {
                } else if constexpr (kF16) {  // K = 16 halves = the same 32 bytes of the swizzled row
                  umma_f16_e(d, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, (kc | ks) != 0);
                  umma_f16_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc, 1);
                  umma_f16_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc, 1);
                } else {
                  umma_tf32_e(d, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, (kc | ks) != 0);
                  umma_tf32_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc, 1);
                  umma_tf32_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc, 1);
                }
              }
              umma_commit_e(&empty[s]);
              if (net == 1 && kc == kcn - 1) umma_commit_e(&tmem_full[hb]);
            }
            __syncwarp();
          }
        }
      }
    }
  }
  __syncthreads();
  if (kMode == 1 && g.bias_sums != nullptr) {
    float* dst = g.bias_sums + (int64_t)blockIdx.x * 2 * H * kEF;
    for (int i = tid; i < 2 * H * kEF; i += kEThreads) {
      const int which = i / (H * kEF), hc = i - which * H * kEF, h = hc / kEF, c = hc - h * kEF;
      dst[i] = (H & 1) ? carry[((0 * 2 + which) * 8 + h) * kEF + c] + carry[((1 * 2 + which) * 8 + h) * kEF + c]
                       : carry[(which * 16 + h) * kEF + c];
    }
  }
  if (warp == kEMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

namespace {
int check_edge_args(int64_t n_atoms, int64_t n_edges, int32_t heads, int32_t f, int32_t hd) {
  if (f != kEF && f != 2 * kEF) return fail(-2, "cgat_edge_attn_*: F must be 128 or 256 (vector attention)");
  if (heads < 1 || heads * (f / kEF) > kEMaxHeads) return fail(-2, "cgat_edge_attn_*: heads * F / 128 must be in [1,16]");
  if (hd <= 0 || (hd & 3)) return fail(-2, "cgat_edge_attn_*: hidden width must be a multiple of 4");
  if (n_atoms * heads * f >= (1ll << 31) || n_edges >= (1ll << 31) - 129) return fail(-2, "cgat_edge_attn_*: size overflow");
  return 0;
}

template <int kMode, bool kF16 = false>
int launch_edge(EdgeArgs a, cudaStream_t stream) {
  const int smem_bytes = kESmemBytes;
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(edge_attn_kernel<kMode, kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kESmemBytes));
    configured = true;
  }
  const int64_t tiles = ceil_div(a.n_edges, kET);
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  edge_attn_kernel<kMode, kF16><<<grid, kEThreads, smem_bytes, stream>>>(a);
  return check_launch(kMode == 0 ? "edge_attn_fwd_kernel" : "edge_attn_bwd_prep_kernel");
}
}  // namespace

namespace {
template <bool kF16>
int edge_fwd_impl(const float* P, const float* T, const int32_t* rowptr, const int32_t* src, const int32_t* dst,
                  const int32_t* rank, const float* w2a_packed, const float* w2m_packed, const float* b2a,
                  const float* b2m, float* out, float* seg_max, float* seg_den, int64_t n_atoms, int64_t n_edges,
                  int32_t heads, int32_t f, int32_t hd, float eps, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_edge_args(n_atoms, n_edges, heads, f, hd)) return e;
  if (kF16 && (hd & 63)) return fail(-2, "cgat_edge_attn_fwd_f16: hidden width must be a multiple of 64");
  if (n_atoms <= 0) return 0;
  const size_t out_bytes = sizeof(float) * (size_t)n_atoms * heads * f;
  CGAT_CUDA(cudaMemsetAsync(out, 0, out_bytes, stream));  // atoms without in-edges aggregate to 0
  if (seg_max) {
    CGAT_CUDA(cudaMemsetAsync(seg_max, 0, out_bytes, stream));
    CGAT_CUDA(cudaMemsetAsync(seg_den, 0, out_bytes, stream));
  }
  if (n_edges <= 0) return 0;
  EdgeArgs a{P, T, rowptr, src, dst, rank, w2a_packed, w2m_packed, b2a, b2m, out, seg_max, seg_den,
             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, (int)n_atoms, (int)n_edges, heads * (f / kEF), hd,
             eps, f / kEF, status_word()};
  return launch_edge<0, kF16>(a, stream);
}

template <bool kF16>
int edge_bwd_prep_impl(const float* P, const float* T, const int32_t* rowptr, const int32_t* src, const int32_t* dst,
                       const int32_t* rank, const float* w2a_packed, const float* w2m_packed, const float* b2a,
                       const float* b2m, const float* out, const float* seg_max, const float* seg_den,
                       const float* g_out, float* d_gate, float* d_msg, uint32_t* signs, float* bias_sums,
                       float* dz_amax, int64_t n_atoms, int64_t n_edges, int32_t heads, int32_t f, int32_t hd, float eps,
                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_edge_args(n_atoms, n_edges, heads, f, hd)) return e;
  if (kF16 && (hd & 63)) return fail(-2, "cgat_edge_attn_bwd_prep_f16: hidden width must be a multiple of 64");
  if (dz_amax) CGAT_CUDA(cudaMemsetAsync(dz_amax, 0, sizeof(float), stream));
  if (n_atoms <= 0 || n_edges <= 0) return 0;
  EdgeArgs a{P, T, rowptr, src, dst, rank, w2a_packed, w2m_packed, b2a, b2m, const_cast<float*>(out),
             const_cast<float*>(seg_max), const_cast<float*>(seg_den), g_out, d_gate, d_msg, signs, bias_sums,
             reinterpret_cast<unsigned int*>(dz_amax), (int)n_atoms, (int)n_edges, heads * (f / kEF), hd, eps,
             f / kEF, status_word()};
  return launch_edge<1, kF16>(a, stream);
}
}  // namespace

// number of CTAs cgat_edge_attn_bwd_prep launches = rows of its bias_sums output
extern "C" int32_t cgat_edge_attn_grid(int64_t n_edges) {
  const int64_t tiles = ceil_div(n_edges, kET);
  return (int32_t)(tiles < kNumSMs ? (tiles > 0 ? tiles : 1) : kNumSMs);
}

extern "C" int cgat_edge_attn_fwd(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                                  const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                                  const float* w2m_packed, const float* b2a, const float* b2m, float* out,
                                  float* seg_max, float* seg_den, int64_t n_atoms, int64_t n_edges, int32_t heads,
                                  int32_t f, int32_t hd, float eps, void* stream_) {
  return edge_fwd_impl<false>(P, T, rowptr, src, dst, rank, w2a_packed, w2m_packed, b2a, b2m, out, seg_max, seg_den,
                              n_atoms, n_edges, heads, f, hd, eps, stream_);
}

// The same kernel on kind::f16 passes; w2a/w2m_packed = cgat_pack_kmajor_f16s(W2, pre_scale 64, lo_scale 1).
extern "C" int cgat_edge_attn_fwd_f16(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                                      const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                                      const float* w2m_packed, const float* b2a, const float* b2m, float* out,
                                      float* seg_max, float* seg_den, int64_t n_atoms, int64_t n_edges, int32_t heads,
                                      int32_t f, int32_t hd, float eps, void* stream_) {
  return edge_fwd_impl<true>(P, T, rowptr, src, dst, rank, w2a_packed, w2m_packed, b2a, b2m, out, seg_max, seg_den,
                             n_atoms, n_edges, heads, f, hd, eps, stream_);
}

// Backward, step 1: per-edge gradients of the second-layer outputs.  Recomputes a_t,h / v_t,h exactly like
// the forward and writes d_gate, d_msg (E, H, F) in destination-sorted edge order plus the LeakyReLU sign
// masks signs[2][H][ceil(hd/32)][E] that the dgrad kernel needs (32 hidden units per word).
extern "C" int cgat_edge_attn_bwd_prep(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                                       const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                                       const float* w2m_packed, const float* b2a, const float* b2m, const float* out,
                                       const float* seg_max, const float* seg_den, const float* g_out, float* d_gate,
                                       float* d_msg, uint32_t* signs, float* bias_sums, float* dz_amax, int64_t n_atoms, int64_t n_edges, int32_t heads,
                                       int32_t f, int32_t hd, float eps, void* stream_) {
  return edge_bwd_prep_impl<false>(P, T, rowptr, src, dst, rank, w2a_packed, w2m_packed, b2a, b2m, out, seg_max, seg_den,
                                   g_out, d_gate, d_msg, signs, bias_sums, dz_amax, n_atoms, n_edges, heads, f, hd, eps, stream_);
}

extern "C" int cgat_edge_attn_bwd_prep_f16(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                                       const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                                       const float* w2m_packed, const float* b2a, const float* b2m, const float* out,
                                       const float* seg_max, const float* seg_den, const float* g_out, float* d_gate,
                                       float* d_msg, uint32_t* signs, float* bias_sums, float* dz_amax, int64_t n_atoms, int64_t n_edges, int32_t heads,
                                       int32_t f, int32_t hd, float eps, void* stream_) {
  return edge_bwd_prep_impl<true>(P, T, rowptr, src, dst, rank, w2a_packed, w2m_packed, b2a, b2m, out, seg_max, seg_den,
                                  g_out, d_gate, d_msg, signs, bias_sums, dz_amax, n_atoms, n_edges, heads, f, hd, eps, stream_);
}
