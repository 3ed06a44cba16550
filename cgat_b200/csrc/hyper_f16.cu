// Fused hypernetwork linear layer on kind::f16 tensor-core passes ("f16x3", tc_common.cuh) — same arithmetic,
// roles and pipeline as hyper_fwd.cu (reference CGAT/Hypernetworksmp.py:243-254 HyperLinear.forward
// + :205-209 BatchLinear), with both MMA operands (the z / y activation tile and the pre-packed Linear weight) as
// scaled fp16 hi/lo pairs instead of tf32 hi/lo pairs:
//   * the same 22 significand bits per operand (parity bar unchanged: tests/test_gpu_kernels.py runs both forms),
//   * kind::f16 issues twice the flops per instruction of kind::tf32 for the same 32-byte K slab, so the operand bytes
//     read from shared memory per flop — the resource that paces the M = N = 128 SS-mode MMAs — halve,
//   * the z tile needs 64 KB instead of 128 KB of shared memory, which pays for a 4-stage weight ring (was 3) and
//     the forward's 32 KB cross-group reduction buffer,
//   * with the MMAs twice as fast the epilogue (one tcgen05.ld + FMA per accumulator element) paces the kernel
//     (ncu r01c: tensor pipe 33 % active, epilogue warps stalled on tcgen05.ld latency), so it runs on SIXTEEN warps:
//     4 lane quadrants x 4 groups of 32 accumulator columns, 32 live row values per thread,
//   * the packed weight image (L2-resident, streamed once per atom tile) is 8.4 MB instead of 16.8 MB.
// Operands here are activations (tanh / LayerNorm outputs, aggregated messages) and weights; the gradient-operand
// kernels (hyper_wgrad, edge dgrad / wgrad) stay on tf32, whose 8-bit exponent needs no scaling.
#include "common.cuh"
#include "tc_common.cuh"

#ifndef CGAT_HYPER_DBG
#define CGAT_HYPER_DBG 0
#endif
#if CGAT_HYPER_DBG & 64   // timing experiment: per-CTA timeline (SM clock) of the last launch
__device__ long long g_hyper_tl[160][16];
#define TL(slot) (g_hyper_tl[blockIdx.x][slot] = clock64())
#define TLV(slot, v) (g_hyper_tl[blockIdx.x][slot] = (long long)(v))
#else
#define TL(slot)
#define TLV(slot, v)
#endif
namespace cgat {
namespace {
using namespace tc;

template <int F>
struct Hyper16Cfg {
  static constexpr int kNH = F / 128;                           // 128-column halves of a predicted-weight row (F = 256: 2)
  static constexpr int kKC = F / kPackChunk16;                  // K chunks of 64 halves (one 128-byte swizzled row)
  static constexpr int kABytes = kKC * (int)kPackStageBytes;    // activation tile, hi+lo per chunk: 64 KB (F = 256: 128 KB)
  static constexpr int kStages = F == 128 ? 4 : 2;              // 32 KB weight stages
  static constexpr int kRedBytes = 4 * 16 * 128 * 4;            // forward: [column group][output of the item][atom row]
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kABytes + kStages * (int)kPackStageBytes + kRedBytes + 1024 + kBarBytes;
  static constexpr int kEpiWarps = 16;                          // 4 column groups x 4 lane quadrants
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kThreads = kEpiThreads + 128 + 64;       // + stagers + MMA warp + weight-stream warp
  static constexpr int kTmemCols = 512;                         // two buffers x (main, correction) x 128 columns
};

// Stage one 128-atom activation tile as fp16 hi/lo K-major images: NT threads (t = 0..NT-1), slot = (K chunk kc, row r,
// 16-byte chunk c) = 8 consecutive K elements: two float4 loads -> 8 hi halves + 8 lo halves; four slots at a time
// (8 float4 in flight) to stay inside the 80-register budget of the CTA.
template <int F, int NT>
__device__ __forceinline__ void hyper16_stage_tile(const float* __restrict__ z, int tile, int n_atoms, uint8_t* a_smem,
                                                   int t) {
  constexpr int kSlots = (F / (int)kPackChunk16) * 1024;
#pragma unroll 1
  for (int base = 0; base < kSlots; base += NT * 4) {
    float4 va[4], vb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = base + t + NT * j, kc = idx >> 10, r = (idx & 1023) >> 3, c = idx & 7;
      const int gr = tile * 128 + r;
      va[j] = vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < kSlots && gr < n_atoms) {
        const float4* p = reinterpret_cast<const float4*>(z + (int64_t)gr * F + kc * kPackChunk16 + c * 8);
        va[j] = __ldg(p), vb[j] = __ldg(p + 1);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = base + t + NT * j, kc = idx >> 10;
      if (idx < kSlots) {
        uint8_t* hi = a_smem + kc * kPackStageBytes;
        const uint32_t off = sw128_offset((idx & 1023) >> 3, idx & 7);
        uint4 h, l;
        split_f16x8(va[j], vb[j], h, l);
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(hi + kPackImageBytes + off) = l;
      }
    }
  }
}

// kMode 0 (forward):   y_out[n,o] = sum_j D_o[n,j] * y_in[n,j] + e_term[n,o]
// kMode 1 (backward):  partial[chunk][n,j] = sum_{o in chunk} y_in[n,o] * D_o[n,j]
//   D_o[n,j] = sum_m z[n,m] Wblk_o[j,m] (+ w_bias[o*F+j]);  see hyper_fwd.cu for the two backward uses.
// F = 256 (BASELINE.json configs[3], the wider net): a predicted-weight row D_o[n, :] has two 128-column halves, each
// its own 128-row tile of the packed weight and its own accumulator pass.  Forward walks both halves inside a work
// item (the half dot products meet in `red`); backward makes (atom tile, half) a "virtual tile" of the item index, so
// that a CTA's register accumulators still cover exactly one 128-column block of the result.
template <int F, int kMode>
__global__ void __launch_bounds__(Hyper16Cfg<F>::kThreads, 1)
hyper_rowdot_f16_kernel(const float* __restrict__ z, const float* __restrict__ y_in, const float* __restrict__ e_term,
                        const float* __restrict__ e_term2, const float* __restrict__ w_bias,
                        const float* __restrict__ w_packed, float* __restrict__ y_out, int n_atoms, int oc, int n_slots,
                        int split, unsigned int* __restrict__ scale_amax) {
  using Cfg = Hyper16Cfg<F>;
  static_assert(F == 128 || F == 256, "instantiated for F = 128 and F = 256");
  constexpr int NH = Cfg::kNH;
  constexpr int kHalvesPerItem = kMode == 0 ? NH : 1;   // forward: both halves inside the item; backward: half = part of the tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                   // [kKC][hi|lo][16 KB]
  uint8_t* b_smem = smem + Cfg::kABytes;                    // [kStages][hi|lo][16 KB]
  float* red = reinterpret_cast<float*>(b_smem + Cfg::kStages * kPackStageBytes);  // [4][16][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + Cfg::kRedBytes);
  uint64_t* full = bars;                                    // [kStages] TMA -> MMA
  uint64_t* empty = bars + Cfg::kStages;                    // [kStages] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;            // [2] MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;                     // [2] epilogue -> MMA
  uint64_t* a_full = tmem_empty + 2;                        // stagers -> MMA
  uint64_t* a_free = a_full + 1;                            // MMA -> stagers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#if CGAT_HYPER_DBG & 64
  if (tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    TL(0); TLV(1, gt);
    for (int i = 2; i < 16; ++i) TLV(i, 0);
  }
#endif
  // A cluster of kCS CTAs works on kCS consecutive atom tiles (a "tile group") side by side and walks the SAME
  // sequence of weight stages: the leader CTA fetches every 32 KB stage ONCE from L2 and the copy engine multicasts it
  // into the same shared-memory offset of every CTA of the cluster (the weight stream from L2, 64 KB per output channel
  // and CTA, was what the MMA warp waited for: ncu r02y, 45 % of all samples on the epilogue's wait for the accumulator).
  constexpr int kCS = kHyperCluster;
  const int crank = (int)cluster_ctarank();
  const int cid = (int)blockIdx.x / kCS, n_clusters = (int)gridDim.x / kCS;
  const int n_tiles = (n_atoms + 127) / 128;
  const int n_tg = (n_tiles + kCS - 1) / kCS;
  const int n_chunks = F / oc;
  const int n_vt = kMode == 0 ? n_tg : n_tg * NH;   // (virtual) tile groups: backward splits a tile into its column halves
  const int n_items = n_vt * n_chunks;
  // Work decomposition, identical in all three roles.
  // split == 0: work item = ((virtual) tile group, chunk of oc output channels), chunk fastest, a contiguous item range
  //   per CLUSTER.  A range may cross into the next tile group: the activation tile is then re-staged with the whole
  //   pipeline drained (per-CTA timeline profiles/r03j: 8 us out of 60, on the CTAs that decide the kernel's duration).
  // split > 0 (there are enough clusters to give every (virtual) tile group `split` of them): cluster -> ONE tile group
  //   and a 4-aligned slice [o_lo, o_hi) of its output channels, walked in items of <= oc outputs; nothing is re-staged
  //   and the backward form writes exactly `split` partials per tile.
  int item_lo, item_hi, vt_fix = 0, o_lo = 0, o_hi = 0;
  if (split > 0) {
    vt_fix = cid / split;
    const int r = cid - vt_fix * split;
    o_lo = 4 * ((F / 4) * r / split), o_hi = 4 * ((F / 4) * (r + 1) / split);
    item_lo = 0, item_hi = (o_hi - o_lo + oc - 1) / oc;
  } else {
    item_lo = (int)((int64_t)n_items * cid / n_clusters);
    item_hi = (int)((int64_t)n_items * (cid + 1) / n_clusters);
  }
  auto item_vt = [&](int item) { return split > 0 ? vt_fix : item / n_chunks; };
  auto item_o0 = [&](int item) { return split > 0 ? o_lo + item * oc : (item % n_chunks) * oc; };
  auto item_cnt = [&](int item) { return split > 0 ? min(oc, o_hi - (o_lo + item * oc)) : oc; };

  if (tid == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kCS);   // the MMAs of EVERY CTA of the cluster have read the stage
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], Cfg::kEpiThreads / 32);
    }
    mbar_init(a_full, 128);
    mbar_init(a_free, 1);
    mbar_init_fence();
  }
  constexpr int kMmaWarp = Cfg::kEpiWarps + 4;
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers are initialised before anyone multicasts into / arrives on them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) { TL(2); TLV(10, item_hi - item_lo); }

  // The FIRST activation tile is staged by the epilogue warps and the stagers together (640 threads, one pass: the
  // epilogue has nothing to do until the first accumulator is ready, and until this tile is in shared memory nothing
  // else can start either — timeline profiles/r03j: 2.5 us with 128 threads); the weight stream starts meanwhile.
  constexpr int kCoop = Cfg::kEpiThreads + 128;
  if (tid < kCoop && item_lo < item_hi) {
    const int vt0 = item_vt(item_lo);
    hyper16_stage_tile<F, kCoop>(z, (kMode == 0 ? vt0 : vt0 / NH) * kCS + crank, n_atoms, a_smem, tid);
    fence_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(kCoop) : "memory");
  }

  if (warp < Cfg::kEpiWarps) {
    // ------------------------------------------------------------------ epilogue
    // 16 warps: a warp reads the TMEM lanes 32*(warp%4)..; column group grp = warp/4 owns columns [32 grp, 32 grp + 32)
    // of every accumulator, so a thread keeps a quarter of its atom's row.  kMode 0: the four partial dot products of
    // an output meet in shared memory (red[grp][output][row], conflict-free) and are summed, together with the e
    // term, by all 512 threads once per work item; kMode 1 needs no exchange.
    constexpr int QF = 32;   // a thread keeps a quarter of a 128-column half of its atom's row
    const int grp = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ocount = 0;
    float y[QF];
    float sc_max = 0.f;   // kMode 1: max |scale| seen by this thread (the gradient operand of the weight-gradient kernel)
    for (int item = item_lo; item < item_hi; ++item) {
      const int vt = item_vt(item), o0 = item_o0(item), ocnt = item_cnt(item);
      const bool first_of_run = split > 0 ? item == 0 : (item == item_lo || o0 == 0);
      const bool last_of_run = split > 0 ? item + 1 == item_hi : (item + 1 == item_hi || o0 + oc == F);
      const int tile = (kMode == 0 ? vt : vt / NH) * kCS + crank, vhalf = kMode == 0 ? 0 : vt % NH;
      const int n = tile * 128 + row;
      const bool valid = n < n_atoms;
      for (int hh = 0; hh < kHalvesPerItem; ++hh) {
        const int half = kMode == 0 ? hh : vhalf;
        // kMode 0: this atom's quarter of the y_in half; kMode 1: the running partial sums over o, carried across the
        // consecutive chunks of one virtual tile (a new run starts with the CTA's first item or a tile's first chunk)
        if (kMode == 0 || first_of_run) {
#pragma unroll
          for (int j = 0; j < QF / 4; ++j) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kMode == 0 && valid)
              t = __ldg(reinterpret_cast<const float4*>(y_in + (int64_t)n * F + half * 128 + grp * QF) + j);
            y[4 * j] = t.x, y[4 * j + 1] = t.y, y[4 * j + 2] = t.z, y[4 * j + 3] = t.w;
          }
        }
        // kMode 1: the scale g[n, o] of output o+1 is requested while output o is consumed — a dependent global load
        // right behind the accumulator barrier was an exposed L2 round trip per output channel
        float sc_next = (kMode == 1 && valid) ? __ldg(y_in + (int64_t)n * F + o0) : 0.f;
        for (int oi = 0; oi < ocnt; ++oi, ++ocount) {
          const int o = o0 + oi;
          const uint32_t b = ocount & 1u;
          const float sc = sc_next;
          if (kMode == 1 && valid && oi + 1 < ocnt) sc_next = __ldg(y_in + (int64_t)n * F + o + 1);
          if (kMode == 1) sc_max = fmaxf(sc_max, fabsf(sc));
          // The bias of the predicted weight row, p[n, o*F + j] = D_o[n, j] + bl[o*F + j], does not depend on the MMAs:
          // its contribution is taken BEFORE waiting for the accumulator (broadcast loads: every thread of a column
          // group reads the same 32 floats), so only tcgen05.ld + two FMAs per element sit behind the barrier.
          float acc = 0.f;
          if (w_bias != nullptr && !(CGAT_HYPER_DBG & 32)) {   // 32 = timing experiment: no bias term
            const float4* bp = reinterpret_cast<const float4*>(w_bias + (int64_t)o * F + half * 128 + grp * QF);
#pragma unroll
            for (int q = 0; q < QF / 4; ++q) {
              const float4 b4 = __ldg(bp + q);
              if (kMode == 0) {
                acc = fmaf(b4.x, y[4 * q], acc), acc = fmaf(b4.y, y[4 * q + 1], acc);
                acc = fmaf(b4.z, y[4 * q + 2], acc), acc = fmaf(b4.w, y[4 * q + 3], acc);
              } else {
                y[4 * q] = fmaf(b4.x, sc, y[4 * q]), y[4 * q + 1] = fmaf(b4.y, sc, y[4 * q + 1]);
                y[4 * q + 2] = fmaf(b4.z, sc, y[4 * q + 2]), y[4 * q + 3] = fmaf(b4.w, sc, y[4 * q + 3]);
              }
            }
          }
          mbar_wait(&tmem_full[b], (ocount >> 1) & 1u);
          tc_fence_after();
          if (tid == 0 && ocount == 0) TL(5);
          const uint32_t tb = tmem + lane_base + b * 256 + grp * QF;
          // 8 columns per batch (80 registers per thread with 21 warps, 32 of them hold the row values), software
          // pipelined: the loads of batch cc+1 are in flight while batch cc is consumed.  v = hi*hi products,
          // w = (hi*lo + lo*hi products) * 2^11.
          float v[2][8], w[2][8];
#if (CGAT_HYPER_DBG & 3) == 1   // timing experiment: no correction-column loads
#pragma unroll
          for (int j = 0; j < 8; ++j) w[0][j] = w[1][j] = 0.f;
#define DBG_LDW(a, b)
#define DBG_LDV(a, b) tmem_ld8(a, b)
#elif (CGAT_HYPER_DBG & 2)  // timing experiment: no accumulator loads at all
#pragma unroll
          for (int j = 0; j < 8; ++j) w[0][j] = w[1][j] = v[0][j] = v[1][j] = sc;
#define DBG_LDW(a, b)
#define DBG_LDV(a, b)
#else
#define DBG_LDW(a, b) tmem_ld8(a, b)
#define DBG_LDV(a, b) tmem_ld8(a, b)
#endif
          DBG_LDV(tb, v[0]);
          DBG_LDW(tb + 128, w[0]);
#pragma unroll
          for (int cc = 0; cc < QF / 8; ++cc) {
            tmem_ld_wait();
            if (cc + 1 < QF / 8) {
              DBG_LDV(tb + (cc + 1) * 8, v[(cc + 1) & 1]);
              DBG_LDW(tb + 128 + (cc + 1) * 8, w[(cc + 1) & 1]);
            }
#pragma unroll
            for (int j = 0; j < ((CGAT_HYPER_DBG & 32) ? 1 : 8); ++j) {   // 32 = timing experiment: (almost) no epilogue math
              const float t = fmaf(w[cc & 1][j], kF16LoInv, v[cc & 1][j]);
              if (kMode == 0) acc = fmaf(t, y[cc * 8 + j], acc);
              else y[cc * 8 + j] = fmaf(t, sc, y[cc * 8 + j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[b]);   // one arrival per warp: 512 per-thread arrivals on one word serialise
          if (tid == 0) { TL(6); TLV(11, ocount + 1); }
          if (kMode == 0) {
            float* rp = red + (grp * 16 + oi) * 128 + row;   // only this thread touches the word: halves add up in place
            *rp = hh == 0 ? acc : *rp + acc;
          }
        }
      }
      if (kMode == 0) {
        asm volatile("bar.sync 2, %0;" ::"n"(Cfg::kEpiThreads) : "memory");  // all partials of this item are in `red`
        // thread -> (atom row, 4 consecutive outputs of the item): y_out = (p0 + p1) + (p2 + p3) + e
        const int q0 = grp * 4;
        if (q0 < ocnt && valid) {
          float r[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float* pr = red + (q0 + q) * 128 + row;
            r[q] = (pr[0] + pr[16 * 128]) + (pr[2 * 16 * 128] + pr[3 * 16 * 128]);
          }
          const int64_t off = (int64_t)n * F + o0 + q0;
          const float4 e1 = __ldg(reinterpret_cast<const float4*>(e_term + off));
          float4 e2 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e_term2 != nullptr) e2 = __ldg(reinterpret_cast<const float4*>(e_term2 + off));
          *reinterpret_cast<float4*>(y_out + off) =
              make_float4(r[0] + (e1.x + e2.x), r[1] + (e1.y + e2.y), r[2] + (e1.z + e2.z), r[3] + (e1.w + e2.w));
        }
        asm volatile("bar.sync 3, %0;" ::"n"(Cfg::kEpiThreads) : "memory");  // `red` may be overwritten by the next item
      }
      if (kMode == 1 && valid && last_of_run) {
        // end of this CTA's run on the (virtual) tile: one partial per (CTA, tile); the CTA that finishes the tile also
        // clears the slots nobody used (the caller sums all n_slots)
        const int slot = split > 0 ? cid - vt_fix * split : cid - hyper_cta_of_item((int64_t)vt * n_chunks, n_items, n_clusters);
        const bool finishes_tile = split > 0 ? slot == split - 1 : o0 + oc == F;
        const int64_t col = vhalf * 128 + grp * QF;
        float4* dst = reinterpret_cast<float4*>(y_out + ((int64_t)slot * n_atoms + n) * F + col);
#pragma unroll
        for (int j = 0; j < QF / 4; ++j) dst[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        if (finishes_tile)
          for (int sl = slot + 1; sl < n_slots; ++sl) {
            float4* z4 = reinterpret_cast<float4*>(y_out + ((int64_t)sl * n_atoms + n) * F + col);
#pragma unroll
            for (int j = 0; j < QF / 4; ++j) z4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
      }
    }
    if (kMode == 1 && scale_amax != nullptr) {
      // order-independent (hence deterministic) maximum: non-negative floats compare like their bit patterns
      sc_max = fmaxf(sc_max, __shfl_xor_sync(0xffffffffu, sc_max, 16));
      sc_max = fmaxf(sc_max, __shfl_xor_sync(0xffffffffu, sc_max, 8));
      sc_max = fmaxf(sc_max, __shfl_xor_sync(0xffffffffu, sc_max, 4));
      sc_max = fmaxf(sc_max, __shfl_xor_sync(0xffffffffu, sc_max, 2));
      sc_max = fmaxf(sc_max, __shfl_xor_sync(0xffffffffu, sc_max, 1));
      if (lane == 0 && grp == 0) atomicMax(scale_amax, __float_as_uint(sc_max));
    }
  } else if (warp < kMmaWarp) {
    // ------------------------------------------------------------------ activation-tile stagers
    const int st = tid - Cfg::kEpiThreads;  // 0..127
    uint32_t it = 0;
    int staged_tile = -1;
    for (int item = item_lo; item < item_hi; ++item) {
      const int vt = item_vt(item);
      const int tile = (kMode == 0 ? vt : vt / NH) * kCS + crank;
      const bool restage = tile != staged_tile;
      staged_tile = tile;
      if (restage && st == 0 && it > 0) TL(14);
      if (restage) mbar_wait(a_free, (it + 1) & 1u);  // the previous tile's MMAs have finished reading the tile
      if (restage && st == 0 && it > 0) TL(12);
      if (restage && it > 0) hyper16_stage_tile<F, 128>(z, tile, n_atoms, a_smem, st);   // (the first tile is already there)
      if (restage) {
        fence_async_smem();
        mbar_arrive(a_full);
        if (st == 0) { if (it == 0) TL(3); else TL(13); }
        ++it;
      }
    }
  } else if (warp == kMmaWarp + 1) {
    // ------------------------------------------------------------------ weight stream (one thread)
    uint32_t cnt = 0;
    for (int item = item_lo; item < item_hi; ++item) {
      const int vt = item_vt(item), o0 = item_o0(item), ocnt = item_cnt(item);
      const int vhalf = kMode == 0 ? 0 : vt % NH;
      if (lane == 0) {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(w_packed);
        for (int hh = 0; hh < kHalvesPerItem; ++hh) {
          const int half = kMode == 0 ? hh : vhalf;
          for (int oi = 0; oi < ocnt; ++oi) {
            const int64_t rt = (int64_t)(o0 + oi) * NH + half;   // 128-row tile of the packed weight
            for (int kc = 0; kc < Cfg::kKC; ++kc, ++cnt) {
              const uint32_t s = cnt % Cfg::kStages, u = cnt / Cfg::kStages;
              mbar_wait(&empty[s], (u + 1) & 1u);   // all CTAs of the cluster are done with the stage's last contents
#if CGAT_HYPER_DBG & 16   // timing experiment: no weight stream (the MMAs read whatever the stage holds)
              mbar_arrive(&full[s]);
              continue;
#endif
              mbar_arrive_expect_tx(&full[s], kPackStageBytes);
              if constexpr (kCS == 1)
                bulk_g2s(b_smem + s * kPackStageBytes, wsrc + (rt * Cfg::kKC + kc) * kPackStageBytes, kPackStageBytes,
                         &full[s]);
              else if (crank == 0)
                bulk_g2s_multicast(b_smem + s * kPackStageBytes, wsrc + (rt * Cfg::kKC + kc) * kPackStageBytes,
                                   kPackStageBytes, &full[s], (uint16_t)((1u << kCS) - 1));
            }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(128, 128), idesc2 = umma_idesc_f16(128, 256);
    uint32_t it = 0, cnt = 0, ocount = 0;
    int staged_tile = -1;
    // (A second issuing warp, one per accumulator buffer, was measured: no change — after the converged issue below the
    // kernel is paced by the tensor pipe's operand path, not by the issuing thread.)
    for (int item = item_lo; item < item_hi; ++item) {
      const int vt = item_vt(item), ocnt = item_cnt(item);
      const int tile = kMode == 0 ? vt : vt / NH;   // tile GROUP: this CTA's tile is tile * kCS + crank
      if (tile != staged_tile) {
        mbar_wait(a_full, it & 1u);
        ++it;
        staged_tile = tile;
      }
      tc_fence_after();
      for (int oi = 0; oi < ocnt * kHalvesPerItem; ++oi, ++ocount) {
        const uint32_t b = ocount & 1u;
        mbar_wait(&tmem_empty[b], ((ocount >> 1) + 1) & 1u);
        tc_fence_after();
        for (int kc = 0; kc < Cfg::kKC; ++kc, ++cnt) {
          const uint32_t s = cnt % Cfg::kStages, u = cnt / Cfg::kStages;
          mbar_wait(&full[s], u & 1u);
          tc_fence_after();
          if (lane == 0 && cnt == 0) TL(4);
          {
            // all 32 lanes, warp-uniform operands; one elected lane issues (tc_common.cuh, "_e" forms)
            const uint32_t b_hi = smem_u32(b_smem + s * kPackStageBytes);
            const uint32_t a_hi = smem_u32(a_smem + kc * kPackStageBytes), a_lo = a_hi + kPackImageBytes;
            const uint32_t d = tmem + b * 256, dc = d + 128;
            // one N = 256 MMA multiplies a_hi with the adjacent [b_hi; b_lo] images (main | correction columns), one
            // N = 128 MMA adds a_lo * b_hi to the correction columns; each K step is 16 halves = 32 bytes of the row
            const uint64_t da_hi = umma_desc_k_sw128(a_hi), da_lo = umma_desc_k_sw128(a_lo), db = umma_desc_k_sw128(b_hi);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t off = (uint64_t)(ks * 2);   // 32 bytes along the swizzled row = 2 units of the address field
#if !(CGAT_HYPER_DBG & 4)   // 4 = timing experiment: no MMAs, only the commits
              umma_f16_e(d, da_hi + off, db + off, idesc2, (kc | ks) != 0);
#if !(CGAT_HYPER_DBG & 8)   // 8 = timing experiment: only the N = 256 MMA
              umma_f16_e(dc, da_lo + off, db + off, idesc, 1);
#endif
#endif
            }
            if constexpr (kCS == 1) umma_commit_e(&empty[s]);
            else umma_commit_multicast_e(&empty[s], (uint16_t)((1u << kCS) - 1));
            if (kc == Cfg::kKC - 1) umma_commit_e(&tmem_full[b]);
          }
          __syncwarp();
        }
      }
      // last item of this atom tile: the activation tile may be overwritten once these MMAs are done
      const int next_tile = item + 1 == item_hi ? -1 : (kMode == 0 ? item_vt(item + 1) : item_vt(item + 1) / NH);
      if (next_tile != tile) umma_commit_e(a_free);
      __syncwarp();
    }
  }
  if (tid == 0) TL(7);
  __syncthreads();
  if (tid == 0) TL(8);
  cluster_sync_all();   // nobody leaves while a peer may still multicast into its shared memory or arrive on its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::kTmemCols);
#if CGAT_HYPER_DBG & 64
    if (lane == 0) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      TL(9); TLV(15, gt);
    }
#endif
  }
}

}  // namespace
}  // namespace cgat
#if CGAT_HYPER_DBG & 64
extern "C" int cgat_debug_hyper_timeline(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_hyper_tl, sizeof(g_hyper_tl));
}
#endif

using namespace cgat;

namespace {
// Launch plan of one f16 hyper kernel: tile-aligned (split > 0) when that is estimated to finish sooner, else the
// contiguous item ranges of common.cuh (see the kernel's "work decomposition" comment).
struct Hyper16Plan {
  int oc, split, n_clusters, n_slots;
};

template <int F, int kMode>
cudaLaunchConfig_t hyper16_config(cudaLaunchAttribute* attr, cudaStream_t stream) {
  using Cfg = Hyper16Cfg<F>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)kNumSMs);
  cfg.blockDim = dim3((unsigned)Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kHyperCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cfg;
}

// a cluster needs its CTAs on SMs of one GPC: GPCs with an odd number of SMs leave one unpaired, so fewer than
// 148 / 2 clusters can be resident at once — a persistent grid must not exceed that or its tail runs as a second wave
template <int F, int kMode>
int hyper16_max_clusters() {
  static int max_clusters = 0;
  if (max_clusters == 0) {
    using Cfg = Hyper16Cfg<F>;
    if (cudaFuncSetAttribute(hyper_rowdot_f16_kernel<F, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::kSmemBytes) != cudaSuccess)
      cudaGetLastError();   // the launch itself reports the failure
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = hyper16_config<F, kMode>(attr, nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, hyper_rowdot_f16_kernel<F, kMode>, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = kNumSMs / kHyperCluster - 4;
    }
    max_clusters = n;
  }
  return max_clusters;
}

template <int F, int kMode>
Hyper16Plan hyper16_plan(int64_t n_atoms) {
  const int maxc = hyper16_max_clusters<F, kMode>();
  const int64_t n_tiles = (n_atoms + 127) / 128;
  const int64_t n_tg = (n_tiles + kHyperCluster - 1) / kHyperCluster;
  const int64_t n_vt = n_tg * (kMode ? F / 128 : 1);
  Hyper16Plan p;
  p.oc = hyper_chunk(n_atoms, F);
  p.split = 0;
  p.n_clusters = hyper_grid(n_atoms, F, kMode, kHyperCluster);
  if (p.n_clusters > maxc) p.n_clusters = maxc;
  p.n_slots = hyper_parts(n_atoms, F);
  if (n_vt <= maxc) {
    int split = (int)(maxc / n_vt);
    if (split > F / 4) split = F / 4;
    // outputs walked by the busiest cluster: a 4-aligned slice of the F channels — against the contiguous ranges plus
    // the cost of one mid-run re-staging (about 8 outputs' worth, profiles/r03j) when ranges cross tiles
    const int64_t n_items = hyper_items(n_atoms, F, kMode, kHyperCluster);
    const int64_t t_aligned = 4 * ((F / 4 + split - 1) / split);
    const int64_t t_ranges = (n_items + p.n_clusters - 1) / p.n_clusters * p.oc + (n_items % p.n_clusters ? 8 : 0);
    if (t_aligned <= t_ranges) {
      p.split = split;
      p.oc = 16;
      p.n_clusters = (int)n_vt * split;
      p.n_slots = split;
    }
  }
  return p;
}

template <int F, int kMode>
int launch_hyper16_f(const float* z, const float* y_in, const float* e_term, const float* e_term2, const float* w_bias,
                     const float* w_packed, float* y_out, int64_t n_atoms, cudaStream_t stream, float* scale_amax) {
  const Hyper16Plan p = hyper16_plan<F, kMode>(n_atoms);
  cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t cfg = hyper16_config<F, kMode>(attr, stream);
  cfg.gridDim = dim3((unsigned)(p.n_clusters * kHyperCluster));
  const int n_atoms_i = (int)n_atoms;
  unsigned int* amax = reinterpret_cast<unsigned int*>(scale_amax);
  CGAT_CUDA(cudaLaunchKernelEx(&cfg, hyper_rowdot_f16_kernel<F, kMode>, z, y_in, e_term, e_term2, w_bias, w_packed, y_out,
                               n_atoms_i, p.oc, p.n_slots, p.split, amax));
  return check_launch(kMode == 0 ? "hyper_rowdot_f16_kernel" : "hyper_rowscale_f16_kernel");
}

template <int kMode>
int launch_hyper16(const float* z, const float* y_in, const float* e_term, const float* e_term2, const float* w_bias,
                   const float* w_packed, float* y_out, int64_t n_atoms, int32_t f, cudaStream_t stream,
                   float* scale_amax = nullptr) {
  if (n_atoms <= 0) return 0;
  if (f != 128 && f != 256) return fail(-2, "cgat_hyper_*_f16: instantiated for F = 128 and F = 256");
  if (n_atoms >= (1ll << 31) - 128) return fail(-2, "cgat_hyper_*_f16: too many atoms");
  if (w_bias && (reinterpret_cast<uintptr_t>(w_bias) & 15)) return fail(-2, "cgat_hyper_*_f16: w_bias must be 16-byte aligned");
  if (f == 128)
    return launch_hyper16_f<128, kMode>(z, y_in, e_term, e_term2, w_bias, w_packed, y_out, n_atoms, stream, scale_amax);
  return launch_hyper16_f<256, kMode>(z, y_in, e_term, e_term2, w_bias, w_packed, y_out, n_atoms, stream, scale_amax);
}
}  // namespace

// partial slots cgat_hyper_rowscale_f16[_amax] writes (and the caller sums) for this problem size
extern "C" int32_t cgat_hyper_rowscale_parts_f16(int64_t n_atoms, int32_t f) {
  if (n_atoms <= 0 || (f != 128 && f != 256)) return 1;
  return f == 128 ? hyper16_plan<128, 1>(n_atoms).n_slots : hyper16_plan<256, 1>(n_atoms).n_slots;
}

// cgat_hyper_rowdot_fwd with w_packed = cgat_pack_kmajor_f16 of W[:F*F, :F]  (same arguments and result)
extern "C" int cgat_hyper_rowdot_fwd_f16(const float* z, const float* y_in, const float* e_term, const float* e_term2,
                                         const float* w_bias, const float* w_packed, float* y_out, int64_t n_atoms,
                                         int32_t f, void* stream_) {
  return launch_hyper16<0>(z, y_in, e_term, e_term2, w_bias, w_packed, y_out, n_atoms, f, (cudaStream_t)stream_);
}

// cgat_hyper_rowscale with w_packed = cgat_pack_kmajor_f16 of the F stacked blocks; partial: (cgat_hyper_rowscale_parts, N, F)
extern "C" int cgat_hyper_rowscale_f16(const float* a, const float* scale, const float* w_bias, const float* w_packed,
                                       float* partial, int64_t n_atoms, int32_t f, void* stream_) {
  return launch_hyper16<1>(a, scale, nullptr, nullptr, w_bias, w_packed, partial, n_atoms, f, (cudaStream_t)stream_);
}

// The same, and additionally scale_amax[0] = max(scale_amax[0], max |scale|) (device float, zero it beforehand): the
// power-of-two range of the gradient operand that cgat_hyper_wgrad_f16 needs, taken while `scale` is being read anyway.
extern "C" int cgat_hyper_rowscale_f16_amax(const float* a, const float* scale, const float* w_bias,
                                            const float* w_packed, float* partial, float* scale_amax, int64_t n_atoms,
                                            int32_t f, void* stream_) {
  return launch_hyper16<1>(a, scale, nullptr, nullptr, w_bias, w_packed, partial, n_atoms, f, (cudaStream_t)stream_,
                           scale_amax);
}
