// Fused edge-attention backward, step 2 (second-layer dgrad) — the product form for F = 128: d_pre on kind::f16.
// (SURVEY.md §8a row A12 for rows A2-A4; reference CGAT/CGAT.py:319-329 differentiated.)
//
//   d_pre[e, net*H*Hd + h*Hd + k] = leaky_relu'(pre[e, ...]) * sum_c dZ_net[e, h, c] * W2_net[h*F + c, k]
//
// Same arithmetic as edge_dgrad_kernel<true, true> (edge_attn_bwd.cu: hidden units on the 128 TMEM lanes, a tile of 128
// edges on the columns, W2^T tiles streamed by the copy engine, dZ converted to scaled fp16 hi/lo pairs while staged) —
// rebuilt around what the ablation runs of that kernel showed (profiles/r03p, r03q: switch a piece off, time the rest):
//   * with the stores, the MMAs, the dZ staging and the W2 stream ALL off it still took 241 of 468 us: the epilogue
//     fetched its LeakyReLU sign words with 16 broadcast loads per batch of 16 columns, eight exposed L2 round trips
//     per tile.  Here a warp fetches the 128 words of a tile with one coalesced load, one step ahead, and hands them
//     out by shuffles; the accumulator loads are software-pipelined.
//   * staging dZ cost 150-250 us: every (net, head) tile of dZ was loaded and converted once per 128-unit HALF of the
//     hidden layer (Hd = 256: twice).  Here the converted tile stays in shared memory for all halves: the ring is split
//     into a W2 ring (3 x 32 KB, copy engine) and two whole-K dZ buffers (2 x 64 KB).
//   * the producers looked up segment metadata nobody reads in this form and met at a 256-thread barrier per tile.
// Measured and dropped: taking the per-destination segment sums of d_pre in this epilogue (thread = hidden unit, running
// sum over the destination-sorted columns, a store at every segment end) so that cgat_edge_attn_reduce only makes its
// by-source pass — the reduce went from 330 to 208 us but this kernel from 260 to 417 us (one warp-uniform branch per
// column splits the unrolled column loop into basic blocks and the epilogue is what the kernel waits for).
#include "common.cuh"
#include "tc_common.cuh"

#ifndef CGAT_EDGE_DBG
#define CGAT_EDGE_DBG 0   // timing experiments: 1 no d_pre stores, 2 no MMAs, 4 no dZ staging, 8 no W2 stream
#endif

namespace cgat {
namespace {
using namespace tc;

constexpr int kT = 128;                 // edges per tile
constexpr int kFz = 128;                // channels per head
constexpr int kKc = kFz / kPackChunk16; // K chunks of 64 channels: 2
constexpr int kEpi = 256;               // two epilogue groups of 4 warps, one per TMEM buffer
constexpr int kProd = 256;
constexpr int kMmaW = (kEpi + kProd) / 32;
constexpr int kStreamW = kMmaW + 1;
constexpr int kThreadsZ = kEpi + kProd + 64;
constexpr int kWStagesZ = 3;
constexpr int kZBytes = kKc * (int)kPackStageBytes;   // one dZ tile, all K: [kc][hi|lo] = 64 KB
constexpr int kSmemZ = 2 * kZBytes + kWStagesZ * (int)kPackStageBytes + 256 + 1024;

struct DgradZArgs {
  const float* d_gate;     // (E, H, F) rows in destination-sorted order
  const float* d_msg;
  const uint32_t* signs;   // [2][H][kcn][E]
  const int32_t* segptr;   // (N+1) CSR pointer (CTA ranges start on segment boundaries, like the other passes)
  const float* wt_a;       // packed f16 (H * ceil(Hd/128)*128, F): row h*Hd+k, col c = W2A[h*F+c, k]
  const float* wt_m;
  float* d_pre;            // (E, 2*H*Hd)
  const float* dz_amax;
  int n_atoms, n_edges, heads, hd;
};

__device__ __forceinline__ int lower_bound_z(const int32_t* a, int n, int64_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kThreadsZ, 1) edge_dgrad_zr_kernel(const DgradZArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* zbuf = smem;                                    // [2][kc][hi|lo][16 KB]
  uint8_t* wring = smem + 2 * kZBytes;                     // [3][hi|lo][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + kWStagesZ * kPackStageBytes);
  uint64_t* z_full = bars;                 // [2] producers -> MMA
  uint64_t* z_empty = bars + 2;            // [2] MMA -> producers
  uint64_t* w_full = bars + 4;             // [3] copy engine -> MMA
  uint64_t* w_empty = bars + 7;            // [3] MMA -> stream warp
  uint64_t* tmem_full = bars + 10;         // [2] MMA -> epilogue group
  uint64_t* tmem_empty = bars + 12;        // [2] epilogue group -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  int32_t* range = reinterpret_cast<int32_t*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = g.heads, hd = g.hd, hhd = H * hd;
  const int kcn = (hd + 31) / 32;       // sign words per (net, head, edge)
  const int nhalf = (hd + 127) / 128;   // M tiles of 128 hidden units per head
  float s_scale = 1.f, s_inv = 1.f;
  {
    const float amax = __ldg(g.dz_amax);
    int ex;
    frexpf(amax, &ex);
    if (amax > 0.f && amax < INFINITY) s_scale = ldexpf(1.f, 4 - ex), s_inv = ldexpf(1.f, ex - 4);
  }

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&z_full[b], kProd);
      mbar_init(&z_empty[b], 1);
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 4);      // one arrival per warp of the owning epilogue group
    }
    for (int s = 0; s < kWStagesZ; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init_fence();
    const int G_ = gridDim.x;
    const int64_t t0 = (int64_t)g.n_edges * blockIdx.x / G_, t1 = (int64_t)g.n_edges * (blockIdx.x + 1) / G_;
    const int a_lo = lower_bound_z(g.segptr, g.n_atoms + 1, t0);
    const int a_hi = (blockIdx.x == G_ - 1) ? g.n_atoms : lower_bound_z(g.segptr, g.n_atoms + 1, t1);
    range[0] = g.segptr[a_lo];
    range[1] = g.segptr[a_hi];
  }
  if (warp == kMmaW) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int e_lo = range[0], e_hi = range[1];
  const int n_tiles = (e_hi - e_lo + kT - 1) / kT;
  const int n_z = 2 * H * n_tiles;            // dZ tiles: (net, head) outer, edge tile inner
  const int n_steps = n_z * nhalf;            // accumulator tiles: (dZ tile, half)

  if (warp < kEpi / 32) {
    // ---------------------------------------------------------------- epilogue
    const int grp = warp >> 2, q = warp & 3;          // group = TMEM buffer; q = lane quadrant
    const int c = q * 32 + lane;                      // hidden unit within the half = TMEM lane
    const uint32_t bitpos = (uint32_t)((lane & 3) * 8 + (lane >> 2));
    const int64_t ldd = 2 * (int64_t)hhd;
    auto load_signs = [&](int step, uint32_t (&sw)[4]) {
      sw[0] = sw[1] = sw[2] = sw[3] = 0u;
      if (step >= n_steps) return;
      const int zc = step / nhalf, half = step - zc * nhalf;
      const int nh = zc / n_tiles, tile = zc - nh * n_tiles;
      if (half * 128 + q * 32 >= hd) return;   // a warp of padded hidden units (warp-uniform: the lanes serve each other)
      const uint32_t* sg = g.signs + ((int64_t)nh * kcn + (half * 4 + q)) * g.n_edges;   // nh = net * H + h
      const int e0 = e_lo + tile * kT, nv = min(kT, e_hi - e0);
      const uint32_t* sp = sg + e0 + 4 * lane;
      if (4 * lane + 4 <= nv && (reinterpret_cast<uintptr_t>(sp) & 15) == 0) {
        const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(sp));
        sw[0] = v4.x, sw[1] = v4.y, sw[2] = v4.z, sw[3] = v4.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (4 * lane + j < nv) sw[j] = __ldg(sp + j);
      }
    };
    uint32_t sw[4], nsw[4];
    load_signs(grp, nsw);
    for (int step = grp; step < n_steps; step += 2) {
      const int zc = step / nhalf, half = step - zc * nhalf;
      const int nh = zc / n_tiles, tile = zc - nh * n_tiles;
      const int kk = half * 128 + c;
      const bool kvalid = kk < hd;
      const int col = nh * hd + kk;                    // = net*H*Hd + h*Hd + kk
      const int e0 = e_lo + tile * kT, nv = min(kT, e_hi - e0);
#pragma unroll
      for (int j = 0; j < 4; ++j) sw[j] = nsw[j];
      load_signs(step + 2, nsw);
      mbar_wait(&tmem_full[grp], (uint32_t)(step >> 1) & 1u);
      tc_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + grp * 256;
      float* prow = g.d_pre + (int64_t)e0 * ldd + col;
      // 16 columns per batch (96-register cap of the CTA), the next batch's accumulator loads in flight during the math
      float v[2][16], w[2][16];
      tmem_ld16(tbase, v[0]);
      tmem_ld16(tbase + 128, w[0]);
#pragma unroll
      for (int cc = 0; cc < kT / 16; ++cc) {   // (stopping at the last column group with edges was measured: the branch
        tmem_ld_wait();                       //  costs the straight-line schedule more than the partial tile saves)
        if (cc + 1 < kT / 16) {
          tmem_ld16(tbase + (cc + 1) * 16, v[(cc + 1) & 1]);
          tmem_ld16(tbase + 128 + (cc + 1) * 16, w[(cc + 1) & 1]);
        }
        const int left = kvalid ? nv - cc * 16 : 0;  // columns of this batch that are real edges
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t word = __shfl_sync(0xffffffffu, sw[j & 3], cc * 4 + (j >> 2));
          const float dp = fmaf(w[cc & 1][j], kF16LoInv, v[cc & 1][j]) * (((word >> bitpos) & 1u) ? s_inv : 0.01f * s_inv);
          if (j < left && !(CGAT_EDGE_DBG & 1)) *prow = dp;
          if ((CGAT_EDGE_DBG & 1) && dp == 12345.f) *prow = dp;
          prow += ldd;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[grp]);
    }
  } else if (warp < kMmaW) {
    // ---------------------------------------------------------------- producers: dZ tiles -> fp16 hi/lo images
    const int pt = tid - kEpi;
    const float* dz[2] = {g.d_gate, g.d_msg};
    // slot = (edge row, 16-byte chunk) = 8 consecutive channels: two float4 per slot, 4 slots per thread and K chunk.
    // The loads of chunk q+1 are in flight while chunk q is converted (q runs over (dZ tile, K chunk)).
    auto load = [&](int zc, int kc, float4* aa, float4* bb) {
      if (zc >= n_z) {   // past the end (also the empty edge range: n_tiles = 0)
#pragma unroll
        for (int j = 0; j < 4; ++j) aa[j] = bb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
      }
      const int nh = zc / n_tiles, tile = zc - nh * n_tiles;
      const int net = nh / H, h = nh - net * H;
      const int e0 = e_lo + tile * kT, nv = min(kT, e_hi - e0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = pt + kProd * j, r = idx >> 3, cch = idx & 7;
        aa[j] = bb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nv && !(CGAT_EDGE_DBG & 4)) {   // padding rows are staged as zeros
          const float4* p = reinterpret_cast<const float4*>(dz[net] + ((int64_t)(e0 + r) * H + h) * kFz + kc * kPackChunk16 + cch * 8);
          aa[j] = __ldg(p), bb[j] = __ldg(p + 1);
        }
      }
    };
    auto store = [&](uint8_t* bh, const float4* xa, const float4* xb, int n16) {
#pragma unroll
      for (int j = 0; j < ((CGAT_EDGE_DBG & 4) ? 0 : 4); ++j) {
        const int idx = pt + kProd * j;
        if ((idx >> 3) >= n16) break;   // rows beyond the tile's last 16-column group with edges are never read
        const float4 a = make_float4(xa[j].x * s_scale, xa[j].y * s_scale, xa[j].z * s_scale, xa[j].w * s_scale);
        const float4 b = make_float4(xb[j].x * s_scale, xb[j].y * s_scale, xb[j].z * s_scale, xb[j].w * s_scale);
        uint4 hi, lo;
        split_f16x8(a, b, hi, lo);
        const uint32_t off = sw128_offset(idx >> 3, idx & 7);
        *reinterpret_cast<uint4*>(bh + off) = hi;
        *reinterpret_cast<uint4*>(bh + kPackImageBytes + off) = lo;
      }
    };
    float4 a0[4], b0[4], a1[4], b1[4];
    load(0, 0, a0, b0);
    for (int zc = 0; zc < n_z; ++zc) {
      const uint32_t zb = (uint32_t)zc & 1u;
      uint8_t* zt = zbuf + zb * kZBytes;
      const int tile_z = zc % n_tiles;
      const int n16 = (min(kT, e_hi - (e_lo + tile_z * kT)) + 15) & ~15;
      load(zc, 1, a1, b1);
      mbar_wait(&z_empty[zb], (((uint32_t)zc >> 1) + 1) & 1u);   // the MMAs of the tile before last have read the buffer
      store(zt, a0, b0, n16);
      load(zc + 1, 0, a0, b0);
      store(zt + kPackStageBytes, a1, b1, n16);
      fence_async_smem();
      mbar_arrive(&z_full[zb]);
    }
  } else if (warp == kStreamW) {
    // ---------------------------------------------------------------- W2^T stream (one thread)
    if (lane == 0) {
      const uint8_t* wt[2] = {reinterpret_cast<const uint8_t*>(g.wt_a), reinterpret_cast<const uint8_t*>(g.wt_m)};
      uint32_t wc = 0;
      for (int zc = 0; zc < n_z; ++zc) {
        const int nh = zc / n_tiles, net = nh / H, h = nh - net * H;
        for (int half = 0; half < nhalf; ++half)
          for (int kc = 0; kc < kKc; ++kc, ++wc) {
            const uint32_t s = wc % kWStagesZ, u = wc / kWStagesZ;
            mbar_wait(&w_empty[s], (u + 1) & 1u);
            if (CGAT_EDGE_DBG & 8) {
              mbar_arrive(&w_full[s]);
              continue;
            }
            mbar_arrive_expect_tx(&w_full[s], kPackStageBytes);
            bulk_g2s(wring + s * kPackStageBytes, wt[net] + ((int64_t)(h * nhalf + half) * kKc + kc) * kPackStageBytes,
                     kPackStageBytes, &w_full[s]);
          }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- MMA issuer (all lanes, one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_f16(128, kT), idesc2 = umma_idesc_f16(128, 2 * kT);
    uint32_t wc = 0, step = 0;
    for (int zc = 0; zc < n_z; ++zc) {
      const uint32_t zb = (uint32_t)zc & 1u;
      mbar_wait(&z_full[zb], ((uint32_t)zc >> 1) & 1u);
      tc_fence_after();
      const uint32_t z_base = smem_u32(zbuf + zb * kZBytes);
      const uint32_t n16 = (uint32_t)((min(kT, e_hi - (e_lo + (zc % n_tiles) * kT)) + 15) & ~15);
      const uint32_t idesc_p = umma_idesc_f16(128, n16);   // partial (last) tile of the CTA: N = n16 < 128
      for (int half = 0; half < nhalf; ++half, ++step) {
        const uint32_t b = step & 1u;
        mbar_wait(&tmem_empty[b], ((step >> 1) + 1) & 1u);
        tc_fence_after();
        const uint32_t d = tmem + b * 256, dc = d + 128;
        for (int kc = 0; kc < kKc; ++kc, ++wc) {
          const uint32_t s = wc % kWStagesZ, u = wc / kWStagesZ;
          mbar_wait(&w_full[s], u & 1u);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(wring + s * kPackStageBytes), a_lo = a_hi + kPackImageBytes;
          const uint32_t b_hi = z_base + kc * kPackStageBytes, b_lo = b_hi + kPackImageBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t off = ks * 32;
            if (!(CGAT_EDGE_DBG & 2)) {
              // the hi and lo images of the dZ chunk are adjacent: one N = 256 MMA multiplies W_hi with both (main |
              // correction columns), one N = 128 MMA adds W_lo * dZ_hi to the correction columns
              if (n16 == kT) {
                umma_f16_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc2, (kc | ks) != 0);
                umma_f16_e(dc, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, 1);
              } else {   // the lo image does not follow the first n16 rows of the hi image: three N = n16 MMAs
                umma_f16_e(dc, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc_p, (kc | ks) != 0);
                umma_f16_e(dc, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc_p, 1);
                umma_f16_e(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc_p, (kc | ks) != 0);
              }
            }
          }
          umma_commit_e(&w_empty[s]);
          if (kc == kKc - 1) umma_commit_e(&tmem_full[b]);
        }
      }
      umma_commit_e(&z_empty[zb]);   // every half of this tile has been issued: the buffer is free once they complete
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == kMmaW) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace
}  // namespace cgat

using namespace cgat;

// Called by cgat_edge_attn_dgrad_f16 (edge_attn_bwd.cu) for F = 128; same arguments and result.
int cgat_edge_dgrad_zr_launch(const float* d_gate, const float* d_msg, const uint32_t* signs, const int32_t* segptr,
                              const float* wt_a_packed, const float* wt_m_packed, const float* dz_amax, float* d_pre,
                              int64_t n_atoms, int64_t n_edges, int32_t heads, int32_t hd, int32_t grid,
                              cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(edge_dgrad_zr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemZ));
    configured = true;
  }
  DgradZArgs a{d_gate, d_msg, signs, segptr, wt_a_packed, wt_m_packed, d_pre, dz_amax, (int)n_atoms, (int)n_edges,
               heads, hd};
  edge_dgrad_zr_kernel<<<grid, kThreadsZ, kSmemZ, stream>>>(a);
  return check_launch("edge_dgrad_zr_kernel");
}
