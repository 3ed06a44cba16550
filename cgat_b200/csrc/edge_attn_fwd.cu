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
// segments), balanced by edge count.  Roles (416 threads):
//   warps 0-3   epilogue: tcgen05.ld gate/message rows, segmented online softmax, write out[d,h,:]
//   warps 4-11  producers: gather + LeakyReLU + split -> shared memory (B operand); thread 0 also issues the
//               cp.async.bulk of the pre-packed W2 chunk (A operand); 3-stage full/empty mbarrier ring
//   warp  12    TMEM allocation (512 columns: 2 heads x {gate, message} x 128) + single-thread MMA issue
#include "common.cuh"
#include "tc_common.cuh"

namespace cgat {
namespace {
using namespace tc;

constexpr int kET = 128;             // edges per tile (MMA N)
constexpr int kEF = 128;             // output channels per head (MMA M) — instantiated for F = 128
constexpr int kEProducers = 256;
constexpr int kEThreads = 128 + kEProducers + 32;
constexpr int kEStages = 3;
constexpr int kEStageBytes = 2 * (int)kPackStageBytes;  // [W2 chunk hi|lo][hidden chunk hi|lo]
constexpr int kEMaxHeads = 8;
constexpr int kEMetaBufs = 4;          // tile metadata ring (producers run ahead of the epilogue)
constexpr int kEMetaBytes = kEMetaBufs * 3 * kET * 4;
constexpr int kECarryBytes = kEMaxHeads * 4 * kEF * 4;  // (max, den, acc, open dst) per head and channel
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
  int n_atoms, n_edges, heads, hd;
  float eps;
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

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.01f * x; }

// kMode 0: forward (segmented online softmax, writes out / max / den)
// kMode 1: backward-prep: recompute a, v and emit  d_msg = alpha * g,  d_gate = alpha * (v - out) * g  per edge
//          (alpha from the saved per-segment max / den) plus the sign masks of the hidden pre-activations.
template <int kMode>
__global__ void __launch_bounds__(kEThreads, 1) edge_attn_kernel(const EdgeArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  int32_t* meta = reinterpret_cast<int32_t*>(smem + kEStages * kEStageBytes);          // [4][3][128]
  float* carry = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(meta) + kEMetaBytes);  // [H][4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(carry) + kECarryBytes);
  uint64_t* full = bars;                        // [3]
  uint64_t* empty = bars + kEStages;            // [3]
  uint64_t* tmem_full = bars + 2 * kEStages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int32_t* range = reinterpret_cast<int32_t*>(tmem_slot + 1);  // e_lo, e_hi

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = g.heads, hd = g.hd, hhd = H * hd;
  const int kcn = (hd + 31) / 32;

  if (tid == 0) {
    for (int s = 0; s < kEStages; ++s) {
      mbar_init(&full[s], kEProducers);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    mbar_init_fence();
    // this CTA's destination-atom range: whole softmax segments, balanced by edge count
    const int G = gridDim.x;
    const int64_t t0 = (int64_t)g.n_edges * blockIdx.x / G, t1 = (int64_t)g.n_edges * (blockIdx.x + 1) / G;
    const int a_lo = lower_bound_i32(g.rowptr, g.n_atoms + 1, t0);
    const int a_hi = (blockIdx.x == G - 1) ? g.n_atoms : lower_bound_i32(g.rowptr, g.n_atoms + 1, t1);
    range[0] = g.rowptr[a_lo];
    range[1] = g.rowptr[a_hi];
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int e_lo = range[0], e_hi = range[1];
  const int n_tiles = (e_hi - e_lo + kET - 1) / kET;

  if (warp < 4) {
    // ---------------------------------------------------------------- epilogue
    const int c = warp * 32 + lane;  // channel = TMEM lane
    uint32_t hcount = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int e0 = e_lo + tile * kET;
      const int nv = min(kET, e_hi - e0);
      const int32_t* mdst = meta + (tile & (kEMetaBufs - 1)) * 3 * kET;
      const bool last_tile = (tile == n_tiles - 1);
      for (int h = 0; h < H; ++h, ++hcount) {
        const uint32_t hb = hcount & 1u;
        float* cs = carry + (h * 4) * kEF;
        float m = -INFINITY, den = 0.f, acc = 0.f;  // kMode 1 reuses them as (seg max, 1/(den+eps), out) of the open segment
        float gg = 0.f;
        int d = -1;
        if (kMode == 0 && tile > 0) {
          m = cs[c], den = cs[kEF + c], acc = cs[2 * kEF + c], d = __float_as_int(cs[3 * kEF + c]);
        }
        const float ba = __ldg(g.b2a + h * kEF + c), bm = __ldg(g.b2m + h * kEF + c);
        mbar_wait(&tmem_full[hb], (hcount >> 1) & 1u);
        tc_fence_after();
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + hb * 256;
#pragma unroll 1
        for (int cc = 0; cc < kET / 32; ++cc) {
          float av[32], vv[32];
          tmem_ld32(tbase + cc * 32, av);
          tmem_ld32(tbase + 128 + cc * 32, vv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int t = cc * 32 + j;
            if (t < nv) {
              const int dc = mdst[t];
              const float a = av[j] + ba, v = vv[j] + bm;
              if (kMode == 0) {
                if (dc != d) {
                  if (d >= 0) {
                    const int64_t o = ((int64_t)d * H + h) * kEF + c;
                    g.out[o] = acc / (den + g.eps);
                    if (g.smax) g.smax[o] = m, g.sden[o] = den;
                  }
                  m = -INFINITY, den = 0.f, acc = 0.f, d = dc;
                }
                const float mn = fmaxf(m, a);
                const float r = expf(m - mn), p = expf(a - mn);
                den = den * r + p;
                acc = acc * r + p * v;
                m = mn;
              } else {
                if (dc != d) {
                  const int64_t o = ((int64_t)dc * H + h) * kEF + c;
                  m = __ldg(g.smax + o);
                  den = 1.f / (__ldg(g.sden + o) + g.eps);
                  acc = __ldg(g.out + o);
                  gg = __ldg(g.g_out + o);
                  d = dc;
                }
                const float alpha = expf(a - m) * den;
                const int64_t o = ((int64_t)(e0 + t) * H + h) * kEF + c;
                g.d_msg[o] = alpha * gg;
                g.d_gate[o] = alpha * (v - acc) * gg;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[hb]);
        if (kMode == 0) {
          if (last_tile) {
            if (d >= 0) {
              const int64_t o = ((int64_t)d * H + h) * kEF + c;
              g.out[o] = acc / (den + g.eps);
              if (g.smax) g.smax[o] = m, g.sden[o] = den;
            }
          } else {
            cs[c] = m, cs[kEF + c] = den, cs[2 * kEF + c] = acc, cs[3 * kEF + c] = __int_as_float(d);
          }
        }
      }
    }
  } else if (warp < 12) {
    // ---------------------------------------------------------------- producers
    const int pt = tid - 128;
    uint32_t cnt = 0;
    const uint8_t* w2[2] = {reinterpret_cast<const uint8_t*>(g.w2a), reinterpret_cast<const uint8_t*>(g.w2m)};
    const int64_t ldp = 4 * (int64_t)hhd, ldt = 2 * (int64_t)hhd;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int e0 = e_lo + tile * kET;
      const int nv = min(kET, e_hi - e0);
      int32_t* mt = meta + (tile & (kEMetaBufs - 1)) * 3 * kET;
      if (pt < kET) {
        const bool ok = pt < nv;
        mt[pt] = ok ? g.dst[e0 + pt] : -1;
        mt[kET + pt] = ok ? g.src[e0 + pt] : 0;
        mt[2 * kET + pt] = ok ? g.rank[e0 + pt] : 0;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEProducers) : "memory");
      // this thread's 4 (edge row, 16-byte chunk) slots of every stage
      int rd[4], rs[4], rr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = (pt + kEProducers * j) >> 3;
        rd[j] = mt[r], rs[j] = mt[kET + r], rr[j] = mt[2 * kET + r];
      }
      for (int h = 0; h < H; ++h) {
        for (int net = 0; net < 2; ++net) {
          for (int kc = 0; kc < kcn; ++kc, ++cnt) {
            const uint32_t s = cnt % kEStages, u = cnt / kEStages;
            mbar_wait(&empty[s], (u + 1) & 1u);
            uint8_t* st = stages + s * kEStageBytes;
            if (pt == 0) {
              mbar_expect_tx(&full[s], kPackStageBytes);
              bulk_g2s(st, w2[net] + ((int64_t)h * kcn + kc) * kPackStageBytes, kPackStageBytes, &full[s]);
            }
            float4 pd[4], ps[4], te[4];
            const int col0 = h * hd + kc * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int cch = (pt + kEProducers * j) & 7;
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
              const int idx = pt + kEProducers * j;
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
                if ((lane & 7) == 0 && r < nv)
                  g.signs[((int64_t)(net * H + h) * kcn + kc) * g.n_edges + e0 + r] = word;
              }
              x.x = lrelu(x.x), x.y = lrelu(x.y), x.z = lrelu(x.z), x.w = lrelu(x.w);
              float4 hi, lo;
              split_tf32(x, hi, lo);
              const uint32_t off = sw128_offset(idx >> 3, idx & 7);
              *reinterpret_cast<float4*>(bh + off) = hi;
              *reinterpret_cast<float4*>(bh + kPackImageBytes + off) = lo;
            }
            fence_async_smem();
            mbar_arrive(&full[s]);
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(kEF, kET);
    uint32_t cnt = 0, hcount = 0;
    for (int tile = 0; tile < n_tiles; ++tile) {
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
            if (lane == 0) {
              const uint32_t a_hi = smem_u32(stages + s * kEStageBytes), a_lo = a_hi + kPackImageBytes;
              const uint32_t b_hi = a_hi + kPackStageBytes, b_lo = b_hi + kPackImageBytes;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t off = ks * 32;
                umma_tf32(d, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, (kc | ks) != 0);
                umma_tf32(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc, 1);
                umma_tf32(d, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc, 1);
              }
              umma_commit(&empty[s]);
              if (net == 1 && kc == kcn - 1) umma_commit(&tmem_full[hb]);
            }
            __syncwarp();
          }
        }
      }
    }
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

namespace {
int check_edge_args(int64_t n_atoms, int64_t n_edges, int32_t heads, int32_t f, int32_t hd) {
  if (f != kEF) return fail(-2, "cgat_edge_attn_*: only F = 128 (vector attention) is instantiated");
  if (heads < 1 || heads > kEMaxHeads) return fail(-2, "cgat_edge_attn_*: heads must be in [1,8]");
  if (hd <= 0 || (hd & 3)) return fail(-2, "cgat_edge_attn_*: hidden width must be a multiple of 4");
  if (n_atoms >= (1ll << 31) - 1 || n_edges >= (1ll << 31) - 129) return fail(-2, "cgat_edge_attn_*: size overflow");
  return 0;
}

template <int kMode>
int launch_edge(const EdgeArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(edge_attn_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, kESmemBytes));
    configured = true;
  }
  const int64_t tiles = ceil_div(a.n_edges, kET);
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  edge_attn_kernel<kMode><<<grid, kEThreads, kESmemBytes, stream>>>(a);
  return check_launch(kMode == 0 ? "edge_attn_fwd_kernel" : "edge_attn_bwd_prep_kernel");
}
}  // namespace

extern "C" int cgat_edge_attn_fwd(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                                  const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                                  const float* w2m_packed, const float* b2a, const float* b2m, float* out,
                                  float* seg_max, float* seg_den, int64_t n_atoms, int64_t n_edges, int32_t heads,
                                  int32_t f, int32_t hd, float eps, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_edge_args(n_atoms, n_edges, heads, f, hd)) return e;
  if (n_atoms <= 0) return 0;
  const size_t out_bytes = sizeof(float) * (size_t)n_atoms * heads * f;
  CGAT_CUDA(cudaMemsetAsync(out, 0, out_bytes, stream));  // atoms without in-edges aggregate to 0
  if (seg_max) {
    CGAT_CUDA(cudaMemsetAsync(seg_max, 0, out_bytes, stream));
    CGAT_CUDA(cudaMemsetAsync(seg_den, 0, out_bytes, stream));
  }
  if (n_edges <= 0) return 0;
  EdgeArgs a{P, T, rowptr, src, dst, rank, w2a_packed, w2m_packed, b2a, b2m, out, seg_max, seg_den,
             nullptr, nullptr, nullptr, nullptr, (int)n_atoms, (int)n_edges, heads, hd, eps};
  return launch_edge<0>(a, stream);
}

// Backward, step 1: per-edge gradients of the second-layer outputs.  Recomputes a_t,h / v_t,h exactly like
// the forward and writes d_gate, d_msg (E, H, F) in destination-sorted edge order plus the LeakyReLU sign
// masks signs[2][H][ceil(hd/32)][E] that the dgrad kernel needs (32 hidden units per word).
extern "C" int cgat_edge_attn_bwd_prep(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                                       const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                                       const float* w2m_packed, const float* b2a, const float* b2m, const float* out,
                                       const float* seg_max, const float* seg_den, const float* g_out, float* d_gate,
                                       float* d_msg, uint32_t* signs, int64_t n_atoms, int64_t n_edges, int32_t heads,
                                       int32_t f, int32_t hd, float eps, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_edge_args(n_atoms, n_edges, heads, f, hd)) return e;
  if (n_atoms <= 0 || n_edges <= 0) return 0;
  EdgeArgs a{P, T, rowptr, src, dst, rank, w2a_packed, w2m_packed, b2a, b2m, const_cast<float*>(out),
             const_cast<float*>(seg_max), const_cast<float*>(seg_den), g_out, d_gate, d_msg, signs,
             (int)n_atoms, (int)n_edges, heads, hd, eps};
  return launch_edge<1>(a, stream);
}
