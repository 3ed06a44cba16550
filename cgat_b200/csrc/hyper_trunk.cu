// Hypernetwork trunks, fused (reference CGAT/Hypernetworksmp.py:36-83 FCBlock,
// :243-254 HyperLinear.forward).
//
// Every HyperLinear of a node layer owns a small MLP ("trunk") that turns the per-atom hyper-input h into
// the code z that its big last Linear expands into a predicted weight matrix:
//     t1 = tanh(W1 h + b1), t2 = tanh(W2 t1 + b2), t3 = tanh(W3 t2 + b3), z = t4 = tanh(W4 t3 + b4)
// and the bias-shaped tail of that last Linear is one more product of the same shape:
//     e = We z + be          (We = last.weight[F*F:], be = last.bias[F*F:])
// A node layer has J = 4 HyperLinears with the SAME input h.  As written in the reference that is
// 4 x 5 cuBLAS calls + 16 tanh launches forward and ~4x that backward, each on an (N x 128) x (128 x 128)
// problem that cannot fill the GPU.  Here one persistent kernel runs the whole chain for a (128-atom tile,
// hyper-layer j) work item: the activation tile never leaves the SM between layers — the epilogue threads
// (one atom row each) read the accumulator from TMEM, apply bias + tanh, store the fp32 row (saved for
// backward) and write it back into shared memory, already split into tf32 hi/lo and swizzled, as the A operand
// of the next layer's MMAs.  Weights stream through a cp.async.bulk ring from a pre-packed image.
//
// The backward chain has the same shape and runs in the same kernel (kMode 1):
//     d4 = (dE We + dZ) * (1 - z^2),  d3 = (d4 W4) * (1 - t3^2),  ...,  d1 = (d2 W2) * (1 - t1^2),  dh_j = d1 W1
// The pre-activation gradients d_i are stored for the weight gradients (cgat_gemm3x_tn_batched).
//
// Roles (416 threads, one persistent CTA per SM):
//   warps 0-7   epilogue (thread = atom row x column half): tcgen05.ld -> bias/tanh (or tanh') -> global + next A operand
//   warps 8-11  stage the chain input tile (h, or dE_j); warp 8 lane 0 also streams the packed weights
//   warp  12    TMEM allocation + MMA issue (3xTF32, one accumulator per K chunk, see below)
#include "common.cuh"
#include "tc_common.cuh"

namespace cgat {
namespace {
using namespace tc;

constexpr int kTF = 128;           // trunk width (instantiated for F = 128)
constexpr int kTMaxJ = 4;          // hyper-layers per call
constexpr int kTSteps = 5;         // GEMMs per chain
constexpr int kTKC = kTF / kPackChunk;
constexpr int kTABytes = kTKC * (int)kPackStageBytes;  // 128 KB: activation tile, hi+lo
constexpr int kTStages = 3;
constexpr int kTSmemBytes = kTABytes + kTStages * (int)kPackStageBytes + 1024 + 512;
constexpr int kTEpiWarps = 8;       // two column halves x four lane quadrants (16 warps: 80-register cap, measured slower)
constexpr int kTMmaWarp = kTEpiWarps + 4;
constexpr int kTThreads = (kTMmaWarp + 1) * 32;
constexpr int64_t kTMatFloats = (int64_t)kTKC * (kPackStageBytes / 4);  // one packed F x F matrix

struct TrunkArgs {
  const float* x0;        // fwd: h (N,F); bwd: dE (J,N,F)
  const float* w_packed;  // [J][5][kKC][hi|lo][16 KB]; fwd order W1..W4,We; bwd order We^T,W4^T..W1^T
  const float* bias[kTMaxJ * kTSteps];  // fwd only
  float* T;               // (J,4,N,F) tanh outputs: fwd written, bwd read
  float* E;               // (J,N,F)   fwd: e
  const float* dZ;        // (J,N,F)   bwd
  float* D;               // (J,4,N,F) bwd: pre-activation gradients d1..d4
  float* dH;              // (J,N,F)   bwd: per hyper-layer dL/dh
  int n_atoms, J;
};

template <int kMode>
__global__ void __launch_bounds__(kTThreads, 1) hyper_trunk_kernel(const TrunkArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                      // [kKC][hi|lo][16 KB]
  uint8_t* b_smem = smem + kTABytes;           // [kStages][hi|lo][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + kTStages * kPackStageBytes);
  uint64_t* full = bars;                       // [kStages] TMA -> MMA
  uint64_t* empty = bars + kTStages;           // [kStages] MMA -> TMA
  uint64_t* a_ready = bars + 2 * kTStages;     // stagers / epilogue -> MMA: A operand of the next step is in place
  uint64_t* acc_full = a_ready + 1;            // MMA -> epilogue
  uint64_t* a_free = acc_full + 1;             // MMA -> stagers: the item's last MMAs have read the A tile
  uint64_t* tmem_free = a_free + 1;            // epilogue -> MMA: the item's last accumulator has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_free + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = g.n_atoms, J = g.J;
  const int n_tiles = (N + 127) / 128;
  const int n_items = n_tiles * J;
  // item = (atom tile, hyper-layer j), j fastest; contiguous ranges per CTA keep the input tile hot in L2
  const int item_lo = (int)((int64_t)n_items * blockIdx.x / gridDim.x);
  const int item_hi = (int)((int64_t)n_items * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(a_ready, 128);
    mbar_init(acc_full, 1);
    mbar_init(a_free, 1);
    mbar_init(tmem_free, 128);
    mbar_init_fence();
  }
  if (warp == kTMmaWarp) tmem_alloc(tmem_slot, 4 * kTF);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < kTEpiWarps) {
    // ------------------------------------------------------------------ epilogue
    // 8 warps: warp -> (lane quadrant qd, column half ch).  A chain step is strictly serial (MMA -> this epilogue ->
    // next MMA), so the epilogue's latency is the step's: with 4 warps (one per scheduler, 128 columns each) it took
    // 7 of the 8.7 us of a forward step (in-graph profile profiles/r03m: 87 us for two 5-step chains).
    const int qd = warp & 3, ch = warp >> 2;
    const int r = qd * 32 + lane;  // row of the tile = TMEM lane
    uint32_t k = 0;                  // global step counter (phase of acc_full)
    for (int item = item_lo; item < item_hi; ++item) {
      const int tile = item / J, j = item - tile * J;
      const int n = tile * 128 + r;
      const bool valid = n < N;
      for (int s = 0; s < kTSteps; ++s, ++k) {
        mbar_wait(acc_full, k & 1u);
        tc_fence_after();
        const bool last = (s == kTSteps - 1);
        // where this step's result goes / which saved activation it needs
        float* dst;
        const float* tt = nullptr;   // bwd: tanh output whose derivative scales this step
        const float* add = nullptr;  // bwd step 0: + dZ
        const float* bias = nullptr;
        if (kMode == 0) {
          dst = last ? g.E + ((int64_t)j * N + n) * kTF : g.T + (((int64_t)j * 4 + s) * N + n) * kTF;
          bias = g.bias[j * kTSteps + s];
        } else {
          dst = last ? g.dH + ((int64_t)j * N + n) * kTF : g.D + (((int64_t)j * 4 + (3 - s)) * N + n) * kTF;
          if (!last) tt = g.T + (((int64_t)j * 4 + (3 - s)) * N + n) * kTF;
          if (s == 0) add = g.dZ + ((int64_t)j * N + n) * kTF;
        }
#pragma unroll 1
        for (int c16 = ch * (kTF / 16 / (kTEpiWarps / 4)); c16 < (ch + 1) * (kTF / 16 / (kTEpiWarps / 4)); ++c16) {
          const int cc = c16 >> 1;  // K chunk of the next step this column group belongs to
          float v0[16], v1[16], v2[16], v3[16];
          const uint32_t tb = tmem + ((uint32_t)(qd * 32) << 16) + c16 * 16;
          tmem_ld16(tb, v0);
          tmem_ld16(tb + kTF, v1);
          tmem_ld16(tb + 2 * kTF, v2);
          tmem_ld16(tb + 3 * kTF, v3);
          tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int q = (c16 & 1) * 4 + q4;  // 16-byte chunk inside the 32-float K chunk
            float4 x = make_float4((v0[4 * q4] + v1[4 * q4]) + (v2[4 * q4] + v3[4 * q4]),
                                   (v0[4 * q4 + 1] + v1[4 * q4 + 1]) + (v2[4 * q4 + 1] + v3[4 * q4 + 1]),
                                   (v0[4 * q4 + 2] + v1[4 * q4 + 2]) + (v2[4 * q4 + 2] + v3[4 * q4 + 2]),
                                   (v0[4 * q4 + 3] + v1[4 * q4 + 3]) + (v2[4 * q4 + 3] + v3[4 * q4 + 3]));
            if (kMode == 0) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + cc * 32) + q);
              x.x += b.x, x.y += b.y, x.z += b.z, x.w += b.w;
              if (!last) x.x = tanhf(x.x), x.y = tanhf(x.y), x.z = tanhf(x.z), x.w = tanhf(x.w);
            } else {
              if (add != nullptr && valid) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(add + cc * 32) + q);
                x.x += a.x, x.y += a.y, x.z += a.z, x.w += a.w;
              }
              if (!last) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) t = __ldg(reinterpret_cast<const float4*>(tt + cc * 32) + q);
                x.x *= 1.f - t.x * t.x, x.y *= 1.f - t.y * t.y, x.z *= 1.f - t.z * t.z, x.w *= 1.f - t.w * t.w;
              }
            }
            if (!valid) x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) reinterpret_cast<float4*>(dst + cc * 32)[q] = x;
            if (!last) {  // next step's A operand: K chunk cc, row r, 16-byte chunk q
              float4 hi, lo;
              split_tf32(x, hi, lo);
              const uint32_t off = sw128_offset(r, q);
              *reinterpret_cast<float4*>(a_smem + cc * kPackStageBytes + off) = hi;
              *reinterpret_cast<float4*>(a_smem + cc * kPackStageBytes + kPackImageBytes + off) = lo;
            }
          }
        }
        tc_fence_before();
        if (!last) fence_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(kTEpiWarps * 32) : "memory");   // both column halves are done
        if (ch == 0) {   // the barriers count 128 arrivals (the stagers use a_ready too)
          if (!last) mbar_arrive(a_ready);
          else mbar_arrive(tmem_free);
        }
      }
    }
  } else if (warp < kTMmaWarp) {
    // ------------------------------------------------------------------ chain-input stagers + weight TMA
    const int st = tid - kTEpiWarps * 32;
    uint32_t it = 0, cnt = 0;
    for (int item = item_lo; item < item_hi; ++item, ++it) {
      const int tile = item / J, j = item - tile * J;
      const float* x0 = g.x0 + (kMode == 1 ? (int64_t)j * N * kTF : 0);
      mbar_wait(a_free, (it + 1) & 1u);  // the previous item's last MMAs have read the A tile
#pragma unroll 1
      for (int kc = 0; kc < kTKC; ++kc) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int idx = st + 128 * q, r = idx >> 3, c = idx & 7;
          const int gr = tile * 128 + r;
          v[q] = gr < N ? __ldg(reinterpret_cast<const float4*>(x0 + (int64_t)gr * kTF + kc * 32 + c * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint8_t* hi = a_smem + kc * kPackStageBytes;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int idx = st + 128 * q;
          const uint32_t off = sw128_offset(idx >> 3, idx & 7);
          float4 h, l;
          split_tf32(v[q], h, l);
          *reinterpret_cast<float4*>(hi + off) = h;
          *reinterpret_cast<float4*>(hi + kPackImageBytes + off) = l;
        }
      }
      fence_async_smem();
      mbar_arrive(a_ready);
      if (st == 0) {
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(g.w_packed) + (int64_t)j * kTSteps * kTKC * kPackStageBytes;
        for (int sk = 0; sk < kTSteps * kTKC; ++sk, ++cnt) {
          const uint32_t s = cnt % kTStages, u = cnt / kTStages;
          mbar_wait(&empty[s], (u + 1) & 1u);
          mbar_arrive_expect_tx(&full[s], kPackStageBytes);
          bulk_g2s(b_smem + s * kPackStageBytes, wsrc + (int64_t)sk * kPackStageBytes, kPackStageBytes, &full[s]);
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(128, kTF);
    uint32_t it = 0, cnt = 0, k = 0;
    for (int item = item_lo; item < item_hi; ++item, ++it) {
      mbar_wait(tmem_free, (it + 1) & 1u);  // the previous item's last accumulator has been drained
      for (int s = 0; s < kTSteps; ++s, ++k) {
        mbar_wait(a_ready, k & 1u);
        tc_fence_after();
        for (int kc = 0; kc < kTKC; ++kc, ++cnt) {
          const uint32_t sg = cnt % kTStages, u = cnt / kTStages;
          mbar_wait(&full[sg], u & 1u);
          tc_fence_after();
          {  // all lanes, warp-uniform operands; one elected lane issues (tc_common.cuh "_e" forms)
            const uint32_t a_hi = smem_u32(a_smem + kc * kPackStageBytes), a_lo = a_hi + kPackImageBytes;
            const uint32_t b_hi = smem_u32(b_smem + sg * kPackStageBytes), b_lo = b_hi + kPackImageBytes;
            // The tensor core truncates (rounds toward zero) the fp32 accumulator on every accumulating MMA, a
            // coherent shrink of ~3.6e-7 for a 16-step K = 128 sum that the hypernetwork amplifies into gradient
            // noise.  So the hi*hi products of K chunk kc go to their OWN accumulator (3 truncations at a quarter
            // of the magnitude); the 2^-11-sized correction products of all chunks share the last one, which
            // receives its hi*hi products (chunk 3) after them.  The epilogue adds the four with round-to-nearest.
            const uint32_t d_main = tmem + kc * kTF, d_corr = tmem + 3 * kTF;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t off = ks * 32;
              umma_tf32_e(d_corr, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, (kc | ks) != 0);
              umma_tf32_e(d_corr, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc, 1);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t off = ks * 32;
              umma_tf32_e(d_main, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc,
                        (kc == kTKC - 1) ? 1u : (uint32_t)(ks != 0));
            }
            umma_commit_e(&empty[sg]);
            if (kc == kTKC - 1) {
              umma_commit_e(acc_full);
              if (s == kTSteps - 1) umma_commit_e(a_free);
            }
          }
          __syncwarp();
        }
      }
    }
  }
  __syncthreads();
  if (warp == kTMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 4 * kTF);
  }
}

// ---- packing of a list of F x F matrices into chain operands ------------------------------------
constexpr int kPackMaxMats = 24;
struct PackListArgs {
  const float* w[kPackMaxMats];
  int64_t ld[kPackMaxMats];
  int transpose;
};

// grid: (kKC, n_mats); thread -> (row r = idx/8, 16-byte chunk c = idx%8) of a 128 x 32 chunk
__global__ void trunk_pack_kernel(const PackListArgs a, float* __restrict__ out) {
  const int kc = blockIdx.x, m = blockIdx.y;
  const float* w = a.w[m];
  const int64_t ld = a.ld[m];
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + ((int64_t)m * kTKC + kc) * kPackStageBytes;
  for (int idx = threadIdx.x; idx < kPackRows * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    const int gk = kc * kPackChunk + c * 4;
    float4 x;
    if (a.transpose == 0) {
      x = __ldg(reinterpret_cast<const float4*>(w + (int64_t)r * ld + gk));
    } else {
      x.x = __ldg(w + (int64_t)gk * ld + r), x.y = __ldg(w + (int64_t)(gk + 1) * ld + r);
      x.z = __ldg(w + (int64_t)(gk + 2) * ld + r), x.w = __ldg(w + (int64_t)(gk + 3) * ld + r);
    }
    float4 hi, lo;
    split_tf32(x, hi, lo);
    const uint32_t off = sw128_offset(r, c);
    *reinterpret_cast<float4*>(dst + off) = hi;
    *reinterpret_cast<float4*>(dst + kPackImageBytes + off) = lo;
  }
}

template <int kMode>
int launch_trunk(const TrunkArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CGAT_CUDA(cudaFuncSetAttribute(hyper_trunk_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmemBytes));
    configured = true;
  }
  const int n_items = ((a.n_atoms + 127) / 128) * a.J;
  const int grid = n_items < kNumSMs ? n_items : kNumSMs;
  hyper_trunk_kernel<kMode><<<grid, kTThreads, kTSmemBytes, stream>>>(a);
  return check_launch(kMode == 0 ? "hyper_trunk_fwd_kernel" : "hyper_trunk_bwd_kernel");
}

int check_trunk(int64_t n_atoms, int32_t f, int32_t J) {
  if (f != kTF) return fail(-2, "cgat_hyper_trunk_*: only F = 128 is instantiated");
  if (J < 1 || J > kTMaxJ) return fail(-2, "cgat_hyper_trunk_*: 1..4 hyper-layers per call");
  if (n_atoms >= (1ll << 31) / (kTF * 4 * kTMaxJ)) return fail(-2, "cgat_hyper_trunk_*: too many atoms");
  return 0;
}

}  // namespace
}  // namespace cgat

using namespace cgat;

extern "C" int64_t cgat_hyper_trunk_packed_floats(int32_t n_mats, int32_t f) {
  return f == kTF ? (int64_t)n_mats * kTMatFloats : 0;
}

// weights: HOST array of n_mats device pointers to F x F row-major matrices (leading dimensions ld[i]);
// out: n_mats packed images (cgat_hyper_trunk_packed_floats).  transpose = 1 packs W^T (backward chain).
extern "C" int cgat_hyper_trunk_pack(const float* const* weights, const int64_t* ld, int32_t n_mats, int32_t f,
                                     int32_t transpose, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (f != kTF) return fail(-2, "cgat_hyper_trunk_pack: only F = 128 is instantiated");
  if (n_mats < 1 || n_mats > kPackMaxMats) return fail(-2, "cgat_hyper_trunk_pack: 1..24 matrices per call");
  PackListArgs a;
  for (int i = 0; i < n_mats; ++i) {
    if ((ld[i] & 3) || (reinterpret_cast<uintptr_t>(weights[i]) & 15))
      return fail(-2, "cgat_hyper_trunk_pack: matrices must be 16-byte aligned with ld % 4 == 0");
    a.w[i] = weights[i], a.ld[i] = ld[i];
  }
  a.transpose = transpose;
  trunk_pack_kernel<<<dim3(kTKC, n_mats), 256, 0, stream>>>(a, out);
  return check_launch("trunk_pack_kernel");
}

// T[j][s] = tanh(W_{j,s+1} T[j][s-1] + b_{j,s+1}) (T[j][-1] = h), s = 0..3;  E[j] = We_j T[j][3] + be_j.
//   w_packed: cgat_hyper_trunk_pack(transpose=0) of [j][W1,W2,W3,W4,We];  biases: HOST array of J*5 device pointers.
extern "C" int cgat_hyper_trunk_fwd(const float* h, const float* w_packed, const float* const* biases, float* T,
                                    float* E, int64_t n_atoms, int32_t f, int32_t J, void* stream_) {
  if (int e = check_trunk(n_atoms, f, J)) return e;
  if (n_atoms <= 0) return 0;
  TrunkArgs a{};
  a.x0 = h, a.w_packed = w_packed, a.T = T, a.E = E, a.n_atoms = (int)n_atoms, a.J = J;
  for (int i = 0; i < J * kTSteps; ++i) a.bias[i] = biases[i];
  return launch_trunk<0>(a, (cudaStream_t)stream_);
}

// D[j][3] = (dE[j] We_j + dZ[j]) * (1 - T[j][3]^2);  D[j][s-1] = (D[j][s] W_{j,s+1}) * (1 - T[j][s-1]^2);
// dH[j] = D[j][0] W_{j,1}.   wt_packed: cgat_hyper_trunk_pack(transpose=1) of [j][We,W4,W3,W2,W1].
extern "C" int cgat_hyper_trunk_bwd(const float* dE, const float* dZ, const float* T, const float* wt_packed, float* D,
                                    float* dH, int64_t n_atoms, int32_t f, int32_t J, void* stream_) {
  if (int e = check_trunk(n_atoms, f, J)) return e;
  if (n_atoms <= 0) return 0;
  TrunkArgs a{};
  a.x0 = dE, a.w_packed = wt_packed, a.T = const_cast<float*>(T), a.dZ = dZ, a.D = D, a.dH = dH;
  a.n_atoms = (int)n_atoms, a.J = J;
  return launch_trunk<1>(a, (cudaStream_t)stream_);
}
