// Train-step glue kernels (VERDICT r01 missing #4, next #8) — all HBM-bound elementwise /
// reduction work that round 1 left to library kernels:
//
//   cgat_sum_parts    out = [out +] sum_p parts[p]   fixed order (deterministic).  The split-K / split-atom partial
//                     results of the tensor-core kernels (hyper_rowscale, hyper_wgrad, edge wgrad, gemm3x_tn...) used
//                     to be summed by torch.sum: 176 `reduce_kernel` launches per cfg2 step at ~12 us each (2.1 ms).
//   cgat_adamw_flat   AdamW over ONE flat fp32 parameter / gradient / moment buffer (reference
//                     CGAT/lightning_module.py:328-344 builds torch.optim.AdamW; default --optim AdamW, lr 1.25e-4,
//                     wd 1e-6): 16 B read + 12 B written per parameter in one launch; learning rate and step count
//                     live in device memory so that a captured CUDA graph replays with the current values; the DDP
//                     average (1 / world size) is folded into the gradient read.
//   cgat_l1_loss      L1 loss on column 0 of the prediction (reference lightning_module.py:237-240, nn.L1Loss on
//                     output vs normalised target) and its gradient w.r.t. the (C, 2) prediction in one launch.
#include "common.cuh"

namespace cgat {
namespace {

constexpr int kSumThreads = 256;

template <bool kAcc>
__global__ void __launch_bounds__(kSumThreads) sum_parts_kernel(const float* __restrict__ parts, int n_parts,
                                                                int64_t stride, float* __restrict__ out, int64_t n4) {
  // one float4 per thread and iteration; the parts of one element are 16-byte loads `stride` floats apart, all issued
  // before the adds (n_parts <= 32 in this library: up to 32 independent loads in flight per thread)
  for (int64_t i = (int64_t)blockIdx.x * kSumThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kSumThreads) {
    float4 acc = kAcc ? reinterpret_cast<const float4*>(out)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = reinterpret_cast<const float4*>(parts) + i;
    int p = 0;
    for (; p + 4 <= n_parts; p += 4) {
      const float4 a = __ldg(src + (int64_t)p * (stride >> 2)), b = __ldg(src + (int64_t)(p + 1) * (stride >> 2));
      const float4 c = __ldg(src + (int64_t)(p + 2) * (stride >> 2)), d = __ldg(src + (int64_t)(p + 3) * (stride >> 2));
      acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
      acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
      acc.x += c.x, acc.y += c.y, acc.z += c.z, acc.w += c.w;
      acc.x += d.x, acc.y += d.y, acc.z += d.z, acc.w += d.w;
    }
    for (; p < n_parts; ++p) {
      const float4 a = __ldg(src + (int64_t)p * (stride >> 2));
      acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

template <bool kAcc>
__global__ void __launch_bounds__(kSumThreads) sum_parts_scalar_kernel(const float* __restrict__ parts, int n_parts,
                                                                       int64_t stride, float* __restrict__ out,
                                                                       int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * kSumThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSumThreads) {
    float acc = kAcc ? out[i] : 0.f;
    for (int p = 0; p < n_parts; ++p) acc += __ldg(parts + (int64_t)p * stride + i);
    out[i] = acc;
  }
}

// out[m, n] = act(sum_p parts[p][m, n] + bias[n]): the epilogue of a split-K GEMM whose partial tiles were written by
// cgat_gemm3x_nt_splitk (bias / activation cannot be applied per part).  n_cols % 4 == 0.
__global__ void __launch_bounds__(kSumThreads) sum_parts_bias_act_kernel(const float* __restrict__ parts, int n_parts,
                                                                         int64_t stride, const float* __restrict__ bias,
                                                                         float* __restrict__ out, int64_t n4,
                                                                         int n_cols4, int act) {
  for (int64_t i = (int64_t)blockIdx.x * kSumThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kSumThreads) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = reinterpret_cast<const float4*>(parts) + i;
    for (int p = 0; p < n_parts; ++p) {
      const float4 a = __ldg(src + (int64_t)p * (stride >> 2));
      acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
    }
    if (bias != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (int)(i % n_cols4));
      acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
    }
    float* a = reinterpret_cast<float*>(&acc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = a[j];
      if (act == 1) v = v > 0.f ? v : 0.01f * v;
      else if (act == 2) v = tanhf(v);
      else if (act == 3) v = fmaxf(v, 0.f);
      a[j] = v;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

__global__ void adamw_tick_kernel(float* __restrict__ step) { step[0] += 1.f; }

constexpr int kAdamThreads = 256;

__global__ void __launch_bounds__(kAdamThreads) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                                  float* __restrict__ m, float* __restrict__ v,
                                                                  int64_t n4, const float* __restrict__ lr_ptr,
                                                                  const float* __restrict__ step_ptr, float beta1,
                                                                  float beta2, float eps, float wd, float grad_scale) {
  // torch.optim.AdamW (decoupled weight decay): p *= 1 - lr wd; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
  // p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
  const float lr = __ldg(lr_ptr), t = __ldg(step_ptr);
  const float bc1 = 1.f - powf(beta1, t), bc2s = sqrtf(1.f - powf(beta2, t));
  const float step_size = lr / bc1, decay = 1.f - lr * wd;
  for (int64_t i = (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kAdamThreads) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    float* ga = reinterpret_cast<float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * grad_scale;
      ma[j] = beta1 * ma[j] + (1.f - beta1) * gr;
      va[j] = beta2 * va[j] + (1.f - beta2) * gr * gr;
      const float denom = sqrtf(va[j]) / bc2s + eps;
      pa[j] = pa[j] * decay - step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

// One CTA: n is the number of crystals of a batch (hundreds to thousands).  Fixed-order tree reduction.
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ out, int64_t ldo,
                                                      const float* __restrict__ target, int n,
                                                      float* __restrict__ loss, float* __restrict__ grad, int64_t ldg,
                                                      int n_rows, int n_cols) {
  __shared__ float red[256];
  float acc = 0.f;
  const float inv = 1.f / (float)n;
  for (int i = threadIdx.x; i < n_rows; i += 256) {
    float gi = 0.f;
    if (i < n) {
      const float d = out[(int64_t)i * ldo] - target[i];
      acc += fabsf(d);
      gi = d > 0.f ? inv : (d < 0.f ? -inv : 0.f);
    }
    if (grad != nullptr) {
      grad[(int64_t)i * ldg] = gi;  // rows >= n (padding crystals) and the other columns get zero gradient
      for (int c = 1; c < n_cols; ++c) grad[(int64_t)i * ldg + c] = 0.f;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] * inv;
}

}  // namespace
}  // namespace cgat

using namespace cgat;

// out[i] = (accumulate ? out[i] : 0) + sum_{p < n_parts} parts[p * part_stride + i], i < n; parts summed in index order.
extern "C" int cgat_sum_parts(const float* parts, int32_t n_parts, int64_t part_stride, float* out, int64_t n,
                              int32_t accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0) return 0;
  if (n_parts < 0) return fail(-2, "cgat_sum_parts: n_parts must be >= 0");
  const bool vec = !(n & 3) && !(part_stride & 3) && !(reinterpret_cast<uintptr_t>(parts) & 15) &&
                   !(reinterpret_cast<uintptr_t>(out) & 15);
  const int64_t work = vec ? n / 4 : n;
  int64_t blocks = ceil_div(work, kSumThreads);
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (vec) {
    if (accumulate) sum_parts_kernel<true><<<(unsigned)blocks, kSumThreads, 0, stream>>>(parts, n_parts, part_stride, out, work);
    else sum_parts_kernel<false><<<(unsigned)blocks, kSumThreads, 0, stream>>>(parts, n_parts, part_stride, out, work);
  } else {
    if (accumulate) sum_parts_scalar_kernel<true><<<(unsigned)blocks, kSumThreads, 0, stream>>>(parts, n_parts, part_stride, out, n);
    else sum_parts_scalar_kernel<false><<<(unsigned)blocks, kSumThreads, 0, stream>>>(parts, n_parts, part_stride, out, n);
  }
  return check_launch("sum_parts_kernel");
}

// out (M, N) = act(sum_p parts[p] + bias[N]) for contiguous (M, N) parts `part_stride` floats apart; N % 4 == 0.
// act: 0 none, 1 LeakyReLU(0.01), 2 tanh, 3 ReLU (as cgat_gemm3x_nt).
extern "C" int cgat_sum_parts_bias_act(const float* parts, int32_t n_parts, int64_t part_stride, const float* bias,
                                       float* out, int64_t M, int64_t N, int32_t act, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M <= 0 || N <= 0) return 0;
  if ((N & 3) || (part_stride & 3) || n_parts < 1 ||
      ((reinterpret_cast<uintptr_t>(parts) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15))
    return fail(-2, "cgat_sum_parts_bias_act: N and part_stride must be multiples of 4, buffers 16-byte aligned");
  const int64_t work = M * N / 4;
  int64_t blocks = ceil_div(work, kSumThreads);
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  sum_parts_bias_act_kernel<<<(unsigned)blocks, kSumThreads, 0, stream>>>(parts, n_parts, part_stride, bias, out, work,
                                                                         (int)(N / 4), act);
  return check_launch("sum_parts_bias_act_kernel");
}

// One AdamW step over flat buffers of n floats (n % 4 == 0, 16-byte aligned).  `step` (device float) is incremented
// first and then read as t; `lr` is a device float.  grad_scale multiplies the gradient on the way in (1 / world size).
extern "C" int cgat_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float* step,
                               float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0) return 0;
  if ((n & 3) || ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                   reinterpret_cast<uintptr_t>(v)) & 15))
    return fail(-2, "cgat_adamw_flat: n must be a multiple of 4 and the buffers 16-byte aligned");
  adamw_tick_kernel<<<1, 1, 0, stream>>>(step);
  if (int e = check_launch("adamw_tick_kernel")) return e;
  int64_t blocks = ceil_div(n / 4, kAdamThreads);
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  adamw_flat_kernel<<<(unsigned)blocks, kAdamThreads, 0, stream>>>(p, g, m, v, n / 4, lr, step, beta1, beta2, eps,
                                                                  weight_decay, grad_scale);
  return check_launch("adamw_flat_kernel");
}

// loss[0] = mean_{i < n} |out[i*ldo] - target[i]|;  grad (n_rows, n_cols; optional): d loss / d out, zero for rows >= n
// (padding crystals) and for columns >= 1.
extern "C" int cgat_l1_loss(const float* out, int64_t ldo, const float* target, int64_t n, float* loss, float* grad,
                            int64_t ldg, int64_t n_rows, int32_t n_cols, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0 || n_rows < n || n_rows >= (1ll << 31)) return fail(-2, "cgat_l1_loss: need 0 < n <= n_rows < 2^31");
  l1_loss_kernel<<<1, 256, 0, stream>>>(out, ldo, target, (int)n, loss, grad, ldg, (int)n_rows, n_cols);
  return check_launch("l1_loss_kernel");
}
