// Blackwell (sm_100a) tensor-core plumbing used by the cgat_b200 GEMM-shaped kernels:
// mbarrier, tcgen05.alloc/mma/commit/ld, UMMA shared-memory and instruction descriptors, and the
// fp32 -> (tf32_hi, tf32_lo) split that makes three kind::tf32 passes reproduce an fp32 product
// to ~2^-21 (the parity bar of BASELINE.json rules out a single TF32 pass).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace cgat {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the warp is parked by the hardware until the phase completes (or the hint
// expires) instead of re-issuing the poll.  In the edge kernels ~30 % of all executed instructions were mbarrier
// polls (BRA + SYNCS + YIELD, ncu r01h); waking is still event-driven, so no latency is added.  Measured neutral on
// B200 (27.5 ms cfg2 step either way: the polls were not what the working warps were short of); kept because it
// removes the instructions.  -DCGAT_MBAR_SPIN restores the plain spin.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
#ifdef CGAT_MBAR_TRAP
// Debug build (python -m cgat_b200.build --trap-barriers): every mbarrier wait is bounded.  After CGAT_MBAR_TRAP polls
// (each parks the warp for up to 20 us) the kernel prints which barrier of which block never completed and traps, so
// a protocol bug surfaces as "unspecified launch failure" + one line of text instead of a hung GPU.
static __device__ __noinline__ void mbar_timeout(const char* name, int line, uint32_t parity) {
  printf("cgat_b200: mbarrier wait timed out: %s (line %d), parity %u, block (%d,%d,%d), thread %d\n", name, line, parity,
         blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait_named(uint64_t* bar, uint32_t parity, const char* name, int line) {
  for (uint32_t polls = 0; !mbar_try_wait_hint(bar, parity, 20000u); ++polls)
    if (polls > (uint32_t)(CGAT_MBAR_TRAP)) mbar_timeout(name, line, parity);
}
#define mbar_wait(bar, parity) mbar_wait_named((bar), (parity), #bar, __LINE__)
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef CGAT_MBAR_SPIN
  while (!mbar_try_wait(bar, parity)) {
  }
#else
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
  }
#endif
}
#endif

// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tensor memory --------------------------------------------------------------------------
// One full warp allocates `cols` (power of two >= 32) TMEM columns; the base address lands in smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// 32 lanes x 32 columns of fp32 accumulator -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -------------------------------------------------------------------------
// K-major operand tile in the canonical SWIZZLE_128B layout: rows of 128 bytes (32 fp32 along K),
// 16-byte chunk c of row r stored at chunk (c ^ (r & 7)); groups of 8 rows are 1024 B apart (SBO).
// The tile base must be 1024-byte aligned.  (cute/arch/mma_sm100_desc.hpp: SmemDescriptor.)
constexpr uint32_t kSwizzleRowBytes = 128;
constexpr uint32_t kSwizzleAtomBytes = 1024;

__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(kSwizzleAtomBytes >> 4) << 32;          // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                 // layout type SWIZZLE_128B
  return d;
}

// MN-major operand (the contiguous dimension of the source is M or N, the contraction runs over rows).
// For 32-bit (tf32) elements the only swizzled MN-major layout is SWIZZLE_128B_BASE32B
// (cutlass/gemm/collective/builders/sm100_common.inl:92; cute Layout_MN_SW128_32B_Atom =
// Swizzle<2,5,2> o (1024 bits x 4)): a stage is a stack of images, one per block of 32 consecutive
// M/N elements; an image holds the K rows of that block, 128 bytes each; atoms are 4 rows (512 B) and
// inside an atom the 32-byte unit u of row r sits at unit (u ^ (r & 3)).
// LBO = byte distance between images (next 32 M/N elements), SBO = byte distance between 4-row K groups.
constexpr uint32_t kMnAtomBytes = 512;

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(kMnAtomBytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // layout type SWIZZLE_128B_BASE32B
  return d;
}
// byte offset, inside an image, of the 16-byte chunk `chunk16` (0..7) of K row r
__device__ __forceinline__ uint32_t mn_sw128_offset(uint32_t r, uint32_t chunk16) {
  return r * kSwizzleRowBytes + ((((chunk16 >> 1) ^ (r & 3u)) << 5) | ((chunk16 & 1u) << 4));
}

// byte offset of fp32 element (row r, k) of a K-major SW128 tile whose K extent is 32 floats
__device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t chunk16) {
  return r * kSwizzleRowBytes + ((chunk16 ^ (r & 7u)) << 4);
}

// kind::tf32, fp32 accumulate, both operands K-major (cute/arch/mma_sm100_desc.hpp: InstrDescriptor)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t m, uint32_t n, uint32_t a_mn_major = 0,
                                                       uint32_t b_mn_major = 0) {
  return (1u << 4)               // c_format  = F32
         | (2u << 7)             // a_format  = TF32
         | (2u << 10)            // b_format  = TF32
         | (a_mn_major << 15)    // A: 0 = K-major, 1 = MN-major
         | (b_mn_major << 16)    // B: 0 = K-major, 1 = MN-major
         | ((n >> 3) << 17)      // N / 8
         | ((m >> 4) << 24);     // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued UMMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32-byte store (sm_100: STG.256): a thread that owns a row of an accumulator tile writes whole 32-byte sectors,
// so row-per-thread epilogues do not leave half-written sectors for L2 to merge.  p must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// 32-byte read-only load (sm_100: LDG.256): 8 consecutive floats of a gathered row in one request, whole sectors.
// p must be 32-byte aligned.
__device__ __forceinline__ void ldg_v8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

// ---- fp32 -> tf32 hi/lo split -------------------------------------------------------------------
// hi = round-to-nearest TF32 of x, lo = round-to-nearest TF32 of (x - hi) (the subtraction is exact in
// fp32).  Both are exactly representable in TF32, so the tensor core's own operand conversion
// (truncation) is lossless and the residual is unbiased: a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with
// a relative error of ~2^-22 per product (the dropped lo*lo term is ~2^-24).
// Round to nearest, ties away from zero, on the bit pattern: add half a TF32 ulp to the magnitude and clear the 13
// low mantissa bits.  Same result as `cvt.rna.tf32.f32` for every finite input, but two integer instructions: ptxas
// expands the cvt into ~6 (it also preserves NaN / Inf payloads), which made the hi/lo split the largest single
// item of the tf32 producers' instruction mix (ncu r01l: 37 % of hyper_wgrad's instructions).
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - hi);
}
__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
  split_tf32(x.x, hi.x, lo.x);
  split_tf32(x.y, hi.y, lo.y);
  split_tf32(x.z, hi.z, lo.z);
  split_tf32(x.w, hi.w, lo.w);
}

}  // namespace tc
}  // namespace cgat

// ---- bulk async copy (TMA engine, no tensor map: contiguous pre-packed operand tiles) -----------
namespace cgat {
namespace tc {

// this thread arrives on `bar` and announces `bytes` of async-copy traffic that must also land
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// global -> shared, completion signalled on `bar` (SASS: UBLKCP).  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- thread-block clusters: multicast of a pre-packed operand stage to every CTA of the cluster ------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// global -> the SAME shared-memory offset in every CTA of `cta_mask`; each destination CTA's mbarrier at the offset of
// `bar` receives complete_tx(bytes) (SASS: UBLKCP with .MULTICAST).  One L2 read feeds all CTAs of the cluster.
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// tcgen05.commit that arrives on the mbarrier at the offset of `bar` in every CTA of `cta_mask`: "this CTA's MMAs have
// finished reading the stage" reaches the CTA that refills it for the whole cluster (and the others, which re-arm)
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// Packed K-major operand: a [rows x k] fp32 matrix cut into tiles of kPackRows rows x 32 floats, each
// stored as two consecutive 16 KB SWIZZLE_128B images (tf32 hi, then tf32 lo) in the order
// [row_tile][k_chunk][hi|lo].  One bulk copy of 32 KB brings a ready-to-multiply stage.
constexpr int kPackRows = 128;
constexpr int kPackChunk = 32;
constexpr uint32_t kPackImageBytes = kPackRows * kPackChunk * 4;  // 16 KB
constexpr uint32_t kPackStageBytes = 2 * kPackImageBytes;         // hi + lo

__host__ __device__ inline int64_t packed_floats(int64_t rows, int64_t k) {
  int64_t rt = (rows + kPackRows - 1) / kPackRows, kc = (k + kPackChunk - 1) / kPackChunk;
  return rt * kc * (kPackStageBytes / 4);
}


// ---- kind::f16 with a SCALED fp16 hi/lo split ("f16x3") --------------------------------------------
// x ~= hi + lo * 2^-11 with hi = rn_f16(x), lo = rn_f16((x - hi) * 2^11): 11 + 11 significand bits, the same
// 22 bits as the tf32 hi/lo split, but the kind::f16 MMA has twice the rate of kind::tf32 and its operands are
// half as wide (half the shared-memory operand traffic, which is what paces the M = N = 128 kernels).  The scale
// keeps lo in fp16's normal range for |x| down to ~2^-14; the hi*lo + lo*hi products go to their own
// accumulator (as in the tf32 kernels) and are multiplied by 2^-11 once, in the epilogue.  Only for operands that
// are ACTIVATIONS or WEIGHTS (|x| well inside [2^-14 * 2^-11, 65504]); gradient operands stay on the tf32 kernels.
constexpr float kF16LoScale = 2048.f;
constexpr float kF16LoInv = 1.f / 2048.f;

__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t m, uint32_t n) {
  return (1u << 4)               // c_format = F32; a_format = b_format = 0 (F16); both operands K-major
         | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands (K = 16 per instruction), fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- warp-converged issue ("_e" forms) --------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit take their operands from UNIFORM registers.  Issued from inside `if (lane == 0)` the
// compiler cannot use the uniform datapath and wraps EVERY instruction in an elect / R2UR.BROADCAST / branch loop
// (≈ 15 instructions and a back edge per MMA: ncu r02y showed the issuing warp never waiting on a barrier while the
// tensor pipe sat at half rate — the issue loop itself was the pace).  These forms are called by ALL 32 lanes of the
// converged MMA warp with warp-uniform arguments; one lane, chosen by elect.sync inside the asm (always the same lane
// of a converged warp, so tcgen05.commit tracks the MMAs that lane issued), executes the instruction.
__device__ __forceinline__ void umma_f16_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit_multicast_e(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// two fp32 -> packed (hi.x | hi.y << 16), (lo.x | lo.y << 16), lo = rn_f16((x - hi) * lo_scale)
// hi is taken as the fp32 value rounded to 11 significant bits on the bit pattern (round_tf32: fp16 and tf32 carry the
// same 11 bits), which IS the fp16 value whenever x lies in fp16's normal range, so no half -> float conversion is
// needed to form the remainder and each pair costs two packed conversions (F2FP) instead of six scalar ones — the
// split was the largest item of the f16 producers' instruction mix.  Below fp16's normal range (|x| < 2^-14) the
// packed conversion rounds hi once more and that second rounding (< 2^-25 absolute) is not carried into lo: an absolute
// error floor far below the fp32 rounding error of the in-range elements it is summed with.  -DCGAT_F16_SPLIT_EXACT
// restores the conversion-based form.
__device__ __forceinline__ uint32_t pack_half2_rn(float x, float y) {
  const __half2 h = __floats2half2_rn(x, y);   // .x = low half
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void split_f16x2s(float x, float y, float lo_scale, uint32_t& hi, uint32_t& lo) {
#ifdef CGAT_F16_SPLIT_EXACT
  const __half hx = __float2half_rn(x), hy = __float2half_rn(y);
  const __half lx = __float2half_rn((x - __half2float(hx)) * lo_scale);
  const __half ly = __float2half_rn((y - __half2float(hy)) * lo_scale);
  hi = (uint32_t)__half_as_ushort(hx) | ((uint32_t)__half_as_ushort(hy) << 16);
  lo = (uint32_t)__half_as_ushort(lx) | ((uint32_t)__half_as_ushort(ly) << 16);
#else
  const float hx = round_tf32(x), hy = round_tf32(y);
  hi = pack_half2_rn(hx, hy);
  lo = pack_half2_rn((x - hx) * lo_scale, (y - hy) * lo_scale);
#endif
}
// 8 consecutive fp32 (two float4) -> one 16-byte chunk of the hi image and one of the lo image
__device__ __forceinline__ void split_f16x8s(const float4& a, const float4& b, float lo_scale, uint4& hi, uint4& lo) {
  split_f16x2s(a.x, a.y, lo_scale, hi.x, lo.x);
  split_f16x2s(a.z, a.w, lo_scale, hi.y, lo.y);
  split_f16x2s(b.x, b.y, lo_scale, hi.z, lo.z);
  split_f16x2s(b.z, b.w, lo_scale, hi.w, lo.w);
}
// the scaled-lo form used with a separate correction accumulator (hyper_f16.cu)
__device__ __forceinline__ void split_f16x8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  split_f16x8s(a, b, kF16LoScale, hi, lo);
}

// ---- MN-major fp16 operands (contraction over the ROWS of the source: weight-gradient shapes) --------------------
// Canonical SWIZZLE_128B MN-major layout for 16-bit elements (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>:
// Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): an "image" holds the K rows of one block of 64
// consecutive M/N elements, 128 bytes per row; 8-row groups (1024 B) are SBO apart, images LBO apart; inside a group
// the 16-byte unit u of row r sits at unit (u ^ (r & 7)) — physically the same pattern as the K-major SW128 tile.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_16b(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // next block of 64 M/N elements
  d |= (uint64_t)(kSwizzleAtomBytes >> 4) << 32;          // next group of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                                 // SWIZZLE_128B
  return d;
}
// byte offset inside an image of the 8-byte half-unit that holds M/N elements [4*q4, 4*q4 + 4) (q4 in 0..15) of K row r
__device__ __forceinline__ uint32_t mn16_offset(uint32_t r, uint32_t q4) {
  return (r >> 3) * kSwizzleAtomBytes + (r & 7u) * kSwizzleRowBytes + ((((q4 >> 1) ^ (r & 7u)) << 4) | ((q4 & 1u) << 3));
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);   // F32 accumulate, A and B MN-major
}
// four fp32 -> four fp16 hi halves + four fp16 lo halves (lo = rn_f16((x - hi) * lo_scale)), 8 bytes each
__device__ __forceinline__ void split_f16x4s(const float4& a, float lo_scale, uint2& hi, uint2& lo) {
  split_f16x2s(a.x, a.y, lo_scale, hi.x, lo.x);
  split_f16x2s(a.z, a.w, lo_scale, hi.y, lo.y);
}

// Packed K-major fp16 operand: tiles of kPackRows rows x kPackChunk16 halves (128-byte rows, SWIZZLE_128B), stored
// as [row_tile][k_chunk][hi|lo] 16 KB images like the tf32 form; one 32 KB stage now covers 64 contraction steps.
constexpr int kPackChunk16 = 64;
__host__ __device__ inline int64_t packed_floats_f16(int64_t rows, int64_t k) {
  int64_t rt = (rows + kPackRows - 1) / kPackRows, kc = (k + kPackChunk16 - 1) / kPackChunk16;
  return rt * kc * (kPackStageBytes / 4);
}

}  // namespace tc
}  // namespace cgat
