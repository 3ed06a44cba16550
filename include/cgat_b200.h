/*
 * cgat_b200 — C ABI of the B200-native CGAT hot path.
 *
 * The reference (hyllios/CGAT) is pure Python and has NO FFI: its model is selected by module name
 * (CGAT/lightning_module.py:165-176, `importlib.import_module(hparams.version).CGAtNet`).  The
 * drop-in boundary is therefore the Python class `cgat_b200.CGAT.CGAtNet`; this header is the
 * compute boundary underneath it.  Each entry point replaces the implicit third-party GPU work the
 * reference dispatches at the cited call site.
 *
 * Conventions: plain device pointers + sizes, an explicit CUDA stream (cudaStream_t passed as
 * void*), no allocation inside, no exceptions across the ABI.  Every function returns 0 on success
 * or a non-zero cudaError_t / negative argument-error code; cgat_last_error() gives the message.
 * All floating point is fp32; all indices are int64 at the Python boundary and int32 after
 * cgat_csr_build / cgat_segment_ptr.
 */
#ifndef CGAT_B200_H_
#define CGAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGAT_B200_ABI_VERSION 1

int cgat_abi_version(void);
const char* cgat_last_error(void);
/* number of kernels launched by this library since load (bench.py's `gpu_launches`) */
int64_t cgat_launch_count(void);

/* ---- integer structure -------------------------------------------------
 * Replaces what PyG's MessagePassing.propagate / torch_scatter derive implicitly from edge_index
 * on every call (reference CGAT/CGAT.py:313-326): a CSR-by-destination view of the edge list.
 *   edge_index : (2,E) int64, row 0 = source, row 1 = destination   (CGAT/data.py:140)
 *   edge_attr  : (E,)  int64 shell rank                               (CGAT/prepare_data.py:163-169)
 * Outputs (int32): perm (E) = stable argsort of destinations, rowptr (N+1) = exclusive cumsum of
 * in-degrees, and the permuted src / dst / rank arrays.  Bit-exact vs torch.sort(stable=True),
 * bincount, cumsum.  workspace: >= cgat_csr_workspace_bytes(E,N) bytes.                           */
size_t cgat_csr_workspace_bytes(int64_t n_edges, int64_t n_nodes);
int cgat_csr_build(const int64_t* edge_index, const int64_t* edge_attr, int64_t n_edges, int64_t n_nodes,
                   int32_t* perm, int32_t* rowptr, int32_t* src_sorted, int32_t* dst_sorted,
                   int32_t* rank_sorted, void* workspace, size_t workspace_bytes, void* stream);

/* ptr (n_seg+1) of a SORTED int64 segment index vector: crystal_ptr from batch.batch
 * (CGAT/CGAT.py:567), Roost row pointers from self_fea_idx / crystal_elem_idx
 * (CGAT/roost_message.py:445-453).  Also writes the index as int32.                              */
int cgat_segment_ptr(const int64_t* index, int64_t n, int64_t n_seg, int32_t* ptr, int32_t* index32,
                     void* stream);

/* ---- segmented softmax + weighted sum ---------------------
 * One kernel family replaces torch_geometric.utils.softmax + scatter_add at CGAT/CGAT.py:323-326
 * and :59-61, and scatter_max/scatter_add at CGAT/roost_message.py:307-315.
 * Rows t of segment s are contiguous: [ptr[s], ptr[s+1]).
 *   gate  (n_rows, H, Fa)  Fa == F (vector attention) or 1 (scalar attention)
 *   value (n_rows, H, F)
 *   u     (n_rows) optional per-row multiplier (Roost weights**pow), may be NULL
 *   alpha = u * exp(gate - segmax(gate)) / (segsum(u * exp(gate - segmax)) + eps)
 *   out   (n_seg, H, F) = segsum(alpha * value);   atomic-free, deterministic.
 * Saved for backward: seg_max, seg_den (n_seg, H, Fa).                                           */
int cgat_seg_softmax_fwd(const float* gate, const float* value, const float* u, const int32_t* ptr,
                         int64_t n_seg, int32_t heads, int32_t f, int32_t fa, float eps,
                         float* out, float* seg_max, float* seg_den, void* stream);
/* d_gate (n_rows,H,Fa), d_value (n_rows,H,F) from d_out (n_seg,H,F).  d_gate is also d(log u). */
int cgat_seg_softmax_bwd(const float* gate, const float* value, const float* u, const int32_t* seg_of_row,
                         const float* out, const float* seg_max, const float* seg_den, const float* d_out,
                         int64_t n_rows, int32_t heads, int32_t f, int32_t fa, float eps,
                         float* d_gate, float* d_value, void* stream);

/* ---- dense contraction on the tensor cores -----------
 * C[M,N] = act(A[M,K] * B[N,K]^T + bias[N]); fp32 in/out; tcgen05 kind::tf32 with hi/lo error
 * compensation (3 passes) so results stay at fp32 accuracy.  Replaces the cuDNN grouped-conv /
 * cuBLAS sgemm calls the reference dispatches from nn.Conv1d / nn.Linear (reference
 * CGAT/CGAT.py:91-109, CGAT/Hypernetworksmp.py:82-83, CGAT/message_changed.py:58-63).
 * act: 0 none, 1 LeakyReLU(0.01), 2 tanh, 3 ReLU.  K, lda, ldb multiples of 4; A, B 16-B aligned. */
int cgat_gemm3x_nt(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C,
                   int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t act, void* stream);

/* ---- packed tensor-core operands ---------------------------------------------------------------
 * fp32 [rows x k] (or its transpose, transpose=1: stored [k x rows]) -> 128-row x 32-float tiles,
 * pre-split into (tf32 hi, tf32 lo) and pre-swizzled; layout [row_tile][k_chunk][hi|lo][16 KB].
 * `out` holds cgat_packed_floats(rows,k) floats.  Repack after every weight update.              */
int64_t cgat_packed_floats(int64_t rows, int64_t k);
int cgat_pack_kmajor(const float* w, int64_t ld, int64_t rows, int64_t k, int32_t transpose, float* out,
                     void* stream);

/* ---- fused hypernetwork linear layer ----------------------------------
 * y_out[n,o] = sum_i (sum_k z[n,k] W[o*F+i,k]) * y_in[n,i] + e_term[n,o]
 * Replaces Linear(F -> F*F+F) + view + BatchLinear (reference CGAT/Hypernetworksmp.py:243-254,
 * 205-209) without materialising the (N, F*F+F) predicted-weight tensor.  w_packed =
 * cgat_pack_kmajor(W[:F*F,:F]); e_term carries the bias-shaped remainder (see hyper_fwd.cu).   */
int cgat_hyper_rowdot_fwd(const float* z, const float* y_in, const float* e_term, const float* w_packed,
                          float* y_out, int64_t n_atoms, int32_t f, void* stream);

/* ---- fused edge attention, forward ---------------------------------
 * out[d,h,:] = sum_{t in in(d)} softmax_t(a_t,h)[:] * v_t,h[:] with
 *   hid_t = leaky_relu(P[dst_t, dst-block] + P[src_t, src-block] + T[rank_t]),
 *   a_t,h = W2A_h hidA_t,h + b2A_h,  v_t,h = W2M_h hidM_t,h + b2M_h.
 * Replaces index_select/cat, the grouped Conv1d MLPs, torch_geometric softmax and scatter_add of
 * GATConvNodes.message/aggregate (reference CGAT/CGAT.py:103-109, 319-326) in one kernel.
 *   P (N, 4*H*Hd) = x [W1A_i; W1M_i; W1A_j; W1M_j]^T,  T (K+1, 2*H*Hd) = e [W1A_e; W1M_e]^T + [b1A; b1M]
 *   rowptr/src/dst/rank: cgat_csr_build outputs;  w2a/w2m_packed: cgat_pack_kmajor of (H*F, Hd)
 *   out, seg_max, seg_den: (N, H, F); seg_max/seg_den may be NULL (inference).  F = 128.          */
int cgat_edge_attn_fwd(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                       const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                       const float* w2m_packed, const float* b2a, const float* b2m, float* out,
                       float* seg_max, float* seg_den, int64_t n_atoms, int64_t n_edges, int32_t heads,
                       int32_t f, int32_t hd, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGAT_B200_H_ */
