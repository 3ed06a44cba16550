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

#define CGAT_B200_ABI_VERSION 7 /* 2: bwd_prep bias_sums; f16 entry points. 3: train-step glue. 4: dz_amax. 5: n_ranks, status flags. 6: parts_f16. 7: reduce workspace */

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
                   int32_t* rank_sorted, int32_t n_ranks, void* workspace, size_t workspace_bytes, void* stream);
/* Input validation without a sync on the hot path: what nn.Embedding / index_select refuse in the reference (a node id
 * outside [0, n_nodes), a shell rank outside [0, n_ranks); n_ranks <= 0 means 32, the kernels' table limit) raises a
 * STICKY flag in device memory and the index is clamped (a bad destination drops the edge), so nothing downstream
 * gathers out of bounds.  cgat_status_flags reads (and optionally clears) the flags — it synchronises the device:
 * bit 0 destination, bit 1 source, bit 2 shell rank out of range; bit 3 a non-finite aggregate left
 * cgat_edge_attn_fwd* (fp16 operand overflow of a diverged network, or NaN / Inf inputs).                       */
int cgat_status_flags(uint32_t* flags_out, int32_t reset);

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

/* Persistent form for a SHORT contraction (K <= 128) against a WIDE pre-packed weight — the per-atom
 * first-layer projections P = x [W1A_i; W1M_i; W1A_j; W1M_j]^T (reference CGAT/CGAT.py:103-109, 320-322):
 * w_packed = cgat_pack_kmajor(W, ld, N, K, 0); the A tile stays resident, accumulators are double-buffered.  */
int cgat_gemm3x_nt_res(const float* A, int64_t lda, const float* w_packed, const float* bias, float* C,
                       int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t act, void* stream);

/* Split-K form for long contractions with few output tiles (dL/dx = dL/dP * W1, K = 4*H*Hd): n_split partial
 * products `split_stride` floats apart, no bias / activation; the caller sums them.                 */
int cgat_gemm3x_nt_splitk(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                          int64_t split_stride, int64_t M, int64_t N, int64_t K, int32_t n_split, void* stream);

/* C[s][M,N] = sum_{k in split s} A[k,m] * B[k,n]  (A [K x M], B [K x N] row-major): the weight-gradient
 * shape dW = dY^T X of every nn.Linear / Conv1d(k=1) backward on this path.  Operands are staged
 * MN-major (no transposed copies).  n_split partial results `split_stride` floats apart.          */
int cgat_gemm3x_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                   int64_t split_stride, int64_t M, int64_t N, int64_t K, int32_t n_split, void* stream);
/* Batched form for `batch` (<= 24) operand pairs of identical shape — all weight gradients of a node layer's
 * hypernetwork trunks in one launch.  A, B: HOST arrays of device pointers.  C: (n_split, batch, M, N).
 * colsum (optional, may be NULL): (n_split, batch, M) column sums of A_e over the split's rows (the bias
 * gradients of the same nn.Linear backward), produced while the operand is staged.                      */
int cgat_gemm3x_tn_batched(const float* const* A, const float* const* B, int32_t batch, int64_t lda, int64_t ldb,
                           float* C, float* colsum, int64_t M, int64_t N, int64_t K, int32_t n_split,
                           void* stream);

/* ---- packed tensor-core operands ---------------------------------------------------------------
 * fp32 [rows x k] (transpose=1: given as [k x rows]; transpose=2: each 128x128 block transposed)
 * -> 128-row x 32-float tiles,
 * pre-split into (tf32 hi, tf32 lo) and pre-swizzled; layout [row_tile][k_chunk][hi|lo][16 KB].
 * `out` holds cgat_packed_floats(rows,k) floats.  Repack after every weight update.              */
int64_t cgat_packed_floats(int64_t rows, int64_t k);
int cgat_pack_kmajor(const float* w, int64_t ld, int64_t rows, int64_t k, int32_t transpose, float* out,
                     void* stream);
/* fp16 hi/lo form ("f16x3": x ~= hi + lo * 2^-11, both fp16; the same 22 significand bits as the tf32 pair) for the
 * kind::f16 kernels below, whose operands are activations and weights: tiles of 128 rows x 64 halves, otherwise the
 * same layout, arguments and repack rule.  `out` holds cgat_packed_floats_f16(rows,k) floats.      */
int64_t cgat_packed_floats_f16(int64_t rows, int64_t k);
int cgat_pack_kmajor_f16(const float* w, int64_t ld, int64_t rows, int64_t k, int32_t transpose, float* out,
                         void* stream);
/* General form: hi = f16(w * pre_scale), lo = f16((w * pre_scale - hi) * lo_scale).  cgat_pack_kmajor_f16 is
 * (1, 2^11); the edge-attention kernels, whose three products share one accumulator, read (2^6, 1).          */
int cgat_pack_kmajor_f16s(const float* w, int64_t ld, int64_t rows, int64_t k, int32_t transpose, float pre_scale,
                          float lo_scale, float* out, void* stream);

/* ---- fused hypernetwork linear layer ----------------------------------
 * y_out[n,o] = sum_i (sum_k z[n,k] W[o*F+i,k] + w_bias[o*F+i]) * y_in[n,i] + e_term[n,o]
 * Replaces Linear(F -> F*F+F) + view + BatchLinear (reference CGAT/Hypernetworksmp.py:243-254,
 * 205-209) without materialising the (N, F*F+F) predicted-weight tensor.  w_packed =
 * cgat_pack_kmajor(W[:F*F,:F]); w_bias (optional) = the first F*F entries of the Linear's bias, added to
 * the predicted weights in the epilogue; e_term (+ optional e_term2, may be NULL) carry the bias-shaped
 * remainder (see hyper_fwd.cu).                                                                   */
int cgat_hyper_rowdot_fwd(const float* z, const float* y_in, const float* e_term, const float* e_term2,
                          const float* w_bias, const float* w_packed, float* y_out, int64_t n_atoms, int32_t f,
                          void* stream);

/* ---- fused hypernetwork trunks -------------------------------------------
 * The J (<= 4) HyperLinears of a node layer share the hyper-input h; each runs
 *   t1 = tanh(W1 h + b1) ... t4 = z = tanh(W4 t3 + b4)     (FCBlock, reference CGAT/Hypernetworksmp.py:36-83)
 *   e  = We z + be   with We = last.weight[F*F:], be = last.bias[F*F:]   (bias tail of :243-254)
 * One persistent kernel runs the 5-GEMM chain per (128-atom tile, j) with the activation tile resident in
 * shared memory / TMEM between layers (replaces 20 cuBLAS + 16 tanh launches per node layer, forward).
 *   cgat_hyper_trunk_pack: HOST array of n_mats (<= 24) device pointers to F x F matrices -> packed chain
 *     operands; forward order per j: W1,W2,W3,W4,We (transpose=0); backward: We,W4,W3,W2,W1 (transpose=1).
 *   fwd: T (J,4,N,F) tanh outputs (saved for backward; T[j][3] = z_j), E (J,N,F).
 *   bwd: D[j][3] = (dE[j] We + dZ[j]) * (1 - T[j][3]^2), D[j][s-1] = (D[j][s] W_{s+1}) * (1 - T[j][s-1]^2),
 *        dH[j] = D[j][0] W1;  D (J,4,N,F) feeds cgat_gemm3x_tn_batched for the weight / bias gradients.   */
int64_t cgat_hyper_trunk_packed_floats(int32_t n_mats, int32_t f);
int cgat_hyper_trunk_pack(const float* const* weights, const int64_t* ld, int32_t n_mats, int32_t f,
                          int32_t transpose, float* out, void* stream);
int cgat_hyper_trunk_fwd(const float* h, const float* w_packed, const float* const* biases, float* T, float* E,
                         int64_t n_atoms, int32_t f, int32_t J, void* stream);
int cgat_hyper_trunk_bwd(const float* dE, const float* dZ, const float* T, const float* wt_packed, float* D,
                         float* dH, int64_t n_atoms, int32_t f, int32_t J, void* stream);

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
/* The same kernel on kind::f16 passes (half the W2 stream from L2, twice the tensor rate): identical arguments and
 * results; w2a/w2m_packed = cgat_pack_kmajor_f16s(W2, .., pre_scale 64, lo_scale 1); Hd % 64 == 0.          */
int cgat_edge_attn_fwd_f16(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                           const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                           const float* w2m_packed, const float* b2a, const float* b2m, float* out,
                           float* seg_max, float* seg_den, int64_t n_atoms, int64_t n_edges, int32_t heads,
                           int32_t f, int32_t hd, float eps, void* stream);

/* Backward companion of cgat_hyper_rowdot_fwd (activation gradients):
 *   partial[c][n,j] = sum_{o in chunk c} scale[n,o] * (sum_m a[n,m] Wblk_o[j,m] + w_bias[o*F+j]);  result =
 *   sum_c partial[c]   (w_bias optional: pass the Linear's bias with a = z, NULL with a = y)
 * (a = z, blocks of W)            -> dL/dy_in[n,i] = sum_o g[n,o] * (W z + ..)[o,i]
 * (a = y, transposed blocks of W) -> dL/dz[n,k]    = sum_o g[n,o] * sum_i y[n,i] W[o*F+i,k]
 * partial holds cgat_hyper_rowscale_parts(n_atoms, f) x n_atoms x f floats.                      */
int32_t cgat_hyper_rowscale_parts(int64_t n_atoms, int32_t f);
int cgat_hyper_rowscale(const float* a, const float* scale, const float* w_bias, const float* w_packed,
                        float* partial, int64_t n_atoms, int32_t f, void* stream);
/* The same two products on kind::f16 passes (twice the tensor rate, half the shared-memory operand traffic):
 * identical arguments and results, w_packed = cgat_pack_kmajor_f16(...) (hyper_f16.cu).           */
int cgat_hyper_rowdot_fwd_f16(const float* z, const float* y_in, const float* e_term, const float* e_term2,
                              const float* w_bias, const float* w_packed, float* y_out, int64_t n_atoms, int32_t f,
                              void* stream);
int cgat_hyper_rowscale_f16(const float* a, const float* scale, const float* w_bias, const float* w_packed,
                            float* partial, int64_t n_atoms, int32_t f, void* stream);
/* Partial slots of the f16 form (its own work split: when there are enough SMs it gives every atom tile a fixed
 * number of CTAs, each with a slice of the output channels): partial holds this many x n_atoms x f floats.     */
int32_t cgat_hyper_rowscale_parts_f16(int64_t n_atoms, int32_t f);

/* Weight gradient of the hyper-linear layer: dL/dW[o*F+i, k] = sum_n g[n,o] y[n,i] z[n,k], contracted over
 * atoms with MN-major operands; the scaled rows g[n,o]*y[n,:] are formed while staging (the reference's
 * autograd materialises the (N, F*F) gradient of the predicted weights instead).
 * out: (cgat_hyper_wgrad_splits(N), F*F, F) partial results; sum over dim 0.                       */
int32_t cgat_hyper_wgrad_splits(int64_t n_atoms);
int cgat_hyper_wgrad(const float* g, const float* y, const float* z, float* out, int64_t n_atoms, int32_t f,
                     void* stream);

/* kind::f16 form of cgat_hyper_wgrad (twice the tensor rate; 48 KB stages): the gradient operand g is multiplied by a
 * power of two derived from g_amax[0] = max |g| (device float), which cgat_hyper_rowscale_f16_amax produces while it
 * reads the same g as `scale` (scale_amax must be zeroed before the launch).  F = 128 or 256.
 * tail (optional, F = 128): (cgat_hyper_wgrad_splits(N), F, 2F) partial [g^T y | g^T z], the bias-shaped gradients of
 * the same Linear (dL/db[:F*F] and dL/dW[F*F:]) summed from the rows the kernel stages anyway.                 */
int cgat_hyper_rowscale_f16_amax(const float* a, const float* scale, const float* w_bias, const float* w_packed,
                                 float* partial, float* scale_amax, int64_t n_atoms, int32_t f, void* stream);
int cgat_hyper_wgrad_f16(const float* g, const float* y, const float* z, const float* g_amax, float* out,
                         float* tail, int64_t n_atoms, int32_t f, void* stream);

/* ---- fused edge attention, backward -------------------------------------
 * Step 1: recompute a, v; write d_gate = dL/da, d_msg = dL/dv (E,H,F; destination-sorted rows) and the
 *         LeakyReLU sign masks signs[2][H][ceil(Hd/32)][E].  g_out = dL/d out (N,H,F).
 * Step 2: d_hid = dZ W2 on the tensor cores, d_pre = d_hid * leaky_relu'(pre), G[seg, col_off + ...] =
 *         per-segment sums (= dL/dP blocks); optionally per-rank partial sums d_rank (grid, n_ranks, 2*H*Hd)
 *         — or, with d_pre != NULL (identity order only), just the per-edge d_pre (E, 2*H*Hd), whose segment
 *         sums cgat_edge_attn_reduce takes (G and d_rank are then not written).
 *         wt_*_packed: cgat_pack_kmajor of W2^T per head, shape (H*Hd, F).
 * Step 3: dL/dW2 = dZ^T hid, contraction over edges with MN-major operands; partial results
 *         (cgat_edge_attn_wgrad_splits(H), 2, H, F, Hd).
 * Together they replace autograd through index_select / grouped Conv1d / softmax / scatter_add.     */
int cgat_edge_attn_bwd_prep(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                            const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                            const float* w2m_packed, const float* b2a, const float* b2m, const float* out,
                            const float* seg_max, const float* seg_den, const float* g_out, float* d_gate,
                            float* d_msg, uint32_t* signs, float* bias_sums, float* dz_amax, int64_t n_atoms,
                            int64_t n_edges, int32_t heads, int32_t f, int32_t hd, float eps, void* stream);
/* dz_amax (optional, device float): receives max |d_gate|, |d_msg| — the power-of-two range that the kind::f16
 * weight-gradient kernel (cgat_edge_attn_wgrad_f16) applies to its gradient operand.
 * bias_sums (optional): (cgat_edge_attn_grid(E), 2, H, F) per-CTA column sums of d_msg | d_gate; their sum over
 * dim 0 is dL/d b2 of the message | gate net, so the caller never re-reads the (E, H, F) tensors for it.       */
int32_t cgat_edge_attn_grid(int64_t n_edges);
/* kind::f16 form of cgat_edge_attn_bwd_prep (w2*_packed as for cgat_edge_attn_fwd_f16). */
int cgat_edge_attn_bwd_prep_f16(const float* P, const float* T, const int32_t* rowptr, const int32_t* src,
                            const int32_t* dst, const int32_t* rank, const float* w2a_packed,
                            const float* w2m_packed, const float* b2a, const float* b2m, const float* out,
                            const float* seg_max, const float* seg_den, const float* g_out, float* d_gate,
                            float* d_msg, uint32_t* signs, float* bias_sums, float* dz_amax, int64_t n_atoms,
                            int64_t n_edges, int32_t heads, int32_t f, int32_t hd, float eps, void* stream);
int32_t cgat_edge_attn_dgrad_grid(int64_t n_edges);
int cgat_edge_attn_dgrad(const float* d_gate, const float* d_msg, const uint32_t* signs, const int32_t* segptr,
                         const int32_t* seg, const int32_t* row, const int32_t* rnk, const float* wt_a_packed,
                         const float* wt_m_packed, float* G, int64_t ldg, int32_t col_off, float* d_rank,
                         int32_t n_ranks, float* d_pre, int64_t n_atoms, int64_t n_edges, int32_t heads, int32_t f,
                         int32_t hd, void* stream);
/* kind::f16 form of the d_pre variant of cgat_edge_attn_dgrad (identity edge order): wt_*_packed =
 * cgat_pack_kmajor_f16 of W2^T per head; the gradient operand is scaled by a power of two derived from dz_amax[0]
 * (left by cgat_edge_attn_bwd_prep).  Writes d_pre (E, 2*H*Hd).                                               */
int cgat_edge_attn_dgrad_f16(const float* d_gate, const float* d_msg, const uint32_t* signs, const int32_t* segptr,
                             const int32_t* seg, const float* wt_a_packed, const float* wt_m_packed,
                             const float* dz_amax, float* d_pre, int64_t n_atoms, int64_t n_edges, int32_t heads,
                             int32_t f, int32_t hd, void* stream);
/* Segment sums of the per-edge pre-activation gradients d_pre (E, cols; destination-sorted rows) written by
 * cgat_edge_attn_dgrad(d_pre != NULL): HBM-bound, deterministic, replaces the segment sums of the dgrad epilogue
 * and a second tensor-core pass over source-grouped edges.
 *   G[a, dst_col_off:+cols] = sum over the in-edges of atom a;  G[a, src_col_off:+cols] = sum over its out-edges
 *   (src_rowptr / src_row / src_rank: the source-grouped order, src_row = row of d_pre);
 *   d_rank (cgat_edge_attn_reduce_groups(N), n_ranks, cols): partial sums per shell rank (sum dim 0 = dL/dT);
 *   rank_scratch (cgat_edge_attn_reduce_chunks(N), n_ranks, cols) floats and counters
 *   (cgat_edge_attn_reduce_counters(N, cols) int32, zeroed by the call): workspace — every CTA chunk leaves its
 *   per-rank partial in the scratch and the last CTA of a group of chunks to finish adds the group's partials in
 *   chunk order (deterministic) into d_rank.                                                                     */
int32_t cgat_edge_attn_reduce_chunks(int64_t n_atoms);
int32_t cgat_edge_attn_reduce_groups(int64_t n_atoms);
int32_t cgat_edge_attn_reduce_counters(int64_t n_atoms, int32_t cols);
int cgat_edge_attn_reduce(const float* d_pre, int64_t ldd, const int32_t* dst_rowptr, const int32_t* src_rowptr,
                          const int32_t* src_row, const int32_t* src_rank, float* G, int64_t ldg,
                          int32_t dst_col_off, int32_t src_col_off, float* d_rank, float* rank_scratch,
                          int32_t* counters, int32_t n_ranks, int64_t n_atoms, int32_t cols, void* stream);
int32_t cgat_edge_attn_wgrad_splits(int32_t heads);
int cgat_edge_attn_wgrad(const float* P, const float* T, const int32_t* src, const int32_t* dst,
                         const int32_t* rank, const float* d_gate, const float* d_msg, float* out,
                         int64_t n_edges, int32_t heads, int32_t f, int32_t hd, void* stream);
/* kind::f16 form (twice the tensor rate, 4 x 48 KB stages, two alternating producer groups): d_gate / d_msg are
 * multiplied by a power of two derived from dz_amax[0] (left by cgat_edge_attn_bwd_prep) and the result divided by it. */
int32_t cgat_edge_attn_wgrad_f16_splits(int32_t heads, int32_t f, int32_t hd);   /* partial results it writes */
int cgat_edge_attn_wgrad_f16(const float* P, const float* T, const int32_t* src, const int32_t* dst,
                             const int32_t* rank, const float* d_gate, const float* d_msg, const float* dz_amax,
                             float* out, int64_t n_edges, int32_t heads, int32_t f, int32_t hd, void* stream);

/* ---- train-step glue ----------------------------------------------------------
 * cgat_sum_parts: out[i] = (accumulate ? out[i] : 0) + sum_{p < n_parts} parts[p * part_stride + i], parts added in
 *   index order (deterministic).  Sums the split-K / split-atom partial results of the kernels above (replaces the
 *   library reductions autograd would run at the same places).
 * cgat_adamw_flat: one AdamW step (decoupled weight decay, torch.optim.AdamW arithmetic) over flat parameter /
 *   gradient / first- and second-moment buffers of n floats — the optimizer the reference builds at
 *   CGAT/lightning_module.py:328-344 (default --optim AdamW).  `lr` and `step` are DEVICE floats (a captured CUDA graph
 *   replays with the current values; `step` is incremented before use); grad_scale multiplies the gradient on the way in
 *   (the data-parallel average 1 / world size).  n % 4 == 0, buffers 16-byte aligned.
 * cgat_l1_loss: loss[0] = mean_{i<n} |out[i*ldo] - target[i]| (nn.L1Loss on the first output column against the
 *   normalised target, reference lightning_module.py:206-210, 237-240) and, if grad != NULL, its gradient w.r.t. the
 *   (n_rows, n_cols) prediction (zero for padding rows >= n and columns >= 1).                                     */
int cgat_sum_parts(const float* parts, int32_t n_parts, int64_t part_stride, float* out, int64_t n,
                   int32_t accumulate, void* stream);
/* out (M, N) = act(sum_p parts[p] + bias[N]): the epilogue of cgat_gemm3x_nt_splitk when the layer has a bias /
 * activation (few-tile GEMMs of the Roost / pool / output MLPs are split over K to fill the SMs).          */
int cgat_sum_parts_bias_act(const float* parts, int32_t n_parts, int64_t part_stride, const float* bias, float* out,
                            int64_t M, int64_t N, int32_t act, void* stream);
int cgat_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float* step,
                    float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream);
int cgat_l1_loss(const float* out, int64_t ldo, const float* target, int64_t n, float* loss, float* grad,
                 int64_t ldg, int64_t n_rows, int32_t n_cols, void* stream);

/* ---- device-side collation ------------------------------------------------------
 * Builds the collated, bucket-padded batch CGAtNet.forward takes from a crystal store resident in HBM (packed ragged
 * arrays, cgat_b200/store.py) and a list of selected crystals — what the reference does per batch on the host with
 * torch_geometric Batch.from_data_list (CGAT/lightning_module.py:199-200), collate_batch (CGAT/roost_message.py:
 * 400-458) and the Roost pair lists of CompositionData.__getitem__ (CGAT/data.py:89-96).  Bit-exact against that.
 *   store: x (A, d) atom features; nbr / rank (A, k) int32 LOCAL neighbour index inside the crystal / shell rank;
 *          atom_ptr (C+1); comp_w (M), comp_fea (M, d) distinct-element weights / features; comp_ptr (C+1); y (C).
 *   cgat_collate_plan: exclusive scans (n_sel+1, int64) of the selected crystals' atom / element / pair counts.
 *   cgat_collate_fill: out_x (n_pad, d), edge_index (2, n_pad*k), edge_attr (n_pad*k), batch (n_pad), out_y (n_sel+1),
 *          out_w (nc_pad), out_fea (nc_pad, d), self_idx / nbr_idx (mc_pad), cry_idx (nc_pad); rows beyond the real
 *          totals form ONE dummy crystal (id n_sel) exactly like cgat_b200/batching.pad_batch.                     */
int cgat_collate_plan(const int64_t* sel, int64_t n_sel, const int64_t* atom_ptr, const int64_t* comp_ptr,
                      int64_t* atom_off, int64_t* comp_off, int64_t* pair_off, void* stream);
int cgat_collate_fill(const int64_t* sel, int64_t n_sel, const float* x, const int32_t* nbr, const int32_t* rank,
                      const int64_t* atom_ptr, const float* comp_w, const float* comp_fea, const int64_t* comp_ptr,
                      const float* y, int32_t d, int32_t k, const int64_t* atom_off, const int64_t* comp_off,
                      const int64_t* pair_off, float* out_x, int64_t* edge_index, int64_t* edge_attr, int64_t* batch,
                      float* out_y, float* out_w, float* out_fea, int64_t* self_idx, int64_t* nbr_idx,
                      int64_t* cry_idx, int64_t n_atoms, int64_t n_pad, int64_t n_comp, int64_t nc_pad,
                      int64_t n_pairs, int64_t mc_pad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGAT_B200_H_ */
