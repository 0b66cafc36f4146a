// aux_kernels.cuh — small HBM-bound kernels around the tcgen05 sweeps: pair logits, statistic finalisation, vector
// preparation for the backward sweeps, partial reductions, L2 normalisation.  All are warp-reduced / vectorised.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "peer.cuh"

namespace flyp {

// Forward preparation: t2[i] = scale * log2(e) * <A[i,:], B[idx(i),:]>, pos[i] = idx(i) (int32), idx(i) = labels ?
// labels[i] : offset + i; rows without a valid positive (and padding rows [n, n_pad)) get t2 = -inf, pos = -1.
// gate (device int, may be null): the kernel returns immediately when *gate == 0.  zero_words / n_zero (<= 256, may be
// null): control words cleared by the kernel.  a16 / b16 (bf16 features only, may be null): fp16 copies of A (n rows)
// and B (n_b rows) written in the same pass.
void launch_pair_dot(const void* A, const void* B, int dtype, const float* scale, int n, int n_pad, int n_b, int dim,
                     const int64_t* labels, int offset, float* t2, int* pos, const int* gate, int* zero_words,
                     int n_zero, void* a16, void* b16, cudaStream_t st);

// Where the forward's flat schedule (sched.h) put the partial column sums of a column block.
struct FwdColSched { int m_tiles, n_units, n_local, rot, workers, mc; };
// Single-rank symmetric loss: outputs finished by the finalize kernels themselves (all null: not requested).
struct FwdFinish { float* col_lse; float* col_nll; void* loss; int loss_bf16; };

// Fast-path finalize: rowpart[P][ld_rows] -> row_lse (natural log) and row_nll (= lse - positive logit, computed
// without cancellation); colpart[slots][ld_cols] -> col_stat[3][n_n].  Sets *flag = 1 when the fixed shift was inadequate.
void launch_fwd_finalize(const float* rowpart, int n_rowparts, int ld_rows, int n_m, const float* colpart,
                         const FwdColSched& cs, int ld_cols, int n_n, const float* scale, float slack, const float* t2,
                         const int* pos, int col_pos_offset, float* row_lse, float* row_nll, float* col_stat,
                         int* flag, const FwdFinish& fin, cudaStream_t st);
// Robust finalize (only status[0] = *flag is written when *flag == 0; status may be null): merges exact (max, sum) pairs.
// col_* may be null (rows only).
void launch_fwd_finalize_robust(const float* rowpart, const float* rowmax, int n_rowparts, int ld_rows, int n_m,
                                const float* colpart, const float* colmax, int n_colparts, int ld_cols, int n_n,
                                const float* t2, const int* pos, float* row_lse, float* row_nll, float* col_stat,
                                const int* flag, int* status, const FwdFinish& fin, cudaStream_t st);
// col_stat_all[world][3*n_cols] -> col_lse; loss[i] = 0.5 (row_nll[i] + col_nll[off+i])
// fused argmax: reduce the per-(column block, half) row maxima / columns written by the forward kernel's argmax mode
void launch_argmax_finalize(const float* pmax, const int* pidx, int n_parts, int ld, int n_m, long long* out,
                            float* out_max, cudaStream_t st);
// wait: readiness of col_stat_all / row_nll when other ranks push them (peer.cuh); flags == nullptr -> no waiting
void launch_clip_finish(const float* col_stat_all, int world, const float* row_nll, int n_rows, int n_cols,
                        int row_offset, float* col_lse, float* col_nll, void* loss, int loss_bf16, PeerWait wait,
                        cudaStream_t st);

// Vectors consumed by bwd_kernel, padded with zeros / -1 to a multiple of 128 entries.
//   w[i] = wmul * g[i]; l2[i] = lse[i] * log2(e); lab[i] = labels ? labels[i] : (i + lab_offset if in [0, lab_range) else -1)
//   d[i] = dmul * (g[i] expm1(-nll[i]) + (g2 ? g2[lab[i]] expm1(-nll2[lab[i]]) : 0))  (0 when lab[i] < 0): the exact
//          value of dS at the positive;  gmax_bits[0..2] (zeroed by the caller) accumulate {bits(max|g|), key(max lse2), ~key(min lse2)}
void launch_bwd_prep(int n, int n_pad, const float* g, float wmul, const float* lse, const float* nll,
                     const int64_t* labels, int lab_offset, int lab_range, const float* g2, const float* nll2,
                     float dmul, float* w, float* l2, int* lab, float* d, uint32_t* gmax_bits, cudaStream_t st);
// Row-sharded symmetric loss (n = global batch, rows [off, off + n_loc) are local): the vectors of both sweeps in one
// pass - w = g / 2, l2c / l2r = column / row logsumexp in log2 units (padded to n_pad), lab[n_loc], d[n_loc].
void launch_bwd_prep_sharded(int n, int n_pad, int off, int n_loc, const void* g, int g_bf16, const float* row_lse_all,
                             const float* row_nll_all, const float* col_lse, const float* col_nll, float* w, float* l2c,
                             float* l2r, int* lab, float* d, uint32_t* words, cudaStream_t st);
// words = {bits(max|g|), key(max lse2), ~key(min lse2)} as accumulated by launch_bwd_prep; computes the centre c0 of the
// lse range, info = {c0, valid}, and f_x[i] = w_x[i] * 2^(c0 - l_x[i]) for both vector sets (f_b may be null).
void launch_bwd_fast_vectors(const uint32_t* words, int n_a, const float* w_a, const float* l_a, float* f_a, int n_b,
                             const float* w_b, const float* l_b, float* f_b, float* info, cudaStream_t st);
// out[0] = sum(parts[0..n))   (single block, fixed order -> deterministic); push (may be null): also publish the
// sum to the other ranks (peer.cuh)
void launch_sum_parts(const float* parts, int n, float* out, const PeerPush* push, cudaStream_t st);
void launch_fill_float(float* p, float v, cudaStream_t st);

// dst (fp16, n_elems) = saturating round-to-nearest of src (bf16 or fp32); n_elems % 8 == 0
void launch_to_f16(const void* src, int dtype, size_t n_elems, void* dst, cudaStream_t st);

// fp32 [n][dim] -> three bf16 planes / two fp16 planes side by side ([n][NP * dp], dp = dim rounded up to 64, zero pad)
// wait (may be null): the rows of x are written by other ranks; every block waits for the owners of its rows first
void launch_split_planes_bf16x3(const float* x, int n, int dim, int dp, void* out, cudaStream_t st,
                                const PeerWait* wait = nullptr);
void launch_split_planes_f16x2(const float* x, int n, int dim, int dp, void* out, cudaStream_t st,
                               const PeerWait* wait = nullptr);

// label-aware variants (clip/loss.py:123-192), see flyp_label_stats / flyp_label_sweep in api.cu
void launch_label_prep(int n, int n_pad, const int* cls_a, const int* cls_b, int* pa, int* pb, const float* lse,
                       float* lse2, cudaStream_t st);
void launch_label_sum3(const float* px, const float* pl, const float* pr, int n_parts, int ld, int n, float* out0,
                       float* out1, float* out2, cudaStream_t st);
void launch_label_sweep_prep(int n, int n_pad, const float* wr, const float* lr, const float* wc, const float* lc,
                             const float* d_diag, const float* mk_r, const float* mk_c, const int* cls_a, const int* cls_b,
                             const float* gmax, float* o_wr, float* o_lr, float* o_d, float* o_kr, int* o_ca, float* o_wc,
                             float* o_lc, float* o_kc, int* o_cb, int* o_pos, uint32_t* words, float* fast_info,
                             cudaStream_t st);

void launch_l2norm_fwd(const void* x, int n, int dim, int dtype, void* y, float* inv_norm, cudaStream_t st);
void launch_l2norm_bwd(const void* y, const void* dy, const float* inv_norm, int n, int dim, int dtype, void* dx,
                       cudaStream_t st);

}  // namespace flyp
