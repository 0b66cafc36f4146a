// clip_kernels.cuh — parameter blocks shared by the kernels (clip_kernels.cu) and the C-ABI host code (api.cu).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "peer.cuh"
#include "sched.h"

namespace flyp {

constexpr int TILE = 128;       // S tile is TILE x TILE (M rows on TMEM lanes, N columns on TMEM columns)
constexpr int KCHUNK = 64;      // bf16 elements per 128-byte swizzle row
constexpr int CHUNK_BYTES = TILE * KCHUNK * 2;  // one [128 x 64] bf16 SW128 operand chunk = 16 KiB

// Fixed exponent shift (log2 units) shared by the forward kernel and its finalize step.  |x| <= |c1| for unit-norm
// rows; c0 = |c1| keeps the dominant terms near exp2(0).  When the range [-2|c1|, 0] would underflow fp32 the window
// is slid up as far as the overflow bound (slack) allows.  Whether the window was adequate is checked a posteriori
// (k_fwd_finalize); if not, the robust (per-tile max) kernels recompute the statistics exactly.
__host__ __device__ inline float fixed_shift(float c1, float slack) {
    float ac = c1 < 0.f ? -c1 : c1;
    float c0 = ac;
    if (2.f * ac > 100.f) { float a = 100.f - ac, b = ac - slack; c0 = a > b ? a : b; }
    return c0;
}

// K-plane schedule of the S contraction.  bf16 features: one term, plane 0.  fp32 features are held as three bf16
// planes x = x1 + x2 + x3 side by side ([n][3 * plane_cols]) and S = sum over the six (p, q) products with p + q <= 4,
// i.e. fp32-accurate logits from bf16 tensor-core products (fp32 accumulation).
struct KPlan {
    int n_terms;       // 1 or 6
    int plane_cols;    // column offset between planes (multiple of 64)
    int pa[6], pb[6];  // plane of the M-side / N-side operand used by term i
};
inline KPlan kplan_bf16() { KPlan k{}; k.n_terms = 1; k.plane_cols = 0; return k; }
inline KPlan kplan_f32(int plane_cols) {
    KPlan k{};
    k.n_terms = 6; k.plane_cols = plane_cols;
    const int a[6] = {0, 0, 1, 0, 1, 2}, b[6] = {0, 1, 0, 2, 1, 0};
    for (int i = 0; i < 6; ++i) { k.pa[i] = a[i]; k.pb[i] = b[i]; }
    return k;
}

// ---- forward: per-row / per-column sum of exp2(c1 * <a_m, b_n> - c0) -------------------------------------------
struct FwdParams {
    int n_m, n_n;            // valid rows of the M-side / N-side operand
    int kc;                  // number of 64-wide K chunks per plane (ceil(D / 64))
    KPlan kplan;             // total chunks of the contraction = kc * kplan.n_terms
    int m_tiles, n_tiles;    // ceil(n / 128)
    int n_slots;             // partial column-sum slots per column block (flat schedule, sched.h: fwd_sched_slots)
    int ld_rows;             // leading dimension (floats) of rowpart / rowmax  (>= m_tiles * 128)
    int ld_cols;             // leading dimension (floats) of colpart / colmax  (>= n_tiles * 128)
    const float* scale;      // device scalar: logit_scale (already exp-ed)
    const int* pos;          // [ld_rows] positive column of each M row (-1: none); that element is EXCLUDED from all
                             // sums (it is added back exactly by the finalize step); nullptr: no exclusion ...
    int pos_arith, pos_off;  // ... unless pos_arith != 0: then the positive of row m is column m + pos_off (if it exists)
    float shift_slack;       // c0 = c1 - shift_slack (log2 units), see DESIGN.md "fixed shift"
    float* rowpart;          // [n_tiles * 2][ld_rows]  partial row sums, one per (N block, column half)
    float* colpart;          // [n_slots][ld_cols]      partial column sums, one per work item that touched the block
    float* rowmax;           // robust mode only: running log2-domain max paired with rowpart
    float* colmax;           // robust mode only
    float* dbg_logits;       // optional [n_m][n_n] fp32 raw dot products (debug path), may be null
    int* argidx;             // robust mode, optional [n_tiles * 2][ld_rows]: column of the per-tile row maximum (lowest
                             // index among equals), written next to rowmax INSTEAD of the sums (fused zero-shot argmax)
    // row-sharded multi-GPU: the N-side rows of other ranks arrive over NVLink while the kernel runs.  Column blocks are
    // visited starting at block nb_rot (this rank's own rows) so that work proceeds in arrival order, and the producer
    // polls wait_b before the first TMA read of a block.
    // label-aware variants (clip/loss.py:123-192; robust kernel only, all nullptr / 0 otherwise): cls_m / cls_n are the
    // class ids of the M-side rows / N-side columns.
    //   mask_mode 1 (exclude): entries with equal class ids are dropped from the sums (the diagonal positive of such a
    //                row is still added back exactly by the finalize step) - the `ignore` variant;
    //   mask_mode 2 (accumulate): instead of (max, sum) the kernel accumulates, over the entries with equal class ids
    //                EXCEPT the positive, rowpart = sum x (log2 units), rowmax = sum log2(1 - P), acc3 = sum P / (1 - P)
    //                with P = 2^(x - acc_lse[row]) - the same-label terms of the soft-label and google_sup variants.
    const int* cls_m;
    const int* cls_n;
    int mask_mode;
    const float* acc_lse;    // [ld_rows] row logsumexp in log2 units (mask_mode 2)
    float* acc3;             // [n_tiles * 2][ld_rows] third accumulator (mask_mode 2)
    int nb_rot;              // first column unit (block; MC kernel: block pair) of the flat schedule
    int n_local;             // column units (from nb_rot on) whose rows are this rank's own: phase A of the schedule
    PeerWait wait_b;
};

// ---- backward sweep: dA[m, :] = s * sum_n dS[m, n] * B[n, :]  with dS recomputed from the S tile ------------------
//  dS[m,n] = wr[m] * exp2(x - lr[m]) + wc[n] * exp2(x - lc[n]), except at the positives, where the value is REPLACED
//  by the exactly precomputed dr[m] (n == labr[m]) / dc[n] (m == labc[n])  (softmax - 1 without cancellation)
//  x = c1 * <a_m, b_n>
struct BwdParams {
    int n_m, n_n;
    int kc;                  // K chunks per plane of the S contraction
    KPlan kplan;
    int f32_mode;            // 1: dS staged as two fp16 planes, features as two fp16 planes (fp32-accurate products)
    int bd_plane_cols;       // column offset between the fp16 planes of the dA operand (f32_mode)
    int m_tiles, n_tiles;
    int d_out;               // number of output columns (= feature dim D)
    int d_parts;             // ceil(d_out / 256)
    const float* scale;
    const float* wr;         // [m] weights of the row-softmax term (nullptr -> term absent)
    const float* lr;         // [m] row logsumexp, log2 domain
    const float* wc;         // [n]
    const float* lc;         // [n] column logsumexp, log2 domain
    const int* labr;         // [m] positive column of row m (or -1), nullptr -> none
    const float* dr;         // [m] value of dS at (m, labr[m])
    const int* labc;         // [n] positive row of column n (or -1), nullptr -> none
    const float* dc;         // [n]
    const void* a_rows;      // M-side operand rows (bf16, ld = lda elements) for the d(scale) reduction; may be null
    int lda;
    void* out;               // [n_m][ld_out] gradient w.r.t. the M-side operand
    int ld_out;
    int out_fp32;            // 1: out is fp32, 0: bf16
    float out_mul;           // extra factor folded into the output (e.g. W for gather_with_grad)
    const uint32_t* gmax_bits;  // device word: bit pattern of max |upstream grad| (dS is staged as scaled fp16)
    const float* fa;         // [m] fast form: wr 2^(c0 - lr)      (see bwd_common.cuh)
    const float* fb;         // [n] fast form: wc 2^(c0 - lc)
    const float* fast_info;  // device {c0, valid}: valid != 0 -> single-exponential epilogue
    // schedule (pair kernel): floor(m_tiles / sched_pairs) rounds of whole row blocks, then the remaining row blocks as a
    // flat tail: their (row block, column step) units, row-block-major, cut into sched_pairs equal contiguous ranges
    // (flat_start).  A range that covers a row block only partly accumulates that part into part_out[2 * pair + (0: the
    // range starts inside the block, 1: it ends inside it)][128][d_out] (fp32), summed by launch_reduce_parts.
    int sched_pairs;
    int n_dh, d_half;        // pair kernel: passes over the output columns (1, or 2 when d_out > 512) and columns per pass;
                             // the schedule runs over the m_tiles * n_dh virtual row blocks (row block, d-half)
    float* part_out;
    PeerWait wait_b, wait_bd; // readiness of the N-side operand rows (tmB) / of their fp16 copy (tmBd), see peer.cuh
    int dbg;                 // debug experiments (FLYP_DBG env): bit 0 = every streamed load reads box (0, 0)
    unsigned long long* prof; // optional debug: per-role wait-cycle counters of cluster 0 (see tools/pair_prof.py)
    float* dscale_part;      // partial sums of dS * <a, b> (unscaled), may be null: [m_tiles * d_parts] (bwd_kernel) or
                             // [2 * sched_pairs] (pair kernel: one per CTA)
    // label-aware variants (single-CTA kernel only; nullptr / 0 otherwise): class ids of the M rows / N columns (cls_n
    // padded with -1), and what happens at the entries with equal ids other than the row's positive:
    //   mask_mode 1: dS = 0 (`ignore`);
    //   mask_mode 2: dS -= mk_r[m] + mk_c[n]                                   (soft labels: the -E/c terms);
    //   mask_mode 3: dS -= mk_r[m] / (1 - P_row) + mk_c[n] / (1 - P_col)       (google_sup_loss).
    const int* cls_m;
    const int* cls_n;
    const float* mk_r;
    const float* mk_c;
    int mask_mode;
    // pair kernel, end-of-sweep reductions by the whole grid (clip_bwd_pair.cu: sweep_tail_reduce):
    int* grid_cnt;           // arrival counter of the grid barrier, zeroed by the host before the launch; nullptr: no
                             // barrier, the fp32 partials of split blocks and the d(scale) partials are left as they are
    uint16_t* ds_out;        // dS kernel (clip_kernels.cu: ds_kernel_mc): the dS matrix [n_m][ds_ld] fp16 it writes
    int ds_ld;
    int keep_ds;             // pair kernel: the staged fp16 dS tiles (scaled by the staging factor) are also written to
                             // global memory through the tmDS store map, for the products over dS (clip_dst_gemm.cu)
    float* dscale_out;       // the sum of all d(scale) partials (written by CTA 0 after the barrier), published to the
                             // other ranks through ds_push
    PeerPush ds_push;
};

// Workers (CTAs; multicast kernel: clusters) of the forward's flat schedule and its partial column-sum slots.
int fwd_workers(int m_tiles, int n_tiles, bool mc, int num_sms);
// robust = exact per-tile (max, sum) row statistics only; gate (device int, may be null): the robust kernel returns
// immediately when *gate == 0.  The grid is fwd_workers(p.m_tiles, p.n_tiles, false, num_sms).
void launch_fwd(const CUtensorMap& tmA, const CUtensorMap& tmB, const FwdParams& p, bool robust, const int* gate,
                int num_sms, cudaStream_t st);
// Both robust passes of the symmetric loss in ONE gated launch: rows of S (tmA x tmB, pr) on the first half of the grid,
// rows of S^T (tmB x tmA, pc) on the second.
void launch_fwd_robust2(const CUtensorMap& tmA, const CUtensorMap& tmB, const FwdParams& pr, const FwdParams& pc,
                        const int* gate, int num_sms, cudaStream_t st);
// Multicast variant of the fast forward (bf16, D <= 512, no debug logits): tmA64 has [64 rows][64 cols] boxes; the
// schedule runs over (n_tiles + 1) / 2 column-block pairs on fwd_workers(.., true, ..) clusters.
void launch_fwd_mc(const CUtensorMap& tmA64, const CUtensorMap& tmB, const FwdParams& p, int num_sms, cudaStream_t st);
// dS kernel of the unfused backward (bf16, D <= 512): the forward's schedule / producer / MMA issuer with an epilogue
// that writes dS (bp.ds_out, fp16 scaled by the staging factor) and one d(scale) partial per CTA (bp.dscale_part).
// Returns the number of CTAs (= partials).  Forward declaration of BwdParams: defined below.
struct BwdParams;
int launch_ds_mc(const CUtensorMap& tmA64, const CUtensorMap& tmB, const FwdParams& p, const BwdParams& bp, int num_sms,
                 cudaStream_t st);
// tmBd: tensor map used for the N-side operand rows as the B operand of the dA MMA (box [64 d][128 n]).
void launch_bwd(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBd, const BwdParams& p,
                int num_sms, cudaStream_t st);
// CTA-pair variant (clip_bwd_pair.cu): requires d_out % 128 == 0, d_out <= 1024; tmA64 has box [64 rows][64 cols];
// column vectors padded to a multiple of 256; dscale_part has 2 entries per M tile.
// tmDS (p.keep_ds != 0): store map of the kept dS matrix ([n_m][n_n] fp16, box [64 rows][64 cols]).
void launch_bwd_pair(const CUtensorMap& tmA64, const CUtensorMap& tmB, const CUtensorMap& tmBd, const CUtensorMap* tmDS,
                     const BwdParams& p, int num_sms, cudaStream_t st);
size_t bwd_pair_smem_bytes();
// number of CTA pairs the schedule of launch_bwd_pair uses (<= num_sms / 2); m_tiles counts VIRTUAL row blocks
int bwd_pair_sched_pairs(int m_tiles, int n_cols, int num_sms);
constexpr int PAIR_NSTEP = 256;      // columns per step of the pair kernel

// ---- products over the kept dS matrix (clip_dst_gemm.cu) -----------------------------------------------------------
constexpr int DST_TILE_ROWS = 256;   // output rows per CTA pair
constexpr int DST_TILE_COLS = 512;   // output columns per pass (the pair's tensor memory: 2 x 256 fp32 columns per CTA)
struct DstParams {
    int n_k;                 // contraction length: rows of dS (transposed) / columns of dS
    int n_out;               // output rows: columns of dS (transposed) / rows of dS
    int dim;                 // feature columns of X16 and of the output (a multiple of 8)
    int col_begin;           // first output column (a multiple of 64): the product covers columns [col_begin, dim) - the
                             // sweep of a wide problem (dim > 512) computes the first columns of its gradient itself and
                             // leaves the others to this product instead of recomputing S for a second pass
    int tile_cols;           // output columns per tile and pass: 512 (one accumulator stage = the whole tensor memory) or
                             // 256 (two stages: the drain of a tile overlaps the MMAs of the next - used when the
                             // epilogue writes to other GPUs over NVLink)
    int out_tiles, n_dh;     // ceil(n_out / 256) tiles x ceil(dim / tile_cols) passes = the virtual tiles of the schedule
    int sched_pairs;         // CTA pairs (dst_gemm_sched_pairs)
    int transposed;          // 1: out = c dS^T X16, 0: out = c dS X16
    const float* scale;      // c = scale * out_mul / G, G the staging factor of dS (gmax_bits, bwd_common.cuh)
    const uint32_t* gmax_bits;
    float out_mul;
    void* out; int ld_out; int out_fp32;
    // rows_per_rank > 0 (row-sharded loss, fp32 only): output row n belongs to rank q = n / rows_per_rank and is written
    // to out_rank[q] + (n - q * rows_per_rank) * ld_out - this rank's slot of rank q's reduce-scatter buffer, local or
    // peer-mapped over NVLink (comm.cu): the product and the scatter half of the reduce-scatter are one kernel
    float* out_rank[PEER_MAXW];
    int rows_per_rank;
    float* part_out;         // [2 * sched_pairs][256][512] fp32 partial tiles of the flat tail
    int* grid_cnt;           // arrival counter of the grid barrier (zeroed before the launch); nullptr: single pair
};
int dst_gemm_sched_pairs(int v_tiles, int k_blocks, int num_sms);
size_t dst_gemm_part_floats(int pairs);
// tmDS: load map of dS ([rows][cols] fp16, box [128][64]); tmX: load map of X16 ([n_k][dim] fp16, box [128][64])
void launch_dst_gemm(const CUtensorMap& tmDS, const CUtensorMap& tmX, const DstParams& p, cudaStream_t st);

size_t fwd_smem_bytes(bool stationary);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is needed once per kernel and device, not once per launch
template <typename K>
inline void ensure_smem_attr(K kernel, size_t smem, bool (&done)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!done[dev]) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        done[dev] = true;
    }
}
size_t bwd_smem_bytes();

}  // namespace flyp
