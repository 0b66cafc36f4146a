/* flyp_clip.h — C ABI of libflypclip.so: the B200 (sm_100a) implementation of FLYP's contrastive-loss hot path.
 *
 * Every entry point replaces a piece of the reference's Python (joliang17/FLYP); citations are file:line in that repo.
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated; the caller owns every buffer,
 *     including the workspace (size from the matching *_workspace_bytes call).  The library allocates nothing.
 *   - all work is enqueued on the cudaStream_t passed in (as void*); no entry point synchronises the device.
 *   - return 0 on success, a negative code on error; flyp_last_error() returns the message (thread-local).
 *   - `scale` is a device pointer to the already-exponentiated logit_scale (clip/model.py:378), so no host sync.
 *   - dtype: FLYP_BF16 features are bf16; FLYP_F32 features are fp32 (evaluated as 3-way bf16 split products, fp32
 *     accumulation).  All statistics, losses and d(scale) are fp32.  Gradients have the feature dtype.
 *   - feature matrices are row-major, contiguous ([n, dim], leading dimension = dim), 16-byte aligned, dim % 8 == 0.
 */
#ifndef FLYP_CLIP_H
#define FLYP_CLIP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FLYP_BF16 = 0, FLYP_F32 = 1 };
enum {
    FLYP_OK = 0,
    FLYP_ERR_ARG = -1,      /* bad argument (shape, alignment, dtype) */
    FLYP_ERR_CUDA = -2,     /* CUDA runtime / driver error */
    FLYP_ERR_WORKSPACE = -3 /* workspace too small */
};

const char* flyp_last_error(void);
int flyp_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Symmetric contrastive loss, row-sharded.  Replaces clip/loss.py:103-118 (logits) + :194-211 (labels, two
 * cross-entropies, per-item average).  A rank holds n_rows local image rows and all n_cols text rows; local row i is
 * global row row_offset + i and its positive text is column row_offset + i (clip/loss.py:66-67 ordering).
 * world_size == 1: n_rows == n_cols, row_offset == 0.
 * ------------------------------------------------------------------------------------------------------------------ */
int flyp_clip_workspace_bytes(int n_rows, int n_cols, int dim, int dtype, size_t* bytes);
/* 1 when a backward over this problem that asks for both feature gradients keeps dS (single rank, square, bf16, dim a
 * multiple of 128, >= 1024 pairs: the first sweep also writes its staged fp16 dS tiles into the workspace - n_rows x
 * n_cols x 2 bytes of it - and the text gradient is the product dS^T . image over them instead of a second sweep that
 * recomputes the logits), 0 when it runs two sweeps.  For callers that account FLOPs; FLYP_KEEP_DS=0 switches it off. */
int flyp_clip_keeps_ds(int n_rows, int n_cols, int dim, int dtype);
/* Which kernels a backward over this problem runs when both feature gradients are wanted - for callers that account
 * FLOPs and time kernels: 0 = two sweeps (S recomputed twice: 10 n_rows n_cols dim executed per fwd+bwd), 1 = one sweep
 * that keeps dS + the product dS^T . image (8, = algorithmic), 2 (single rank, bf16, dim <= 512) = the unfused backward:
 * a dS kernel with the forward's tensor-core pipeline + the two products dS . text and dS^T . image (8).  On several
 * ranks plan 2 is not used (1 or 0).  Plan 2 is off unless FLYP_UNFUSED=1 (measured slower: its dS kernel is bound by
 * its epilogue); FLYP_KEEP_DS=0 forces plan 0 (A/B measurements). */
int flyp_clip_backward_plan(int n_rows, int n_cols, int dim, int dtype);

/* Forward, local part.  The positive logit of every row is kept out of the tensor-core sums and added back exactly,
 * so losses much smaller than the logits keep full relative accuracy.  Out:
 *   row_lse[n_rows]      natural-log logsumexp_j S[i, j]                  (image->text direction, exact per rank)
 *   row_nll[n_rows]      row_lse[i] - S[i, row_offset + i]                (image->text cross-entropy of item i)
 *   col_stat[3 * n_cols] partial column statistics over the LOCAL rows, log2 units:
 *                        [0, n) m_j reference, [n, 2n) sum_i exp2(S[i,j] log2e - m_j) over local rows excluding the
 *                        positive, [2n, 3n) positive logit of column j if its row is local, else -inf.
 *                        Triples from different ranks merge exactly in flyp_clip_fwd_finish (the only cross-rank
 *                        exchange of the loss).
 *   status[1]            device int (may be NULL): 0 = fixed-shift fast path used, 1 = robust recomputation used */
int flyp_clip_fwd_local(const void* img, const void* txt, const float* scale, int n_rows, int n_cols, int dim,
                        int dtype, int row_offset, float* row_lse, float* row_nll, float* col_stat, int* status,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Forward, finish.  col_stat_all = world x [3 * n_cols] (the gathered col_stat of every rank, rank-major).  Out:
 *   col_lse[n_cols] natural-log logsumexp_i S[i, j] over ALL rows;
 *   col_nll[n_cols] col_lse[j] - S[j, j]                                        (text->image cross-entropy of item j)
 *   loss[n_rows]    0.5 * (row_nll[i] + col_nll[row_offset + i])                (clip/loss.py:208-209) */
int flyp_clip_fwd_finish(const float* col_stat_all, int world, const float* row_nll, int n_rows, int n_cols,
                         int row_offset, float* col_lse, float* col_nll, float* loss, void* stream);

/* Backward, local part (autograd of clip/loss.py:117-118,208-209).  row_lse/row_nll/col_lse/col_nll are the saved
 * forward outputs; g_row[n_rows] / g_col[n_cols] are the upstream gradients on the loss entries of the local rows / of
 * all columns' items.  Out (each may be NULL to skip):
 *   d_img[n_rows, dim]  complete gradient of the local image rows (grad_dtype: FLYP_BF16 or FLYP_F32)
 *   d_txt[n_cols, dim]  this rank's PARTIAL gradient of every text row (sum over ranks = full gradient)
 *   d_scale[1]          this rank's partial d loss / d logit_scale (fp32; requires d_img)
 * grad_mul is folded into d_img / d_txt (1 for gather_with_grad=False, world for True). */
int flyp_clip_bwd_local(const void* img, const void* txt, const float* scale, int n_rows, int n_cols, int dim,
                        int dtype, int row_offset, const float* row_lse, const float* row_nll, const float* col_lse,
                        const float* col_nll, const float* g_row, const float* g_col, float grad_mul, int grad_dtype,
                        void* d_img, void* d_txt, float* d_scale, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * One-directional softmax cross-entropy over logits = scale * A . B^T with integer targets.  Replaces
 * src/models/ce_ablation.py:122-123 (A = image features [n, dim], B = class text features [n_classes, dim]) and
 * serves clip/loss.py:109-111 (local_loss blocks).  labels: int64 device vector or NULL (then target = label_offset + i).
 * ------------------------------------------------------------------------------------------------------------------ */
int flyp_ce_workspace_bytes(int n, int n_classes, int dim, int dtype, size_t* bytes);
/* Out: loss[n] = lse[i] - logit[i, target_i] (reduction='none'), lse[n] natural-log logsumexp (saved for backward). */
int flyp_ce_fwd(const void* a, const void* b, const float* scale, int n, int n_classes, int dim, int dtype,
                const int64_t* labels, int label_offset, float* loss, float* lse, void* workspace,
                size_t workspace_bytes, void* stream);
/* lse / loss: the saved forward outputs; g[n] upstream gradient on loss.  Out (each may be NULL): d_a[n, dim],
 * d_b[n_classes, dim] (grad_dtype), d_scale[1] fp32 (requires d_a).
 * Targets follow F.cross_entropy: label -100 (ignore_index) gives loss 0 and no gradient for that row; any other label
 * outside [0, n_classes) traps the kernel (torch raises a device-side assertion for the same input). */
int flyp_ce_bwd(const void* a, const void* b, const float* scale, int n, int n_classes, int dim, int dtype,
                const int64_t* labels, int label_offset, const float* lse, const float* loss, const float* g,
                int grad_dtype, void* d_a, void* d_b, float* d_scale, void* workspace, size_t workspace_bytes,
                void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Peer-memory exchange between the GPUs of one node (NVLink / NVSwitch).  Replaces the collectives of the row-sharded
 * loss: the two feature all-gathers of clip/loss.py:19-69 (gather_features; rank-major ordering of :66-67) and the
 * O(B) statistics this implementation exchanges instead of replicating logits (clip/loss.py:103-114).
 * One communicator per (rank, process); its exchange segment is device memory owned by the library (the only
 * allocation the library makes, at creation time, never on the hot path).  All calls must be made by every rank in the
 * same order.  Readiness of remotely written rows is signalled through device flag words described by flyp_ready_t;
 * the *_ex entry points below make the tensor-core kernels poll them right before the first read of a rank's rows, so
 * the transfers overlap the kernels.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct flyp_comm flyp_comm;
#define FLYP_IPC_HANDLE_BYTES 64
#define FLYP_COMM_MAX_WORLD 16

/* Rows [k * rows_per_flag, (k + 1) * rows_per_flag) are valid once (int32)(flags[k * stride + j] - seq) >= 0, j < sub.
 * flags == NULL: ready. */
typedef struct {
    const uint32_t* flags; /* device: producer k is done when flags[k * stride + j], j < sub, all reached seq */
    uint32_t seq;
    int n_flags;
    int rows_per_flag;
    int sub;               /* flag words per producer (0 is read as 1) */
    int stride;            /* words between the flag groups of consecutive producers (0 is read as sub) */
    uint32_t timeout_ms;   /* a kernel waiting longer than this for a producer traps (0: waits for ever) */
    uint32_t* err;         /* device-visible word set to 1 + k before the trap when flags[k] timed out; may be NULL */
} flyp_ready_t;

/* The gathered matrices of one step, [world * n_rows, dim] rank-major, in this rank's exchange segment. */
typedef struct {
    const void *img_all, *txt_all;      /* bf16 */
    const void *img16_all, *txt16_all;  /* fp16 copies (operands of the backward's second GEMM) */
    flyp_ready_t img_ready, txt_ready, img16_ready, txt16_ready;
    uint32_t seq;                       /* sequence number of this step (pass it to the push / sum calls) */
} flyp_gathered_t;

/* The gathered statistics of one step. */
typedef struct {
    const float* col_stat_all;          /* [world][3 * n_cols] column triples of every rank (flyp_clip_fwd_local) */
    const float *row_lse_all, *row_nll_all; /* [world * n_rows] */
    flyp_ready_t ready;
} flyp_stats_t;

/* Create the communicator of `rank` on the current device for blocks of at most max_rows x dim bf16 features. */
int flyp_comm_create(int rank, int world, int max_rows, int dim, flyp_comm** comm);
int flyp_comm_segment_bytes(const flyp_comm* comm, size_t* bytes);
/* Size of the exchange segment a communicator of this shape needs (for callers that allocate it themselves). */
int flyp_comm_layout_bytes(int world, int max_rows, int dim, size_t* bytes);
/* Communicator over caller-owned, ZEROED segments that are already mapped in this process: segments[q] = address of
 * rank q's segment (segments[rank] = the own one), multicast = address of an NVSwitch multicast mapping that aliases all
 * of them (or NULL).  This is what torch.distributed._symmetric_memory hands out (buffer_ptrs, multicast_ptr); with a
 * multicast mapping every push is ONE copy / store stream that the switch replicates to all ranks. */
int flyp_comm_create_external(int rank, int world, int max_rows, int dim, void* const* segments, void* multicast,
                              flyp_comm** comm);
int flyp_comm_has_multicast(const flyp_comm* comm);
/* handle_out: FLYP_IPC_HANDLE_BYTES host bytes to be all-gathered by the caller (e.g. through torch.distributed). */
int flyp_comm_ipc_handle(flyp_comm* comm, void* handle_out);
/* all_handles: world x FLYP_IPC_HANDLE_BYTES host bytes, rank-major.  The caller must barrier before the first step. */
int flyp_comm_connect_ipc(flyp_comm* comm, const void* all_handles);
/* Same-process peers (single-process multi-GPU, or several emulated ranks on one GPU in tests). */
int flyp_comm_connect_local(flyp_comm* comm, flyp_comm* const* peers);
/* 0, or 1 + k if a kernel gave up waiting for rank k (host read of a mapped word; no synchronisation).  A kernel that
 * gives up TRAPS right after setting the word - it never continues on rows that have not arrived - so every later CUDA
 * call of the process fails, like after an NCCL watchdog abort. */
int flyp_comm_error(const flyp_comm* comm);
int flyp_comm_reset_error(flyp_comm* comm);
/* How long a kernel waits for a peer's rows before it traps (default 600000 = 10 min, or FLYP_PEER_TIMEOUT_MS at
 * creation; 0 = for ever). */
int flyp_comm_set_timeout_ms(flyp_comm* comm, uint32_t timeout_ms);
/* Rows per rank from which a backward over this communicator computes the text gradient as the product over the kept dS
 * with its NVLink reduce-scatter instead of the transposed sweep (default 6144, or FLYP_RS_MIN_ROWS at creation; 0 =
 * whenever the shape keeps dS, see flyp_clip_keeps_ds): the scatter moves (W - 1) / W x B x D x 4 bytes per rank however
 * small the rank's share, the sweep it replaces shrinks with it.  Must be the same on every rank. */
int flyp_comm_set_rs_min_rows(flyp_comm* comm, int rows);
int flyp_comm_destroy(flyp_comm* comm);

/* clip/loss.py:59-67 without torch.cat: a pack kernel on `stream` copies the local rows (and their fp16 copies) into
 * the own slots of the gathered matrices; the copy engines then push the slots to every peer on the communicator's side
 * stream, text first - with an NVSwitch multicast mapping ONE copy per matrix reaches all ranks, otherwise one copy per
 * peer in ring order - each block copy followed by a 4-byte copy of the step's sequence number into the flag word the
 * consumers poll.  Returns at once; `out` says where the gathered matrices will be and which flags announce them. */
int flyp_comm_gather_features(flyp_comm* comm, const void* img, const void* txt, int n_rows, int dim, int dtype,
                              flyp_gathered_t* out, void* stream);
/* Push this rank's column triples / row statistics (outputs of flyp_clip_fwd_local) into every rank's segment. */
int flyp_comm_push_stats(flyp_comm* comm, uint32_t seq, const float* col_stat, const float* row_lse,
                         const float* row_nll, int n_rows, int n_cols, flyp_stats_t* out, void* stream);
/* All-reduce (sum, fixed rank order) of one fp32 device scalar: push, then sum once every rank's value has arrived. */
int flyp_comm_push_scalar(flyp_comm* comm, uint32_t seq, const float* value, void* stream);
int flyp_comm_sum_scalar(flyp_comm* comm, uint32_t seq, float* out, void* stream);

/* Variants of the loss entry points for operands that other ranks are still writing.
 *   txt_ready        readiness of the rows of `txt` (forward: the schedule starts at this rank's own column block and
 *                    follows the ring order of the push)
 *   txt16 (may be NULL -> converted here), txt16_ready: the fp16 copy of `txt` and its readiness
 *   stats_ready      readiness of col_stat_all / row_nll (all ranks)
 *   loss_dtype       FLYP_F32 or FLYP_BF16: element type of the loss vector written by the finish step
 *   (g_dtype of flyp_clip_bwd_sharded likewise: element type of the upstream gradient vector) */
int flyp_clip_fwd_local_ex(const void* img, const void* txt, const float* scale, int n_rows, int n_cols, int dim,
                           int dtype, int row_offset, float* row_lse, float* row_nll, float* col_stat, int* status,
                           void* workspace, size_t workspace_bytes, const flyp_ready_t* txt_ready, void* stream);
int flyp_clip_fwd_finish_ex(const float* col_stat_all, int world, const float* row_nll, int n_rows, int n_cols,
                            int row_offset, float* col_lse, float* col_nll, void* loss, int loss_dtype,
                            const flyp_ready_t* stats_ready, void* stream);
int flyp_clip_bwd_local_ex(const void* img, const void* txt, const float* scale, int n_rows, int n_cols, int dim,
                           int dtype, int row_offset, const float* row_lse, const float* row_nll, const float* col_lse,
                           const float* col_nll, const float* g_row, const float* g_col, float grad_mul, int grad_dtype,
                           void* d_img, void* d_txt, float* d_scale, void* workspace, size_t workspace_bytes,
                           const void* txt16, const flyp_ready_t* txt_ready, const flyp_ready_t* txt16_ready,
                           void* stream);
/* The one-directional cross-entropy on a class matrix `b` that other ranks are still writing (the local_loss blocks of
 * clip/loss.py:109-111 over the gathered features): b_ready as above; b16 (may be NULL -> converted here) the fp16 copy
 * of b and its readiness. */
int flyp_ce_fwd_ex(const void* a, const void* b, const float* scale, int n, int n_classes, int dim, int dtype,
                   const int64_t* labels, int label_offset, float* loss, float* lse, void* workspace,
                   size_t workspace_bytes, const flyp_ready_t* b_ready, void* stream);
int flyp_ce_bwd_ex(const void* a, const void* b, const float* scale, int n, int n_classes, int dim, int dtype,
                   const int64_t* labels, int label_offset, const float* lse, const float* loss, const float* g,
                   int grad_dtype, void* d_a, void* d_b, float* d_scale, void* workspace, size_t workspace_bytes,
                   const void* b16, const flyp_ready_t* b_ready, const flyp_ready_t* b16_ready, void* stream);

/* Both backward sweeps of a rank of the row-sharded symmetric loss with one shared preparation pass: d_img from the
 * row block img . txt_all^T, d_txt from the transposed block txt . img_all^T (complete gradients of the local rows,
 * no B x D reduce-scatter), d_scale = this rank's share.  *_all / *16_all: gathered features and their fp16 copies
 * (flyp_gathered_t); row_lse_all / row_nll_all / col_lse / col_nll: statistics of all n_cols global rows / columns;
 * g[n_cols]: upstream gradient on the (replicated) loss vector.  Any of d_img, d_txt, d_scale may be NULL. */
int flyp_clip_bwd_sharded(const void* img, const void* txt, const void* img_all, const void* txt_all,
                          const void* img16_all, const void* txt16_all, const float* scale, int n_rows, int n_cols,
                          int dim, int dtype, int row_offset, const float* row_lse_all, const float* row_nll_all,
                          const float* col_lse, const float* col_nll, const void* g, int g_dtype, float grad_mul,
                          int grad_dtype, void* d_img, void* d_txt, float* d_scale, void* workspace,
                          size_t workspace_bytes, const flyp_ready_t* img_ready, const flyp_ready_t* txt_ready,
                          const flyp_ready_t* img16_ready, const flyp_ready_t* txt16_ready, void* stream);

/* Whole-step entry points of a rank - what the drop-in module's forward / backward call, ONE call per direction.
 * comm == NULL (world must be 1): the single-GPU loss of the FLYP loop (src/models/flyp_loss.py:365,496): no exchange,
 * 5 kernel launches forward (preparation, tcgen05 sweep, finalize incl. the loss vector, two gated robust-path stubs)
 * and 5 backward (control-word memset, 2 vector kernels, 2 tcgen05 sweeps that reduce their own partial sums).
 * comm != NULL: flyp_clip_fwd_step = gather (the pack kernel also prepares the positive logits) + forward statistics +
 * flyp_comm_push_stats + flyp_clip_fwd_finish_ex; flyp_clip_bwd_step = flyp_clip_bwd_sharded with the d(logit_scale)
 * all-reduce folded in: the last CTA of the first sweep publishes this rank's partial, the W partials are summed (fixed
 * rank order) after the second sweep.
 * Out: loss[world * n_rows] (the full per-item vector on every rank, clip/loss.py:113-114,208) and the statistics the
 * backward needs (row_lse, row_nll [n_rows]; col_lse, col_nll [world * n_rows]); col_stat (3 * world * n_rows floats)
 * may be NULL (workspace scratch is used).  `step` keeps where the gathered data lives (valid until the next
 * flyp_clip_fwd_step on this communicator: a step's backward must be issued before the next forward).
 * feat16 (comm == NULL, bf16 features, may be NULL): 2 * n_rows * dim fp16 elements that receive the fp16 copies of img
 * and txt for the backward (written by the forward's preparation pass; without it the backward converts again).
 * status (may be NULL): device int, 1 when the robust recomputation ran.
 * d_scale (the global sum) and d_scale_partial (scratch, 1 float) are both NULL or both given when comm != NULL; with
 * comm == NULL only d_scale is used. */
typedef struct {
    flyp_gathered_t gathered;
    flyp_stats_t stats;
} flyp_step_t;
int flyp_clip_fwd_step(flyp_comm* comm, const void* img, const void* txt, const float* scale, int n_rows, int dim,
                       int dtype, int rank, int world, float* row_lse, float* row_nll, float* col_stat, float* col_lse,
                       float* col_nll, void* loss, int loss_dtype, void* feat16, int* status, void* workspace,
                       size_t workspace_bytes, flyp_step_t* step, void* stream);
int flyp_clip_bwd_step(flyp_comm* comm, const flyp_step_t* step, const void* img, const void* txt, const float* scale,
                       int n_rows, int dim, int dtype, int rank, int world, const float* col_lse, const float* col_nll,
                       const void* g, int g_dtype, float grad_mul, int grad_dtype, void* d_img, void* d_txt,
                       float* d_scale_partial, float* d_scale, void* workspace, size_t workspace_bytes, void* stream);

/* flyp_clip_bwd_step in two phases, for callers that step several ranks from ONE thread (a kernel must never wait for
 * a kernel that has not been enqueued): phases = 1 enqueues everything a rank computes and publishes (the sweeps; with
 * the kept-dS backward the product dS^T . image whose fp32 partials go straight into the owners' reduce-scatter buffers
 * over NVLink, and the release of the step's flags), phases = 2 what waits for the other ranks (the sum of the W
 * partials of the text gradient into d_txt, the sum of the d(logit_scale) partials), 3 = both = flyp_clip_bwd_step. */
int flyp_clip_bwd_step_phase(flyp_comm* comm, const flyp_step_t* step, const void* img, const void* txt,
                             const float* scale, int n_rows, int dim, int dtype, int rank, int world,
                             const float* col_lse, const float* col_nll, const void* g, int g_dtype, float grad_mul,
                             int grad_dtype, void* d_img, void* d_txt, float* d_scale_partial, float* d_scale,
                             void* workspace, size_t workspace_bytes, int phases, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Row-wise L2 normalisation x / ||x||_2 (no epsilon), clip/model.py:375-376, src/models/ce_ablation.py:115-118.
 * ------------------------------------------------------------------------------------------------------------------ */
/* y[n, dim] = x / ||x||; inv_norm[n] = 1 / ||x|| (fp32, saved for backward). */
int flyp_l2norm_fwd(const void* x, int n, int dim, int dtype, void* y, float* inv_norm, void* stream);
/* dx = (dy - y <y, dy>) * inv_norm */
int flyp_l2norm_bwd(const void* y, const void* dy, const float* inv_norm, int n, int dim, int dtype, void* dx,
                    void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Label-aware ClipLoss variants - clip/loss.py:123-192: soft labels over the items of equal ground label (:188-192),
 * `ignore` (:132-159: same-label non-diagonal logits leave the softmax), `google_sup_loss` (:160-187).  The same S tiles
 * as the default loss with a class-equality mask in the epilogues; flyp_b200/labeled.py combines the two building
 * blocks below into the three losses and their gradients.  Square problems: a, b are [n, dim], the positive of row i is
 * column i, cls_a / cls_b (int32 device vectors) are the class ids of the rows of a / b.
 * Workspace: flyp_clip_workspace_bytes(n, n, dim, dtype).
 *   flyp_label_stats, mode 1: out0[i] = log sum_j exp S_ij over the columns of a DIFFERENT class plus the diagonal,
 *                              out1[i] = out0[i] - S_ii                             (masked cross-entropy of row i);
 *                     mode 2: over the columns j != i of the SAME class: out0[i] = sum S_ij, out1[i] = sum ln(1 - P_ij),
 *                              out2[i] = sum P_ij / (1 - P_ij), P_ij = exp(S_ij - lse_rows[i]).
 *   flyp_label_sweep: out[m, :] = scale * sum_n dS[m, n] b[n, :] (grad_dtype), d_scale = sum dS[m, n] <a_m, b_n>, with
 *       dS[m, n] = wr[m] exp(S - lr[m]) + wc[n] exp(S - lc[n]) off the same-class entries, d_diag[m] at n = m (if given),
 *       and at the other same-class entries: mode 0 the same; 1: 0; 2: that minus (mk_r[m] + mk_c[n]);
 *       3: that minus (mk_r[m] / (1 - exp(S - lr[m])) + mk_c[n] / (1 - exp(S - lc[n]))).  lr / lc natural-log units;
 *       gmax[0] (device) >= max |dS| sets the fp16 staging scale.
 * ------------------------------------------------------------------------------------------------------------------ */
int flyp_label_stats(const void* a, const void* b, const float* scale, int n, int dim, int dtype, const int* cls_a,
                     const int* cls_b, int mode, const float* lse_rows, float* out0, float* out1, float* out2,
                     void* workspace, size_t workspace_bytes, void* stream);
int flyp_label_sweep(const void* a, const void* b, const float* scale, int n, int dim, int dtype, const float* wr,
                     const float* lr, const float* wc, const float* lc, const float* d_diag, const int* cls_a,
                     const int* cls_b, const float* mk_r, const float* mk_c, int mode, const float* gmax, int grad_dtype,
                     void* out, float* d_scale, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Encoder tail, fused: the final projection of a tower followed by the L2 normalisation that feeds the loss -
 * clip/model.py:242-243 (`x @ self.proj`), :359 (`x[eot] @ self.text_projection`), :375-376 (x / x.norm(dim=-1,
 * keepdim=True)).  y[n, n_out] = z / ||z||_2, z = x[n, k] . w[k, n_out]: one tcgen05 GEMM whose accumulator tile stays
 * in tensor memory, normalised in the epilogue; z is never written.  x, w: bf16, or fp32 (evaluated as 3-way bf16 split
 * products, fp32-accurate; needs the workspace).  y: y_dtype FLYP_BF16 or FLYP_F32; y16 (bf16 y only, may be NULL): fp16
 * copy of the rounded features, the operand format of the loss's backward; inv_norm[n] (may be NULL) = 1 / ||z|| for
 * the backward (flyp_l2norm_bwd, then two plain GEMMs).  n_out % 64 == 0, n_out <= 1024; k % 8 == 0 (fp32: k % 64 == 0).
 * ------------------------------------------------------------------------------------------------------------------ */
int flyp_project_normalize_workspace_bytes(int n, int k, int n_out, int dtype, size_t* bytes);
int flyp_project_normalize_fwd(const void* x, const void* w, int n, int k, int n_out, int dtype, void* y, int y_dtype,
                               void* y16, float* inv_norm, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Zero-shot prediction, src/models/eval.py:150-158 (`logits = ...; pred = logits.argmax(dim=1)`) and
 * src/models/zeroshot.py:56-81: out_index[i] = argmax_j <a_i, b_j> (ties -> lowest j, like torch.argmax), fused into the
 * forward kernel's epilogue - the [n_m, n_n] logits are never written.  out_max (optional) = the maximal dot product.
 * Workspace: flyp_clip_workspace_bytes(n_m, n_n, dim, dtype).
 * ------------------------------------------------------------------------------------------------------------------ */
int flyp_argmax(const void* a, const void* b, int n_m, int n_n, int dim, int dtype, int64_t* out_index, float* out_max,
                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Debug / evaluation: raw dot products <a_i, b_j> as fp32 [n_m, n_n] through the same tensor-core path (used by the
 * parity tests).  Not used by the loss.
 * ------------------------------------------------------------------------------------------------------------------ */
int flyp_debug_logits(const void* a, const void* b, int n_m, int n_n, int dim, int dtype, float* out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Debug: while set (non-NULL, 16 x uint64 device words), the backward sweep of cluster 0 records per-role wait-cycle
 * counters there (tools/pair_prof.py).  Pass NULL to switch it off.  Not thread-safe; never used by the product path. */
int flyp_debug_profile(void* device_buffer_16_u64);
/* Measurement: while set (cudaEvent_t handles, both of a pair or neither; NULL switches off), the library records
 * fwd_start / fwd_stop around the launch of the forward tcgen05 sweep and sweep_start / sweep_stop around the backward
 * sweep number `sweep` (0: d image / first operand, 1: d text) of every call, on the launching stream - kernel-only
 * timings inside a real step (bench.py's roofline).  Not thread-safe; never used by the product path. */
int flyp_debug_kernel_events(void* fwd_start, void* fwd_stop, void* sweep_start, void* sweep_stop, int sweep);

#ifdef __cplusplus
}
#endif
#endif /* FLYP_CLIP_H */
