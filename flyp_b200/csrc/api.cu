// api.cu — the C ABI of libflypclip.so (include/flyp_clip.h): argument checking, workspace carving, TMA descriptor
// construction and kernel orchestration.  No allocation, no device synchronisation, no state kept between calls.
#include "../../include/flyp_clip.h"
#include "aux_kernels.cuh"
#include "clip_kernels.cuh"
#include "comm_internal.h"
#include "tail_kernel.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

flyp::PeerWait to_wait(const flyp_ready_t* r) {
    flyp::PeerWait w;
    memset(&w, 0, sizeof(w));
    if (r != nullptr && r->flags != nullptr) {
        w.flags = r->flags; w.seq = r->seq; w.n_flags = r->n_flags; w.rows_per_flag = r->rows_per_flag > 0 ? r->rows_per_flag : 1;
        w.sub = r->sub > 0 ? r->sub : 1; w.stride = r->stride > 0 ? r->stride : w.sub; w.err = r->err;
        w.timeout_ms = r->timeout_ms;
    }
    return w;
}

// measurement switches: read from the environment ONCE per process (never on the hot path)
int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}
int env_bwd_impl() { static const int v = env_int("FLYP_BWD_IMPL", 0); return v; }   // 1: never the pair sweep, 2: whenever possible
int env_fwd_mc() { static const int v = env_int("FLYP_FWD_MC", 1); return v; }
int env_dbg() { static const int v = env_int("FLYP_DBG", 0); return v; }

#define CUDA_OK(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) return fail(FLYP_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));   \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 row-major [rows][cols] (leading dimension ld elements) -> tiles of [128 rows][64 cols], 128-byte swizzle,
// out-of-bounds elements read as zero.
int make_tmap(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld, bool f16 = false, int box_rows = flyp::TILE) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(FLYP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(FLYP_ERR_ARG, "feature pointer not 16-byte aligned");
    if ((ld % 8) != 0) return fail(FLYP_ERR_ARG, "leading dimension %d not a multiple of 8", ld);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)flyp::KCHUNK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FLYP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

unsigned long long* g_prof_buf = nullptr;   // debug only (flyp_debug_profile)
// measurement only (flyp_debug_kernel_events): events recorded right before / after the forward sweep kernel and the
// backward sweep kernel of the selected sweep, on the launching stream
cudaEvent_t g_ev_fwd[2] = {nullptr, nullptr}, g_ev_sweep[2] = {nullptr, nullptr};
int g_ev_sweep_idx = 0;

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = v > 0 ? v : 148;
    }
    return cached[dev];
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// The CTA-pair backward sweep handles feature dims that are multiples of 128 up to 1024; FLYP_BWD_IMPL=1 forces the
// single-CTA kernel (kept for other dims and for A/B measurements).
int num_sms();
bool use_pair_kernel(int dim, int dtype, int n_m, int n_n) {
    const int forced = env_bwd_impl();                                    // A/B switch: 1 never, 2 whenever the shape allows
    if (forced == 1 || dtype != FLYP_BF16 || dim % 128 != 0 || dim > 1024) return false;
    if (dim <= 512 || forced == 2) return true;
    // two passes over the column halves: pays off once the sweep is long enough to amortise the extra pipeline fills and
    // partial sums (measured: B = 8192, D = 1024 1.02 ms vs 1.35 ms; B = 4096, D = 768 0.42 ms vs 0.35 ms) ...
    const int m_tiles = ceil_div(n_m, flyp::TILE);
    if ((long long)m_tiles * ceil_div(n_n, flyp::TILE) >= 2048) return true;
    // ... or when the single-CTA kernel, which cannot split a row block along the columns, would leave most SMs idle
    // (a rank's share of B = 4096, D = 768 on 8 GPUs: 12 work items, 110 us however few rows the rank owns)
    return 2 * m_tiles * ceil_div(dim, 256) < num_sms();
}
// passes over the output columns of the pair sweep and columns per pass (a multiple of 128, <= 512)
inline int pair_n_dh(int dim) { return dim > 512 ? 2 : 1; }
inline int pair_d_half(int dim) { const int n = pair_n_dh(dim); return ((dim + n - 1) / n + 127) / 128 * 128; }
// d(scale) partial slots and fp32 tail-partial floats a backward sweep over n_m rows x n_n columns may use
size_t sweep_dscale_slots(int n_m, int n_n, int dim, int dtype) {
    const int m_tiles = ceil_div(n_m, flyp::TILE);
    const size_t single = (size_t)m_tiles * ceil_div(dim, 256);      // bwd_kernel: one per (row block, 256 output columns)
    if (!use_pair_kernel(dim, dtype, n_m, n_n)) return single;
    // (the label-aware sweeps always run the single-CTA kernel: room for either)
    size_t pair = (size_t)2 * flyp::bwd_pair_sched_pairs(m_tiles * pair_n_dh(dim), n_n, num_sms());
    if (pair < (size_t)num_sms()) pair = (size_t)num_sms();      // (the dS kernel: one per CTA)
    return pair > single ? pair : single;
}
size_t sweep_part_floats(int n_m, int n_n, int dim, int dtype) {
    if (!use_pair_kernel(dim, dtype, n_m, n_n)) return 0;
    const int m_tiles = ceil_div(n_m, flyp::TILE);
    return (size_t)2 * flyp::bwd_pair_sched_pairs(m_tiles * pair_n_dh(dim), n_n, num_sms()) * flyp::TILE * pair_d_half(dim);
}
// Kept-dS backward (single rank, both feature gradients wanted): the first sweep also writes its staged fp16 dS tiles to
// the workspace and the second gradient is the product dS^T . A over them (clip_dst_gemm.cu) instead of a second sweep.
// FLYP_KEEP_DS=0 switches it off (A/B), FLYP_KEEP_DS_MAX_MB bounds the n_rows x n_cols fp16 matrix (default 24 GiB).
int env_keep_ds() { static const int v = env_int("FLYP_KEEP_DS", 1); return v; }
int env_keep_ds_max_mb() { static const int v = env_int("FLYP_KEEP_DS_MAX_MB", 24576); return v; }
inline int keep_ds_ld(int n_cols) { return ceil_div(n_cols, flyp::PAIR_NSTEP) * flyp::PAIR_NSTEP; }
bool keep_ds_eligible(int n_rows, int n_cols, int dim, int dtype) {
    // square: one rank; n_rows < n_cols: the row block of one of n_cols / n_rows ranks (row-sharded loss)
    if (env_keep_ds() == 0 || n_rows < 1024 || n_cols % n_rows != 0 || n_cols / n_rows > FLYP_COMM_MAX_WORLD) return false;
    if (dtype != FLYP_BF16 || dim % 128 != 0 || dim > 1024) return false;
    if (!use_pair_kernel(dim, dtype, n_rows, n_cols)) return false;
    return (size_t)n_rows * keep_ds_ld(n_cols) * 2 <= (size_t)env_keep_ds_max_mb() << 20;
}
constexpr int VEC_PAD = 256;   // per-row / per-column vectors are padded to this many entries
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int plane_cols(int dim) { return ceil_div(dim, 64) * 64; }

float shift_slack(int n_m, int n_n) {
    int n = n_m > n_n ? n_m : n_n;
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    return 126.f - 2.f - (float)lg;
}

struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return r;
    }
};

// ---- workspace of one forward statistics pass over S = A . B^T ([n_m] x [n_n]) ------------------------------------
struct StatsWs {
    int m_tiles, n_tiles, ld_rows, ld_cols;
    bool use_mc;                         // multicast (2-CTA cluster) forward kernel
    int workers, n_slots;                // flat schedule of the fast forward (sched.h)
    float *rowpart, *rowmax, *colpart;   // fast pass + robust row pass
    float *rowpart2, *rowmax2;           // robust column pass (roles swapped): [m_tiles*2][ld_cols]
    float* t2;                           // [ld_rows] positive logit of each row, log2 units
    int* pos;                            // [ld_rows] positive column of each row, -1 = none, -2 = ignored row
    int* flag;                           // [4] control words; [0]: the fixed-shift fast path was inadequate
    uint16_t *planes_a, *planes_b;       // fp32 features: three bf16 planes each ([n][3 * plane_cols])
};
void carve_stats(Carver& c, int n_m, int n_n, int dim, int dtype, bool want_cols, StatsWs& w) {
    w.m_tiles = ceil_div(n_m, flyp::TILE);
    w.n_tiles = ceil_div(n_n, flyp::TILE);
    w.use_mc = dtype == FLYP_BF16 && dim <= 512 && w.n_tiles >= 2 && env_fwd_mc() != 0;
    const int sms = num_sms();
    w.workers = flyp::fwd_workers(w.m_tiles, w.n_tiles, w.use_mc, sms);
    // partial column-sum slots: the worst case over the phase splits run_stats may choose (none; the rank's own rows
    // first) and over both kernels (the debug-logits pass always uses the plain one)
    w.n_slots = 1;
    for (int mc = 0; mc < 2; ++mc) {
        if (mc && !w.use_mc) continue;
        const int units = mc ? (w.n_tiles + 1) / 2 : w.n_tiles;
        const int wk = flyp::fwd_workers(w.m_tiles, w.n_tiles, mc != 0, sms);
        const int loc = n_m / (flyp::TILE * (mc ? 2 : 1));
        const int cand[2] = {0, loc < units ? loc : 0};
        for (int k = 0; k < 2; ++k) {
            const int s2 = flyp::fwd_sched_slots(w.m_tiles, units, cand[k], wk);
            if (s2 > w.n_slots) w.n_slots = s2;
        }
    }
    w.ld_rows = w.m_tiles * flyp::TILE;
    w.ld_cols = w.n_tiles * flyp::TILE;
    w.rowpart = c.take<float>((size_t)w.n_tiles * 2 * w.ld_rows);
    w.rowmax = c.take<float>((size_t)w.n_tiles * 2 * w.ld_rows);
    w.colpart = c.take<float>((size_t)w.n_slots * (w.ld_cols + flyp::TILE));
    if (want_cols) {
        w.rowpart2 = c.take<float>((size_t)w.m_tiles * 2 * w.ld_cols);
        w.rowmax2 = c.take<float>((size_t)w.m_tiles * 2 * w.ld_cols);
    } else {
        w.rowpart2 = w.rowmax2 = nullptr;
    }
    w.t2 = c.take<float>(w.ld_rows);
    w.pos = c.take<int>(w.ld_rows);
    w.flag = c.take<int>(4);
    if (dtype == FLYP_F32) {
        w.planes_a = c.take<uint16_t>((size_t)n_m * 3 * plane_cols(dim));
        w.planes_b = c.take<uint16_t>((size_t)n_n * 3 * plane_cols(dim));
    } else {
        w.planes_a = w.planes_b = nullptr;
    }
}

// One forward statistics pass.
struct StatsIO {
    const void *A, *B;
    const float* scale;
    int n_m, n_n, dim, dtype;
    const int64_t* labels;               // integer targets, or null: the positive of row i is column pos_offset + i
    int pos_offset;
    float *row_lse, *row_nll, *col_stat; // col_stat null: row direction only (ce head)
    int* status;                         // optional device int: 1 when the robust path ran
    float* dbg_logits;                   // debug: raw dot products instead of statistics
    const flyp_ready_t* b_ready;         // multi-GPU: arrival flags of the rows of B
    void *a16, *b16;                     // optional (bf16 features): fp16 copies of A / B to produce on the way
    bool pre_done;                       // t2 / pos / control words were already prepared (pack kernel of the gather)
    flyp::FwdFinish fin;                 // single-rank symmetric loss: finish col_lse / col_nll / loss in the same pass
};

// Statistics of S = scale * A . B^T with one positive per row (labels, or column pos_offset + i):
//   row_lse[n_m] natural-log logsumexp, row_nll[n_m] = row_lse - positive logit, and (col_stat != null) the column
//   triples (m_j, sum_j, t_j), see k_fwd_finalize.  The positives are excluded from the tensor-core sums and added back
//   exactly, so a loss much smaller than the logits keeps full relative accuracy.
// Launches: preparation (pair logits, fp16 copies, control words), the tcgen05 sweep, finalize, and the robust
// recomputation + its finalize, both gated on a device flag (they return at once when the fast path was adequate).
int run_stats(const StatsIO& io, const StatsWs& w, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    int rc;
    const int n_m = io.n_m, n_n = io.n_n, dim = io.dim, dtype = io.dtype;
    flyp::KPlan kplan = flyp::kplan_bf16();
    const bool dbg = io.dbg_logits != nullptr;
    // fp32 features whose rows other ranks are still writing: the kernel that splits them into planes waits for the
    // owners of its rows; the tensor-core kernels then read complete, local planes
    const flyp_ready_t* kernel_b_ready = dtype == FLYP_F32 ? nullptr : io.b_ready;
    if (dtype == FLYP_F32) {
        const int dp = plane_cols(dim);
        const flyp::PeerWait wb = to_wait(io.b_ready);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(io.A), n_m, dim, dp, w.planes_a, st);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(io.B), n_n, dim, dp, w.planes_b, st, &wb);
        CUDA_OK(cudaGetLastError());
        if ((rc = make_tmap(&tmA, w.planes_a, n_m, 3 * dp, 3 * dp)) != 0) return rc;
        if ((rc = make_tmap(&tmB, w.planes_b, n_n, 3 * dp, 3 * dp)) != 0) return rc;
        kplan = flyp::kplan_f32(dp);
    } else {
        if ((rc = make_tmap(&tmA, io.A, n_m, dim, dim)) != 0) return rc;
        if ((rc = make_tmap(&tmB, io.B, n_n, dim, dim)) != 0) return rc;
    }
    const int sms = num_sms();
    const float slack = shift_slack(n_m, n_n);
    const bool mc = w.use_mc && !dbg;

    flyp::FwdParams p;
    memset(&p, 0, sizeof(p));
    p.n_m = n_m; p.n_n = n_n; p.kc = ceil_div(dim, flyp::KCHUNK); p.kplan = kplan;
    p.m_tiles = w.m_tiles; p.n_tiles = w.n_tiles; p.n_slots = w.n_slots;
    p.ld_rows = w.ld_rows; p.ld_cols = w.ld_cols;
    p.scale = io.scale; p.shift_slack = slack;
    p.rowpart = w.rowpart; p.colpart = w.colpart; p.rowmax = w.rowmax; p.colmax = nullptr;
    p.dbg_logits = io.dbg_logits;
    if (!dbg) {
        if (!io.pre_done) {
            flyp::launch_pair_dot(io.A, io.B, dtype, io.scale, n_m, w.ld_rows, n_n, dim, io.labels, io.pos_offset, w.t2,
                                  w.pos, nullptr, w.flag, 4, io.a16, io.b16, st);
            CUDA_OK(cudaGetLastError());
        }
        p.pos = w.pos;
    }
    flyp::FwdColSched cs;
    {
        // multi-GPU: rows of B owned by other ranks are still arriving; start at this rank's own column block
        flyp::FwdParams pf = p;
        pf.wait_b = to_wait(kernel_b_ready);
        const int unit_cols = flyp::TILE * (mc ? 2 : 1);
        cs.n_units = mc ? (w.n_tiles + 1) / 2 : w.n_tiles;
        if (pf.wait_b.flags != nullptr) {
            // the schedule starts at this rank's own rows and follows the ring order of the pushes.  When the own rows
            // are a large share of the work (<= 4 ranks) every worker first takes its part of them (phase A) so that
            // nobody waits for the gather; with more ranks the local part is too short to hide the transfer and the
            // extra work items (each reloads a 128 KB stationary operand) cost more than the overlap gains.
            pf.nb_rot = io.pos_offset / unit_cols;
            const int loc = n_m / unit_cols;
            pf.n_local = (loc < cs.n_units && 4 * loc >= cs.n_units) ? loc : 0;
        }
        cs.m_tiles = w.m_tiles; cs.rot = pf.nb_rot; cs.n_local = pf.n_local; cs.mc = mc ? 1 : 0;
        cs.workers = flyp::fwd_workers(w.m_tiles, w.n_tiles, mc, sms);
        if (g_ev_fwd[0] != nullptr && !dbg) cudaEventRecord(g_ev_fwd[0], st);
        if (mc) {
            CUtensorMap tmA64;
            if ((rc = make_tmap(&tmA64, io.A, n_m, dim, dim, false, 64)) != 0) return rc;
            flyp::launch_fwd_mc(tmA64, tmB, pf, sms, st);
        } else {
            flyp::launch_fwd(tmA, tmB, pf, /*robust=*/false, nullptr, sms, st);
        }
        if (g_ev_fwd[1] != nullptr && !dbg) cudaEventRecord(g_ev_fwd[1], st);
    }
    CUDA_OK(cudaGetLastError());
    if (dbg) return 0;
    flyp::launch_fwd_finalize(w.rowpart, w.n_tiles * 2, w.ld_rows, n_m, w.colpart, cs, w.ld_cols, n_n, io.scale, slack,
                              w.t2, w.pos, io.pos_offset, io.row_lse, io.row_nll, io.col_stat, w.flag, io.fin, st);
    CUDA_OK(cudaGetLastError());

    // Robust recomputation, gated on the device flag (the kernels return immediately when the fast path was adequate).
    if (io.col_stat != nullptr) {
        // second pass with the roles swapped: rows of S^T are the columns of S; the positive of column j is local row
        // j - pos_offset (explicit labels never come with a column direction)
        flyp::FwdParams pt = p;
        pt.n_m = n_n; pt.n_n = n_m; pt.m_tiles = w.n_tiles; pt.n_tiles = w.m_tiles;
        pt.ld_rows = w.ld_cols; pt.ld_cols = w.ld_rows;
        pt.rowpart = w.rowpart2; pt.rowmax = w.rowmax2; pt.colpart = nullptr;
        pt.pos = nullptr; pt.pos_arith = 1; pt.pos_off = -io.pos_offset;
        flyp::launch_fwd_robust2(tmA, tmB, p, pt, w.flag, sms, st);
    } else {
        flyp::launch_fwd(tmA, tmB, p, /*robust=*/true, w.flag, sms, st);
    }
    CUDA_OK(cudaGetLastError());
    flyp::launch_fwd_finalize_robust(w.rowpart, w.rowmax, w.n_tiles * 2, w.ld_rows, n_m, w.rowpart2, w.rowmax2,
                                     w.m_tiles * 2, w.ld_cols, n_n, w.t2, w.pos, io.row_lse, io.row_nll, io.col_stat,
                                     w.flag, io.status, io.fin, st);
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- workspace of the backward sweeps -----------------------------------------------------------------------------
struct VecSet { float *w, *l2, *d, *f; int* lab; };
void carve_vecs(Carver& c, int n_pad, VecSet& v) {
    v.w = c.take<float>(n_pad); v.l2 = c.take<float>(n_pad); v.d = c.take<float>(n_pad); v.f = c.take<float>(n_pad);
    v.lab = c.take<int>(n_pad);
}

// Control words of one backward call, cleared by ONE memset at its start:
//   [0..3]  {bits(max|g|), key(max lse2), ~key(min lse2), -} accumulated by the prep kernels
//   [4 + s] arrival counter of the end-of-kernel grid barrier of sweep / product s (0, 1, 2)
constexpr int CTRL_WORDS = 8;
struct BwdCtrl {
    uint32_t* words;
    int* grid_cnt(int sweep) const { return reinterpret_cast<int*>(words) + 4 + sweep; }
};

int check_common(int n_m, int n_n, int dim, int dtype) {
    if (n_m <= 0 || n_n <= 0) return fail(FLYP_ERR_ARG, "empty operand (%d x %d)", n_m, n_n);
    if (dim <= 0 || dim % 8 != 0) return fail(FLYP_ERR_ARG, "dim=%d must be a positive multiple of 8", dim);
    if (dtype != FLYP_BF16 && dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "dtype %d not supported (bf16 / fp32)", dtype);
    return 0;
}

// One backward sweep: gradient w.r.t. the rows of A.
struct SweepIO {
    const void *A, *B, *B_f16;           // bf16: the features and the fp16 copy of B.  fp32: A_planes / B_planes are the
    const void *A_planes, *B_planes;     // 3-plane bf16 splits, B_f16 the 2-plane fp16 split of B; output is fp32
    int dtype;
    const float* scale;
    int n_m, n_n, dim;
    const float *wr, *lr, *wc, *lc;      // row / column softmax terms (null: term absent)
    const int* labr; const float* dr;    // exact dS at the positive of each row
    const int* labc; const float* dc;    // ... of each column
    const float *fa, *fb, *fast_info;
    void* out; int out_fp32; float out_mul;
    float* dscale_part;                  // null: no d(scale)
    float* dscale_out;                   // the sweep's total (summed in-kernel by the pair sweep, by k_sum_parts otherwise)
    size_t n_dscale;
    float* part_scratch;
    const BwdCtrl* ctrl; int sweep;      // control words and which counter set this sweep uses
    const flyp::PeerPush* ds_push;       // multi-GPU: where the total is published (may be null)
    const flyp_ready_t *b_ready, *b16_ready;
    const int *cls_m, *cls_n;            // label-aware variants (clip_kernels.cuh BwdParams): class ids, weights, mode
    const float *mk_r, *mk_c;
    int mask_mode;
    uint16_t* ds_keep; int ds_ld;        // pair sweep: also write the staged dS tiles here ([n_m][ds_ld] fp16); may be null
    int first_pass_only;                 // wide problems (dim > 512) that keep dS: only the first pair_d_half(dim) output
                                         // columns; the others come from a product over the kept dS (no second S recompute)
};

int run_sweep(const SweepIO& io, cudaStream_t st) {
    CUtensorMap tmA, tmB, tmBd;
    int rc;
    const int n_m = io.n_m, n_n = io.n_n, dim = io.dim, dtype = io.dtype;
    const bool f32 = dtype == FLYP_F32;
    const int dp = plane_cols(dim);
    flyp::KPlan kplan = flyp::kplan_bf16();
    if (f32) {
        if (!io.out_fp32) return fail(FLYP_ERR_ARG, "fp32 features need fp32 gradients");
        if ((rc = make_tmap(&tmA, io.A_planes, n_m, 3 * dp, 3 * dp)) != 0) return rc;
        if ((rc = make_tmap(&tmB, io.B_planes, n_n, 3 * dp, 3 * dp)) != 0) return rc;
        if ((rc = make_tmap(&tmBd, io.B_f16, n_n, 2 * dp, 2 * dp, true)) != 0) return rc;
        kplan = flyp::kplan_f32(dp);
    } else {
        if ((rc = make_tmap(&tmA, io.A, n_m, dim, dim)) != 0) return rc;
        if ((rc = make_tmap(&tmB, io.B, n_n, dim, dim)) != 0) return rc;
        if ((rc = make_tmap(&tmBd, io.B_f16, n_n, dim, dim, true)) != 0) return rc;
    }
    flyp::BwdParams p;
    memset(&p, 0, sizeof(p));
    p.n_m = n_m; p.n_n = n_n; p.kc = ceil_div(dim, flyp::KCHUNK); p.kplan = kplan;
    p.f32_mode = f32 ? 1 : 0; p.bd_plane_cols = dp;
    p.m_tiles = ceil_div(n_m, flyp::TILE); p.n_tiles = ceil_div(n_n, flyp::TILE);
    p.d_out = dim; p.d_parts = ceil_div(dim, 256);
    p.scale = io.scale; p.wr = io.wr; p.lr = io.lr; p.wc = io.wc; p.lc = io.lc;
    p.labr = io.labr; p.dr = io.dr; p.labc = io.labc; p.dc = io.dc;
    p.fa = io.fa; p.fb = io.fb; p.fast_info = io.fast_info;
    p.a_rows = io.dscale_part ? io.A : nullptr; p.lda = dim;
    p.out = io.out; p.ld_out = dim; p.out_fp32 = io.out_fp32; p.out_mul = io.out_mul; p.gmax_bits = io.ctrl->words;
    p.dscale_part = io.dscale_part;
    p.wait_b = to_wait(io.b_ready); p.wait_bd = to_wait(io.b16_ready);
    p.prof = g_prof_buf;
    p.dbg = env_dbg();
    const bool push = io.dscale_part != nullptr && io.ds_push != nullptr && io.ds_push->n_dst > 0;
    const bool timed = g_ev_sweep[0] != nullptr && g_ev_sweep_idx == io.sweep;
    if (timed) cudaEventRecord(g_ev_sweep[0], st);
    p.cls_m = io.cls_m; p.cls_n = io.cls_n; p.mk_r = io.mk_r; p.mk_c = io.mk_c; p.mask_mode = io.mask_mode;
    if (io.mask_mode == 0 && use_pair_kernel(dim, dtype, n_m, n_n)) {
        CUtensorMap tmA64;
        if ((rc = make_tmap(&tmA64, io.A, n_m, dim, dim, false, 64)) != 0) return rc;
        if (io.part_scratch == nullptr) return fail(FLYP_ERR_ARG, "the pair sweep needs its partial-sum scratch");
        p.n_dh = io.first_pass_only ? 1 : pair_n_dh(dim); p.d_half = pair_d_half(dim);
        p.sched_pairs = flyp::bwd_pair_sched_pairs(p.m_tiles * p.n_dh, n_n, num_sms());
        p.part_out = io.part_scratch;
        p.grid_cnt = (env_dbg() & 2) ? nullptr : io.ctrl->grid_cnt(io.sweep);   // (debug bit 1: partials left unreduced)
        if (io.dscale_part != nullptr) {
            p.dscale_out = io.dscale_out;
            if (push) p.ds_push = *io.ds_push;
        }
        CUtensorMap tmDS;
        if (io.ds_keep != nullptr) {
            if ((rc = make_tmap(&tmDS, io.ds_keep, n_m, n_n, io.ds_ld, true, 64)) != 0) return rc;
            p.keep_ds = 1;
        }
        flyp::launch_bwd_pair(tmA64, tmB, tmBd, io.ds_keep != nullptr ? &tmDS : nullptr, p, num_sms(), st);
        if (timed) cudaEventRecord(g_ev_sweep[1], st);
        CUDA_OK(cudaGetLastError());
    } else {
        flyp::launch_bwd(tmA, tmB, tmBd, p, num_sms(), st);
        if (timed) cudaEventRecord(g_ev_sweep[1], st);
        CUDA_OK(cudaGetLastError());
        if (io.dscale_part != nullptr) {
            if ((size_t)p.m_tiles * p.d_parts > io.n_dscale) return fail(FLYP_ERR_WORKSPACE, "d(scale) partial slots");
            flyp::launch_sum_parts(io.dscale_part, p.m_tiles * p.d_parts, io.dscale_out, push ? io.ds_push : nullptr, st);
            CUDA_OK(cudaGetLastError());
        }
    }
    return 0;
}

}  // namespace

namespace flyp {
// used by comm.cu: same thread-local message buffer as the rest of the C ABI
void set_error(int code, const char* fmt, ...) {
    (void)code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace flyp

extern "C" {

const char* flyp_last_error(void) { return g_err; }
int flyp_version(void) { return 100; }

// ------------------------------------------------------------------------------------------------ clip
struct ClipWs {
    StatsWs stats;
    VecSet rows, cols;       // bwd vectors: by local image row / by text column
    float* dscale_part;
    size_t n_dscale;
    float* part_scratch;      // fp32 partial outputs of split tail blocks (pair kernel)
    BwdCtrl ctrl;
    float* fast_info;         // {c0, valid}
    float* col_stat;          // [3 * n_cols] scratch for the column triples when the caller does not want them
    uint16_t *img16, *txt16;  // fp16 staging copies of the features when the caller keeps none (backward only)
    uint16_t* ds_keep;        // kept-dS backward (keep_ds_eligible): [n_rows][ds_ld] fp16, else nullptr
    int ds_ld, gemm_pairs;
    float* gemm_part;         // fp32 partial tiles of the dS^T . A product
    size_t bytes;
};
static void carve_clip(void* base, int n_rows, int n_cols, int dim, int dtype, ClipWs& w) {
    Carver c(base);
    carve_stats(c, n_rows, n_cols, dim, dtype, true, w.stats);
    const int rp = ceil_div(n_rows, VEC_PAD) * VEC_PAD, cp = ceil_div(n_cols, VEC_PAD) * VEC_PAD;
    carve_vecs(c, rp, w.rows);
    carve_vecs(c, cp, w.cols);
    w.n_dscale = sweep_dscale_slots(n_rows, n_cols, dim, dtype);
    w.dscale_part = c.take<float>(w.n_dscale);
    {
        const size_t a = sweep_part_floats(n_rows, n_cols, dim, dtype), b = sweep_part_floats(n_cols, n_rows, dim, dtype);
        const size_t n = a > b ? a : b;
        w.part_scratch = n ? c.take<float>(n) : nullptr;
    }
    w.ctrl.words = c.take<uint32_t>(CTRL_WORDS);
    w.fast_info = c.take<float>(2);
    w.col_stat = c.take<float>((size_t)3 * n_cols);
    // fp16 staging copy (bf16 features) or two fp16 planes (fp32 features)
    const size_t w16 = dtype == FLYP_F32 ? (size_t)2 * plane_cols(dim) : (size_t)dim;
    w.img16 = c.take<uint16_t>((size_t)n_rows * w16);
    w.txt16 = c.take<uint16_t>((size_t)n_cols * w16);
    w.ds_keep = nullptr; w.gemm_part = nullptr; w.ds_ld = 0; w.gemm_pairs = 0;
    if (keep_ds_eligible(n_rows, n_cols, dim, dtype)) {
        w.ds_ld = keep_ds_ld(n_cols);
        w.gemm_pairs = num_sms() / 2;                                 // (upper bound: the schedule is chosen per launch)
        w.gemm_part = c.take<float>(flyp::dst_gemm_part_floats(w.gemm_pairs));
        w.ds_keep = c.take<uint16_t>((size_t)n_rows * w.ds_ld);
    }
    w.bytes = align_up(c.off, 256);
}

// out[n, :] = scale * out_mul / G * sum_m dS[m, n] x16[m, :] over the dS matrix the first sweep kept (n_m x n_n)
// Unfused backward (single rank, bf16, D <= 512, dS kept; OFF by default, FLYP_UNFUSED=1 switches it on): the dS kernel -
// the forward's tensor-core pipeline with an epilogue that writes dS and the d(scale) partials - then the two plain
// products dS . T and dS^T . I.  The idea: the fused sweep does its S product with 64 rows per SM (the accumulators of
// 128 rows x D take an SM's whole tensor memory) and runs at 71 % of burst where plain products run at 80 - 85 %.
// Measured (B = 32768, D = 512): the two products take 0.82 ms each as expected, but the dS kernel takes 1.41 ms against
// 0.76 ms for the forward - its epilogue (exponential, weights, fp16 packing, transposition, 32 KiB of stores per tile)
// has to keep up with a full-rate 128 x 128 S tile every ~2000 clocks, twice the element rate the fused sweep's epilogue
// sustains; the step is 3.93 ms against 3.49 ms fused.  Kept, tested, as the measured alternative.
int env_unfused() { static const int v = env_int("FLYP_UNFUSED", 0); return v; }
static bool unfused_eligible(const ClipWs& w, int n_rows, int n_cols, int dim) {
    return env_unfused() != 0 && w.ds_keep != nullptr && n_rows == n_cols && dim <= 512 && w.stats.use_mc;
}
// io: the fields of the d-image sweep (A = image rows, B = text rows, the row / column vectors)
static int run_ds_kernel(const SweepIO& io, const ClipWs& w, int* n_parts, cudaStream_t st) {
    CUtensorMap tmA64, tmB;
    int rc;
    if ((rc = make_tmap(&tmA64, io.A, io.n_m, io.dim, io.dim, false, 64)) != 0) return rc;
    if ((rc = make_tmap(&tmB, io.B, io.n_n, io.dim, io.dim)) != 0) return rc;
    flyp::FwdParams p;
    memset(&p, 0, sizeof(p));
    p.n_m = io.n_m; p.n_n = io.n_n; p.kc = ceil_div(io.dim, flyp::KCHUNK); p.kplan = flyp::kplan_bf16();
    p.m_tiles = ceil_div(io.n_m, flyp::TILE); p.n_tiles = ceil_div(io.n_n, flyp::TILE);
    p.scale = io.scale; p.wait_b = to_wait(io.b_ready);
    flyp::BwdParams b;
    memset(&b, 0, sizeof(b));
    b.n_m = io.n_m; b.n_n = io.n_n; b.scale = io.scale;
    b.wr = io.wr; b.lr = io.lr; b.wc = io.wc; b.lc = io.lc; b.labr = io.labr; b.dr = io.dr; b.labc = io.labc; b.dc = io.dc;
    b.fa = io.fa; b.fb = io.fb; b.fast_info = io.fast_info; b.gmax_bits = io.ctrl->words;
    b.dscale_part = io.dscale_part; b.ds_out = w.ds_keep; b.ds_ld = w.ds_ld;
    const bool timed = g_ev_sweep[0] != nullptr && g_ev_sweep_idx == 0;
    if (timed) cudaEventRecord(g_ev_sweep[0], st);
    const int n = flyp::launch_ds_mc(tmA64, tmB, p, b, num_sms(), st);
    if (timed) cudaEventRecord(g_ev_sweep[1], st);
    CUDA_OK(cudaGetLastError());
    if ((size_t)n > io.n_dscale) return fail(FLYP_ERR_WORKSPACE, "d(scale) partial slots");
    *n_parts = n;
    return 0;
}

static int run_dst_gemm(const ClipWs& w, int n_m, int n_n, int dim, const void* x16, const float* scale, float out_mul,
                        void* out, int out_fp32, cudaStream_t st, float* const* out_rank, int rows_per_rank, int transposed,
                        int counter, int col_begin = 0);
// dS kernel, d(scale), d_img = s dS T, d_txt = s dS^T I.  io: the d-image sweep's fields; i16: fp16 copy of the image rows
static int run_unfused_backward(const SweepIO& io, const ClipWs& w, const void* i16, void* d_img, void* d_txt,
                                float* d_scale, cudaStream_t st) {
    int rc, n_parts = 0;
    if ((rc = run_ds_kernel(io, w, &n_parts, st)) != 0) return rc;
    if (d_scale != nullptr) {
        flyp::launch_sum_parts(io.dscale_part, n_parts, d_scale, nullptr, st);
        CUDA_OK(cudaGetLastError());
    }
    if ((rc = run_dst_gemm(w, io.n_m, io.n_n, io.dim, io.B_f16, io.scale, io.out_mul, d_img, io.out_fp32, st, nullptr, 0,
                           /*transposed=*/0, /*counter=*/0)) != 0) return rc;
    return run_dst_gemm(w, io.n_m, io.n_n, io.dim, i16, io.scale, io.out_mul, d_txt, io.out_fp32, st, nullptr, 0, 1, 1);
}

// Wide problems (dim > 512) that keep dS: the sweep covered the output columns [0, pair_d_half(dim)); the others are the
// product dS . x16[:, d_half:] - S is recomputed ONCE per sweep, not once per 512 output columns (FLYP_WIDE_PRODUCT=0: A/B)
static bool wide_product(int dim) {
    static const int on = env_int("FLYP_WIDE_PRODUCT", 1);
    return on != 0 && pair_n_dh(dim) > 1;
}
static int run_wide_rest(const SweepIO& io, const ClipWs& w, cudaStream_t st) {
    return run_dst_gemm(w, io.n_m, io.n_n, io.dim, io.B_f16, io.scale, io.out_mul, io.out, io.out_fp32, st, nullptr, 0,
                        /*transposed=*/0, /*counter=*/2, /*col_begin=*/pair_d_half(io.dim));
}

// transposed = 0: out[m, :] = scale * out_mul / G * sum_n dS[m, n] x16[n, :] (x16 has n_n rows)
static int run_dst_gemm(const ClipWs& w, int n_m, int n_n, int dim, const void* x16, const float* scale, float out_mul,
                        void* out, int out_fp32, cudaStream_t st, float* const* out_rank, int rows_per_rank,
                        int transposed, int counter, int col_begin) {
    CUtensorMap tmDS, tmX;
    int rc;
    const int n_k = transposed ? n_m : n_n, n_o = transposed ? n_n : n_m;
    if ((rc = make_tmap(&tmDS, w.ds_keep, n_m, n_n, w.ds_ld, true)) != 0) return rc;
    if ((rc = make_tmap(&tmX, x16, n_k, dim, dim, true)) != 0) return rc;
    flyp::DstParams p;
    memset(&p, 0, sizeof(p));
    p.n_k = n_k; p.n_out = n_o; p.dim = dim; p.col_begin = col_begin;
    // partials that go to other GPUs: 256-column tiles in two accumulator stages, so that the NVLink-bound drain of a
    // tile overlaps the MMAs of the next (dS is then read ceil(dim / 256) times - from this rank's HBM, cheap beside it)
    p.tile_cols = out_rank != nullptr ? 256 : flyp::DST_TILE_COLS;
    if (env_int("FLYP_GEMM_TILE_COLS", 0) == 256 || env_int("FLYP_GEMM_TILE_COLS", 0) == 512)      // A/B switch
        p.tile_cols = env_int("FLYP_GEMM_TILE_COLS", 0);
    p.out_tiles = ceil_div(n_o, flyp::DST_TILE_ROWS); p.n_dh = ceil_div(dim - col_begin, p.tile_cols);
    p.sched_pairs = flyp::dst_gemm_sched_pairs(p.out_tiles * p.n_dh, ceil_div(n_k, 128), num_sms());
    if (p.sched_pairs > w.gemm_pairs) return fail(FLYP_ERR_WORKSPACE, "partial-tile scratch of the dS product");
    p.transposed = transposed;
    p.scale = scale; p.gmax_bits = w.ctrl.words; p.out_mul = out_mul;
    p.out = out; p.ld_out = dim; p.out_fp32 = out_fp32;
    if (out_rank != nullptr) {
        if (n_o % rows_per_rank != 0 || n_o / rows_per_rank > flyp::PEER_MAXW) return fail(FLYP_ERR_ARG, "bad rank split");
        for (int q = 0; q < n_o / rows_per_rank; ++q) p.out_rank[q] = out_rank[q];
        p.rows_per_rank = rows_per_rank; p.out_fp32 = 1; p.out = nullptr;
    }
    // (no flat tail when the pairs divide the tiles: no partial tiles, no grid barrier, an ordinary launch)
    p.part_out = w.gemm_part;
    p.grid_cnt = (p.out_tiles * p.n_dh) % p.sched_pairs == 0 ? nullptr : w.ctrl.grid_cnt(counter);
    const bool timed = g_ev_sweep[0] != nullptr && g_ev_sweep_idx == (transposed ? 1 : 2) && col_begin == 0;
    if (timed) cudaEventRecord(g_ev_sweep[0], st);
    flyp::launch_dst_gemm(tmDS, tmX, p, st);
    if (timed) cudaEventRecord(g_ev_sweep[1], st);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int flyp_clip_workspace_bytes(int n_rows, int n_cols, int dim, int dtype, size_t* bytes) {
    int rc = check_common(n_rows, n_cols, dim, dtype);
    if (rc) return rc;
    if (!bytes) return fail(FLYP_ERR_ARG, "bytes is null");
    ClipWs w;
    carve_clip(nullptr, n_rows, n_cols, dim, dtype, w);
    *bytes = w.bytes;
    return 0;
}

int flyp_clip_keeps_ds(int n_rows, int n_cols, int dim, int dtype) {
    if (check_common(n_rows, n_cols, dim, dtype) != 0) return 0;
    return keep_ds_eligible(n_rows, n_cols, dim, dtype) ? 1 : 0;
}

int flyp_clip_backward_plan(int n_rows, int n_cols, int dim, int dtype) {
    if (check_common(n_rows, n_cols, dim, dtype) != 0) return 0;
    if (!keep_ds_eligible(n_rows, n_cols, dim, dtype)) return 0;
    ClipWs w;
    carve_clip(nullptr, n_rows, n_cols, dim, dtype, w);
    w.ds_keep = reinterpret_cast<uint16_t*>(1);                    // (size query: only "is it there" matters)
    return unfused_eligible(w, n_rows, n_cols, dim) ? 2 : 1;
}

int flyp_clip_fwd_local(const void* img, const void* txt, const float* scale, int n_rows, int n_cols, int dim,
                        int dtype, int row_offset, float* row_lse, float* row_nll, float* col_stat, int* status,
                        void* workspace, size_t workspace_bytes, void* stream) {
    return flyp_clip_fwd_local_ex(img, txt, scale, n_rows, n_cols, dim, dtype, row_offset, row_lse, row_nll, col_stat,
                                  status, workspace, workspace_bytes, nullptr, stream);
}

int flyp_clip_fwd_local_ex(const void* img, const void* txt, const float* scale, int n_rows, int n_cols, int dim,
                           int dtype, int row_offset, float* row_lse, float* row_nll, float* col_stat, int* status,
                           void* workspace, size_t workspace_bytes, const flyp_ready_t* txt_ready, void* stream) {
    int rc = check_common(n_rows, n_cols, dim, dtype);
    if (rc) return rc;
    if (!img || !txt || !scale || !row_lse || !row_nll || !col_stat || !workspace)
        return fail(FLYP_ERR_ARG, "null pointer argument");
    if (row_offset < 0 || row_offset + n_rows > n_cols)
        return fail(FLYP_ERR_ARG, "row_offset %d + n_rows %d exceeds n_cols %d", row_offset, n_rows, n_cols);
    ClipWs w;
    carve_clip(workspace, n_rows, n_cols, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    StatsIO io;
    memset(&io, 0, sizeof(io));
    io.A = img; io.B = txt; io.scale = scale; io.n_m = n_rows; io.n_n = n_cols; io.dim = dim; io.dtype = dtype;
    io.pos_offset = row_offset; io.row_lse = row_lse; io.row_nll = row_nll; io.col_stat = col_stat; io.status = status;
    io.b_ready = txt_ready;
    return run_stats(io, w.stats, static_cast<cudaStream_t>(stream));
}

int flyp_clip_fwd_finish(const float* col_stat_all, int world, const float* row_nll, int n_rows, int n_cols,
                         int row_offset, float* col_lse, float* col_nll, float* loss, void* stream) {
    return flyp_clip_fwd_finish_ex(col_stat_all, world, row_nll, n_rows, n_cols, row_offset, col_lse, col_nll, loss,
                                   FLYP_F32, nullptr, stream);
}

int flyp_clip_fwd_finish_ex(const float* col_stat_all, int world, const float* row_nll, int n_rows, int n_cols,
                            int row_offset, float* col_lse, float* col_nll, void* loss, int loss_dtype,
                            const flyp_ready_t* stats_ready, void* stream) {
    if (!col_stat_all || !row_nll || !col_lse || !col_nll || !loss) return fail(FLYP_ERR_ARG, "null pointer argument");
    if (world < 1 || n_rows <= 0 || n_cols <= 0) return fail(FLYP_ERR_ARG, "bad sizes");
    if (loss_dtype != FLYP_BF16 && loss_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad loss_dtype %d", loss_dtype);
    flyp::launch_clip_finish(col_stat_all, world, row_nll, n_rows, n_cols, row_offset, col_lse, col_nll, loss,
                             loss_dtype == FLYP_BF16, to_wait(stats_ready), static_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int flyp_clip_bwd_local(const void* img, const void* txt, const float* scale, int n_rows, int n_cols, int dim,
                        int dtype, int row_offset, const float* row_lse, const float* row_nll, const float* col_lse,
                        const float* col_nll, const float* g_row, const float* g_col, float grad_mul, int grad_dtype,
                        void* d_img, void* d_txt, float* d_scale, void* workspace, size_t workspace_bytes,
                        void* stream) {
    return flyp_clip_bwd_local_ex(img, txt, scale, n_rows, n_cols, dim, dtype, row_offset, row_lse, row_nll, col_lse,
                                  col_nll, g_row, g_col, grad_mul, grad_dtype, d_img, d_txt, d_scale, workspace,
                                  workspace_bytes, nullptr, nullptr, nullptr, stream);
}

// common fields of the two sweeps of a backward call
static SweepIO sweep_base(const float* scale, int dtype, int dim, const ClipWs& w, int grad_dtype, float grad_mul) {
    SweepIO io;
    memset(&io, 0, sizeof(io));
    io.scale = scale; io.dtype = dtype; io.dim = dim; io.fast_info = w.fast_info;
    io.out_fp32 = grad_dtype; io.out_mul = grad_mul;
    io.part_scratch = w.part_scratch; io.ctrl = &w.ctrl; io.n_dscale = w.n_dscale;
    return io;
}

int flyp_clip_bwd_local_ex(const void* img, const void* txt, const float* scale, int n_rows, int n_cols, int dim,
                           int dtype, int row_offset, const float* row_lse, const float* row_nll, const float* col_lse,
                           const float* col_nll, const float* g_row, const float* g_col, float grad_mul, int grad_dtype,
                           void* d_img, void* d_txt, float* d_scale, void* workspace, size_t workspace_bytes,
                           const void* txt16, const flyp_ready_t* txt_ready, const flyp_ready_t* txt16_ready,
                           void* stream) {
    int rc = check_common(n_rows, n_cols, dim, dtype);
    if (rc) return rc;
    if (!img || !txt || !scale || !row_lse || !row_nll || !col_lse || !col_nll || !g_row || !g_col || !workspace)
        return fail(FLYP_ERR_ARG, "null pointer argument");
    if (grad_dtype != FLYP_BF16 && grad_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad grad_dtype %d", grad_dtype);
    if (row_offset < 0 || row_offset + n_rows > n_cols)
        return fail(FLYP_ERR_ARG, "row_offset %d + n_rows %d exceeds n_cols %d", row_offset, n_rows, n_cols);
    const bool f32 = dtype == FLYP_F32;
    if (f32 && grad_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "fp32 features need grad_dtype = FLYP_F32");
    if (txt16 != nullptr && f32) return fail(FLYP_ERR_ARG, "a precomputed fp16 copy is only accepted for bf16 features");
    if (d_scale && !d_img) return fail(FLYP_ERR_ARG, "d_scale requires d_img");
    ClipWs w;
    carve_clip(workspace, n_rows, n_cols, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rp = ceil_div(n_rows, VEC_PAD) * VEC_PAD, cp = ceil_div(n_cols, VEC_PAD) * VEC_PAD;
    CUDA_OK(cudaMemsetAsync(w.ctrl.words, 0, CTRL_WORDS * sizeof(uint32_t), st));
    // rows: w = g_row/2, positive column row_offset + i, exact dS there from the saved cross-entropies
    flyp::launch_bwd_prep(n_rows, rp, g_row, 0.5f, row_lse, row_nll, nullptr, row_offset, n_cols, g_col, col_nll, 0.5f,
                          w.rows.w, w.rows.l2, w.rows.lab, w.rows.d, w.ctrl.words, st);
    // columns: w = g_col/2, positive local row j - row_offset
    flyp::launch_bwd_prep(n_cols, cp, g_col, 0.5f, col_lse, col_nll, nullptr, -row_offset, n_rows, g_row, row_nll, 0.5f,
                          w.cols.w, w.cols.l2, w.cols.lab, w.cols.d, w.ctrl.words, st);
    flyp::launch_bwd_fast_vectors(w.ctrl.words, rp, w.rows.w, w.rows.l2, w.rows.f, cp, w.cols.w, w.cols.l2, w.cols.f,
                                  w.fast_info, st);
    CUDA_OK(cudaGetLastError());
    const int dp = plane_cols(dim);
    if (f32) {
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(img), n_rows, dim, dp, w.stats.planes_a, st);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(txt), n_cols, dim, dp, w.stats.planes_b, st);
        CUDA_OK(cudaGetLastError());
    }
    const bool keep = w.ds_keep != nullptr && d_img && d_txt;      // second gradient from the kept dS (clip_dst_gemm.cu)
    if (d_img) {
        const void* t16 = txt16;
        if (t16 == nullptr) {
            if (f32) flyp::launch_split_planes_f16x2(static_cast<const float*>(txt), n_cols, dim, dp, w.txt16, st);
            else flyp::launch_to_f16(txt, dtype, (size_t)n_cols * dim, w.txt16, st);
            CUDA_OK(cudaGetLastError());
            t16 = w.txt16;
        }
        SweepIO io = sweep_base(scale, dtype, dim, w, grad_dtype, grad_mul);
        io.A = img; io.B = txt; io.B_f16 = t16; io.A_planes = w.stats.planes_a; io.B_planes = w.stats.planes_b;
        io.n_m = n_rows; io.n_n = n_cols;
        io.wr = w.rows.w; io.lr = w.rows.l2; io.wc = w.cols.w; io.lc = w.cols.l2; io.labr = w.rows.lab; io.dr = w.rows.d;
        io.fa = w.rows.f; io.fb = w.cols.f;
        io.out = d_img; io.sweep = 0;
        if (d_scale) { io.dscale_part = w.dscale_part; io.dscale_out = d_scale; }
        io.b_ready = txt_ready; io.b16_ready = txt16_ready;
        if (keep) { io.ds_keep = w.ds_keep; io.ds_ld = w.ds_ld; }
        if (keep && unfused_eligible(w, n_rows, n_cols, dim) && txt_ready == nullptr) {
            flyp::launch_to_f16(img, dtype, (size_t)n_rows * dim, w.img16, st);
            CUDA_OK(cudaGetLastError());
            return run_unfused_backward(io, w, w.img16, d_img, d_txt, d_scale, st);
        }
        if (keep && wide_product(dim)) io.first_pass_only = 1;
        if ((rc = run_sweep(io, st)) != 0) return rc;
        if (io.first_pass_only && (rc = run_wide_rest(io, w, st)) != 0) return rc;
    }
    if (d_txt && keep) {
        flyp::launch_to_f16(img, dtype, (size_t)n_rows * dim, w.img16, st);
        CUDA_OK(cudaGetLastError());
        if ((rc = run_dst_gemm(w, n_rows, n_cols, dim, w.img16, scale, grad_mul, d_txt, grad_dtype, st, nullptr, 0, 1, 1)) != 0) return rc;
    } else if (d_txt) {
        if (f32) flyp::launch_split_planes_f16x2(static_cast<const float*>(img), n_rows, dim, dp, w.img16, st);
        else flyp::launch_to_f16(img, dtype, (size_t)n_rows * dim, w.img16, st);
        CUDA_OK(cudaGetLastError());
        SweepIO io = sweep_base(scale, dtype, dim, w, grad_dtype, grad_mul);
        io.A = txt; io.B = img; io.B_f16 = w.img16; io.A_planes = w.stats.planes_b; io.B_planes = w.stats.planes_a;
        io.n_m = n_cols; io.n_n = n_rows;
        io.wr = w.cols.w; io.lr = w.cols.l2; io.wc = w.rows.w; io.lc = w.rows.l2; io.labr = w.cols.lab; io.dr = w.cols.d;
        io.fa = w.cols.f; io.fb = w.rows.f;
        io.out = d_txt; io.sweep = 1;
        if ((rc = run_sweep(io, st)) != 0) return rc;
    }
    return 0;
}

// Both sweeps of a rank of the row-sharded symmetric loss (world == 1: the whole loss).  comm (may be null): the
// d(scale) partial is published to the other ranks by the first sweep's last CTA and the W partials are summed after
// the second sweep.
static int bwd_sharded_impl(const void* img, const void* txt, const void* img_all, const void* txt_all,
                            const void* img16_all, const void* txt16_all, const float* scale, int n_rows, int n_cols,
                            int dim, int dtype, int row_offset, const float* row_lse_all, const float* row_nll_all,
                            const float* col_lse, const float* col_nll, const void* g, int g_dtype, float grad_mul,
                            int grad_dtype, void* d_img, void* d_txt, float* d_scale, void* workspace,
                            size_t workspace_bytes, const flyp_ready_t* img_ready, const flyp_ready_t* txt_ready,
                            const flyp_ready_t* img16_ready, const flyp_ready_t* txt16_ready, void* stream,
                            flyp_comm* comm, uint32_t seq, float* d_scale_total, int phases = 3) {
    int rc = check_common(n_rows, n_cols, dim, dtype);
    if (rc) return rc;
    if (g_dtype != FLYP_BF16 && g_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad g_dtype %d", g_dtype);
    const bool f32 = dtype == FLYP_F32;
    if (f32 && grad_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "fp32 features need grad_dtype = FLYP_F32");
    if (f32 && (img16_all || txt16_all)) return fail(FLYP_ERR_ARG, "fp16 copies are only accepted for bf16 features");
    if (!img || !txt || !scale || !row_lse_all || !row_nll_all || !col_lse || !col_nll || !g || !workspace)
        return fail(FLYP_ERR_ARG, "null pointer argument");
    if ((d_img || d_scale) && !txt_all) return fail(FLYP_ERR_ARG, "d_img needs the gathered text features");
    if (d_txt && !img_all) return fail(FLYP_ERR_ARG, "d_txt needs the gathered image features");
    if (d_scale && !d_img) return fail(FLYP_ERR_ARG, "d_scale requires d_img");
    if (grad_dtype != FLYP_BF16 && grad_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad grad_dtype %d", grad_dtype);
    if (row_offset < 0 || row_offset + n_rows > n_cols)
        return fail(FLYP_ERR_ARG, "row_offset %d + n_rows %d exceeds n_cols %d", row_offset, n_rows, n_cols);
    ClipWs w;
    carve_clip(workspace, n_rows, n_cols, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cp = ceil_div(n_cols, VEC_PAD) * VEC_PAD;
    // Kept-dS backward (carve_clip reserved the dS matrix): the second gradient is a product over the first sweep's dS.
    //   single rank (square problem): d_txt = s dS^T I directly;
    //   row-sharded over a communicator: every rank's partial s dS_r^T I_r of ALL text rows is scattered, by the product
    //   kernel itself, into the owners' reduce-scatter buffers over NVLink and summed there (keep_rs).
    const int world = flyp::comm_world(comm);
    const bool keep = w.ds_keep != nullptr && comm == nullptr && n_rows == n_cols && d_img && d_txt && img_all != nullptr;
    static const int rs_on = env_int("FLYP_KEEP_DS_RS", 1);       // A/B switch: 0 = two sweeps on several GPUs
    const bool keep_rs = rs_on != 0 && w.ds_keep != nullptr && comm != nullptr && world > 1 && n_cols == n_rows * world &&
                         d_img && d_txt && n_rows >= flyp::comm_rs_min_rows(comm);
    if ((phases & 1) == 0) {
        // finish phase only: everything below was enqueued by an earlier call with phase 1
        if (keep_rs && (rc = flyp::comm_rs_reduce(comm, seq, n_rows, dim, d_txt, grad_dtype, grad_mul, stream)) != 0) return rc;
        if (comm != nullptr && d_scale != nullptr && d_scale_total != nullptr)
            return flyp_comm_sum_scalar(comm, seq, d_scale_total, stream);
        return 0;
    }
    // global vectors live in the column set (its d / lab arrays hold the row-statistics l2 / f), local ones in the row set
    float *wg = w.cols.w, *l2c = w.cols.l2, *fc = w.cols.f, *l2r = w.cols.d, *fr = reinterpret_cast<float*>(w.cols.lab);
    CUDA_OK(cudaMemsetAsync(w.ctrl.words, 0, CTRL_WORDS * sizeof(uint32_t), st));
    flyp::launch_bwd_prep_sharded(n_cols, cp, row_offset, n_rows, g, g_dtype == FLYP_BF16, row_lse_all, row_nll_all, col_lse, col_nll, wg, l2c,
                                  l2r, w.rows.lab, w.rows.d, w.ctrl.words, st);
    flyp::launch_bwd_fast_vectors(w.ctrl.words, cp, wg, l2c, fc, cp, wg, l2r, fr, w.fast_info, st);
    CUDA_OK(cudaGetLastError());
    const int off = row_offset;
    const int dp = plane_cols(dim);
    flyp::PeerPush push;
    memset(&push, 0, sizeof(push));
    if (comm != nullptr && d_scale != nullptr && (rc = flyp::comm_scalar_push_target(comm, seq, &push)) != 0) return rc;
    if (d_img) {
        // image rows of this rank against all texts: complete d_img and this row block's share of d(scale)
        const void* t16 = txt16_all;
        if (f32) {
            // split planes of the operands of this sweep (the gathered rows have arrived once the split kernel's own
            // wait on their flags is over: the sweep itself then needs no flags)
            const flyp::PeerWait wt = to_wait(txt_ready);
            flyp::launch_split_planes_bf16x3(static_cast<const float*>(img), n_rows, dim, dp, w.stats.planes_a, st);
            flyp::launch_split_planes_bf16x3(static_cast<const float*>(txt_all), n_cols, dim, dp, w.stats.planes_b, st, &wt);
            flyp::launch_split_planes_f16x2(static_cast<const float*>(txt_all), n_cols, dim, dp, w.txt16, st, &wt);
            CUDA_OK(cudaGetLastError());
            t16 = w.txt16;
        } else if (t16 == nullptr) {             // the caller kept no fp16 copy: make one
            flyp::launch_to_f16(txt_all, dtype, (size_t)n_cols * dim, w.txt16, st);
            CUDA_OK(cudaGetLastError());
            t16 = w.txt16;
        }
        SweepIO io = sweep_base(scale, dtype, dim, w, grad_dtype, grad_mul);
        io.A = img; io.B = txt_all; io.B_f16 = t16; io.n_m = n_rows; io.n_n = n_cols;
        io.A_planes = w.stats.planes_a; io.B_planes = w.stats.planes_b;
        io.wr = wg + off; io.lr = l2r + off; io.wc = wg; io.lc = l2c; io.labr = w.rows.lab; io.dr = w.rows.d;
        io.fa = fr + off; io.fb = fc;
        io.out = d_img; io.sweep = 0;
        if (d_scale) { io.dscale_part = w.dscale_part; io.dscale_out = d_scale; io.ds_push = &push; }
        if (!f32) { io.b_ready = txt_ready; io.b16_ready = txt16_all ? txt16_ready : nullptr; }
        if (keep || keep_rs) { io.ds_keep = w.ds_keep; io.ds_ld = w.ds_ld; }
        if (keep && unfused_eligible(w, n_rows, n_cols, dim)) {
            const void* i16 = img16_all;
            if (i16 == nullptr) {
                flyp::launch_to_f16(img_all, dtype, (size_t)n_cols * dim, w.img16, st);
                CUDA_OK(cudaGetLastError());
                i16 = w.img16;
            }
            return run_unfused_backward(io, w, i16, d_img, d_txt, d_scale, st);
        }
        if ((keep || keep_rs) && wide_product(dim)) io.first_pass_only = 1;
        if ((rc = run_sweep(io, st)) != 0) return rc;
        if (io.first_pass_only && (rc = run_wide_rest(io, w, st)) != 0) return rc;
    }
    if (d_txt && keep_rs) {
        // this rank's partial of every text row's gradient, written into the owners' buffers by the product kernel
        float* out_rank[FLYP_COMM_MAX_WORLD];
        if ((rc = flyp::comm_rs_targets(comm, seq, n_rows, dim, out_rank)) != 0) return rc;
        const void* i16 = img16_all != nullptr ? static_cast<const uint16_t*>(img16_all) + (size_t)off * dim : nullptr;
        if (i16 == nullptr) {                         // (own rows: nothing to wait for)
            flyp::launch_to_f16(img, dtype, (size_t)n_rows * dim, w.img16, st);
            CUDA_OK(cudaGetLastError());
            i16 = w.img16;
        }
        // (grad_mul is the RECEIVER's factor - gather_with_grad scales the gradients of a rank's own rows - applied by
        // the sum)
        if ((rc = run_dst_gemm(w, n_rows, n_cols, dim, i16, scale, 1.0f, nullptr, 1, st, out_rank, n_rows, 1, 1)) != 0) return rc;
        if ((rc = flyp::comm_rs_signal(comm, seq, stream)) != 0) return rc;
        if ((phases & 2) != 0 &&
            (rc = flyp::comm_rs_reduce(comm, seq, n_rows, dim, d_txt, grad_dtype, grad_mul, stream)) != 0) return rc;
    } else if (d_txt && keep) {
        // single rank: d_txt = s dS^T I over the dS values the first sweep kept - no second recompute of the logits
        const void* i16 = img16_all;
        if (i16 == nullptr) {
            flyp::launch_to_f16(img_all, dtype, (size_t)n_cols * dim, w.txt16, st);
            CUDA_OK(cudaGetLastError());
            i16 = w.txt16;
        }
        if ((rc = run_dst_gemm(w, n_rows, n_cols, dim, i16, scale, grad_mul, d_txt, grad_dtype, st, nullptr, 0, 1, 1)) != 0) return rc;
    } else if (d_txt) {
        // the transposed problem: text rows of this rank against all images (no B x D reduce-scatter)
        const void* i16 = img16_all;
        if (f32) {                               // (the first sweep is done with the plane buffers: stream order)
            const flyp::PeerWait wi = to_wait(img_ready);
            flyp::launch_split_planes_bf16x3(static_cast<const float*>(txt), n_rows, dim, dp, w.stats.planes_a, st);
            flyp::launch_split_planes_bf16x3(static_cast<const float*>(img_all), n_cols, dim, dp, w.stats.planes_b, st, &wi);
            flyp::launch_split_planes_f16x2(static_cast<const float*>(img_all), n_cols, dim, dp, w.txt16, st, &wi);
            CUDA_OK(cudaGetLastError());
            i16 = w.txt16;
        } else if (i16 == nullptr) {
            flyp::launch_to_f16(img_all, dtype, (size_t)n_cols * dim, w.txt16, st);
            CUDA_OK(cudaGetLastError());
            i16 = w.txt16;
        }
        SweepIO io = sweep_base(scale, dtype, dim, w, grad_dtype, grad_mul);
        io.A = txt; io.B = img_all; io.B_f16 = i16; io.n_m = n_rows; io.n_n = n_cols;
        io.A_planes = w.stats.planes_a; io.B_planes = w.stats.planes_b;
        io.wr = wg + off; io.lr = l2c + off; io.wc = wg; io.lc = l2r; io.labr = w.rows.lab; io.dr = w.rows.d;
        io.fa = fc + off; io.fb = fr;
        io.out = d_txt; io.sweep = 1;
        if (!f32) { io.b_ready = img_ready; io.b16_ready = img16_all ? img16_ready : nullptr; }
        if ((rc = run_sweep(io, st)) != 0) return rc;
    }
    if ((phases & 2) != 0 && comm != nullptr && d_scale != nullptr && d_scale_total != nullptr)
        return flyp_comm_sum_scalar(comm, seq, d_scale_total, stream);
    return 0;
}

int flyp_clip_bwd_sharded(const void* img, const void* txt, const void* img_all, const void* txt_all,
                          const void* img16_all, const void* txt16_all, const float* scale, int n_rows, int n_cols,
                          int dim, int dtype, int row_offset, const float* row_lse_all, const float* row_nll_all,
                          const float* col_lse, const float* col_nll, const void* g, int g_dtype, float grad_mul,
                          int grad_dtype, void* d_img, void* d_txt, float* d_scale, void* workspace,
                          size_t workspace_bytes, const flyp_ready_t* img_ready, const flyp_ready_t* txt_ready,
                          const flyp_ready_t* img16_ready, const flyp_ready_t* txt16_ready, void* stream) {
    return bwd_sharded_impl(img, txt, img_all, txt_all, img16_all, txt16_all, scale, n_rows, n_cols, dim, dtype,
                            row_offset, row_lse_all, row_nll_all, col_lse, col_nll, g, g_dtype, grad_mul, grad_dtype,
                            d_img, d_txt, d_scale, workspace, workspace_bytes, img_ready, txt_ready, img16_ready,
                            txt16_ready, stream, nullptr, 0, nullptr);
}

// ---- whole-step entry points: what ClipLoss.forward / backward of a rank call ---------------------------------------
int flyp_clip_fwd_step(flyp_comm* comm, const void* img, const void* txt, const float* scale, int n_rows, int dim,
                       int dtype, int rank, int world, float* row_lse, float* row_nll, float* col_stat, float* col_lse,
                       float* col_nll, void* loss, int loss_dtype, void* feat16, int* status, void* workspace,
                       size_t workspace_bytes, flyp_step_t* step, void* stream) {
    if (!step) return fail(FLYP_ERR_ARG, "null pointer argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(FLYP_ERR_ARG, "bad rank %d / world %d", rank, world);
    if (!comm && world != 1) return fail(FLYP_ERR_ARG, "world %d needs a communicator", world);
    if (loss_dtype != FLYP_BF16 && loss_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad loss_dtype %d", loss_dtype);
    const int n_cols = n_rows * world, off = rank * n_rows;
    int rc = check_common(n_rows, n_cols, dim, dtype);
    if (rc) return rc;
    if (!img || !txt || !scale || !row_lse || !row_nll || !col_lse || !col_nll || !loss || !workspace)
        return fail(FLYP_ERR_ARG, "null pointer argument");
    ClipWs w;
    carve_clip(workspace, n_rows, n_cols, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StatsIO io;
    memset(&io, 0, sizeof(io));
    io.A = img; io.scale = scale; io.n_m = n_rows; io.n_n = n_cols; io.dim = dim; io.dtype = dtype;
    io.pos_offset = off; io.row_lse = row_lse; io.row_nll = row_nll; io.status = status;
    io.col_stat = col_stat ? col_stat : w.col_stat;
    if (comm == nullptr) {
        // single rank: no exchange; the finalize kernels finish the loss themselves
        memset(step, 0, sizeof(*step));
        step->gathered.img_all = img; step->gathered.txt_all = txt;
        if (feat16 != nullptr && dtype == FLYP_BF16) {
            io.a16 = feat16;
            io.b16 = static_cast<uint16_t*>(feat16) + (size_t)n_rows * dim;
            step->gathered.img16_all = io.a16; step->gathered.txt16_all = io.b16;
        }
        step->stats.col_stat_all = io.col_stat; step->stats.row_lse_all = row_lse; step->stats.row_nll_all = row_nll;
        io.B = txt;
        io.fin.col_lse = col_lse; io.fin.col_nll = col_nll; io.fin.loss = loss; io.fin.loss_bf16 = loss_dtype == FLYP_BF16;
        return run_stats(io, w.stats, st);
    }
    flyp::PackExtra ex;
    ex.scale = scale; ex.t2 = w.stats.t2; ex.pos = w.stats.pos; ex.n_pad = w.stats.ld_rows; ex.row_offset = off;
    ex.zero_words = w.stats.flag; ex.n_zero = 4;
    rc = flyp::comm_gather(comm, img, txt, n_rows, dim, dtype, &ex, &step->gathered, stream);
    if (rc) return rc;
    io.B = step->gathered.txt_all; io.b_ready = &step->gathered.txt_ready; io.pre_done = true;
    if ((rc = run_stats(io, w.stats, st)) != 0) return rc;
    rc = flyp_comm_push_stats(comm, step->gathered.seq, io.col_stat, row_lse, row_nll, n_rows, n_cols, &step->stats, stream);
    if (rc) return rc;
    return flyp_clip_fwd_finish_ex(step->stats.col_stat_all, world, step->stats.row_nll_all, n_cols, n_cols, 0, col_lse,
                                   col_nll, loss, loss_dtype, &step->stats.ready, stream);
}

int flyp_clip_bwd_step(flyp_comm* comm, const flyp_step_t* step, const void* img, const void* txt, const float* scale,
                       int n_rows, int dim, int dtype, int rank, int world, const float* col_lse, const float* col_nll,
                       const void* g, int g_dtype, float grad_mul, int grad_dtype, void* d_img, void* d_txt,
                       float* d_scale_partial, float* d_scale, void* workspace, size_t workspace_bytes, void* stream) {
    return flyp_clip_bwd_step_phase(comm, step, img, txt, scale, n_rows, dim, dtype, rank, world, col_lse, col_nll, g,
                                    g_dtype, grad_mul, grad_dtype, d_img, d_txt, d_scale_partial, d_scale, workspace,
                                    workspace_bytes, 3, stream);
}

int flyp_clip_bwd_step_phase(flyp_comm* comm, const flyp_step_t* step, const void* img, const void* txt,
                             const float* scale, int n_rows, int dim, int dtype, int rank, int world,
                             const float* col_lse, const float* col_nll, const void* g, int g_dtype, float grad_mul,
                             int grad_dtype, void* d_img, void* d_txt, float* d_scale_partial, float* d_scale,
                             void* workspace, size_t workspace_bytes, int phases, void* stream) {
    if (!step) return fail(FLYP_ERR_ARG, "null pointer argument");
    if (phases < 1 || phases > 3) return fail(FLYP_ERR_ARG, "phases %d (1: compute and publish, 2: finish, 3: both)", phases);
    if (world < 1 || rank < 0 || rank >= world) return fail(FLYP_ERR_ARG, "bad rank %d / world %d", rank, world);
    if (!comm && world != 1) return fail(FLYP_ERR_ARG, "world %d needs a communicator", world);
    const flyp_gathered_t& gg = step->gathered;
    if (comm == nullptr) {
        // single rank: the row block's share of d(scale) is the whole gradient
        return bwd_sharded_impl(img, txt, gg.img_all, gg.txt_all, gg.img16_all, gg.txt16_all, scale, n_rows, n_rows, dim,
                                dtype, 0, step->stats.row_lse_all, step->stats.row_nll_all, col_lse, col_nll, g, g_dtype,
                                grad_mul, grad_dtype, d_img, d_txt, d_scale, workspace, workspace_bytes, nullptr, nullptr,
                                nullptr, nullptr, stream, nullptr, 0, nullptr);
    }
    if ((d_scale != nullptr) != (d_scale_partial != nullptr))
        return fail(FLYP_ERR_ARG, "d_scale and d_scale_partial go together");
    return bwd_sharded_impl(img, txt, gg.img_all, gg.txt_all, gg.img16_all, gg.txt16_all, scale, n_rows, n_rows * world,
                            dim, dtype, rank * n_rows, step->stats.row_lse_all, step->stats.row_nll_all, col_lse, col_nll,
                            g, g_dtype, grad_mul, grad_dtype, d_img, d_txt, d_scale_partial, workspace, workspace_bytes,
                            &gg.img_ready, &gg.txt_ready, &gg.img16_ready, &gg.txt16_ready, stream, comm, gg.seq,
                            d_scale, phases);
}

// ------------------------------------------------------------------------------------------------ ce head
struct CeWs {
    StatsWs stats;
    VecSet v;
    float* dscale_part;
    size_t n_dscale;
    float* part_scratch;
    BwdCtrl ctrl;
    float* fast_info;
    uint16_t *a16, *b16;
    size_t bytes;
};
static void carve_ce(void* base, int n, int n_classes, int dim, int dtype, CeWs& w) {
    Carver c(base);
    carve_stats(c, n, n_classes, dim, dtype, false, w.stats);
    const int np = ceil_div(n, VEC_PAD) * VEC_PAD;
    carve_vecs(c, np, w.v);
    w.n_dscale = sweep_dscale_slots(n, n_classes, dim, dtype);
    w.dscale_part = c.take<float>(w.n_dscale);
    {
        const size_t a = sweep_part_floats(n, n_classes, dim, dtype), b = sweep_part_floats(n_classes, n, dim, dtype);
        const size_t nn = a > b ? a : b;
        w.part_scratch = nn ? c.take<float>(nn) : nullptr;
    }
    w.ctrl.words = c.take<uint32_t>(CTRL_WORDS);
    w.fast_info = c.take<float>(2);
    const size_t w16 = dtype == FLYP_F32 ? (size_t)2 * plane_cols(dim) : (size_t)dim;
    w.a16 = c.take<uint16_t>((size_t)n * w16);
    w.b16 = c.take<uint16_t>((size_t)n_classes * w16);
    w.bytes = align_up(c.off, 256);
}

int flyp_ce_workspace_bytes(int n, int n_classes, int dim, int dtype, size_t* bytes) {
    int rc = check_common(n, n_classes, dim, dtype);
    if (rc) return rc;
    if (!bytes) return fail(FLYP_ERR_ARG, "bytes is null");
    CeWs w;
    carve_ce(nullptr, n, n_classes, dim, dtype, w);
    *bytes = w.bytes;
    return 0;
}

int flyp_ce_fwd(const void* a, const void* b, const float* scale, int n, int n_classes, int dim, int dtype,
                const int64_t* labels, int label_offset, float* loss, float* lse, void* workspace,
                size_t workspace_bytes, void* stream) {
    return flyp_ce_fwd_ex(a, b, scale, n, n_classes, dim, dtype, labels, label_offset, loss, lse, workspace,
                          workspace_bytes, nullptr, stream);
}

int flyp_ce_fwd_ex(const void* a, const void* b, const float* scale, int n, int n_classes, int dim, int dtype,
                   const int64_t* labels, int label_offset, float* loss, float* lse, void* workspace,
                   size_t workspace_bytes, const flyp_ready_t* b_ready, void* stream) {
    int rc = check_common(n, n_classes, dim, dtype);
    if (rc) return rc;
    if (!a || !b || !scale || !loss || !lse || !workspace) return fail(FLYP_ERR_ARG, "null pointer argument");
    CeWs w;
    carve_ce(workspace, n, n_classes, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    StatsIO io;
    memset(&io, 0, sizeof(io));
    io.A = a; io.B = b; io.scale = scale; io.n_m = n; io.n_n = n_classes; io.dim = dim; io.dtype = dtype;
    io.labels = labels; io.pos_offset = label_offset; io.row_lse = lse; io.row_nll = loss; io.b_ready = b_ready;
    return run_stats(io, w.stats, static_cast<cudaStream_t>(stream));
}

int flyp_ce_bwd(const void* a, const void* b, const float* scale, int n, int n_classes, int dim, int dtype,
                const int64_t* labels, int label_offset, const float* lse, const float* loss, const float* g,
                int grad_dtype, void* d_a, void* d_b, float* d_scale, void* workspace, size_t workspace_bytes,
                void* stream) {
    return flyp_ce_bwd_ex(a, b, scale, n, n_classes, dim, dtype, labels, label_offset, lse, loss, g, grad_dtype, d_a, d_b,
                          d_scale, workspace, workspace_bytes, nullptr, nullptr, nullptr, stream);
}

int flyp_ce_bwd_ex(const void* a, const void* b, const float* scale, int n, int n_classes, int dim, int dtype,
                   const int64_t* labels, int label_offset, const float* lse, const float* loss, const float* g,
                   int grad_dtype, void* d_a, void* d_b, float* d_scale, void* workspace, size_t workspace_bytes,
                   const void* b16, const flyp_ready_t* b_ready, const flyp_ready_t* b16_ready, void* stream) {
    int rc = check_common(n, n_classes, dim, dtype);
    if (rc) return rc;
    if (!a || !b || !scale || !lse || !loss || !g || !workspace) return fail(FLYP_ERR_ARG, "null pointer argument");
    if (grad_dtype != FLYP_BF16 && grad_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad grad_dtype %d", grad_dtype);
    const bool f32 = dtype == FLYP_F32;
    if (f32 && grad_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "fp32 features need grad_dtype = FLYP_F32");
    if (f32 && b16 != nullptr) return fail(FLYP_ERR_ARG, "a precomputed fp16 copy is only accepted for bf16 features");
    if (d_scale && !d_a) return fail(FLYP_ERR_ARG, "d_scale requires d_a");
    CeWs w;
    carve_ce(workspace, n, n_classes, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int np = ceil_div(n, VEC_PAD) * VEC_PAD;
    CUDA_OK(cudaMemsetAsync(w.ctrl.words, 0, CTRL_WORDS * sizeof(uint32_t), st));
    // w = g, positive = label, dS there = g * expm1(-loss)
    flyp::launch_bwd_prep(n, np, g, 1.0f, lse, loss, labels, label_offset, n_classes, nullptr, nullptr, 1.0f, w.v.w,
                          w.v.l2, w.v.lab, w.v.d, w.ctrl.words, st);
    flyp::launch_bwd_fast_vectors(w.ctrl.words, np, w.v.w, w.v.l2, w.v.f, 0, nullptr, nullptr, nullptr, w.fast_info, st);
    CUDA_OK(cudaGetLastError());
    const int dp = plane_cols(dim);
    if (f32) {
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(a), n, dim, dp, w.stats.planes_a, st);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(b), n_classes, dim, dp, w.stats.planes_b, st);
        CUDA_OK(cudaGetLastError());
    }
    SweepIO base;
    memset(&base, 0, sizeof(base));
    base.scale = scale; base.dtype = dtype; base.dim = dim; base.fast_info = w.fast_info;
    base.out_fp32 = grad_dtype; base.out_mul = 1.0f;
    base.part_scratch = w.part_scratch; base.ctrl = &w.ctrl; base.n_dscale = w.n_dscale;
    if (d_a) {
        const void* bb16 = b16;
        if (bb16 == nullptr) {
            if (f32) flyp::launch_split_planes_f16x2(static_cast<const float*>(b), n_classes, dim, dp, w.b16, st);
            else flyp::launch_to_f16(b, dtype, (size_t)n_classes * dim, w.b16, st);
            CUDA_OK(cudaGetLastError());
            bb16 = w.b16;
        }
        SweepIO io = base;
        io.A = a; io.B = b; io.B_f16 = bb16; io.A_planes = w.stats.planes_a; io.B_planes = w.stats.planes_b;
        io.n_m = n; io.n_n = n_classes;
        io.wr = w.v.w; io.lr = w.v.l2; io.labr = w.v.lab; io.dr = w.v.d; io.fa = w.v.f;
        io.out = d_a; io.sweep = 0;
        if (d_scale) { io.dscale_part = w.dscale_part; io.dscale_out = d_scale; }
        io.b_ready = b_ready; io.b16_ready = b16 ? b16_ready : nullptr;
        if ((rc = run_sweep(io, st)) != 0) return rc;
    }
    if (d_b) {
        // rows = classes, columns = samples: only the column (sample) softmax term exists
        if (f32) flyp::launch_split_planes_f16x2(static_cast<const float*>(a), n, dim, dp, w.a16, st);
        else flyp::launch_to_f16(a, dtype, (size_t)n * dim, w.a16, st);
        CUDA_OK(cudaGetLastError());
        SweepIO io = base;
        io.A = b; io.B = a; io.B_f16 = w.a16; io.A_planes = w.stats.planes_b; io.B_planes = w.stats.planes_a;
        io.n_m = n_classes; io.n_n = n;
        io.wc = w.v.w; io.lc = w.v.l2; io.labc = w.v.lab; io.dc = w.v.d; io.fb = w.v.f;
        io.out = d_b; io.sweep = 1;
        // (rows of b that other ranks wrote were all consumed - and waited for - by the forward of the same step)
        if ((rc = run_sweep(io, st)) != 0) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ label-aware variants
// Building blocks of the label-aware ClipLoss variants (clip/loss.py:123-192: soft labels, `ignore`, `google_sup_loss`),
// orchestrated by flyp_b200/labeled.py.  Square problems (n x n), positives on the diagonal, one class id per item.
int flyp_label_stats(const void* a, const void* b, const float* scale, int n, int dim, int dtype, const int* cls_a,
                     const int* cls_b, int mode, const float* lse_rows, float* out0, float* out1, float* out2,
                     void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(n, n, dim, dtype);
    if (rc) return rc;
    if (!a || !b || !scale || !cls_a || !cls_b || !out0 || !out1 || !workspace) return fail(FLYP_ERR_ARG, "null pointer argument");
    if (mode != 1 && mode != 2) return fail(FLYP_ERR_ARG, "mode %d (1: exclude same-class entries, 2: accumulate over them)", mode);
    if (mode == 2 && (!lse_rows || !out2)) return fail(FLYP_ERR_ARG, "mode 2 needs lse_rows and out2");
    ClipWs w;
    carve_clip(workspace, n, n, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const StatsWs& s = w.stats;
    const int rp = ceil_div(n, VEC_PAD) * VEC_PAD;
    CUtensorMap tmA, tmB;
    flyp::KPlan kplan = flyp::kplan_bf16();
    if (dtype == FLYP_F32) {
        const int dp = plane_cols(dim);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(a), n, dim, dp, s.planes_a, st);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(b), n, dim, dp, s.planes_b, st);
        CUDA_OK(cudaGetLastError());
        if ((rc = make_tmap(&tmA, s.planes_a, n, 3 * dp, 3 * dp)) != 0) return rc;
        if ((rc = make_tmap(&tmB, s.planes_b, n, 3 * dp, 3 * dp)) != 0) return rc;
        kplan = flyp::kplan_f32(dp);
    } else {
        if ((rc = make_tmap(&tmA, a, n, dim, dim)) != 0) return rc;
        if ((rc = make_tmap(&tmB, b, n, dim, dim)) != 0) return rc;
    }
    // positives on the diagonal: t2 / pos; class ids padded with -1 / -2 (never equal); lse in log2 units
    flyp::launch_pair_dot(a, b, dtype, scale, n, s.ld_rows, n, dim, nullptr, 0, s.t2, s.pos, nullptr, s.flag, 4, nullptr, nullptr, st);
    flyp::launch_label_prep(n, rp, cls_a, cls_b, w.rows.lab, w.cols.lab, lse_rows, w.rows.l2, st);
    CUDA_OK(cudaGetLastError());
    flyp::FwdParams p;
    memset(&p, 0, sizeof(p));
    p.n_m = n; p.n_n = n; p.kc = ceil_div(dim, flyp::KCHUNK); p.kplan = kplan;
    p.m_tiles = s.m_tiles; p.n_tiles = s.n_tiles; p.n_slots = s.n_slots; p.ld_rows = s.ld_rows; p.ld_cols = s.ld_cols;
    p.scale = scale; p.shift_slack = shift_slack(n, n);
    p.rowpart = s.rowpart; p.rowmax = s.rowmax; p.pos = s.pos;
    p.cls_m = w.rows.lab; p.cls_n = w.cols.lab; p.mask_mode = mode; p.acc_lse = w.rows.l2; p.acc3 = s.rowpart2;
    flyp::launch_fwd(tmA, tmB, p, /*robust=*/true, nullptr, num_sms(), st);
    CUDA_OK(cudaGetLastError());
    if (mode == 1) {
        CUDA_OK(cudaMemsetAsync(s.flag, 1, sizeof(int), st));             // non-zero: the robust finalize does its work
        flyp::FwdFinish nofin;
        memset(&nofin, 0, sizeof(nofin));
        flyp::launch_fwd_finalize_robust(s.rowpart, s.rowmax, s.n_tiles * 2, s.ld_rows, n, nullptr, nullptr, 0, s.ld_cols, n,
                                         s.t2, s.pos, out0, out1, nullptr, s.flag, nullptr, nofin, st);
    } else {
        flyp::launch_label_sum3(s.rowpart, s.rowmax, s.rowpart2, s.n_tiles * 2, s.ld_rows, n, out0, out1, out2, st);
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int flyp_label_sweep(const void* a, const void* b, const float* scale, int n, int dim, int dtype, const float* wr,
                     const float* lr, const float* wc, const float* lc, const float* d_diag, const int* cls_a,
                     const int* cls_b, const float* mk_r, const float* mk_c, int mode, const float* gmax, int grad_dtype,
                     void* out, float* d_scale, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(n, n, dim, dtype);
    if (rc) return rc;
    if (!a || !b || !scale || !wr || !lr || !wc || !lc || !gmax || !out || !workspace)
        return fail(FLYP_ERR_ARG, "null pointer argument");
    if (mode < 0 || mode > 3) return fail(FLYP_ERR_ARG, "bad mode %d", mode);
    if (mode != 0 && (!cls_a || !cls_b)) return fail(FLYP_ERR_ARG, "class ids are required for mode %d", mode);
    if (mode >= 2 && (!mk_r || !mk_c)) return fail(FLYP_ERR_ARG, "mode %d needs the same-class weights", mode);
    if (grad_dtype != FLYP_BF16 && grad_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad grad_dtype %d", grad_dtype);
    const bool f32 = dtype == FLYP_F32;
    if (f32 && grad_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "fp32 features need grad_dtype = FLYP_F32");
    ClipWs w;
    carve_clip(workspace, n, n, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rp = ceil_div(n, VEC_PAD) * VEC_PAD;
    // vectors padded into the workspace (log2 units for the logsumexps), staging scale from the caller's bound on |dS|
    CUDA_OK(cudaMemsetAsync(w.ctrl.words, 0, CTRL_WORDS * sizeof(uint32_t), st));
    flyp::launch_label_sweep_prep(n, rp, wr, lr, wc, lc, d_diag, mk_r, mk_c, cls_a, cls_b, gmax, w.rows.w, w.rows.l2, w.rows.d,
                                  w.rows.f, w.rows.lab, w.cols.w, w.cols.l2, w.cols.f, w.cols.lab,
                                  reinterpret_cast<int*>(w.cols.d), w.ctrl.words, w.fast_info, st);
    CUDA_OK(cudaGetLastError());
    const int dp = plane_cols(dim);
    if (f32) {
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(a), n, dim, dp, w.stats.planes_a, st);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(b), n, dim, dp, w.stats.planes_b, st);
        flyp::launch_split_planes_f16x2(static_cast<const float*>(b), n, dim, dp, w.txt16, st);
    } else {
        flyp::launch_to_f16(b, dtype, (size_t)n * dim, w.txt16, st);
    }
    CUDA_OK(cudaGetLastError());
    SweepIO io = sweep_base(scale, dtype, dim, w, grad_dtype, 1.0f);
    io.A = a; io.B = b; io.B_f16 = w.txt16; io.A_planes = w.stats.planes_a; io.B_planes = w.stats.planes_b;
    io.n_m = n; io.n_n = n;
    io.wr = w.rows.w; io.lr = w.rows.l2; io.wc = w.cols.w; io.lc = w.cols.l2;
    if (d_diag != nullptr) { io.labr = reinterpret_cast<int*>(w.cols.d); io.dr = w.rows.d; }   // positive of row i: column i
    io.fa = nullptr; io.fb = nullptr;
    io.out = out; io.sweep = 0;
    if (d_scale) { io.dscale_part = w.dscale_part; io.dscale_out = d_scale; }
    if (mode != 0) { io.cls_m = w.rows.lab; io.cls_n = w.cols.lab; io.mk_r = w.rows.f; io.mk_c = w.cols.f; io.mask_mode = mode; }
    return run_sweep(io, st);
}

// ------------------------------------------------------------------------------------------------ encoder tail
int flyp_project_normalize_workspace_bytes(int n, int k, int n_out, int dtype, size_t* bytes) {
    if (!bytes) return fail(FLYP_ERR_ARG, "bytes is null");
    if (n <= 0 || k <= 0 || n_out <= 0) return fail(FLYP_ERR_ARG, "bad shape");
    if (dtype != FLYP_BF16 && dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad dtype %d", dtype);
    // fp32 inputs: three bf16 planes of x ([n][3 k]) and of W ([3][k][n_out])
    *bytes = dtype == FLYP_F32 ? align_up((size_t)n * 3 * k * 2, 256) + align_up((size_t)3 * k * n_out * 2, 256) + 256 : 256;
    return 0;
}

int flyp_project_normalize_fwd(const void* x, const void* w, int n, int k, int n_out, int dtype, void* y, int y_dtype,
                               void* y16, float* inv_norm, void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !w || !y) return fail(FLYP_ERR_ARG, "null pointer argument");
    if (n <= 0 || k <= 0 || k % 8 != 0) return fail(FLYP_ERR_ARG, "bad shape [%d, %d] (k %% 8 must be 0)", n, k);
    if (n_out <= 0 || n_out % 64 != 0 || n_out > 1024)
        return fail(FLYP_ERR_ARG, "output dim %d must be a multiple of 64, at most 1024", n_out);
    if (dtype != FLYP_BF16 && dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad dtype %d", dtype);
    if (y_dtype != FLYP_BF16 && y_dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad y_dtype %d", y_dtype);
    if (y16 != nullptr && y_dtype != FLYP_BF16) return fail(FLYP_ERR_ARG, "the fp16 copy goes with bf16 features");
    if (dtype == FLYP_F32 && k % 64 != 0) return fail(FLYP_ERR_ARG, "fp32 inputs need k %% 64 == 0 (k = %d)", k);
    size_t need = 0;
    flyp_project_normalize_workspace_bytes(n, k, n_out, dtype, &need);
    if (dtype == FLYP_F32 && (!workspace || workspace_bytes < need))
        return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, need);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUtensorMap tmX, tmW;
    int rc;
    flyp::TailParams p;
    memset(&p, 0, sizeof(p));
    p.n = n; p.n_out = n_out; p.kc = ceil_div(k, flyp::KCHUNK);
    const int split = flyp::tail_n_split(n_out);
    p.n_cta = ceil_div(ceil_div(n_out, split), 64) * 64;
    p.y = y; p.y_fp32 = y_dtype == FLYP_F32; p.y16 = y16; p.inv_norm = inv_norm;
    if (dtype == FLYP_F32) {
        uint16_t* px = static_cast<uint16_t*>(workspace);
        uint16_t* pw = reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(workspace) + align_up((size_t)n * 3 * k * 2, 256));
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(x), n, k, k, px, st);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(w), 1, k * n_out, k * n_out, pw, st);   // [3][k][n_out]
        CUDA_OK(cudaGetLastError());
        if ((rc = make_tmap(&tmX, px, n, 3 * k, 3 * k)) != 0) return rc;
        if ((rc = make_tmap(&tmW, pw, 3 * k, n_out, n_out, false, 64)) != 0) return rc;
        p.kplan = flyp::kplan_f32(k);
        p.w_plane_rows = k;
    } else {
        if ((rc = make_tmap(&tmX, x, n, k, k)) != 0) return rc;
        if ((rc = make_tmap(&tmW, w, k, n_out, n_out, false, 64)) != 0) return rc;
        p.kplan = flyp::kplan_bf16();
    }
    flyp::launch_tail(tmX, tmW, p, st);
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ l2 normalise
int flyp_l2norm_fwd(const void* x, int n, int dim, int dtype, void* y, float* inv_norm, void* stream) {
    if (!x || !y) return fail(FLYP_ERR_ARG, "null pointer argument");
    if (n <= 0 || dim <= 0 || dim % 8 != 0) return fail(FLYP_ERR_ARG, "bad shape [%d, %d] (dim %% 8 must be 0)", n, dim);
    if (dtype != FLYP_BF16 && dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad dtype %d", dtype);
    flyp::launch_l2norm_fwd(x, n, dim, dtype, y, inv_norm, static_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int flyp_l2norm_bwd(const void* y, const void* dy, const float* inv_norm, int n, int dim, int dtype, void* dx,
                    void* stream) {
    if (!y || !dy || !inv_norm || !dx) return fail(FLYP_ERR_ARG, "null pointer argument");
    if (n <= 0 || dim <= 0 || dim % 8 != 0) return fail(FLYP_ERR_ARG, "bad shape [%d, %d] (dim %% 8 must be 0)", n, dim);
    if (dtype != FLYP_BF16 && dtype != FLYP_F32) return fail(FLYP_ERR_ARG, "bad dtype %d", dtype);
    flyp::launch_l2norm_bwd(y, dy, inv_norm, n, dim, dtype, dx, static_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int flyp_debug_profile(void* device_buffer_16_u64) {
    g_prof_buf = static_cast<unsigned long long*>(device_buffer_16_u64);
    return 0;
}

int flyp_debug_kernel_events(void* fwd_start, void* fwd_stop, void* sweep_start, void* sweep_stop, int sweep) {
    if ((fwd_start == nullptr) != (fwd_stop == nullptr) || (sweep_start == nullptr) != (sweep_stop == nullptr))
        return fail(FLYP_ERR_ARG, "events come in pairs");
    g_ev_fwd[0] = static_cast<cudaEvent_t>(fwd_start); g_ev_fwd[1] = static_cast<cudaEvent_t>(fwd_stop);
    g_ev_sweep[0] = static_cast<cudaEvent_t>(sweep_start); g_ev_sweep[1] = static_cast<cudaEvent_t>(sweep_stop);
    g_ev_sweep_idx = sweep;
    return 0;
}

// ------------------------------------------------------------------------------------------------ fused argmax
int flyp_argmax(const void* a, const void* b, int n_m, int n_n, int dim, int dtype, int64_t* out_index, float* out_max,
                void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(n_m, n_n, dim, dtype);
    if (rc) return rc;
    if (!a || !b || !out_index || !workspace) return fail(FLYP_ERR_ARG, "null pointer argument");
    ClipWs w;
    carve_clip(workspace, n_m, n_n, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* one = w.dscale_part;                       // the kernel multiplies by scale * log2(e) > 0: order-preserving
    flyp::launch_fill_float(one, 1.0f, st);           // (a kernel, not a staged pageable H2D copy: graph-capturable)
    CUDA_OK(cudaGetLastError());
    CUtensorMap tmA, tmB;
    flyp::KPlan kplan = flyp::kplan_bf16();
    if (dtype == FLYP_F32) {
        const int dp = plane_cols(dim);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(a), n_m, dim, dp, w.stats.planes_a, st);
        flyp::launch_split_planes_bf16x3(static_cast<const float*>(b), n_n, dim, dp, w.stats.planes_b, st);
        CUDA_OK(cudaGetLastError());
        if ((rc = make_tmap(&tmA, w.stats.planes_a, n_m, 3 * dp, 3 * dp)) != 0) return rc;
        if ((rc = make_tmap(&tmB, w.stats.planes_b, n_n, 3 * dp, 3 * dp)) != 0) return rc;
        kplan = flyp::kplan_f32(dp);
    } else {
        if ((rc = make_tmap(&tmA, a, n_m, dim, dim)) != 0) return rc;
        if ((rc = make_tmap(&tmB, b, n_n, dim, dim)) != 0) return rc;
    }
    flyp::FwdParams p;
    memset(&p, 0, sizeof(p));
    const StatsWs& s = w.stats;
    p.n_m = n_m; p.n_n = n_n; p.kc = ceil_div(dim, flyp::KCHUNK); p.kplan = kplan;
    p.m_tiles = s.m_tiles; p.n_tiles = s.n_tiles; p.n_slots = s.n_slots; p.ld_rows = s.ld_rows; p.ld_cols = s.ld_cols;
    p.scale = one; p.shift_slack = shift_slack(n_m, n_n);
    p.rowpart = s.rowpart; p.rowmax = s.rowmax; p.colpart = nullptr; p.colmax = nullptr; p.pos = nullptr;
    p.argidx = reinterpret_cast<int*>(s.rowpart);     // the sums are not needed: their buffer carries the columns
    flyp::launch_fwd(tmA, tmB, p, /*robust=*/true, nullptr, num_sms(), st);
    CUDA_OK(cudaGetLastError());
    flyp::launch_argmax_finalize(s.rowmax, p.argidx, s.n_tiles * 2, s.ld_rows, n_m, reinterpret_cast<long long*>(out_index),
                                 out_max, st);
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ debug logits
int flyp_debug_logits(const void* a, const void* b, int n_m, int n_n, int dim, int dtype, float* out,
                      void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(n_m, n_n, dim, dtype);
    if (rc) return rc;
    if (!a || !b || !out || !workspace) return fail(FLYP_ERR_ARG, "null pointer argument");
    ClipWs w;
    carve_clip(workspace, n_m, n_n, dim, dtype, w);
    if (workspace_bytes < w.bytes) return fail(FLYP_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);
    // scale is irrelevant for raw dot products but the kernel reads it: park a 1.0f in the (unused) d(scale) scratch
    float* one = w.dscale_part;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    flyp::launch_fill_float(one, 1.0f, st);
    CUDA_OK(cudaGetLastError());
    StatsIO io;
    memset(&io, 0, sizeof(io));
    io.A = a; io.B = b; io.scale = one; io.n_m = n_m; io.n_n = n_n; io.dim = dim; io.dtype = dtype; io.dbg_logits = out;
    return run_stats(io, w.stats, st);
}

}  // extern "C"

// the ctypes mirrors in flyp_b200/_lib.py assume these layouts (tests/test_comm_cpu.py pins the Python side)
static_assert(sizeof(flyp_ready_t) == 40, "flyp_ready_t layout");
static_assert(sizeof(flyp_gathered_t) == 4 * 8 + 4 * 40 + 8, "flyp_gathered_t layout");
static_assert(sizeof(flyp_stats_t) == 3 * 8 + 40, "flyp_stats_t layout");
static_assert(sizeof(flyp_step_t) == sizeof(flyp_gathered_t) + sizeof(flyp_stats_t), "flyp_step_t layout");
