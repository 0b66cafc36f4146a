// comm.cu — peer-memory exchange between the GPUs of one node (NVLink 5 / NVSwitch), C ABI flyp_comm_*.
//
// Replaces, for the row-sharded loss, the collectives of the reference: the two feature all-gathers of
// clip/loss.py:19-69 (gather_features) and - because no rank ever holds the B x B logits here - the exchange of the
// per-column softmax statistics, of the per-row statistics and of the d(logit_scale) partial sums.
//
// Every rank owns one exchange segment that all peers map (torch symmetric memory with an NVSwitch multicast mapping
// through flyp_comm_create_external, or cudaMalloc + CUDA IPC through flyp_comm_create / connect_ipc).  The
// segment holds, double-buffered by step parity, the GATHERED matrices of this rank: text / image features in bf16 and
// their fp16 copies (the operand format of the backward's second GEMM), [world * rows, dim] rank-major - the ordering of
// clip/loss.py:66-67.  A step is
//   pack   (kernel, caller's stream)  local rows -> own slots of the four gathered matrices (+ fp16 conversion)
//   push   (copy engines, side stream) own slots -> the same slots of every rank.  With an NVSwitch multicast mapping of
//                                     the segments (torch symmetric memory, flyp_comm_create_external) ONE copy per
//                                     matrix to the multicast address reaches all ranks, followed by ONE 4-byte copy
//                                     of the sequence number to the multicast address of the flag word: 8 copy-engine
//                                     operations per step, text first.  Without multicast (CUDA IPC segments) the
//                                     same is done peer by peer in ring order (W - 1 times the operations; measured
//                                     on 8 GPUs: ~4 us + bytes / 750 GB/s per copy on one engine, 3.5 us per flag copy).
//                                     A push by remote stores from a few SMs was also measured and rejected (36 GB/s
//                                     per CTA, and it slows the concurrent tensor-core kernels by 15 %).
//   the tensor-core kernels poll those flag words (peer.cuh) right before their first TMA read of a rank's rows: the
//   forward starts on its own column block while the other blocks are still in flight.
// The small vectors (column triples, row statistics, d(scale)) are pushed by a kernel with remote stores (to the
// multicast address when there is one) and flagged the same way.  No NCCL call is on this path.
#include "../../include/flyp_clip.h"
#include "aux_kernels.cuh"
#include "comm_internal.h"
#include "peer.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace {

constexpr int MAXW = 16;
constexpr int MAXG = 8;             // flag words are spaced MAXG words (32 bytes) apart: one word per (flag set, rank)
enum { ARR_TXT = 0, ARR_TXT16 = 1, ARR_IMG = 2, ARR_IMG16 = 3, N_ARR = 4 };
enum { FLAG_STAT = N_ARR, FLAG_DS = N_ARR + 1, FLAG_RS = N_ARR + 2, N_FLAGSETS = N_ARR + 3 };

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#define COMM_CUDA_OK(expr)                                                                       \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            flyp::set_error(FLYP_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));            \
            return FLYP_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

struct SegPtrs { uint8_t* seg[MAXW]; };

}  // namespace

struct flyp_comm {
    int rank, world, dev, max_rows, dim;
    size_t seg_bytes;
    uint8_t* seg[MAXW];
    bool ipc_mapped[MAXW];
    bool connected;
    uint8_t* mc;                // multicast mapping of all segments (NVSwitch), or nullptr
    bool owns_seg;              // the own segment was cudaMalloc'ed here (else it belongs to the caller)
    cudaStream_t side;
    cudaEvent_t ev_packed, ev_pushed;
    uint32_t seq;
    uint32_t* err_host;
    uint32_t* err_dev;
    uint32_t timeout_ms;        // how long a kernel waits for a peer before it traps (0: for ever)
    int rs_min_rows;            // rows per rank from which the text gradient goes through the product + reduce-scatter
    // byte offsets inside a segment (identical on every rank)
    size_t off_feat[2][N_ARR], off_colstat[2], off_rowstat[2], off_dscale[2], off_flags, off_seqword[2], off_counter;
    size_t off_rs[2];           // reduce-scatter buffer of the text gradient: [world (source rank)][rows][dim] fp32
};

namespace {

void layout(flyp_comm* c) {
    const size_t cap = (size_t)c->world * c->max_rows;
    size_t off = 0;
    auto take = [&](size_t bytes) { off = align_up(off, 1024); const size_t r = off; off += bytes; return r; };
    for (int par = 0; par < 2; ++par) {
        // a matrix and its fp16 copy are adjacent: for fp32 features the pair holds ONE gathered fp32 matrix instead
        for (int a = 0; a < N_ARR; a += 2) {
            c->off_feat[par][a] = take(cap * c->dim * 4);
            c->off_feat[par][a + 1] = c->off_feat[par][a] + cap * c->dim * 2;
        }
        c->off_colstat[par] = take((size_t)c->world * 3 * cap * sizeof(float));
        c->off_rowstat[par] = take(2 * cap * sizeof(float));
        c->off_dscale[par] = take(MAXW * sizeof(float));
        c->off_rs[par] = take(cap * c->dim * sizeof(float));
    }
    c->off_flags = take((size_t)N_FLAGSETS * MAXW * MAXG * sizeof(uint32_t));
    c->off_seqword[0] = take(sizeof(uint32_t));
    c->off_seqword[1] = take(sizeof(uint32_t));
    c->off_counter = take(sizeof(uint32_t));
    c->seg_bytes = align_up(off, 1 << 20);
}

// flag word j of producer k in flag set `set` of rank q's segment
inline uint32_t* flag_ptr(const flyp_comm* c, int q, int set, int k) {
    return reinterpret_cast<uint32_t*>(c->seg[q] + c->off_flags) + ((size_t)set * MAXW + k) * MAXG;
}

// ---- pack: local bf16 rows -> own slots (bf16 copy + fp16 conversion), sequence word, own flags ---------------------
// One warp per row.  With `ex` the same pass also prepares the forward that follows: the positive logit of every local
// row (its positive is the LOCAL text row of the same index) and the control words of the step.
__global__ void k_pack(const uint4* __restrict__ img, const uint4* __restrict__ txt, int n_rows, int dim8,
                       uint4* __restrict__ img_bf, uint4* __restrict__ img_h, uint4* __restrict__ txt_bf,
                       uint4* __restrict__ txt_h, uint32_t* seqword, uint32_t* own_flags, int rank, uint32_t seq,
                       flyp::PackExtra ex) {
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            *seqword = seq;
            for (int a = 0; a < N_ARR; ++a) own_flags[(a * MAXW + rank) * MAXG] = seq;
        }
        if (ex.zero_words != nullptr && (int)threadIdx.x < ex.n_zero) ex.zero_words[threadIdx.x] = 0;
    }
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    auto conv = [](uint4 v) {
        uint4 u;
        const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
        uint32_t* d = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float lo = __uint_as_float(s[e] << 16), hi = __uint_as_float(s[e] & 0xffff0000u);
            uint32_t r;
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
            d[e] = r;
        }
        return u;
    };
    float acc = 0.f;
    if (row < n_rows) {
        for (int c = lane; c < dim8; c += 32) {
            const size_t i = (size_t)row * dim8 + c;
            const uint4 a = img[i], b = txt[i];
            img_bf[i] = a; txt_bf[i] = b;
            img_h[i] = conv(a); txt_h[i] = conv(b);
            const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc = fmaf(__uint_as_float(wa[e] << 16), __uint_as_float(wb[e] << 16), acc);
                acc = fmaf(__uint_as_float(wa[e] & 0xffff0000u), __uint_as_float(wb[e] & 0xffff0000u), acc);
            }
        }
    }
    if (ex.t2 != nullptr && row < ex.n_pad) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) {
            ex.t2[row] = row < n_rows ? acc * ex.scale[0] * 1.4426950408889634f : -INFINITY;
            ex.pos[row] = row < n_rows ? ex.row_offset + row : -1;
        }
    }
}

// fp32 features: the rows go out as they are (the consumers split them into bf16 / fp16 planes on arrival)
__global__ void k_pack32(const float4* __restrict__ img, const float4* __restrict__ txt, int n_rows, int dim4,
                         float4* __restrict__ img_o, float4* __restrict__ txt_o, uint32_t* seqword, uint32_t* own_flags,
                         int rank, uint32_t seq, flyp::PackExtra ex) {
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            *seqword = seq;
            for (int a = 0; a < N_ARR; ++a) own_flags[(a * MAXW + rank) * MAXG] = seq;
        }
        if (ex.zero_words != nullptr && (int)threadIdx.x < ex.n_zero) ex.zero_words[threadIdx.x] = 0;
    }
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    float acc = 0.f;
    if (row < n_rows) {
        for (int c = lane; c < dim4; c += 32) {
            const size_t i = (size_t)row * dim4 + c;
            const float4 a = img[i], b = txt[i];
            img_o[i] = a; txt_o[i] = b;
            acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
        }
    }
    if (ex.t2 != nullptr && row < ex.n_pad) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) {
            ex.t2[row] = row < n_rows ? acc * ex.scale[0] * 1.4426950408889634f : -INFINITY;
            ex.pos[row] = row < n_rows ? ex.row_offset + row : -1;
        }
    }
}

// ---- statistics push: this rank's column triples and row statistics -> every rank's segment, then the flag ----------
// ptrs.seg[0 .. n_dst): the destinations (every rank's segment, or the single multicast mapping); own_seg: this rank's
__global__ void k_push_stats(SegPtrs ptrs, int n_dst, SegPtrs uni, int world, uint8_t* own_seg, int rank, size_t off_colstat, size_t off_rowstat, size_t off_flags,
                             size_t off_counter, const float* __restrict__ col_stat, const float* __restrict__ row_lse,
                             const float* __restrict__ row_nll, int n_rows, int n_cols, size_t cap, uint32_t seq) {
    const int q = blockIdx.y;
    float* cs = reinterpret_cast<float*>(ptrs.seg[q] + off_colstat) + (size_t)rank * 3 * n_cols;
    float* rl = reinterpret_cast<float*>(ptrs.seg[q] + off_rowstat) + (size_t)rank * n_rows;
    float* rn = rl + cap;
    const int n_cs = 3 * n_cols;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if ((n_cols & 3) == 0 && (n_rows & 3) == 0 && (cap & 3) == 0) {      // 16-byte remote stores
        const float4* s4 = reinterpret_cast<const float4*>(col_stat);
        float4* d4 = reinterpret_cast<float4*>(cs);
        for (int i = tid; i < n_cs / 4; i += nth) d4[i] = s4[i];
        const float4 *l4 = reinterpret_cast<const float4*>(row_lse), *n4 = reinterpret_cast<const float4*>(row_nll);
        float4 *dl = reinterpret_cast<float4*>(rl), *dn = reinterpret_cast<float4*>(rn);
        for (int i = tid; i < n_rows / 4; i += nth) { dl[i] = l4[i]; dn[i] = n4[i]; }
    } else {
        for (int i = tid; i < n_cs; i += nth) cs[i] = col_stat[i];
        for (int i = tid; i < n_rows; i += nth) { rl[i] = row_lse[i]; rn[i] = row_nll[i]; }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        uint32_t* counter = reinterpret_cast<uint32_t*>(own_seg + off_counter);
        const uint32_t total = gridDim.x * gridDim.y;
        s_last = (atomicInc(counter, total - 1) == total - 1) ? 1 : 0;      // last block: all stores above are out
        if (s_last) __threadfence_system();
    }
    __syncthreads();
    // the flag goes out through the UNICAST mappings with release semantics (multicast stores are only weakly ordered:
    // a flag stored through the multicast address is not guaranteed to land behind the data), one thread per rank so
    // that the W round trips overlap
    if (s_last && (int)threadIdx.x < world)
        flyp::st_release_sys_u32(reinterpret_cast<uint32_t*>(uni.seg[threadIdx.x] + off_flags) + (FLAG_STAT * MAXW + rank) * MAXG, seq);
}

__global__ void k_push_scalar(flyp::PeerPush push, const float* __restrict__ value) {
    if (threadIdx.x == 0) flyp::peer_push_value(push, value[0]);
}

__global__ void k_sum_scalar(const float* parts, int world, flyp::PeerWait w, float* __restrict__ out) {
    if (threadIdx.x != 0) return;
    flyp::peer_wait_all(w);
    float s = 0.f;
    for (int q = 0; q < world; ++q) s += __ldcg(parts + q);      // fixed rank order: identical bits on every rank
    out[0] = s;
}

// ---- reduce-scatter of the text gradient (kept-dS backward, api.cu) --------------------------------------------------
// Scatter half: the product kernel dS^T . I of every rank writes its fp32 partial of rank q's rows straight into slot
// [own rank] of rank q's buffer (remote stores over NVLink, clip_dst_gemm.cu).  k_rs_signal, enqueued behind that kernel,
// releases the sequence number into every rank's flag word; k_rs_reduce waits for the W flags and sums the W slots in
// rank order (identical bits whatever the arrival order).
__global__ void k_rs_signal(SegPtrs uni, int world, int rank, size_t off_flags, uint32_t seq) {
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        flyp::st_release_sys_u32(reinterpret_cast<uint32_t*>(uni.seg[threadIdx.x] + off_flags) + (FLAG_RS * MAXW + rank) * MAXG, seq);
    }
}

__global__ void k_rs_reduce(const float* __restrict__ slots, int world, size_t slot_floats, flyp::PeerWait w, void* out,
                            int out_fp32, float mul) {
    if (threadIdx.x == 0) flyp::peer_wait_all(w);
    __syncthreads();
    const size_t n4 = slot_floats / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < world; ++q) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(slots + (size_t)q * slot_floats) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= mul; acc.y *= mul; acc.z *= mul; acc.w *= mul;
        if (out_fp32) {
            reinterpret_cast<float4*>(out)[i] = acc;
        } else {
            __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(out)[i] = u;
        }
    }
}

int check_comm(const flyp_comm* c, bool need_connected) {
    if (c == nullptr) { flyp::set_error(FLYP_ERR_ARG, "comm is null"); return FLYP_ERR_ARG; }
    if (need_connected && !c->connected) { flyp::set_error(FLYP_ERR_ARG, "comm is not connected"); return FLYP_ERR_ARG; }
    return 0;
}

// ---- feature push schedule (copy engines) ------------------------------------------------------------------------
// Matrices in the order the consumers need them: text (forward), its fp16 copy (first backward sweep), image and its
// fp16 copy (second sweep).  Each block copy is followed by the 4-byte copy of the sequence number into the flag word.
struct PushOp { void* dst; const void* src; size_t bytes; };
constexpr int MAX_PUSH_OPS = 2 * N_ARR * MAXW;
int build_push_ops(const flyp_comm* c, int par, int n_rows, int dim, bool f32, PushOp* ops) {
    int n = 0;
    const size_t esz = f32 ? 4 : 2;
    const size_t slot_bytes = (size_t)n_rows * dim * esz, slot_off = (size_t)c->rank * slot_bytes;
    uint8_t* own = c->seg[c->rank];
    const void* seqword = own + c->off_seqword[par];
    const size_t flag_stride = (size_t)MAXG * sizeof(uint32_t);
    for (int a = 0; a < N_ARR; ++a) {
        // fp32 features: arrays 1 and 3 (the fp16 copies) do not exist - only their flags are raised, together with
        // the flags of the fp32 matrix that occupies the pair
        const bool data = !f32 || (a & 1) == 0;
        const size_t off_data = c->off_feat[par][a] + slot_off;
        const size_t off_flag = c->off_flags + ((size_t)a * MAXW + c->rank) * flag_stride;
        if (c->mc != nullptr) {                                  // one multicast copy reaches every rank
            if (data) ops[n++] = {c->mc + off_data, own + off_data, slot_bytes};
            ops[n++] = {c->mc + off_flag, seqword, sizeof(uint32_t)};
        } else {
            for (int k = 1; k < c->world; ++k) {                 // ring order: one sender per receiver at a time
                const int q = (c->rank - k + c->world) % c->world;
                if (data) ops[n++] = {c->seg[q] + off_data, own + off_data, slot_bytes};
                ops[n++] = {c->seg[q] + off_flag, seqword, sizeof(uint32_t)};
            }
        }
    }
    return n;
}

}  // namespace

extern "C" {

static int comm_new(int rank, int world, int max_rows, int dim, flyp_comm** out) {
    if (!out) { flyp::set_error(FLYP_ERR_ARG, "out is null"); return FLYP_ERR_ARG; }
    if (world < 1 || world > MAXW || rank < 0 || rank >= world || max_rows <= 0 || dim <= 0 || dim % 8 != 0) {
        flyp::set_error(FLYP_ERR_ARG, "bad comm shape: rank %d world %d (max %d) rows %d dim %d", rank, world, MAXW, max_rows,
                        dim);
        return FLYP_ERR_ARG;
    }
    flyp_comm* c = new flyp_comm();
    memset(c, 0, sizeof(*c));
    c->rank = rank; c->world = world; c->max_rows = max_rows; c->dim = dim;
    COMM_CUDA_OK(cudaGetDevice(&c->dev));
    layout(c);
    COMM_CUDA_OK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    COMM_CUDA_OK(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
    COMM_CUDA_OK(cudaEventCreateWithFlags(&c->ev_pushed, cudaEventDisableTiming));
    COMM_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&c->err_host), sizeof(uint32_t), cudaHostAllocMapped));
    *c->err_host = 0;
    COMM_CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->err_dev), c->err_host, 0));
    // Ordinary rank skew (a dataloader respawn, rank 0 evaluating or writing a checkpoint) is minutes, not seconds:
    // the default is 10 minutes, FLYP_PEER_TIMEOUT_MS overrides it (0 = wait for ever), read once at creation.
    c->timeout_ms = 600000u;
    if (const char* e = getenv("FLYP_PEER_TIMEOUT_MS")) c->timeout_ms = (uint32_t)strtoul(e, nullptr, 10);
    // The product + NVLink scatter of the text gradient costs its MMAs (proportional to the rows per rank) plus a
    // scatter of (W - 1) / W x B x D x 4 bytes that does not shrink with the rank's share; the transposed sweep it replaces
    // is proportional to the rows per rank.  Measured at B = 32768, D = 512: 0.43 against 0.80 ms at 16384 rows per rank,
    // 0.28 against 0.44 ms at 8192, 0.20 (+ 0.05 ms of flag release, slot sum and waiting) against 0.26 ms at 4096.
    c->rs_min_rows = 6144;
    if (const char* e = getenv("FLYP_RS_MIN_ROWS")) c->rs_min_rows = atoi(e);
    *out = c;
    return 0;
}

int flyp_comm_layout_bytes(int world, int max_rows, int dim, size_t* bytes) {
    if (!bytes || world < 1 || world > MAXW || max_rows <= 0 || dim <= 0 || dim % 8 != 0) {
        flyp::set_error(FLYP_ERR_ARG, "bad comm shape");
        return FLYP_ERR_ARG;
    }
    flyp_comm tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.world = world; tmp.max_rows = max_rows; tmp.dim = dim;
    layout(&tmp);
    *bytes = tmp.seg_bytes;
    return 0;
}

int flyp_comm_create(int rank, int world, int max_rows, int dim, flyp_comm** out) {
    int rc = comm_new(rank, world, max_rows, dim, out);
    if (rc) return rc;
    flyp_comm* c = *out;
    uint8_t* p = nullptr;
    cudaError_t e = cudaMalloc(&p, c->seg_bytes);
    if (e != cudaSuccess) {
        flyp::set_error(FLYP_ERR_CUDA, "cudaMalloc of the %zu-byte exchange segment: %s", c->seg_bytes, cudaGetErrorString(e));
        flyp_comm_destroy(c);
        *out = nullptr;
        return FLYP_ERR_CUDA;
    }
    c->seg[rank] = p;
    c->owns_seg = true;
    COMM_CUDA_OK(cudaMemset(p, 0, c->seg_bytes));
    c->connected = (world == 1);
    COMM_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int flyp_comm_create_external(int rank, int world, int max_rows, int dim, void* const* segments, void* multicast,
                              flyp_comm** out) {
    if (!segments) { flyp::set_error(FLYP_ERR_ARG, "segments is null"); return FLYP_ERR_ARG; }
    int rc = comm_new(rank, world, max_rows, dim, out);
    if (rc) return rc;
    flyp_comm* c = *out;
    for (int q = 0; q < world; ++q) {
        if (!segments[q]) { flyp_comm_destroy(c); *out = nullptr; flyp::set_error(FLYP_ERR_ARG, "segment %d is null", q); return FLYP_ERR_ARG; }
        c->seg[q] = static_cast<uint8_t*>(segments[q]);
    }
    c->mc = static_cast<uint8_t*>(multicast);
    if (const char* e = getenv("FLYP_COMM_MULTICAST")) {       // measurement knob: 0 = ignore the multicast mapping
        if (e[0] == '0') c->mc = nullptr;
    }
    c->owns_seg = false;
    c->connected = true;
    return 0;
}

int flyp_comm_has_multicast(const flyp_comm* c) { return (c != nullptr && c->mc != nullptr) ? 1 : 0; }

int flyp_comm_segment_bytes(const flyp_comm* c, size_t* bytes) {
    int rc = check_comm(c, false);
    if (rc) return rc;
    if (bytes) *bytes = c->seg_bytes;
    return 0;
}

int flyp_comm_ipc_handle(flyp_comm* c, void* handle_out) {
    int rc = check_comm(c, false);
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == FLYP_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    COMM_CUDA_OK(cudaSetDevice(c->dev));
    COMM_CUDA_OK(cudaIpcGetMemHandle(&h, c->seg[c->rank]));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int flyp_comm_connect_ipc(flyp_comm* c, const void* all_handles) {
    int rc = check_comm(c, false);
    if (rc) return rc;
    COMM_CUDA_OK(cudaSetDevice(c->dev));
    for (int q = 0; q < c->world; ++q) {
        if (q == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t*>(all_handles) + (size_t)q * FLYP_IPC_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        COMM_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->seg[q] = static_cast<uint8_t*>(p);
        c->ipc_mapped[q] = true;
    }
    c->connected = true;
    return 0;
}

int flyp_comm_connect_local(flyp_comm* c, flyp_comm* const* peers) {
    int rc = check_comm(c, false);
    if (rc) return rc;
    for (int q = 0; q < c->world; ++q) {
        if (q == c->rank) continue;
        if (!peers[q] || peers[q]->rank != q || peers[q]->world != c->world || peers[q]->seg_bytes != c->seg_bytes) {
            flyp::set_error(FLYP_ERR_ARG, "peer %d does not match this communicator", q);
            return FLYP_ERR_ARG;
        }
        c->seg[q] = peers[q]->seg[q];
    }
    c->connected = true;
    return 0;
}

int flyp_comm_error(const flyp_comm* c) {
    if (!c) return 0;
    return (int)*reinterpret_cast<volatile uint32_t*>(c->err_host);
}

int flyp_comm_reset_error(flyp_comm* c) {
    if (!c) return 0;
    *reinterpret_cast<volatile uint32_t*>(c->err_host) = 0;
    return 0;
}

int flyp_comm_set_timeout_ms(flyp_comm* c, uint32_t timeout_ms) {
    int rc = check_comm(c, false);
    if (rc) return rc;
    c->timeout_ms = timeout_ms;
    return 0;
}

int flyp_comm_set_rs_min_rows(flyp_comm* c, int rows) {
    int rc = check_comm(c, false);
    if (rc) return rc;
    c->rs_min_rows = rows;
    return 0;
}

int flyp_comm_destroy(flyp_comm* c) {
    if (!c) return 0;
    cudaSetDevice(c->dev);
    cudaDeviceSynchronize();
    for (int q = 0; q < c->world; ++q)
        if (c->ipc_mapped[q]) cudaIpcCloseMemHandle(c->seg[q]);
    if (c->owns_seg && c->seg[c->rank]) cudaFree(c->seg[c->rank]);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ev_packed) cudaEventDestroy(c->ev_packed);
    if (c->ev_pushed) cudaEventDestroy(c->ev_pushed);
    if (c->err_host) cudaFreeHost(c->err_host);
    cudaGetLastError();
    delete c;
    return 0;
}

int flyp_comm_gather_features(flyp_comm* c, const void* img, const void* txt, int n_rows, int dim, int dtype,
                              flyp_gathered_t* out, void* stream) {
    return flyp::comm_gather(c, img, txt, n_rows, dim, dtype, nullptr, out, stream);
}

}  // extern "C"

namespace flyp {

static void fill_ready(const flyp_comm* c, flyp_ready_t* r, int set, uint32_t seq, int rows_per_flag) {
    r->flags = flag_ptr(c, c->rank, set, 0);
    r->seq = seq; r->n_flags = c->world; r->rows_per_flag = rows_per_flag; r->err = c->err_dev;
    r->sub = 1; r->stride = MAXG; r->timeout_ms = c->timeout_ms;
}

int comm_gather(flyp_comm* c, const void* img, const void* txt, int n_rows, int dim, int dtype, const PackExtra* extra,
                flyp_gathered_t* out, void* stream) {
    int rc = check_comm(c, true);
    if (rc) return rc;
    if (!img || !txt || !out) { flyp::set_error(FLYP_ERR_ARG, "null pointer argument"); return FLYP_ERR_ARG; }
    if (dtype != FLYP_BF16 && dtype != FLYP_F32) { flyp::set_error(FLYP_ERR_ARG, "bad dtype %d", dtype); return FLYP_ERR_ARG; }
    const bool f32 = dtype == FLYP_F32;
    if (n_rows <= 0 || dim != c->dim || n_rows > c->max_rows) {
        flyp::set_error(FLYP_ERR_ARG, "gather shape [%d, %d] does not fit the communicator (max rows %d, dim %d)", n_rows, dim,
                        c->max_rows, c->dim);
        return FLYP_ERR_ARG;
    }
    if (((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(txt)) & 15) != 0) {
        flyp::set_error(FLYP_ERR_ARG, "feature pointers must be 16-byte aligned");
        return FLYP_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    COMM_CUDA_OK(cudaSetDevice(c->dev));
    const uint32_t seq = ++c->seq;
    const int par = (int)(seq & 1u);
    uint8_t* own = c->seg[c->rank];
    const size_t slot_bytes = (size_t)n_rows * dim * (f32 ? 4 : 2), slot_off = (size_t)c->rank * slot_bytes;
    // the copy engines may still be reading the own slots / sequence word of the previous step
    COMM_CUDA_OK(cudaStreamWaitEvent(st, c->ev_pushed, 0));
    PackExtra ex;
    if (extra != nullptr) ex = *extra; else memset(&ex, 0, sizeof(ex));
    const int rows = (ex.t2 != nullptr && ex.n_pad > n_rows) ? ex.n_pad : n_rows;
    if (f32)
        k_pack32<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(
            static_cast<const float4*>(img), static_cast<const float4*>(txt), n_rows, dim / 4,
            reinterpret_cast<float4*>(own + c->off_feat[par][ARR_IMG] + slot_off),
            reinterpret_cast<float4*>(own + c->off_feat[par][ARR_TXT] + slot_off),
            reinterpret_cast<uint32_t*>(own + c->off_seqword[par]), reinterpret_cast<uint32_t*>(own + c->off_flags), c->rank,
            seq, ex);
    else
        k_pack<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(
            static_cast<const uint4*>(img), static_cast<const uint4*>(txt), n_rows, dim / 8,
            reinterpret_cast<uint4*>(own + c->off_feat[par][ARR_IMG] + slot_off),
            reinterpret_cast<uint4*>(own + c->off_feat[par][ARR_IMG16] + slot_off),
            reinterpret_cast<uint4*>(own + c->off_feat[par][ARR_TXT] + slot_off),
            reinterpret_cast<uint4*>(own + c->off_feat[par][ARR_TXT16] + slot_off),
            reinterpret_cast<uint32_t*>(own + c->off_seqword[par]), reinterpret_cast<uint32_t*>(own + c->off_flags), c->rank,
            seq, ex);
    COMM_CUDA_OK(cudaGetLastError());
    if (c->world > 1) {
        COMM_CUDA_OK(cudaEventRecord(c->ev_packed, st));
        COMM_CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_packed, 0));
        PushOp ops[MAX_PUSH_OPS];
        const int n = build_push_ops(c, par, n_rows, dim, f32, ops);
        for (int i = 0; i < n; ++i)
            COMM_CUDA_OK(cudaMemcpyAsync(ops[i].dst, ops[i].src, ops[i].bytes, cudaMemcpyDefault, c->side));
        COMM_CUDA_OK(cudaEventRecord(c->ev_pushed, c->side));
    }
    memset(out, 0, sizeof(*out));
    out->txt_all = own + c->off_feat[par][ARR_TXT];
    out->txt16_all = f32 ? nullptr : own + c->off_feat[par][ARR_TXT16];
    out->img_all = own + c->off_feat[par][ARR_IMG];
    out->img16_all = f32 ? nullptr : own + c->off_feat[par][ARR_IMG16];
    flyp_ready_t* r[N_ARR] = {&out->txt_ready, &out->txt16_ready, &out->img_ready, &out->img16_ready};
    for (int a = 0; a < N_ARR; ++a) fill_ready(c, r[a], a, seq, n_rows);
    out->seq = seq;
    return 0;
}

int comm_scalar_push_target(flyp_comm* c, uint32_t seq, PeerPush* out) {
    int rc = check_comm(c, true);
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    const size_t off = c->off_dscale[seq & 1u] + (size_t)c->rank * sizeof(float);
    if (c->mc != nullptr) {
        out->n_dst = 1;
        out->dst[0] = reinterpret_cast<float*>(c->mc + off);
    } else {
        out->n_dst = c->world;
        for (int q = 0; q < c->world; ++q) out->dst[q] = reinterpret_cast<float*>(c->seg[q] + off);
    }
    out->n_flag = c->world;
    for (int q = 0; q < c->world; ++q) out->flag[q] = flag_ptr(c, q, FLAG_DS, c->rank);
    out->seq = seq;
    return 0;
}

// Reduce-scatter buffer of step `seq`: out_rank[q] = this rank's slot ([n_rows][dim] fp32) in rank q's segment.
int comm_rs_targets(flyp_comm* c, uint32_t seq, int n_rows, int dim, float** out_rank) {
    int rc = check_comm(c, true);
    if (rc) return rc;
    if (n_rows <= 0 || n_rows > c->max_rows || dim != c->dim) {
        flyp::set_error(FLYP_ERR_ARG, "reduce-scatter block [%d, %d] does not fit the communicator", n_rows, dim);
        return FLYP_ERR_ARG;
    }
    const size_t off = c->off_rs[seq & 1u] + (size_t)c->rank * n_rows * dim * sizeof(float);
    for (int q = 0; q < c->world; ++q) out_rank[q] = reinterpret_cast<float*>(c->seg[q] + off);
    return 0;
}

int comm_world(const flyp_comm* c) { return c != nullptr ? c->world : 1; }
int comm_rs_min_rows(const flyp_comm* c) { return c != nullptr ? c->rs_min_rows : 0; }

int comm_rs_signal(flyp_comm* c, uint32_t seq, void* stream) {
    int rc = check_comm(c, true);
    if (rc) return rc;
    SegPtrs uni;
    for (int q = 0; q < MAXW; ++q) uni.seg[q] = q < c->world ? c->seg[q] : nullptr;
    k_rs_signal<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(uni, c->world, c->rank, c->off_flags, seq);
    COMM_CUDA_OK(cudaGetLastError());
    return 0;
}

int comm_rs_reduce(flyp_comm* c, uint32_t seq, int n_rows, int dim, void* out, int out_fp32, float mul, void* stream) {
    int rc = check_comm(c, true);
    if (rc) return rc;
    flyp::PeerWait w;
    w.flags = flag_ptr(c, c->rank, FLAG_RS, 0); w.seq = seq; w.n_flags = c->world; w.rows_per_flag = 1; w.err = c->err_dev;
    w.sub = 1; w.stride = MAXG; w.timeout_ms = c->timeout_ms;
    const size_t slot = (size_t)n_rows * dim;
    int blocks = (int)((slot / 4 + 255) / 256);
    if (blocks > 592) blocks = 592;                 // 4 per SM: an HBM-bound sum of W x n_rows x dim floats
    if (blocks < 1) blocks = 1;
    k_rs_reduce<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float*>(c->seg[c->rank] + c->off_rs[seq & 1u]), c->world, slot, w, out, out_fp32, mul);
    COMM_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace flyp

extern "C" {

int flyp_comm_push_stats(flyp_comm* c, uint32_t seq, const float* col_stat, const float* row_lse, const float* row_nll,
                         int n_rows, int n_cols, flyp_stats_t* out, void* stream) {
    int rc = check_comm(c, true);
    if (rc) return rc;
    if (!col_stat || !row_lse || !row_nll || !out) { flyp::set_error(FLYP_ERR_ARG, "null pointer argument"); return FLYP_ERR_ARG; }
    const size_t cap = (size_t)c->world * c->max_rows;
    if (n_rows <= 0 || n_rows > c->max_rows || n_cols != n_rows * c->world) {
        flyp::set_error(FLYP_ERR_ARG, "statistics shape (%d rows, %d columns) does not fit the communicator", n_rows, n_cols);
        return FLYP_ERR_ARG;
    }
    COMM_CUDA_OK(cudaSetDevice(c->dev));
    const int par = (int)(seq & 1u);
    SegPtrs ptrs, uni;
    int n_dst = c->world;
    for (int q = 0; q < MAXW; ++q) uni.seg[q] = ptrs.seg[q] = q < c->world ? c->seg[q] : nullptr;
    if (c->mc != nullptr) { n_dst = 1; ptrs.seg[0] = c->mc; }      // one store stream reaches every rank
    int bx = (3 * n_cols + 1023) / 1024;
    if (bx > 16) bx = 16;
    if (bx < 1) bx = 1;
    k_push_stats<<<dim3(bx, n_dst), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        ptrs, n_dst, uni, c->world, c->seg[c->rank], c->rank, c->off_colstat[par], c->off_rowstat[par], c->off_flags, c->off_counter, col_stat, row_lse,
        row_nll, n_rows, n_cols, cap, seq);
    COMM_CUDA_OK(cudaGetLastError());
    uint8_t* own = c->seg[c->rank];
    out->col_stat_all = reinterpret_cast<const float*>(own + c->off_colstat[par]);
    out->row_lse_all = reinterpret_cast<const float*>(own + c->off_rowstat[par]);
    out->row_nll_all = out->row_lse_all + cap;
    flyp::fill_ready(c, &out->ready, FLAG_STAT, seq, n_rows);
    return 0;
}

int flyp_comm_push_scalar(flyp_comm* c, uint32_t seq, const float* value, void* stream) {
    int rc = check_comm(c, true);
    if (rc) return rc;
    if (!value) { flyp::set_error(FLYP_ERR_ARG, "null pointer argument"); return FLYP_ERR_ARG; }
    COMM_CUDA_OK(cudaSetDevice(c->dev));
    flyp::PeerPush push;
    if ((rc = flyp::comm_scalar_push_target(c, seq, &push)) != 0) return rc;
    k_push_scalar<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(push, value);
    COMM_CUDA_OK(cudaGetLastError());
    return 0;
}

int flyp_comm_sum_scalar(flyp_comm* c, uint32_t seq, float* out, void* stream) {
    int rc = check_comm(c, true);
    if (rc) return rc;
    if (!out) { flyp::set_error(FLYP_ERR_ARG, "null pointer argument"); return FLYP_ERR_ARG; }
    COMM_CUDA_OK(cudaSetDevice(c->dev));
    flyp::PeerWait w;
    w.flags = flag_ptr(c, c->rank, FLAG_DS, 0); w.seq = seq; w.n_flags = c->world; w.rows_per_flag = 1; w.err = c->err_dev;
    w.sub = 1; w.stride = MAXG; w.timeout_ms = c->timeout_ms;
    k_sum_scalar<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float*>(c->seg[c->rank] + c->off_dscale[seq & 1u]), c->world, w, out);
    COMM_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
