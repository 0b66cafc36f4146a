// aux_kernels.cu — the HBM-bound helper kernels of the ClipLoss path (see aux_kernels.cuh).
#include "aux_kernels.cuh"
#include "clip_kernels.cuh"
#include "sched.h"
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace flyp {

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// 8 consecutive elements -> fp32
template <bool F32>
__device__ __forceinline__ void load8(const void* base, size_t elem, float (&v)[8]) {
    if (F32) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem);
        float4 a = __ldg(p), b = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[2 * e] = __uint_as_float(w[e] << 16);
            v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
        }
    }
}
template <bool F32>
__device__ __forceinline__ void store8(void* base, size_t elem, const float (&v)[8]) {
    if (F32) {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem);
        p[0] = make_float4(v[0], v[1], v[2], v[3]);
        p[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint4 u;
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(v[4], v[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(v[6], v[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + elem) = u;
    }
}

// ------------------------------------------------------------------------------------------------ pair logits
__device__ __forceinline__ uint4 bf16x8_to_f16x8(uint4 v) {
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
}

// The forward's preparation pass, one warp per row index r:
//   r < n_pad: t2[r] = scale * log2(e) * <A[r,:], B[idx(r),:]> (-inf when row r has no valid positive), pos[r] = idx(r)
//              or -1; rows [n, n_pad) are padding (pos = -1, t2 = -inf);
//   bf16 features, a16 / b16 given: the fp16 copies of A (rows < n) and B (rows < n_b) that the backward's second GEMM
//              multiplies (tcgen05 kind::f16 needs both operands in one 16-bit format) - written here so that the
//              features cross HBM once for both purposes;
//   zero_words (n_zero ints): control words of the step (fast-path flag) cleared by the first thread.
template <bool F32>
__global__ void k_pair_dot(const void* __restrict__ A, const void* __restrict__ B, const float* __restrict__ scale,
                           int n, int n_pad, int n_b, int dim, const int64_t* __restrict__ labels, int offset,
                           float* __restrict__ t2, int* __restrict__ pos, const int* __restrict__ gate,
                           int* __restrict__ zero_words, int n_zero, uint4* __restrict__ a16, uint4* __restrict__ b16) {
    if (gate != nullptr && *gate == 0) return;      // helper of the robust path: nothing to do
    if (zero_words != nullptr && blockIdx.x == 0 && (int)threadIdx.x < n_zero) zero_words[threadIdx.x] = 0;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (!F32 && b16 != nullptr && row < n_b) {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(B) + (size_t)row * dim);
        uint4* dst = b16 + (size_t)row * (dim / 8);
        for (int c = lane; c < dim / 8; c += 32) dst[c] = bf16x8_to_f16x8(__ldg(src + c));
    }
    if (row >= n_pad) return;
    long long idx = -1;
    bool ignored = false;
    if (row < n) {
        idx = labels ? labels[row] : (long long)offset + row;
        if (labels != nullptr) {
            // F.cross_entropy semantics (src/models/ce_ablation.py:123): ignore_index = -100 gives loss 0 / gradient 0,
            // any other target outside [0, n_classes) is a device-side assertion failure
            if (idx == -100) ignored = true;
            else if (idx < 0 || idx >= n_b) __trap();
        }
    }
    const bool ok = idx >= 0 && idx < n_b;
    float acc = 0.f;
    if (!F32 && a16 != nullptr && row < n) {
        // bf16 rows: one pass serves the conversion and the dot product
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(A) + (size_t)row * dim);
        uint4* dst = a16 + (size_t)row * (dim / 8);
        for (int c = lane; c < dim / 8; c += 32) {
            const uint4 v = __ldg(src + c);
            dst[c] = bf16x8_to_f16x8(v);
            if (ok) {
                float b[8];
                load8<F32>(B, (size_t)idx * dim + c * 8, b);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc = fmaf(__uint_as_float(w[e] << 16), b[2 * e], acc);
                    acc = fmaf(__uint_as_float(w[e] & 0xffff0000u), b[2 * e + 1], acc);
                }
            }
        }
    } else if (ok) {
        for (int d = lane * 8; d < dim; d += 256) {
            float a[8], b[8];
            load8<F32>(A, (size_t)row * dim + d, a);
            load8<F32>(B, (size_t)idx * dim + d, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(a[e], b[e], acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        t2[row] = ok ? acc * scale[0] * LOG2E_F : -INFINITY;
        pos[row] = ok ? (int)idx : (ignored ? -2 : -1);
    }
}

void launch_pair_dot(const void* A, const void* B, int dtype, const float* scale, int n, int n_pad, int n_b, int dim,
                     const int64_t* labels, int offset, float* t2, int* pos, const int* gate, int* zero_words,
                     int n_zero, void* a16, void* b16, cudaStream_t st) {
    int rows = n_pad;
    if (dtype != 1 && b16 != nullptr && n_b > rows) rows = n_b;
    if (rows <= 0) return;
    const int wpb = 8;
    dim3 grid((rows + wpb - 1) / wpb), block(wpb * 32);
    if (dtype == 1)
        k_pair_dot<true><<<grid, block, 0, st>>>(A, B, scale, n, n_pad, n_b, dim, labels, offset, t2, pos, gate, zero_words,
                                                n_zero, nullptr, nullptr);
    else
        k_pair_dot<false><<<grid, block, 0, st>>>(A, B, scale, n, n_pad, n_b, dim, labels, offset, t2, pos, gate,
                                                 zero_words, n_zero, static_cast<uint4*>(a16), static_cast<uint4*>(b16));
}

// log2-domain logaddexp of a (off-positive mass) and t (positive logit): returns lse2 and nll = ln2 * (lse2 - t)
__device__ __forceinline__ void lse_with_positive(float a, float t, float& lse2, float& nll) {
    if (t == -INFINITY) { lse2 = a; nll = a * LN2_F; return; }      // no positive: nll degenerates to lse
    if (a == -INFINITY) { lse2 = t; nll = 0.f; return; }
    if (t >= a) {
        const float r = exp2f(a - t);
        lse2 = t + log2f(1.f + r);
        nll = log1pf(r);
    } else {
        const float r = exp2f(t - a);
        lse2 = a + log2f(1.f + r);
        nll = (a - t) * LN2_F + log1pf(r);
    }
}

// ------------------------------------------------------------------------------------------------ finalize (fast)
// col_stat layout: [3][n_n] = (m_j log2 reference, sum_j of exp2(x - m_j) over the local rows EXCLUDING positives,
//                              t_j log2 positive logit of column j if its row is local else -inf)
// Block = 8 warps x 32 lanes: lane = row / column within a group of 32, the warps split the partials 8 ways
// (coalesced 128-byte reads, 8x more loads in flight than one thread per row), then one smem reduction.
// fin (single-rank symmetric loss, n_m == n_n, positives on the diagonal): the column triples need no merge across
// ranks, so col_lse / col_nll and loss[i] = (row_nll[i] + col_nll[i]) / 2 (clip/loss.py:208-209) are finished here.
__device__ __forceinline__ void finish_item(const FwdFinish& fin, int i, float col_a, float col_t, float row_nll_i) {
    float lse2, nll;
    lse_with_positive(col_a, col_t, lse2, nll);
    fin.col_lse[i] = lse2 * LN2_F;
    fin.col_nll[i] = nll;
    const float v = 0.5f * (row_nll_i + nll);
    if (fin.loss_bf16) reinterpret_cast<__nv_bfloat16*>(fin.loss)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(fin.loss)[i] = v;
}

__global__ void k_fwd_finalize(const float* __restrict__ rowpart, int n_rowparts, int ld_rows, int n_m,
                               const float* __restrict__ colpart, FwdColSched cs, int ld_cols, int n_n,
                               const float* __restrict__ scale, float slack, const float* __restrict__ t2,
                               const int* __restrict__ pos, int col_pos_offset, float* __restrict__ row_lse,
                               float* __restrict__ row_nll, float* __restrict__ col_stat, int* __restrict__ flag,
                               FwdFinish fin) {
    __shared__ float sm_r[8][32], sm_c[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const float c1 = scale[0] * LOG2E_F;
    const float c0 = fixed_shift(c1, slack);
    // A sum is trustworthy when the mass lost to flush-to-zero (< n * 2^-126 in shifted units) is below fp32
    // resolution of either the sum itself or of the (exactly known) positive term it is added to.
    const float lg_r = log2f((float)n_n), lg_c = log2f((float)n_m);
    float pr = 0.f, pc = 0.f;
    if (i < n_m)
        for (int p = w; p < n_rowparts; p += 8) pr += rowpart[(size_t)p * ld_rows + i];
    if (i < n_n && col_stat != nullptr) {
        // the slots the flat schedule wrote for this column's block (sched.h)
        const int nb = i >> 7, u = cs.mc ? nb >> 1 : nb;
        int ub = u - cs.rot;
        if (ub < 0) ub += cs.n_units;
        const int n_colparts = fwd_unit_slots(ub, cs.m_tiles, cs.n_units, cs.n_local, cs.workers);
        for (int p = w; p < n_colparts; p += 8) pc += colpart[(size_t)p * ld_cols + i];
    }
    sm_r[w][lane] = pr; sm_c[w][lane] = pc;
    __syncthreads();
    if (w != 0) return;
    bool bad = false;
    float nll_row = 0.f;
    if (i < n_m) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += sm_r[k][lane];
        const float t = t2 ? t2[i] : -INFINITY;
        const float a = sum > 0.f ? log2f(sum) + c0 : -INFINITY;
        float lse2;
        lse_with_positive(a, t, lse2, nll_row);
        if (pos != nullptr && pos[i] == -2) nll_row = 0.f;       // ignored target (F.cross_entropy ignore_index)
        row_lse[i] = lse2 * LN2_F;
        if (row_nll) row_nll[i] = nll_row;
        const bool ok = !isinf(sum) && (sum >= exp2f(lg_r - 101.f) || t >= c0 - 101.f + lg_r);
        bad |= !ok;
    }
    if (i < n_n && col_stat != nullptr) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += sm_c[k][lane];
        const int r = i - col_pos_offset;                    // local row whose positive is column i
        const float t = (r >= 0 && r < n_m && t2 && pos[r] == i) ? t2[r] : -INFINITY;
        col_stat[i] = c0;
        col_stat[n_n + i] = sum;
        col_stat[2 * n_n + i] = t;
        const bool ok = !isinf(sum) && (sum >= exp2f(lg_c - 101.f) || t >= c0 - 101.f + lg_c);
        bad |= !ok;
        if (fin.col_lse != nullptr && i < n_m) finish_item(fin, i, sum > 0.f ? log2f(sum) + c0 : -INFINITY, t, nll_row);
    }
    if (bad) atomicOr(flag, 1);
}

void launch_fwd_finalize(const float* rowpart, int n_rowparts, int ld_rows, int n_m, const float* colpart,
                         const FwdColSched& cs, int ld_cols, int n_n, const float* scale, float slack, const float* t2,
                         const int* pos, int col_pos_offset, float* row_lse, float* row_nll, float* col_stat,
                         int* flag, const FwdFinish& fin, cudaStream_t st) {
    const int n = n_m > n_n ? n_m : n_n;
    k_fwd_finalize<<<(n + 31) / 32, 256, 0, st>>>(rowpart, n_rowparts, ld_rows, n_m, colpart, cs, ld_cols, n_n, scale,
                                                    slack, t2, pos, col_pos_offset, row_lse, row_nll, col_stat, flag, fin);
}

// ------------------------------------------------------------------------------------------------ finalize (robust)
__device__ __forceinline__ void merge_pairs(const float* __restrict__ part, const float* __restrict__ pmax, int np,
                                            int ld, int i, float& m_out, float& s_out) {
    float m = -INFINITY;
    for (int p = 0; p < np; ++p) m = fmaxf(m, pmax[(size_t)p * ld + i]);
    float s = 0.f;
    if (m > -INFINITY) {
        for (int p = 0; p < np; ++p) {
            const float pm = pmax[(size_t)p * ld + i];
            if (pm > -INFINITY) s += part[(size_t)p * ld + i] * exp2f(pm - m);
        }
    }
    m_out = m;
    s_out = s;
}

__global__ void k_fwd_finalize_robust(const float* __restrict__ rowpart, const float* __restrict__ rowmax,
                                      int n_rowparts, int ld_rows, int n_m, const float* __restrict__ colpart,
                                      const float* __restrict__ colmax, int n_colparts, int ld_cols, int n_n,
                                      const float* __restrict__ t2, const int* __restrict__ pos,
                                      float* __restrict__ row_lse, float* __restrict__ row_nll,
                                      float* __restrict__ col_stat, const int* __restrict__ flag,
                                      int* __restrict__ status, FwdFinish fin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && status != nullptr) status[0] = *flag;
    if (*flag == 0) return;
    float nll_row = 0.f;
    if (i < n_m) {
        float m, s;
        merge_pairs(rowpart, rowmax, n_rowparts, ld_rows, i, m, s);
        const float a = s > 0.f ? log2f(s) + m : -INFINITY;
        float lse2;
        lse_with_positive(a, t2 ? t2[i] : -INFINITY, lse2, nll_row);
        if (pos != nullptr && pos[i] == -2) nll_row = 0.f;
        row_lse[i] = lse2 * LN2_F;
        if (row_nll) row_nll[i] = nll_row;
    }
    if (i < n_n && col_stat != nullptr) {
        float m, s;
        merge_pairs(colpart, colmax, n_colparts, ld_cols, i, m, s);
        col_stat[i] = m;          // the positive-logit row (col_stat[2 n_n + i]) was already written by the fast finalize
        col_stat[n_n + i] = s;
        if (fin.col_lse != nullptr && i < n_m)
            finish_item(fin, i, s > 0.f ? log2f(s) + m : -INFINITY, col_stat[2 * n_n + i], nll_row);
    }
}

void launch_fwd_finalize_robust(const float* rowpart, const float* rowmax, int n_rowparts, int ld_rows, int n_m,
                                const float* colpart, const float* colmax, int n_colparts, int ld_cols, int n_n,
                                const float* t2, const int* pos, float* row_lse, float* row_nll, float* col_stat,
                                const int* flag, int* status, const FwdFinish& fin, cudaStream_t st) {
    const int n = n_m > n_n ? n_m : n_n;
    k_fwd_finalize_robust<<<(n + 255) / 256, 256, 0, st>>>(rowpart, rowmax, n_rowparts, ld_rows, n_m, colpart, colmax,
                                                           n_colparts, ld_cols, n_n, t2, pos, row_lse, row_nll,
                                                           col_stat, flag, status, fin);
}

// ------------------------------------------------------------------------------------------------ argmax finalize
// out[i] = column of the maximum of row i over all (column block, half) partials; partials are in increasing column
// order and a later one wins only when strictly greater, so ties go to the lowest index like torch.argmax.
__global__ void k_argmax_finalize(const float* __restrict__ pmax, const int* __restrict__ pidx, int n_parts, int ld,
                                  int n_m, long long* __restrict__ out, float* __restrict__ out_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_m) return;
    float best = pmax[i];
    int bi = pidx[i];
    for (int p = 1; p < n_parts; ++p) {
        const float v = pmax[(size_t)p * ld + i];
        if (v > best) { best = v; bi = pidx[(size_t)p * ld + i]; }
    }
    out[i] = bi;
    if (out_max != nullptr) out_max[i] = best;
}
void launch_argmax_finalize(const float* pmax, const int* pidx, int n_parts, int ld, int n_m, long long* out,
                            float* out_max, cudaStream_t st) {
    k_argmax_finalize<<<(n_m + 255) / 256, 256, 0, st>>>(pmax, pidx, n_parts, ld, n_m, out, out_max);
}

// ------------------------------------------------------------------------------------------------ clip finish
// col_stat_all[world][3 * n_cols] -> col_lse[n_cols]; loss[i] = 0.5 * (row_nll[i] + col_nll[row_offset + i])
__global__ void k_clip_finish(const float* col_stat_all, int world, const float* row_nll, int n_rows, int n_cols,
                              int row_offset, float* __restrict__ col_lse, float* __restrict__ col_nll,
                              void* __restrict__ loss, int loss_bf16, PeerWait wait) {
    // multi-GPU: the triples / row statistics of the other ranks are pushed into this rank's memory over NVLink; poll
    // their flags first, then read through L2 (__ldcg: no stale non-coherent lines)
    if (wait.flags != nullptr) {
        if (threadIdx.x == 0) peer_wait_all(wait);
        __syncthreads();
    }
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    const size_t ldw = (size_t)3 * n_cols;
    float m = -INFINITY, t = -INFINITY;
    for (int w = 0; w < world; ++w) {
        if (__ldcg(col_stat_all + w * ldw + n_cols + j) > 0.f) m = fmaxf(m, __ldcg(col_stat_all + w * ldw + j));
        t = fmaxf(t, __ldcg(col_stat_all + w * ldw + 2 * n_cols + j));
    }
    float s = 0.f;
    for (int w = 0; w < world; ++w) {
        const float sw = __ldcg(col_stat_all + w * ldw + n_cols + j);
        if (sw > 0.f) s += sw * exp2f(__ldcg(col_stat_all + w * ldw + j) - m);
    }
    const float a = s > 0.f ? log2f(s) + m : -INFINITY;
    float lse2, nll;
    lse_with_positive(a, t, lse2, nll);
    col_lse[j] = lse2 * LN2_F;
    col_nll[j] = nll;
    const int i = j - row_offset;
    if (i >= 0 && i < n_rows) {
        const float v = 0.5f * (__ldcg(row_nll + i) + nll);
        if (loss_bf16) reinterpret_cast<__nv_bfloat16*>(loss)[i] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(loss)[i] = v;
    }
}

void launch_clip_finish(const float* col_stat_all, int world, const float* row_nll, int n_rows, int n_cols,
                        int row_offset, float* col_lse, float* col_nll, void* loss, int loss_bf16, PeerWait wait,
                        cudaStream_t st) {
    k_clip_finish<<<(n_cols + 255) / 256, 256, 0, st>>>(col_stat_all, world, row_nll, n_rows, n_cols, row_offset,
                                                        col_lse, col_nll, loss, loss_bf16, wait);
}

// ------------------------------------------------------------------------------------------------ bwd vectors
__global__ void k_bwd_prep(int n, int n_pad, const float* __restrict__ g, float wmul, const float* __restrict__ lse,
                           const float* __restrict__ nll, const int64_t* __restrict__ labels, int lab_offset,
                           int lab_range, const float* __restrict__ g2, const float* __restrict__ nll2, float dmul, float* __restrict__ w, float* __restrict__ l2,
                           int* __restrict__ lab, float* __restrict__ d, uint32_t* __restrict__ gmax_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float wi = 0.f, li = 0.f, di = 0.f;
    int lb = -1;
    // words (all zero before the first prep kernel of a backward): [0] bits of max|g|, [1] order-preserving key of the
    // max of lse (log2 units) over live entries, [2] COMPLEMENT of the key of their min (so that zero is neutral)
    uint32_t gb = 0u, khi = 0u, klo = 0xffffffffu;
    if (i < n) {
        gb = __float_as_uint(fabsf(g[i]));
        if (g[i] != 0.f) {
            const uint32_t u = __float_as_uint(lse[i] * LOG2E_F);
            khi = klo = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        }
    }
    gb = __reduce_max_sync(0xffffffffu, gb);
    khi = __reduce_max_sync(0xffffffffu, khi);
    klo = __reduce_min_sync(0xffffffffu, klo);
    if ((threadIdx.x & 31) == 0 && gmax_bits) {
        if (gb != 0u) atomicMax(gmax_bits, gb);
        if (khi != 0u) atomicMax(gmax_bits + 1, khi);
        if (klo != 0xffffffffu) atomicMax(gmax_bits + 2, ~klo);
    }
    if (i >= n_pad) return;
    if (i < n) {
        const float gi = g[i];
        wi = wmul * gi;
        li = lse[i] * LOG2E_F;
        long long t = labels ? labels[i] : (long long)i + lab_offset;
        if (labels != nullptr && t == -100) wi = 0.f;            // ignored target: no gradient from this row
        if (t >= 0 && t < lab_range) {
            lb = (int)t;
            // softmax - 1 at the positive = expm1(-nll), free of cancellation
            di = dmul * (gi * expm1f(-nll[i]) + (g2 ? g2[lb] * expm1f(-nll2[lb]) : 0.f));
        }
    }
    if (w) w[i] = wi;
    if (l2) l2[i] = li;
    if (lab) lab[i] = lb;
    if (d) d[i] = di;
}

void launch_bwd_prep(int n, int n_pad, const float* g, float wmul, const float* lse, const float* nll,
                     const int64_t* labels, int lab_offset, int lab_range, const float* g2, const float* nll2,
                     float dmul, float* w, float* l2, int* lab, float* d, uint32_t* gmax_bits, cudaStream_t st) {
    k_bwd_prep<<<(n_pad + 255) / 256, 256, 0, st>>>(n, n_pad, g, wmul, lse, nll, labels, lab_offset, lab_range, g2,
                                                    nll2, dmul, w, l2, lab, d, gmax_bits);
}

// Row-sharded symmetric loss: one pass prepares the vectors of BOTH backward sweeps of a rank.  Global (padded) vectors:
// w = g / 2, l2c = col_lse log2 e, l2r = row_lse_all log2 e; local: lab[i] = off + i, d[i] = exact dS at the positive of
// local row i (the same value serves the transposed sweep).
__global__ void k_bwd_prep_sharded(int n, int n_pad, int off, int n_loc, const void* __restrict__ g, int g_bf16,
                                   const float* row_lse_all, const float* row_nll_all, const float* __restrict__ col_lse,
                                   const float* __restrict__ col_nll, float* __restrict__ w, float* __restrict__ l2c,
                                   float* __restrict__ l2r, int* __restrict__ lab, float* __restrict__ d,
                                   uint32_t* __restrict__ words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t gb = 0u, khi = 0u, klo = 0xffffffffu;
    float wi = 0.f, lc = 0.f, lr = 0.f;
    if (i < n) {
        const float gi = g_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g)[i])
                                : reinterpret_cast<const float*>(g)[i];
        gb = __float_as_uint(fabsf(gi));
        wi = 0.5f * gi;
        lc = col_lse[i] * LOG2E_F;
        lr = __ldcg(row_lse_all + i) * LOG2E_F;
        if (gi != 0.f) {
            const uint32_t u = __float_as_uint(lc), v = __float_as_uint(lr);
            const uint32_t ku = (u & 0x80000000u) ? ~u : (u | 0x80000000u), kv = (v & 0x80000000u) ? ~v : (v | 0x80000000u);
            khi = ku > kv ? ku : kv;
            klo = ku < kv ? ku : kv;
        }
        const int r = i - off;
        if (r >= 0 && r < n_loc) {
            lab[r] = i;
            d[r] = 0.5f * gi * (expm1f(-__ldcg(row_nll_all + i)) + expm1f(-col_nll[i]));
        }
    }
    gb = __reduce_max_sync(0xffffffffu, gb);
    khi = __reduce_max_sync(0xffffffffu, khi);
    klo = __reduce_min_sync(0xffffffffu, klo);
    if ((threadIdx.x & 31) == 0) {
        if (gb != 0u) atomicMax(words, gb);
        if (khi != 0u) atomicMax(words + 1, khi);
        if (klo != 0xffffffffu) atomicMax(words + 2, ~klo);
    }
    if (i < n_pad) { w[i] = wi; l2c[i] = lc; l2r[i] = lr; }
}
void launch_bwd_prep_sharded(int n, int n_pad, int off, int n_loc, const void* g, int g_bf16, const float* row_lse_all,
                             const float* row_nll_all, const float* col_lse, const float* col_nll, float* w, float* l2c,
                             float* l2r, int* lab, float* d, uint32_t* words, cudaStream_t st) {
    k_bwd_prep_sharded<<<(n_pad + 255) / 256, 256, 0, st>>>(n, n_pad, off, n_loc, g, g_bf16, row_lse_all, row_nll_all, col_lse,
                                                            col_nll, w, l2c, l2r, lab, d, words);
}

// Fast (single-exponential) form of the backward epilogue: c0 = centre of the lse range, valid when the range is at
// most 200 log2 units wide; f[i] = w[i] * 2^(c0 - l2[i]).
__device__ __forceinline__ float key_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void k_bwd_fast_vectors(const uint32_t* __restrict__ words, int n_a, const float* __restrict__ w_a,
                                   const float* __restrict__ l_a, float* __restrict__ f_a, int n_b,
                                   const float* __restrict__ w_b, const float* __restrict__ l_b,
                                   float* __restrict__ f_b, float* __restrict__ info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t khi = words[1], klo = ~words[2];
    float c0 = 0.f, valid = 1.f;
    if (khi != 0u && klo != 0xffffffffu) {
        const float hi = key_to_float(khi), lo = key_to_float(klo);
        c0 = 0.5f * (hi + lo);
        valid = (hi - lo <= 200.f && isfinite(hi) && isfinite(lo)) ? 1.f : 0.f;
    }
    if (i == 0) { info[0] = c0; info[1] = valid; }
    if (i < n_a) { const float w = w_a[i]; f_a[i] = (w == 0.f) ? 0.f : w * exp2f(c0 - l_a[i]); }
    if (f_b != nullptr && i < n_b) { const float w = w_b[i]; f_b[i] = (w == 0.f) ? 0.f : w * exp2f(c0 - l_b[i]); }
}
void launch_bwd_fast_vectors(const uint32_t* words, int n_a, const float* w_a, const float* l_a, float* f_a, int n_b,
                             const float* w_b, const float* l_b, float* f_b, float* info, cudaStream_t st) {
    const int n = n_a > n_b ? n_a : n_b;
    k_bwd_fast_vectors<<<(n + 255) / 256, 256, 0, st>>>(words, n_a, w_a, l_a, f_a, n_b, w_b, l_b, f_b, info);
}

__global__ void k_sum_parts(const float* __restrict__ parts, int n, float* __restrict__ out, PeerPush push) {
    __shared__ float sm[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += parts[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            out[0] = v;
            peer_push_value(push, v);
        }
    }
}
void launch_sum_parts(const float* parts, int n, float* out, const PeerPush* push, cudaStream_t st) {
    PeerPush p;
    if (push != nullptr) p = *push; else memset(&p, 0, sizeof(p));
    k_sum_parts<<<1, 256, 0, st>>>(parts, n, out, p);
}

__global__ void k_fill_float(float* p, float v) { p[0] = v; }
void launch_fill_float(float* p, float v, cudaStream_t st) { k_fill_float<<<1, 1, 0, st>>>(p, v); }

// ------------------------------------------------------------------------------------------------ label-aware variants
// class ids padded to n_pad (rows with -2, columns with -1: padding never matches), lse -> log2 units
__global__ void k_label_prep(int n, int n_pad, const int* __restrict__ cls_a, const int* __restrict__ cls_b,
                             int* __restrict__ pa, int* __restrict__ pb, const float* __restrict__ lse,
                             float* __restrict__ lse2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    pa[i] = i < n ? cls_a[i] : -2;
    pb[i] = i < n ? cls_b[i] : -1;
    if (lse != nullptr) lse2[i] = i < n ? lse[i] * LOG2E_F : 0.f;
}
void launch_label_prep(int n, int n_pad, const int* cls_a, const int* cls_b, int* pa, int* pb, const float* lse,
                       float* lse2, cudaStream_t st) {
    k_label_prep<<<(n_pad + 255) / 256, 256, 0, st>>>(n, n_pad, cls_a, cls_b, pa, pb, lse, lse2);
}

// sums of the three accumulators of the forward kernel's mask_mode 2 over the (column block, half) partials;
// out0 = sum of the same-class logits (natural units), out1 = sum ln(1 - P), out2 = sum P / (1 - P)
__global__ void k_label_sum3(const float* __restrict__ px, const float* __restrict__ pl, const float* __restrict__ pr,
                             int n_parts, int ld, int n, float* __restrict__ out0, float* __restrict__ out1,
                             float* __restrict__ out2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int p = 0; p < n_parts; ++p) {
        a += px[(size_t)p * ld + i]; b += pl[(size_t)p * ld + i]; c += pr[(size_t)p * ld + i];
    }
    out0[i] = a * LN2_F; out1[i] = b * LN2_F; out2[i] = c;
}
void launch_label_sum3(const float* px, const float* pl, const float* pr, int n_parts, int ld, int n, float* out0,
                       float* out1, float* out2, cudaStream_t st) {
    k_label_sum3<<<(n + 255) / 256, 256, 0, st>>>(px, pl, pr, n_parts, ld, n, out0, out1, out2);
}

// vectors of a label-aware backward sweep, padded to n_pad; words[0] = bits(gmax[0]); fast path off
__global__ void k_label_sweep_prep(int n, int n_pad, const float* __restrict__ wr, const float* __restrict__ lr,
                                   const float* __restrict__ wc, const float* __restrict__ lc,
                                   const float* __restrict__ d_diag, const float* __restrict__ mk_r,
                                   const float* __restrict__ mk_c, const int* __restrict__ cls_a,
                                   const int* __restrict__ cls_b, const float* __restrict__ gmax, float* __restrict__ o_wr,
                                   float* __restrict__ o_lr, float* __restrict__ o_d, float* __restrict__ o_kr,
                                   int* __restrict__ o_ca, float* __restrict__ o_wc, float* __restrict__ o_lc,
                                   float* __restrict__ o_kc, int* __restrict__ o_cb, int* __restrict__ o_pos,
                                   uint32_t* __restrict__ words, float* __restrict__ fast_info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { words[0] = __float_as_uint(fabsf(gmax[0])); fast_info[0] = 0.f; fast_info[1] = 0.f; }
    if (i >= n_pad) return;
    const bool live = i < n;
    o_wr[i] = live ? wr[i] : 0.f; o_lr[i] = live ? lr[i] * LOG2E_F : 0.f;
    o_wc[i] = live ? wc[i] : 0.f; o_lc[i] = live ? lc[i] * LOG2E_F : 0.f;
    o_d[i] = (live && d_diag) ? d_diag[i] : 0.f;
    o_kr[i] = (live && mk_r) ? mk_r[i] : 0.f; o_kc[i] = (live && mk_c) ? mk_c[i] : 0.f;
    o_ca[i] = (live && cls_a) ? cls_a[i] : -2; o_cb[i] = (live && cls_b) ? cls_b[i] : -1;
    o_pos[i] = live ? i : -1;
}
void launch_label_sweep_prep(int n, int n_pad, const float* wr, const float* lr, const float* wc, const float* lc,
                             const float* d_diag, const float* mk_r, const float* mk_c, const int* cls_a, const int* cls_b,
                             const float* gmax, float* o_wr, float* o_lr, float* o_d, float* o_kr, int* o_ca, float* o_wc,
                             float* o_lc, float* o_kc, int* o_cb, int* o_pos, uint32_t* words, float* fast_info,
                             cudaStream_t st) {
    k_label_sweep_prep<<<(n_pad + 255) / 256, 256, 0, st>>>(n, n_pad, wr, lr, wc, lc, d_diag, mk_r, mk_c, cls_a, cls_b, gmax,
                                                            o_wr, o_lr, o_d, o_kr, o_ca, o_wc, o_lc, o_kc, o_cb, o_pos,
                                                            words, fast_info);
}

// ------------------------------------------------------------------------------------------------ fp16 staging copy
// The dA MMA multiplies the fp16-staged dS tile with the features, and tcgen05 kind::f16 needs both operands in the
// same 16-bit format, so the backward keeps an fp16 copy of the features (exact for bf16 inputs within fp16 range).
template <bool F32>
__global__ void k_to_f16(const void* __restrict__ src, size_t n8, void* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    float v[8];
    load8<F32>(src, i * 8, v);
    uint4 u;
    uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        uint32_t r;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[2 * e + 1]), "f"(v[2 * e]));
        w[e] = r;
    }
    reinterpret_cast<uint4*>(dst)[i] = u;
}
void launch_to_f16(const void* src, int dtype, size_t n_elems, void* dst, cudaStream_t st) {
    const size_t n8 = n_elems / 8;
    if (n8 == 0) return;
    const unsigned grid = (unsigned)((n8 + 255) / 256);
    if (dtype == 1) k_to_f16<true><<<grid, 256, 0, st>>>(src, n8, dst);
    else k_to_f16<false><<<grid, 256, 0, st>>>(src, n8, dst);
}

// ------------------------------------------------------------------------------------------------ fp32 -> planes
// x (fp32 [n][dim]) -> NP planes of 16-bit floats side by side: out[n][NP * dp], plane p at columns [p*dp, p*dp+dim),
// x = x_1 + x_2 (+ x_3) with x_1 = round(x), x_2 = round(x - x_1), ...  Columns [dim, dp) of every plane are zero.
template <int NP, bool BF16>
__global__ void k_split_planes(const float* __restrict__ x, int n, int dim, int dp, uint16_t* __restrict__ out,
                               PeerWait wait) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per 8 columns of dp
    const int cols8 = dp / 8;
    if (wait.flags != nullptr) {
        // multi-GPU: the rows are being written by other ranks - wait for the owners of this block's rows (peer.cuh)
        if (threadIdx.x == 0) {
            const size_t i0 = (size_t)blockIdx.x * blockDim.x, i1 = i0 + blockDim.x - 1;
            const int r0 = (int)(i0 / cols8), r1 = (int)(i1 / cols8);
            peer_wait_rows(wait, r0, (r1 < n ? r1 : n - 1) + 1);
        }
        __syncthreads();
    }
    if (i >= (size_t)n * cols8) return;
    const int row = (int)(i / cols8), c0 = (int)(i - (size_t)row * cols8) * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (c0 + e < dim) ? __ldcg(x + (size_t)row * dim + c0 + e) : 0.f;
#pragma unroll
    for (int pl = 0; pl < NP; ++pl) {
        uint4 u;
        uint16_t* h = reinterpret_cast<uint16_t*>(&u);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float r;
            if (BF16) { const __nv_bfloat16 b = __float2bfloat16_rn(v[e]); h[e] = *reinterpret_cast<const uint16_t*>(&b); r = __bfloat162float(b); }
            else { const __half b = __float2half_rn(v[e]); h[e] = *reinterpret_cast<const uint16_t*>(&b); r = __half2float(b); }
            v[e] -= r;
        }
        *reinterpret_cast<uint4*>(out + (size_t)row * NP * dp + (size_t)pl * dp + c0) = u;
    }
}
static PeerWait wait_or_none(const PeerWait* w) {
    PeerWait r;
    if (w != nullptr) r = *w; else memset(&r, 0, sizeof(r));
    return r;
}
void launch_split_planes_bf16x3(const float* x, int n, int dim, int dp, void* out, cudaStream_t st, const PeerWait* wait) {
    const size_t t = (size_t)n * (dp / 8);
    if (t) k_split_planes<3, true><<<(unsigned)((t + 255) / 256), 256, 0, st>>>(x, n, dim, dp, static_cast<uint16_t*>(out),
                                                                                 wait_or_none(wait));
}
void launch_split_planes_f16x2(const float* x, int n, int dim, int dp, void* out, cudaStream_t st, const PeerWait* wait) {
    const size_t t = (size_t)n * (dp / 8);
    if (t) k_split_planes<2, false><<<(unsigned)((t + 255) / 256), 256, 0, st>>>(x, n, dim, dp, static_cast<uint16_t*>(out),
                                                                                  wait_or_none(wait));
}

// ------------------------------------------------------------------------------------------------ L2 normalise
// One warp per RPW rows, 128-bit accesses.  Rows of up to CH * 256 elements (CH <= 4) stay in registers - in their
// STORAGE format (a bf16 chunk of 8 elements is one uint4) - between the reduction and the scaling, so every operand
// crosses HBM exactly once: 2 n D e bytes forward (read x, write y), 3 n D e bytes backward (read y and dy, write dx).
// RPW is chosen so that every lane keeps >= 64 bytes per operand in flight (short bf16 rows: four rows per warp); rows
// longer than 1024 elements re-read the tail (L2 hit).
template <bool F32> struct Pk8;
template <> struct Pk8<true> { float4 a, b; };
template <> struct Pk8<false> { uint4 u; };
template <bool F32>
__device__ __forceinline__ Pk8<F32> pk_load(const void* base, size_t elem);
template <>
__device__ __forceinline__ Pk8<true> pk_load<true>(const void* base, size_t elem) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem);
    Pk8<true> r; r.a = __ldg(p); r.b = __ldg(p + 1); return r;
}
template <>
__device__ __forceinline__ Pk8<false> pk_load<false>(const void* base, size_t elem) {
    Pk8<false> r; r.u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem)); return r;
}
__device__ __forceinline__ void pk_unpack(const Pk8<true>& p, float (&v)[8]) {
    v[0] = p.a.x; v[1] = p.a.y; v[2] = p.a.z; v[3] = p.a.w; v[4] = p.b.x; v[5] = p.b.y; v[6] = p.b.z; v[7] = p.b.w;
}
__device__ __forceinline__ void pk_unpack(const Pk8<false>& p, float (&v)[8]) {
    const uint32_t w[4] = {p.u.x, p.u.y, p.u.z, p.u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(w[e] << 16); v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
}
template <bool F32>
__device__ __forceinline__ Pk8<F32> pk_zero() { Pk8<F32> r; memset(&r, 0, sizeof(r)); return r; }

template <bool F32, int CH, int RPW>
__global__ void __launch_bounds__(256) k_l2norm_fwd(const void* __restrict__ x, int n, int dim, void* __restrict__ y,
                                                    float* __restrict__ inv_norm) {
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    const int lane = threadIdx.x & 31;
    if (row0 >= n) return;
    Pk8<F32> v[RPW][CH];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const bool live = row0 + r < n;
        const size_t base = (size_t)(row0 + r) * dim;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int d = c * 256 + lane * 8;
            v[r][c] = (live && d < dim) ? pk_load<F32>(x, base + d) : pk_zero<F32>();
        }
    }
    float ss[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        ss[r] = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float t[8];
            pk_unpack(v[r][c], t);
#pragma unroll
            for (int e = 0; e < 8; ++e) ss[r] = fmaf(t[e], t[e], ss[r]);
        }
        if (row0 + r < n)
            for (int d = CH * 256 + lane * 8; d < dim; d += 256) {       // rows longer than CH * 256 (CH == 4 only)
                float t[8];
                load8<F32>(x, (size_t)(row0 + r) * dim + d, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) ss[r] = fmaf(t[e], t[e], ss[r]);
            }
        ss[r] = warp_sum(ss[r]);
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        if (row0 + r >= n) break;
        const size_t base = (size_t)(row0 + r) * dim;
        const float inv = 1.0f / sqrtf(ss[r]);   // no epsilon: matches x / x.norm(dim=-1, keepdim=True)
        if (lane == 0 && inv_norm) inv_norm[row0 + r] = inv;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int d = c * 256 + lane * 8;
            if (d < dim) {
                float t[8];
                pk_unpack(v[r][c], t);
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] *= inv;
                store8<F32>(y, base + d, t);
            }
        }
        for (int d = CH * 256 + lane * 8; d < dim; d += 256) {
            float t[8];
            load8<F32>(x, base + d, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] *= inv;
            store8<F32>(y, base + d, t);
        }
    }
}

template <bool F32, int CH, int RPW>
__global__ void __launch_bounds__(256) k_l2norm_bwd(const void* __restrict__ y, const void* __restrict__ dy,
                                                    const float* __restrict__ inv_norm, int n, int dim,
                                                    void* __restrict__ dx) {
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    const int lane = threadIdx.x & 31;
    if (row0 >= n) return;
    Pk8<F32> a[RPW][CH], b[RPW][CH];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const bool live = row0 + r < n;
        const size_t base = (size_t)(row0 + r) * dim;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int d = c * 256 + lane * 8;
            const bool ok = live && d < dim;
            a[r][c] = ok ? pk_load<F32>(y, base + d) : pk_zero<F32>();
            b[r][c] = ok ? pk_load<F32>(dy, base + d) : pk_zero<F32>();
        }
    }
    float dot[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        dot[r] = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float p[8], q[8];
            pk_unpack(a[r][c], p);
            pk_unpack(b[r][c], q);
#pragma unroll
            for (int e = 0; e < 8; ++e) dot[r] = fmaf(p[e], q[e], dot[r]);
        }
        if (row0 + r < n)
            for (int d = CH * 256 + lane * 8; d < dim; d += 256) {
                float p[8], q[8];
                load8<F32>(y, (size_t)(row0 + r) * dim + d, p);
                load8<F32>(dy, (size_t)(row0 + r) * dim + d, q);
#pragma unroll
                for (int e = 0; e < 8; ++e) dot[r] = fmaf(p[e], q[e], dot[r]);
            }
        dot[r] = warp_sum(dot[r]);
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        if (row0 + r >= n) break;
        const size_t base = (size_t)(row0 + r) * dim;
        const float inv = inv_norm[row0 + r];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int d = c * 256 + lane * 8;
            if (d < dim) {
                float p[8], q[8], o[8];
                pk_unpack(a[r][c], p);
                pk_unpack(b[r][c], q);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (q[e] - p[e] * dot[r]) * inv;
                store8<F32>(dx, base + d, o);
            }
        }
        for (int d = CH * 256 + lane * 8; d < dim; d += 256) {
            float p[8], q[8], o[8];
            load8<F32>(y, base + d, p);
            load8<F32>(dy, base + d, q);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (q[e] - p[e] * dot[r]) * inv;
            store8<F32>(dx, base + d, o);
        }
    }
}

template <bool F32, int CH, int RPW>
static void l2norm_launch_fwd(const void* x, int n, int dim, void* y, float* inv_norm, cudaStream_t st) {
    const int rows_per_block = 8 * RPW;
    k_l2norm_fwd<F32, CH, RPW><<<(n + rows_per_block - 1) / rows_per_block, 256, 0, st>>>(x, n, dim, y, inv_norm);
}
template <bool F32, int CH, int RPW>
static void l2norm_launch_bwd(const void* y, const void* dy, const float* inv_norm, int n, int dim, void* dx,
                              cudaStream_t st) {
    const int rows_per_block = 8 * RPW;
    k_l2norm_bwd<F32, CH, RPW><<<(n + rows_per_block - 1) / rows_per_block, 256, 0, st>>>(y, dy, inv_norm, n, dim, dx);
}

// rows per warp: fp32 2 / 2 / 1 / 1, bf16 4 / 2 / 2 / 2 for CH = 1 / 2 / 3 / 4 (measured: more rows per warp cost
// occupancy through registers and lose bandwidth)
void launch_l2norm_fwd(const void* x, int n, int dim, int dtype, void* y, float* inv_norm, cudaStream_t st) {
    if (n <= 0) return;
    const int ch = (dim + 255) / 256;
    if (dtype == 1) {
        if (ch <= 1) l2norm_launch_fwd<true, 1, 2>(x, n, dim, y, inv_norm, st);
        else if (ch == 2) l2norm_launch_fwd<true, 2, 2>(x, n, dim, y, inv_norm, st);
        else if (ch == 3) l2norm_launch_fwd<true, 3, 1>(x, n, dim, y, inv_norm, st);
        else l2norm_launch_fwd<true, 4, 1>(x, n, dim, y, inv_norm, st);
    } else {
        if (ch <= 1) l2norm_launch_fwd<false, 1, 4>(x, n, dim, y, inv_norm, st);
        else if (ch == 2) l2norm_launch_fwd<false, 2, 2>(x, n, dim, y, inv_norm, st);
        else if (ch == 3) l2norm_launch_fwd<false, 3, 2>(x, n, dim, y, inv_norm, st);
        else l2norm_launch_fwd<false, 4, 2>(x, n, dim, y, inv_norm, st);
    }
}

void launch_l2norm_bwd(const void* y, const void* dy, const float* inv_norm, int n, int dim, int dtype, void* dx,
                       cudaStream_t st) {
    if (n <= 0) return;
    const int ch = (dim + 255) / 256;
    if (dtype == 1) {
        if (ch <= 1) l2norm_launch_bwd<true, 1, 2>(y, dy, inv_norm, n, dim, dx, st);
        else if (ch == 2) l2norm_launch_bwd<true, 2, 2>(y, dy, inv_norm, n, dim, dx, st);
        else if (ch == 3) l2norm_launch_bwd<true, 3, 1>(y, dy, inv_norm, n, dim, dx, st);
        else l2norm_launch_bwd<true, 4, 1>(y, dy, inv_norm, n, dim, dx, st);
    } else {
        if (ch <= 1) l2norm_launch_bwd<false, 1, 4>(y, dy, inv_norm, n, dim, dx, st);
        else if (ch == 2) l2norm_launch_bwd<false, 2, 2>(y, dy, inv_norm, n, dim, dx, st);
        else if (ch == 3) l2norm_launch_bwd<false, 3, 2>(y, dy, inv_norm, n, dim, dx, st);
        else l2norm_launch_bwd<false, 4, 2>(y, dy, inv_norm, n, dim, dx, st);
    }
}

}  // namespace flyp
