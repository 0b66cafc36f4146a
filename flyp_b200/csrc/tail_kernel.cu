// tail_kernel.cu — the encoder tail of the FLYP towers, fused (sm_100a only):
//
//     clip/model.py:242-243   x = x @ self.proj                  (vision tower;  :359  x[eot] @ self.text_projection)
//     clip/model.py:375-376   features / features.norm(dim=-1, keepdim=True)
//
// y[n, N] = z / ||z||_2 with z = x[n, K] . W[K, N]: a tcgen05 GEMM whose 128 x N fp32 accumulator tile stays in tensor
// memory (N <= 512 columns per CTA; N up to 1024 on a cluster of two CTAs that exchange the row norms through
// distributed shared memory), and an epilogue that reduces the sum of squares per row, scales, and writes the features
// in the loss's storage format (bf16 or fp32, optionally the fp16 copy the backward sweeps multiply) - z never exists in
// HBM.  fp32 inputs are evaluated as 3-way bf16 split products (six terms, fp32 accumulation) like the loss itself.
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..11 epilogue
// (thread = TMEM lane = row; two warps per lane quarter split the columns).
#include "clip_kernels.cuh"
#include "sm100.cuh"
#include "tail_kernel.cuh"

namespace flyp {
using namespace sm100;

namespace {
constexpr int TAIL_THREADS = 384;
constexpr int XBOX = TILE * KCHUNK * 2;      // x chunk [128 rows][64 k] bf16 = 16 KiB (K-major)
constexpr int WBOX = KCHUNK * 64 * 2;        // W box [64 k rows][64 n cols] bf16 = 8 KiB (MN-major)

DEVI uint8_t* align1024t(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
DEVI void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
DEVI float ld_dsmem_f32(uint32_t cluster_addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
    return v;
}
}  // namespace

// NSPLIT CTAs (a cluster when NSPLIT == 2) share a 128-row tile: CTA c computes the output columns
// [c * n_cta, (c + 1) * n_cta).
template <int NSPLIT>
__device__ __forceinline__ void tail_body(const CUtensorMap& tmX, const CUtensorMap& tmW, const TailParams& p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024t(smem_raw);
    const int nbox = p.n_cta / 64;                         // W boxes per K chunk
    const int stage_bytes = XBOX + nbox * WBOX;
    const int stages = p.stages;
    uint8_t* ring = smem;
    float* red = reinterpret_cast<float*>(ring + (size_t)stages * stage_bytes);      // [2][128] partial sums of squares
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 2 * TILE);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (8 + s); };
    const uint32_t ACCFULL = bar0 + 8u * 16;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = NSPLIT == 2 ? (int)cluster_ctarank() : 0;
    const int tile = NSPLIT == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int row0 = tile * TILE, col0 = cta * p.n_cta;
    const int kc_total = p.kc * p.kplan.n_terms;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(ACCFULL, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW);
    }
    if (warp == 2) { tmem_alloc(smem_u32(tmem_holder), 512); tmem_relinquish(); }
    tc_fence_before();
    if (NSPLIT == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int term = 0; term < p.kplan.n_terms; ++term) {
                const int xa = p.kplan.pa[term] * p.kplan.plane_cols;      // column offset of the x plane
                const int wb = p.kplan.pb[term] * p.w_plane_rows;          // row offset of the W plane
                for (int c = 0; c < p.kc; ++c) {
                    mbar_wait(EMPTY(stage), phase ^ 1);
                    mbar_expect_tx(FULL(stage), stage_bytes);
                    uint8_t* dst = ring + (size_t)stage * stage_bytes;
                    tma_load_2d(smem_u32(dst), &tmX, FULL(stage), xa + c * KCHUNK, row0);
                    for (int j = 0; j < nbox; ++j)
                        tma_load_2d(smem_u32(dst + XBOX + j * WBOX), &tmW, FULL(stage), col0 + j * 64, wb + c * KCHUNK);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int c = 0; c < kc_total; ++c) {
                mbar_wait(FULL(stage), phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(ring + (size_t)stage * stage_bytes);
                const uint32_t b_addr = a_addr + XBOX;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = umma_desc_sw128(a_addr + k * 32, 16, 1024);
                    // up to 256 output columns per instruction: W [16 k rows][N], MN-major, 64-column boxes WBOX apart
                    for (int n0 = 0; n0 < p.n_cta; n0 += 256) {
                        const int nn = p.n_cta - n0 < 256 ? p.n_cta - n0 : 256;
                        const uint64_t bd = umma_desc_sw128(b_addr + (n0 / 64) * WBOX + k * 16 * 128, WBOX, 1024);
                        umma_bf16(tmem_base + n0, ad, bd, umma_idesc(TILE, nn, 1, 1, 0, 1), (c | k) != 0);
                    }
                }
                umma_commit(EMPTY(stage));
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(ACCFULL);
        }
    }
    // ---- epilogue, pass 1: sum of squares of this thread's row over this warp's columns
    const int q = warp & 3, h = warp >= 4 ? (warp - 4) >> 2 : 0;
    const int rloc = q * 32 + lane, row = row0 + rloc;
    const int ncols = p.n_cta / 2, c_begin = h * ncols;       // this warp's columns of the CTA's tile
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + c_begin;
    if (warp >= 4) {
        mbar_wait(ACCFULL, 0);
        tc_fence_after();
        float ss = 0.f;
        for (int cc = 0; cc < ncols; cc += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + cc, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) { const float v = __uint_as_float(r[k]); ss = fmaf(v, v, ss); }
        }
        red[h * TILE + rloc] = ss;
        epi_bar();
    }
    // the other CTA of the cluster holds the other half of every row: the partial sums are exchanged through
    // distributed shared memory (every thread of both CTAs takes part in the cluster barrier)
    if (NSPLIT == 2) cluster_sync_all();
    if (warp >= 4) {
        // ---- pass 2: scale and store
        float tot = red[rloc] + red[TILE + rloc];
        if (NSPLIT == 2) {
            const uint32_t peer = mapa(smem_u32(red), (uint32_t)(cta ^ 1));
            tot += ld_dsmem_f32(peer + 4u * rloc) + ld_dsmem_f32(peer + 4u * (TILE + rloc));
        }
        const float inv = 1.0f / sqrtf(tot);               // no epsilon: x / x.norm(dim=-1, keepdim=True)
        if (row < p.n && h == 0 && cta == 0 && p.inv_norm != nullptr) p.inv_norm[row] = inv;
        for (int cc = 0; cc < ncols; cc += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + cc, r);
            tmem_ld_wait();
            if (row >= p.n) continue;
            const int col = col0 + c_begin + cc;
            if (col >= p.n_out) continue;
            const size_t o = (size_t)row * p.n_out + col;
            if (p.y_fp32) {
                float* out = reinterpret_cast<float*>(p.y) + o;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(out + 4 * j) =
                        make_float4(__uint_as_float(r[4 * j]) * inv, __uint_as_float(r[4 * j + 1]) * inv,
                                    __uint_as_float(r[4 * j + 2]) * inv, __uint_as_float(r[4 * j + 3]) * inv);
            } else {
                uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + o);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
                    u.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
                    u.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
                    u.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
                    out[j] = u;
                }
            }
            if (p.y16 != nullptr) {
                // fp16 copy of the ROUNDED bf16 features (what the loss's backward sweeps multiply)
                uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y16) + o);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t b2 = pack_bf16x2(__uint_as_float(r[8 * j + 2 * e]) * inv,
                                                        __uint_as_float(r[8 * j + 2 * e + 1]) * inv);
                        w[e] = pack_f16x2(__uint_as_float(b2 << 16), __uint_as_float(b2 & 0xffff0000u));
                    }
                    out[j] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tc_fence_before();
    if (NSPLIT == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

__global__ void __launch_bounds__(TAIL_THREADS, 1)
tail_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ TailParams p) {
    tail_body<1>(tmX, tmW, p);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TAIL_THREADS, 1)
tail_kernel_c2(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ TailParams p) {
    tail_body<2>(tmX, tmW, p);
}

int tail_n_split(int n_out) { return n_out > 512 ? 2 : 1; }

size_t tail_smem_bytes(int n_cta, int* stages_out) {
    const int stage = XBOX + (n_cta / 64) * WBOX;
    int stages = (200 * 1024) / stage;
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    if (stages_out) *stages_out = stages;
    return (size_t)stages * stage + 2 * TILE * sizeof(float) + 256 + 1024;
}

void launch_tail(const CUtensorMap& tmX, const CUtensorMap& tmW, TailParams p, cudaStream_t st) {
    const int split = tail_n_split(p.n_out);
    int stages = 2;
    const size_t smem = tail_smem_bytes(p.n_cta, &stages);
    p.stages = stages;
    const int tiles = (p.n + TILE - 1) / TILE;
    if (split == 1) {
        static bool attr_done[64] = {false};
        ensure_smem_attr(tail_kernel, smem, attr_done);
        tail_kernel<<<tiles, TAIL_THREADS, smem, st>>>(tmX, tmW, p);
    } else {
        static bool attr_done[64] = {false};
        ensure_smem_attr(tail_kernel_c2, smem, attr_done);
        tail_kernel_c2<<<tiles * 2, TAIL_THREADS, smem, st>>>(tmX, tmW, p);
    }
}

}  // namespace flyp
