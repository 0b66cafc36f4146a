// clip_dst_gemm.cu — gradient products over the KEPT dS matrix (tcgen05 cta_group::2, fp16 x fp16 -> fp32).
//
// The backward sweep (clip_bwd_pair.cu, keep_ds) leaves dS = d(loss)/d(logits) in global memory as the staged fp16
// values its own dS . B product consumed ([n_m rows][n_n columns], scaled by the staging factor G).  The gradient of the
// OTHER operand is then a plain product over those values,
//
//     transposed = 1:  out[n, :] = c * sum_m dS[m, n] X16[m, :]      (d_text  = s dS^T I: contraction over the rows)
//     transposed = 0:  out[m, :] = c * sum_n dS[m, n] X16[n, :]      (d_image = s dS T:   contraction over the columns)
//
// c = scale * out_mul / G, instead of a second sweep that recomputes every logit, its exponentials and its dS: the step
// executes 8 B^2 D FLOPs, exactly the algorithmic count (forward S, one recompute, dI, dT), not 10 B^2 D.
//
// One cluster of two CTAs owns an output tile of 256 rows (128 per CTA = its 128 TMEM lanes) x up to 512 columns
// (the whole tensor memory: 2 x 256 fp32 columns) and walks over the contraction index in blocks of 128.
// Per block and CTA the TMA producer loads two [128][64] boxes of dS (its half of the 256 output rows) and, for every
// 256-column accumulator, two [128][64] boxes of X16 (its half of the columns: the B operand of a pair MMA is split
// across the two CTAs by halves of N), 16 KiB each, through one ring of 12 slots.  The operands need no transposition:
//   transposed:      dS boxes are [128 contraction rows][64 output rows]  -> A operand MN-major (two boxes = M 128)
//   not transposed:  dS boxes are [128 output rows][64 contraction cols]  -> A operand K-major  (two boxes = K 128)
//   X16 boxes are   [128 contraction rows][64 feature columns]            -> B operand MN-major (two boxes = N 128)
// Schedule: the (tile, contraction block) units are cut like the sweep's (sched.h: SweepItems): whole tiles in rounds,
// then a flat tail in P equal ranges whose fp32 partial tiles are summed, in pair order, by the whole grid behind a grid
// barrier (cooperative launch).
#include "clip_kernels.cuh"
#include "sm100.cuh"
#include "bwd_common.cuh"
#include "sched.h"
#include <cstring>

namespace flyp {
using namespace sm100;

namespace {
constexpr int GT = 384;                 // threads: warp 0 producer, 1 MMA issuer, 2 TMEM, 4..11 epilogue
constexpr int G_SLOT = 16384;
constexpr int G_NSLOT = 12;
constexpr int G_ROWS = DST_TILE_ROWS;   // output rows per tile (pair)
constexpr int G_COLS = DST_TILE_COLS;   // output columns per pass
constexpr int G_KB = 128;               // contraction rows per block
constexpr int G_STAGE = 8 * 32 * 32 * 4;   // epilogue: one [32][32] fp32 transposition buffer per warp
constexpr int G_SMEM = G_NSLOT * G_SLOT + G_STAGE + 1024 /*barriers, scratch*/ + 1024 /*alignment*/;
static_assert(G_SMEM <= 232448, "shared memory budget of one CTA");

DEVI uint8_t* g_align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
DEVI void g_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
DEVI int g_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// start of output row `orow` when the output is fp32 (possibly scattered over the ranks' buffers)
DEVI float* out_row_f32(const DstParams& p, int orow) {
    if (p.rows_per_rank > 0) {
        const int q = orow / p.rows_per_rank;
        return p.out_rank[q] + (size_t)(orow - q * p.rows_per_rank) * p.ld_out;
    }
    return reinterpret_cast<float*>(p.out) + (size_t)orow * p.ld_out;
}

// The fp32 partial tiles of the split tail tiles, summed by the whole grid (cf. sweep_tail_reduce in clip_bwd_pair.cu).
__device__ __noinline__ void gemm_tail_reduce(const DstParams& p, const int et, const int NJ, int* red_i) {
    __threadfence();
    g_epi_bar();
    if (et == 0) {
        atomicAdd(p.grid_cnt, 1);
        while (g_ld_acquire(p.grid_cnt) < (int)gridDim.x) __nanosleep(40);
    }
    g_epi_bar();
    __threadfence();
    const int v_tiles = p.out_tiles * p.n_dh, P = p.sched_pairs;
    const int first = (v_tiles / P) * P, n_tail = v_tiles - first;
    const int COLS = p.tile_cols;
    const int CPB = (G_ROWS / 128) * (COLS / 128);              // [128 x 128] chunks per tile
    for (int c = blockIdx.x; c < n_tail * CPB; c += gridDim.x) {
        const int tb = c / CPB, ci = c - tb * CPB;
        const int r0 = (ci / (COLS / 128)) * 128, dl0 = (ci % (COLS / 128)) * 128;
        if (et == 0) {
            TailParts parts;
            int np = 0;
            if (parts.init(v_tiles, NJ, P, tb))
                for (int sl = parts.next(); sl >= 0; sl = parts.next()) red_i[2 + np++] = sl;
            red_i[1] = np;
        }
        g_epi_bar();
        const int np = red_i[1];
        const int vb = first + tb, ob = vb / p.n_dh, dh = vb - ob * p.n_dh;
        if (np > 0 && p.col_begin + dh * COLS + dl0 < p.dim) {
#pragma unroll 1
            for (int f0 = et; f0 < 128 * 32; f0 += 256 * 8) {
                float4 acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < np; ++k) {
                    const float* base = p.part_out + ((size_t)red_i[2 + k] * G_ROWS + r0) * COLS + dl0;
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int f = f0 + j * 256, ri = f >> 5;
                        v[j] = (ob * G_ROWS + r0 + ri < p.n_out)
                                   ? __ldcg(reinterpret_cast<const float4*>(base + (size_t)ri * COLS + (f & 31) * 4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { acc[j].x += v[j].x; acc[j].y += v[j].y; acc[j].z += v[j].z; acc[j].w += v[j].w; }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int f = f0 + j * 256, ri = f >> 5, oi = ob * G_ROWS + r0 + ri;
                    const int d = p.col_begin + dh * COLS + dl0 + (f & 31) * 4;
                    if (oi >= p.n_out) continue;
                    if (p.out_fp32) {
                        *reinterpret_cast<float4*>(out_row_f32(p, oi) + d) = acc[j];
                    } else {
                        uint2 u;
                        u.x = pack_bf16x2(acc[j].x, acc[j].y); u.y = pack_bf16x2(acc[j].z, acc[j].w);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)oi * p.ld_out + d) = u;
                    }
                }
            }
        }
        g_epi_bar();
    }
}

}  // namespace

int dst_gemm_sched_pairs(int v_tiles, int k_blocks, int num_sms) { return dst_sched_pairs(v_tiles, k_blocks, num_sms / 2); }
size_t dst_gemm_part_floats(int pairs) { return (size_t)2 * pairs * G_ROWS * G_COLS; }

template <bool TRANSPOSED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GT, 1)
dst_gemm_kernel(const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ DstParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = g_align1024(smem_raw);
    float* stage = reinterpret_cast<float*>(ring + G_NSLOT * G_SLOT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + G_NSLOT * G_SLOT + G_STAGE);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (G_NSLOT + s); };
    auto ACCFULL = [&](int s) { return bar0 + 8u * (2 * G_NSLOT + s); };
    auto ACCEMPTY = [&](int s) { return bar0 + 8u * (2 * G_NSLOT + 2 + s); };
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * G_NSLOT + 4);
    int* red_i = reinterpret_cast<int*>(bars + 2 * G_NSLOT + 6);      // [1] slot count, [2 ..] slot list (<= 2 P entries)
    // tile width: 512 columns = the whole tensor memory, one accumulator stage; or 256 columns in TWO stages, so that the
    // drain of a tile - slow when it goes to another GPU over NVLink - overlaps the MMAs of the next one
    const int COLS = p.tile_cols, STAGES = G_COLS / COLS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int NJ = (p.n_k + G_KB - 1) / G_KB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_NSLOT; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(ACCFULL(s), 1); mbar_init(ACCEMPTY(s), 512); }
        fence_mbar_init();
        tma_prefetch_desc(&tmDS); tma_prefetch_desc(&tmX);
    }
    if (warp == 2) { tmem_alloc_cg2(smem_u32(tmem_holder), 512); tmem_relinquish_cg2(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    // accumulators (256 columns each) of pass dh
    auto n_acc = [&](int dh) { const int left = p.dim - p.col_begin - dh * COLS; return left >= COLS ? COLS / 256 : (left + 255) / 256; };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs, own halves)
        if (elect_one()) {
            int slot = 0; uint32_t ph = 0;
            auto put = [&](const CUtensorMap* tm, int c0, int c1) {
                mbar_wait(EMPTY(slot), ph ^ 1);
                if (cta == 0) mbar_expect_tx(FULL(slot), 2 * G_SLOT);
                tma_load_2d_cg2(smem_u32(ring + slot * G_SLOT), tm, mapa(FULL(slot), 0), c0, c1);
                if (++slot == G_NSLOT) { slot = 0; ph ^= 1; }
            };
            SweepItems iter(p.out_tiles, p.n_dh, p.sched_pairs, NJ, pair);
            ItemInfo ii;
            while (iter.next(ii)) {
                const int o0 = ii.mb * G_ROWS + (int)cta * 128;      // first output row of this CTA
                const int na = n_acc(ii.dh);
                for (int kb = ii.t0; kb < ii.t1; ++kb) {
                    const int k0 = kb * G_KB;
                    if (TRANSPOSED) { put(&tmDS, o0, k0); put(&tmDS, o0 + 64, k0); }
                    else { put(&tmDS, k0, o0); put(&tmDS, k0 + 64, o0); }
                    for (int a = 0; a < na; ++a)
                        for (int ds = 0; ds < 2; ++ds)
                            put(&tmX, p.col_begin + ii.dh * COLS + a * 256 + (int)cta * 128 + ds * 64, k0);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (cta == 0 && elect_one()) {
            constexpr uint32_t IDESC = umma_idesc(256, 256, 0, 0, TRANSPOSED ? 1 : 0, 1);   // fp16 x fp16, B MN-major
            int slot = 0; uint32_t ph = 0; uint32_t it = 0;
            auto adv = [&]() { if (++slot == G_NSLOT) { slot = 0; ph ^= 1; } };
            constexpr uint32_t HI = umma_desc_hi(1024);
            const uint32_t ring_lo_mn = umma_desc_lo(smem_u32(ring), G_SLOT), ring_lo_k = umma_desc_lo(smem_u32(ring), 16);
            SweepItems iter(p.out_tiles, p.n_dh, p.sched_pairs, NJ, pair);
            ItemInfo ii;
            for (; iter.next(ii); ++it) {
                const int na = n_acc(ii.dh);
                const int as = (int)(it % (uint32_t)STAGES);
                const uint32_t use = it / (uint32_t)STAGES;
                mbar_wait(ACCEMPTY(as), (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t tm_acc = tmem_base + as * COLS;
                for (int kb = ii.t0; kb < ii.t1; ++kb) {
                    mbar_wait(FULL(slot), ph);
                    const int a_slot = slot;
                    adv();
                    mbar_wait(FULL(slot), ph);
                    adv();
                    // (descriptors: base word + an integer add per MMA, sm100.cuh)
                    const uint32_t a_lo = (TRANSPOSED ? ring_lo_mn : ring_lo_k) + a_slot * (G_SLOT >> 4);
                    for (int a = 0; a < na; ++a) {
                        mbar_wait(FULL(slot), ph);
                        const int b_slot = slot;
                        adv();
                        mbar_wait(FULL(slot), ph);
                        adv();
                        tc_fence_after();
                        const uint32_t b_lo = ring_lo_mn + b_slot * (G_SLOT >> 4);
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            // A, transposed: [16 contraction rows][128 output rows as two 64-wide boxes], MN-major;
                            //    otherwise:  [128 output rows][16 contraction columns] of box kk / 4, K-major
                            const uint64_t ad = umma_desc_join(
                                TRANSPOSED ? a_lo + kk * (2048 >> 4) : a_lo + (kk >> 2) * (G_SLOT >> 4) + (kk & 3) * 2, HI);
                            // B: [16 contraction rows][128 feature columns as two 64-wide boxes], MN-major
                            const uint64_t bd = umma_desc_join(b_lo + kk * (2048 >> 4), HI);
                            umma_f16_cg2(tm_acc + a * 256, ad, bd, IDESC, !(kb == ii.t0 && kk == 0));
                        }
                        umma_commit_cg2(EMPTY(b_slot));
                        umma_commit_cg2(EMPTY(b_slot + 1));
                    }
                    umma_commit_cg2(EMPTY(a_slot));
                    umma_commit_cg2(EMPTY(a_slot + 1));
                }
                umma_commit_cg2(ACCFULL(as));
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (both CTAs)
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;
        float G, invG;
        staging_scale(p.gmax_bits, G, invG);
        const float omul = *p.scale * p.out_mul * invG;
        const uint32_t R_ACCEMPTY0 = mapa(ACCEMPTY(0), 0), R_ACCEMPTY1 = mapa(ACCEMPTY(1), 0);
        float* const stg = stage + (warp - 4) * 1024;        // this warp's [32][32] fp32 transposition buffer
        uint32_t it = 0;
        SweepItems iter(p.out_tiles, p.n_dh, p.sched_pairs, NJ, pair);
        ItemInfo ii;
        for (; iter.next(ii); ++it) {
            const int na = n_acc(ii.dh);
            const int as = (int)(it % (uint32_t)STAGES);
            const uint32_t use = it / (uint32_t)STAGES;
            mbar_wait(ACCFULL(as), use & 1);
            tc_fence_after();
            for (int a = 0; a < na; ++a) {
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int dl = a * 256 + h * 128 + cc * 32;      // first of this warp's 32 columns within the pass
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * COLS + dl, r);
                    tmem_ld_wait();
                    // [32 rows (lanes)][32 columns] -> shared memory (XOR-swizzled: conflict-free both ways) -> each
                    // store instruction of the warp writes four full 128-byte row segments instead of 32 scattered
                    // 16-byte pieces (what a lane = row mapping gives): it matters for the partials that go to another
                    // GPU's memory over NVLink, where small scattered writes run at a fraction of the link rate
#pragma unroll
                    for (int k = 0; k < 32; ++k) stg[lane * 32 + (k ^ lane)] = __uint_as_float(r[k]) * omul;
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = i * 4 + (lane >> 3), c0 = (lane & 7) * 4;
                        float4 v;
                        v.x = stg[rr * 32 + ((c0 + 0) ^ rr)]; v.y = stg[rr * 32 + ((c0 + 1) ^ rr)];
                        v.z = stg[rr * 32 + ((c0 + 2) ^ rr)]; v.w = stg[rr * 32 + ((c0 + 3) ^ rr)];
                        const int rt = (int)cta * 128 + q * 32 + rr;         // row within the tile
                        const int orow = ii.mb * G_ROWS + rt;
                        const int d = p.col_begin + ii.dh * COLS + dl + c0;
                        if (orow >= p.n_out || d >= p.dim) continue;
                        if (ii.part >= 0) {
                            *reinterpret_cast<float4*>(p.part_out + ((size_t)ii.part * G_ROWS + rt) * COLS + dl + c0) = v;
                        } else if (p.out_fp32) {
                            *reinterpret_cast<float4*>(out_row_f32(p, orow) + d) = v;
                        } else {
                            uint2 u;
                            u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)orow * p.ld_out + d) = u;
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(as ? R_ACCEMPTY1 : R_ACCEMPTY0);
        }
        if (p.grid_cnt != nullptr) gemm_tail_reduce(p, et, NJ, red_i);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_cg2(tmem_base, 512);
}

void launch_dst_gemm(const CUtensorMap& tmDS, const CUtensorMap& tmX, const DstParams& p, cudaStream_t st) {
    const int grid = p.sched_pairs * 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(GT); cfg.dynamicSmemBytes = G_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;       // the grid barrier of the tail reduction needs every CTA resident
    attr[0].val.cooperative = p.grid_cnt != nullptr ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (p.transposed) {
        static bool done[64] = {false};
        ensure_smem_attr(dst_gemm_kernel<true>, G_SMEM, done);
        cudaLaunchKernelEx(&cfg, dst_gemm_kernel<true>, tmDS, tmX, p);
    } else {
        static bool done[64] = {false};
        ensure_smem_attr(dst_gemm_kernel<false>, G_SMEM, done);
        cudaLaunchKernelEx(&cfg, dst_gemm_kernel<false>, tmDS, tmX, p);
    }
}

}  // namespace flyp
