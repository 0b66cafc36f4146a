// clip_kernels.cu — the tcgen05 / TMEM / TMA kernels of the ClipLoss hot path (sm_100a only).
//
// Reference semantics being replaced (joliang17/FLYP):
//   clip/loss.py:117-118   logits = logit_scale * image_features @ text_features.T      -> S tiles in TMEM
//   clip/loss.py:208-209   F.cross_entropy(logits, arange, reduction='none') both ways  -> fused exp2/sum epilogue
//   autograd of the above  (softmax - onehot) @ features                                 -> bwd_kernel
// The B x B logit matrix only ever exists as 128 x 128 fp32 tiles in tensor memory.
//
// Warp roles in both kernels (384 threads, one CTA per SM, persistent over a static work list):
//   warp 0   TMA producer (one elected lane)        warp 1   tcgen05.mma issuer (one elected lane)
//   warp 2   TMEM allocator                         warps 4..11  epilogue (TMEM -> registers -> statistics / dS)
#include "clip_kernels.cuh"
#include "sm100.cuh"
#include "bwd_common.cuh"

namespace flyp {
using namespace sm100;

constexpr int NTHREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;

DEVI uint8_t* align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
DEVI void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ===================================================================================================================
// Forward statistics kernel
// ===================================================================================================================
template <bool STAT>
struct FwdCfg {
    static constexpr int STAGE_BYTES = STAT ? CHUNK_BYTES : 2 * CHUNK_BYTES;
    static constexpr int STAGES = STAT ? 5 : 6;
    static constexpr int KC_MAX = 8;
    static constexpr int STAT_BYTES = STAT ? KC_MAX * CHUNK_BYTES : 0;
    static constexpr int RED_BYTES = 4 * TILE * 4;
    static constexpr int ACC_STAGES = 4;
    static constexpr int SMEM_BYTES = STAT_BYTES + STAGES * STAGE_BYTES + RED_BYTES + 256 + 1024;
};

size_t fwd_smem_bytes(bool stationary) {
    return stationary ? FwdCfg<true>::SMEM_BYTES : FwdCfg<false>::SMEM_BYTES;
}

// Work items come from the flat schedule of sched.h (FwdItems): worker `worker` of `nworkers` walks its contiguous range
// of (column unit, row tile) units; a column unit is one block of 128 columns, or for MC a pair of adjacent blocks.
//
// MC: clusters of two CTAs sweep two adjacent column blocks over the SAME stream of A tiles; each CTA fetches half of
// every A chunk and multicasts it to both (halves the L2 -> SM traffic, the measured limiter of the 1-CTA kernel).
// DS (with STAT, MC, not ROBUST): instead of the statistics the epilogue turns every S tile into dS = d(loss)/d(logits)
// (bwd_common.cuh: ds_tile, the same code as the backward sweep's) and writes it, as fp16 scaled by the staging factor,
// to the dS matrix bp.ds_out ([n_m][bp.ds_ld]) - the first of the three kernels of the unfused backward (api.cu): this one
// at the forward's tensor-core efficiency (M = 128 rows per SM, stationary B), then two plain products over dS
// (clip_dst_gemm.cu).  d(scale) = sum dS . <a, b> comes out of the same epilogue loop (one partial per CTA).
constexpr int DS_TRANS_BYTES = 8 * 2048;      // DS: one [16 rows][128 bytes] transposition buffer per epilogue warp

template <bool STAT, bool ROBUST, bool MC, bool DS = false>
__device__ __forceinline__ void
fwd_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const FwdParams& p, const int worker, const int nworkers,
         const BwdParams* bp = nullptr) {
    using Cfg = FwdCfg<STAT>;
    static_assert(!DS || (STAT && MC && !ROBUST), "the dS mode is built on the multicast stationary forward");
    constexpr int RED_BYTES = DS ? DS_TRANS_BYTES : Cfg::RED_BYTES;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int ACC_STAGES = Cfg::ACC_STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint8_t* stat_b = smem;
    uint8_t* ring = smem + Cfg::STAT_BYTES;
    float* red = reinterpret_cast<float*>(ring + STAGES * Cfg::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + RED_BYTES);
    // barrier map: [0,S) full  [S,2S) empty  2S bfull  2S+1 bfree  [2S+2, +ACC) tfull  [.., +ACC) tempty
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t BFULL = bar0 + 8u * (2 * STAGES), BFREE = bar0 + 8u * (2 * STAGES + 1);
    auto TFULL = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    auto TEMPTY = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + ACC_STAGES + s); };
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 + 2 * ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = MC ? (int)cluster_ctarank() : 0;
    // MC: a column unit covers the blocks (2 u, 2 u + 1); the second may lie beyond n_tiles (all-zero B)
    const int n_units = MC ? (p.n_tiles + 1) / 2 : p.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), MC ? 2 : 1); }
        mbar_init(BFULL, 1); mbar_init(BFREE, 1);
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(TFULL(s), 1); mbar_init(TEMPTY(s), EPI_THREADS); }
        fence_mbar_init();
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    }
    if (warp == 2) { tmem_alloc(smem_u32(tmem_holder), 512); tmem_relinquish(); }
    tc_fence_before();
    if (MC) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            FwdItems items(p.m_tiles, n_units, p.n_local, p.nb_rot, nworkers, worker);
            FwdItem fi;
            for (; items.next(fi); ++it) {
                const int nb = MC ? 2 * fi.u + cta : fi.u;
                const int mt0 = fi.mt0, mt1 = fi.mt1;
                peer_wait_rows(p.wait_b, nb * TILE, min((nb + 1) * TILE, p.n_n));
                if (STAT) {
                    mbar_wait(BFREE, (it & 1) ^ 1);
                    mbar_expect_tx(BFULL, p.kc * CHUNK_BYTES);
                    for (int c = 0; c < p.kc; ++c)
                        tma_load_2d(smem_u32(stat_b + c * CHUNK_BYTES), &tmB, BFULL, c * KCHUNK, nb * TILE);
                }
                for (int mt = mt0; mt < mt1; ++mt) {
                    for (int term = 0; term < p.kplan.n_terms; ++term) {
                        const int ca = p.kplan.pa[term] * p.kplan.plane_cols, cb = p.kplan.pb[term] * p.kplan.plane_cols;
                        for (int c = 0; c < p.kc; ++c) {
                            mbar_wait(EMPTY(stage), phase ^ 1);
                            mbar_expect_tx(FULL(stage), Cfg::STAGE_BYTES);
                            uint8_t* dst = ring + stage * Cfg::STAGE_BYTES;
                            if (MC)   // tmA has 64-row boxes: this CTA's half of the tile, delivered to both CTAs
                                tma_load_2d_mc(smem_u32(dst + cta * (CHUNK_BYTES / 2)), &tmA, FULL(stage),
                                               ca + c * KCHUNK, mt * TILE + cta * 64, (uint16_t)3);
                            else
                                tma_load_2d(smem_u32(dst), &tmA, FULL(stage), ca + c * KCHUNK, mt * TILE);
                            if (!STAT)
                                tma_load_2d(smem_u32(dst + CHUNK_BYTES), &tmB, FULL(stage), cb + c * KCHUNK, nb * TILE);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            constexpr uint32_t IDESC = umma_idesc_bf16(TILE, TILE, 0, 0);
            int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0; uint32_t it = 0;
            FwdItems items(p.m_tiles, n_units, p.n_local, p.nb_rot, nworkers, worker);
            FwdItem fi;
            // The issuing thread's own instruction stream bounds a kernel whose MMAs are 64 clocks long: descriptors are a
            // base word plus an integer add per MMA (sm100.cuh), and the poll of the NEXT stage's barrier is issued before
            // the MMAs of the current one so that its latency overlaps their issue.
            constexpr uint32_t HI = umma_desc_hi(1024);
            const uint32_t ring_lo = umma_desc_lo(smem_u32(ring), 16), stat_lo = umma_desc_lo(smem_u32(stat_b), 16);
            const int kc_total = p.kc * p.kplan.n_terms;
            for (; items.next(fi); ++it) {
                const int mt0 = fi.mt0, mt1 = fi.mt1;
                if (STAT) { mbar_wait(BFULL, it & 1); tc_fence_after(); }
                uint32_t ok = mbar_try_wait(FULL(stage), phase);
                for (int mt = mt0; mt < mt1; ++mt) {
                    mbar_wait(TEMPTY(as), aphase ^ 1);
                    const uint32_t d_tmem = tmem_base + as * TILE;
#pragma unroll 1
                    for (int c = 0; c < kc_total; ++c) {
                        if (!ok) mbar_wait(FULL(stage), phase);
                        tc_fence_after();
                        const int cur = stage;
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        // (the next chunk of this item; the last poll of an item is redone by the next item)
                        ok = (c + 1 < kc_total || mt + 1 < mt1) ? mbar_try_wait(FULL(stage), phase) : 0u;
                        const uint32_t a_lo = ring_lo + cur * (Cfg::STAGE_BYTES >> 4);
                        const uint32_t b_lo = STAT ? stat_lo + c * (CHUNK_BYTES >> 4) : a_lo + (CHUNK_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, umma_desc_join(a_lo + k * 2, HI), umma_desc_join(b_lo + k * 2, HI), IDESC,
                                      (c | k) != 0);
                        if (MC) umma_commit_mc(EMPTY(cur), (uint16_t)3); else umma_commit(EMPTY(cur));
                    }
                    umma_commit(TFULL(as));
                    if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
                }
                if (STAT) umma_commit(BFREE);
            }
        }
    } else if (warp >= 4 && DS) {
        // ------------------------------------------------------------------ epilogue, dS mode
        if constexpr (DS) {
            const BwdParams& b = *bp;
            const int q = warp & 3, h = (warp - 4) >> 2;
            const float s = *p.scale;
            const float c1 = s * LOG2E;
            float G, invG;
            staging_scale(b.gmax_bits, G, invG);
            const bool fast = b.fast_info != nullptr && b.fast_info[1] != 0.f;
            const float c0 = fast ? b.fast_info[0] : 0.f;
            const bool want_ds = b.dscale_part != nullptr;
            float dsum = 0.f;
            uint4* const tb = reinterpret_cast<uint4*>(red) + (warp - 4) * 128;      // [16 rows][8 x 16 bytes]
            int as = 0; uint32_t aphase = 0;
            FwdItems items(p.m_tiles, n_units, p.n_local, p.nb_rot, nworkers, worker);
            FwdItem fi;
            while (items.next(fi)) {
                const int nb = 2 * fi.u + cta;
                const bool nb_live = nb < p.n_tiles;
                const int col0 = nb * TILE + h * 64;
                for (int mt = fi.mt0; mt < fi.mt1; ++mt) {
                    mbar_wait(TFULL(as), aphase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * TILE + h * 64;
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(taddr, r0);
                    tmem_ld_32x32b_x32(taddr + 32, r1);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(TEMPTY(as));
                    if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
                    if (!nb_live) continue;                       // (warp-uniform)
                    const int row = mt * TILE + q * 32 + lane;
                    const RowCtx rc = load_row_ctx<true>(b, row, fast, G);
                    float v[64];
                    ds_tile<true, true>(r0, r1, b, rc, col0, c1, fast, c0, G, v, want_ds, dsum);
                    uint32_t pk[32];
                    pack_ds(v, pk);
                    // lane = row holds 128 contiguous bytes of its dS row: through a small transposition buffer every
                    // store instruction writes four full 128-byte row segments (two passes of 16 rows)
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if ((lane >> 4) == half) {
                            const int r = lane & 15;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                tb[r * 8 + (j ^ (r & 7))] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                        }
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int idx = i * 32 + lane, r = idx >> 3, j = idx & 7;
                            const uint4 val = tb[r * 8 + (j ^ (r & 7))];
                            const int grow = mt * TILE + q * 32 + half * 16 + r;
                            if (grow < p.n_m)
                                *reinterpret_cast<uint4*>(b.ds_out + (size_t)grow * b.ds_ld + col0 + j * 8) = val;
                        }
                        __syncwarp();
                    }
                }
            }
            if (want_ds) {
                // one d(scale) partial per CTA (fixed order: warps, then the host-side sum over CTAs)
                float* redf = reinterpret_cast<float*>(red);
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, off);
                epi_bar_sync();                                   // (the transposition buffers are done with)
                if (lane == 0) redf[warp - 4] = dsum;
                epi_bar_sync();
                if (threadIdx.x == 128) {
                    float tot = 0.f;
                    for (int k = 0; k < 8; ++k) tot += redf[k];
                    b.dscale_part[worker * 2 + cta] = tot * invG;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3;          // TMEM lane quarter this warp may access
        const int h = (warp - 4) >> 2;   // column half handled by this warp
        const float s = *p.scale;
        const float c1 = s * LOG2E;
        const float c0 = fixed_shift(c1, p.shift_slack);
        int as = 0; uint32_t aphase = 0;
        FwdItems items(p.m_tiles, n_units, p.n_local, p.nb_rot, nworkers, worker);
        FwdItem fi;
        while (items.next(fi)) {
            const int nb = MC ? 2 * fi.u + cta : fi.u;
            const bool nb_live = nb < p.n_tiles;           // MC: the odd block of the last pair may not exist
            const int mt0 = fi.mt0, mt1 = fi.mt1;
            const int col0 = nb * TILE + h * 64;
            const bool n_edge = (nb == p.n_tiles - 1) && (p.n_n % TILE != 0);
            float ca[64];
#pragma unroll
            for (int k = 0; k < 64; ++k) ca[k] = 0.f;

            for (int mt = mt0; mt < mt1; ++mt) {
                mbar_wait(TFULL(as), aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * TILE + h * 64;
                const int row = mt * TILE + q * 32 + lane;
                const bool rowvalid = row < p.n_m;
                const bool edge = n_edge || ((mt == p.m_tiles - 1) && (p.n_m % TILE != 0));
                float rs = 0.f, rmax = -INFINITY;
                // rel: index (within this thread's 64 columns) of the row's positive, excluded from the sums
                int rel = -1;
                if (p.pos != nullptr) rel = p.pos[row] - col0;
                else if (p.pos_arith) {
                    const int pc = row + p.pos_off;
                    if (rowvalid && pc >= 0 && pc < p.n_n) rel = pc - col0;
                }
                const bool has_pos = __any_sync(0xffffffffu, rel >= 0 && rel < 64);
                if (!ROBUST) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(taddr + cc * 32, r);
                        tmem_ld_wait();
                        if (p.dbg_logits != nullptr && rowvalid) {
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                int col = col0 + cc * 32 + k;
                                if (col < p.n_n) p.dbg_logits[(size_t)row * p.n_n + col] = __uint_as_float(r[k]);
                            }
                        }
                        if (!edge && !has_pos) {
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                float e = ex2f(fmaf(__uint_as_float(r[k]), c1, -c0));
                                rs += e;
                                ca[cc * 32 + k] += e;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                float e = ex2f(fmaf(__uint_as_float(r[k]), c1, -c0));
                                const bool keep = rowvalid && (col0 + cc * 32 + k) < p.n_n && (cc * 32 + k) != rel;
                                e = keep ? e : 0.f;
                                rs += e;
                                ca[cc * 32 + k] += e;
                            }
                        }
                    }
                } else {
                    // robust: exact per-tile (max, sum) pair for this row over this thread's 64 columns
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(taddr, r0);
                    tmem_ld_32x32b_x32(taddr + 32, r1);
                    tmem_ld_wait();
                    float x[64];
                    const float cx = (p.argidx != nullptr) ? 1.f : c1;      // argmax mode works on the raw dot products
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        x[k] = __uint_as_float(r0[k]) * cx;
                        x[32 + k] = __uint_as_float(r1[k]) * cx;
                    }
                    if (edge || has_pos) {
#pragma unroll
                        for (int k = 0; k < 64; ++k)
                            if (!rowvalid || (col0 + k) >= p.n_n || k == rel) x[k] = -INFINITY;
                    }
                    float acc_r = 0.f;
                    if (p.mask_mode != 0) {
                        // label-aware variants: same-class entries (cls_n padded to ld_cols with -1, never equal)
                        const int crow = rowvalid ? p.cls_m[row] : -2;
                        const float lrow = (p.mask_mode == 2 && rowvalid) ? p.acc_lse[row] : 0.f;
                        float acc_x = 0.f, acc_l = 0.f;
#pragma unroll
                        for (int k4 = 0; k4 < 16; ++k4) {
                            const int4 c4 = __ldg(reinterpret_cast<const int4*>(p.cls_n + col0) + k4);
                            const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int k = k4 * 4 + j;
                                const bool same = cc[j] == crow && x[k] != -INFINITY;      // (-inf: padding / the positive)
                                if (p.mask_mode == 1) {
                                    if (same) x[k] = -INFINITY;
                                } else if (same) {
                                    const float pr = ex2f(x[k] - lrow);
                                    const float om = 1.f - pr;
                                    acc_x += x[k];
                                    acc_l += __log2f(om);
                                    acc_r += __fdividef(pr, om);
                                }
                            }
                        }
                        if (p.mask_mode == 2) { rs = acc_x; rmax = acc_l; }
                    }
                    if (p.mask_mode == 2) {
                        if (nb_live) p.acc3[(size_t)(nb * 2 + h) * p.ld_rows + row] = acc_r;
                    } else {
#pragma unroll
                        for (int k = 0; k < 64; ++k) rmax = fmaxf(rmax, x[k]);
                    }
                    if (p.mask_mode == 2) {
                        // (sums already in rs / rmax)
                    } else if (p.argidx != nullptr) {
                        // fused argmax: first column (lowest index) that attains the maximum of this thread's 64 columns
                        int bi = 63;
#pragma unroll
                        for (int k = 62; k >= 0; --k) bi = (x[k] == rmax) ? k : bi;
                        if (nb_live) p.argidx[(size_t)(nb * 2 + h) * p.ld_rows + row] = col0 + bi;
                    } else {
                        const float mref = (rmax == -INFINITY) ? 0.f : rmax;
#pragma unroll
                        for (int k = 0; k < 64; ++k) rs += ex2f(x[k] - mref);
                    }
                }
                tc_fence_before();
                mbar_arrive(TEMPTY(as));
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
                if (nb_live) {
                    if (!ROBUST || p.argidx == nullptr) p.rowpart[(size_t)(nb * 2 + h) * p.ld_rows + row] = rs;
                    if (ROBUST) p.rowmax[(size_t)(nb * 2 + h) * p.ld_rows + row] = rmax;
                }
            }

            if (!ROBUST) {
                // column sums: butterfly transpose-reduce over the 32 lanes -> lane l owns columns 2l, 2l+1
#pragma unroll
                for (int off = 16, n = 64; off >= 1; off >>= 1, n >>= 1) {
                    const int half = n >> 1;
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int k = 0; k < half; ++k) {
                        float send = up ? ca[k] : ca[k + half];
                        float keep = up ? ca[k + half] : ca[k];
                        ca[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                red[q * TILE + h * 64 + 2 * lane] = ca[0];
                red[q * TILE + h * 64 + 2 * lane + 1] = ca[1];
                epi_bar_sync();
                const int et = threadIdx.x - 128;
                if (et < TILE && nb_live) {
                    float v = (red[et] + red[TILE + et]) + (red[2 * TILE + et] + red[3 * TILE + et]);
                    p.colpart[(size_t)fi.slot * p.ld_cols + nb * TILE + et] = v;
                }
                epi_bar_sync();
            }
        }
    }
    tc_fence_before();
    if (MC) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <bool STAT, bool ROBUST>
__global__ void __launch_bounds__(NTHREADS, 1)
fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwdParams p,
           const int* __restrict__ gate) {
    if (ROBUST && gate != nullptr && *gate == 0) return;  // fast path was adequate: nothing to do (grid-uniform)
    fwd_body<STAT, ROBUST, false>(tmA, tmB, p, (int)blockIdx.x, (int)gridDim.x);
}

// Both robust passes of the symmetric loss in one gated launch: CTAs [0, split) sweep the rows of S, the others the rows
// of S^T (operand roles swapped).
template <bool STAT>
__global__ void __launch_bounds__(NTHREADS, 1)
fwd_kernel_robust2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwdParams pr,
                   const FwdParams pc, const int* __restrict__ gate, const int split) {
    if (gate != nullptr && *gate == 0) return;
    if ((int)blockIdx.x < split) fwd_body<STAT, true, false>(tmA, tmB, pr, (int)blockIdx.x, split);
    else fwd_body<STAT, true, false>(tmB, tmA, pc, (int)blockIdx.x - split, (int)gridDim.x - split);
}

// tmA64: the A operand with [64 rows][64 cols] boxes
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
fwd_kernel_mc(const __grid_constant__ CUtensorMap tmA64, const __grid_constant__ CUtensorMap tmB, const FwdParams p) {
    fwd_body<true, false, true>(tmA64, tmB, p, (int)(blockIdx.x >> 1), (int)(gridDim.x >> 1));
}

// The dS kernel of the unfused backward: same schedule, producer and MMA issuer as fwd_kernel_mc.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
ds_kernel_mc(const __grid_constant__ CUtensorMap tmA64, const __grid_constant__ CUtensorMap tmB, const FwdParams p,
             const __grid_constant__ BwdParams bp) {
    fwd_body<true, false, true, true>(tmA64, tmB, p, (int)(blockIdx.x >> 1), (int)(gridDim.x >> 1), &bp);
}

int fwd_workers(int m_tiles, int n_tiles, bool mc, int num_sms) {
    return mc ? fwd_sched_workers(m_tiles, (n_tiles + 1) / 2, num_sms / 2) : fwd_sched_workers(m_tiles, n_tiles, num_sms);
}

void launch_fwd_mc(const CUtensorMap& tmA64, const CUtensorMap& tmB, const FwdParams& p, int num_sms, cudaStream_t st) {
    const int workers = fwd_workers(p.m_tiles, p.n_tiles, true, num_sms);
    const size_t smem = fwd_smem_bytes(true);
    static bool attr_done[64] = {false};
    ensure_smem_attr(fwd_kernel_mc, smem, attr_done);
    fwd_kernel_mc<<<workers * 2, NTHREADS, smem, st>>>(tmA64, tmB, p);
}

int launch_ds_mc(const CUtensorMap& tmA64, const CUtensorMap& tmB, const FwdParams& p, const BwdParams& bp, int num_sms,
                 cudaStream_t st) {
    const int workers = fwd_workers(p.m_tiles, p.n_tiles, true, num_sms);
    const size_t smem = fwd_smem_bytes(true) - FwdCfg<true>::RED_BYTES + DS_TRANS_BYTES;
    static bool attr_done[64] = {false};
    ensure_smem_attr(ds_kernel_mc, smem, attr_done);
    ds_kernel_mc<<<workers * 2, NTHREADS, smem, st>>>(tmA64, tmB, p, bp);
    return workers * 2;          // CTAs = d(scale) partials written
}

static bool fwd_stationary(const FwdParams& p) { return p.kplan.n_terms == 1 && p.kc <= FwdCfg<true>::KC_MAX; }

void launch_fwd(const CUtensorMap& tmA, const CUtensorMap& tmB, const FwdParams& p, bool robust, const int* gate,
                int num_sms, cudaStream_t st) {
    const bool stat = fwd_stationary(p);
    const int grid = fwd_workers(p.m_tiles, p.n_tiles, false, num_sms);
    const size_t smem = fwd_smem_bytes(stat);
#define FLYP_LAUNCH_FWD(S, R)                                                                                  \
    do {                                                                                                       \
        static bool attr_done[64] = {false};                                                                   \
        ensure_smem_attr(fwd_kernel<S, R>, smem, attr_done);                                                   \
        fwd_kernel<S, R><<<grid, NTHREADS, smem, st>>>(tmA, tmB, p, gate);                                     \
    } while (0)
    if (stat) { if (robust) FLYP_LAUNCH_FWD(true, true); else FLYP_LAUNCH_FWD(true, false); }
    else      { if (robust) FLYP_LAUNCH_FWD(false, true); else FLYP_LAUNCH_FWD(false, false); }
#undef FLYP_LAUNCH_FWD
}

void launch_fwd_robust2(const CUtensorMap& tmA, const CUtensorMap& tmB, const FwdParams& pr, const FwdParams& pc,
                        const int* gate, int num_sms, cudaStream_t st) {
    const bool stat = fwd_stationary(pr);
    // each pass gets half of the SMs (the robust path is the rare one; simplicity over balance)
    const int half = num_sms / 2 > 0 ? num_sms / 2 : 1;
    const int g_r = fwd_workers(pr.m_tiles, pr.n_tiles, false, half);
    const int g_c = fwd_workers(pc.m_tiles, pc.n_tiles, false, half);
    const size_t smem = fwd_smem_bytes(stat);
    if (stat) {
        static bool attr_done[64] = {false};
        ensure_smem_attr(fwd_kernel_robust2<true>, smem, attr_done);
        fwd_kernel_robust2<true><<<g_r + g_c, NTHREADS, smem, st>>>(tmA, tmB, pr, pc, gate, g_r);
    } else {
        static bool attr_done[64] = {false};
        ensure_smem_attr(fwd_kernel_robust2<false>, smem, attr_done);
        fwd_kernel_robust2<false><<<g_r + g_c, NTHREADS, smem, st>>>(tmA, tmB, pr, pc, gate, g_r);
    }
}

// ===================================================================================================================
// Backward sweep kernel: gradient w.r.t. the M-side operand, 256 output columns per work item.
// ===================================================================================================================
struct BwdCfg {
    static constexpr int STAGE_BYTES = 2 * CHUNK_BYTES;   // A chunk + B chunk of the S contraction
    static constexpr int STAGES = 4;
    static constexpr int TB_BYTES = 4 * CHUNK_BYTES;      // B operand of the dA MMA: [128 n][256 d] as 4 MN-major boxes
    static constexpr int DS_BYTES = 2 * CHUNK_BYTES;      // dS tile [128 m][128 n] fp16, K-major SW128 (one plane)
    static constexpr int RED_BYTES = 64;
    // bf16 features: 4 ring stages + 1 dS plane; fp32 features (f32_mode): 3 ring stages + 2 dS planes (hi, lo)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + TB_BYTES + DS_BYTES + RED_BYTES + 256 + 1024;
    static constexpr int DPART = 256;
};
size_t bwd_smem_bytes() { return BwdCfg::SMEM_BYTES; }

template <bool ROW_TERM, bool COL_TERM, bool MASK>
__global__ void __launch_bounds__(NTHREADS, 1)
bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           const __grid_constant__ CUtensorMap tmBd, const BwdParams p) {
    using Cfg = BwdCfg;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const bool f32m = p.f32_mode != 0;
    const int nstages = f32m ? STAGES - 1 : STAGES;       // one ring stage is traded for the second dS plane
    const int npass = f32m ? 3 : 1;                        // dA = dS_hi.B_hi + dS_hi.B_lo + dS_lo.B_hi
    uint8_t* ring = smem;
    uint8_t* tb = ring + nstages * Cfg::STAGE_BYTES;
    uint8_t* ds = tb + Cfg::TB_BYTES;
    float* red = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::TB_BYTES + Cfg::DS_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + Cfg::RED_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto SFULL = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto SEMPTY = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    const uint32_t TBFULL = bar0 + 8u * (2 * STAGES + 4), TBEMPTY = bar0 + 8u * (2 * STAGES + 5);
    const uint32_t DSFULL = bar0 + 8u * (2 * STAGES + 6), DSEMPTY = bar0 + 8u * (2 * STAGES + 7);
    const uint32_t ACCFULL = bar0 + 8u * (2 * STAGES + 8), ACCEMPTY = bar0 + 8u * (2 * STAGES + 9);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = p.m_tiles * p.d_parts;
    const int NT = p.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(SFULL(s), 1); mbar_init(SEMPTY(s), EPI_THREADS); }
        mbar_init(TBFULL, 1); mbar_init(TBEMPTY, 1);
        mbar_init(DSFULL, EPI_THREADS); mbar_init(DSEMPTY, 1);
        mbar_init(ACCFULL, 1); mbar_init(ACCEMPTY, EPI_THREADS);
        fence_mbar_init();
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmBd);
    }
    if (warp == 2) { tmem_alloc(smem_u32(tmem_holder), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t TM_ACC = tmem_base, TM_S = tmem_base + Cfg::DPART;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0; uint32_t g = 0;  // g: running TB-load counter (TB buffer phase)
            auto load_s = [&](int mb, int n) {
                for (int term = 0; term < p.kplan.n_terms; ++term) {
                    const int ca = p.kplan.pa[term] * p.kplan.plane_cols, cb = p.kplan.pb[term] * p.kplan.plane_cols;
                    for (int c = 0; c < p.kc; ++c) {
                        mbar_wait(EMPTY(stage), phase ^ 1);
                        mbar_expect_tx(FULL(stage), Cfg::STAGE_BYTES);
                        uint8_t* dst = ring + stage * Cfg::STAGE_BYTES;
                        tma_load_2d(smem_u32(dst), &tmA, FULL(stage), ca + c * KCHUNK, mb * TILE);
                        tma_load_2d(smem_u32(dst + CHUNK_BYTES), &tmB, FULL(stage), cb + c * KCHUNK, n * TILE);
                        if (++stage == nstages) { stage = 0; phase ^= 1; }
                    }
                }
            };
            auto load_tb = [&](int dp, int n) {
                for (int pass = 0; pass < npass; ++pass) {
                    const int pcol = (pass == 1) ? p.bd_plane_cols : 0;      // feature plane: hi, lo, hi
                    mbar_wait(TBEMPTY, (g & 1) ^ 1);
                    mbar_expect_tx(TBFULL, Cfg::TB_BYTES);
                    for (int j = 0; j < 4; ++j)
                        tma_load_2d(smem_u32(tb + j * CHUNK_BYTES), &tmBd, TBFULL,
                                    pcol + dp * Cfg::DPART + j * KCHUNK, n * TILE);
                    ++g;
                }
            };
            peer_wait_all(p.wait_b);      // multi-GPU: the N-side rows (and their fp16 copy) of every rank have arrived
            peer_wait_all(p.wait_bd);
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int mb = item / p.d_parts, dp = item % p.d_parts;
                // same order as the MMA issuer consumes: S(0), S(1), dA(0), S(2), dA(1), ...
                load_s(mb, 0);
                for (int t = 1; t < NT; ++t) { load_s(mb, t); load_tb(dp, t - 1); }
                load_tb(dp, NT - 1);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            constexpr uint32_t IDESC_S = umma_idesc_bf16(TILE, TILE, 0, 0);
            // A = dS staged as scaled fp16, B = fp16 copy of the features (MN-major): a_format = b_format = 0 (F16)
            constexpr uint32_t IDESC_D = umma_idesc_bf16(TILE, Cfg::DPART, 0, 1) & ~((7u << 7) | (7u << 10));
            int stage = 0; uint32_t phase = 0; uint32_t gs = 0, gd = 0, gtb = 0, it = 0;
            const int kc_total = p.kc * p.kplan.n_terms;
            // (lean issue path as in the forward: descriptor base words + integer adds, next stage polled ahead)
            constexpr uint32_t HI = umma_desc_hi(1024);
            const uint32_t ring_lo = umma_desc_lo(smem_u32(ring), 16);
            auto mma_s = [&]() {
                const int sb = gs & 1;
                mbar_wait(SEMPTY(sb), ((gs >> 1) & 1) ^ 1);
                const uint32_t d_tmem = TM_S + sb * TILE;
                uint32_t ok = mbar_try_wait(FULL(stage), phase);
#pragma unroll 1
                for (int c = 0; c < kc_total; ++c) {
                    if (!ok) mbar_wait(FULL(stage), phase);
                    tc_fence_after();
                    const int cur = stage;
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                    if (c + 1 < kc_total) ok = mbar_try_wait(FULL(stage), phase);
                    const uint32_t a_lo = ring_lo + cur * (Cfg::STAGE_BYTES >> 4), b_lo = a_lo + (CHUNK_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, umma_desc_join(a_lo + k * 2, HI), umma_desc_join(b_lo + k * 2, HI), IDESC_S,
                                  (c | k) != 0);
                    umma_commit(EMPTY(cur));
                }
                umma_commit(SFULL(sb));
                ++gs;
            };
            auto mma_d = [&](bool first) {
                if (first) { mbar_wait(ACCEMPTY, (it & 1) ^ 1); }
                mbar_wait(DSFULL, gd & 1);
                const uint32_t tb_addr = smem_u32(tb);
                for (int pass = 0; pass < npass; ++pass) {
                    mbar_wait(TBFULL, gtb & 1);
                    tc_fence_after();
                    const uint32_t ds_addr = smem_u32(ds) + (pass == 2 ? Cfg::DS_BYTES : 0);   // dS plane: hi, hi, lo
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        // A: dS rows m, K = n (K-major, two 64-wide chunks). B: [K = n rows][N = d], MN-major boxes.
                        const uint64_t ad = umma_desc_sw128(ds_addr + (kk >> 2) * CHUNK_BYTES + (kk & 3) * 32, 16, 1024);
                        const uint64_t bd = umma_desc_sw128(tb_addr + kk * 16 * 128, CHUNK_BYTES, 1024);
                        umma_bf16(TM_ACC, ad, bd, IDESC_D, (!first) || (pass | kk) != 0);
                    }
                    umma_commit(TBEMPTY);
                    ++gtb;
                }
                umma_commit(DSEMPTY);
                ++gd;
            };
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                mma_s();
                for (int t = 1; t < NT; ++t) { mma_s(); mma_d(t == 1); }
                mma_d(NT == 1);
                umma_commit(ACCFULL);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;
        const float s = *p.scale;
        const float c1 = s * LOG2E;
        float G, invG;
        staging_scale(p.gmax_bits, G, invG);
        const bool fast = !MASK && p.fast_info != nullptr && p.fast_info[1] != 0.f;    // (masks use the two-exponential form)
        const float c0 = fast ? p.fast_info[0] : 0.f;
        uint32_t gs = 0, it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int mb = item / p.d_parts, dp = item % p.d_parts;
            const int rloc = q * 32 + lane;
            const int m = mb * TILE + rloc;
            const bool rowvalid = m < p.n_m;
            const RowCtx rc = load_row_ctx<ROW_TERM>(p, m, fast, G);
            // the S tile is recomputed once per 256-column output part: only part 0 contributes to d(scale)
            const bool want_ds = p.dscale_part != nullptr && dp == 0;
            float dsum = 0.f;
            for (int t = 0; t < NT; ++t, ++gs) {
                const int sb = gs & 1;
                mbar_wait(SFULL(sb), (gs >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = TM_S + ((uint32_t)(q * 32) << 16) + sb * TILE + h * 64;
                uint32_t r0[32], r1[32];
                tmem_ld_32x32b_x32(taddr, r0);
                tmem_ld_32x32b_x32(taddr + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(SEMPTY(sb));
                const int n0 = t * TILE + h * 64;
                float v[64];
                ds_tile<ROW_TERM, COL_TERM, MASK>(r0, r1, p, rc, n0, c1, fast, c0, G, v, want_ds, dsum);
                uint32_t pk[32];
                pack_ds(v, pk);
                // dS buffer is free once the dA MMA of the previous tile has completed
                mbar_wait(DSEMPTY, (gs & 1) ^ 1);
                store_ds_row(ds + h * CHUNK_BYTES + rloc * 128, rloc, pk);
                if (f32m) {
                    // second plane: what the fp16 rounding of the first one lost
                    residual_ds(v, pk);
                    pack_ds(v, pk);
                    store_ds_row(ds + Cfg::DS_BYTES + h * CHUNK_BYTES + rloc * 128, rloc, pk);
                }
                fence_proxy_async_smem();
                mbar_arrive(DSFULL);
            }
            // -------- item done: drain the accumulator (this thread: row m, 128 of the 256 columns)
            mbar_wait(ACCFULL, it & 1);
            tc_fence_after();
            const float omul = s * p.out_mul * invG;
            const int dbase = dp * Cfg::DPART + h * 128;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(TM_ACC + ((uint32_t)(q * 32) << 16) + h * 128 + cc * 32, r);
                tmem_ld_wait();
                const int d0 = dbase + cc * 32;
                if (rowvalid) {
                    if (p.out_fp32) {
                        float* orow = reinterpret_cast<float*>(p.out) + (size_t)m * p.ld_out;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (d0 + j * 4 < p.d_out) {
                                float4 o = make_float4(__uint_as_float(r[4 * j]) * omul, __uint_as_float(r[4 * j + 1]) * omul,
                                                       __uint_as_float(r[4 * j + 2]) * omul, __uint_as_float(r[4 * j + 3]) * omul);
                                *reinterpret_cast<float4*>(orow + d0 + j * 4) = o;
                            }
                        }
                    } else {
                        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.ld_out;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (d0 + j * 8 < p.d_out) {
                                uint4 o;
                                o.x = pack_bf16x2(__uint_as_float(r[8 * j]) * omul, __uint_as_float(r[8 * j + 1]) * omul);
                                o.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * omul, __uint_as_float(r[8 * j + 3]) * omul);
                                o.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * omul, __uint_as_float(r[8 * j + 5]) * omul);
                                o.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * omul, __uint_as_float(r[8 * j + 7]) * omul);
                                *reinterpret_cast<uint4*>(orow + d0 + j * 8) = o;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(ACCEMPTY);
            if (p.dscale_part != nullptr) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, off);
                if (lane == 0) red[warp - 4] = dsum;
                epi_bar_sync();
                if (et == 0) {
                    float tot = 0.f;
                    for (int w = 0; w < 8; ++w) tot += red[w];
                    p.dscale_part[item] = tot * invG;
                }
                epi_bar_sync();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

void launch_bwd(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBd, const BwdParams& p,
                int num_sms, cudaStream_t st) {
    const int n_items = p.m_tiles * p.d_parts;
    const int grid = n_items < num_sms ? n_items : num_sms;
    const size_t smem = bwd_smem_bytes();
    const bool row_term = p.wr != nullptr, col_term = p.wc != nullptr;
#define FLYP_LAUNCH_BWD(R, C, M)                                                                               \
    do {                                                                                                       \
        static bool attr_done[64] = {false};                                                                   \
        ensure_smem_attr(bwd_kernel<R, C, M>, smem, attr_done);                                                \
        bwd_kernel<R, C, M><<<grid, NTHREADS, smem, st>>>(tmA, tmB, tmBd, p);                                  \
    } while (0)
    if (p.mask_mode != 0) FLYP_LAUNCH_BWD(true, true, true);         // label-aware variants always carry both terms
    else if (row_term && col_term) FLYP_LAUNCH_BWD(true, true, false);
    else if (row_term) FLYP_LAUNCH_BWD(true, false, false);
    else FLYP_LAUNCH_BWD(false, true, false);
#undef FLYP_LAUNCH_BWD
}

}  // namespace flyp
