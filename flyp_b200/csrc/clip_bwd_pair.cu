// clip_bwd_pair.cu — backward sweep on CTA pairs (tcgen05 cta_group::2), feature dim a multiple of 128 up to 1024.
//
// dim > 512: the dA^T accumulators of a 128-row block (128 x dim fp32) exceed the pair's tensor memory, so a row block is
// swept once per HALF of the output columns ("virtual row blocks" (row block, d-half), p.n_dh = 2, p.d_half columns each),
// S being recomputed per pass (6 B^2 D executed per sweep instead of the 10 B^2 D of the single-CTA kernel at D = 1024);
// and only the first 8 of the kc K-chunks of the A half-block are stationary - the others are streamed through the ring,
// two 8 KiB chunks per slot, right before the B chunks they multiply.
//
// One cluster of two CTAs (an SM pair) owns a block of 128 M-rows and walks over the N side in steps of 256 columns.
// Per step
//   (1) S = A_blk . B_step^T as ONE pair MMA (M = 128: each CTA supplies 64 rows of A; N = 256: each CTA supplies
//       128 rows of B).  Each CTA receives the logits of ITS 64 rows for all 256 columns
//       (TMEM lanes 0..63 = columns 0..127, lanes 64..127 = columns 128..255).
//   (2) the epilogue warps of each CTA turn their S slice into dS (softmax weights, exact value at the positive),
//       stage it as scaled fp16 in shared memory as [64 rows][256 cols] K-major.
//   (3) dA^T[d, m] += sum_n B16[n, d] * dS[m, n] as pair MMAs with M = 256 feature columns (128 per CTA), N = 128 rows
//       (the B operand of a pair MMA is split across the two CTAs by halves of N, i.e. exactly the 64-row dS slices
//       the two CTAs produced - the tensor core shares them, no software exchange), K = 256.
// Every logit tile is therefore recomputed once per sweep (not once per 256 output columns as in bwd_kernel) and both
// CTAs run at full pair-MMA rate.  The accumulators (2 x 128 columns) live in TMEM for the whole row block together
// with two S stages (2 x 128 columns).
//
// Shared memory per CTA: stationary A half-block (64 rows x D, <= 64 KiB), one ring of eight 16 KiB slots that carries,
// in consumption order, the B chunks of the S product and the fp16 feature boxes of the dA^T product, and the dS tile.
#include "clip_kernels.cuh"
#include "sm100.cuh"
#include "bwd_common.cuh"
#include "sched.h"
#include <cstdlib>
#include <cstring>

namespace flyp {
using namespace sm100;

namespace {
constexpr int NTHREADS2 = 384;
constexpr int EPI_ALL = 512;          // epilogue threads of both CTAs
constexpr float LOG2E2 = 1.4426950408889634f;

struct PairCfg {
    static constexpr int SLOT = 16384;
    // Ring slots x dS tiles: 8 x 1 or 6 x 2 fit (the code below handles both).  Measured at B = 32768, D = 512: two tiles
    // (the epilogue writes tile t + 1 while the MMAs and the spill still read tile t) do not pay for the two ring slots
    // they cost - 1.61 / 1.96 ms per sweep without / with the dS spill against 1.52 / 1.80 ms.
    static constexpr int NSLOT = 8;                    // (even: the dA^T operand boxes are consumed as (even, odd) pairs)
    static constexpr int IST_BYTES = 65536;            // 8 chunks [64 rows][64] bf16
    static constexpr int DS_TILE = 32768;              // one dS tile: 4 chunks [64 rows][64] fp16
    static constexpr int DS_TILES = 1;
    static constexpr int DS_BYTES = DS_TILES * DS_TILE;
    static constexpr int RED_BYTES = 512;              // 8 d(scale) partials, a flag, the slot list of a split block
    static constexpr int SMEM_BYTES = IST_BYTES + NSLOT * SLOT + DS_BYTES + RED_BYTES + 256 + 1024;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
    static constexpr int NSTEP = 256;                  // N columns per step
};

DEVI uint8_t* align1024p(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
DEVI void epi_bar_sync2() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

}  // namespace

// Schedule of a sweep over m_tiles row blocks on P CTA pairs: floor(m_tiles / P) rounds of whole row blocks (all pairs
// walk the columns in lockstep: one stream of B tiles through L2 serves them all), then the remaining blocks as a FLAT
// tail - their (row block, column step) units cut into P equal contiguous ranges - so that the last, partial wave keeps
// every pair busy (a per-row-block schedule idles 14 % of the pairs when a rank owns 32 or 64 row blocks, i.e.
// B = 32768 on 8 or 4 GPUs).
int bwd_pair_sched_pairs(int m_tiles, int n_cols, int num_sms) {
    int npairs = num_sms / 2;
    static const int forced = [] { const char* e = getenv("FLYP_SCHED_PAIRS"); return e ? atoi(e) : 0; }();   // A/B switch
    if (forced > 0 && forced < npairs) npairs = forced;
    const int NJ = (n_cols + PairCfg::NSTEP - 1) / PairCfg::NSTEP;
    const long long S = (long long)m_tiles * NJ;
    // A range shorter than MIN_STEPS column steps does not pay for its share of the end-of-sweep reduction (fp32
    // partials, grid barrier): small problems use fewer pairs - down to one per row block, i.e. no partials at all
    // (B = 512: 4 pairs x 2 steps instead of 8 x 1 - measured 30 us -> 12 us per sweep).
    constexpr int MIN_STEPS = 4;
    long long cap = S / MIN_STEPS;
    if (cap < m_tiles) cap = m_tiles;
    if (cap < npairs) npairs = (int)cap;
    return (int)(S < npairs ? S : npairs);
}

size_t bwd_pair_smem_bytes() { return PairCfg::SMEM_BYTES; }

// ---- end-of-sweep reductions, by the whole grid ----------------------------------------------------------------------
// The flat tail of the schedule leaves fp32 partial accumulators of the split row blocks in part_out, and every CTA a
// partial of d(scale).  Rather than a second kernel (launch gap) or the last arriver of each block (one CTA per block,
// latency-bound: measured +35 us per sweep at 32 blocks), ALL CTAs meet at a grid barrier - the launch is cooperative, so
// every CTA of the grid is resident - and then share the work: one [128 rows x 128 columns] output chunk of a split
// block at a time, its partial slots summed in pair order (deterministic), 8 independent 16-byte loads in flight per
// thread.  Called by the 256 epilogue threads of every CTA; out of line so that it costs the main loop no registers.
__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __noinline__ void sweep_tail_reduce(const BwdParams& p, const int et, const int NJ, int* red_i) {
    // ---- grid barrier (counter zeroed by the host before the launch)
    __threadfence();
    epi_bar_sync2();
    if (et == 0) {
        atomicAdd(p.grid_cnt, 1);
        while (ld_acquire_gpu_s32(p.grid_cnt) < (int)gridDim.x) __nanosleep(40);
    }
    epi_bar_sync2();
    __threadfence();
    const int v_tiles = p.m_tiles * p.n_dh, P = p.sched_pairs;
    const int first = (v_tiles / P) * P, n_tail = v_tiles - first;
    const int cpb = p.d_half / 128;                         // 128-column chunks per (virtual) row block
    for (int c = blockIdx.x; c < n_tail * cpb; c += gridDim.x) {
        const int tb = c / cpb, dl0 = (c - tb * cpb) * 128;
        if (et == 0) {
            TailParts parts;
            int np = 0;
            if (parts.init(v_tiles, NJ, P, tb))
                for (int sl = parts.next(); sl >= 0; sl = parts.next()) red_i[18 + np++] = sl;
            red_i[17] = np;                                  // 0: the block was swept whole and written directly
        }
        epi_bar_sync2();
        const int np = red_i[17];
        const int vb = first + tb, mb = vb / p.n_dh, dh = vb - mb * p.n_dh;
        if (np > 0 && dh * p.d_half + dl0 < p.d_out) {
#pragma unroll 1
            for (int f0 = et; f0 < TILE * 32; f0 += 256 * 8) {
                float4 acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < np; ++k) {
                    const float* base = p.part_out + (size_t)red_i[18 + k] * TILE * p.d_half + dl0;
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int f = f0 + j * 256, ri = f >> 5;
                        v[j] = (mb * TILE + ri < p.n_m)
                                   ? __ldcg(reinterpret_cast<const float4*>(base + (size_t)ri * p.d_half + (f & 31) * 4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { acc[j].x += v[j].x; acc[j].y += v[j].y; acc[j].z += v[j].z; acc[j].w += v[j].w; }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int f = f0 + j * 256, ri = f >> 5, mi = mb * TILE + ri;
                    const int d = dh * p.d_half + dl0 + (f & 31) * 4;
                    if (mi >= p.n_m) continue;
                    if (p.out_fp32) {
                        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)mi * p.ld_out + d) = acc[j];
                    } else {
                        uint2 u;
                        u.x = pack_bf16x2(acc[j].x, acc[j].y); u.y = pack_bf16x2(acc[j].z, acc[j].w);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)mi * p.ld_out + d) = u;
                    }
                }
            }
        }
        epi_bar_sync2();          // red_i is rewritten for the next chunk
    }
    // ---- d(scale): CTA 0 sums the partials of the whole grid in a fixed order and publishes the total
    if (p.dscale_out != nullptr && blockIdx.x == 0 && et < 32) {
        float acc = 0.f;
        for (int i = et; i < (int)gridDim.x; i += 32) acc += __ldcg(p.dscale_part + i);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (et == 0) p.dscale_out[0] = acc;
        peer_push_value_warp(p.ds_push, acc, et);
    }
}

// WIDE: feature dim > 512 (streamed A chunks, two passes over the output columns); the narrow instantiation compiles
// that logic away.
// PROF: per-role wait-cycle counters (tools/pair_prof.py); a template parameter so that the MMA issuer's hot loop carries
// no profiling branches otherwise.
template <bool ROW_TERM, bool COL_TERM, bool WIDE, bool PROF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS2, 1)
bwd_pair_kernel(const __grid_constant__ CUtensorMap tmA64, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmBd, const __grid_constant__ CUtensorMap tmDS,
                const __grid_constant__ BwdParams p) {
    using Cfg = PairCfg;
    constexpr int NSLOT = Cfg::NSLOT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024p(smem_raw);
    uint8_t* ist = smem;
    uint8_t* ring = ist + Cfg::IST_BYTES;
    uint8_t* ds = ring + NSLOT * Cfg::SLOT;
    float* red = reinterpret_cast<float*>(ds + Cfg::DS_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + Cfg::RED_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NSLOT + s); };
    auto SFULL = [&](int s) { return bar0 + 8u * (2 * NSLOT + s); };
    auto SEMPTY = [&](int s) { return bar0 + 8u * (2 * NSLOT + 2 + s); };
    auto DSFULL = [&](int b) { return bar0 + 8u * (2 * NSLOT + 4 + b); };
    auto DSEMPTY = [&](int b) { return bar0 + 8u * (2 * NSLOT + 6 + b); };
    auto DSLOC = [&](int b) { return bar0 + 8u * (2 * NSLOT + 8 + b); };   // keep_ds: this CTA's 256 epilogue threads wrote tile b
    const uint32_t ACCFULL = bar0 + 8u * (2 * NSLOT + 10), ACCEMPTY = bar0 + 8u * (2 * NSLOT + 11);
    const uint32_t IFULL = bar0 + 8u * (2 * NSLOT + 12), IFREE = bar0 + 8u * (2 * NSLOT + 13);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * NSLOT + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int NJ = (p.n_n + Cfg::NSTEP - 1) / Cfg::NSTEP;
    const int KC = p.kc;                         // 64-wide K chunks of the S contraction (even, <= 16)
    const int KCS = WIDE ? (KC < 8 ? KC : 8) : KC;   // ... of which the first KCS are stationary, the others streamed
    const int ND = (p.d_half + 255) / 256;       // pair MMAs of the dA^T product (256 feature columns each) per pass
    // the dA^T operand boxes are consumed as ADJACENT slot pairs (even, odd): when the streamed A chunks take an odd
    // number of slots per step (dim = 640, 896) an empty bubble slot restores the alignment
    const bool pad_slot = WIDE && (((KC - KCS) / 2) & 1) != 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(SFULL(s), 1); mbar_init(SEMPTY(s), EPI_ALL); }
        for (int b = 0; b < 2; ++b) {
            // keep_ds: a tile is free again once the dA^T MMAs have read it AND its spill to global memory has
            mbar_init(DSFULL(b), EPI_ALL); mbar_init(DSEMPTY(b), p.keep_ds ? 2 : 1); mbar_init(DSLOC(b), 256);
        }
        mbar_init(ACCFULL, 1); mbar_init(ACCEMPTY, EPI_ALL);
        mbar_init(IFULL, 1); mbar_init(IFREE, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmA64); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmBd);
        if (p.keep_ds) tma_prefetch_desc(&tmDS);
    }
    if (warp == 2) { tmem_alloc_cg2(smem_u32(tmem_holder), 512); tmem_relinquish_cg2(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t TM_ACC = tmem_base, TM_S = tmem_base + 256;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs, own halves)
        if (elect_one()) {
            int slot = 0; uint32_t ph = 0; uint32_t it = 0;
            const bool prof = p.prof != nullptr && pair == 0;
            long long w_empty = 0;
            const long long t_begin = clock64();
            auto put = [&](const CUtensorMap* tm, int c0, int c1) {
                if (prof) { const long long t0 = clock64(); mbar_wait(EMPTY(slot), ph ^ 1); w_empty += clock64() - t0; }
                else mbar_wait(EMPTY(slot), ph ^ 1);
                if (cta == 0) mbar_expect_tx(FULL(slot), 2 * Cfg::SLOT);
                if (p.dbg & 1) { c0 = 0; c1 = 0; }
                tma_load_2d_cg2(smem_u32(ring + slot * Cfg::SLOT), tm, mapa(FULL(slot), 0), c0, c1);
                if (++slot == NSLOT) { slot = 0; ph ^= 1; }
            };
            // two streamed K-chunks (c, c + 1) of this CTA's 64 A rows share one slot (2 x 8 KiB)
            auto put_a2 = [&](int c, int mb) {
                mbar_wait(EMPTY(slot), ph ^ 1);
                if (cta == 0) mbar_expect_tx(FULL(slot), 2 * Cfg::SLOT);
                const uint32_t dst = smem_u32(ring + slot * Cfg::SLOT), bar = mapa(FULL(slot), 0);
                tma_load_2d_cg2(dst, &tmA64, bar, c * KCHUNK, mb * TILE + (int)cta * 64);
                tma_load_2d_cg2(dst + 8192, &tmA64, bar, (c + 1) * KCHUNK, mb * TILE + (int)cta * 64);
                if (++slot == NSLOT) { slot = 0; ph ^= 1; }
            };
            // measurement switches (FLYP_DBG bits 8 / 16): the loads of the dA^T operand / of the S operand are skipped
            // altogether - the slot just changes hands with stale contents (wrong results, right timing)
            auto skip_slot = [&]() {
                mbar_wait(EMPTY(slot), ph ^ 1);
                if (cta == 0) mbar_arrive(FULL(slot));
                if (++slot == NSLOT) { slot = 0; ph ^= 1; }
            };
            auto load_s = [&](int t, int mb) {
                for (int c = 0; c < KC; ++c) {
                    if (WIDE && c >= KCS && (c & 1) == 0) put_a2(c, mb);
                    if (p.dbg & 16) skip_slot();
                    else put(&tmB, c * KCHUNK, t * Cfg::NSTEP + (int)cta * 128);
                }
                if (pad_slot) {                                // bubble: no data, the slot just changes hands
                    mbar_wait(EMPTY(slot), ph ^ 1);
                    if (cta == 0) mbar_arrive(FULL(slot));
                    if (++slot == NSLOT) { slot = 0; ph ^= 1; }
                }
            };
            auto load_t = [&](int t, int dcol0) {
                for (int dblk = 0; dblk < ND; ++dblk)
                    for (int jh = 0; jh < 2; ++jh)
                        for (int dsub = 0; dsub < 2; ++dsub) {
                            if (p.dbg & 8) skip_slot();
                            else put(&tmBd, dcol0 + (dblk * 2 + (int)cta) * 128 + dsub * 64, t * Cfg::NSTEP + jh * 128);
                        }
            };
            peer_wait_all(p.wait_b);      // multi-GPU: the N-side rows (and their fp16 copy) of every rank have arrived
            peer_wait_all(p.wait_bd);
            SweepItems iter(p.m_tiles, p.n_dh, p.sched_pairs, NJ, pair);
            ItemInfo ii;
            for (; iter.next(ii); ++it) {
                mbar_wait(IFREE, (it & 1) ^ 1);
                if (cta == 0) mbar_expect_tx(IFULL, 2 * KCS * 8192);
                for (int c = 0; c < KCS; ++c)
                    tma_load_2d_cg2(smem_u32(ist + c * 8192), &tmA64, mapa(IFULL, 0), c * KCHUNK,
                                    ii.mb * TILE + (int)cta * 64);
                const int dcol0 = ii.dh * p.d_half;
                load_s(ii.t0, ii.mb);
                for (int t = ii.t0 + 1; t < ii.t1; ++t) { load_s(t, ii.mb); load_t(t - 1, dcol0); }
                load_t(ii.t1 - 1, dcol0);
            }
            if (prof) { p.prof[8 + cta * 2] = (unsigned long long)(clock64() - t_begin); p.prof[9 + cta * 2] = w_empty; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (cta == 0 && elect_one()) {
            constexpr uint32_t IDESC_S = umma_idesc(128, 256, 1, 1, 0, 0);     // bf16 x bf16, both K-major
            constexpr uint32_t IDESC_D = umma_idesc(256, 128, 0, 0, 1, 0);     // fp16 x fp16, A MN-major, B K-major
            int slot = 0; uint32_t ph = 0; uint32_t gs = 0, gd = 0, it = 0;
            const bool prof = PROF && p.prof != nullptr && pair == 0;
            long long w_sempty = 0, w_full_s = 0, w_dsfull = 0, w_full_t = 0, w_acc = 0, w_ifull = 0;
            const long long t_begin = PROF ? clock64() : 0;
            auto twait = [&](uint32_t bar, uint32_t parity, long long& acc) {
                if constexpr (PROF) {
                    if (prof) { const long long t0 = clock64(); mbar_wait(bar, parity); acc += clock64() - t0; return; }
                }
                mbar_wait(bar, parity);
            };
            auto adv = [&]() { if (++slot == NSLOT) { slot = 0; ph ^= 1; } };
            // descriptor halves (sm100.cuh): per MMA only an integer add on the low word
            constexpr uint32_t HI = umma_desc_hi(1024);
            const uint32_t ist_lo = umma_desc_lo(smem_u32(ist), 16);            // stationary A chunks, K-major
            const uint32_t ring_lo = umma_desc_lo(smem_u32(ring), 16);          // streamed B chunks, K-major
            const uint32_t ring_lo_mn = umma_desc_lo(smem_u32(ring), Cfg::SLOT); // feature boxes, MN-major (LBO = next box)
            const uint32_t ds_lo = umma_desc_lo(smem_u32(ds), 16);              // dS tile, K-major
            // narrow features: the poll of the NEXT slot's barrier is issued before the MMAs of the current one, so that
            // its latency (and the branch on it) overlaps their issue instead of preceding the next chunk
            auto mma_s_narrow = [&]() {
                const int sb = gs & 1;
                twait(SEMPTY(sb), ((gs >> 1) & 1) ^ 1, w_sempty);
                const uint32_t d_tmem = TM_S + sb * 128;
                uint32_t ok = mbar_try_wait(FULL(slot), ph);
#pragma unroll 1
                for (int c = 0; c < KC; ++c) {
                    if (!ok) twait(FULL(slot), ph, w_full_s);
                    tc_fence_after();
                    const int cur = slot;
                    adv();
                    if (c + 1 < KC) ok = mbar_try_wait(FULL(slot), ph);
                    const uint32_t a_lo = ist_lo + c * (8192 >> 4), b_lo = ring_lo + cur * (Cfg::SLOT >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_cg2(d_tmem, umma_desc_join(a_lo + k * 2, HI), umma_desc_join(b_lo + k * 2, HI), IDESC_S,
                                     (c | k) != 0);
                    umma_commit_cg2(EMPTY(cur));
                }
                umma_commit_cg2(SFULL(sb));
                ++gs;
            };
            auto mma_s = [&]() {
                if constexpr (!WIDE) { mma_s_narrow(); return; }
                const int sb = gs & 1;
                twait(SEMPTY(sb), ((gs >> 1) & 1) ^ 1, w_sempty);
                tc_fence_after();
                const uint32_t d_tmem = TM_S + sb * 128;
                int a_slot = 0;
                uint32_t ok = mbar_try_wait(FULL(slot), ph);       // (polls issued one slot ahead, as in mma_s_narrow)
#pragma unroll 1
                for (int c = 0; c < KC; ++c) {
                    uint32_t a_lo;
                    if (c < KCS) {
                        a_lo = ist_lo + c * (8192 >> 4);
                    } else {
                        if ((c & 1) == 0) {                    // the slot holding the streamed A chunks c and c + 1
                            if (!ok) twait(FULL(slot), ph, w_full_s);
                            a_slot = slot;
                            adv();
                            ok = mbar_try_wait(FULL(slot), ph);
                        }
                        a_lo = ring_lo + a_slot * (Cfg::SLOT >> 4) + (c & 1) * (8192 >> 4);
                    }
                    if (!ok) twait(FULL(slot), ph, w_full_s);
                    tc_fence_after();
                    const int cur = slot;
                    adv();
                    ok = mbar_try_wait(FULL(slot), ph);
                    const uint32_t b_lo = ring_lo + cur * (Cfg::SLOT >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_cg2(d_tmem, umma_desc_join(a_lo + k * 2, HI), umma_desc_join(b_lo + k * 2, HI), IDESC_S,
                                     (c | k) != 0);
                    umma_commit_cg2(EMPTY(cur));
                    if (c >= KCS && (c & 1) == 1) umma_commit_cg2(EMPTY(a_slot));
                }
                if (pad_slot) {
                    if (!ok) twait(FULL(slot), ph, w_full_s);
                    umma_commit_cg2(EMPTY(slot));
                    adv();
                }
                umma_commit_cg2(SFULL(sb));
                ++gs;
            };
            auto mma_d = [&](bool first) {
                if (first) twait(ACCEMPTY, (it & 1) ^ 1, w_acc);
                const int db = gd % Cfg::DS_TILES;
                twait(DSFULL(db), (gd / Cfg::DS_TILES) & 1, w_dsfull);
                // (the polls of the next group's two slots are issued before the MMAs of the current group)
                uint32_t ok0 = mbar_try_wait(FULL(slot), ph), ok1 = mbar_try_wait(FULL(slot + 1), ph);
#pragma unroll 1
                for (int g = 0; g < 2 * ND; ++g) {                  // g = dblk * 2 + jh
                    const int dblk = g >> 1, jh = g & 1;
                    if (!ok0) twait(FULL(slot), ph, w_full_t);      // the two feature boxes (64 + 64 columns) of this K half
                    if (!ok1) twait(FULL(slot + 1), ph, w_full_t);
                    tc_fence_after();
                    const int cur = slot;
                    adv(); adv();
                    if (g + 1 < 2 * ND) { ok0 = mbar_try_wait(FULL(slot), ph); ok1 = mbar_try_wait(FULL(slot + 1), ph); }
                    // A: [K = 16 n-rows][M = 128 feature columns as two 64-wide boxes], MN-major
                    const uint32_t a_lo = ring_lo_mn + cur * (Cfg::SLOT >> 4);
                    // B: dS [N = 64 rows per CTA][K], K-major, four 64-wide chunks of 8 KiB
                    const uint32_t b_lo = ds_lo + db * (Cfg::DS_TILE >> 4) + jh * (2 * 8192 >> 4);
                    const uint32_t d_tmem = TM_ACC + dblk * 128;
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8)
                        umma_f16_cg2(d_tmem, umma_desc_join(a_lo + k8 * (2048 >> 4), HI),
                                     umma_desc_join(b_lo + (k8 >> 2) * (8192 >> 4) + (k8 & 3) * 2, HI), IDESC_D,
                                     !(first && jh == 0 && k8 == 0));
                    umma_commit_cg2(EMPTY(cur));
                    umma_commit_cg2(EMPTY(cur + 1));
                }
                umma_commit_cg2(DSEMPTY(db));
                ++gd;
            };
            SweepItems iter(p.m_tiles, p.n_dh, p.sched_pairs, NJ, pair);
            ItemInfo ii;
            for (; iter.next(ii); ++it) {
                const int nj = ii.t1 - ii.t0;
                twait(IFULL, it & 1, w_ifull);
                tc_fence_after();
                mma_s();
                for (int t = 1; t < nj; ++t) { mma_s(); mma_d(t == 1); }
                mma_d(nj == 1);
                umma_commit_cg2(ACCFULL);
                umma_commit_cg2(IFREE);
            }
            if (PROF && prof) {
                p.prof[0] = (unsigned long long)(clock64() - t_begin);
                p.prof[1] = w_sempty; p.prof[2] = w_full_s; p.prof[3] = w_dsfull; p.prof[4] = w_full_t;
                p.prof[5] = w_acc; p.prof[6] = w_ifull; p.prof[7] = gs;
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ dS spill (both CTAs, keep_ds only)
        // The staged fp16 dS tile of every step ([64 rows][256 columns] per CTA, exactly what the dA^T MMAs read) is also
        // written to global memory: the gradient of the OTHER operand is then a plain product dS^T . A over these values
        // (clip_dst_gemm.cu) instead of a second sweep that recomputes every logit.  Four cp.async.bulk.tensor stores per
        // step (the TMA undoes the swizzle), with an L2 evict-first policy: 2 GiB of dS stream through the L2 per sweep and
        // must not displace the 64 MiB of operands every pair re-reads (L2 hit rate of the operand loads 95 % without the
        // spill, 77 % with a plain store).  (Measured alternatives: 32-byte stores from the epilogue's registers 1.92 ms
        // per sweep, a copy warp with ordinary loads / stores 2.00 ms, this store without the policy 1.80 ms; no spill
        // 1.53 ms.)
        if (p.keep_ds && elect_one()) {
            uint32_t gs = 0;
            const uint64_t pol_first = l2_policy_evict_first();
            SweepItems iter(p.m_tiles, p.n_dh, p.sched_pairs, NJ, pair);
            ItemInfo ii;
            while (iter.next(ii)) {
                for (int t = ii.t0; t < ii.t1; ++t, ++gs) {
                    const int db = gs % Cfg::DS_TILES;
                    mbar_wait(DSLOC(db), (gs / Cfg::DS_TILES) & 1);
                    if (ii.dh == 0 && !(p.dbg & 32)) {         // (FLYP_DBG bit 5: measurement - the spill is skipped)
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4)
                            tma_store_2d_hint(&tmDS, smem_u32(ds + db * Cfg::DS_TILE + c4 * 8192), t * Cfg::NSTEP + c4 * 64,
                                              ii.mb * TILE + (int)cta * 64, pol_first);
                        tma_store_commit();
                        tma_store_wait_read0();
                    }
                    mbar_arrive(DSEMPTY(db));
                }
            }
            tma_store_wait_all0();
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (both CTAs)
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;
        const float s = *p.scale;
        const float c1 = s * LOG2E2;
        float G, invG;
        staging_scale(p.gmax_bits, G, invG);
        const bool fast = p.fast_info != nullptr && p.fast_info[1] != 0.f;
        const float c0 = fast ? p.fast_info[0] : 0.f;
        const uint32_t R_SEMPTY0 = mapa(SEMPTY(0), 0), R_SEMPTY1 = mapa(SEMPTY(1), 0);
        const uint32_t R_DSFULL0 = mapa(DSFULL(0), 0), R_DSFULL1 = mapa(DSFULL(1), 0), R_ACCEMPTY = mapa(ACCEMPTY, 0);
        const int rloc = (q & 1) * 32 + lane;       // row within this CTA's 64-row slice
        const int jq = q >> 1;                      // which 128-column half of the step this lane quarter holds
        uint32_t gs = 0, it = 0;
        const bool prof = p.prof != nullptr && pair == 0 && cta == 0 && threadIdx.x == 128;
        long long w_sfull = 0, w_dsempty = 0, w_accfull = 0;
        const long long t_begin = prof ? clock64() : 0;
        auto ewait = [&](uint32_t bar, uint32_t parity, long long& acc) {
            if (prof) { const long long t0 = clock64(); mbar_wait(bar, parity); acc += clock64() - t0; }
            else mbar_wait(bar, parity);
        };
        int* const red_i = reinterpret_cast<int*>(red);    // [16] last-arriver flag, [17] slot count, [18 ..] slot list
        const bool want_ds = p.dscale_part != nullptr;
        float dsum = 0.f;                           // d(scale) share of this thread over all items of the pair
        SweepItems iter(p.m_tiles, p.n_dh, p.sched_pairs, NJ, pair);
        ItemInfo ii;
        for (; iter.next(ii); ++it) {
            const int m = ii.mb * TILE + (int)cta * 64 + rloc;
            const RowCtx rc = load_row_ctx<ROW_TERM>(p, m, fast, G);
            for (int t = ii.t0; t < ii.t1; ++t, ++gs) {
                const int sb = gs & 1;
                ewait(SFULL(sb), (gs >> 1) & 1, w_sfull);
                tc_fence_after();
                const uint32_t taddr = TM_S + ((uint32_t)(q * 32) << 16) + sb * 128 + h * 64;
                uint32_t r0[32], r1[32];
                tmem_ld_32x32b_x32(taddr, r0);
                tmem_ld_32x32b_x32(taddr + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_cluster(sb ? R_SEMPTY1 : R_SEMPTY0);
                const int n0 = t * Cfg::NSTEP + jq * 128 + h * 64;
                float v[64];
                ds_tile<ROW_TERM, COL_TERM>(r0, r1, p, rc, n0, c1, fast, c0, G, v, want_ds && ii.dh == 0, dsum);
                uint32_t pk[32];
                pack_ds(v, pk);
                const int db = gs % Cfg::DS_TILES;          // dS tile of this step
                ewait(DSEMPTY(db), ((gs / Cfg::DS_TILES) & 1) ^ 1, w_dsempty);
                store_ds_row(ds + db * Cfg::DS_TILE + (jq * 2 + h) * 8192 + rloc * 128, rloc, pk);
                fence_proxy_async_smem();
                mbar_arrive_cluster(db ? R_DSFULL1 : R_DSFULL0);
                if (p.keep_ds) mbar_arrive(DSLOC(db));
            }
            // -------- row block done: drain dA^T (lanes = feature columns, TMEM columns = the 128 rows of the block)
            ewait(ACCFULL, it & 1, w_accfull);
            tc_fence_after();
            const float omul = s * p.out_mul * invG;
            for (int dblk = 0; dblk < ND; ++dblk) {
                const int dl = (dblk * 2 + (int)cta) * 128 + q * 32 + lane;      // column within this pass
                const int d = ii.dh * p.d_half + dl;
                const bool dvalid = dl < p.d_half && d < p.d_out;
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(TM_ACC + ((uint32_t)(q * 32) << 16) + dblk * 128 + h * 64 + cc * 32, r);
                    tmem_ld_wait();
                    if (dvalid) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const int ri = h * 64 + cc * 32 + k;
                            const int mi = ii.mb * TILE + ri;
                            if (mi < p.n_m) {
                                const float a = __uint_as_float(r[k]);
                                if (ii.part >= 0)
                                    p.part_out[((size_t)ii.part * TILE + ri) * p.d_half + dl] = a * omul;
                                else if (p.out_fp32)
                                    reinterpret_cast<float*>(p.out)[(size_t)mi * p.ld_out + d] = a * omul;
                                else
                                    reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)mi * p.ld_out + d] =
                                        __float2bfloat16_rn(a * omul);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(R_ACCEMPTY);
        }
        if (want_ds) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, off);
            if (lane == 0) red[warp - 4] = dsum;
            epi_bar_sync2();
            if (warp == 4) {
                float tot = lane < 8 ? red[lane] : 0.f;
#pragma unroll
                for (int off = 4; off >= 1; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
                if (lane == 0) p.dscale_part[pair * 2 + (int)cta] = tot * invG;
            }
        }
        if (p.grid_cnt != nullptr) sweep_tail_reduce(p, et, NJ, red_i);
        if (prof) {
            p.prof[12] = (unsigned long long)(clock64() - t_begin);
            p.prof[13] = w_sfull; p.prof[14] = w_dsempty; p.prof[15] = w_accfull;
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_cg2(tmem_base, 512);
}

void launch_bwd_pair(const CUtensorMap& tmA64, const CUtensorMap& tmB, const CUtensorMap& tmBd, const CUtensorMap* tmDS_opt,
                     const BwdParams& p, int num_sms, cudaStream_t st) {
    const CUtensorMap& tmDS = tmDS_opt != nullptr ? *tmDS_opt : tmB;      // (unused unless p.keep_ds)
    (void)num_sms;
    const int grid = p.sched_pairs * 2;
    const size_t smem = bwd_pair_smem_bytes();
    const bool row_term = p.wr != nullptr, col_term = p.wc != nullptr;
    // cooperative: the end-of-sweep grid barrier needs every CTA resident (the launch fails instead of hanging otherwise)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NTHREADS2); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = p.grid_cnt != nullptr ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
#define FLYP_LAUNCH_BWD2_P(R, C, W, P)                                                                         \
    do {                                                                                                       \
        static bool attr_done[64] = {false};                                                                  \
        ensure_smem_attr(bwd_pair_kernel<R, C, W, P>, smem, attr_done);                                        \
        cudaLaunchKernelEx(&cfg, bwd_pair_kernel<R, C, W, P>, tmA64, tmB, tmBd, tmDS, p);                      \
    } while (0)
#define FLYP_LAUNCH_BWD2(R, C, W)                                                                              \
    do {                                                                                                       \
        if (p.prof != nullptr) FLYP_LAUNCH_BWD2_P(R, C, W, true); else FLYP_LAUNCH_BWD2_P(R, C, W, false);     \
    } while (0)
    const bool wide = p.kc > 8 || p.n_dh > 1;
    if (wide) {
        if (row_term && col_term) FLYP_LAUNCH_BWD2(true, true, true);
        else if (row_term) FLYP_LAUNCH_BWD2(true, false, true);
        else FLYP_LAUNCH_BWD2(false, true, true);
    } else {
        if (row_term && col_term) FLYP_LAUNCH_BWD2(true, true, false);
        else if (row_term) FLYP_LAUNCH_BWD2(true, false, false);
        else FLYP_LAUNCH_BWD2(false, true, false);
    }
#undef FLYP_LAUNCH_BWD2
#undef FLYP_LAUNCH_BWD2_P
}

}  // namespace flyp
