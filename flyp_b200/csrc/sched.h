// sched.h — integer arithmetic of the work schedules, shared by the tensor-core kernels (device), the reductions that
// consume their partial results (device) and the host (workspace sizing).  tests/test_sched_host.py compiles this header
// with g++ and checks that every unit of work is covered exactly once and that the consumers find exactly the partial
// slots the producers wrote.  No CUDA dependencies.
#pragma once
#if defined(__CUDACC__)
#define FLYP_HD __host__ __device__ __forceinline__
#else
#define FLYP_HD inline
#endif

namespace flyp {

// first flat unit of pair q when S units are cut into `pairs` equal contiguous ranges
FLYP_HD long long flat_start(int q, long long S, int pairs) { return (long long)q * S / pairs; }

// One work item of a CTA pair: row block mb, pass dh over the output columns, column steps [t0, t1).
// part = -1: the item covers its whole virtual row block and writes the output directly; otherwise the index of the fp32
// partial slot it accumulates into.
struct ItemInfo { int mb, dh, t0, t1, part; };

// Schedule of a sweep over v_tiles = m_tiles * n_dh virtual row blocks (row block, d-half; d-half fastest) of NJ column
// steps each on P CTA pairs: floor(v_tiles / P) rounds of whole blocks, one per pair and round, in lockstep over the
// columns; then the remaining blocks as a FLAT tail - their (block, step) units, block-major, cut into P equal
// contiguous ranges, each range cut again at block boundaries.  A range that covers a block only partly accumulates
// that part into slot 2 * pair + (0: it is the pair's first tail item, 1: a later one).
struct SweepItems {
    long long pos, end;
    int NJ, pair, ord, round, full_rounds, P, n_dh;
    FLYP_HD SweepItems(int m_tiles, int n_dh_, int P_, int NJ_, int pair_) {
        P = P_; NJ = NJ_; pair = pair_; ord = 0; round = 0; n_dh = n_dh_;
        const int vtiles = m_tiles * n_dh;
        full_rounds = vtiles / P;
        const long long S = (long long)(vtiles - full_rounds * P) * NJ;      // units of the flat tail
        pos = flat_start(pair_, S, P);
        end = flat_start(pair_ + 1, S, P);
    }
    FLYP_HD bool next(ItemInfo& r) {
        if (round < full_rounds) {
            const int vb = round * P + pair;
            r.mb = vb / n_dh; r.dh = vb - r.mb * n_dh; r.t0 = 0; r.t1 = NJ; r.part = -1;
            ++round;
            return true;
        }
        if (pos >= end) return false;
        const int tb = (int)(pos / NJ);
        const int vb = full_rounds * P + tb;
        r.mb = vb / n_dh; r.dh = vb - r.mb * n_dh;
        r.t0 = (int)(pos - (long long)tb * NJ);
        const long long room = NJ - r.t0, len = end - pos;
        r.t1 = r.t0 + (int)(len < room ? len : room);
        r.part = (r.t0 == 0 && r.t1 == NJ) ? -1 : 2 * pair + (ord == 0 ? 0 : 1);
        pos += r.t1 - r.t0;
        ++ord;
        return true;
    }
};

// The partial slots of tail block tb (0-based among the blocks beyond the whole-block rounds), in pair order, as the
// reduction derives them without any table.  Returns false when the block was swept whole (nothing to reduce).
struct TailParts {
    long long S, lo, hi;
    int pairs, q;
    FLYP_HD bool init(int v_tiles, int NJ, int pairs_, int tb) {
        pairs = pairs_;
        const int first = (v_tiles / pairs) * pairs;
        S = (long long)(v_tiles - first) * NJ; lo = (long long)tb * NJ; hi = lo + NJ;
        q = (int)(lo * pairs / S);
        while (q + 1 < pairs && flat_start(q + 1, S, pairs) <= lo) ++q;
        while (q > 0 && flat_start(q, S, pairs) > lo) --q;
        return !(flat_start(q, S, pairs) <= lo && flat_start(q + 1, S, pairs) >= hi);
    }
    // next slot, or -1 when done
    FLYP_HD int next() {
        // pairs whose range is empty (fewer tail units than pairs) own no item and no slot
        while (q < pairs && flat_start(q, S, pairs) < hi && flat_start(q + 1, S, pairs) == flat_start(q, S, pairs)) ++q;
        if (!(q < pairs && flat_start(q, S, pairs) < hi)) return -1;
        const int slot = 2 * q + (flat_start(q, S, pairs) >= lo ? 0 : 1);
        ++q;
        return slot;
    }
};

// CTA pairs of the products over the kept dS (clip_dst_gemm.cu): v_tiles output tiles of k_blocks contraction blocks
// each, walked with SweepItems / TailParts like the sweep.  Whole tiles only (ceil(v_tiles / pairs) rounds on a number
// of pairs that divides v_tiles) when that costs no more than the flat schedule's equal shares PLUS its tail - partial
// tiles written to scratch, a grid barrier, the sum: about TAIL_BLOCKS contraction blocks' worth of time during which no
// MMA runs (and, when the output goes to other GPUs, all of the split tiles' NVLink traffic comes at the very end).
// Otherwise every pair takes part, but no range is shorter than MIN_BLOCKS blocks.
FLYP_HD int dst_sched_pairs(int v_tiles, int k_blocks, int max_pairs) {
    constexpr int TAIL_BLOCKS = 48, MIN_BLOCKS = 8;
    int npairs = max_pairs;
    const long long S = (long long)v_tiles * k_blocks;
    {
        const int rounds = (v_tiles + npairs - 1) / npairs;
        const int pw = (v_tiles + rounds - 1) / rounds;
        if (v_tiles % pw == 0 && (long long)rounds * k_blocks <= S / npairs + TAIL_BLOCKS) return pw;
    }
    long long cap = S / MIN_BLOCKS;
    if (cap < v_tiles) cap = v_tiles;
    if (cap < npairs) npairs = (int)cap;
    return (int)(S < npairs ? S : npairs);
}

// ---------------------------------------------------------------------------------------------------- forward
// Flat schedule of the forward statistics sweep.  A unit = (column unit u, row tile mt): u is one 128-column block of
// the N side (the multicast kernel: a pair of adjacent blocks, swept by the two CTAs of a cluster), mt one 128-row tile
// of the M side.  Units are ordered column-unit-major, the column units in ROTATED order (ordinal ub = (u - rot) mod NU:
// multi-GPU ranks start on their own rows and follow the order in which the other ranks' rows arrive).
// The ordinals are split in two PHASES - [0, NA): the rank's own column units, whose operand is already in memory;
// [NA, NU): the units of the other ranks, still arriving over NVLink (NA = 0: a single phase) - and the units of each
// phase are cut into equal contiguous ranges, one per worker (CTA or cluster; a phase with fewer units than workers uses
// the first workers only).  So EVERY worker starts on local data while the gather is in flight.  A worker's range is cut
// again at column-unit boundaries: each piece is one work item (the column unit's operand stays in shared memory while
// the row tiles stream by).  The column sums of an item go to partial slot `slot` of its column unit: the ordinal of the
// worker among the workers that touch the unit.
FLYP_HD int flat_worker_of(long long x, long long S, int P) { return (int)(((x + 1) * P + S - 1) / S) - 1; }

struct FwdItem { int u, mt0, mt1, slot; };

struct FwdPhase {
    long long S;      // units of the phase
    int P;            // workers that take part (each gets at least one unit)
    int ub0;          // first column-unit ordinal
};
FLYP_HD FwdPhase fwd_phase(int ph, int m_tiles, int n_units, int n_local, int workers) {
    FwdPhase f;
    const int nu = ph == 0 ? n_local : n_units - n_local;
    f.ub0 = ph == 0 ? 0 : n_local;
    f.S = (long long)nu * m_tiles;
    f.P = (int)(f.S < workers ? f.S : workers);
    return f;
}

struct FwdItems {
    long long pos, end;
    FwdPhase ph;
    int MT, NU, NA, rot, W, q, phase;
    FLYP_HD FwdItems(int m_tiles, int n_units, int n_local, int rot_, int workers, int q_) {
        MT = m_tiles; NU = n_units; NA = n_local; rot = rot_; W = workers; q = q_;
        phase = -1; pos = end = 0;
    }
    FLYP_HD bool next(FwdItem& r) {
        while (pos >= end) {
            if (++phase > 1) return false;
            ph = fwd_phase(phase, MT, NU, NA, W);
            if (q < ph.P) { pos = flat_start(q, ph.S, ph.P); end = flat_start(q + 1, ph.S, ph.P); }
            else pos = end = 0;
        }
        const int ubl = (int)(pos / MT);                       // ordinal within the phase
        r.mt0 = (int)(pos - (long long)ubl * MT);
        const long long room = MT - r.mt0, len = end - pos;
        r.mt1 = r.mt0 + (int)(len < room ? len : room);
        int u = ph.ub0 + ubl + rot;
        if (u >= NU) u -= NU;
        r.u = u;
        r.slot = q - flat_worker_of((long long)ubl * MT, ph.S, ph.P);
        pos += r.mt1 - r.mt0;
        return true;
    }
};
// workers to launch (every one of them gets at least one unit in some phase)
FLYP_HD int fwd_sched_workers(int m_tiles, int n_units, int max_workers) {
    const long long S = (long long)m_tiles * n_units;
    return (int)(S < max_workers ? S : max_workers);
}
// upper bound on the partial column-sum slots of a column unit
FLYP_HD int fwd_sched_slots(int m_tiles, int n_units, int n_local, int workers) {
    int best = 1;
    for (int p = 0; p < 2; ++p) {
        const FwdPhase f = fwd_phase(p, m_tiles, n_units, n_local, workers);
        if (f.S == 0) continue;
        const long long L = f.S / f.P;                           // shortest range, >= 1
        long long s = (m_tiles - 1 + L - 1) / L + 1;
        if (s > m_tiles) s = m_tiles;
        if (s > best) best = (int)s;
    }
    return best;
}
// slots actually written for the column unit with rotated ordinal ub
FLYP_HD int fwd_unit_slots(int ub, int m_tiles, int n_units, int n_local, int workers) {
    const FwdPhase f = fwd_phase(ub < n_local ? 0 : 1, m_tiles, n_units, n_local, workers);
    const long long lo = (long long)(ub - f.ub0) * m_tiles;
    return flat_worker_of(lo + m_tiles - 1, f.S, f.P) - flat_worker_of(lo, f.S, f.P) + 1;
}

}  // namespace flyp
