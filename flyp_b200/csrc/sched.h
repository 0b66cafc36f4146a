// sched.h — integer arithmetic of the backward sweep's work schedule, shared by the pair kernel (device), the partial-sum
// reduction kernel (device) and the host (workspace sizing; tests/test_sched_host.py compiles this header with g++ and
// checks that every (virtual row block, column step) unit is covered exactly once and that the reduction finds exactly
// the partial accumulators the kernel wrote).  No CUDA dependencies.
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

}  // namespace flyp
