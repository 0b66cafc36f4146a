// Host-side property check of flyp_b200/csrc/sched.h (compiled with g++ by tests/test_sched_host.py).
// For many (m_tiles, n_dh, NJ, P): every (virtual row block, column step) unit is covered by exactly one item; whole items
// are the only ones with part = -1; partial slots are unique, below 2 * P; and TailParts (what the reduction kernel uses)
// enumerates, for every tail block, exactly the slots of the partial items that cover it, in pair order - or reports the
// block as swept whole exactly when a single whole item covers it.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#include "sched.h"

using namespace flyp;

static int check_p(int m_tiles, int n_dh, int NJ, int P);
static int check(int m_tiles, int n_dh, int NJ, int npairs) {
    const int v_tiles = m_tiles * n_dh;
    const long long S_total = (long long)v_tiles * NJ;
    return check_p(m_tiles, n_dh, NJ, (int)(S_total < npairs ? S_total : npairs));      // bwd_pair_sched_pairs
}
static int check_p(int m_tiles, int n_dh, int NJ, int P) {
    const int v_tiles = m_tiles * n_dh;
    std::vector<int> cover((size_t)v_tiles * NJ, 0);
    std::map<int, std::vector<int>> slots_of_block;                     // virtual block -> partial slots, in pair order
    std::set<int> used_slots, whole_blocks;
    for (int pair = 0; pair < P; ++pair) {
        SweepItems it(m_tiles, n_dh, P, NJ, pair);
        ItemInfo ii;
        int n_items = 0;
        while (it.next(ii)) {
            ++n_items;
            if (ii.mb < 0 || ii.mb >= m_tiles || ii.dh < 0 || ii.dh >= n_dh || ii.t0 < 0 || ii.t1 > NJ || ii.t0 >= ii.t1) return 1;
            const int vb = ii.mb * n_dh + ii.dh;
            for (int t = ii.t0; t < ii.t1; ++t) cover[(size_t)vb * NJ + t]++;
            const bool whole = ii.t0 == 0 && ii.t1 == NJ;
            if (whole != (ii.part == -1)) return 2;
            if (whole) { if (!whole_blocks.insert(vb).second) return 3; }
            else {
                if (ii.part < 0 || ii.part >= 2 * P) return 4;
                if (!used_slots.insert(ii.part).second) return 5;
                slots_of_block[vb].push_back(ii.part);
            }
        }
        if (n_items == 0) return 6;                                     // every scheduled pair has work
    }
    for (int c : cover) if (c != 1) return 7;
    const int first = (v_tiles / P) * P;
    for (int vb = 0; vb < v_tiles; ++vb) {
        if (vb < first) { if (!whole_blocks.count(vb) || slots_of_block.count(vb)) return 8; continue; }
        TailParts parts;
        const bool split = parts.init(v_tiles, NJ, P, vb - first);
        if (split == (whole_blocks.count(vb) != 0)) return 9;
        if (!split) continue;
        std::vector<int> got;
        for (int s = parts.next(); s >= 0; s = parts.next()) got.push_back(s);
        if (got != slots_of_block[vb]) return 10;
    }
    return 0;
}

// Forward flat schedule: every (column unit, row tile) unit covered exactly once; the slots written for a column unit are
// exactly 0 .. fwd_unit_slots() - 1 (what the finalize kernel sums), all below fwd_sched_slots() (the workspace bound);
// every worker has work; items of one worker do not repeat a column unit.
static int check_fwd(int m_tiles, int n_units, int n_local, int rot, int max_workers) {
    const int P = fwd_sched_workers(m_tiles, n_units, max_workers);
    const int bound = fwd_sched_slots(m_tiles, n_units, n_local, P);
    std::vector<int> cover((size_t)m_tiles * n_units, 0);
    std::map<int, std::set<int>> slots_of_unit;
    for (int q = 0; q < P; ++q) {
        FwdItems it(m_tiles, n_units, n_local, rot, P, q);
        FwdItem fi;
        int n_items = 0;
        std::set<int> seen;
        bool remote_seen = false;
        while (it.next(fi)) {
            ++n_items;
            if (fi.u < 0 || fi.u >= n_units || fi.mt0 < 0 || fi.mt1 > m_tiles || fi.mt0 >= fi.mt1) return 21;
            if (fi.slot < 0 || fi.slot >= bound) return 22;
            if (!seen.insert(fi.u).second) return 23;
            if (!slots_of_unit[fi.u].insert(fi.slot).second) return 24;
            for (int mt = fi.mt0; mt < fi.mt1; ++mt) cover[(size_t)fi.u * m_tiles + mt]++;
            int ub = fi.u - rot; if (ub < 0) ub += n_units;
            if (ub >= n_local) remote_seen = true;
            else if (remote_seen) return 29;                   // local units always come first
        }
        if (n_items == 0 && n_local == 0) return 25;         // (two phases: a tiny phase may leave late workers idle)
    }
    for (int c : cover) if (c != 1) return 26;
    for (int u = 0; u < n_units; ++u) {
        int ub = u - rot; if (ub < 0) ub += n_units;
        const int ns = fwd_unit_slots(ub, m_tiles, n_units, n_local, P);
        const std::set<int>& got = slots_of_unit[u];
        if ((int)got.size() != ns) return 27;
        int k = 0;
        for (int sl : got) if (sl != k++) return 28;
    }
    return 0;
}

int main() {
    long long n = 0;
    for (int maxw : {1, 2, 5, 37, 74, 148})
        for (int m_tiles = 1; m_tiles <= 300; m_tiles += (m_tiles < 40 ? 1 : 29))
            for (int n_units : {1, 2, 3, 4, 8, 15, 16, 31, 32, 128, 256, 512}) {
                const int rots[3] = {0, n_units / 3, n_units - 1};
                const int locals[4] = {0, 1, n_units / 8, n_units / 2};
                for (int rot : rots)
                    for (int n_local : locals) {
                        if (n_local > n_units) continue;
                        const int rc = check_fwd(m_tiles, n_units, n_local, rot, maxw);
                        if (rc) { printf("FAIL fwd rc=%d m_tiles=%d n_units=%d n_local=%d rot=%d maxw=%d\n", rc, m_tiles, n_units, n_local, rot, maxw); return 1; }
                        ++n;
                    }
            }
    // products over the kept dS (clip_dst_gemm.cu): the pair count comes from dst_sched_pairs (whole-tile rounds or a flat
    // tail); out_tiles x n_dh virtual tiles of k_blocks contraction blocks
    for (int max_pairs : {1, 2, 37, 66, 74})
        for (int n_dh = 1; n_dh <= 4; ++n_dh)
            for (int out_tiles = 1; out_tiles <= 520; out_tiles += (out_tiles < 140 ? 1 : 31))
                for (int kb : {1, 2, 7, 8, 9, 32, 64, 128, 256, 512}) {
                    const int P = dst_sched_pairs(out_tiles * n_dh, kb, max_pairs);
                    if (P < 1 || P > max_pairs) { printf("FAIL dst pairs %d\n", P); return 1; }
                    const int rc = check_p(out_tiles, n_dh, kb, P);
                    if (rc) { printf("FAIL dst rc=%d out_tiles=%d n_dh=%d kb=%d P=%d\n", rc, out_tiles, n_dh, kb, P); return 1; }
                    ++n;
                }
    // the cases the design notes quote: 8 ranks x 4096 rows, 256-column tiles -> four whole-tile rounds on 64 pairs;
    // one GPU at B = 32768 -> the flat schedule on all 74 pairs; a 2-GPU B = 8192 problem -> one whole-tile round
    if (dst_sched_pairs(256, 32, 74) != 64 || dst_sched_pairs(128, 256, 74) != 74 || dst_sched_pairs(64, 32, 74) != 64) {
        printf("FAIL dst_sched_pairs reference cases\n");
        return 1;
    }
    const int npairs_list[] = {1, 2, 3, 7, 64, 66, 70, 74};
    for (int npairs : npairs_list)
        for (int n_dh = 1; n_dh <= 2; ++n_dh)
            for (int m_tiles = 1; m_tiles <= 300; m_tiles += (m_tiles < 80 ? 1 : 37))
                for (int NJ : {1, 2, 3, 5, 16, 31, 128, 256}) {
                    const int rc = check(m_tiles, n_dh, NJ, npairs);
                    if (rc) { printf("FAIL rc=%d m_tiles=%d n_dh=%d NJ=%d npairs=%d\n", rc, m_tiles, n_dh, NJ, npairs); return 1; }
                    ++n;
                }
    printf("OK %lld schedules\n", n);
    return 0;
}
