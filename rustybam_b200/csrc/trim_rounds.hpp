// trim_rounds.hpp — host side of `rb trim-paf`: the round structure of Paf::overlapping_paf_recs (paf.rs:210-305).
//
// The reference sorts the records by query name (stable), lists every pair of records of one query whose query
// intervals overlap without containment, orders the pairs by overlap (largest first, stable), trims the FIRST pair of
// every query name and — if any pair had to wait — starts over on the trimmed set.  Only the query spans matter for
// that bookkeeping.  This file is the whole-set statement of one round, literal to the reference's loop; the GPU runs the
// per-name form of it (trim_select_group / k_trim_select in trim_core.cuh / trim_kernels.cu), and the CPU fuzz harness
// (tests/native/trim_core_check.cpp) checks the two against each other every round and the result against the oracle.
// Plain C++ (no CUDA).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace rb {

struct TrimSpan {        // one record, in the order of the name-sorted set
    uint64_t q_st, q_en; // current query span
    uint32_t name;       // dense rank of its query name (equal names <=> equal rank; non-decreasing over the set)
};
struct TrimPairSel { uint32_t left, right; uint64_t st_ovl, en_ovl; };

// One pass of paf.rs:229-287: fills `contained` (reset, like the reference's fresh vector), the pairs to trim in this
// round (one per query name) and returns how many pairs have to wait for the next round.
inline size_t trim_round(const std::vector<TrimSpan>& recs, std::vector<uint8_t>& contained, std::vector<TrimPairSel>& sel) {
    struct Cand { uint64_t overlap; uint32_t l, r; };
    std::vector<Cand> pairs;
    contained.assign(recs.size(), 0);
    sel.clear();
    if (recs.size() < 2) return 0;
    for (size_t i = 0; i + 1 < recs.size(); i++) {
        const TrimSpan& a = recs[i];
        for (size_t j = i + 1; j < recs.size() && recs[j].name == a.name; j++) {
            const TrimSpan& b = recs[j];
            const uint64_t mn = std::min(a.q_en, b.q_en), mx = std::max(a.q_st, b.q_st);  // bed.rs:74-85
            const uint64_t overlap = mn < mx ? 0 : mn - mx;
            if (overlap < 1) continue;
            if (overlap == b.q_en - b.q_st) contained[j] = 1;
            else if (overlap == a.q_en - a.q_st) contained[i] = 1;
            else if (a.q_st <= b.q_st) pairs.push_back(Cand{overlap, (uint32_t)i, (uint32_t)j});
            else pairs.push_back(Cand{overlap, (uint32_t)j, (uint32_t)i});
        }
    }
    std::stable_sort(pairs.begin(), pairs.end(), [](const Cand& x, const Cand& y) { return x.overlap > y.overlap; });  // paf.rs:261
    std::vector<uint8_t> seen;  // per name rank
    size_t waiting = 0;
    for (const Cand& c : pairs) {
        const uint32_t nm = recs[c.l].name;
        if (nm >= seen.size()) seen.resize((size_t)nm + 1, 0);
        if (seen[nm]) { waiting++; continue; }
        seen[nm] = 1;
        const TrimSpan &l = recs[c.l], &r = recs[c.r];
        sel.push_back(TrimPairSel{c.l, c.r, std::max(l.q_st, r.q_st), std::min(l.q_en, r.q_en)});
    }
    return waiting;
}

}  // namespace rb
