// lift_core.cuh — run-length ("closed form") statement of one (BED window, PAF record) liftover.
//
// Replaces, per pair, the reference's per-base machinery:
//   liftover.rs:17-105   trim_paf_rec_to_rgn
//   paf.rs:541-561       tpos_to_idx / tpos_to_idx_match   (binary search over 8 B/column + slide)
//   paf.rs:593-620       subset_cigar / collapse_long_cigar
//   paf.rs:825-857       check_integrity (recomputes nmatch / aln_len)
//   bamstats.rs:107-142  add_stats_from_cigar on the trimmed CIGAR (fused: prefix-sum differences)
// with two searches over the sampled target prefix sums (one 48-byte sample per 32 ops) plus a
// <=32-op walk each.  Column indices of the reference's arrays are never built: column c of a
// record == (op index, offset in op), and A (columns before an op) is one of the counters.
//
// __host__ __device__: the same code is fuzzed on the CPU against the literal per-base oracle
// (tests/native/lift_core_check.cpp) and runs inside k_lift on the GPU.
#pragma once
#include "rb_common.cuh"

namespace rb {

// Read access to ops / samples.  A thread block may stage a window of both in shared memory
// (s_ops / s_smp cover ops [so_lo, so_hi) and chunks [sc_lo, sc_hi)); everything else comes from HBM.
struct OpsView {
    const uint32_t* ops;
    const Ctr* samples;
    const uint32_t* s_ops = nullptr;
    const Ctr* s_smp = nullptr;
    uint64_t so_lo = 0, so_hi = 0, sc_lo = 0, sc_hi = 0;
    RB_HD uint32_t op(uint64_t k) const { return (k - so_lo < so_hi - so_lo) ? s_ops[k - so_lo] : ops[k]; }
    RB_HD const Ctr& ent(uint64_t c, uint32_t s) const {  // entry s of chunk c (0 = absolute sample, 1.. = sub-samples)
        return (c - sc_lo < sc_hi - sc_lo) ? s_smp[(c - sc_lo) * SUBS + s] : samples[c * SUBS + s];
    }
    // the SUBS entries of chunk c are contiguous in either place: one range check for all of them
    RB_HD const Ctr* chunk(uint64_t c) const { return &ent(c, 0); }
    // ops [k, k + n) as one pointer when the whole run is staged, else nullptr (callers then go through op())
    RB_HD const uint32_t* op_run(uint64_t k, uint32_t n) const {
        return (k - so_lo < so_hi - so_lo && so_hi - k >= n) ? s_ops + (k - so_lo) : nullptr;
    }
    RB_HD Ctr smp(uint64_t c) const { return ent(c, 0); }
    RB_HD uint32_t smp_T(uint64_t c) const { return ent(c, 0).T; }
    // target bases of the record before op 32c + 8s (s >= 1; the position must lie inside the record)
    RB_HD uint32_t sub_T(uint64_t c, uint32_t s) const {
        const Ctr& e = ent(c, s);
        return (e.aux & SUB_ABS) ? e.T : e.T + ent(c, 0).T;
    }
    // counters of the record before op 32c + 8s (the position must lie inside the record, after its first op)
    // k_samples leaves the sub-samples out when windows are wide (one boundary per ~100 ops makes them dead weight: three
    // quarters of its stores) and says so in the chunk's absolute sample: SUB_ABS set in entry 0 = "no entries 1..3 here"
    RB_HD bool no_subs(uint64_t c) const { return (ent(c, 0).aux & SUB_ABS) != 0; }
    RB_HD Ctr at(uint64_t c, uint32_t s) const {
        Ctr e = ent(c, s);
        if (s == 0) { e.aux &= ~SUB_ABS; return e; }
        const bool abs = (e.aux & SUB_ABS) != 0;
        e.aux = abs ? (e.aux & ~SUB_ABS) : 0u;  // absolute entries of k_tok_scan carry the overflow bit and the slow-op count themselves
        if (!abs) ctr_add(e, ent(c, 0));
        return e;
    }
};

enum : int { POLICY_RIGHTMOST = 0, POLICY_EARLY_EXIT = 1 };
enum : uint32_t { LIFT_OK = 0, LIFT_ERR_NOT_FOUND = 1 };  // NOT_FOUND == the reference's "Problem getting index in cigar" panic

// Counters accumulated from the record's first op up to (excluding) op k (op_first <= k < op_end).
RB_HD Ctr ctr_before(const OpsView& v, const RecInfo& r, uint64_t k, ClassAcc& acc) {
    uint64_t base = (k >> SUB_LOG2) << SUB_LOG2;
    const uint64_t cbase = (k >> SAMPLE_LOG2) << SAMPLE_LOG2;
    if (base > cbase && base > r.op_first && v.no_subs(k >> SAMPLE_LOG2)) base = cbase;  // the chunk carries its absolute sample only
    Ctr c;
    uint64_t j;
    if (base > r.op_first) { c = v.at(k >> SAMPLE_LOG2, (uint32_t)(base & (SAMPLE - 1)) >> SUB_LOG2); j = base; }
    else { c = ctr_zero(); j = r.op_first; }
    acc_reset(acc);
    const uint64_t j0 = j;
    for (; j < k; j++) acc_add_op(acc, v.op(j));
    if (acc.big >= ACC_BIG) {  // a class sum might have wrapped: exact (slow) accumulation
        for (j = j0; j < k; j++) ctr_add_op(c, v.op(j));
        return c;
    }
    acc_flush(acc, c);
    return c;
}

// Warp reconvergence point (device only).  lift_pair is written in stages without early returns so that
// every lane executes the same sequence of these barriers: lanes leave the walk loops at different
// times, and without them the straight-line code after a loop runs once per straggler group.
#if defined(__CUDA_ARCH__)
#define RB_CONVERGE() __syncwarp()
#else
#define RB_CONVERGE() ((void)0)
#endif

// The 32-op chunk whose sampled target prefix is the last one <= p (the chunk find_op starts its walk in).
RB_HD uint64_t chunk_of(const OpsView& v, const RecInfo& r, uint32_t p) {
    uint64_t lo = r.op_first >> SAMPLE_LOG2, hi = (r.op_end - 1) >> SAMPLE_LOG2;
    // narrow to the staged chunk window when the answer provably lies inside it (all probes then hit smem)
    if (v.sc_hi > v.sc_lo) {
        const uint64_t a = v.sc_lo, z = v.sc_hi - 1;
        if (a > lo && a <= hi && v.smp_T(a) <= p) lo = a;
        if (z > lo && z <= hi && v.smp_T(z) > p) hi = z - 1;
    }
    if (lo >= v.sc_lo && hi < v.sc_hi) {  // the whole search range is staged: 32-bit indices, no range checks
        uint32_t a = (uint32_t)(lo - v.sc_lo), z = (uint32_t)(hi - v.sc_lo);
        while (a < z) {
            const uint32_t mid = (a + z + 1) >> 1;
            if (v.s_smp[mid * SUBS].T <= p) a = mid;
            else z = mid - 1;
        }
        return v.sc_lo + a;
    }
    // (An interpolation step in front of this bisection — guess the chunk from the position's share of the record's span, bracket,
    // then 3 probes — was measured and made k_lift 38 % SLOWER at 10 kb windows: the upper levels of the bisection are the same
    // for every pair of a record and hit L1, only the last few probes travel; the guess costs a 64-bit division and four loads.
    // Staging just the T word of every chunk a wide-window block spans (4 B per 32 ops) and bisecting in shared memory changed nothing
    // either, 1.04 against 1.02 ms at 8 haplotypes x 10 kb: what the pairs wait for is the chunk's entries, its ops and the
    // neighbours the slide rules look at — one or two dependent trips each — not the bisection.)
    while (lo < hi) {  // largest chunk whose starting T is <= p (chunk `lo` always qualifies)
        const uint64_t mid = (lo + hi + 1) >> 1;
        if (v.smp_T(mid) <= p) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// The reference-consuming op i (len > 0) with T_i <= p < T_i + L_i ; o = p - T_i ; `before` = counters before op i.
// `live` == false: does nothing (keeps the lane in step with its warp).
// FAST (k_emit's fused path): the record is known to hold no op of ACC_BIG bases or more (such records are RF_SLOW), so the
// class sums of a short walk cannot wrap and the exact re-accumulation is compiled out.
template <bool FAST = false>
RB_HD bool find_op(const OpsView& v, const RecInfo& r, bool live, uint32_t p, uint64_t& i, uint32_t& o, Ctr& before, ClassAcc& acc) {
    bool found = false;
    Ctr c = ctr_zero();
    uint64_t k0 = 0;
    uint32_t n = 0, j = 0;
    if (live && r.op_end > r.op_first) {
        const uint64_t lo = chunk_of(v, r, p);
        k0 = lo << SAMPLE_LOG2;
        const Ctr* e = v.chunk(lo);  // entry 0 = absolute sample, 1.. = sub-samples (relative unless SUB_ABS)
        const uint32_t T0 = e[0].T;
        uint32_t s = 0;  // last sub-sample of the chunk that lies inside the record and whose target prefix is <= p
        for (uint32_t t = (e[0].aux & SUB_ABS) ? 0u : SUBS - 1; t >= 1; t--) {  // (SUB_ABS in entry 0: the chunk has none)
            const uint64_t pos = k0 + t * SUB_OPS;
            if (pos > r.op_first && pos < r.op_end && ((e[t].aux & SUB_ABS) ? e[t].T : e[t].T + T0) <= p) { s = t; break; }
        }
        if (s) {
            k0 += s * SUB_OPS;
            c = e[s];
            const bool abs = (c.aux & SUB_ABS) != 0;
            c.aux = 0;
            if (!abs) ctr_add(c, e[0]);
        }
        else if (k0 > r.op_first) { c = e[0]; c.aux &= ~SUB_ABS; }
        else k0 = r.op_first;
        const uint64_t left = r.op_end - k0;
        n = left > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)left;
    }
    RB_CONVERGE();
    const uint32_t rel = p - c.T;  // target offset relative to the chunk start
    acc_reset(acc);
    const uint32_t near = n < SUB_OPS ? n : SUB_OPS;  // the op is normally within the sub-sample block
    const uint32_t* run = near ? v.op_run(k0, near) : nullptr;
    if (run) {  // staged run: plain shared-memory reads
        for (; j < near; j++) {
            const uint32_t w = run[j];
            const uint32_t L = op_len(w);
            if (is_ref(op_code(w)) && L > 0 && rel - acc.T < L) { found = true; break; }
            acc_add_op(acc, w);
        }
    }
    for (; !found && j < n; j++) {
        const uint32_t w = v.op(k0 + j);
        const uint32_t L = op_len(w);
        if (is_ref(op_code(w)) && L > 0 && rel - acc.T < L) { found = true; break; }
        acc_add_op(acc, w);
    }
    RB_CONVERGE();
    if (found) {
        if (!FAST && acc.big >= ACC_BIG) {  // a class sum might have wrapped: exact (slow) accumulation
            for (uint32_t t = 0; t < j; t++) ctr_add_op(c, v.op(k0 + t));
        } else {
            acc_flush(acc, c);
        }
        i = k0 + j; o = p - c.T; before = c;
    }
    return found;
}

// core::slice::binary_search of Rust 1.52..=1.81 over a column array whose entries equal to the
// target are exactly columns [ca, cb]: returns the first probe that compares Equal.
RB_HD uint32_t early_exit_probe(uint32_t n, uint32_t ca, uint32_t cb) {
    uint64_t size = n, left = 0, right = n;
    while (left < right) {
        const uint64_t mid = left + size / 2;
        if (mid < ca) left = mid + 1;
        else if (mid > cb) right = mid;
        else return (uint32_t)mid;
        size = right - left;
    }
    return ca;  // unreachable when ca <= cb < n
}

// Merge walk over the trimmed op range (collapse_long_cigar semantics, paf.rs:602-620): zero-length
// ops vanish, adjacent same-class ops fuse.  `emit(len, code)` is called once per printed op.
template <class Emit>
RB_HD void merged_walk(const OpsView& v, uint64_t si, uint64_t ei, uint32_t s_len, uint32_t e_len, Emit&& emit) {
    uint32_t prev = 0xFFFFFFFFu, run = 0;
    for (uint64_t k = si; k <= ei; k++) {
        const uint32_t w = v.op(k);
        const uint32_t code = op_code(w);
        const uint32_t L = (k == si) ? s_len : (k == ei ? e_len : op_len(w));
        if (L == 0) continue;
        if (code == prev) run += L;
        else {
            if (prev != 0xFFFFFFFFu) emit(run, prev);
            prev = code;
            run = L;
        }
    }
    if (prev != 0xFFFFFFFFu) emit(run, prev);
}

RB_HD void fill_stats(PairRes& out, const Ctr& d) {
    out.equal = d.EQ; out.diff = d.X + d.M; out.matches = d.M;
    out.ins = d.I; out.del = d.D; out.ins_ev = d.IEV; out.del_ev = d.DEV;
}

// ---- the three stages of one pair ----------------------------------------------------------------
// Split so that two drivers can share them: lift_pair (search per pair: k_lift, general path) and the
// streaming driver of stream_core.cuh (k_scan_lift: every 32-op chunk resolves the window boundaries
// that fall into it while it scans, and leaves one half-result per boundary).

// START: tpos_to_idx_match(t_st, search_right = true) given the found op (i, o, counters before i).
// Returns false when the start index would be the number of columns (liftover.rs:52-54 -> None).
RB_HD bool lift_start(const OpsView& v, uint64_t eo1, uint32_t a_lead, uint32_t tot_A, int policy, uint64_t i, uint32_t o,
                      const Ctr& before, uint64_t& si, uint32_t& so, Ctr& cs) {
    const uint32_t w = v.op(i);
    const uint32_t L = op_len(w), code = op_code(w);
    bool slide = false;
    if (o == L - 1) {  // the right-most column holding this target position may be an insertion column
        uint64_t k2 = i + 1;
        while (k2 < eo1 && op_len(v.op(k2)) == 0) k2++;
        if (k2 < eo1 && !is_ref(op_code(v.op(k2)))) slide = true;
    }
    if (slide && policy == POLICY_EARLY_EXIT && is_match(code)) {
        const uint32_t ca = before.A + o - a_lead;
        uint32_t extra = 0;
        for (uint64_t k = i + 1; k < eo1; k++) {
            const uint32_t w2 = v.op(k);
            if (op_len(w2) == 0) continue;
            if (is_ref(op_code(w2))) break;
            extra += op_len(w2);
        }
        if (early_exit_probe(tot_A, ca, ca + extra) == ca) slide = false;
    }
    if (!slide && is_match(code)) {
        si = i; so = o; cs = before;
        ctr_add_bases(cs, code, o);
        return true;
    }
    Ctr c = before;
    ctr_add_op(c, w);
    uint64_t k = i + 1;
    bool found = false;
    for (; k < eo1; k++) {
        const uint32_t w2 = v.op(k);
        if (is_match(op_code(w2)) && op_len(w2) > 0) { found = true; break; }
        ctr_add_op(c, w2);
    }
    si = k; so = 0; cs = c;
    return found;  // not found: start index == number of columns > any end index
}

// END: tpos_to_idx_match(t_en - 1, search_right = false) given the found op.  Returns false when the
// search slid to column 0 without meeting a match column (start > end -> None).
RB_HD bool lift_end(const OpsView& v, uint64_t eo0, uint64_t i, uint32_t o, const Ctr& before, uint64_t& ei, uint32_t& eo, Ctr& ce,
                    uint32_t& txt_before_ei) {
    const uint32_t w = v.op(i);
    const uint32_t code = op_code(w);
    if (is_match(code)) {
        ei = i; eo = o; ce = before; txt_before_ei = before.TXT;
        ctr_add_bases(ce, code, o + 1);
        return true;
    }
    Ctr c = before;
    uint64_t k = i;
    bool found = false;
    while (k > eo0) {
        k--;
        const uint32_t w2 = v.op(k);
        ctr_sub_op(c, w2);
        if (is_match(op_code(w2)) && op_len(w2) > 0) { found = true; break; }
    }
    ei = k; eo = found ? op_len(v.op(k)) - 1 : 0; ce = c; txt_before_ei = c.TXT;
    if (found) ctr_add_bases(ce, op_code(v.op(k)), eo + 1);
    return found;
}

// FINISH: coordinates, nmatch / aln_len, fused stats and the trimmed CIGAR's byte count from the two ends.
// L_si = length of op si.  Leaves out.kind == PK_DROP when start column > end column (Q7).
template <bool FAST = false>
RB_HD void lift_finish(const OpsView& v, const RecInfo& r, uint64_t si, uint32_t so, const Ctr& cs, uint32_t L_si, uint64_t ei,
                       uint32_t eo, const Ctr& ce, uint32_t txt_before_ei, PairRes& out) {
    if (cs.A >= ce.A) return;  // start column > end column: window lies inside an indel (Q7)
    out.kind = PK_TRIM;
    out.t_st = r.t_st + cs.T;
    out.t_en = r.t_st + ce.T;
    if (r.flags & RF_MINUS) {  // Q8: columns walk the query downward from q_en
        out.q_st = r.q_en0 - ce.Q;
        out.q_en = r.q_en0 - cs.Q;
    } else {
        out.q_st = r.q_st0 + cs.Q;
        out.q_en = r.q_st0 + ce.Q;
    }
    Ctr d = ce;
    ctr_sub(d, cs);
    out.nmatch = d.EQ + d.X + d.M;
    out.aln_len = d.A;
    fill_stats(out, d);
    out.si = si; out.ei = ei;
    if (si == ei) {
        out.s_len = eo - so + 1; out.e_len = 0;
        out.cg_bytes = ndigits32(out.s_len) + 1;
    } else {
        out.s_len = L_si - so; out.e_len = eo + 1;
        const uint32_t first = ndigits32(L_si) + 1;  // text bytes of op si as the input spells it (canonical)
        out.mid_len = txt_before_ei - cs.TXT - first;
        out.mid_off = r.text_off + cs.TXT + first;
        out.cg_bytes = ndigits32(out.s_len) + 1 + out.mid_len + ndigits32(out.e_len) + 1;
    }
    if (!FAST && (r.flags & RF_SLOW)) {  // Q15: re-collapse (rare: zero-length or adjacent same-class ops in the input)
        uint32_t bytes = 0, iev = 0, dev = 0;
        merged_walk(v, si, ei, out.s_len, out.e_len, [&](uint32_t len, uint32_t c2) {
            bytes += ndigits32(len) + 1;
            iev += (c2 == OP_I);
            dev += (c2 == OP_D);
        });
        out.cg_bytes = bytes; out.ins_ev = iev; out.del_ev = dev;
    }
}

RB_HD void pair_clear(PairRes& out) {
    out.mid_len = 0; out.mid_off = 0;
    out.kind = PK_DROP;
    out.t_st = out.t_en = out.q_st = out.q_en = out.nmatch = out.aln_len = out.si = out.ei = 0;
    out.s_len = out.e_len = out.cg_bytes = 0;
    out.equal = out.diff = out.ins = out.del = out.ins_ev = out.del_ev = out.matches = 0;
}
// liftover.rs:22-25 (Q3): record strictly inside the window -> the record itself, uncollapsed
RB_HD void pair_early(const RecInfo& r, PairRes& out) {
    out.kind = PK_EARLY;
    out.t_st = r.t_st; out.t_en = r.t_en; out.q_st = r.q_st; out.q_en = r.q_en;
    out.nmatch = r.tot.EQ + r.tot.X + r.tot.M;
    out.aln_len = r.tot.A;
    out.si = r.eo0; out.ei = r.eo1 - 1;
    out.cg_bytes = r.tot.TXT;
    out.mid_len = r.tot.TXT; out.mid_off = r.text_off + r.lead_txt;
    fill_stats(out, r.tot);
}

// One pair, searching for both ends.  Returns LIFT_OK (out.kind says DROP / TRIM / EARLY) or LIFT_ERR_NOT_FOUND.
// Staged, no early returns (see RB_CONVERGE): a lane that is done just stops being `live`.
// `enabled` == false (a candidate pair that does not overlap): produces PK_DROP, keeps the lane in step.
RB_HD uint32_t lift_pair(const OpsView& v, const RecInfo& r, uint64_t w_st, uint64_t w_en, int policy, bool enabled, PairRes& out,
                         ClassAcc& acc) {
    uint32_t status = LIFT_OK;
    bool live = enabled;
    pair_clear(out);
    if (live && r.t_st > w_st && r.t_en < w_en) {
        pair_early(r, out);
        live = false;
    } else if (live && r.t_en <= r.t_st) {
        status = LIFT_ERR_NOT_FOUND;
        live = false;
    }
    const uint32_t ps = live ? (uint32_t)((w_st > r.t_st ? w_st : r.t_st) - r.t_st) : 0u;      // liftover.rs:28
    const uint32_t pe = live ? (uint32_t)((w_en < r.t_en ? w_en : r.t_en) - 1 - r.t_st) : 0u;  // liftover.rs:38-40

    uint64_t i = 0; uint32_t o = 0; Ctr before = ctr_zero();
    if (!find_op(v, r, live, ps, i, o, before, acc) && live) { status = LIFT_ERR_NOT_FOUND; live = false; }
    uint64_t si = 0; uint32_t so = 0; Ctr cs = ctr_zero();
    if (live && !lift_start(v, r.eo1, r.a_lead, r.tot.A, policy, i, o, before, si, so, cs)) live = false;
    RB_CONVERGE();

    if (!find_op(v, r, live, pe, i, o, before, acc) && live) { status = LIFT_ERR_NOT_FOUND; live = false; }
    if (live) {
        uint64_t ei; uint32_t eo; Ctr ce; uint32_t txt_before_ei;
        if (lift_end(v, r.eo0, i, o, before, ei, eo, ce, txt_before_ei))
            lift_finish(v, r, si, so, cs, op_len(v.op(si)), ei, eo, ce, txt_before_ei, out);
    }
    RB_CONVERGE();
    return status;
}

// ---- half results of the streaming driver (one per window boundary, 64 B each) ----
struct HalfS {       // start of the trimmed alignment
    Ctr c;           // counters before the start column; c.aux = 1 if a start column exists
    uint32_t so;     // offset of the start column in op si
    uint32_t L_si;   // length of op si
    uint64_t si;
};
struct HalfE {       // end of the trimmed alignment
    Ctr c;           // counters up to and including the end column; c.aux = 1 if an end column exists
    uint32_t eo;     // offset of the end column in op ei
    uint32_t txt_before_ei;
    uint64_t ei;
};

// Pair = (record, window) from its two halves.  Same result as lift_pair for every pair whose two
// boundaries were resolved (all overlapping pairs of a record that passes check_integrity).
RB_HD uint32_t combine_pair(const OpsView& v, const RecInfo& r, uint64_t w_st, uint64_t w_en, const HalfS& hs, const HalfE& he,
                            PairRes& out) {
    pair_clear(out);
    if (r.t_st > w_st && r.t_en < w_en) { pair_early(r, out); return LIFT_OK; }
    if (r.t_en <= r.t_st) return LIFT_ERR_NOT_FOUND;
    // (the index guard only matters for unwritten halves of a record that fails check_integrity: no wild reads)
    if (hs.c.aux == 1u && he.c.aux == 1u && hs.si >= r.eo0 && he.ei < r.eo1 && hs.si <= he.ei)
        lift_finish(v, r, hs.si, hs.so, hs.c, hs.L_si, he.ei, he.eo, he.c, he.txt_before_ei, out);
    return LIFT_OK;
}

// Rows whose untouched ops are a long verbatim run of the input text (whole-record rows of trim-paf, early rows and wide
// windows over long records): when the rows of a call are few and long, k_serialise leaves that run to k_copy_mid, which
// spreads it over the grid instead of one warp per row.  Both kernels decide with this predicate.
constexpr uint32_t MID_BIG = 8192;
RB_HD bool mid_is_big(const RecInfo& r, const PairRes& p) {
    return (r.flags & RF_CANON) && p.mid_len >= MID_BIG &&
           (p.kind == PK_EARLY || (p.kind == PK_TRIM && p.ei > p.si && !(r.flags & RF_SLOW)));
}

// Bytes of the printed PAF line (paf.rs:923-943), '\n' included = a part that depends on the record only ...
RB_HD uint32_t line_const_bytes(const RecInfo& r, uint32_t q_name_len, uint32_t t_name_len) {
    return q_name_len + t_name_len + ndigits64(r.q_len) + 1 /*strand*/ + ndigits64(r.t_len) + ndigits64(r.mapq) +
           11 /*tabs between the 12 columns*/ + 6 /*\tid:Z:*/ + 6 /*\tcg:Z:*/ + 1 /*\n*/;
}
// ... plus the lifted coordinates, nmatch / aln_len, the id and the CIGAR
RB_HD uint32_t line_var_bytes(const PairRes& p, uint32_t id_len) {
    return id_len + ndigits64(p.q_st) + ndigits64(p.q_en) + ndigits64(p.t_st) + ndigits64(p.t_en) + ndigits32(p.nmatch) +
           ndigits32(p.aln_len) + p.cg_bytes;
}
RB_HD uint32_t line_bytes(const RecInfo& r, const PairRes& p, uint32_t q_name_len, uint32_t t_name_len, uint32_t id_len) {
    return line_const_bytes(r, q_name_len, t_name_len) + line_var_bytes(p, id_len);
}

}  // namespace rb
