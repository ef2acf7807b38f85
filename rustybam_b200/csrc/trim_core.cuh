// trim_core.cuh — run-length ("closed form") statement of `rb trim-paf`'s per-pair work.
//
// Replaces, per pair of query-overlapping records, the reference's per-base machinery:
//   trim_overlap.rs:6-20    score_of_qpos            (binary search over 8 B/column per overlap base, per record)
//   trim_overlap.rs:36-86   trim_overlapping_pafs    (two per-base score vectors, two cumulative passes, arg-max)
//   paf.rs:564-591          qpos_to_idx / qpos_to_idx_match
//   paf.rs:785-823          truncate_record_by_query (subset_cigar + collapse_long_cigar + check_integrity)
// with per-op prefix arrays in QUERY space (qp = query bases before the op, wp = score of the query positions before
// the op; 12 B per op, built once per call by k_trim_scan) and searches over them:
//   * the score of a query position depends only on the op that owns it, except for the LAST position of a
//     query-consuming op that is followed by columns that do not advance the query (D, N, H, P): all of them carry
//     the same query position, binary_search returns the right-most one (Rust < 1.52 / >= 1.82), so that
//     position scores as the last such column;
//   * cumulative-left + cumulative-right (trim_overlap.rs:60-69) is, up to a constant, a piecewise linear function of
//     the split point whose break points are op boundaries of either record: the first arg-max lies on one of them;
//   * truncation = two column look-ups + the slide to the nearest M/=/X column, all in (op, offset) space; a record
//     that has been truncated is a VIEW (first / last column) on its unchanged ops, so later rounds reuse the arrays.
// Both duplicate policies of core::slice::binary_search are implemented; the early-exit one makes the score
// prefix a function of the record's current truncation (see trim_probe_op).
//
// __host__ __device__: fuzzed on the CPU against the literal per-base oracle (tests/native/trim_core_check.cpp) and
// run by k_trim_pairs / k_trim_rows on the GPU.
#pragma once
#include "lift_core.cuh"

namespace rb {

struct TrimScores { int32_t match, diff, indel; };  // trim_overlap.rs:8-10 (CLI defaults 1 / 1 / 1)
struct TrimArr {
    const uint32_t* qp;    // [global op] query bases of the record before this op (counted from op_first, like Ctr::Q)
    const long long* wp;   // [global op] score of the query positions owned by the effective ops before this op
    const uint32_t* ap = nullptr;  // [global op] alignment columns of the record before this op (early-exit policy only)
    int policy = POLICY_RIGHTMOST; // which duplicate core::slice::binary_search returns (in query space here)
};
struct TrimView {          // 64 bytes per record
    uint64_t si, ei;       // ops holding the first / last alignment column of the (possibly truncated) record
    uint32_t so, eo;       // offsets of those columns inside their ops
    uint64_t q_st, q_en;   // current query span
    long long w_tot;       // score of all query positions of the effective (stripped) record
    uint32_t x_end;        // query bases before op eo1 (end of the effective range in the qp frame)
    uint32_t trimmed;      // truncations applied so far
    uint32_t bad;          // effective record does not start and end on an M/=/X op of positive length: unsupported here
    uint32_t pad;
};
enum : uint32_t { TRIM_OK = 0, TRIM_ABORT = 1 };  // ABORT == the reference panics (check_integrity().unwrap(), paf.rs:822)
constexpr uint32_t TRIM_NO_TAIL = 0xFFu;

// trim_overlap.rs:14-19
RB_HD int32_t trim_col_score(uint32_t code, const TrimScores& sc) {
    return code == OP_EQ ? sc.match : ((code == OP_I || code == OP_D) ? -sc.indel : -sc.diff);
}
// The last op (len > 0) of the run of non-query-consuming ops that follows op k, looking no further than op `last`
// (inclusive); TRIM_NO_TAIL if the next op of positive length advances the query (or there is none).
RB_HD uint32_t trim_tail(const OpsView& v, uint64_t k, uint64_t last, uint64_t* tail_op = nullptr) {
    uint32_t tail = TRIM_NO_TAIL;
    for (uint64_t j = k + 1; j <= last; j++) {
        const uint32_t w = v.op(j);
        if (op_len(w) == 0) continue;
        if (is_qry(op_code(w))) break;
        tail = op_code(w);
        if (tail_op) *tail_op = j;
    }
    return tail;
}
// score of the query positions owned by op k of a record whose effective range ends at eo1 (exclusive)
RB_HD long long trim_w_op(const OpsView& v, uint64_t k, uint64_t eo1, const TrimScores& sc) {
    const uint32_t w = v.op(k), L = op_len(w), code = op_code(w);
    if (L == 0 || !is_qry(code)) return 0;
    const int32_t s = trim_col_score(code, sc);
    long long t = (long long)L * s;
    const uint32_t tail = trim_tail(v, k, eo1 - 1);
    if (tail != TRIM_NO_TAIL) t += trim_col_score(tail, sc) - s;
    return t;
}
// ---- early-exit search policy (Rust 1.52 ..= 1.81) in query space ------------------------------------------------------
// The last query position of a query-consuming op that is followed by non-query columns is held by columns [ca, cb] of the
// record AS IT IS NOW (a truncated record's arrays are rebuilt, paf.rs:807-812: column 0 is the view's first column and
// the column count n is the view's); binary_search returns the first probe that lands in that range, so which class the
// position scores as — the op's own or one of the tail's — depends on (n, ca, cb): early_exit_probe (lift_core.cuh) replays it.
// Consequence: wp is a function of the view.  Under this policy it is re-scanned for the two records of every trimmed pair
// (trim_w_op_view below, the tail limited to the view: no view-end correction is needed then), everything else is shared.
RB_HD uint32_t trim_view_c0(const TrimArr& a, const TrimView& tv) { return a.ap[tv.si] + tv.so; }            // first column of the view
RB_HD uint32_t trim_view_n(const TrimArr& a, const TrimView& tv) { return a.ap[tv.ei] + tv.eo + 1u - trim_view_c0(a, tv); }
// Column (as an op of the tail run + the query-consuming base op) the search returns for the last position of op k;
// `tail_end` = last op of the non-query run behind k inside the view (k itself if there is none).  Returns the op whose class
// the position scores as.
RB_HD uint64_t trim_probe_op(const OpsView& v, const TrimArr& a, const TrimView& tv, uint64_t k, uint64_t* tail_end = nullptr) {
    uint32_t tail_cols = 0;
    uint64_t te = k;
    for (uint64_t j = k + 1; j <= tv.ei; j++) {
        const uint32_t w = v.op(j);
        if (op_len(w) == 0) continue;
        if (is_qry(op_code(w))) break;
        tail_cols += op_len(w);
        te = j;
    }
    if (tail_end) *tail_end = te;
    if (tail_cols == 0) return k;
    if (a.policy != POLICY_EARLY_EXIT) return te;  // right-most duplicate: the end of the run
    const uint32_t ca = a.ap[k] + op_len(v.op(k)) - 1u - trim_view_c0(a, tv), cb = ca + tail_cols;
    uint32_t pr = early_exit_probe(trim_view_n(a, tv), ca, cb);
    if (pr == ca) return k;
    pr -= ca;  // 1-based column inside the tail run
    for (uint64_t j = k + 1; j <= te; j++) {
        const uint32_t w = v.op(j);
        if (op_len(w) == 0) continue;
        if (pr <= op_len(w)) return j;
        pr -= op_len(w);
    }
    return te;
}
// score of the query positions owned by op k inside view tv (the view's last op counts in full: positions behind the view's
// last column are never asked for)
RB_HD long long trim_w_op_view(const OpsView& v, const TrimArr& a, const TrimView& tv, uint64_t k, const TrimScores& sc) {
    const uint32_t w = v.op(k), L = op_len(w), code = op_code(w);
    if (L == 0 || !is_qry(code)) return 0;
    const int32_t s = trim_col_score(code, sc);
    long long t = (long long)L * s;
    if (k >= tv.si && k < tv.ei) {
        const uint64_t j = trim_probe_op(v, a, tv, k);
        if (j != k) t += trim_col_score(op_code(v.op(j)), sc) - s;
    }
    return t;
}

// right-most op k in [lo, hi) with qp[k] <= x  (== the query-consuming op that owns query offset x)
RB_HD uint64_t trim_find_q(const TrimArr& a, uint64_t lo, uint64_t hi, uint32_t x) {
    hi--;
    while (lo < hi) {
        const uint64_t mid = (lo + hi + 1) >> 1;
        if (a.qp[mid] <= x) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
// query offset (qp frame) of absolute query position p
RB_HD uint32_t trim_x(const RecInfo& r, uint64_t p) {
    return (r.flags & RF_MINUS) ? (uint32_t)(r.q_en0 - 1 - p) : (uint32_t)(p - r.q_st0);
}
// score of the first x query positions in column order (x in the qp frame, >= qp[eo0])
RB_HD long long trim_G(const OpsView& v, const TrimArr& a, const RecInfo& r, const TrimView& tv, uint32_t x, const TrimScores& sc) {
    if (x >= tv.x_end) return tv.w_tot;
    const uint64_t k = trim_find_q(a, r.eo0, r.eo1, x);
    return a.wp[k] + (long long)(x - a.qp[k]) * trim_col_score(op_code(v.op(k)), sc);
}
// sum of score_of_qpos over the absolute query positions [qa, qb) of the record as it is now (view tv)
RB_HD long long trim_S(const OpsView& v, const TrimArr& a, const RecInfo& r, const TrimView& tv, uint64_t qa, uint64_t qb,
                       const TrimScores& sc) {
    if (qb <= qa) return 0;
    long long s;
    if (r.flags & RF_MINUS) s = trim_G(v, a, r, tv, (uint32_t)(r.q_en0 - qa), sc) - trim_G(v, a, r, tv, (uint32_t)(r.q_en0 - qb), sc);
    else s = trim_G(v, a, r, tv, (uint32_t)(qb - r.q_st0), sc) - trim_G(v, a, r, tv, (uint32_t)(qa - r.q_st0), sc);
    // the view's last column lost the non-query columns that followed it in the untruncated record
    // (early-exit policy: wp was scanned for this very view, nothing to correct)
    const uint32_t we = v.op(tv.ei);
    if (a.policy != POLICY_EARLY_EXIT && tv.eo == op_len(we) - 1u) {
        const uint32_t tail = trim_tail(v, tv.ei, r.eo1 - 1);
        if (tail != TRIM_NO_TAIL) {
            const uint32_t xe = a.qp[tv.ei] + tv.eo;
            const uint64_t pe = (r.flags & RF_MINUS) ? r.q_en0 - 1 - xe : r.q_st0 + xe;
            if (pe >= qa && pe < qb) s += trim_col_score(op_code(we), sc) - trim_col_score(tail, sc);
        }
    }
    return s;
}

struct TrimBest {       // arg-max state: larger total wins, ties go to the smaller split point
    long long total;
    uint64_t c;
};
RB_HD void trim_best_merge(TrimBest& b, long long total, uint64_t c) {
    if (total > b.total || (total == b.total && c < b.c)) { b.total = total; b.c = c; }
}
// total(c) - Rtot for a split at absolute query position c in [A, B]: score of the left record over [A, c) minus the
// score of the right record over [A, c)   (trim_overlap.rs:60-77 with the constant Rtot = right over [A, B) taken out)
RB_HD long long trim_val(const OpsView& v, const TrimArr& a, const RecInfo& rl, const TrimView& tl, const RecInfo& rr, const TrimView& tr,
                         uint64_t A, uint64_t c, const TrimScores& sc) {
    return trim_S(v, a, rl, tl, A, c, sc) - trim_S(v, a, rr, tr, A, c, sc);
}
// What the arg-max needs to know about one record of a pair, computed once per pair (and thread): the cumulative score at
// the overlap start, and the view-end correction of trim_S as a (position, delta) pair.
struct TrimSide {
    long long g_A;     // trim_G at the overlap start A (in the record's own column order)
    long long delta;   // 0, or what the view's last column gains by having lost its non-query tail
    uint64_t pe;       // absolute query position of that column
    uint32_t minus;
};
RB_HD TrimSide trim_side(const OpsView& v, const TrimArr& a, const RecInfo& r, const TrimView& tv, uint64_t A, const TrimScores& sc) {
    TrimSide s;
    s.minus = (r.flags & RF_MINUS) ? 1u : 0u;
    s.g_A = trim_G(v, a, r, tv, s.minus ? (uint32_t)(r.q_en0 - A) : (uint32_t)(A - r.q_st0), sc);
    s.delta = 0; s.pe = 0;
    const uint32_t we = v.op(tv.ei);
    if (a.policy != POLICY_EARLY_EXIT && tv.eo == op_len(we) - 1u) {
        const uint32_t tail = trim_tail(v, tv.ei, r.eo1 - 1);
        if (tail != TRIM_NO_TAIL) {
            const uint32_t xe = a.qp[tv.ei] + tv.eo;
            s.pe = s.minus ? r.q_en0 - 1 - xe : r.q_st0 + xe;
            s.delta = trim_col_score(op_code(we), sc) - trim_col_score(tail, sc);
        }
    }
    return s;
}
// trim_S over [A, c) from the cumulative score g_c at c (== trim_S(v, a, r, tv, A, c, sc) for A < c)
RB_HD long long trim_S_at(const TrimSide& s, uint64_t A, uint64_t c, long long g_c) {
    long long x = s.minus ? s.g_A - g_c : g_c - s.g_A;
    if (s.delta != 0 && s.pe >= A && s.pe < c) x += s.delta;
    return x;
}
// cumulative score of record r at absolute query position c (one search)
RB_HD long long trim_G_at(const OpsView& v, const TrimArr& a, const RecInfo& r, const TrimView& tv, uint64_t c, const TrimScores& sc) {
    return trim_G(v, a, r, tv, (r.flags & RF_MINUS) ? (uint32_t)(r.q_en0 - c) : (uint32_t)(c - r.q_st0), sc);
}
// Candidate split points contributed by record `rx` (`x_is_left`: which of the pair it is): the op boundaries (and the
// one-base segments in front of non-query runs) that fall into [A, B].  Work item t of n (a thread of the grid; 0 of 1 on
// the host).  The record's own cumulative score at its op boundaries comes straight from wp — only the OTHER record is
// searched, once per candidate.
RB_HD void trim_scan_candidates(const OpsView& v, const TrimArr& a, bool x_is_left, const RecInfo& rl, const TrimView& tl, const TrimSide& sl,
                                const RecInfo& rr, const TrimView& tr, const TrimSide& sr, uint64_t A, uint64_t B, const TrimScores& sc,
                                uint32_t t, uint32_t n, TrimBest& best) {
    const RecInfo& rx = x_is_left ? rl : rr;
    const RecInfo& ry = x_is_left ? rr : rl;
    const TrimView& tx = x_is_left ? tl : tr;
    const TrimView& ty = x_is_left ? tr : tl;
    const TrimSide& sx = x_is_left ? sl : sr;
    const TrimSide& sy = x_is_left ? sr : sl;
    uint64_t k0 = trim_find_q(a, rx.eo0, rx.eo1, trim_x(rx, A)), k1 = trim_find_q(a, rx.eo0, rx.eo1, trim_x(rx, B - 1));
    if (k0 > k1) { const uint64_t tmp = k0; k0 = k1; k1 = tmp; }
    const bool minus = (rx.flags & RF_MINUS) != 0;
    for (uint64_t k = k0 + t; k <= k1; k += n) {
        const uint32_t w = v.op(k), L = op_len(w);
        if (L == 0 || !is_qry(op_code(w))) continue;
        const long long s1 = trim_col_score(op_code(w), sc);
        const long long g0 = a.wp[k], gm = g0 + (long long)(L - 1u) * s1,
                        g1 = g0 + (a.policy == POLICY_EARLY_EXIT ? trim_w_op_view(v, a, tx, k, sc) : trim_w_op(v, k, rx.eo1, sc));
        uint64_t cand[3];
        long long gx[3];
        if (minus) {  // columns run against the query: the op's first column is its highest query position
            const uint64_t hi = rx.q_en0 - a.qp[k], lo = hi - L;
            cand[0] = hi; gx[0] = g0; cand[1] = lo + 1; gx[1] = gm; cand[2] = lo; gx[2] = g1;
        } else {
            const uint64_t lo = rx.q_st0 + a.qp[k], hi = lo + L;
            cand[0] = lo; gx[0] = g0; cand[1] = hi - 1; gx[1] = gm; cand[2] = hi; gx[2] = g1;
        }
        for (int i = 0; i < 3; i++) {
            const uint64_t c = cand[i];
            if (c <= A || c > B) continue;  // c == A is one of the fixed candidates (value 0)
            const long long own = trim_S_at(sx, A, c, gx[i]);
            const long long other = trim_S_at(sy, A, c, trim_G_at(v, a, ry, ty, c, sc));
            trim_best_merge(best, x_is_left ? own - other : other - own, c);
        }
    }
}
RB_HD void trim_fixed_candidates(const OpsView& v, const TrimArr& a, const RecInfo& rl, const TrimView& tl, const RecInfo& rr,
                                 const TrimView& tr, uint64_t A, uint64_t B, const TrimScores& sc, TrimBest& best) {
    const uint64_t cand[4] = {A, A + 1, B - 1, B};
    for (int i = 0; i < 4; i++)
        if (cand[i] >= A && cand[i] <= B) trim_best_merge(best, trim_val(v, a, rl, tl, rr, tr, A, cand[i], sc), cand[i]);
}
// The arg-max as one 64-bit key for atomicMax across the blocks of a pair: larger total first, then the smaller split point.
// |total| < 2^31: total = L[A,c) - R[A,c) is bounded by 2 x overlap x |score|, and k_trim_select refuses pairs with
// 2 x overlap x max|score| >= 2^31 (the reference's own i32 sums allow half of that) and
// c - A < 2^32 (per-record query sums are below 2^32).
RB_HD unsigned long long trim_key(const TrimBest& b, uint64_t A) {
    return ((unsigned long long)(b.total + 2147483648ll) << 32) | (unsigned long long)(~(uint32_t)(b.c - A));
}
RB_HD TrimBest trim_unkey(unsigned long long key, uint64_t A) {
    TrimBest b;
    b.total = (long long)(key >> 32) - 2147483648ll;
    b.c = A + (uint64_t)(~(uint32_t)key);
    return b;
}
// split point from the arg-max over all candidates (trim_overlap.rs:71-77: `if l + r > max` starting from max = 0)
RB_HD uint64_t trim_split(const TrimBest& best, long long r_tot, uint64_t A) { return (best.total + r_tot > 0) ? best.c : A; }

// ---- which pair of a query name is trimmed in a round (paf.rs:229-275) ---------------------------------------------
// The reference lists the overlapping pairs (i < j, file order within the name), sorts them by overlap (largest first,
// stable) and trims the first one of every query name; all the others wait for the next round.  "First after a stable sort"
// == largest overlap, ties to the pair generated first == one max over the key below.
enum : uint32_t { TRIM_PAIR_NONE = 0, TRIM_PAIR_J_CONTAINED = 1, TRIM_PAIR_I_CONTAINED = 2, TRIM_PAIR_PARTIAL = 3 };
RB_HD uint32_t trim_pair_class(uint64_t i_st, uint64_t i_en, uint64_t j_st, uint64_t j_en, uint64_t& overlap) {
    const uint64_t mn = i_en < j_en ? i_en : j_en, mx = i_st > j_st ? i_st : j_st;  // bed.rs:74-85
    overlap = mn < mx ? 0 : mn - mx;
    if (overlap < 1) return TRIM_PAIR_NONE;
    if (overlap == j_en - j_st) return TRIM_PAIR_J_CONTAINED;  // paf.rs:243-245 (checked first)
    if (overlap == i_en - i_st) return TRIM_PAIR_I_CONTAINED;
    return TRIM_PAIR_PARTIAL;
}
// i, j = positions inside the name's group of m records (m < 65536), overlap < 2^32
RB_HD unsigned long long trim_sel_key(uint64_t overlap, uint32_t i, uint32_t j, uint32_t m) {
    return ((unsigned long long)overlap << 32) | (unsigned long long)(~(i * m + j));
}
RB_HD void trim_sel_unkey(unsigned long long key, uint32_t m, uint32_t& i, uint32_t& j) {
    const uint32_t g = ~(uint32_t)key;
    i = g / m; j = g % m;
}
struct TrimPairSelDev { uint32_t left, right; uint64_t st_ovl, en_ovl; };  // == TrimPairSel of trim_rounds.hpp
constexpr uint32_t TRIM_SEL_NONE = 0xFFFFFFFFu;
// the selected pair of a group from the winning key (records lo + i, lo + j): left = the one that starts first (paf.rs:251-255)
RB_HD TrimPairSelDev trim_sel_make(const TrimView* views, uint32_t lo, uint32_t m, unsigned long long key) {
    uint32_t i, j;
    trim_sel_unkey(key, m, i, j);
    const TrimView &a = views[lo + i], &b = views[lo + j];
    TrimPairSelDev s;
    if (a.q_st <= b.q_st) { s.left = lo + i; s.right = lo + j; }
    else { s.left = lo + j; s.right = lo + i; }
    s.st_ovl = a.q_st > b.q_st ? a.q_st : b.q_st;
    s.en_ovl = a.q_en < b.q_en ? a.q_en : b.q_en;
    return s;
}
// sequential statement of one group's selection (the kernel spreads the same loop over a block and reduces with atomics)
RB_HD uint32_t trim_select_group(const TrimView* views, uint32_t lo, uint32_t hi, uint8_t* contained, TrimPairSelDev& sel) {
    const uint32_t m = hi - lo;
    unsigned long long best = 0;
    uint32_t n_pairs = 0;
    for (uint32_t r = lo; r < hi; r++) contained[r] = 0;
    for (uint32_t i = 0; i + 1 < m; i++)
        for (uint32_t j = i + 1; j < m; j++) {
            uint64_t ov;
            const uint32_t c = trim_pair_class(views[lo + i].q_st, views[lo + i].q_en, views[lo + j].q_st, views[lo + j].q_en, ov);
            if (c == TRIM_PAIR_J_CONTAINED) contained[lo + j] = 1;
            else if (c == TRIM_PAIR_I_CONTAINED) contained[lo + i] = 1;
            else if (c == TRIM_PAIR_PARTIAL) {
                n_pairs++;
                const unsigned long long k = trim_sel_key(ov, i, j, m);
                if (k > best) best = k;
            }
        }
    sel.left = TRIM_SEL_NONE; sel.right = 0; sel.st_ovl = sel.en_ovl = 0;
    if (n_pairs) sel = trim_sel_make(views, lo, m, best);
    return n_pairs;
}

// ---- truncate_record_by_query (paf.rs:785-823) in (op, offset) space ---------------------------------------------
struct TrimCol { uint64_t k; uint32_t o; };
RB_HD bool trim_col_lt(const TrimCol& x, const TrimCol& y) { return x.k < y.k || (x.k == y.k && x.o < y.o); }

// qpos_to_idx (right-most policy): `base` = the query-consuming column of position p, `last` = the right-most column
// that repeats that position (the end of the non-query run behind it, inside the view)
RB_HD void trim_col_of(const OpsView& v, const TrimArr& a, const RecInfo& r, const TrimView& tv, uint64_t p, TrimCol& base, TrimCol& last) {
    const uint32_t x = trim_x(r, p);
    base.k = trim_find_q(a, r.eo0, r.eo1, x);
    base.o = x - a.qp[base.k];
    last = base;
    if (base.o == op_len(v.op(base.k)) - 1u && base.k < tv.ei) {
        uint64_t j = 0;
        if (a.policy == POLICY_EARLY_EXIT) {
            // the probe either returns the query-consuming column itself or one inside the run: the slides only need to know which
            // (from any column of the run they pass the rest of it, no column of it is a match)
            uint64_t te = base.k;
            if (trim_probe_op(v, a, tv, base.k, &te) != base.k) { last.k = te; last.o = op_len(v.op(te)) - 1u; }
        } else if (trim_tail(v, base.k, tv.ei, &j) != TRIM_NO_TAIL) { last.k = j; last.o = op_len(v.op(j)) - 1u; }
    }
}
// `while idx < max_idx && !match { idx += 1 }` from column `last` (whose query-consuming column is `base`)
RB_HD bool trim_slide_right(const OpsView& v, const TrimView& tv, const TrimCol& base, const TrimCol& last, TrimCol& out) {
    if (last.k == base.k && is_match(op_code(v.op(base.k)))) { out = base; return true; }
    for (uint64_t j = last.k + 1; j <= tv.ei; j++) {
        const uint32_t w = v.op(j);
        if (op_len(w) > 0 && is_match(op_code(w))) { out.k = j; out.o = 0; return true; }
    }
    return false;  // idx == number of columns: the reference indexes out of bounds
}
// `while idx > 0 && !match { idx -= 1 }`
RB_HD void trim_slide_left(const OpsView& v, const TrimView& tv, const TrimCol& base, TrimCol& out) {
    if (is_match(op_code(v.op(base.k)))) { out = base; return; }
    for (uint64_t j = base.k; j > tv.si;) {
        j--;
        const uint32_t w = v.op(j);
        if (op_len(w) > 0 && is_match(op_code(w))) { out.k = j; out.o = op_len(w) - 1u; return; }
    }
    out.k = tv.si; out.o = tv.so;  // column 0 of the view
}
// Truncates the view to the query interval [nqs, nqe) (q_st <= nqs < nqe <= q_en).
RB_HD uint32_t trim_truncate(const OpsView& v, const TrimArr& a, const RecInfo& r, TrimView& tv, uint64_t nqs, uint64_t nqe) {
    TrimCol b0, l0, b1, l1, cs, ce;
    trim_col_of(v, a, r, tv, nqs, b0, l0);
    trim_col_of(v, a, r, tv, nqe - 1, b1, l1);
    if (r.flags & RF_MINUS) {  // paf.rs:582: the directions swap on '-'; columns run against the query
        trim_slide_left(v, tv, b0, ce);                       // aln_st (from new_q_st): the higher column
        if (!trim_slide_right(v, tv, b1, l1, cs)) return TRIM_ABORT;  // aln_en (from new_q_en - 1): the lower column
    } else {
        if (!trim_slide_right(v, tv, b0, l0, cs)) return TRIM_ABORT;
        trim_slide_left(v, tv, b1, ce);
    }
    if (trim_col_lt(ce, cs)) return TRIM_ABORT;  // crossed: spans and CIGAR disagree -> check_integrity().unwrap() panics
    tv.si = cs.k; tv.so = cs.o; tv.ei = ce.k; tv.eo = ce.o;
    const uint32_t xs = a.qp[cs.k] + cs.o, xe = a.qp[ce.k] + ce.o + 1u;
    if (r.flags & RF_MINUS) { tv.q_st = r.q_en0 - xe; tv.q_en = r.q_en0 - xs; }
    else { tv.q_st = r.q_st0 + xs; tv.q_en = r.q_st0 + xe; }
    tv.trimmed++;
    return TRIM_OK;
}

// The untruncated view of a record (after k_rec_prep's strip); w_tot / x_end are filled by the scan.
RB_HD void trim_view_init(const OpsView& v, const RecInfo& r, TrimView& tv) {
    const uint32_t w0 = v.op(r.eo0), w1 = v.op(r.eo1 - 1);
    tv.si = r.eo0; tv.so = 0;
    tv.ei = r.eo1 - 1; tv.eo = op_len(w1) - 1u;
    tv.q_st = r.q_st; tv.q_en = r.q_en;
    tv.trimmed = 0; tv.pad = 0;
    tv.bad = (op_len(w0) > 0 && is_match(op_code(w0)) && op_len(w1) > 0 && is_match(op_code(w1))) ? 0u : 1u;
}

// The printed row of one record after all rounds: untouched -> the (stripped) record itself, uncollapsed;
// truncated -> the view, re-collapsed (paf.rs:807-808), nmatch / aln_len from check_integrity (paf.rs:822).
RB_HD void trim_row(const OpsView& v, const RecInfo& r, const TrimView& tv, ClassAcc& acc, PairRes& out) {
    pair_clear(out);
    if (!tv.trimmed) { pair_early(r, out); return; }
    const uint32_t ws = v.op(tv.si), we = v.op(tv.ei);
    Ctr cs = ctr_before(v, r, tv.si, acc);
    ctr_add_bases(cs, op_code(ws), tv.so);
    Ctr ce = ctr_before(v, r, tv.ei, acc);
    const uint32_t txt_before_ei = ce.TXT;
    ctr_add_bases(ce, op_code(we), tv.eo + 1u);
    lift_finish(v, r, tv.si, tv.so, cs, op_len(ws), tv.ei, tv.eo, ce, txt_before_ei, out);
}

}  // namespace rb
