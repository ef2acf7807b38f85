// stream_core.cuh — streaming resolution of window boundaries inside one run of ops.
//
// The reference answers "which alignment column holds target position p" with a binary search
// over per-base arrays, twice per (window, record) pair (paf.rs:541-561, liftover.rs:26-49).  On
// the sorted-BED fast path both sides are monotone — the windows of a record by `st`/`en`, the
// ops by their target prefix sum T — so the two sequences are MERGED instead: while a thread of
// k_scan_lift walks its 32-op chunk for the prefix scan it also owns every boundary whose target
// position falls into the chunk:
//     start boundaries  PS_j = max(st_j, t_st) - t_st          in [T0, T0 + Tseg)      (liftover.rs:28)
//     end   boundaries  E_j  = min(en_j, t_en) - t_st  (= pe+1) in (T0, T0 + Tseg]      (liftover.rs:38-40)
// with T0 = target bases of the record before the segment.  Each boundary becomes one half result
// (HalfS / HalfE, lift_core.cuh); combine_pair() joins the two halves of a pair.
//
// __host__ __device__: fuzzed on the CPU against lift_pair and the literal oracle
// (tests/native/lift_core_check.cpp), runs inside k_scan_lift on the GPU.
#pragma once
#include "lift_core.cuh"

namespace rb {

// first index in [lo, hi) for which pred(index) is true (pred monotone false..true); hi if none
template <class Pred>
RB_HD uint32_t first_true(uint32_t lo, uint32_t hi, Pred&& pred) {
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (pred(mid)) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// Windows [wlo, whi) of one record as relative boundary positions (u32: a record's target span is < 2^32).
//   WA::ps(j)  = max(st_j, t_st) - t_st      non-decreasing in j
//   WA::en(j)  = min(en_j, t_en) - t_st      non-decreasing in j (fast path: en monotone)
struct WinGlobal {
    const uint64_t* st;
    const uint64_t* en;
    uint64_t t_st, t_en;
    RB_HD uint32_t ps(uint32_t j) const { const uint64_t s = st[j]; return (uint32_t)((s > t_st ? s : t_st) - t_st); }
    RB_HD uint32_t pe(uint32_t j) const { const uint64_t e = en[j]; return (uint32_t)((e < t_en ? e : t_en) - t_st); }
};

struct SegRec {          // what a segment needs to know about its record
    uint64_t eo0, eo1;   // stripped op range
    uint32_t wlo, whi;   // overlapping windows
    uint32_t s_lo, s_hi; // search range of the start boundaries ([wlo, whi), or a narrower range known to hold them)
    uint32_t e_lo, e_hi; // same for the end boundaries
    uint64_t pair0;      // index of pair (record, wlo) in emission order
};

// Walk ops [k0, k0 + n) — a run inside ONE record, `base` = the record's counters before op k0,
// Tseg = target bases of the run — and emit the halves of every boundary that falls into it.
//   seg_op(j)       op word k0 + j (the caller's staged copy)
//   v               global view for the look-ahead / look-behind of the slide rules
//   emit_s(pair, HalfS), emit_e(pair, HalfE)
template <class SegOp, class WA, class EmitS, class EmitE>
RB_HD void stream_segment(const OpsView& v, const SegRec& r, const WA& wa, uint64_t k0, uint32_t n, SegOp&& seg_op, const Ctr& base,
                          uint32_t Tseg, ClassAcc& acc, EmitS&& emit_s, EmitE&& emit_e) {
    if (Tseg == 0 || r.whi <= r.wlo) return;
    const uint32_t T0 = base.T, T1 = T0 + Tseg;
    uint32_t js = first_true(r.s_lo, r.s_hi, [&](uint32_t j) { return wa.ps(j) >= T0; });
    const uint32_t js_end = first_true(js, r.s_hi, [&](uint32_t j) { return wa.ps(j) >= T1; });
    uint32_t je = first_true(r.e_lo, r.e_hi, [&](uint32_t j) { return wa.pe(j) > T0; });
    const uint32_t je_end = first_true(je, r.e_hi, [&](uint32_t j) { return wa.pe(j) > T1; });
    if (js == js_end && je == je_end) return;
    uint32_t next_s = js < js_end ? wa.ps(js) : 0xFFFFFFFFu;
    uint32_t next_e = je < je_end ? wa.pe(je) : 0xFFFFFFFFu;
    bool more_s = js < js_end, more_e = je < je_end;
    acc_reset(acc);
    uint32_t Tcur = T0;
    for (uint32_t j = 0; j < n && (more_s || more_e); j++) {
        const uint32_t w = seg_op(j);
        const uint32_t L = op_len(w), code = op_code(w);
        if (is_ref(code) && L > 0) {
            const uint32_t Tend = Tcur + L;
            if ((more_s && next_s < Tend) || (more_e && next_e <= Tend)) {
                Ctr before = base;  // counters before op k0 + j
                if (acc.big >= ACC_BIG) { for (uint32_t t = 0; t < j; t++) ctr_add_op(before, seg_op(t)); }
                else acc_flush(acc, before);
                while (more_s && next_s < Tend) {  // start boundary inside this op: column offset o = PS - Tcur
                    HalfS h;
                    const bool ok = lift_start(v, r.eo1, 0u, 0u, POLICY_RIGHTMOST, k0 + j, next_s - Tcur, before, h.si, h.so, h.c);
                    h.c.aux = ok ? 1u : 0u;
                    h.L_si = ok ? op_len(v.op(h.si)) : 0u;
                    emit_s(r.pair0 + (js - r.wlo), h);
                    js++;
                    more_s = js < js_end;
                    if (more_s) next_s = wa.ps(js);
                }
                while (more_e && next_e <= Tend) {  // end boundary: pe = E - 1, column offset o = E - 1 - Tcur
                    HalfE h;
                    const bool ok = lift_end(v, r.eo0, k0 + j, next_e - 1u - Tcur, before, h.ei, h.eo, h.c, h.txt_before_ei);
                    h.c.aux = ok ? 1u : 0u;
                    emit_e(r.pair0 + (je - r.wlo), h);
                    je++;
                    more_e = je < je_end;
                    if (more_e) next_e = wa.pe(je);
                }
            }
            Tcur = Tend;
        }
        acc_add_op(acc, w);
    }
}

}  // namespace rb
