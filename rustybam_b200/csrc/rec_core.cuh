// rec_core.cuh — per-record preparation shared by the kernels and the CPU fuzz harness.
//
// Restates, at op granularity:
//   paf.rs:656-783  remove_trailing_indels  (called by aligned_pairs at paf.rs:503 on every record
//                   of every contig, liftover.rs:119-121) — including its panics
//   paf.rs:825-857  check_integrity         (span checks; nmatch / aln_len come from the counters)
//   paf.rs:622-627  paf_overlaps_rgn        (strict half-open, on the STRIPPED coordinates, Q6)
#pragma once
#include "rb_common.cuh"

namespace rb {

// error codes raised per record (first one wins); numbering shared with rbcuda.h rb_status (negated)
enum : uint32_t {
    RE_OK = 0,
    RE_CIGAR_PARSE = 4,     // rust-htslib CigarString::try_from error -> `expect` panic (paf.rs:399)
    RE_INTEGRITY = 5,       // check_integrity().unwrap() (paf.rs:70)
    RE_STRIP_PANIC = 6,     // remove_trailing_indels: leading deletion / all-indel / empty CIGAR (paf.rs:663,782)
    RE_INDEX_PANIC = 7,     // "Problem getting index in cigar" (liftover.rs:31-49)
    RE_UNSUPPORTED = 8,     // input outside the documented domain of the B200 path (op len >= 2^28, sums >= 2^32)
};

// Strip leading/trailing I/D ops.  In: r.op_first/op_end, r.t_st/t_en, r.q_st0/q_en0, RF_MINUS.
// Out: eo0/eo1, stripped coords, a_lead, n_lead/n_trail, id_len, RF_STRIPPED.
RB_HD uint32_t strip_record(const uint32_t* ops, RecInfo& r) {
    r.eo0 = r.op_first; r.eo1 = r.op_end;
    r.q_st = r.q_st0; r.q_en = r.q_en0;
    r.a_lead = 0; r.n_lead = 0; r.n_trail = 0; r.id_len = 0; r.lead_txt = 0;
    if (r.op_end <= r.op_first) return RE_STRIP_PANIC;  // self.cigar.first().unwrap() on an empty CIGAR
    uint32_t lead_i = 0, trail_i = 0, trail_d = 0, txt = 0;
    uint64_t k = r.op_first;
    for (; k < r.op_end; k++) {
        const uint32_t w = ops[k], code = op_code(w);
        if (code == OP_D) return RE_STRIP_PANIC;  // q is bumped by 1 per leading D (paf.rs:673) -> integrity panic (Q9)
        if (code != OP_I) break;
        lead_i += op_len(w);
        txt += ndigits32(op_len(w)) + 1;
        r.n_lead++;
    }
    if (k == r.op_end) return RE_STRIP_PANIC;  // nothing but indels: both ends strip everything, spans break
    r.lead_txt = txt;
    uint64_t k2 = r.op_end;
    while (k2 > k) {
        const uint32_t w = ops[k2 - 1], code = op_code(w);
        if (code == OP_D) trail_d += op_len(w);
        else if (code == OP_I) trail_i += op_len(w);
        else break;
        txt += ndigits32(op_len(w)) + 1;
        r.n_trail++;
        k2--;
    }
    r.eo0 = k; r.eo1 = k2;
    r.a_lead = lead_i;
    r.t_en -= trail_d;
    if (r.flags & RF_MINUS) {  // paf.rs:764-766: the query fix-ups swap on '-'
        r.q_st = r.q_st0 + trail_i;
        r.q_en = r.q_en0 - lead_i;
    } else {
        r.q_st = r.q_st0 + lead_i;
        r.q_en = r.q_en0 - trail_i;
    }
    if (r.n_lead | r.n_trail) {
        r.flags |= RF_STRIPPED;
        r.id_len = 4 + txt + 1;  // "_TO." lead-ops "." trail-ops
    }
    return RE_OK;
}

}  // namespace rb
