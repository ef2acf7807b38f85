// rb_common.cuh — shared types of the B200 liftover/stats path (device + host-side unit tests).
//
// Data layout in HBM:
//   text      u8[N]            concatenated cg:Z: payloads of all records (read once by the tokeniser)
//   ops       u32[n_ops]       one word per CIGAR op: (len << 4) | code, BAM op codes, len < 2^28
//   heads     u32[n_ops/32+1]  bit k%32 of word k/32 set  <=>  op k is the first op of a record
//   samples   Ctr[n_ops/32+1]  sampled segmented prefix sums: samples[c] = counters accumulated from
//                              the first op of the record containing op 32c up to (excluding) op 32c
//   recs      RecInfo[n_rec]   per-record state after the leading/trailing indel strip
// Nothing per alignment column is ever materialised (the reference keeps 24 B per column,
// paf.rs:362-364,501-538); per-op state is 4 B + 1.5 B of samples.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline
#endif

namespace rb {

// BAM op codes; text alphabet "MIDNSHP=X" (rust-htslib Cigar enum order, paf.rs:946-975 classes)
enum : uint32_t { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
constexpr uint32_t MASK_REF = (1u << OP_M) | (1u << OP_D) | (1u << OP_N) | (1u << OP_EQ) | (1u << OP_X);  // paf.rs:946-951
constexpr uint32_t MASK_QRY = (1u << OP_M) | (1u << OP_I) | (1u << OP_S) | (1u << OP_EQ) | (1u << OP_X);  // paf.rs:958-963
constexpr uint32_t MASK_MATCH = (1u << OP_M) | (1u << OP_EQ) | (1u << OP_X);                              // paf.rs:973-975
constexpr uint32_t MAX_OP_LEN = (1u << 28) - 1;
constexpr int SAMPLE_LOG2 = 5;
constexpr uint32_t SAMPLE = 1u << SAMPLE_LOG2;  // ops per sample chunk (== bits of a heads word)

RB_HD uint32_t op_len(uint32_t w) { return w >> 4; }
RB_HD uint32_t op_code(uint32_t w) { return w & 15u; }
RB_HD bool is_ref(uint32_t code) { return (MASK_REF >> code) & 1u; }
RB_HD bool is_qry(uint32_t code) { return (MASK_QRY >> code) & 1u; }
RB_HD bool is_match(uint32_t code) { return (MASK_MATCH >> code) & 1u; }

RB_HD uint32_t ndigits32(uint32_t v) {
    return 1u + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) +
           (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
}
RB_HD uint32_t ndigits64(uint64_t v) {
    if (v < 4294967296ull) return ndigits32((uint32_t)v);
    return 10u + (v >= 10000000000ull) + (v >= 100000000000ull) + (v >= 1000000000000ull) +
           (v >= 10000000000000ull) + (v >= 100000000000000ull) + (v >= 1000000000000000ull) +
           (v >= 10000000000000000ull) + (v >= 100000000000000000ull) + (v >= 1000000000000000000ull) +
           (v >= 10000000000000000000ull);
}

// Prefix counters of one record, in u32 (a record whose sums reach 2^32 wraps the reference's own
// u32 accumulators, paf.rs:632-635 / bamstats.rs:26-32, and is rejected up front).
struct Ctr {
    uint32_t T;    // target bases  (M D N = X)
    uint32_t Q;    // query bases   (M I S = X)
    uint32_t A;    // alignment columns (all ops)            -> aln_len
    uint32_t EQ;   // '=' bases                               -> stats.equal
    uint32_t X;    // 'X' bases
    uint32_t M;    // 'M' bases                               -> stats.matches (+ diff)
    uint32_t I;    // inserted bases
    uint32_t D;    // deleted bases
    uint32_t IEV;  // insertion events (ops)
    uint32_t DEV;  // deletion events (ops)
    uint32_t TXT;  // bytes of the canonical "{len}{op}" text
    uint32_t aux;  // bit 31: sticky, the A sum wrapped past 2^32 ; bits 30..0: number of "slow" ops
                   // (zero-length, or same class as the op before it in the record) -> RF_SLOW
};
constexpr uint32_t AUX_OVF = 0x80000000u, AUX_CNT = 0x7FFFFFFFu;

RB_HD Ctr ctr_zero() {
    Ctr c;
    c.T = c.Q = c.A = c.EQ = c.X = c.M = c.I = c.D = c.IEV = c.DEV = c.TXT = c.aux = 0;
    return c;
}
RB_HD void ctr_add(Ctr& a, const Ctr& b) {
    a.T += b.T; a.Q += b.Q;
    const uint32_t s = a.A + b.A;
    a.aux = (((a.aux & AUX_CNT) + (b.aux & AUX_CNT)) & AUX_CNT) | ((a.aux | b.aux) & AUX_OVF) | (s < a.A ? AUX_OVF : 0u);
    a.A = s;
    a.EQ += b.EQ; a.X += b.X; a.M += b.M; a.I += b.I; a.D += b.D;
    a.IEV += b.IEV; a.DEV += b.DEV; a.TXT += b.TXT;
}
RB_HD void ctr_sub(Ctr& a, const Ctr& b) {  // aux is left alone
    a.T -= b.T; a.Q -= b.Q; a.A -= b.A; a.EQ -= b.EQ; a.X -= b.X; a.M -= b.M; a.I -= b.I; a.D -= b.D;
    a.IEV -= b.IEV; a.DEV -= b.DEV; a.TXT -= b.TXT;
}
// `n` bases of an op of class `code`; `whole` adds the event and the text of the complete op
RB_HD void ctr_add_bases(Ctr& c, uint32_t code, uint32_t n) {
    const uint32_t s = c.A + n;
    if (s < c.A) c.aux |= AUX_OVF;
    c.A = s;
    if (is_ref(code)) c.T += n;
    if (is_qry(code)) c.Q += n;
    if (code == OP_EQ) c.EQ += n;
    else if (code == OP_X) c.X += n;
    else if (code == OP_M) c.M += n;
    else if (code == OP_I) c.I += n;
    else if (code == OP_D) c.D += n;
}
RB_HD void ctr_add_op(Ctr& c, uint32_t w) {
    const uint32_t code = op_code(w), n = op_len(w);
    ctr_add_bases(c, code, n);
    c.IEV += (code == OP_I);
    c.DEV += (code == OP_D);
    c.TXT += ndigits32(n) + 1u;
}
RB_HD void ctr_sub_op(Ctr& c, uint32_t w) {
    Ctr t = ctr_zero();
    ctr_add_op(t, w);
    ctr_sub(c, t);
}

// record flags
enum : uint32_t {
    RF_MINUS = 1u,        // strand '-'
    RF_SLOW = 2u,         // has zero-length ops or adjacent same-class ops: trimmed rows need the merge walk (Q15)
    RF_STRIPPED = 4u,     // leading/trailing indels were stripped (id gets "_TO.<st>.<en>", paf.rs:726-732)
};

// Per-record state after remove_trailing_indels (paf.rs:656-783).  Prefix counters in `samples`
// are relative to op_first (the unstripped record); the lifted coordinates only ever need
// differences against the ORIGINAL q_st/q_en/t_st, the strip fix-ups cancel.
struct RecInfo {
    uint64_t op_first;   // global index of the record's first op (unstripped)
    uint64_t op_end;     // one past the record's last op (unstripped)
    uint64_t eo0, eo1;   // effective (stripped) op range [eo0, eo1)
    uint64_t t_st, t_en; // stripped target span (t_st never moves: a leading D aborts like the reference)
    uint64_t q_st, q_en; // stripped query span (what an early-return row prints)
    uint64_t q_st0, q_en0;  // original query span (lift arithmetic)
    uint64_t q_len, t_len, mapq;
    uint32_t q_name, t_name;  // indices into the name table
    uint32_t flags;
    uint32_t a_lead;     // alignment columns removed from the front (early-exit policy emulation)
    uint32_t n_lead, n_trail;  // number of stripped ops at either end
    uint32_t id_len;     // bytes of the "_TO.x.y" id suffix (0 if not stripped)
    Ctr tot;             // counters of the effective op range (early-return rows, integrity, rb stats)
    uint32_t wlo, whi;   // overlapping window range in the contig-sorted window arrays
    uint32_t pad;
};

// One (window, record) pair after lift: everything the serialiser and the numeric mirror need.
struct PairRes {
    uint64_t t_st, t_en, q_st, q_en;
    uint64_t nmatch, aln_len;
    uint64_t si, ei;      // global op indices of the first / last op of the trimmed CIGAR
    uint32_t s_len;       // length printed for op si (L_si - s.o, or e.o - s.o + 1 when si == ei)
    uint32_t e_len;       // length printed for op ei (e.o + 1); unused when si == ei
    uint32_t cg_bytes;    // bytes of the trimmed CIGAR text
    uint32_t kind;        // PK_*
    uint32_t equal, diff, ins, del, ins_ev, del_ev, matches;  // bamstats.rs:107-127 on the trimmed CIGAR
    uint32_t pad;
};
enum : uint32_t { PK_DROP = 0, PK_TRIM = 1, PK_EARLY = 2 };

}  // namespace rb
