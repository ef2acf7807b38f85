// rb_common.cuh — shared types of the B200 liftover/stats path (device + host-side unit tests).
//
// Data layout in HBM:
//   text      u8[N]            concatenated cg:Z: payloads of all records (read once by the tokeniser)
//   ops       u32[n_ops]       one word per CIGAR op: (len << 4) | code, BAM op codes, len < 2^28
//   heads     u32[n_ops/32+1]  bit k%32 of word k/32 set  <=>  op k is the first op of a record
//   samples   Ctr[(n_ops/32+1)*4] sampled segmented prefix sums: samples[4c] = counters accumulated from
//                              the first op of the record containing op 32c up to (excluding) op 32c;
//                              samples[4c+s] (s = 1..3) = the same at op 32c + 8s, relative to the chunk start
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
#ifndef RB_SAMPLE_LOG2
#define RB_SAMPLE_LOG2 5
#endif
constexpr int SAMPLE_LOG2 = RB_SAMPLE_LOG2;  // 5 or 4
constexpr uint32_t SAMPLE = 1u << SAMPLE_LOG2;  // ops per sample chunk (== bits of a heads word)
// Sub-samples: every chunk keeps SUBS counters, one per SUB_OPS ops.  Entry 0 is the absolute sample (counters of
// the record before the chunk's first op); entries 1.. are RELATIVE to the chunk start — or, when a record starts
// inside the chunk before that position (SUB_ABS set in aux), relative to that record's first op, i.e. absolute.
// They cut the per-boundary walk of find_op / ctr_before from <= 31 ops to <= 7.
constexpr int SUB_LOG2 = 3;
constexpr uint32_t SUB_OPS = 1u << SUB_LOG2;
constexpr uint32_t SUBS = SAMPLE / SUB_OPS;
constexpr uint32_t SUB_ABS = 0x40000000u;

RB_HD uint32_t op_len(uint32_t w) { return w >> 4; }
RB_HD uint32_t op_code(uint32_t w) { return w & 15u; }
RB_HD bool is_ref(uint32_t code) { return (MASK_REF >> code) & 1u; }
RB_HD bool is_qry(uint32_t code) { return (MASK_QRY >> code) & 1u; }
RB_HD bool is_match(uint32_t code) { return (MASK_MATCH >> code) & 1u; }

RB_HD uint32_t clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)v);
#else
    return v ? (uint32_t)__builtin_clz(v) : 32u;
#endif
}
// decimal digits of v (v < 2^32): bit length -> estimate of log10 -> one compare against a power of ten
#if defined(__CUDACC__)
static __constant__ uint32_t c_digits_pow10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};
#endif
RB_HD uint32_t ndigits32(uint32_t v) {
    const uint32_t x = v | 1u;                         // 0 prints as "0": one digit (x has the digit count of v for every v)
    const uint32_t t = ((32u - clz32(x)) * 1233u) >> 12;  // floor(bits * log10(2)): digits - 1 or digits
#if defined(__CUDA_ARCH__)
    return t + (x >= c_digits_pow10[t] ? 1u : 0u);
#else
    static const uint32_t p10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};
    return t + (x >= p10[t] ? 1u : 0u);
#endif
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
// class flags per BAM code, 7 bits each (9 codes = 63 bits of one constant):
//   bit0 target-consuming  bit1 query-consuming  bit2 '='  bit3 'X'  bit4 'M'  bit5 'I'  bit6 'D'
constexpr uint64_t class_flags_table() {
    const uint64_t f[9] = {0x13 /*M*/, 0x22 /*I*/, 0x41 /*D*/, 0x01 /*N*/, 0x02 /*S*/, 0x00 /*H*/, 0x00 /*P*/, 0x07 /*=*/, 0x0B /*X*/};
    uint64_t t = 0;
    for (int c = 0; c < 9; c++) t |= f[c] << (7 * c);
    return t;
}
constexpr uint64_t CLASS_FLAGS = class_flags_table();
RB_HD uint32_t class_flags(uint32_t code) { return (uint32_t)(CLASS_FLAGS >> (7u * code)) & 0x7Fu; }

// `n` bases of an op of class `code` (no event, no text).  Written as bit tests so that the
// compiler emits one predicate + one predicated add per counter.
RB_HD void ctr_add_bases(Ctr& c, uint32_t code, uint32_t n) {
    const uint32_t f = class_flags(code);
    const uint32_t s = c.A + n;
    if (s < n) c.aux |= AUX_OVF;
    c.A = s;
    if (f & 1u) c.T += n;
    if (f & 2u) c.Q += n;
    if (f & 4u) c.EQ += n;
    if (f & 8u) c.X += n;
    if (f & 16u) c.M += n;
    if (f & 32u) c.I += n;
    if (f & 64u) c.D += n;
}
RB_HD void ctr_add_op(Ctr& c, uint32_t w) {
    const uint32_t code = op_code(w), n = op_len(w);
    const uint32_t f = class_flags(code);
    const uint32_t s = c.A + n;
    if (s < n) c.aux |= AUX_OVF;
    c.A = s;
    if (f & 1u) c.T += n;
    if (f & 2u) c.Q += n;
    if (f & 4u) c.EQ += n;
    if (f & 8u) c.X += n;
    if (f & 16u) c.M += n;
    if (f & 32u) { c.I += n; c.IEV++; }
    if (f & 64u) { c.D += n; c.DEV++; }
    c.TXT += ndigits32(n) + 1u;
}
RB_HD void ctr_sub_op(Ctr& c, uint32_t w) {
    Ctr t = ctr_zero();
    ctr_add_op(t, w);
    ctr_sub(c, t);
}

// ---- cheap accumulation for the hot walks -------------------------------------------------------
// Exactly one class sum changes per op, so the walks keep nine per-thread class sums in an indexed
// array (shared memory on the device: sum[code * stride], stride = threads per block, conflict-free;
// a local array on the host) and derive the 11 counters only when a chunk / boundary is flushed.
struct ClassAcc {
    uint32_t* sum;     // sum[code * stride], codes 0..8
    uint32_t stride;
    uint32_t T;        // running target bases (the find test needs it every op)
    uint32_t iev, dev, txt;
    uint32_t big;      // OR of all lengths seen: class sums of <= 32 ops cannot wrap while every len < 2^26
};
RB_HD void acc_reset(ClassAcc& a) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t c = 0; c < 9u; c++) a.sum[c * a.stride] = 0u;
    a.T = a.iev = a.dev = a.txt = a.big = 0u;
}
RB_HD void acc_add_op(ClassAcc& a, uint32_t w) {
    const uint32_t code = op_code(w), n = op_len(w);
    a.sum[code * a.stride] += n;
    if (is_ref(code)) a.T += n;
    a.iev += (code == OP_I);
    a.dev += (code == OP_D);
    a.txt += ndigits32(n) + 1u;
    a.big |= n;
}
// c += everything accumulated since acc_reset
RB_HD void acc_flush(const ClassAcc& a, Ctr& c) {
    const uint32_t M = a.sum[OP_M * a.stride], I = a.sum[OP_I * a.stride], D = a.sum[OP_D * a.stride],
                   N = a.sum[OP_N * a.stride], S = a.sum[OP_S * a.stride], H = a.sum[OP_H * a.stride],
                   P = a.sum[OP_P * a.stride], E = a.sum[OP_EQ * a.stride], X = a.sum[OP_X * a.stride];
    const uint64_t all = (uint64_t)M + I + D + N + S + H + P + E + X + c.A;
    if (all >> 32) c.aux |= AUX_OVF;
    c.A = (uint32_t)all;
    c.T += M + D + N + E + X;
    c.Q += M + I + S + E + X;
    c.EQ += E; c.X += X; c.M += M; c.I += I; c.D += D;
    c.IEV += a.iev; c.DEV += a.dev; c.TXT += a.txt;
}
constexpr uint32_t ACC_BIG = 1u << 26;

// record flags
enum : uint32_t {
    RF_MINUS = 1u,        // strand '-'
    RF_SLOW = 2u,         // has zero-length ops or adjacent same-class ops: trimmed rows need the merge walk (Q15); also set for
                          // records holding an op of >= 2^26 bases (the walks' class sums could wrap): such records stay on
                          // the general kernels, the fused fast path never sees them
    RF_STRIPPED = 4u,     // leading/trailing indels were stripped (id gets "_TO.<st>.<en>", paf.rs:726-732)
    RF_CANON = 8u,        // the CIGAR text is canonical (no leading zeros): op k's text starts TXT_prefix(k) bytes into it,
                          // so untouched ops of a trimmed row are copied from the input text instead of being re-formatted
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
    uint32_t lead_txt;   // canonical text bytes of the stripped leading ops (text offset of op eo0 inside the CIGAR)
    uint64_t text_off;   // byte offset of the record's CIGAR in the device text buffer
    uint32_t line_const; // bytes of a printed row that depend on the record only (names, q_len, t_len, mapq, separators)
    uint32_t pad2;
};

// One (window, record) pair after lift: everything the serialiser and the numeric mirror need.
struct PairRes {          // 112 bytes
    uint64_t t_st, t_en, q_st, q_en;
    uint64_t si, ei;      // global op indices of the first / last op of the trimmed CIGAR
    uint64_t mid_off;     // RF_CANON rows: where the untouched ops (strictly between si and ei; early rows: all ops) start in
                          // the device text buffer ...
    uint32_t mid_len;     // ... and how many bytes they are
    uint32_t nmatch, aln_len;  // u32 like the reference's own accumulators (paf.rs:632-635)
    uint32_t s_len;       // length printed for op si (L_si - s.o, or e.o - s.o + 1 when si == ei)
    uint32_t e_len;       // length printed for op ei (e.o + 1); unused when si == ei
    uint32_t cg_bytes;    // bytes of the trimmed CIGAR text
    uint32_t kind;        // PK_*
    uint32_t equal, diff, ins, del, ins_ev, del_ev, matches;  // bamstats.rs:107-127 on the trimmed CIGAR
};
static_assert(sizeof(PairRes) == 112, "PairRes layout");
enum : uint32_t { PK_DROP = 0, PK_TRIM = 1, PK_EARLY = 2,
                  PK_WHOLE = 3 };  // rb invert: the record as it is; k_serialise leaves the CIGAR bytes of the row to k_whole_text

}  // namespace rb
