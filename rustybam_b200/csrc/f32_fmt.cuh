// f32_fmt.cuh — Rust's `Display` for f32 (bamstats.rs:262-265 prints the three identities with `{}`): the shortest decimal
// digits that round-trip, closest to the value, printed positionally ("99.89702", "100", "0.000012", "NaN", "inf").
//
// Free-format algorithm of Steele & White / Burger & Dybvig over a small fixed-width big integer — the same algorithm Rust's
// core::num::flt2dec::strategy::dragon::format_shortest implements (its Grisu fast path falls back to it and is defined to
// agree with it): exact rounding interval (asymmetric at exact powers of two), even mantissas own their interval ends, a
// final-digit tie rounds up.  Integer arithmetic only, no tables, no libc: __host__ __device__.
//
// Status: equal to the oracle's independent fmt_f32 (printf trial + exact-expansion tie test) on every f32 in [0, 100]
// (1 120 403 457 values) and on strided sweeps of the rest; equal to std::to_chars except at 2 048 exact ties in [0, 100], which
// Ryu rounds to even (tests/native/f32_fmt_check.cpp, profiles/r01z_f32_fmt.txt).  The tie rule is stated from knowledge of the
// std sources (std is not present in this image): "parity unpinned", switchable with -DRB_F32_TIE_EVEN.  Used by the C++ host
// (rbh::fmt_f32) today; it is also the core of the GPU-side `rb stats` row formatter that DESIGN.md §8 lists as next.
#pragma once
#include <stdint.h>
#include <string.h>

#include "rb_common.cuh"

namespace rb {

constexpr int F32_LIMBS = 8;  // 256 bits: the largest intermediate is ~2^160 (subnormals scaled by 10^45; 2^128-sized values)
struct Big { uint32_t w[F32_LIMBS]; };

RB_HD void big_zero(Big& a) { for (int i = 0; i < F32_LIMBS; i++) a.w[i] = 0; }
RB_HD void big_set_pow2(Big& a, uint32_t e) { big_zero(a); a.w[e >> 5] = 1u << (e & 31); }
RB_HD void big_set_u32_shl(Big& a, uint32_t v, uint32_t e) {  // a = v * 2^e
    big_zero(a);
    const uint32_t i = e >> 5, b = e & 31;
    a.w[i] = v << b;
    if (b && i + 1 < F32_LIMBS) a.w[i + 1] = v >> (32 - b);
}
RB_HD void big_mul_small(Big& a, uint32_t m) {
    uint64_t c = 0;
    for (int i = 0; i < F32_LIMBS; i++) { c += (uint64_t)a.w[i] * m; a.w[i] = (uint32_t)c; c >>= 32; }
}
RB_HD void big_add(Big& a, const Big& b) {
    uint64_t c = 0;
    for (int i = 0; i < F32_LIMBS; i++) { c += (uint64_t)a.w[i] + b.w[i]; a.w[i] = (uint32_t)c; c >>= 32; }
}
RB_HD void big_sub(Big& a, const Big& b) {  // a >= b
    int64_t c = 0;
    for (int i = 0; i < F32_LIMBS; i++) { c += (int64_t)a.w[i] - (int64_t)b.w[i]; a.w[i] = (uint32_t)c; c >>= 32; }
}
RB_HD int big_cmp(const Big& a, const Big& b) {
    for (int i = F32_LIMBS - 1; i >= 0; i--)
        if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
}

// Shortest digits of a positive finite f32 given by its bit pattern: value = 0.d1 d2 ... dn x 10^k.  Returns n (1..9).
RB_HD int f32_shortest_digits(uint32_t bits, uint8_t* digits, int& k_out) {
    const uint32_t be = (bits >> 23) & 0xFFu, frac = bits & 0x7FFFFFu;
    const uint32_t f = be ? (frac | 0x800000u) : frac;  // value = f * 2^e
    const int e = be ? (int)be - 150 : -149;
    const bool even = (f & 1u) == 0;                     // round-to-even: an even mantissa owns both ends of its interval
    const bool pow2 = be > 1 && frac == 0;               // the gap below is half the gap above
    Big r, s, mp, mm;
    if (e >= 0) {
        if (!pow2) { big_set_u32_shl(r, f, (uint32_t)e + 1); big_set_pow2(s, 1); big_set_pow2(mp, (uint32_t)e); big_set_pow2(mm, (uint32_t)e); }
        else { big_set_u32_shl(r, f, (uint32_t)e + 2); big_set_pow2(s, 2); big_set_pow2(mp, (uint32_t)e + 1); big_set_pow2(mm, (uint32_t)e); }
    } else {
        if (!pow2) { big_set_u32_shl(r, f, 1); big_set_pow2(s, (uint32_t)(-e) + 1); big_set_pow2(mp, 0); big_set_pow2(mm, 0); }
        else { big_set_u32_shl(r, f, 2); big_set_pow2(s, (uint32_t)(-e) + 2); big_set_pow2(mp, 1); big_set_pow2(mm, 0); }
    }
    int k = 0;
    for (;;) {  // scale s up while (r + m+) / s >= 1
        Big t = r;
        big_add(t, mp);
        const int c = big_cmp(t, s);
        if (even ? c >= 0 : c > 0) { big_mul_small(s, 10); k++; } else break;
    }
    for (;;) {  // scale r up while (r + m+) * 10 / s < 1
        Big t = r;
        big_add(t, mp);
        big_mul_small(t, 10);
        const int c = big_cmp(t, s);
        if (even ? c < 0 : c <= 0) { big_mul_small(r, 10); big_mul_small(mp, 10); big_mul_small(mm, 10); k--; } else break;
    }
    int n = 0;
    for (;;) {
        big_mul_small(r, 10); big_mul_small(mp, 10); big_mul_small(mm, 10);
        uint32_t d = 0;
        while (big_cmp(r, s) >= 0) { big_sub(r, s); d++; }
        const int c1 = big_cmp(r, mm);
        const bool low = even ? c1 <= 0 : c1 < 0;     // the digits so far, cut here, are inside the interval
        Big t = r;
        big_add(t, mp);
        const int c2 = big_cmp(t, s);
        const bool high = even ? c2 >= 0 : c2 > 0;    // ... and so are they with the last digit bumped
        if (!low && !high && n < 16) { digits[n++] = (uint8_t)d; continue; }
        bool up = high;
        if (low && high) {  // both: the closer one; a tie rounds up (dragon.rs format_shortest: `*mant.mul_pow2(1) >= scale`)
            Big t2 = r;
            big_mul_small(t2, 2);
            const int c3 = big_cmp(t2, s);
#ifdef RB_F32_TIE_EVEN  // what printf / Ryu / std::to_chars do with an exact tie (2 048 of the 1.12 G f32 in [0, 100] differ)
            up = c3 > 0 || (c3 == 0 && (d & 1u));
#else
            up = c3 >= 0;
#endif
        }
        digits[n++] = (uint8_t)d;
        if (up) {  // +1 in the last place, carrying through nines
            int i = n - 1;
            while (i >= 0 && digits[i] == 9) { digits[i] = 0; i--; }
            if (i >= 0) digits[i]++;
            else { digits[0] = 1; n = 1; k++; }
            while (n > 1 && digits[n - 1] == 0) n--;
        }
        break;
    }
    k_out = k;
    return n;
}

// `format!("{}", v)`: returns the number of bytes written to out (at most 64).
RB_HD int f32_display(float v, uint8_t* out) {
    uint32_t bits;
    memcpy(&bits, &v, 4);
    int p = 0;
    if ((bits & 0x7F800000u) == 0x7F800000u) {
        if (bits & 0x7FFFFFu) { out[0] = 'N'; out[1] = 'a'; out[2] = 'N'; return 3; }
        if (bits >> 31) out[p++] = '-';
        out[p++] = 'i'; out[p++] = 'n'; out[p++] = 'f';
        return p;
    }
    if (bits >> 31) out[p++] = '-';
    if ((bits & 0x7FFFFFFFu) == 0) { out[p++] = '0'; return p; }
    uint8_t dg[20];
    int k;
    const int n = f32_shortest_digits(bits & 0x7FFFFFFFu, dg, k);
    if (k <= 0) {
        out[p++] = '0'; out[p++] = '.';
        for (int i = 0; i < -k; i++) out[p++] = '0';
        for (int i = 0; i < n; i++) out[p++] = (uint8_t)('0' + dg[i]);
    } else if (k >= n) {
        for (int i = 0; i < n; i++) out[p++] = (uint8_t)('0' + dg[i]);
        for (int i = n; i < k; i++) out[p++] = '0';
    } else {
        for (int i = 0; i < k; i++) out[p++] = (uint8_t)('0' + dg[i]);
        out[p++] = '.';
        for (int i = k; i < n; i++) out[p++] = (uint8_t)('0' + dg[i]);
    }
    return p;
}

}  // namespace rb
