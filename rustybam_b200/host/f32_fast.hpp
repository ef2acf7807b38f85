// f32_fast.hpp — host-side fast path of Rust's f32 `Display` (bamstats.rs:262-265): std::to_chars gives the shortest
// round-trip digits in ~75 ns and agrees with csrc/f32_fmt.cuh (the Burger-Dybvig core, ~1.3 us on a host core) everywhere
// except at exact ties, which Ryu rounds to even and Rust's Dragon rounds up.  A tie is detected exactly, in integers, and only
// then the core is asked.  Checked against the core on every f32 in [0, 100] (tests/native/f32_fmt_check.cpp).
#pragma once
#include <charconv>
#include <cstdint>
#include <cstring>
#include <string>

#include "../csrc/f32_fmt.cuh"

namespace rbh {

inline std::string f32_display_core(float v) {
    uint8_t buf[96];
    return std::string(reinterpret_cast<const char*>(buf), (size_t)rb::f32_display(v, buf));
}

inline std::string f32_display_fast(float v) {
    if (!(v > 0.0f) || !(v < 16777216.0f)) return f32_display_core(v);  // zero, negative, NaN, inf, integers printed padded
    char buf[128];
    const auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    // v = m * 2^e ; printed c = D * 10^-q.  Tie (half-even kept the lower candidate) <=> v * 10^q == D + 1/2 exactly
    //   <=> m_odd * 5^q * 2^(tz + e + 1 + q) == 2 D + 1  <=> tz + e + 1 + q == 0 and m_odd * 5^q == 2 D + 1
    uint64_t D = 0;
    int q = 0;
    bool point = false;
    for (const char* c = buf; c < r.ptr; c++) {
        if (*c == '.') { point = true; continue; }
        D = D * 10 + (uint64_t)(*c - '0');
        if (point) q++;
        if (D > 4000000000ull) return std::string(buf, r.ptr);  // more than 9 significant digits never happens for f32
    }
    if (q >= 14) return std::string(buf, r.ptr);  // 5^14 > 2 D + 1 for every D < 10^9: no tie
    uint32_t bits;
    memcpy(&bits, &v, 4);
    const uint32_t be = (bits >> 23) & 0xFFu, frac = bits & 0x7FFFFFu;
    const uint32_t m = be ? (frac | 0x800000u) : frac;
    const int e = be ? (int)be - 150 : -149;
    const int tz = __builtin_ctz(m);
    if (tz + e + 1 + q != 0) return std::string(buf, r.ptr);
    uint64_t p5 = 1;
    for (int i = 0; i < q; i++) p5 *= 5;
    if ((uint64_t)(m >> tz) * p5 == 2 * D + 1) return f32_display_core(v);
    return std::string(buf, r.ptr);
}

}  // namespace rbh
