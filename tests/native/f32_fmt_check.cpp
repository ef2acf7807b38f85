// Cross-check of the f32 `Display` restatements (Rust prints f32 with the shortest digits that round-trip, positional):
//   oracle/rb_oracle.cpp  fmt_f32     — trial precisions with "%.*e" + strtof round trip, exact-expansion tie test
//   std::to_chars (shortest round trip, fixed notation; Ryu: exact ties to even) — third opinion, differs at ties only
//   rustybam_b200/csrc/f32_fmt.cuh  rb::f32_display — Burger & Dybvig free-format digits over a 256-bit integer (device-ready)
// over every f32 in [lo, hi] (bit patterns), on all host threads.  The identities `rb stats` prints live in [0, 100].
//   usage: f32_fmt_check <lo> <hi> [stride] [no-oracle]
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "f32_fmt.cuh"
#include "../../rustybam_b200/host/f32_fast.hpp"
#include "rb_oracle.hpp"

static std::string host_fmt(float v) {  // std::to_chars (what the host used before f32_fmt.cuh)
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

int main(int argc, char** argv) {
    const float lo = argc > 1 ? (float)atof(argv[1]) : 0.0f, hi = argc > 2 ? (float)atof(argv[2]) : 100.0f;
    const uint32_t stride = argc > 3 ? (uint32_t)atoi(argv[3]) : 1u;
    const bool no_oracle = argc > 4;  // 4th argument: skip the (slow, printf-based) oracle, compare core vs std::to_chars only
    uint32_t b0, b1;
    memcpy(&b0, &lo, 4); memcpy(&b1, &hi, 4);
    const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> n_diff{0}, n_all{0}, n_ties{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            uint64_t diff = 0, all = 0, ties = 0;
            for (uint64_t b = (uint64_t)b0 + (uint64_t)t * stride; b <= b1; b += (uint64_t)nt * stride) {
                float v;
                const uint32_t bb = (uint32_t)b;
                memcpy(&v, &bb, 4);
                const std::string a = no_oracle ? std::string() : orc::fmt_f32(v), c = host_fmt(v);
                uint8_t buf[96];
                const std::string d((const char*)buf, (size_t)rb::f32_display(v, buf));
                all++;
                if (!no_oracle && a != d) {
                    if (diff < 5) fprintf(stderr, "DIFF bits %08x: oracle %s core %s (to_chars %s)\n", bb, a.c_str(), d.c_str(), c.c_str());
                    diff++;
                }
                {   // the device formatter's fast path (96-bit fixed point; k_emit's stats rows) must equal the core
                    const rb::F32Dec fd = rb::f32_shortest_fast(bb);
                    uint8_t fb[96];
                    const std::string g((const char*)fb, (size_t)(rb::f32_display_put(fb, bb, fd) - fb));
                    if (g != d || g.size() != rb::f32_display_len(bb, fd)) {
                        if (diff < 5) fprintf(stderr, "DIFF bits %08x: device fast path %s (len %u) core %s\n", bb, g.c_str(), rb::f32_display_len(bb, fd), d.c_str());
                        diff++;
                    }
                }
                if (rbh::f32_display_fast(v) != d) {  // the host's fast path (to_chars + exact tie test) must equal the core
                    if (diff < 5) fprintf(stderr, "DIFF bits %08x: host fast path %s core %s\n", bb, rbh::f32_display_fast(v).c_str(), d.c_str());
                    diff++;
                }
                if (d != c) {  // must be an exact tie: same length, the last digit one higher than std::to_chars' (half-even) choice
                    bool tie_shape = d.size() == c.size() && d.compare(0, d.size() - 1, c, 0, c.size() - 1) == 0 && d.back() == c.back() + 1;
                    // (a carry into earlier digits keeps the length but changes more than the last digit)
                    if (!tie_shape && !(d.size() == c.size() && d > c)) {
                        if (diff < 5) fprintf(stderr, "DIFF (not a tie) bits %08x: core %s to_chars %s\n", bb, d.c_str(), c.c_str());
                        diff++;
                    }
                    ties++;
                }
            }
            n_diff += diff; n_all += all; n_ties += ties;
        });
    for (auto& x : th) x.join();
    printf("checked=%llu DIFF=%llu ties_rounded_up_where_to_chars_rounds_to_even=%llu\n", (unsigned long long)n_all.load(),
           (unsigned long long)n_diff.load(), (unsigned long long)n_ties.load());
    return n_diff ? 1 : 0;
}
