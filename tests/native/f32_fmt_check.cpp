// Cross-check of the two f32 `Display` restatements (Rust prints f32 with the shortest digits that round-trip, positional):
//   oracle/rb_oracle.cpp  fmt_f32     — trial precisions with "%.*e" + strtof round trip
//   rustybam_b200/host    rbh::fmt_f32 — std::to_chars (shortest round trip, fixed notation)
// over every f32 in [lo, hi] (bit patterns), on all host threads.  The identities `rb stats` prints live in [0, 100].
//   usage: f32_fmt_check <lo> <hi> [stride]
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rb_oracle.hpp"

static std::string host_fmt(float v) {  // == rbh::fmt_f32 (rustybam_b200/host/rbhost.cpp)
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

int main(int argc, char** argv) {
    const float lo = argc > 1 ? (float)atof(argv[1]) : 0.0f, hi = argc > 2 ? (float)atof(argv[2]) : 100.0f;
    const uint32_t stride = argc > 3 ? (uint32_t)atoi(argv[3]) : 1u;
    uint32_t b0, b1;
    memcpy(&b0, &lo, 4); memcpy(&b1, &hi, 4);
    const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> n_diff{0}, n_all{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            uint64_t diff = 0, all = 0;
            for (uint64_t b = (uint64_t)b0 + (uint64_t)t * stride; b <= b1; b += (uint64_t)nt * stride) {
                float v;
                const uint32_t bb = (uint32_t)b;
                memcpy(&v, &bb, 4);
                const std::string a = orc::fmt_f32(v), c = host_fmt(v);
                all++;
                if (a != c) {
                    if (diff < 5) fprintf(stderr, "DIFF bits %08x: oracle %s host %s\n", bb, a.c_str(), c.c_str());
                    diff++;
                }
            }
            n_diff += diff; n_all += all;
        });
    for (auto& x : th) x.join();
    printf("checked=%llu DIFF=%llu\n", (unsigned long long)n_all.load(), (unsigned long long)n_diff.load());
    return n_diff ? 1 : 0;
}
