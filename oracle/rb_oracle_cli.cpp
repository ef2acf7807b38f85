// rb_oracle — command-line face of the CPU restatement (TEST INFRASTRUCTURE ONLY):
//   rb_oracle [-t N] [--policy rightmost|early-exit] liftover --bed B [--qbed] [--largest] in.paf
//   rb_oracle stats --paf [--qbed] in.paf
// Plain-text inputs only ('-' = stdin).  Exit status 101 where the reference would panic.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "rb_oracle.hpp"

static std::string slurp(const std::string& path) {
    std::ostringstream ss;
    if (path == "-") ss << std::cin.rdbuf();
    else {
        std::ifstream f(path, std::ios::binary);
        if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
        ss << f.rdbuf();
    }
    return ss.str();
}

int main(int argc, char** argv) {
    int threads = 8, policy = orc::POLICY_RIGHTMOST;
    bool qbed = false, largest = false, paf_flag = false;
    std::string cmd, bed, input = "-";
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-t" || a == "--threads") threads = atoi(argv[++i]);
        else if (a == "--policy") policy = strcmp(argv[++i], "early-exit") == 0 ? orc::POLICY_EARLY_EXIT : orc::POLICY_RIGHTMOST;
        else if (a == "--bed" || a == "-b") bed = argv[++i];
        else if (a == "--qbed" || a == "-q") qbed = true;
        else if (a == "--largest" || a == "-l") largest = true;
        else if (a == "--paf" || a == "-p") paf_flag = true;
        else if (cmd.empty()) cmd = a;
        else input = a;
    }
    try {
        std::string text = slurp(input), out;
        if (cmd == "liftover") {
            std::string b = slurp(bed);
            out = orc::run_liftover(text.data(), text.size(), b.data(), b.size(), qbed, largest, policy, threads);
        } else if (cmd == "stats" && paf_flag) {
            out = orc::run_stats(text.data(), text.size(), qbed);
        } else {
            fprintf(stderr, "usage: rb_oracle [-t N] liftover --bed B in.paf | stats --paf in.paf\n");
            return 2;
        }
        fwrite(out.data(), 1, out.size(), stdout);
    } catch (const orc::Abort& e) {
        fprintf(stderr, "thread 'main' panicked: %s\n", e.what());
        return 101;
    }
    return 0;
}
