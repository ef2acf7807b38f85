// synth.cpp — synthetic whole-genome-scale eqx PAFs and tiling BED windows (configs C2-C5
// of BASELINE.json).  There is no network for real HPRC/CHM13 alignments, so the benchmark inputs are
// generated: CHM13-like contig lengths, records that tile each contig with log-uniform spans
// (20 kb - 30 Mbp), '-' strand with p = 0.1, '=' runs ~ Geometric(mean 120) separated by
// X (p 0.88, len 1) / I / D (p 0.06 each, len 1 + Geometric(mean 3), 1 % heavy tail <= 50 kb):
// ~16 ops/kb, ~50 M ops and ~0.14 GB of CIGAR text per haplotype at scale 1.  Records start and end on
// '=' and their coordinates are consistent with the CIGAR, so the reference's check_integrity passes.
#include <cmath>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <thread>

#include "rbhost.hpp"

namespace rbh {

static const struct { const char* name; uint64_t len; } CHM13[] = {
    {"chr1", 248387328}, {"chr2", 242696752}, {"chr3", 201105948}, {"chr4", 193574945}, {"chr5", 182045439},
    {"chr6", 172126628}, {"chr7", 160567428}, {"chr8", 146259331}, {"chr9", 150617247}, {"chr10", 134758134},
    {"chr11", 135127769}, {"chr12", 133324548}, {"chr13", 113566686}, {"chr14", 101161492}, {"chr15", 99753195},
    {"chr16", 96330374}, {"chr17", 84276897}, {"chr18", 80542538}, {"chr19", 61707364}, {"chr20", 66210255},
    {"chr21", 45090682}, {"chr22", 51324926}, {"chrX", 154259566}, {"chrY", 62460029}, {"chrM", 16569}};
static const int N_CONTIG = sizeof(CHM13) / sizeof(CHM13[0]);

struct Rng {  // xoshiro256** seeded through splitmix64
    uint64_t s[4];
    explicit Rng(uint64_t seed) {
        for (int i = 0; i < 4; i++) {
            seed += 0x9E3779B97F4A7C15ull;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }  // (0,1)
    uint64_t geo(double mean_minus_1) { return 1 + (uint64_t)(-std::log(uni()) * mean_minus_1); }
};

struct ContigOut {
    std::vector<uint8_t> cigar;
    std::vector<uint64_t> cigar_end, q_st, q_en, t_st, t_en;
    std::vector<uint8_t> strand;
    uint64_t q_len = 0, t_len = 0;
};

static inline void put_op(std::vector<uint8_t>& out, uint64_t len, char op) {
    char buf[24];
    int n = 0;
    do { buf[n++] = (char)('0' + len % 10); len /= 10; } while (len);
    while (n) out.push_back((uint8_t)buf[--n]);
    out.push_back((uint8_t)op);
}

static void gen_contig(uint64_t seed, int hap, int contig, double scale, ContigOut& o) {
    Rng rng(seed * 0x100000001B3ull + (uint64_t)hap * 1000003ull + (uint64_t)contig * 7919ull + 1);
    const uint64_t tlen = std::max<uint64_t>(2000, (uint64_t)((double)CHM13[contig].len * scale));
    o.t_len = tlen;
    const double lo = std::log(std::min<double>(20e3, (double)tlen / 4 + 1)), hi = std::log(std::min<double>(30e6, (double)tlen));
    uint64_t pos = 0, qpos = 1000 + rng.next() % 5000;
    o.cigar.reserve((size_t)((double)tlen * 0.05));
    while (pos + 1000 < tlen) {
        uint64_t span = (uint64_t)std::exp(lo + (hi - lo) * rng.uni());
        span = std::min<uint64_t>(std::max<uint64_t>(span, 50), tlen - pos);
        uint64_t T = 0, Q = 0;
        for (;;) {
            uint64_t e = std::min<uint64_t>(rng.geo(119.0), 150000);
            if (T + e + 2 >= span) {  // close the record on '='
                e = span - T;
                put_op(o.cigar, e, '=');
                T += e; Q += e;
                break;
            }
            put_op(o.cigar, e, '=');
            T += e; Q += e;
            const double r = rng.uni();
            if (r < 0.88) {
                put_op(o.cigar, 1, 'X');
                T += 1; Q += 1;
            } else {
                uint64_t len = rng.geo(3.0);
                if (rng.uni() < 0.01) len = (uint64_t)std::exp(std::log(50.0) + (std::log(50000.0) - std::log(50.0)) * rng.uni());
                if (r < 0.94) {
                    put_op(o.cigar, len, 'I');
                    Q += len;
                } else {
                    len = std::min<uint64_t>(len, span - T - 1);
                    put_op(o.cigar, len, 'D');
                    T += len;
                }
            }
        }
        o.cigar_end.push_back(o.cigar.size());
        o.t_st.push_back(pos); o.t_en.push_back(pos + T);
        o.q_st.push_back(qpos); o.q_en.push_back(qpos + Q);
        o.strand.push_back(rng.uni() < 0.1 ? '-' : '+');
        pos += T + rng.next() % 200;   // small unaligned gaps between records
        qpos += Q + rng.next() % 200;
    }
    o.q_len = qpos + 1000;
}

Paf synth_paf(const SynthParams& p) {
    const int n_task = p.n_hap * N_CONTIG;
    std::vector<ContigOut> outs((size_t)n_task);
    std::vector<std::thread> pool;
    std::atomic<int> next{0};
    const int nt = std::max(1, std::min(p.threads, n_task));
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&] {
            for (;;) {
                const int k = next.fetch_add(1);
                if (k >= n_task) break;
                if (!((p.contig_mask >> (k % N_CONTIG)) & 1u)) continue;  // (one independent stream per haplotype and contig)
                gen_contig(p.seed, k / N_CONTIG, k % N_CONTIG, p.scale, outs[(size_t)k]);
            }
        });
    for (auto& th : pool) th.join();

    Paf paf;
    size_t total = 0, nrec = 0;
    for (auto& o : outs) { total += o.cigar.size(); nrec += o.strand.size(); }
    paf.cigar.reserve(total);
    paf.cigar_off.reserve(nrec + 1);
    for (int k = 0; k < n_task; k++) {
        ContigOut& o = outs[(size_t)k];
        const int hap = k / N_CONTIG, c = k % N_CONTIG;
        if (!((p.contig_mask >> c) & 1u)) continue;
        const uint32_t tid = paf.name_id(CHM13[c].name);
        const uint32_t qid = paf.name_id("hap" + std::to_string(hap + 1) + "#" + CHM13[c].name);
        const uint64_t base = paf.cigar.size();
        paf.cigar.insert(paf.cigar.end(), o.cigar.begin(), o.cigar.end());
        for (size_t i = 0; i < o.strand.size(); i++) {
            paf.cigar_off.push_back(base + o.cigar_end[i]);
            paf.q_len.push_back(o.q_len); paf.q_st.push_back(o.q_st[i]); paf.q_en.push_back(o.q_en[i]);
            paf.t_len.push_back(o.t_len); paf.t_st.push_back(o.t_st[i]); paf.t_en.push_back(o.t_en[i]);
            paf.mapq.push_back(60);
            paf.strand.push_back(o.strand[i]);
            paf.q_id.push_back(qid); paf.t_id.push_back(tid);
        }
        o = ContigOut();
    }
    return paf;
}

static void target_lengths(const Paf& paf, std::vector<std::pair<uint32_t, uint64_t>>& out) {
    std::vector<int64_t> seen(paf.names.size(), -1);
    for (size_t i = 0; i < paf.size(); i++) {
        const uint32_t t = paf.t_id[i];
        if (seen[t] < 0) { seen[t] = (int64_t)out.size(); out.emplace_back(t, paf.t_len[i]); }
        else out[(size_t)seen[t]].second = std::max(out[(size_t)seen[t]].second, paf.t_len[i]);
    }
}

std::vector<Region> tiling_windows(const Paf& paf, uint64_t width) {
    std::vector<std::pair<uint32_t, uint64_t>> tl;
    target_lengths(paf, tl);
    std::vector<Region> out;
    for (auto& c : tl)
        for (uint64_t st = 0; st < c.second; st += width) {
            Region r;
            r.name = paf.names[c.first]; r.st = st; r.en = std::min(st + width, c.second);
            r.id = r.name + ":" + std::to_string(r.st + 1) + "-" + std::to_string(r.en);
            out.push_back(std::move(r));
        }
    return out;
}

// same rows as Windows::pack(tiling_windows(paf, width), paf) without the per-row std::string traffic
Windows tiling_windows_packed(const Paf& paf, uint64_t width) {
    std::vector<std::pair<uint32_t, uint64_t>> tl;
    target_lengths(paf, tl);
    std::vector<uint32_t> row0(tl.size() + 1, 0);  // BED file order = first-appearance order of the contigs
    for (size_t k = 0; k < tl.size(); k++) row0[k + 1] = row0[k] + (uint32_t)((tl[k].second + width - 1) / width);
    std::vector<size_t> by_tid(tl.size());
    for (size_t k = 0; k < tl.size(); k++) by_tid[k] = k;
    std::sort(by_tid.begin(), by_tid.end(), [&](size_t a, size_t b) { return tl[a].first < tl[b].first; });
    Windows w;
    const size_t n = row0.back();
    w.t_id.reserve(n); w.st.reserve(n); w.en.reserve(n); w.bed_row.reserve(n);
    w.default_ids = true;  // 3-column BED: the GPU formats "{chrom}:{st+1}-{en}" itself
    for (size_t k : by_tid) {
        uint32_t row = row0[k];
        for (uint64_t st = 0; st < tl[k].second; st += width, row++) {
            const uint64_t en = std::min(st + width, tl[k].second);
            w.t_id.push_back(tl[k].first); w.st.push_back(st); w.en.push_back(en); w.bed_row.push_back(row);
        }
    }
    if (w.ids.empty()) w.ids.push_back(0);
    return w;
}

}  // namespace rbh
