// C API over rb_oracle for the Python test-suite (ctypes).  TEST INFRASTRUCTURE ONLY.
// Return codes: 0 = ok / Some, 1 = None (row dropped), 101 = the reference would panic.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rb_oracle.hpp"

using namespace orc;

static char* dup_out(const std::string& s, size_t* n) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    if (n) *n = s.size();
    return p;
}
static void set_err(char* err, size_t cap, const char* msg) {
    if (err && cap) {
        strncpy(err, msg, cap - 1);
        err[cap - 1] = 0;
    }
}

extern "C" {

void orc_free(char* p) { free(p); }

int orc_run_liftover(const char* paf, size_t paf_n, const char* bed, size_t bed_n, int qbed,
                     int largest, int policy, int threads, char** out, size_t* out_n, char* err,
                     size_t err_cap) {
    try {
        *out = dup_out(run_liftover(paf, paf_n, bed, bed_n, qbed != 0, largest != 0, policy, threads), out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

int orc_run_break_paf(const char* paf, size_t paf_n, uint32_t max_size, int policy, char** out, size_t* out_n, char* err,
                      size_t err_cap) {
    try {
        *out = dup_out(run_break_paf(paf, paf_n, max_size, policy), out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

int orc_run_invert(const char* paf, size_t paf_n, char** out, size_t* out_n, char* err, size_t err_cap) {
    try {
        *out = dup_out(run_invert(paf, paf_n), out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

// trim_overlap.rs:22-34 (doctest shape): two record lines -> aligned_pairs on both -> trim_overlapping_pafs -> two lines
int orc_trim_pair(const char* left_line, const char* right_line, int match_score, int diff_score, int indel_score, int policy,
                  char** out, size_t* out_n, char* err, size_t err_cap) {
    try {
        PafRecord l = PafRecord::parse(left_line), r = PafRecord::parse(right_line);
        l.aligned_pairs();
        r.aligned_pairs();
        trim_overlapping_pafs(l, r, match_score, diff_score, indel_score, policy);
        *out = dup_out(l.to_line() + "\n" + r.to_line() + "\n", out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    } catch (const ParseSkip&) {
        set_err(err, err_cap, "line skipped");
        return 2;
    }
}

// stepping form of `rb trim-paf` for the multi-rank tests: begin -> round* -> end
struct TrimHandle { Paf paf; std::vector<bool> contained; };
void* orc_trim_begin(const char* paf, size_t paf_n, char* err, size_t err_cap) {
    try {
        TrimHandle* h = new TrimHandle();
        h->paf = Paf::from_text(paf, paf_n);
        return h;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return nullptr;
    }
}
int orc_trim_round(void* hv, int match_score, int diff_score, int indel_score, int policy, int* waiting, char* err, size_t err_cap) {
    TrimHandle* h = static_cast<TrimHandle*>(hv);
    try {
        *waiting = h->paf.trim_round(match_score, diff_score, indel_score, policy, h->contained) > 0 ? 1 : 0;
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}
int orc_trim_end(void* hv, int remove_contained, char** out, size_t* out_n) {
    TrimHandle* h = static_cast<TrimHandle*>(hv);
    if (remove_contained) h->paf.drop_contained(h->contained);
    std::string text;
    for (const PafRecord& r : h->paf.records) { text += r.to_line(); text += '\n'; }
    *out = dup_out(text, out_n);
    delete h;
    return 0;
}

int orc_run_trim_paf(const char* paf, size_t paf_n, int match_score, int diff_score, int indel_score, int remove_contained,
                     int policy, char** out, size_t* out_n, char* err, size_t err_cap) {
    try {
        *out = dup_out(run_trim_paf(paf, paf_n, match_score, diff_score, indel_score, remove_contained != 0, policy), out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

int orc_run_stats(const char* paf, size_t paf_n, int qbed, char** out, size_t* out_n, char* err,
                  size_t err_cap) {
    try {
        *out = dup_out(run_stats(paf, paf_n, qbed != 0), out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

// `rb liftover .. | rb stats --paf` with wall-clock seconds of each half (CPU baseline leg).
int orc_bench_pipeline(const char* paf, size_t paf_n, const char* bed, size_t bed_n, int policy,
                       int threads, double* secs_liftover, double* secs_stats, uint64_t* n_rows,
                       uint64_t* out_bytes, char* err, size_t err_cap) {
    try {
        auto t0 = std::chrono::steady_clock::now();
        std::string lifted = run_liftover(paf, paf_n, bed, bed_n, false, false, policy, threads);
        auto t1 = std::chrono::steady_clock::now();
        std::string st = run_stats(lifted.data(), lifted.size(), false);
        auto t2 = std::chrono::steady_clock::now();
        uint64_t rows = 0;
        for (char c : lifted) rows += (c == '\n');
        *secs_liftover = std::chrono::duration<double>(t1 - t0).count();
        *secs_stats = std::chrono::duration<double>(t2 - t1).count();
        *n_rows = rows;
        *out_bytes = lifted.size();
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

// The same pipeline, keeping the bytes both halves print (bench.py compares them with the GPU rows of the same contigs).
int orc_bench_pipeline_keep(const char* paf, size_t paf_n, const char* bed, size_t bed_n, int policy, int threads,
                            double* secs_liftover, double* secs_stats, uint64_t* n_rows, char** lifted_out, size_t* lifted_n,
                            char** stats_out, size_t* stats_n, char* err, size_t err_cap) {
    try {
        auto t0 = std::chrono::steady_clock::now();
        std::string lifted = run_liftover(paf, paf_n, bed, bed_n, false, false, policy, threads);
        auto t1 = std::chrono::steady_clock::now();
        std::string st = run_stats(lifted.data(), lifted.size(), false);
        auto t2 = std::chrono::steady_clock::now();
        uint64_t rows = 0;
        for (char c : lifted) rows += (c == '\n');
        *secs_liftover = std::chrono::duration<double>(t1 - t0).count();
        *secs_stats = std::chrono::duration<double>(t2 - t1).count();
        *n_rows = rows;
        *lifted_out = dup_out(lifted, lifted_n);
        *stats_out = dup_out(st, stats_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

// one (record line, region) pair through aligned_pairs + trim_paf_rec_to_rgn
int orc_trim_line(const char* paf_line, const char* rgn_name, uint64_t st, uint64_t en,
                  const char* rgn_id, int policy, char** out, size_t* out_n, char* err,
                  size_t err_cap) {
    try {
        PafRecord rec = PafRecord::parse(paf_line);
        rec.aligned_pairs();
        Region r;
        r.name = rgn_name; r.st = st; r.en = en; r.id = rgn_id;
        PafRecord t;
        if (!trim_paf_rec_to_rgn(r, rec, policy, t)) return 1;
        *out = dup_out(t.to_line(), out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    } catch (const ParseSkip&) {
        set_err(err, err_cap, "ParseSkip");
        return 2;
    }
}

int orc_break_paf(const char* paf_line, uint32_t break_length, int policy, char** out, size_t* out_n,
                  char* err, size_t err_cap) {
    try {
        PafRecord rec = PafRecord::parse(paf_line);
        rec.aligned_pairs();
        std::string s;
        for (const PafRecord& p : break_paf_on_indels(rec, break_length, policy)) {
            s += p.to_line();
            s += '\n';
        }
        *out = dup_out(s, out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

// aligned_pairs() on a line, return the record's cigar string afterwards (paf.rs:375-376)
int orc_aligned_pairs_cigar(const char* paf_line, char** out, size_t* out_n, char* err, size_t err_cap) {
    try {
        PafRecord rec = PafRecord::parse(paf_line);
        rec.aligned_pairs();
        *out = dup_out(rec.to_line(), out_n);
        return 0;
    } catch (const Abort& e) {
        set_err(err, err_cap, e.what());
        return 101;
    }
}

int orc_parse_cigar(const char* s, size_t n, uint32_t* lens, uint8_t* ops, size_t cap, size_t* n_ops) {
    try {
        CigarString c = parse_cigar(s, n);
        *n_ops = c.size();
        for (size_t i = 0; i < c.size() && i < cap; i++) {
            lens[i] = c[i].len;
            ops[i] = (uint8_t)OP_CHARS[c[i].op];
        }
        return 0;
    } catch (const Abort&) {
        return 101;
    }
}

// counts[7] = equal diff ins del matches ins_events del_events ; ids[3] = by_matches by_events by_all
int orc_cigar_stats(const char* s, size_t n, uint32_t* counts, float* ids) {
    try {
        Stats st;
        add_stats_from_cigar(parse_cigar(s, n), st);
        counts[0] = st.equal; counts[1] = st.diff; counts[2] = st.ins; counts[3] = st.del;
        counts[4] = st.matches; counts[5] = st.ins_events; counts[6] = st.del_events;
        ids[0] = st.id_by_matches; ids[1] = st.id_by_events; ids[2] = st.id_by_all;
        return 0;
    } catch (const Abort&) {
        return 101;
    }
}

void orc_fmt_f32(float v, char* buf, size_t cap) {
    std::string s = fmt_f32(v);
    strncpy(buf, s.c_str(), cap - 1);
    buf[cap - 1] = 0;
}

int orc_parse_bed(const char* bed, size_t n, char** out, size_t* out_n) {
    std::string s;
    for (const Region& r : parse_bed_text(bed, n))
        s += r.name + "\t" + std::to_string(r.st) + "\t" + std::to_string(r.en) + "\t" + r.id + "\n";
    *out = dup_out(s, out_n);
    return 0;
}

}  // extern "C"
