// C face of the host helpers for ctypes (tests, bench.py): packed PAF / window builders and the
// synthetic generator.  No CUDA here; the buffers these return are handed to rbcuda.h calls.
#include <cstdlib>
#include <cstring>

#include "rbhost.hpp"

using namespace rbh;

static char* dup_str(const std::string& s, size_t* n) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    if (n) *n = s.size();
    return p;
}

extern "C" {

void* rbh_synth_paf(uint64_t seed, double scale, int n_hap, int threads) {
    SynthParams p;
    p.seed = seed; p.scale = scale; p.n_hap = n_hap; p.threads = threads;
    return new Paf(synth_paf(p));
}
void* rbh_synth_paf_mask(uint64_t seed, double scale, int n_hap, int threads, uint32_t contig_mask) {
    SynthParams p;
    p.seed = seed; p.scale = scale; p.n_hap = n_hap; p.threads = threads; p.contig_mask = contig_mask;
    return new Paf(synth_paf(p));
}
void* rbh_paf_from_text(const char* text, size_t n, char* err, size_t err_cap) {
    try {
        return new Paf(Paf::from_text(text, n));
    } catch (const Panic& e) {
        if (err && err_cap) { strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return nullptr;
    }
}
// paf.rs:62-78 Paf::from_file: plain files are mapped and parsed in place, "-" / .gz / .bgz go through read_all
void* rbh_paf_from_file(const char* path, char* err, size_t err_cap) {
    try {
        return new Paf(Paf::from_file(path));
    } catch (const Panic& e) {
        if (err && err_cap) { strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return nullptr;
    }
}
// file contents as the readers of myio.rs:41-64 deliver them (plain / .gz / .bgz); nullptr on I/O errors
char* rbh_read_all(const char* path, size_t* n) {
    try {
        return dup_str(read_all(path), n);
    } catch (const Panic&) {
        return nullptr;
    }
}
void rbh_paf_view(void* paf, rb_records* out) { *out = static_cast<Paf*>(paf)->view(); }
uint64_t rbh_paf_size(void* paf) { return static_cast<Paf*>(paf)->size(); }
uint64_t rbh_paf_skipped(void* paf) { return static_cast<Paf*>(paf)->skipped; }
void rbh_paf_free(void* paf) { delete static_cast<Paf*>(paf); }
int64_t rbh_paf_find_name(void* paf, const char* name) { return static_cast<Paf*>(paf)->find_name(name); }
char* rbh_paf_text(void* paf, uint64_t lo, uint64_t hi, size_t* n) { return dup_str(paf_text(*static_cast<Paf*>(paf), lo, hi), n); }
// PAF text of every record whose target is name id `tid`
char* rbh_paf_text_of_contig(void* pafv, uint32_t tid, size_t* n, uint64_t* n_rec) {
    Paf& paf = *static_cast<Paf*>(pafv);
    std::string s;
    uint64_t cnt = 0;
    for (size_t i = 0; i < paf.size(); i++)
        if (paf.t_id[i] == tid) { s += paf_text(paf, i, i + 1); cnt++; }
    if (n_rec) *n_rec = cnt;
    return dup_str(s, n);
}

void* rbh_tiling_windows(void* paf, uint64_t width) { return new Windows(tiling_windows_packed(*static_cast<Paf*>(paf), width)); }
void* rbh_windows_from_bed_text(void* paf, const char* bed, size_t n) {
    return new Windows(Windows::pack_text(bed, n, *static_cast<Paf*>(paf)));
}
// the two-step form (a Region per row), kept as the statement the fast one is compared with
void* rbh_windows_from_bed_text_slow(void* paf, const char* bed, size_t n) {
    return new Windows(Windows::pack(parse_bed_text(bed, n), *static_cast<Paf*>(paf)));
}
void rbh_windows_view(void* w, rb_windows* out) { *out = static_cast<Windows*>(w)->view(); }
void rbh_windows_free(void* w) { delete static_cast<Windows*>(w); }
// 3-column tiling BED text; tid < 0 = all contigs, else only that target name id
char* rbh_tiling_bed_text(void* pafv, uint64_t width, int64_t tid, size_t* n) {
    Paf& paf = *static_cast<Paf*>(pafv);
    std::vector<Region> all = tiling_windows(paf, width), sel;
    for (Region& r : all)
        if (tid < 0 || paf.find_name(r.name) == tid) sel.push_back(std::move(r));
    return dup_str(bed_text(sel, false), n);
}
// `rb stats --paf` rows (bamstats.rs:239-270) of the records of `paf` with the counters st_cols[0..7) (equal diff ins del ins_ev
// del_ev matches) and identities id_cols[0..3) (by_matches by_events by_all) of rows row0 .. row0 + paf.size()
char* rbh_stats_text(void* pafv, const uint32_t* const* st_cols, const float* const* id_cols, uint64_t row0, int qbed, int header, size_t* n) {
    Paf& paf = *static_cast<Paf*>(pafv);
    rb_stats_out st{};
    st.n = paf.size();
    st.equal = const_cast<uint32_t*>(st_cols[0]) + row0; st.diff = const_cast<uint32_t*>(st_cols[1]) + row0;
    st.ins = const_cast<uint32_t*>(st_cols[2]) + row0; st.del = const_cast<uint32_t*>(st_cols[3]) + row0;
    st.ins_events = const_cast<uint32_t*>(st_cols[4]) + row0; st.del_events = const_cast<uint32_t*>(st_cols[5]) + row0;
    st.matches = const_cast<uint32_t*>(st_cols[6]) + row0;
    st.id_by_matches = const_cast<float*>(id_cols[0]) + row0; st.id_by_events = const_cast<float*>(id_cols[1]) + row0;
    st.id_by_all = const_cast<float*>(id_cols[2]) + row0;
    std::string out = header ? stats_header(qbed != 0) : std::string();
    for (size_t i = 0; i < paf.size(); i++) append_stats_row(out, paf, i, st, qbed != 0);
    return dup_str(out, n);
}
void rbh_free_str(char* p) { free(p); }
void rbh_fmt_f32(float v, char* buf, size_t cap) {
    std::string s = fmt_f32(v);
    strncpy(buf, s.c_str(), cap - 1);
    buf[cap - 1] = 0;
}

}  // extern "C"
