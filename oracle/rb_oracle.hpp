// =============================================================================
// rb_oracle — CPU restatement of rustybam's PAF liftover + stats path.
//
// *** TEST INFRASTRUCTURE ONLY ***  Nothing under rustybam_b200/ may include,
// link or call this code.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker (or
// as the timed CPU baseline), never as the product.
//
// What it restates (file:line into the reference tree, rustybam v0.1.33):
//   src/paf.rs:379-430   PafRecord::new           -> PafRecord::parse
//   src/paf.rs:62-78     Paf::from_file           -> Paf::from_text
//   src/paf.rs:501-538   aligned_pairs            -> PafRecord::aligned_pairs   (LITERAL, per base, 24 B/column)
//   src/paf.rs:541-561   tpos_to_idx(_match)      -> PafRecord::tpos_to_idx(_match)
//   src/paf.rs:593-620   subset/collapse          -> subset_cigar / collapse_long_cigar
//   src/paf.rs:622-627   paf_overlaps_rgn         -> PafRecord::overlaps
//   src/paf.rs:631-654   infer_n_bases            -> PafRecord::infer_n_bases   (u32 wrapping accumulators)
//   src/paf.rs:656-783   remove_trailing_indels   -> PafRecord::remove_trailing_indels
//   src/paf.rs:825-857   check_integrity          -> PafRecord::check_integrity
//   src/paf.rs:923-944   Display for PafRecord    -> PafRecord::to_line
//   src/paf.rs:946-975   op classes               -> consumes_reference / consumes_query / is_match
//   src/paf.rs:1050-1094 swap query/target        -> paf_swap_query_and_target
//   src/liftover.rs:17-105   trim_paf_rec_to_rgn  -> trim_paf_rec_to_rgn
//   src/liftover.rs:107-167  trim_helper/_by_rgns -> trim_helper / trim_paf_by_rgns
//   src/liftover.rs:182-226  break_paf_on_indels  -> break_paf_on_indels
//   src/paf.rs:564-591       qpos_to_idx(_match)  -> PafRecord::qpos_to_idx(_match)
//   src/paf.rs:785-823       truncate_record_by_query -> PafRecord::truncate_record_by_query
//   src/paf.rs:210-305       overlapping_paf_recs -> Paf::overlapping_paf_recs            (LITERAL rounds)
//   src/trim_overlap.rs:6-86 score / split        -> score_of_qpos / trim_overlapping_pafs (LITERAL, per base)
//   src/bamstats.rs:91-154   stats                -> stats_from_paf / add_stats_from_cigar (u32 + f32)
//   src/bamstats.rs:225-270  printers             -> stats_header / stats_row
//   src/bed.rs:140-194       BED -> Region        -> parse_bed_text
//   src/main.rs:50-58,186-214 drivers             -> run_stats / run_liftover
//
// Arithmetic that lives in UNVENDORED third-party crates (Cargo.lock pins), restated
// from their published behaviour:
//   rust-htslib 0.44.1  CigarString::try_from(&[u8]) / Display   -> parse_cigar / cigar_to_string
//   core::slice::binary_search (toolchain dependent)             -> bsearch_rightmost / bsearch_early_exit
//   bio 1.6.0 bed::Reader (csv, tab, '#' comments)               -> parse_bed_text
//   Rust f32 Display (shortest round trip, positional)           -> fmt_f32
//
// PARITY PINNING: the reference cannot be built here (no cargo/rustc, no network).
// The oracle is pinned against every known-answer vector the reference's own tests
// hold for this path (tests/test_oracle_vectors.py lists them with file:line).
// NOT pinned by any reference test ("parity unpinned"): binary_search's choice among
// duplicate target positions, par_bridge emission order, f32 Display digits, and the
// end-to-end output of `rb liftover | rb stats --paf` on the bundled fixture.
// =============================================================================
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// Models a Rust panic (process would exit with status 101).
struct Abort : std::runtime_error {
    explicit Abort(const std::string& m) : std::runtime_error(m) {}
};

enum Op : uint8_t { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };
extern const char OP_CHARS[10];  // "MIDNSHP=X"

struct Cig {
    uint32_t len;
    uint32_t op;  // 8 bytes per unit, like the reference's `Cigar` enum
};
typedef std::vector<Cig> CigarString;

// binary_search duplicate policy
enum Policy : int { POLICY_RIGHTMOST = 0, POLICY_EARLY_EXIT = 1 };

bool consumes_reference(uint32_t op);
bool consumes_query(uint32_t op);
bool is_match(uint32_t op);

CigarString parse_cigar(const char* s, size_t n);  // throws Abort like the `expect` at paf.rs:399
std::string cigar_to_string(const CigarString& c);

struct Region {
    std::string name;
    uint64_t st = 0, en = 0;
    std::string id;
    std::string display() const;  // bed.rs:41-45
};

struct ParseSkip {};  // PafRecord::new returned Err -> the line is skipped (paf.rs:73)

struct PafRecord {
    std::string q_name;
    uint64_t q_len = 0, q_st = 0, q_en = 0;
    char strand = '+';
    std::string t_name;
    uint64_t t_len = 0, t_st = 0, t_en = 0;
    uint64_t nmatch = 0, aln_len = 0, mapq = 0;
    CigarString cigar;
    std::string tags;
    std::vector<uint64_t> tpos_aln, qpos_aln;
    CigarString long_cigar;
    std::string id;

    static PafRecord parse(const std::string& line);  // throws ParseSkip or Abort
    PafRecord small_copy() const;
    void aligned_pairs();
    // Ok(idx) -> returns true and sets idx ; Err(_) -> false
    bool tpos_to_idx(uint64_t tpos, int policy, size_t& idx) const;
    bool tpos_to_idx_match(uint64_t tpos, bool search_right, int policy, size_t& idx) const;
    // paf.rs:564-591 (query space; the array runs downward on '-' records, binary_search_by with a reversed comparator)
    bool qpos_to_idx(uint64_t qpos, int policy, size_t& idx) const;
    bool qpos_to_idx_match(uint64_t qpos, bool search_right, int policy, size_t& idx) const;
    void make_long_cigar();                                                             // paf.rs:489-498
    void truncate_record_by_query(uint64_t new_q_st, uint64_t new_q_en, int policy);   // paf.rs:785-823
    CigarString subset_cigar(size_t start_idx, size_t end_idx) const;
    static CigarString collapse_long_cigar(const CigarString& c);
    bool overlaps(const Region& r) const;
    void infer_n_bases(uint64_t& t, uint64_t& q, uint64_t& nm, uint64_t& al) const;
    void remove_trailing_indels();
    bool check_integrity(std::string* why = nullptr);  // false == Err(..)
    std::string to_line() const;
};

struct Paf {
    std::vector<PafRecord> records;
    // `skipped` counts lines that hit ParseSkip (stderr note in the reference)
    static Paf from_text(const char* text, size_t n, size_t* skipped = nullptr);
    // paf.rs:210-305 — `rb trim-paf`: rounds of pairwise overlap trimming in query space
    void overlapping_paf_recs(int match_score, int diff_score, int indel_score, bool remove_contained, int policy);
    // the same, one recursion level at a time (the multi-rank tests step several sets in lockstep)
    size_t trim_round(int match_score, int diff_score, int indel_score, int policy, std::vector<bool>& contained);
    void drop_contained(const std::vector<bool>& contained);
};

std::vector<Region> parse_bed_text(const char* text, size_t n);

bool trim_paf_rec_to_rgn(const Region& rgn, const PafRecord& paf, int policy, PafRecord& out);
std::vector<PafRecord> trim_helper(const std::string& name, const std::vector<PafRecord>& recs,
                                   const std::vector<Region>& rgns, int policy, int threads);
std::vector<PafRecord> trim_paf_by_rgns(const std::vector<Region>& rgns,
                                        const std::vector<PafRecord>& recs, bool invert_query,
                                        int policy, int threads);
std::vector<PafRecord> break_paf_on_indels(const PafRecord& paf, uint32_t break_length, int policy);
PafRecord paf_swap_query_and_target(const PafRecord& paf);
// trim_overlap.rs:6-86 — split point of two query-overlapping records by cumulative per-base scores, then truncation
int score_of_qpos(const PafRecord& rec, uint64_t pos, int match_score, int diff_score, int indel_score, int policy);
void trim_overlapping_pafs(PafRecord& left, PafRecord& right, int match_score, int diff_score, int indel_score, int policy);

struct Stats {
    std::string q_nm, r_nm;
    int64_t q_len = 0, q_st = 0, q_en = 0, r_len = 0, r_st = 0, r_en = 0;
    char strand = '\0';
    uint32_t equal = 0, diff = 0, ins = 0, del = 0, matches = 0, ins_events = 0, del_events = 0;
    float id_by_all = 0, id_by_events = 0, id_by_matches = 0;
};
void add_stats_from_cigar(const CigarString& c, Stats& s);
Stats stats_from_paf(const PafRecord& p);
std::string stats_header(bool qbed);
std::string stats_row(const Stats& s, bool qbed);
std::string fmt_f32(float v);  // Rust `{}` for f32

// main.rs drivers: whole-command restatements returning what the reference prints on stdout
std::string run_liftover(const char* paf, size_t paf_n, const char* bed, size_t bed_n, bool qbed,
                         bool largest, int policy, int threads);
std::string run_stats(const char* paf, size_t paf_n, bool qbed);
std::string run_break_paf(const char* paf, size_t paf_n, uint32_t max_size, int policy);
std::string run_invert(const char* paf, size_t paf_n);
std::string run_trim_paf(const char* paf, size_t paf_n, int match_score, int diff_score, int indel_score,
                         bool remove_contained, int policy);

}  // namespace orc
