// rb_oracle.cpp — see rb_oracle.hpp.  TEST INFRASTRUCTURE ONLY (checker / CPU baseline).
// Every function cites the reference lines it restates (paths relative to the
// reference checkout, rustybam v0.1.33).
#include "rb_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>

namespace orc {

const char OP_CHARS[10] = "MIDNSHP=X";

// paf.rs:946-951
bool consumes_reference(uint32_t op) {
    return op == OP_M || op == OP_D || op == OP_N || op == OP_X || op == OP_EQ;
}
// paf.rs:958-963
bool consumes_query(uint32_t op) {
    return op == OP_M || op == OP_I || op == OP_S || op == OP_X || op == OP_EQ;
}
// paf.rs:973-975
bool is_match(uint32_t op) { return op == OP_M || op == OP_X || op == OP_EQ; }

static inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// rust-htslib 0.44.1 `impl TryFrom<&[u8]> for CigarString` (unvendored dep; called at
// paf.rs:398-399 behind `.expect(..)`, so every error is a panic).  Digits -> u32
// (`str::parse::<u32>`: overflow is an error, leading zeros are fine), then one op byte
// out of MIDNSHP=X.  H only as first or last op; S only at the ends or next to H.
CigarString parse_cigar(const char* s, size_t n) {
    CigarString out;
    size_t i = 0;
    while (i < n) {
        size_t j = i;
        while (j < n && is_digit(s[j])) j++;
        if (i == j) throw Abort("Unable to parse cigar string.: expected length before cigar operation");
        uint64_t v = 0;
        bool ovf = false;
        for (size_t k = i; k < j; k++) {
            v = v * 10 + (uint64_t)(s[k] - '0');
            if (v > 0xFFFFFFFFull) ovf = true, v = 0x100000000ull;
        }
        if (ovf) throw Abort("Unable to parse cigar string.: length does not fit u32");
        if (j >= n) throw Abort("index out of bounds: cigar ends in a number");  // `&bytes[j]`
        char op = s[j];
        uint32_t code;
        switch (op) {
            case 'M': code = OP_M; break;
            case 'I': code = OP_I; break;
            case 'D': code = OP_D; break;
            case 'N': code = OP_N; break;
            case 'H':
                if (i == 0 || j + 1 == n) code = OP_H;
                else throw Abort("Unable to parse cigar string.: H only valid at start or end");
                break;
            case 'S': {
                bool ok = (i == 0) || (j + 1 == n) || (s[i - 1] == 'H');
                if (!ok) {
                    ok = true;
                    for (size_t k = j + 1; k < n; k++)
                        if (!(is_digit(s[k]) || s[k] == 'H')) { ok = false; break; }
                }
                if (!ok) throw Abort("Unable to parse cigar string.: S only valid at the ends");
                code = OP_S;
                break;
            }
            case 'P': code = OP_P; break;
            case '=': code = OP_EQ; break;
            case 'X': code = OP_X; break;
            default: throw Abort(std::string("Unable to parse cigar string.: bad op '") + op + "'");
        }
        out.push_back(Cig{(uint32_t)v, code});
        i = j + 1;
    }
    return out;
}

// rust-htslib `impl fmt::Display for CigarString`: "{len}{char}" per op, concatenated.
std::string cigar_to_string(const CigarString& c) {
    std::string s;
    s.reserve(c.size() * 4);
    char buf[16];
    for (const Cig& x : c) {
        int k = snprintf(buf, sizeof buf, "%u", x.len);
        s.append(buf, k);
        s.push_back(OP_CHARS[x.op]);
    }
    return s;
}

// bed.rs:41-45
std::string Region::display() const {
    return name + ":" + std::to_string(st + 1) + "-" + std::to_string(en);
}

// ---- Rust `str::parse::<u64>` : optional '+', >=1 digit, overflow is an error ----------
static bool parse_u64(const std::string& t, uint64_t& out) {
    size_t i = 0;
    if (!t.empty() && t[0] == '+') i = 1;
    if (i >= t.size()) return false;
    uint64_t v = 0;
    for (; i < t.size(); i++) {
        if (!is_digit(t[i])) return false;
        uint64_t d = (uint64_t)(t[i] - '0');
        if (v > (UINT64_MAX - d) / 10) return false;
        v = v * 10 + d;
    }
    out = v;
    return true;
}

static inline bool is_ascii_ws(char c) {  // char::is_ascii_whitespace
    return c == ' ' || c == '\t' || c == '\n' || c == '\x0C' || c == '\r';
}

// paf.rs:379-430.  Order of effects matters: tag/CIGAR handling (asserts, `expect`) runs
// before the numeric columns are parsed, so a bad CIGAR panics even on a line whose
// numeric columns would have been skipped.
PafRecord PafRecord::parse(const std::string& line) {
    std::vector<std::string> t;
    size_t i = 0, n = line.size();
    while (i < n) {
        while (i < n && is_ascii_ws(line[i])) i++;
        size_t j = i;
        while (j < n && !is_ascii_ws(line[j])) j++;
        if (j > i) t.emplace_back(line, i, j - i);
        i = j;
    }
    if (t.size() < 12) throw Abort("assertion failed: t.len() >= 12");
    PafRecord r;
    for (size_t k = 12; k < t.size(); k++) {
        const std::string& tok = t[k];
        // regex "(..):(.):(.*)" — unanchored, leftmost match (bytes; ASCII assumed)
        size_t m = std::string::npos;
        for (size_t p = 0; p + 5 <= tok.size(); p++)
            if (tok[p + 2] == ':' && tok[p + 4] == ':') { m = p; break; }
        if (m == std::string::npos) throw Abort("assertion failed: PAF_TAG.is_match(token)");
        if (tok.compare(m, 2, "cg") == 0 && r.cigar.empty()) {
            r.cigar = parse_cigar(tok.data() + m + 5, tok.size() - (m + 5));
        } else {
            r.tags.push_back('\t');
            r.tags += tok;
        }
    }
    r.q_name = t[0];
    bool ok = parse_u64(t[1], r.q_len) && parse_u64(t[2], r.q_st) && parse_u64(t[3], r.q_en);
    if (ok) {
        if (t[4].size() != 1) ok = false;  // parse::<char>()
        else r.strand = t[4][0];
    }
    r.t_name = t[5];
    ok = ok && parse_u64(t[6], r.t_len) && parse_u64(t[7], r.t_st) && parse_u64(t[8], r.t_en) &&
         parse_u64(t[9], r.nmatch) && parse_u64(t[10], r.aln_len) && parse_u64(t[11], r.mapq);
    if (!ok) throw ParseSkip{};
    return r;
}

// paf.rs:433-456
PafRecord PafRecord::small_copy() const {
    PafRecord c;
    c.q_name = q_name; c.q_len = q_len; c.q_st = q_st; c.q_en = q_en; c.strand = strand;
    c.t_name = t_name; c.t_len = t_len; c.t_st = t_st; c.t_en = t_en;
    c.nmatch = nmatch; c.aln_len = aln_len; c.mapq = mapq;
    c.tags = tags; c.id = id;
    return c;
}

// paf.rs:631-654 — u32 accumulators (release build: wrapping), widened on return.
void PafRecord::infer_n_bases(uint64_t& t, uint64_t& q, uint64_t& nm, uint64_t& al) const {
    uint32_t tb = 0, qb = 0, m = 0, a = 0;
    for (const Cig& o : cigar) {
        if (consumes_reference(o.op)) tb += o.len;
        if (consumes_query(o.op)) qb += o.len;
        if (is_match(o.op)) m += o.len;
        a += o.len;
    }
    t = tb; q = qb; nm = m; al = a;
}

// paf.rs:825-857 — verifies the spans and OVERWRITES nmatch / aln_len.
bool PafRecord::check_integrity(std::string* why) {
    uint64_t tb, qb, nm, al;
    infer_n_bases(tb, qb, nm, al);
    if (t_en - t_st != tb) {
        if (why) *why = "target bases " + std::to_string(tb) + " from cigar does not equal " +
                        std::to_string(t_en) + "-" + std::to_string(t_st);
        return false;
    }
    if (q_en - q_st != qb) {
        if (why) *why = "query bases " + std::to_string(qb) + " from cigar does not equal " +
                        std::to_string(q_en) + "-" + std::to_string(q_st);
        return false;
    }
    nmatch = nm;
    aln_len = al;
    return true;
}

// paf.rs:656-783
void PafRecord::remove_trailing_indels() {
    const size_t cigar_len = cigar.size();
    if (cigar_len == 0) throw Abort("called `Option::unwrap()` on a `None` value (empty cigar)");

    Cig st_opt = cigar.front();
    uint32_t remove_st_t = 0, remove_st_q = 0;
    size_t remove_st_opts = 0;
    CigarString removed_st;
    while (st_opt.op == OP_I || st_opt.op == OP_D) {
        if (st_opt.op == OP_D) {
            remove_st_t += st_opt.len;
            remove_st_q += 1;  // reference quirk (paf.rs:672-673)
        } else {
            remove_st_q += st_opt.len;
        }
        remove_st_opts++;
        removed_st.push_back(st_opt);
        if (remove_st_opts < cigar_len) st_opt = cigar[remove_st_opts];
        else break;
    }
    if (removed_st.size() > 1) {
        for (size_t k = 0; k + 1 < removed_st.size(); k++) {
            uint32_t a = removed_st[k].op, b = removed_st[k + 1].op;
            if ((a == OP_D && b == OP_I) || (a == OP_I && b == OP_D)) {
                remove_st_t += 1;
                remove_st_q -= 1;
            }
        }
    }

    Cig en_opt = cigar.back();
    uint32_t remove_en_t = 0, remove_en_q = 0;
    size_t remove_en_opts = 0;
    CigarString removed_en;
    while (en_opt.op == OP_I || en_opt.op == OP_D) {
        if (en_opt.op == OP_D) remove_en_t += en_opt.len;
        else remove_en_q += en_opt.len;
        remove_en_opts++;
        removed_en.push_back(en_opt);
        if (cigar_len - remove_en_opts > 0) en_opt = cigar[cigar_len - 1 - remove_en_opts];
        else break;
    }

    if (remove_en_opts > 0 || remove_st_opts > 0)
        id += "_TO." + cigar_to_string(removed_st) + "." + cigar_to_string(removed_en);

    cigar.erase(cigar.begin(), cigar.begin() + (ptrdiff_t)remove_st_opts);
    if (cigar.size() >= remove_en_opts) cigar.resize(cigar.size() - remove_en_opts);
    else  // all-indel CIGAR: `self.cigar.len() - remove_en_opts` underflows usize (paf.rs:757) — a panic in
          // debug builds; a release build would carry on with wrapped spans.  Modelled as the panic.
        throw Abort("attempt to subtract with overflow (all-indel CIGAR, paf.rs:757)");

    t_st += remove_st_t;
    t_en -= remove_en_t;
    if (strand == '-') std::swap(remove_st_q, remove_en_q);
    q_st += remove_st_q;
    q_en -= remove_en_q;

    std::string why;
    if (!check_integrity(&why)) throw Abort("remove_trailing_indels: check_integrity failed: " + why);
}

// paf.rs:501-538 — LITERAL per-base expansion: 8 B tpos + 8 B qpos + 8 B long_cigar per column.
void PafRecord::aligned_pairs() {
    remove_trailing_indels();
    int64_t t_pos = (int64_t)t_st - 1;
    int64_t q_pos = (int64_t)q_st - 1;
    long_cigar.clear();
    tpos_aln.clear();
    qpos_aln.clear();
    if (strand == '-') q_pos = (int64_t)q_en;
    for (const Cig& o : cigar) {
        const bool moves_t = consumes_reference(o.op);
        const bool moves_q = consumes_query(o.op);
        for (uint32_t k = 0; k < o.len; k++) {
            long_cigar.push_back(Cig{1, o.op});
            if (moves_t) t_pos += 1;
            if (moves_q && strand == '+') q_pos += 1;
            if (moves_q && strand == '-') q_pos -= 1;
            tpos_aln.push_back((uint64_t)t_pos);
            qpos_aln.push_back((uint64_t)q_pos);
        }
    }
}

// core::slice::binary_search as shipped by Rust < 1.52 and >= 1.82: no early exit; `base`
// moves to `mid` on Less OR Equal, so among equal elements the RIGHT-MOST one is returned.
static bool bsearch_rightmost(const std::vector<uint64_t>& a, uint64_t x, size_t& idx) {
    size_t size = a.size();
    if (size == 0) return false;
    size_t base = 0;
    while (size > 1) {
        size_t half = size / 2, mid = base + half;
        if (!(a[mid] > x)) base = mid;
        size -= half;
    }
    if (a[base] == x) { idx = base; return true; }
    return false;
}
// core::slice::binary_search as shipped by Rust 1.52 ..= 1.81: returns the first probe
// (mid = left + size/2) that compares Equal.
static bool bsearch_early_exit(const std::vector<uint64_t>& a, uint64_t x, size_t& idx) {
    size_t size = a.size(), left = 0, right = size;
    while (left < right) {
        size_t mid = left + size / 2;
        if (a[mid] < x) left = mid + 1;
        else if (a[mid] > x) right = mid;
        else { idx = mid; return true; }
        size = right - left;
    }
    return false;
}

// paf.rs:541-544
bool PafRecord::tpos_to_idx(uint64_t tpos, int policy, size_t& idx) const {
    return policy == POLICY_EARLY_EXIT ? bsearch_early_exit(tpos_aln, tpos, idx)
                                       : bsearch_rightmost(tpos_aln, tpos, idx);
}

// paf.rs:547-561
bool PafRecord::tpos_to_idx_match(uint64_t tpos, bool search_right, int policy, size_t& idx) const {
    if (!tpos_to_idx(tpos, policy, idx)) return false;
    const size_t max_idx = long_cigar.size();
    if (search_right) {
        while (idx < max_idx && !is_match(long_cigar[idx].op)) idx++;
    } else {
        while (idx > 0 && !is_match(long_cigar[idx].op)) idx--;
    }
    return true;
}

// ---- rb trim-paf (query space) -------------------------------------------------------------------
// core::slice::binary_search_by with an arbitrary three-way comparator `cmp(probe)` (<0 Less, 0 Equal, >0 Greater),
// in the two shapes std has shipped (see bsearch_rightmost / bsearch_early_exit above).
template <class F>
static bool bsearch_by(size_t n, int policy, F&& cmp, size_t& idx) {
    if (policy == POLICY_EARLY_EXIT) {
        size_t size = n, left = 0, right = n;
        while (left < right) {
            size_t mid = left + size / 2;
            const int c = cmp(mid);
            if (c < 0) left = mid + 1;
            else if (c > 0) right = mid;
            else { idx = mid; return true; }
            size = right - left;
        }
        return false;
    }
    size_t size = n;
    if (size == 0) return false;
    size_t base = 0;
    while (size > 1) {
        size_t half = size / 2, mid = base + half;
        if (!(cmp(mid) > 0)) base = mid;
        size -= half;
    }
    if (cmp(base) == 0) { idx = base; return true; }
    return false;
}

// paf.rs:564-574
bool PafRecord::qpos_to_idx(uint64_t qpos, int policy, size_t& idx) const {
    const std::vector<uint64_t>& a = qpos_aln;
    if (strand == '-')  // probe.cmp(&qpos).reverse()
        return bsearch_by(a.size(), policy, [&](size_t m) { return a[m] > qpos ? -1 : (a[m] < qpos ? 1 : 0); }, idx);
    return bsearch_by(a.size(), policy, [&](size_t m) { return a[m] < qpos ? -1 : (a[m] > qpos ? 1 : 0); }, idx);
}

// paf.rs:577-591
bool PafRecord::qpos_to_idx_match(uint64_t qpos, bool search_right, int policy, size_t& idx) const {
    if (!qpos_to_idx(qpos, policy, idx)) return false;
    const size_t max_idx = long_cigar.size();
    if ((search_right && strand == '+') || (!search_right && strand == '-')) {
        while (idx < max_idx && !is_match(long_cigar[idx].op)) idx++;
    } else {
        while (idx > 0 && !is_match(long_cigar[idx].op)) idx--;
    }
    return true;
}

// paf.rs:489-498
void PafRecord::make_long_cigar() {
    long_cigar.clear();
    for (const Cig& o : cigar)
        for (uint32_t k = 0; k < o.len; k++) long_cigar.push_back(Cig{1, o.op});
}

// paf.rs:785-823.  Every panic of the reference (assert!, unwrap on Err, slice index out of range, integer
// underflow in a debug build / the search miss it causes in a release build) is an Abort here.
void PafRecord::truncate_record_by_query(uint64_t new_q_st, uint64_t new_q_en, int policy) {
    if (!(new_q_st >= q_st)) throw Abort("New start is less than old start.");
    if (!(new_q_en <= q_en)) throw Abort("New end is greater than old end.");
    make_long_cigar();
    if (new_q_en == 0) throw Abort("truncate_record_by_query: new_q_en - 1 underflows");
    size_t aln_st, aln_en;
    if (!qpos_to_idx_match(new_q_st, true, policy, aln_st)) throw Abort("truncate_record_by_query: qpos_to_idx_match(start) is Err");
    if (!qpos_to_idx_match(new_q_en - 1, false, policy, aln_en)) throw Abort("truncate_record_by_query: qpos_to_idx_match(end) is Err");
    if (aln_st >= qpos_aln.size() || aln_en >= qpos_aln.size()) throw Abort("truncate_record_by_query: index out of bounds");
    const uint64_t new_new_q_st = qpos_aln[aln_st];
    const uint64_t new_new_q_en = qpos_aln[aln_en] + 1;
    if (aln_st > aln_en) std::swap(aln_st, aln_en);
    const uint64_t new_t_st = tpos_aln[aln_st];
    const uint64_t new_t_en = tpos_aln[aln_en] + 1;
    long_cigar = subset_cigar(aln_st, aln_en);
    cigar = collapse_long_cigar(long_cigar);
    t_st = new_t_st; t_en = new_t_en;
    q_st = new_new_q_st; q_en = new_new_q_en;
    remove_trailing_indels();
    std::string why;
    if (!check_integrity(&why)) throw Abort("truncate_record_by_query: check_integrity failed: " + why);
}

// trim_overlap.rs:6-20
int score_of_qpos(const PafRecord& rec, uint64_t pos, int match_score, int diff_score, int indel_score, int policy) {
    size_t idx;
    if (!rec.qpos_to_idx(pos, policy, idx)) throw Abort("score_of_qpos: qpos_to_idx is Err");
    const uint32_t op = rec.long_cigar[idx].op;
    if (op == OP_EQ) return match_score;
    if (op == OP_I || op == OP_D) return -indel_score;
    return -diff_score;
}

// trim_overlap.rs:36-86
void trim_overlapping_pafs(PafRecord& left, PafRecord& right, int match_score, int diff_score, int indel_score, int policy) {
    const uint64_t st_ovl = std::max(left.q_st, right.q_st);
    const uint64_t en_ovl = std::min(left.q_en, right.q_en);
    std::vector<int32_t> l_score{0}, r_score;
    for (uint64_t pos = st_ovl; pos < en_ovl; pos++) {
        l_score.push_back(score_of_qpos(left, pos, match_score, diff_score, indel_score, policy));
        r_score.push_back(score_of_qpos(right, pos, match_score, diff_score, indel_score, policy));
    }
    r_score.push_back(0);
    int32_t acc = 0;
    for (size_t k = 0; k < l_score.size(); k++) { acc += l_score[k]; l_score[k] = acc; }
    acc = 0;
    for (size_t k = r_score.size(); k-- > 0;) { acc += r_score[k]; r_score[k] = acc; }
    uint64_t max_idx = 0;
    int32_t mx = 0;
    for (size_t k = 0; k < l_score.size() && k < r_score.size(); k++) {
        if (l_score[k] + r_score[k] > mx) { mx = l_score[k] + r_score[k]; max_idx = k; }
    }
    left.truncate_record_by_query(left.q_st, st_ovl + max_idx, policy);
    right.truncate_record_by_query(st_ovl + max_idx, right.q_en, policy);
}

// paf.rs:210-287 — one call of overlapping_paf_recs up to the decision to recurse: strip, sort, list the pairs, trim the first
// pair of every query name.  Returns `unseen` (pairs that have to wait for the next call); `contained` is this call's fresh
// vector (paf.rs:226).
size_t Paf::trim_round(int match_score, int diff_score, int indel_score, int policy, std::vector<bool>& contained) {
    for (PafRecord& rec : records) rec.remove_trailing_indels();
    struct Pair { uint64_t overlap; size_t i, j; };
    std::vector<Pair> overlap_pairs;
    std::stable_sort(records.begin(), records.end(), [](const PafRecord& a, const PafRecord& b) { return a.q_name < b.q_name; });
    contained.assign(records.size(), false);
    if (records.size() < 2) return 0;
    for (size_t i = 0; i + 1 < records.size(); i++) {
        const PafRecord& rec1 = records[i];
        for (size_t j = i + 1; j < records.size() && rec1.q_name == records[j].q_name; j++) {
            const PafRecord& rec2 = records[j];
            const uint64_t mn = std::min(rec1.q_en, rec2.q_en), mxs = std::max(rec1.q_st, rec2.q_st);  // bed.rs:74-85
            const uint64_t overlap = mn < mxs ? 0 : mn - mxs;
            if (overlap < 1) continue;
            else if (overlap == rec2.q_en - rec2.q_st) contained[j] = true;
            else if (overlap == rec1.q_en - rec1.q_st) contained[i] = true;
            else if (rec1.q_st <= rec2.q_st) overlap_pairs.push_back(Pair{overlap, i, j});
            else overlap_pairs.push_back(Pair{overlap, j, i});
        }
    }
    std::stable_sort(overlap_pairs.begin(), overlap_pairs.end(),
                     [](const Pair& a, const Pair& b) { return UINT64_MAX - a.overlap < UINT64_MAX - b.overlap; });
    std::vector<std::string> q_seen;  // HashSet<String>: membership only
    size_t unseen = 0;
    for (const Pair& p : overlap_pairs) {
        PafRecord left = records[p.i], right = records[p.j];
        const std::string q_name = left.q_name;
        if (std::find(q_seen.begin(), q_seen.end(), q_name) == q_seen.end()) {
            left.aligned_pairs();
            right.aligned_pairs();
            trim_overlapping_pafs(left, right, match_score, diff_score, indel_score, policy);
            records[p.i] = std::move(left);
            records[p.j] = std::move(right);
            q_seen.push_back(q_name);
        } else {
            unseen++;
        }
    }
    return unseen;
}

// paf.rs:290-300 — the `else if remove_contained` arm of the LAST call
void Paf::drop_contained(const std::vector<bool>& contained) {
    if (records.size() < 2) return;  // (paf.rs:229-231 returned before anything was flagged)
    std::vector<PafRecord> kept;
    for (size_t i = 0; i < records.size(); i++)
        if (!contained[i]) kept.push_back(records[i]);
    records.swap(kept);
}

// paf.rs:210-305 (the tail recursion written as a loop)
void Paf::overlapping_paf_recs(int match_score, int diff_score, int indel_score, bool remove_contained, int policy) {
    std::vector<bool> contained;
    while (trim_round(match_score, diff_score, indel_score, policy, contained) > 0) {}
    if (remove_contained) drop_contained(contained);
}

// paf.rs:593-600
CigarString PafRecord::subset_cigar(size_t start_idx, size_t end_idx) const {
    return CigarString(long_cigar.begin() + (ptrdiff_t)start_idx,
                       long_cigar.begin() + (ptrdiff_t)end_idx + 1);
}

// paf.rs:602-620
CigarString PafRecord::collapse_long_cigar(const CigarString& c) {
    CigarString rtn;
    Cig pre = c.at(0);
    uint32_t pre_len = 1;
    for (size_t k = 1; k < c.size(); k++) {
        if (c[k].op == pre.op) pre_len++;
        else {
            rtn.push_back(Cig{pre_len, pre.op});
            pre = c[k];
            pre_len = 1;
        }
    }
    rtn.push_back(Cig{pre_len, pre.op});
    return rtn;
}

// paf.rs:622-627
bool PafRecord::overlaps(const Region& r) const {
    if (t_name != r.name) return false;
    return t_en > r.st && t_st < r.en;
}

// paf.rs:923-944
std::string PafRecord::to_line() const {
    std::string s;
    std::string cg = cigar_to_string(cigar);
    s.reserve(cg.size() + 160 + q_name.size() + t_name.size() + id.size());
    s += q_name; s += '\t';
    s += std::to_string(q_len); s += '\t';
    s += std::to_string(q_st); s += '\t';
    s += std::to_string(q_en); s += '\t';
    s += strand; s += '\t';
    s += t_name; s += '\t';
    s += std::to_string(t_len); s += '\t';
    s += std::to_string(t_st); s += '\t';
    s += std::to_string(t_en); s += '\t';
    s += std::to_string(nmatch); s += '\t';
    s += std::to_string(aln_len); s += '\t';
    s += std::to_string(mapq);
    s += "\tid:Z:"; s += id;
    s += "\tcg:Z:"; s += cg;
    return s;
}

// paf.rs:62-78 — BufRead::lines(): split at '\n', drop one trailing '\r'.
Paf Paf::from_text(const char* text, size_t n, size_t* skipped) {
    Paf paf;
    size_t nskip = 0, i = 0;
    while (i < n) {
        const char* nl = (const char*)memchr(text + i, '\n', n - i);
        size_t j = nl ? (size_t)(nl - text) : n;
        size_t e = j;
        if (nl && e > i && text[e - 1] == '\r') e--;
        std::string line(text + i, e - i);
        try {
            PafRecord rec = PafRecord::parse(line);
            std::string why;
            if (!rec.check_integrity(&why)) throw Abort("check_integrity().unwrap(): " + why);
            paf.records.push_back(std::move(rec));
        } catch (const ParseSkip&) {
            nskip++;
        }
        i = nl ? j + 1 : n;
    }
    if (skipped) *skipped = nskip;
    return paf;
}

// bed.rs:140-194 over bio 1.6.0 `bed::Reader` (csv: tab delimiter, no header, '#' comment
// lines, NOT flexible: a row whose field count differs from the first row is an Err and
// parse_bed skips it with a warning; start/end must deserialize as u64).
std::vector<Region> parse_bed_text(const char* text, size_t n) {
    std::vector<Region> out;
    size_t i = 0;
    size_t nfields0 = 0;
    while (i < n) {
        size_t j = i;
        while (j < n && text[j] != '\n' && text[j] != '\r') j++;
        std::string line(text + i, j - i);
        i = j;
        while (i < n && (text[i] == '\n' || text[i] == '\r')) i++;  // empty lines are skipped by csv
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> f;
        size_t p = 0;
        for (;;) {
            size_t q = line.find('\t', p);
            if (q == std::string::npos) { f.emplace_back(line, p); break; }
            f.emplace_back(line, p, q - p);
            p = q + 1;
        }
        if (nfields0 == 0) nfields0 = f.size();
        else if (f.size() != nfields0) continue;  // csv UnequalLengths -> Err -> skipped
        if (f.size() < 3) continue;               // cannot deserialize (String,u64,u64,..)
        Region r;
        r.name = f[0];
        auto strict_u64 = [](const std::string& t, uint64_t& v) {  // serde/csv u64: digits only
            if (t.empty()) return false;
            uint64_t x = 0;
            for (char c : t) {
                if (!is_digit(c)) return false;
                uint64_t d = (uint64_t)(c - '0');
                if (x > (UINT64_MAX - d) / 10) return false;
                x = x * 10 + d;
            }
            v = x;
            return true;
        };
        if (!strict_u64(f[1], r.st) || !strict_u64(f[2], r.en)) continue;
        if (f.size() > 3) r.id = f[3];                   // record.name() == 4th column
        else r.id = r.name + ":" + std::to_string(r.st + 1) + "-" + std::to_string(r.en);
        out.push_back(std::move(r));
    }
    return out;
}

// liftover.rs:17-105
bool trim_paf_rec_to_rgn(const Region& rgn, const PafRecord& paf, int policy, PafRecord& out) {
    PafRecord trimmed = paf.small_copy();
    trimmed.id = rgn.id;

    if (paf.t_st > rgn.st && paf.t_en < rgn.en) {
        out = paf;  // paf.clone(): full copy including the per-base arrays
        return true;
    }

    trimmed.t_st = std::max(rgn.st, paf.t_st);
    size_t start_idx = 0, end_idx = 0;
    if (!paf.tpos_to_idx_match(trimmed.t_st, true, policy, start_idx))
        throw Abort("Problem getting index in cigar (start) for " + rgn.display());
    trimmed.t_en = std::min(rgn.en, paf.t_en);
    if (!paf.tpos_to_idx_match(trimmed.t_en - 1, false, policy, end_idx))
        throw Abort("Problem getting index in cigar (end) for " + rgn.display());

    if (start_idx > end_idx) return false;

    if (start_idx >= paf.tpos_aln.size() || end_idx >= paf.tpos_aln.size())
        throw Abort("index out of bounds in tpos_aln");
    trimmed.t_st = paf.tpos_aln[start_idx];
    trimmed.q_st = paf.qpos_aln[start_idx];
    trimmed.t_en = paf.tpos_aln[end_idx];
    trimmed.q_en = paf.qpos_aln[end_idx];

    trimmed.cigar = PafRecord::collapse_long_cigar(paf.subset_cigar(start_idx, end_idx));

    bool no_match_opts = true;
    for (const Cig& o : trimmed.cigar)
        if (is_match(o.op)) { no_match_opts = false; break; }
    if (no_match_opts) return false;

    if (paf.strand == '-') std::swap(trimmed.q_en, trimmed.q_st);
    trimmed.t_en += 1;
    trimmed.q_en += 1;

    trimmed.remove_trailing_indels();

    if (trimmed.cigar.empty()) return false;
    if (trimmed.q_st > trimmed.q_en || trimmed.t_st > trimmed.t_en) return false;
    if (!trimmed.check_integrity()) return false;
    out = std::move(trimmed);
    return true;
}

static void parallel_for(size_t n, int threads, const std::function<void(size_t)>& fn) {
    if (threads <= 1 || n < 2) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    std::string err;
    std::vector<std::thread> pool;
    int nt = (int)std::min<size_t>((size_t)threads, n);
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&] {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n || failed.load()) break;
                try { fn(i); }
                catch (const Abort& e) {
                    if (!failed.exchange(true)) err = e.what();
                }
            }
        });
    for (auto& th : pool) th.join();
    if (failed.load()) throw Abort(err);
}

// liftover.rs:107-132.  Output order is the reference's order under `-t 1` (record-major,
// BED-file-order-minor); with more threads the reference's par_bridge order is
// nondeterministic, here the results are always collected in the -t 1 order.
std::vector<PafRecord> trim_helper(const std::string& name, const std::vector<PafRecord>& recs,
                                   const std::vector<Region>& rgns, int policy, int threads) {
    std::vector<PafRecord> cur_recs;
    for (const PafRecord& r : recs)
        if (r.t_name == name) cur_recs.push_back(r);
    std::vector<const Region*> cur_rgns;
    for (const Region& g : rgns)
        if (g.name == name) cur_rgns.push_back(&g);

    parallel_for(cur_recs.size(), threads, [&](size_t i) { cur_recs[i].aligned_pairs(); });

    std::vector<std::pair<uint32_t, uint32_t>> pairs;  // cartesian product + overlap filter
    for (size_t i = 0; i < cur_recs.size(); i++)
        for (size_t j = 0; j < cur_rgns.size(); j++)
            if (cur_recs[i].overlaps(*cur_rgns[j])) pairs.emplace_back((uint32_t)i, (uint32_t)j);

    std::vector<PafRecord> res(pairs.size());
    std::vector<uint8_t> ok(pairs.size(), 0);
    parallel_for(pairs.size(), threads, [&](size_t k) {
        ok[k] = trim_paf_rec_to_rgn(*cur_rgns[pairs[k].second], cur_recs[pairs[k].first], policy, res[k]);
        // the per-base arrays of an early-return clone are never printed; drop them now
        res[k].tpos_aln = std::vector<uint64_t>();
        res[k].qpos_aln = std::vector<uint64_t>();
        res[k].long_cigar = CigarString();
    });
    std::vector<PafRecord> out;
    for (size_t k = 0; k < pairs.size(); k++)
        if (ok[k]) out.push_back(std::move(res[k]));
    return out;
}

// paf.rs:1050-1065
static CigarString cigar_swap_target_query(const CigarString& c, char strand) {
    CigarString n;
    n.reserve(c.size());
    for (const Cig& o : c) {
        Cig x = o;
        if (o.op == OP_I) x.op = OP_D;
        else if (o.op == OP_D) x.op = OP_I;
        n.push_back(x);
    }
    if (strand == '-') std::reverse(n.begin(), n.end());
    return n;
}

// paf.rs:1068-1094
PafRecord paf_swap_query_and_target(const PafRecord& paf) {
    PafRecord f = paf;
    f.t_name = paf.q_name; f.t_len = paf.q_len; f.t_st = paf.q_st; f.t_en = paf.q_en;
    f.q_name = paf.t_name; f.q_len = paf.t_len; f.q_st = paf.t_st; f.q_en = paf.t_en;
    std::swap(f.qpos_aln, f.tpos_aln);
    f.cigar = cigar_swap_target_query(paf.cigar, paf.strand);
    f.long_cigar = cigar_swap_target_query(paf.long_cigar, paf.strand);
    if (!f.tpos_aln.empty()) f.aligned_pairs();
    return f;
}

// liftover.rs:134-167
std::vector<PafRecord> trim_paf_by_rgns(const std::vector<Region>& rgns,
                                        const std::vector<PafRecord>& paf_recs, bool invert_query,
                                        int policy, int threads) {
    std::vector<PafRecord> newvec;
    const std::vector<PafRecord>* recs = &paf_recs;
    if (invert_query) {
        for (const PafRecord& r : paf_recs) newvec.push_back(paf_swap_query_and_target(r));
        recs = &newvec;
    }
    std::vector<std::string> names;  // itertools `unique()`: first-appearance order
    for (const PafRecord& r : *recs)
        if (std::find(names.begin(), names.end(), r.t_name) == names.end()) names.push_back(r.t_name);
    std::vector<PafRecord> out;
    for (const std::string& name : names) {
        std::vector<PafRecord> tmp = trim_helper(name, *recs, rgns, policy, threads);
        for (PafRecord& r : tmp) out.push_back(std::move(r));
    }
    return out;
}

// liftover.rs:182-226
std::vector<PafRecord> break_paf_on_indels(const PafRecord& paf, uint32_t break_length, int policy) {
    std::vector<PafRecord> rtn;
    uint64_t cur_tpos = paf.t_st, pre_tpos = paf.t_st;
    for (const Cig& o : paf.cigar) {
        if (o.len > break_length && (o.op == OP_D || o.op == OP_I)) {
            if (cur_tpos > pre_tpos) {
                Region rgn;
                rgn.name = paf.t_name; rgn.st = pre_tpos; rgn.en = cur_tpos; rgn.id = paf.id;
                PafRecord x;
                if (trim_paf_rec_to_rgn(rgn, paf, policy, x)) {
                    if (!x.check_integrity()) throw Abort("break_paf_on_indels: check_integrity().unwrap()");
                    rtn.push_back(std::move(x));
                }
            }
            pre_tpos = cur_tpos;
            if (consumes_reference(o.op)) pre_tpos += o.len;
        }
        if (consumes_reference(o.op)) cur_tpos += o.len;
    }
    if (cur_tpos > pre_tpos) {
        Region rgn;
        rgn.name = paf.t_name; rgn.st = pre_tpos; rgn.en = cur_tpos; rgn.id = paf.id;
        PafRecord x;
        if (trim_paf_rec_to_rgn(rgn, paf, policy, x)) rtn.push_back(std::move(x));
    }
    return rtn;
}

// bamstats.rs:107-142 — u32 counters (wrapping), three f32 identities:
// `100.0 * equal as f32 / (sum) as f32` == (100.0f * (f32)equal) / (f32)(u32 sum)
void add_stats_from_cigar(const CigarString& c, Stats& s) {
    for (const Cig& o : c) {
        switch (o.op) {
            case OP_D: s.del_events += 1; s.del += o.len; break;
            case OP_I: s.ins_events += 1; s.ins += o.len; break;
            case OP_EQ: s.equal += o.len; break;
            case OP_X: s.diff += o.len; break;
            case OP_M: s.diff += o.len; s.matches += o.len; break;
            default: break;
        }
    }
    volatile float eq = (float)s.equal;  // volatile: forbid contraction / wider intermediates
    volatile float num = 100.0f * eq;
    volatile float d_all = (float)(uint32_t)(s.equal + s.diff + s.del + s.ins);
    volatile float d_ev = (float)(uint32_t)(s.equal + s.diff + s.del_events + s.ins_events);
    volatile float d_m = (float)(uint32_t)(s.equal + s.diff);
    s.id_by_all = num / d_all;
    s.id_by_events = num / d_ev;
    s.id_by_matches = num / d_m;
}

// bamstats.rs:91-105
Stats stats_from_paf(const PafRecord& p) {
    Stats s;
    add_stats_from_cigar(p.cigar, s);
    s.r_nm = p.t_name; s.r_len = (int64_t)p.t_len; s.r_st = (int64_t)p.t_st; s.r_en = (int64_t)p.t_en;
    s.q_nm = p.q_name; s.q_len = (int64_t)p.q_len; s.q_st = (int64_t)p.q_st; s.q_en = (int64_t)p.q_en;
    s.strand = p.strand;
    return s;
}

// Rust `impl Display for f32`: shortest decimal digit string that parses back to the same
// f32, printed positionally (never an exponent), no trailing ".0"; "NaN", "inf", "-inf".
// Restated by search: the correctly rounded p-significant-digit decimal for p = 1..9, the
// first that round-trips.  (tests cross-check this against libstdc++'s Ryu-based to_chars.)
std::string fmt_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    if (v == 0.0f) return std::signbit(v) ? "-0" : "0";
    char buf[64];
    std::string digits;
    int e10 = 0;
    bool neg = v < 0;
    const float av = std::fabs(v);
    for (int p = 1; p <= 9; p++) {
        snprintf(buf, sizeof buf, "%.*e", p - 1, (double)av);  // the p-digit decimal closest to v
        bool ok = strtof(buf, nullptr) == av || p == 9;
        if (!ok && strtod(buf, nullptr) < (double)av) {
            // At an exact power of two the rounding interval reaches twice as far up as down: the closest p-digit decimal can
            // fall short of it from below while its upper neighbour is inside.  The shortest-digits algorithms (Grisu / Dragon4
            // in Rust's flt2dec, with `minus = 1, plus = 2` for such mantissas) print that neighbour.  Found by the exhaustive
            // cross-check against std::to_chars (tests/native/f32_fmt_check.cpp): 2^-96 is the one f32 in [0, 100] it affects.
            unsigned long long m = 0;
            const char* e = strchr(buf, 'e');
            for (const char* c = buf; c < e; c++)
                if (is_digit(*c)) m = m * 10 + (unsigned long long)(*c - '0');
            char up[64];
            snprintf(up, sizeof up, "%llue%d", m + 1, atoi(e + 1) - (p - 1));
            if (strtof(up, nullptr) == av) {
                snprintf(buf, sizeof buf, "%.*e", p - 1, strtod(up, nullptr));  // same digits, canonical "d.ddde+xx" spelling
                ok = true;
            }
        }
        if (ok && p < 9) {
            // printf rounds an exact tie to even; Rust's shortest formatter (core::num::flt2dec::strategy::dragon::format_shortest,
            // which its Grisu fast path defers to whenever the two neighbours are equally close) rounds it UP:
            // `if up && (!down || *mant.mul_pow2(1) >= scale)`.  A tie = the exact decimal expansion of v has p + 1 significant
            // digits and ends in 5.  (Stated from knowledge of the std sources — std is not present in this image to confirm;
            // SURVEY lists the f32 digits as "parity unpinned".)
            char exact[256];
            snprintf(exact, sizeof exact, "%.150e", (double)av);  // glibc prints the exact binary value
            std::string dg;
            for (const char* c = exact; *c && *c != 'e'; c++)
                if (is_digit(*c)) dg.push_back(*c);
            bool tie = dg.size() > (size_t)p && dg[(size_t)p] == '5';
            for (size_t k = (size_t)p + 1; tie && k < dg.size(); k++) tie = dg[k] == '0';
            if (tie) {
                unsigned long long m = 0;
                for (int k = 0; k < p; k++) m = m * 10 + (unsigned long long)(dg[(size_t)k] - '0');  // truncated p digits
                const char* e = strchr(exact, 'e');
                char up[64];
                snprintf(up, sizeof up, "%llue%d", m + 1, atoi(e + 1) - (p - 1));
                if (strtof(up, nullptr) == av) snprintf(buf, sizeof buf, "%.*e", p - 1, strtod(up, nullptr));
            }
        }
        if (ok) {
            const char* e = strchr(buf, 'e');
            e10 = atoi(e + 1);
            for (const char* c = buf; c < e; c++)
                if (is_digit(*c)) digits.push_back(*c);
            break;
        }
    }
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    std::string s = neg ? "-" : "";
    int nd = (int)digits.size();
    if (e10 >= 0) {
        if (nd <= e10 + 1) {
            s += digits;
            s.append((size_t)(e10 + 1 - nd), '0');
        } else {
            s.append(digits, 0, (size_t)e10 + 1);
            s.push_back('.');
            s.append(digits, (size_t)e10 + 1, std::string::npos);
        }
    } else {
        s += "0.";
        s.append((size_t)(-e10 - 1), '0');
        s += digits;
    }
    return s;
}

// bamstats.rs:225-236
std::string stats_header(bool qbed) {
    std::string s;
    if (qbed) {
        s += "#query_name\tquery_start\tquery_end\tquery_length\t";
        s += "strand\t";
        s += "reference_name\treference_start\treference_end\treference_length\t";
    } else {
        s += "#reference_name\treference_start\treference_end\treference_length\t";
        s += "strand\t";
        s += "query_name\tquery_start\tquery_end\tquery_length\t";
    }
    s += "perID_by_matches\tperID_by_events\tperID_by_all\tmatches\tmismatches\tdeletion_events\t"
         "insertion_events\tdeletions\tinsertions\n";
    return s;
}

// bamstats.rs:239-270
std::string stats_row(const Stats& st, bool qbed) {
    std::string s;
    auto q = [&] {
        s += st.q_nm; s += '\t'; s += std::to_string(st.q_st); s += '\t';
        s += std::to_string(st.q_en); s += '\t'; s += std::to_string(st.q_len); s += '\t';
    };
    auto r = [&] {
        s += st.r_nm; s += '\t'; s += std::to_string(st.r_st); s += '\t';
        s += std::to_string(st.r_en); s += '\t'; s += std::to_string(st.r_len); s += '\t';
    };
    if (qbed) { q(); s += st.strand; s += '\t'; r(); }
    else { r(); s += st.strand; s += '\t'; q(); }
    s += fmt_f32(st.id_by_matches); s += '\t';
    s += fmt_f32(st.id_by_events); s += '\t';
    s += fmt_f32(st.id_by_all); s += '\t';
    s += std::to_string(st.equal); s += '\t';
    s += std::to_string(st.diff); s += '\t';
    s += std::to_string(st.del_events); s += '\t';
    s += std::to_string(st.ins_events); s += '\t';
    s += std::to_string(st.del); s += '\t';
    s += std::to_string(st.ins); s += '\n';
    return s;
}

// main.rs:186-214
std::string run_liftover(const char* paf, size_t paf_n, const char* bed, size_t bed_n, bool qbed,
                         bool largest, int policy, int threads) {
    std::vector<Region> rgns = parse_bed_text(bed, bed_n);
    Paf p = Paf::from_text(paf, paf_n);
    std::vector<PafRecord> recs = trim_paf_by_rgns(rgns, p.records, qbed, policy, threads);
    std::string out;
    if (largest) {
        // sorted_by_key(id) is a stable sort; group_by(id); max_by_key keeps the LAST maximum
        std::stable_sort(recs.begin(), recs.end(),
                         [](const PafRecord& a, const PafRecord& b) { return a.id < b.id; });
        size_t i = 0;
        while (i < recs.size()) {
            size_t j = i, best = i;
            while (j < recs.size() && recs[j].id == recs[i].id) {
                if (recs[j].t_en - recs[j].t_st >= recs[best].t_en - recs[best].t_st) best = j;
                j++;
            }
            out += recs[best].to_line();
            out += '\n';
            i = j;
        }
    } else {
        for (const PafRecord& r : recs) {
            out += r.to_line();
            out += '\n';
        }
    }
    return out;
}

// main.rs:50-58
// main.rs:271-281 — `rb break-paf --max-size N`: records in FILE order, aligned_pairs (strip) then break_paf_on_indels
std::string run_break_paf(const char* paf, size_t paf_n, uint32_t max_size, int policy) {
    Paf p = Paf::from_text(paf, paf_n);
    std::string out;
    for (PafRecord& r : p.records) {
        r.aligned_pairs();
        for (const PafRecord& x : break_paf_on_indels(r, max_size, policy)) { out += x.to_line(); out += '\n'; }
    }
    return out;
}

// main.rs:176-182 — `rb invert`: every record as read (integrity-checked, not stripped), query and target swapped
std::string run_invert(const char* paf, size_t paf_n) {
    Paf p = Paf::from_text(paf, paf_n);
    std::string out;
    for (const PafRecord& r : p.records) { out += paf_swap_query_and_target(r).to_line(); out += '\n'; }
    return out;
}

// main.rs:218-230 — `rb trim-paf`: every record of the (sorted, trimmed) set, in the set's order
std::string run_trim_paf(const char* paf, size_t paf_n, int match_score, int diff_score, int indel_score,
                         bool remove_contained, int policy) {
    Paf p = Paf::from_text(paf, paf_n);
    p.overlapping_paf_recs(match_score, diff_score, indel_score, remove_contained, policy);
    std::string out;
    for (const PafRecord& r : p.records) { out += r.to_line(); out += '\n'; }
    return out;
}

std::string run_stats(const char* paf, size_t paf_n, bool qbed) {
    std::string out = stats_header(qbed);
    Paf p = Paf::from_text(paf, paf_n);
    for (const PafRecord& r : p.records) out += stats_row(stats_from_paf(r), qbed);
    return out;
}

}  // namespace orc
