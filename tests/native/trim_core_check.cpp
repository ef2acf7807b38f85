// CPU fuzz harness: the product's closed-form `rb trim-paf` (rustybam_b200/csrc/trim_core.cuh — the exact code
// k_trim_pairs / k_trim_rows compile for the GPU — plus the host round structure of trim_rounds.hpp) against the
// literal per-base oracle (oracle/rb_oracle.cpp: Paf::overlapping_paf_recs, trim_overlapping_pafs,
// truncate_record_by_query).  Random groups of records that share a query name and overlap on it, both strands,
// random scores, CIGARs with indels next to every kind of op, N/P/S in the middle, zero-length ops, adjacent
// same-class ops, stripped leading/trailing insertions.
//   usage: trim_core_check <seed> <n_groups>
#include <climits>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "trim_core.cuh"
#include "trim_rounds.hpp"
#include "rb_oracle.hpp"
#include "rec_core.cuh"

using namespace rb;

static const char* OPC = "MIDNSHP=X";

struct TestRec {
    std::vector<uint32_t> ops;
    uint64_t t_st, t_en, q_st, q_en, q_len, t_len;
    char strand;
    std::string q_name, t_name, line;
    uint64_t mapq = 60;
};

static std::string cig_text(const std::vector<uint32_t>& ops, size_t a, size_t b) {
    std::string s;
    for (size_t k = a; k < b; k++) s += std::to_string(op_len(ops[k])) + OPC[op_code(ops[k])];
    return s;
}


struct Counts { long n_fail = 0, n_abort = 0, n_trimmed = 0, n_rows = 0, n_rounds_max = 0, n_unsup = 0; };

// one PAF text through the oracle and through the product's closed form; `recs` come from the oracle's own line parser
static void check_text(const std::string& text, TrimScores sc, bool remove_contained, std::mt19937_64& rng, Counts& C, int g, int policy = POLICY_RIGHTMOST) {
    auto U = [&](uint64_t lo, uint64_t hi) { return lo + rng() % (hi - lo + 1); };
    long &n_fail = C.n_fail, &n_abort = C.n_abort, &n_trimmed = C.n_trimmed, &n_rows = C.n_rows, &n_rounds_max = C.n_rounds_max, &n_unsup = C.n_unsup;
    std::vector<TestRec> recs;
    try {
        orc::Paf parsed = orc::Paf::from_text(text.data(), text.size());
        for (const orc::PafRecord& p : parsed.records) {
            TestRec tr;
            for (const orc::Cig& c : p.cigar) tr.ops.push_back((c.len << 4) | c.op);
            tr.t_st = p.t_st; tr.t_en = p.t_en; tr.q_st = p.q_st; tr.q_en = p.q_en; tr.q_len = p.q_len; tr.t_len = p.t_len;
            tr.strand = p.strand; tr.q_name = p.q_name; tr.t_name = p.t_name; tr.mapq = p.mapq;
            recs.push_back(std::move(tr));
        }
    } catch (const orc::Abort&) { n_abort++; return; }
    do {
        // ---- oracle ----
        std::string want;
        bool aborted = false;
        try {
            want = orc::run_trim_paf(text.data(), text.size(), sc.match, sc.diff, sc.indel, remove_contained, policy == POLICY_EARLY_EXIT ? orc::POLICY_EARLY_EXIT : orc::POLICY_RIGHTMOST);
        } catch (const orc::Abort& e) {
            aborted = true;
            if (getenv("RB_DBG")) fprintf(stderr, "abort: %s\n", e.what());
        }

        // ---- product: name-sorted set, device-layout arrays as the kernels define them ----
        std::vector<size_t> order(recs.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return recs[a].q_name < recs[b].q_name; });
        std::vector<uint32_t> ops;
        std::vector<uint64_t> op_off;
        ops.resize(U(0, 40), (7u << 4) | OP_EQ);
        op_off.push_back(ops.size());
        for (size_t i : order) { ops.insert(ops.end(), recs[i].ops.begin(), recs[i].ops.end()); op_off.push_back(ops.size()); }
        std::vector<uint8_t> head(ops.size() + 1, 0);
        for (size_t r = 0; r < order.size(); r++) head[op_off[r]] = 1;
        std::vector<Ctr> samples((ops.size() / SAMPLE + 2) * SUBS);
        {
            Ctr run = ctr_zero(), rel = ctr_zero();
            bool head_in_chunk = false;
            for (size_t k = 0; k < ops.size(); k++) {
                if (k % SAMPLE == 0) { rel = ctr_zero(); head_in_chunk = false; }
                if (head[k]) { run = ctr_zero(); rel = ctr_zero(); if (k % SAMPLE) head_in_chunk = true; }
                if (k % SAMPLE == 0) samples[(k / SAMPLE) * SUBS] = run;
                else if (k % SUB_OPS == 0) {
                    Ctr e = rel;
                    e.aux = head_in_chunk ? SUB_ABS : 0u;
                    samples[(k / SAMPLE) * SUBS + (k % SAMPLE) / SUB_OPS] = e;
                }
                ctr_add_op(run, ops[k]);
                ctr_add_op(rel, ops[k]);
            }
        }
        OpsView view;
        view.ops = ops.data(); view.samples = samples.data();
        uint32_t acc_mem[16];
        ClassAcc acc;
        acc.sum = acc_mem; acc.stride = 1;

        std::vector<RecInfo> ri(order.size());
        std::vector<TrimView> tv(order.size());
        std::vector<uint32_t> qp(ops.size() + 1, 0);
        std::vector<long long> wp(ops.size() + 1, 0);
        std::vector<uint32_t> ap(ops.size() + 1, 0);
        TrimArr arr{qp.data(), wp.data(), ap.data(), policy};
        // what k_trim_scan / k_trim_rescan compute for one record and its current view
        auto scan_record = [&](size_t r) {
            const RecInfo& R = ri[r];
            uint32_t q = 0, acol = 0;
            long long w = 0;
            for (uint64_t k = R.op_first; k < R.op_end; k++) {
                qp[k] = q; ap[k] = acol;
                if (is_qry(op_code(ops[k]))) q += op_len(ops[k]);
                acol += op_len(ops[k]);
                if (k + 1 == R.eo1) tv[r].x_end = q;
            }
            for (uint64_t k = R.op_first; k < R.op_end; k++) {
                wp[k] = w;
                if (k >= R.eo0 && k < R.eo1) w += (policy == POLICY_EARLY_EXIT) ? trim_w_op_view(view, arr, tv[r], k, sc) : trim_w_op(view, k, R.eo1, sc);
                if (k + 1 == R.eo1) tv[r].w_tot = w;
            }
        };
        std::vector<TrimSpan> spans(order.size());
        bool strip_abort = false, unsupported = false;
        for (size_t r = 0; r < order.size(); r++) {
            const TestRec& tr = recs[order[r]];
            RecInfo& R = ri[r];
            R = RecInfo{};
            R.op_first = op_off[r]; R.op_end = op_off[r + 1];
            R.t_st = tr.t_st; R.t_en = tr.t_en; R.q_st0 = tr.q_st; R.q_en0 = tr.q_en;
            R.q_len = tr.q_len; R.t_len = tr.t_len; R.mapq = tr.mapq;
            R.flags = tr.strand == '-' ? RF_MINUS : 0;
            if (strip_record(ops.data(), R) != RE_OK) { strip_abort = true; break; }
            for (uint64_t k = R.op_first; k < R.op_end; k++) {
                if (op_len(ops[k]) == 0) R.flags |= RF_SLOW;
                if (k > R.op_first && op_code(ops[k]) == op_code(ops[k - 1])) R.flags |= RF_SLOW;
            }
            R.tot = ctr_before(view, R, R.eo1 - 1, acc);
            ctr_add_op(R.tot, ops[R.eo1 - 1]);
            Ctr lead = ctr_before(view, R, R.eo0, acc);
            ctr_sub(R.tot, lead);
            trim_view_init(view, R, tv[r]);
            scan_record(r);
            if (tv[r].bad) unsupported = true;
            spans[r] = TrimSpan{R.q_st, R.q_en, 0};
            spans[r].name = (r > 0 && recs[order[r - 1]].q_name == tr.q_name) ? spans[r - 1].name : (r > 0 ? spans[r - 1].name + 1 : 0);
        }
        if (strip_abort) {
            if (!aborted) { n_fail++; fprintf(stderr, "FAIL strip abort but oracle ran\n%s", text.c_str()); }
            else n_abort++;
            break;
        }
        if (unsupported) { n_unsup++; break; }
        std::vector<uint8_t> contained;
        std::vector<TrimPairSel> sel;
        bool p_abort = false;
        long rounds = 0;
        for (;;) {
            const size_t waiting = trim_round(spans, contained, sel);
            rounds++;
            {   // the per-group selection the GPU runs (trim_select_group) picks the same pairs, flags and waiting count
                std::vector<uint8_t> c2(spans.size(), 7);
                std::vector<TrimPairSelDev> s2;
                size_t w2 = 0;
                for (uint32_t lo = 0; lo < spans.size();) {
                    uint32_t hi = lo + 1;
                    while (hi < spans.size() && spans[hi].name == spans[lo].name) hi++;
                    TrimPairSelDev one;
                    const uint32_t np = trim_select_group(tv.data(), lo, hi, c2.data(), one);
                    if (np) { s2.push_back(one); w2 += np - 1; }
                    lo = hi;
                }
                bool same = w2 == waiting && s2.size() == sel.size() && c2 == contained;
                // trim_round lists its pairs by overlap; per name there is one, so compare as sets keyed by `left`
                for (size_t x = 0; same && x < sel.size(); x++) {
                    bool found = false;
                    for (const TrimPairSelDev& y : s2)
                        if (y.left == sel[x].left && y.right == sel[x].right && y.st_ovl == sel[x].st_ovl && y.en_ovl == sel[x].en_ovl) found = true;
                    same = found;
                }
                if (!same) { n_fail++; if (n_fail < 10) fprintf(stderr, "FAIL selection differs from trim_round (waiting %zu vs %zu, %zu vs %zu pairs)\n", w2, waiting, s2.size(), sel.size()); }
            }
            for (const TrimPairSel& p : sel) {
                const RecInfo &rl = ri[p.left], &rr = ri[p.right];
                TrimView &tl = tv[p.left], &tr = tv[p.right];
                const uint64_t A = p.st_ovl, B = p.en_ovl;
                TrimBest best{LLONG_MIN, 0};
                trim_fixed_candidates(view, arr, rl, tl, rr, tr, A, B, sc, best);
                const TrimSide sl = trim_side(view, arr, rl, tl, A, sc), sr = trim_side(view, arr, rr, tr, A, sc);
                const uint32_t nt = (uint32_t)U(1, 5);  // any split of the work over threads gives the same arg-max
                for (uint32_t t = 0; t < nt; t++) {
                    trim_scan_candidates(view, arr, true, rl, tl, sl, rr, tr, sr, A, B, sc, t, nt, best);
                    trim_scan_candidates(view, arr, false, rl, tl, sl, rr, tr, sr, A, B, sc, t, nt, best);
                }
                {   // the key's field limits: totals of +-(2^31 - 1) survive, order by (total, then the smaller c)
                    const TrimBest ext[4] = {{2147483647ll, A}, {-2147483647ll, A + 5}, {0, A + 4294967295ull}, {-1, A}};
                    for (const TrimBest& x : ext) {
                        const TrimBest y = trim_unkey(trim_key(x, A), A);
                        if (y.total != x.total || y.c != x.c) { n_fail++; fprintf(stderr, "FAIL key extremes\n"); }
                    }
                    if (!(trim_key(ext[0], A) > trim_key(ext[2], A) && trim_key(ext[2], A) > trim_key(ext[3], A) &&
                          trim_key(TrimBest{7, A + 1}, A) > trim_key(TrimBest{7, A + 2}, A))) { n_fail++; fprintf(stderr, "FAIL key order\n"); }
                }
                {   // the packed 64-bit key the kernels reduce with atomicMax decodes to the same arg-max
                    const unsigned long long key = trim_key(best, A);
                    TrimBest back = trim_unkey(key, A);
                    if (back.total != best.total || back.c != best.c) { n_fail++; fprintf(stderr, "FAIL key round trip\n"); }
                }
                const long long r_tot = trim_S(view, arr, rr, tr, A, B, sc);
                const uint64_t s = trim_split(best, r_tot, A);
                if (trim_truncate(view, arr, rl, tl, tl.q_st, s) != TRIM_OK) { p_abort = true; break; }
                if (trim_truncate(view, arr, rr, tr, s, tr.q_en) != TRIM_OK) { p_abort = true; break; }
                if (policy == POLICY_EARLY_EXIT) { scan_record(p.left); scan_record(p.right); }  // the score prefix follows the view
                spans[p.left].q_st = tl.q_st; spans[p.left].q_en = tl.q_en;
                spans[p.right].q_st = tr.q_st; spans[p.right].q_en = tr.q_en;
                n_trimmed++;
            }
            if (p_abort || waiting == 0) break;
            if (rounds > 1000) { fprintf(stderr, "FAIL no convergence\n"); n_fail++; break; }
        }
        n_rounds_max = std::max(n_rounds_max, rounds);
        if (p_abort != aborted) {
            n_fail++;
            if (n_fail < 10) fprintf(stderr, "FAIL abort mismatch product=%d oracle=%d scores %d %d %d\n%s", (int)p_abort, (int)aborted, sc.match, sc.diff, sc.indel, text.c_str());
            break;
        }
        if (aborted) { n_abort++; break; }
        std::string got;
        for (size_t r = 0; r < order.size(); r++) {
            if (remove_contained && contained[r]) continue;
            const TestRec& tr = recs[order[r]];
            const RecInfo& R = ri[r];
            PairRes pr;
            trim_row(view, R, tv[r], acc, pr);
            std::string cg, id;
            if (R.flags & RF_STRIPPED) {
                id = "_TO." + cig_text(ops, R.op_first, R.eo0) + ".";
                for (uint64_t k = R.op_end; k > R.eo1; k--) id += cig_text(ops, k - 1, k);
            }
            if (pr.kind == PK_EARLY) cg = cig_text(ops, pr.si, pr.ei + 1);
            else if (pr.kind == PK_TRIM) {
                if (R.flags & RF_SLOW) merged_walk(view, pr.si, pr.ei, pr.s_len, pr.e_len, [&](uint32_t len, uint32_t c) { cg += std::to_string(len) + OPC[c]; });
                else if (pr.si == pr.ei) cg = std::to_string(pr.s_len) + OPC[op_code(ops[pr.si])];
                else cg = std::to_string(pr.s_len) + OPC[op_code(ops[pr.si])] + cig_text(ops, pr.si + 1, pr.ei) + std::to_string(pr.e_len) + OPC[op_code(ops[pr.ei])];
            } else { n_fail++; fprintf(stderr, "FAIL dropped row\n"); }
            if (cg.size() != pr.cg_bytes) { n_fail++; fprintf(stderr, "FAIL cg_bytes %u vs %zu\n", pr.cg_bytes, cg.size()); }
            got += tr.q_name + "\t" + std::to_string(tr.q_len) + "\t" + std::to_string(pr.q_st) + "\t" + std::to_string(pr.q_en) + "\t" +
                   tr.strand + "\t" + tr.t_name + "\t" + std::to_string(tr.t_len) + "\t" + std::to_string(pr.t_st) + "\t" + std::to_string(pr.t_en) + "\t" +
                   std::to_string(pr.nmatch) + "\t" + std::to_string(pr.aln_len) + "\t" + std::to_string(tr.mapq) + "\tid:Z:" + id + "\tcg:Z:" + cg + "\n";
            n_rows++;
        }
        if (got != want) {
            n_fail++;
            if (n_fail < 10)
                fprintf(stderr, "FAIL group %d scores %d %d %d remove_contained %d\n in:\n%s got:\n%s want:\n%s\n", g, sc.match, sc.diff, sc.indel,
                        (int)remove_contained, text.c_str(), got.c_str(), want.c_str());
        }
    
    } while (0);
}
int main(int argc, char** argv) {
    Counts C;
    if (argc > 1 && strcmp(argv[1], "--paf") == 0) {  // trim_core_check --paf FILE [match diff indel remove_contained]
        FILE* f = fopen(argv[2], "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
        std::string text;
        char buf[1 << 16];
        size_t k;
        while ((k = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, k);
        fclose(f);
        TrimScores sc{argc > 3 ? atoi(argv[3]) : 1, argc > 4 ? atoi(argv[4]) : 1, argc > 5 ? atoi(argv[5]) : 1};
        std::mt19937_64 rng(7);
        check_text(text, sc, argc > 6 && atoi(argv[6]) != 0, rng, C, 0, argc > 7 ? atoi(argv[7]) : POLICY_RIGHTMOST);
        printf("aborts=%ld unsupported=%ld trimmed_pairs=%ld rows=%ld max_rounds=%ld FAIL=%ld\n", C.n_abort, C.n_unsup, C.n_trimmed, C.n_rows, C.n_rounds_max, C.n_fail);
        return C.n_fail ? 1 : 0;
    }

    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
    const int n_groups = argc > 2 ? atoi(argv[2]) : 500;
    std::mt19937_64 rng(seed);
    auto U = [&](uint64_t lo, uint64_t hi) { return lo + rng() % (hi - lo + 1); };

    for (int g = 0; g < n_groups; g++) {
        // ---- one PAF: a few query names, a few records each ----
        std::vector<TestRec> recs;
        const int n_names = (int)U(1, 3);
        const int style = (int)U(0, 9);
        for (int nm = 0; nm < n_names; nm++) {
            const int n_rec = (int)U(1, 5);
            for (int r = 0; r < n_rec; r++) {
                TestRec tr;
                const int n_ops = (int)U(1, style == 0 ? 60 : 16);
                std::vector<uint32_t> body;
                auto push = [&](uint32_t code, uint32_t len) { body.push_back((len << 4) | code); };
                if (U(0, 5) == 0) push(OP_I, (uint32_t)U(1, 4));  // stripped at load
                static const uint32_t ends[] = {OP_EQ, OP_EQ, OP_X, OP_M};
                static const uint32_t codes_eqx[] = {OP_EQ, OP_EQ, OP_EQ, OP_X, OP_I, OP_D};
                static const uint32_t codes_all[] = {OP_EQ, OP_X, OP_M, OP_I, OP_D, OP_N, OP_P, OP_EQ, OP_I, OP_D, OP_X};
                push(ends[U(0, 3)], (uint32_t)U(1, 6));
                uint32_t prev = op_code(body.back());
                for (int j = 0; j < n_ops; j++) {
                    uint32_t code = style < 6 ? codes_eqx[U(0, 5)] : codes_all[U(0, 10)];
                    if (style < 8 && code == prev) code = (code == OP_EQ) ? OP_X : OP_EQ;
                    uint32_t len = (uint32_t)(U(0, 3) == 0 ? U(1, 30) : U(1, 4));
                    if (style == 9 && U(0, 6) == 0) len = 0;
                    push(code, len);
                    prev = code;
                }
                push(ends[U(0, 3)], (uint32_t)U(1, 6));
                if (U(0, 5) == 0) push(U(0, 1) ? OP_D : OP_I, (uint32_t)U(1, 4));
                tr.ops = body;
                uint64_t T = 0, Q = 0;
                for (uint32_t w : body) {
                    if (is_ref(op_code(w))) T += op_len(w);
                    if (is_qry(op_code(w))) Q += op_len(w);
                }
                tr.t_st = U(1, 500); tr.t_en = tr.t_st + T; tr.t_len = tr.t_en + U(0, 20);
                tr.q_st = U(0, 80); tr.q_en = tr.q_st + Q; tr.q_len = 400;
                tr.strand = U(0, 1) ? '+' : '-';
                tr.q_name = "q" + std::to_string(nm); tr.t_name = "chrT";
                tr.line = tr.q_name + "\t" + std::to_string(tr.q_len) + "\t" + std::to_string(tr.q_st) + "\t" +
                          std::to_string(tr.q_en) + "\t" + tr.strand + "\t" + tr.t_name + "\t" + std::to_string(tr.t_len) + "\t" +
                          std::to_string(tr.t_st) + "\t" + std::to_string(tr.t_en) + "\t0\t0\t60\tcg:Z:" + cig_text(body, 0, body.size());
                recs.push_back(std::move(tr));
            }
        }
        std::shuffle(recs.begin(), recs.end(), rng);  // file order mixes the names: the stable sort has work to do
        TrimScores sc{(int32_t)U(1, 3), (int32_t)U(0, 3), (int32_t)U(0, 3)};
        if (U(0, 2) == 0) sc = TrimScores{1, 1, 1};
        const bool remove_contained = U(0, 1) == 1;

        std::string text;
        for (auto& tr : recs) text += tr.line + "\n";
        check_text(text, sc, remove_contained, rng, C, g, (argc > 3) ? atoi(argv[3]) : (int)(g & 1));
    }
    printf("groups=%d aborts=%ld unsupported=%ld trimmed_pairs=%ld rows=%ld max_rounds=%ld FAIL=%ld\n", n_groups, C.n_abort, C.n_unsup, C.n_trimmed, C.n_rows,
           C.n_rounds_max, C.n_fail);
    return C.n_fail ? 1 : 0;
}
