// CPU fuzz harness: the product's closed-form pair resolution (rustybam_b200/csrc/lift_core.cuh,
// rec_core.cuh — the exact code k_lift / k_rec_prep compile for the GPU) against the literal
// per-base oracle (oracle/rb_oracle.cpp), both binary_search policies, random CIGARs that include
// zero-length ops, adjacent same-class ops, N/P, leading/trailing indels and clips.
//   usage: lift_core_check <seed> <n_records>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "lift_core.cuh"
#include "rb_oracle.hpp"
#include "rec_core.cuh"
#include "stream_core.cuh"

using namespace rb;

static const char* OPC = "MIDNSHP=X";

struct TestRec {
    std::vector<uint32_t> ops;  // packed
    uint64_t t_st, t_en, q_st, q_en, q_len, t_len;
    char strand;
    std::string line;
};

static std::string cig_text(const std::vector<uint32_t>& ops, size_t a, size_t b) {
    std::string s;
    for (size_t k = a; k < b; k++) s += std::to_string(op_len(ops[k])) + OPC[op_code(ops[k])];
    return s;
}

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
    const int n_rec = argc > 2 ? atoi(argv[2]) : 2000;
    std::mt19937_64 rng(seed);
    auto U = [&](uint64_t lo, uint64_t hi) { return lo + rng() % (hi - lo + 1); };

    std::vector<TestRec> recs;
    for (int r = 0; r < n_rec; r++) {
        TestRec tr;
        const int style = (int)U(0, 9);
        int n_ops = (int)U(1, style == 0 ? 120 : 24);
        if (U(0, 11) == 0) n_ops *= (int)U(8, 30);  // long records (10 .. 100+ sample chunks): chunk_of's interpolation step and bisection
        std::vector<uint32_t> body;
        auto push = [&](uint32_t code, uint32_t len) { body.push_back((len << 4) | code); };
        if (U(0, 9) == 0) push(OP_H, (uint32_t)U(1, 5));
        if (U(0, 7) == 0) push(OP_S, (uint32_t)U(1, 5));
        if (U(0, 3) == 0) {  // leading indel run
            int k = (int)U(1, 3);
            for (int j = 0; j < k; j++) push(U(0, 5) == 0 ? OP_D : OP_I, (uint32_t)U(1, 4));
        }
        static const uint32_t codes_eqx[] = {OP_EQ, OP_EQ, OP_EQ, OP_X, OP_I, OP_D};
        static const uint32_t codes_all[] = {OP_EQ, OP_X, OP_M, OP_I, OP_D, OP_N, OP_P, OP_EQ, OP_I, OP_D};
        uint32_t prev = 99;
        for (int j = 0; j < n_ops; j++) {
            uint32_t code = style < 6 ? codes_eqx[U(0, 5)] : codes_all[U(0, 9)];
            if (style < 8 && code == prev) code = (code == OP_EQ) ? OP_X : OP_EQ;  // canonical in most styles
            uint32_t len = (uint32_t)(U(0, 3) == 0 ? U(1, 40) : U(1, 4));
            if (style == 9 && U(0, 6) == 0) len = 0;
            push(code, len);
            prev = code;
        }
        if (U(0, 3) == 0) {
            int k = (int)U(1, 3);
            for (int j = 0; j < k; j++) push(U(0, 1) ? OP_D : OP_I, (uint32_t)U(1, 4));
        }
        if (U(0, 7) == 0) push(OP_S, (uint32_t)U(1, 5));
        if (U(0, 9) == 0) push(OP_H, (uint32_t)U(1, 5));
        tr.ops = body;
        uint64_t T = 0, Q = 0;
        for (uint32_t w : body) {
            if (is_ref(op_code(w))) T += op_len(w);
            if (is_qry(op_code(w))) Q += op_len(w);
        }
        tr.t_st = U(1, 50); tr.t_en = tr.t_st + T; tr.t_len = tr.t_en + U(0, 20);
        tr.q_st = U(0, 50); tr.q_en = tr.q_st + Q; tr.q_len = tr.q_en + U(0, 20);
        tr.strand = U(0, 1) ? '+' : '-';
        tr.line = "q" + std::to_string(r) + "\t" + std::to_string(tr.q_len) + "\t" + std::to_string(tr.q_st) + "\t" +
                  std::to_string(tr.q_en) + "\t" + tr.strand + "\tchrT\t" + std::to_string(tr.t_len) + "\t" +
                  std::to_string(tr.t_st) + "\t" + std::to_string(tr.t_en) + "\t0\t0\t60\tcg:Z:" +
                  cig_text(body, 0, body.size());
        recs.push_back(std::move(tr));
    }

    // ---- device-layout arrays built on the host exactly as the kernels define them ----
    std::vector<uint32_t> ops;
    std::vector<uint64_t> op_off(1, 0);
    ops.resize(U(0, 40), (7u << 4) | OP_EQ);  // junk ops of a fake preceding record shift the chunk phase
    op_off[0] = ops.size();
    for (auto& tr : recs) {
        ops.insert(ops.end(), tr.ops.begin(), tr.ops.end());
        op_off.push_back(ops.size());
    }
    std::vector<uint8_t> head(ops.size() + 1, 0);
    for (size_t r = 0; r < recs.size(); r++) head[op_off[r]] = 1;
    std::vector<Ctr> samples((ops.size() / SAMPLE + 2) * SUBS);
    {
        Ctr run = ctr_zero(), rel = ctr_zero();  // run: since the record's first op; rel: since the chunk start / last head in the chunk
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
    // k_samples' wide-window form, on a random subset of the chunks (some cases: all of them): the absolute sample only, marked
    // SUB_ABS, entries 1..3 never written (poisoned here) — every lookup has to come out the same
    {
        const uint64_t mode = U(0, 3);  // 0, 1: every chunk keeps its sub-samples; 2: a random half does; 3: none does
        for (size_t c = 0; mode >= 2 && c * SAMPLE < ops.size(); c++) {
            if (mode == 2 && U(0, 1)) continue;
            samples[c * SUBS].aux |= SUB_ABS;
            for (uint32_t t = 1; t < SUBS; t++) memset(&samples[c * SUBS + t], 0xA5, sizeof(Ctr));
        }
    }
    OpsView view;
    view.ops = ops.data(); view.samples = samples.data();
    uint32_t acc_mem[16];
    ClassAcc acc;
    acc.sum = acc_mem; acc.stride = 1;

    long n_stream = 0;
    long n_pairs = 0, n_trim = 0, n_early = 0, n_drop = 0, n_abort = 0, n_slow = 0, n_fail = 0;
    for (size_t r = 0; r < recs.size(); r++) {
        const TestRec& tr = recs[r];
        // oracle side
        orc::PafRecord orec;
        bool aborted = false;
        try {
            orec = orc::PafRecord::parse(tr.line);
            std::string why;
            if (!orec.check_integrity(&why)) { fprintf(stderr, "generator bug: %s\n", why.c_str()); return 2; }
            orec.aligned_pairs();
        } catch (const orc::Abort& e) { aborted = true; if (getenv("RB_DBG")) fprintf(stderr, "abort: %s :: %s\n", e.what(), tr.line.c_str()); }

        RecInfo ri{};
        ri.op_first = op_off[r]; ri.op_end = op_off[r + 1];
        ri.t_st = tr.t_st; ri.t_en = tr.t_en; ri.q_st0 = tr.q_st; ri.q_en0 = tr.q_en;
        ri.q_len = tr.q_len; ri.t_len = tr.t_len; ri.mapq = 60;
        ri.flags = tr.strand == '-' ? RF_MINUS : 0;
        uint32_t err = strip_record(ops.data(), ri);
        if ((err != RE_OK) != aborted) {
            fprintf(stderr, "FAIL strip/abort mismatch rec %zu err=%u aborted=%d\n%s\n", r, err, (int)aborted, tr.line.c_str());
            n_fail++;
            continue;
        }
        if (aborted) { n_abort++; continue; }
        for (uint64_t k = ri.op_first; k < ri.op_end; k++) {
            if (op_len(ops[k]) == 0) ri.flags |= RF_SLOW;
            if (k > ri.op_first && op_code(ops[k]) == op_code(ops[k - 1])) ri.flags |= RF_SLOW;
        }
        ri.tot = ctr_before(view, ri, ri.eo1 - 1, acc);
        ctr_add_op(ri.tot, ops[ri.eo1 - 1]);
        Ctr lead = ctr_before(view, ri, ri.eo0, acc);
        ctr_sub(ri.tot, lead);
        if (ri.flags & RF_SLOW) n_slow++;
        std::string rec_id;
        if (ri.flags & RF_STRIPPED) {
            rec_id = "_TO." + cig_text(ops, ri.op_first, ri.eo0) + ".";
            for (uint64_t k = ri.op_end; k > ri.eo1; k--) rec_id += cig_text(ops, k - 1, k);
        }
        if (rec_id != orec.id || rec_id.size() != ri.id_len) {
            fprintf(stderr, "FAIL id suffix rec %zu: '%s' vs '%s' (id_len %u)\n", r, rec_id.c_str(), orec.id.c_str(), ri.id_len);
            n_fail++;
        }
        if (ri.t_st != orec.t_st || ri.t_en != orec.t_en || ri.q_st != orec.q_st || ri.q_en != orec.q_en) {
            fprintf(stderr, "FAIL stripped coords rec %zu\n%s\n", r, tr.line.c_str());
            n_fail++;
        }

        const int n_win = 6;
        for (int wdx = 0; wdx < n_win; wdx++) {
            uint64_t a = U(tr.t_st > 3 ? tr.t_st - 3 : 0, tr.t_en + 2), b = U(tr.t_st > 3 ? tr.t_st - 3 : 0, tr.t_en + 3);
            if (a > b) std::swap(a, b);
            if (wdx == 0) { a = 0; b = tr.t_en + 5; }
            orc::Region rg;
            rg.name = "chrT"; rg.st = a; rg.en = b; rg.id = "W" + std::to_string(wdx);
            if (!orec.overlaps(rg)) continue;
            if (!(ri.t_en > a && ri.t_st < b)) { fprintf(stderr, "FAIL overlap mismatch\n"); n_fail++; continue; }
            for (int policy = 0; policy < 2; policy++) {
                n_pairs++;
                orc::PafRecord ot;
                bool some = false, opanic = false;
                try { some = orc::trim_paf_rec_to_rgn(rg, orec, policy, ot); }
                catch (const orc::Abort&) { opanic = true; }
                PairRes pr;
                uint32_t lerr = lift_pair(view, ri, a, b, policy, true, pr, acc);
                if ((lerr != LIFT_OK) != opanic) {
                    fprintf(stderr, "FAIL panic mismatch rec %zu win %lu-%lu lerr=%u opanic=%d\n%s\n", r, a, b, lerr, (int)opanic, tr.line.c_str());
                    n_fail++;
                    continue;
                }
                if (opanic) continue;
                if ((pr.kind != PK_DROP) != some) {
                    fprintf(stderr, "FAIL some/none mismatch rec %zu win %lu-%lu policy %d kind=%u some=%d\n%s\n", r, a, b, policy, pr.kind, (int)some, tr.line.c_str());
                    n_fail++;
                    continue;
                }
                if (!some) { n_drop++; continue; }
                std::string cg, id;
                if (pr.kind == PK_EARLY) {
                    n_early++;
                    cg = cig_text(ops, pr.si, pr.ei + 1);
                    id = rec_id;
                } else {
                    n_trim++;
                    id = rg.id;
                    if (ri.flags & RF_SLOW) {
                        merged_walk(view, pr.si, pr.ei, pr.s_len, pr.e_len,
                                    [&](uint32_t len, uint32_t c) { cg += std::to_string(len) + OPC[c]; });
                    } else if (pr.si == pr.ei) {
                        cg = std::to_string(pr.s_len) + OPC[op_code(ops[pr.si])];
                    } else {
                        cg = std::to_string(pr.s_len) + OPC[op_code(ops[pr.si])] + cig_text(ops, pr.si + 1, pr.ei) +
                             std::to_string(pr.e_len) + OPC[op_code(ops[pr.ei])];
                    }
                }
                std::string line = "q" + std::to_string(r) + "\t" + std::to_string(tr.q_len) + "\t" + std::to_string(pr.q_st) +
                                   "\t" + std::to_string(pr.q_en) + "\t" + tr.strand + "\tchrT\t" + std::to_string(tr.t_len) +
                                   "\t" + std::to_string(pr.t_st) + "\t" + std::to_string(pr.t_en) + "\t" +
                                   std::to_string(pr.nmatch) + "\t" + std::to_string(pr.aln_len) + "\t60\tid:Z:" + id +
                                   "\tcg:Z:" + cg;
                const std::string want = ot.to_line();
                bool ok = line == want;
                if (ok && cg.size() != pr.cg_bytes) { ok = false; fprintf(stderr, "cg_bytes %u != %zu\n", pr.cg_bytes, cg.size()); }
                if (ok) {
                    uint32_t lb = line_bytes(ri, pr, (uint32_t)("q" + std::to_string(r)).size(), 4, (uint32_t)id.size());
                    if (lb != line.size() + 1) { ok = false; fprintf(stderr, "line_bytes %u != %zu\n", lb, line.size() + 1); }
                }
                if (ok) {  // fused stats == rb stats --paf on the emitted row
                    orc::Stats st;
                    orc::add_stats_from_cigar(ot.cigar, st);
                    if (st.equal != pr.equal || st.diff != pr.diff || st.ins != pr.ins || st.del != pr.del ||
                        st.matches != pr.matches || st.ins_events != pr.ins_ev || st.del_events != pr.del_ev) {
                        ok = false;
                        fprintf(stderr, "stats mismatch\n");
                    }
                }
                if (!ok) {
                    n_fail++;
                    if (n_fail < 20)
                        fprintf(stderr, "FAIL rec %zu win %lu-%lu policy %d\n in : %s\n got: %s\n want:%s\n", r, a, b, policy,
                                tr.line.c_str(), line.c_str(), want.c_str());
                }
            }
        }

        // ---- streaming driver (stream_core.cuh, what k_scan_lift runs): sorted windows merged against the op stream,
        //      chunked exactly like the kernel (32-op chunks of the GLOBAL op index); must equal lift_pair bit for bit ----
        {
            std::vector<std::pair<uint64_t, uint64_t>> wins;
            const int nw = (int)U(1, 12);
            for (int k = 0; k < nw; k++) {
                uint64_t a = U(ri.t_st > 3 ? ri.t_st - 3 : 0, ri.t_en + 1), b = U(ri.t_st > 3 ? ri.t_st - 3 : 0, ri.t_en + 3);
                if (a > b) std::swap(a, b);
                if (k == 0 && U(0, 3) == 0) { a = 0; b = ri.t_en + 5; }
                if (ri.t_en > a && ri.t_st < b) wins.push_back({a, b});
            }
            std::sort(wins.begin(), wins.end());
            for (size_t k = 1; k < wins.size(); k++) wins[k].second = std::max(wins[k].second, wins[k - 1].second);  // en monotone
            std::vector<uint64_t> wst, wen;
            for (auto& w : wins) if (ri.t_en > w.first && ri.t_st < w.second) { wst.push_back(w.first); wen.push_back(w.second); }
            const uint32_t W = (uint32_t)wst.size();
            HalfS junk_s; HalfE junk_e;
            memset(&junk_s, 0xAB, sizeof junk_s); memset(&junk_e, 0xCD, sizeof junk_e);
            std::vector<HalfS> hs(W, junk_s);
            std::vector<HalfE> he(W, junk_e);
            std::vector<int> ws(W, 0), we(W, 0);
            SegRec sr{ri.eo0, ri.eo1, 0u, W, 0u, W, 0u, W, 0ull};
            WinGlobal wa{wst.data(), wen.data(), ri.t_st, ri.t_en};
            Ctr run = ctr_zero();
            uint64_t k = ri.op_first;
            while (k < ri.op_end) {
                const uint64_t kend = std::min<uint64_t>(ri.op_end, ((k >> SAMPLE_LOG2) + 1) << SAMPLE_LOG2);
                uint32_t Tseg = 0;
                for (uint64_t t = k; t < kend; t++) if (is_ref(op_code(ops[t]))) Tseg += op_len(ops[t]);
                stream_segment(view, sr, wa, k, (uint32_t)(kend - k), [&](uint32_t j) { return ops[k + j]; }, run, Tseg, acc,
                               [&](uint64_t p, const HalfS& h) { hs[p] = h; ws[p]++; },
                               [&](uint64_t p, const HalfE& h) { he[p] = h; we[p]++; });
                for (uint64_t t = k; t < kend; t++) ctr_add_op(run, ops[t]);
                k = kend;
            }
            for (uint32_t j = 0; j < W; j++) {
                n_stream++;
                PairRes a, b;
                const uint32_t e1 = lift_pair(view, ri, wst[j], wen[j], POLICY_RIGHTMOST, true, a, acc);
                const uint32_t e2 = combine_pair(view, ri, wst[j], wen[j], hs[j], he[j], b);
                const bool early = ri.t_st > wst[j] && ri.t_en < wen[j];
                // every boundary of a non-early pair is resolved exactly once (early rows never read their halves)
                bool ok = e1 == e2 && memcmp(&a, &b, sizeof a) == 0 && ws[j] <= 1 && we[j] <= 1 && (early || (ws[j] == 1 && we[j] == 1));
                if (!ok) {
                    n_fail++;
                    if (n_fail < 20)
                        fprintf(stderr, "FAIL stream rec %zu win %lu-%lu: e %u/%u kind %u/%u writes %d/%d t %lu-%lu / %lu-%lu si %lu/%lu ei %lu/%lu\n%s\n", r,
                                wst[j], wen[j], e1, e2, a.kind, b.kind, ws[j], we[j], a.t_st, a.t_en, b.t_st, b.t_en, a.si, b.si, a.ei, b.ei,
                                tr.line.c_str());
                }
            }
        }
    }
    printf("records=%zu aborts=%ld slow=%ld stream=%ld pairs=%ld trim=%ld early=%ld drop=%ld FAIL=%ld\n", recs.size(), n_abort, n_slow,
           n_stream, n_pairs, n_trim, n_early, n_drop, n_fail);
    return n_fail ? 1 : 0;
}
