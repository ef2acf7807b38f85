// trim_kernels.cu — sm_100a kernels of `rb trim-paf` (reference: trim_overlap.rs:6-86, paf.rs:210-305,
// paf.rs:564-591, paf.rs:785-823).  The arithmetic is trim_core.cuh (fuzzed on the CPU against the literal oracle); this
// file only spreads it over the machine:
//   k_trim_scan   per record: segmented exclusive scan of (query advance, position score) over its ops -> qp / wp
//                 (12 B per op), the record's total score and its untruncated view
//   k_trim_select per query name and round: containment flags, overlapping pairs, the one pair trimmed this round
//   k_trim_pairs  per selected pair of one round (a column of blocks each): every thread evaluates the split-point
//                 candidates of a stride of both records' ops in the overlap; warp arg-max, atomicMax on a packed key
//   k_trim_cut    per selected pair: the split point from the key, then the two truncations
//   k_trim_round_end  one thread: nothing had to wait -> done (later rounds of the batch are no-ops)
//   k_trim_rows   per record after the last round: the printed row (PairRes + line size) for the shared serialiser
#include <climits>

#include "rb_kernels.cuh"
#include "trim_core.cuh"

namespace rb {

constexpr int TSCAN_THREADS = 256;
constexpr int TSCAN_ITEMS = 4;
constexpr int TPAIR_THREADS = 128;

// Block-wide exclusive scan of the position scores of record `ri` -> wp[op_first .. op_end); returns the total (every thread).
// view == nullptr: right-most policy, the scores do not depend on the record's truncation (trim_w_op);
// else: early-exit policy, the scores of the positions in front of non-query runs follow the view (trim_w_op_view).
__device__ __forceinline__ long long trim_scan_w(const OpsView& v, const RecInfo& ri, const TrimScores& sc, const TrimArr& arr,
                                                 const TrimView* view, long long* wp) {
    __shared__ long long s_w[TSCAN_THREADS / 32];
    const uint64_t first = ri.op_first, end = ri.op_end, eo0 = ri.eo0, eo1 = ri.eo1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long carry_w = 0;
    for (uint64_t base = first; base < end; base += (uint64_t)TSCAN_THREADS * TSCAN_ITEMS) {
        const uint64_t k0 = base + (uint64_t)tid * TSCAN_ITEMS;
        long long dw[TSCAN_ITEMS];
        long long tw = 0;
#pragma unroll
        for (int j = 0; j < TSCAN_ITEMS; j++) {
            const uint64_t k = k0 + j;
            dw[j] = 0;
            if (k < end && k >= eo0 && k < eo1) dw[j] = view ? trim_w_op_view(v, arr, *view, k, sc) : trim_w_op(v, k, eo1, sc);
            tw += dw[j];
        }
        long long iw = tw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long ow = __shfl_up_sync(0xffffffffu, iw, d);
            if (lane >= d) iw += ow;
        }
        if (lane == 31) s_w[warp] = iw;
        __syncthreads();
        long long pw = carry_w, allw = 0;
        for (int x = 0; x < TSCAN_THREADS / 32; x++) {
            if (x < warp) pw += s_w[x];
            allw += s_w[x];
        }
        pw += iw - tw;
#pragma unroll
        for (int j = 0; j < TSCAN_ITEMS; j++) {
            const uint64_t k = k0 + j;
            if (k < end) wp[k] = pw;
            pw += dw[j];
        }
        carry_w += allw;
        __syncthreads();
    }
    return carry_w;
}

// per record: qp (query bases before each op), ap (alignment columns before each op), the untruncated view, then wp
__global__ void __launch_bounds__(TSCAN_THREADS)
k_trim_scan(const uint32_t* __restrict__ ops, const RecInfo* __restrict__ recs, uint32_t n_rec, TrimScores sc, uint32_t* qp, uint32_t* ap,
            long long* wp, TrimView* __restrict__ views, int policy) {
    __shared__ uint32_t s_q[TSCAN_THREADS / 32], s_a[TSCAN_THREADS / 32];
    __shared__ TrimView s_tv;
    const uint32_t r = blockIdx.x;
    if (r >= n_rec) return;
    const RecInfo& ri = recs[r];
    const uint64_t first = ri.op_first, end = ri.op_end, eo0 = ri.eo0, eo1 = ri.eo1;
    OpsView v;
    v.ops = ops; v.samples = nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t carry_q = 0, carry_a = 0;
    for (uint64_t base = first; base < end; base += (uint64_t)TSCAN_THREADS * TSCAN_ITEMS) {
        const uint64_t k0 = base + (uint64_t)tid * TSCAN_ITEMS;
        uint32_t dq[TSCAN_ITEMS], da[TSCAN_ITEMS];
        uint32_t tq = 0, ta = 0;
#pragma unroll
        for (int j = 0; j < TSCAN_ITEMS; j++) {
            const uint64_t k = k0 + j;
            dq[j] = 0; da[j] = 0;
            if (k < end) {
                const uint32_t w = ops[k];
                if (is_qry(op_code(w))) dq[j] = op_len(w);
                da[j] = op_len(w);
            }
            tq += dq[j]; ta += da[j];
        }
        uint32_t iq = tq, ia = ta;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t oq = __shfl_up_sync(0xffffffffu, iq, d);
            const uint32_t oa = __shfl_up_sync(0xffffffffu, ia, d);
            if (lane >= d) { iq += oq; ia += oa; }
        }
        if (lane == 31) { s_q[warp] = iq; s_a[warp] = ia; }
        __syncthreads();
        uint32_t pq = carry_q, pa = carry_a, allq = 0, alla = 0;
        for (int x = 0; x < TSCAN_THREADS / 32; x++) {
            if (x < warp) { pq += s_q[x]; pa += s_a[x]; }
            allq += s_q[x]; alla += s_a[x];
        }
        pq += iq - tq; pa += ia - ta;
#pragma unroll
        for (int j = 0; j < TSCAN_ITEMS; j++) {
            const uint64_t k = k0 + j;
            if (k < end) { qp[k] = pq; ap[k] = pa; }
            pq += dq[j]; pa += da[j];
        }
        carry_q += allq; carry_a += alla;
        __syncthreads();
    }
    const bool live = end > first && eo1 > eo0;  // (block-uniform)
    if (tid == 0 && live) {
        trim_view_init(v, ri, s_tv);
        s_tv.x_end = (eo1 < end) ? qp[eo1] : carry_q;  // written by this block before the barrier above
    }
    __syncthreads();  // s_tv, and this block's qp / ap writes, are visible to all of its threads
    const TrimArr arr{qp, wp, ap, policy};
    const long long w_tot = trim_scan_w(v, ri, sc, arr, (policy == POLICY_EARLY_EXIT && live) ? &s_tv : nullptr, wp);
    if (tid == 0 && live) {
        s_tv.w_tot = w_tot;
        views[r] = s_tv;
    }
}

typedef TrimPairSelDev TrimPairDev;

// per-call round state on the device (zeroed before the first round)
struct TrimInfo {
    uint32_t waiting;  // != 0: some pair of the current round has to wait for the next one (paf.rs:283-285)
    uint32_t done;     // a round ended with nothing waiting: every later launch of this call is a no-op
    uint32_t rounds;   // rounds run so far
    uint32_t status;   // 0, TRIM_ST_ABORT (a truncation the reference panics on) or TRIM_ST_RANGE (overlap x score >= 2^31)
    uint32_t err_l, err_r;  // the pair that set `status`
    uint32_t last_waiting;  // `waiting` of the round that ended last (read by the host in the stepping form)
    uint32_t pad;
};
enum : uint32_t { TRIM_ST_ABORT = 1, TRIM_ST_RANGE = 2 };

// One block per query name (records [grp_off[g], grp_off[g+1]) of the name-sorted set): containment flags, the number of
// overlapping pairs and the one trimmed this round — max over trim_sel_key (trim_core.cuh: trim_select_group is the
// sequential statement, fuzzed against the host's trim_round).  O(m^2) span comparisons per group, like the reference.
__global__ void __launch_bounds__(128)
k_trim_select(const uint32_t* __restrict__ grp_off, uint32_t n_groups, const TrimView* __restrict__ views, uint8_t* __restrict__ contained,
              TrimPairDev* __restrict__ sel, unsigned long long* __restrict__ keys, TrimInfo* __restrict__ info, unsigned long long max_score) {
    __shared__ unsigned long long s_best;
    __shared__ uint32_t s_cnt, s_skip;
    // (another block of this launch may set `status` at any time: one thread reads it for the whole block)
    if (threadIdx.x == 0) s_skip = info->done | info->status;
    __syncthreads();
    if (s_skip) return;
    const uint32_t g = blockIdx.x;
    if (g >= n_groups) return;
    const uint32_t lo = grp_off[g], hi = grp_off[g + 1], m = hi - lo;
    if (threadIdx.x == 0) { s_best = 0; s_cnt = 0; keys[g] = 0; }
    for (uint32_t r = lo + threadIdx.x; r < hi; r += blockDim.x) contained[r] = 0;
    __syncthreads();
    unsigned long long best = 0;
    uint32_t cnt = 0;
    for (uint32_t i = 0; i + 1 < m; i++) {
        const uint64_t i_st = views[lo + i].q_st, i_en = views[lo + i].q_en;
        bool i_contained = false;
        for (uint32_t j = i + 1 + threadIdx.x; j < m; j += blockDim.x) {
            uint64_t ov;
            const uint32_t c = trim_pair_class(i_st, i_en, views[lo + j].q_st, views[lo + j].q_en, ov);
            if (c == TRIM_PAIR_J_CONTAINED) contained[lo + j] = 1;
            else if (c == TRIM_PAIR_I_CONTAINED) i_contained = true;
            else if (c == TRIM_PAIR_PARTIAL) {
                cnt++;
                const unsigned long long k = trim_sel_key(ov, i, j, m);
                if (k > best) best = k;
            }
        }
        if (i_contained) contained[lo + i] = 1;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, d);
        if (ob > best) best = ob;
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0 && cnt) { atomicMax(&s_best, best); atomicAdd(&s_cnt, cnt); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    TrimPairDev one;
    one.left = TRIM_SEL_NONE; one.right = 0; one.st_ovl = one.en_ovl = 0;
    if (s_cnt) {
        one = trim_sel_make(views, lo, m, s_best);
        if (s_cnt > 1) atomicOr(&info->waiting, 1u);
        // the reference sums the scores in i32 (trim_overlap.rs:52-69); the candidate key packs L[A,c) - R[A,c), which can reach
        // TWICE overlap x score, into 32 bits (trim_core.cuh trim_key): refuse what could leave that range
        if (2ull * (one.en_ovl - one.st_ovl) * max_score >= (1ull << 31)) {
            if (atomicCAS(&info->status, 0u, TRIM_ST_RANGE) == 0u) { info->err_l = one.left; info->err_r = one.right; }
            one.left = TRIM_SEL_NONE;
        }
    }
    sel[g] = one;
}

// grid = (query names, slices): every thread of every slice of a name's selected pair evaluates a stride of both records'
// ops in the overlap (coalesced: consecutive threads take consecutive ops); warp shuffle arg-max, then one atomicMax per
// warp on the pair's packed key (trim_key; zeroed by k_trim_select).  Slice 0 contributes the fixed candidates, so the key
// of a selected pair never stays zero.
__global__ void __launch_bounds__(TPAIR_THREADS)
k_trim_pairs(const TrimPairDev* __restrict__ sel, uint32_t n_sel, const uint32_t* __restrict__ ops, const RecInfo* __restrict__ recs,
             const uint32_t* __restrict__ qp, const long long* __restrict__ wp, const uint32_t* __restrict__ ap, int policy, TrimScores sc,
             const TrimView* __restrict__ views, unsigned long long* __restrict__ keys, const TrimInfo* __restrict__ info) {
    if (info->done | info->status) return;
    const uint32_t p = blockIdx.x;
    if (p >= n_sel) return;
    const TrimPairDev ps = sel[p];
    if (ps.left == TRIM_SEL_NONE) return;
    const RecInfo& rl = recs[ps.left];
    const RecInfo& rr = recs[ps.right];
    const TrimView tl = views[ps.left], tr = views[ps.right];
    OpsView v;
    v.ops = ops; v.samples = nullptr;
    const TrimArr a{qp, wp, ap, policy};
    const uint64_t A = ps.st_ovl, B = ps.en_ovl;
    const TrimSide sl = trim_side(v, a, rl, tl, A, sc), sr = trim_side(v, a, rr, tr, A, sc);
    const uint32_t t = blockIdx.y * TPAIR_THREADS + threadIdx.x, n = gridDim.y * TPAIR_THREADS;
    TrimBest best{LLONG_MIN, 0};
    if (t == 0) trim_fixed_candidates(v, a, rl, tl, rr, tr, A, B, sc, best);
    trim_scan_candidates(v, a, true, rl, tl, sl, rr, tr, sr, A, B, sc, t, n, best);
    trim_scan_candidates(v, a, false, rl, tl, sl, rr, tr, sr, A, B, sc, t, n, best);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const long long ot = __shfl_xor_sync(0xffffffffu, best.total, d);
        const unsigned long long oc = __shfl_xor_sync(0xffffffffu, (unsigned long long)best.c, d);
        trim_best_merge(best, ot, oc);
    }
    if ((threadIdx.x & 31) == 0 && best.total != LLONG_MIN) atomicMax(&keys[p], trim_key(best, A));
}

// one thread per selected pair: split point from the reduced key, then the two truncations (trim_overlap.rs:71-79)
__global__ void __launch_bounds__(128)
k_trim_cut(const TrimPairDev* __restrict__ sel, uint32_t n_sel, const uint32_t* __restrict__ ops, const RecInfo* __restrict__ recs,
           const uint32_t* __restrict__ qp, const long long* __restrict__ wp, const uint32_t* __restrict__ ap, int policy, TrimScores sc,
           TrimView* __restrict__ views, const unsigned long long* __restrict__ keys, TrimInfo* __restrict__ info) {
    if (info->done | info->status) return;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_sel) return;
    const TrimPairDev ps = sel[p];
    if (ps.left == TRIM_SEL_NONE) return;
    const RecInfo& rl = recs[ps.left];
    const RecInfo& rr = recs[ps.right];
    OpsView v;
    v.ops = ops; v.samples = nullptr;
    const TrimArr a{qp, wp, ap, policy};
    const uint64_t A = ps.st_ovl, B = ps.en_ovl;
    TrimView nl = views[ps.left], nr = views[ps.right];
    const TrimBest best = trim_unkey(keys[p], A);
    const long long r_tot = trim_S(v, a, rr, nr, A, B, sc);
    const uint64_t s = trim_split(best, r_tot, A);
    uint32_t st = trim_truncate(v, a, rl, nl, nl.q_st, s);                // trim_overlap.rs:78
    if (st == TRIM_OK) st = trim_truncate(v, a, rr, nr, s, nr.q_en);      // trim_overlap.rs:79
    if (st == TRIM_OK) { views[ps.left] = nl; views[ps.right] = nr; }
    else if (atomicCAS(&info->status, 0u, TRIM_ST_ABORT) == 0u) { info->err_l = ps.left; info->err_r = ps.right; }
}
// NOTE on k_trim_cut's early exit: `status` may be set by another block of the same launch; the pairs of a round are
// disjoint and a failed call produces no output, so it does not matter which of them still get cut.

// early-exit policy only: the score prefix of a record follows its view (trim_core.cuh, trim_probe_op) — the two records of every
// pair cut in this round are scanned again.  Block 2g / 2g + 1 = left / right record of query name g's pair.
__global__ void __launch_bounds__(TSCAN_THREADS)
k_trim_rescan(const TrimPairDev* __restrict__ sel, uint32_t n_sel, const uint32_t* __restrict__ ops, const RecInfo* __restrict__ recs,
              const uint32_t* __restrict__ qp, long long* wp, const uint32_t* __restrict__ ap, TrimScores sc, TrimView* views,
              const TrimInfo* __restrict__ info) {
    __shared__ TrimView s_tv;
    __shared__ uint32_t s_r;
    if (threadIdx.x == 0) {
        s_r = TRIM_SEL_NONE;
        const uint32_t g = blockIdx.x >> 1;
        if (!(info->done | info->status) && g < n_sel) {
            const TrimPairDev ps = sel[g];
            if (ps.left != TRIM_SEL_NONE) { s_r = (blockIdx.x & 1u) ? ps.right : ps.left; s_tv = views[s_r]; }
        }
    }
    __syncthreads();
    const uint32_t r = s_r;
    if (r == TRIM_SEL_NONE) return;
    OpsView v;
    v.ops = ops; v.samples = nullptr;
    const TrimArr arr{qp, wp, ap, POLICY_EARLY_EXIT};
    const long long w_tot = trim_scan_w(v, recs[r], sc, arr, &s_tv, wp);
    if (threadIdx.x == 0) views[r].w_tot = w_tot;
}

// after the cuts of a round: nothing waiting -> the call is done (paf.rs:283-300); else the next round starts over
__global__ void k_trim_round_end(TrimInfo* info, int auto_done) {
    if (info->done | info->status) return;
    info->rounds++;
    info->last_waiting = info->waiting;
    if (auto_done && info->waiting == 0) info->done = 1;
    info->waiting = 0;
}

__global__ void __launch_bounds__(128)
k_trim_rows(uint32_t n_rec, const RecInfo* __restrict__ recs, const TrimView* __restrict__ views, const uint32_t* __restrict__ ops,
            const Ctr* __restrict__ samples, const uint8_t* __restrict__ dropped, PairRes* __restrict__ res, uint32_t* __restrict__ line_len,
            uint64_t* __restrict__ pair_off, LiftPlan* __restrict__ plans) {
    __shared__ uint32_t s_acc[9 * 128];
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) pair_off[n_rec] = n_rec;
    if (r >= n_rec) return;
    OpsView v;
    v.ops = ops; v.samples = samples;
    ClassAcc acc;
    acc.sum = s_acc + threadIdx.x; acc.stride = 128;
    const RecInfo& ri = recs[r];
    PairRes pr;
    if (dropped && dropped[r]) pair_clear(pr);  // --remove-contained (paf.rs:290-300)
    else trim_row(v, ri, views[r], acc, pr);
    res[r] = pr;
    line_len[r] = pr.kind == PK_DROP ? 0u : ri.line_const + line_var_bytes(pr, ri.id_len);
    pair_off[r] = r;
    if (r % LIFT_THREADS == 0) {
        LiftPlan pl;
        pl.k0 = r; pl.uniform = 0u; pl.c_lo = pl.c_hi = ~0ull;
        plans[r / LIFT_THREADS] = pl;
    }
}

void launch_trim_scan(const uint32_t* ops, const RecInfo* recs, uint32_t n_rec, TrimScores sc, uint32_t* qp, uint32_t* ap, long long* wp,
                      TrimView* views, int policy, cudaStream_t s) {
    if (n_rec) k_trim_scan<<<n_rec, TSCAN_THREADS, 0, s>>>(ops, recs, n_rec, sc, qp, ap, wp, views, policy);
}
void launch_trim_rounds(int n_rounds, bool auto_done, const uint32_t* grp_off, uint32_t n_groups, const uint32_t* ops, const RecInfo* recs, const uint32_t* qp,
                        long long* wp, const uint32_t* ap, int policy, TrimScores sc, unsigned long long max_score, TrimView* views,
                        uint8_t* contained, void* sel, unsigned long long* keys, void* info, cudaStream_t s) {
    static_assert(sizeof(TrimPairDev) == 24 && sizeof(TrimInfo) == 32, "layouts");
    TrimInfo* inf = reinterpret_cast<TrimInfo*>(info);
    TrimPairDev* sl = reinterpret_cast<TrimPairDev*>(sel);
    // enough slices per pair to fill the machine a few times over (148 SMs): few names -> many slices each
    uint32_t slices = n_groups ? (148u * 8u + n_groups - 1) / n_groups : 1u;
    if (slices > 64u) slices = 64u;
    if (slices < 1u) slices = 1u;
    for (int r = 0; r < n_rounds; r++) {
        if (n_groups) {
            k_trim_select<<<n_groups, 128, 0, s>>>(grp_off, n_groups, views, contained, sl, keys, inf, max_score);
            k_trim_pairs<<<dim3(n_groups, slices), TPAIR_THREADS, 0, s>>>(sl, n_groups, ops, recs, qp, wp, ap, policy, sc, views, keys, inf);
            k_trim_cut<<<(n_groups + 127) / 128, 128, 0, s>>>(sl, n_groups, ops, recs, qp, wp, ap, policy, sc, views, keys, inf);
            if (policy == POLICY_EARLY_EXIT) k_trim_rescan<<<2 * n_groups, TSCAN_THREADS, 0, s>>>(sl, n_groups, ops, recs, qp, wp, ap, sc, views, inf);
        }
        k_trim_round_end<<<1, 1, 0, s>>>(inf, auto_done ? 1 : 0);
    }
}
void launch_trim_rows(uint32_t n_rec, const RecInfo* recs, const TrimView* views, const uint32_t* ops, const Ctr* samples,
                      const uint8_t* dropped, PairRes* res, uint32_t* line_len, uint64_t* pair_off, LiftPlan* plans, cudaStream_t s) {
    k_trim_rows<<<n_rec / 128 + 1, 128, 0, s>>>(n_rec, recs, views, ops, samples, dropped, res, line_len, pair_off, plans);
}

}  // namespace rb
