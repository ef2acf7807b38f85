// trim_kernels.cu — sm_100a kernels of `rb trim-paf` (reference: trim_overlap.rs:6-86, paf.rs:210-305,
// paf.rs:564-591, paf.rs:785-823).  The arithmetic is trim_core.cuh (fuzzed on the CPU against the literal oracle); this
// file only spreads it over the machine:
//   k_trim_scan   per record: segmented exclusive scan of (query advance, position score) over its ops -> qp / wp
//                 (12 B per op), the record's total score and its untruncated view
//   k_trim_pairs  per selected pair of one round (a column of blocks each): every thread evaluates the split-point
//                 candidates of a stride of both records' ops in the overlap; warp arg-max, atomicMax on a packed key
//   k_trim_cut    per selected pair: the split point from the key, then the two truncations
//   k_trim_rows   per record after the last round: the printed row (PairRes + line size) for the shared serialiser
#include <climits>

#include "rb_kernels.cuh"
#include "trim_core.cuh"

namespace rb {

constexpr int TSCAN_THREADS = 256;
constexpr int TSCAN_ITEMS = 4;
constexpr int TPAIR_THREADS = 128;

__global__ void __launch_bounds__(TSCAN_THREADS)
k_trim_scan(const uint32_t* __restrict__ ops, const RecInfo* __restrict__ recs, uint32_t n_rec, TrimScores sc, uint32_t* __restrict__ qp,
            long long* __restrict__ wp, TrimView* __restrict__ views) {
    __shared__ uint32_t s_q[TSCAN_THREADS / 32];
    __shared__ long long s_w[TSCAN_THREADS / 32];
    const uint32_t r = blockIdx.x;
    if (r >= n_rec) return;
    const RecInfo& ri = recs[r];
    const uint64_t first = ri.op_first, end = ri.op_end, eo0 = ri.eo0, eo1 = ri.eo1;
    OpsView v;
    v.ops = ops; v.samples = nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t carry_q = 0;
    long long carry_w = 0;
    for (uint64_t base = first; base < end; base += (uint64_t)TSCAN_THREADS * TSCAN_ITEMS) {
        const uint64_t k0 = base + (uint64_t)tid * TSCAN_ITEMS;
        uint32_t dq[TSCAN_ITEMS];
        long long dw[TSCAN_ITEMS];
        uint32_t tq = 0;
        long long tw = 0;
#pragma unroll
        for (int j = 0; j < TSCAN_ITEMS; j++) {
            const uint64_t k = k0 + j;
            dq[j] = 0; dw[j] = 0;
            if (k < end) {
                const uint32_t w = ops[k];
                if (is_qry(op_code(w))) dq[j] = op_len(w);
                if (k >= eo0 && k < eo1) dw[j] = trim_w_op(v, k, eo1, sc);
            }
            tq += dq[j]; tw += dw[j];
        }
        // block-wide exclusive scan of the per-thread sums: warp shuffles, then the warp totals
        uint32_t iq = tq;
        long long iw = tw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t oq = __shfl_up_sync(0xffffffffu, iq, d);
            const long long ow = __shfl_up_sync(0xffffffffu, iw, d);
            if (lane >= d) { iq += oq; iw += ow; }
        }
        if (lane == 31) { s_q[warp] = iq; s_w[warp] = iw; }
        __syncthreads();
        uint32_t pq = carry_q;
        long long pw = carry_w;
        for (int x = 0; x < warp; x++) { pq += s_q[x]; pw += s_w[x]; }
        uint32_t allq = 0;
        long long allw = 0;
        for (int x = 0; x < TSCAN_THREADS / 32; x++) { allq += s_q[x]; allw += s_w[x]; }
        pq += iq - tq; pw += iw - tw;
#pragma unroll
        for (int j = 0; j < TSCAN_ITEMS; j++) {
            const uint64_t k = k0 + j;
            if (k < end) { qp[k] = pq; wp[k] = pw; }
            pq += dq[j]; pw += dw[j];
        }
        carry_q += allq; carry_w += allw;
        __syncthreads();
    }
    if (tid == 0 && end > first && eo1 > eo0) {
        TrimView tv;
        trim_view_init(v, ri, tv);
        tv.w_tot = carry_w;
        tv.x_end = (eo1 < end) ? qp[eo1] : carry_q;  // written by this block before the barrier above
        views[r] = tv;
    }
}

struct TrimPairDev { uint32_t left, right; uint64_t st_ovl, en_ovl; };  // == TrimPairSel (trim_rounds.hpp)
struct TrimPairOut { uint64_t l_st, l_en, r_st, r_en; uint32_t status, pad; };

// grid = (pairs of this round, slices): every thread of every slice of a pair evaluates a stride of both records' ops in the
// overlap (coalesced: consecutive threads take consecutive ops); warp shuffle arg-max, then one atomicMax per warp on the
// pair's packed key (trim_key).  keys[] is zeroed before the launch; slice 0 contributes the fixed candidates, so no key
// stays zero.
__global__ void __launch_bounds__(TPAIR_THREADS)
k_trim_pairs(const TrimPairDev* __restrict__ sel, uint32_t n_sel, const uint32_t* __restrict__ ops, const RecInfo* __restrict__ recs,
             const uint32_t* __restrict__ qp, const long long* __restrict__ wp, TrimScores sc, const TrimView* __restrict__ views,
             unsigned long long* __restrict__ keys) {
    const uint32_t p = blockIdx.x;
    if (p >= n_sel) return;
    const TrimPairDev ps = sel[p];
    const RecInfo& rl = recs[ps.left];
    const RecInfo& rr = recs[ps.right];
    const TrimView tl = views[ps.left], tr = views[ps.right];
    OpsView v;
    v.ops = ops; v.samples = nullptr;
    const TrimArr a{qp, wp};
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

// one thread per pair: split point from the reduced key, then the two truncations (trim_overlap.rs:71-79)
__global__ void __launch_bounds__(128)
k_trim_cut(const TrimPairDev* __restrict__ sel, uint32_t n_sel, const uint32_t* __restrict__ ops, const RecInfo* __restrict__ recs,
           const uint32_t* __restrict__ qp, const long long* __restrict__ wp, TrimScores sc, TrimView* __restrict__ views,
           const unsigned long long* __restrict__ keys, TrimPairOut* __restrict__ out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_sel) return;
    const TrimPairDev ps = sel[p];
    const RecInfo& rl = recs[ps.left];
    const RecInfo& rr = recs[ps.right];
    OpsView v;
    v.ops = ops; v.samples = nullptr;
    const TrimArr a{qp, wp};
    const uint64_t A = ps.st_ovl, B = ps.en_ovl;
    TrimView nl = views[ps.left], nr = views[ps.right];
    const TrimBest best = trim_unkey(keys[p], A);
    const long long r_tot = trim_S(v, a, rr, nr, A, B, sc);
    const uint64_t s = trim_split(best, r_tot, A);
    TrimPairOut o;
    o.pad = 0;
    o.status = trim_truncate(v, a, rl, nl, nl.q_st, s);                           // trim_overlap.rs:78
    if (o.status == TRIM_OK) o.status = trim_truncate(v, a, rr, nr, s, nr.q_en);  // trim_overlap.rs:79
    if (o.status == TRIM_OK) { views[ps.left] = nl; views[ps.right] = nr; }
    o.l_st = nl.q_st; o.l_en = nl.q_en; o.r_st = nr.q_st; o.r_en = nr.q_en;
    out[p] = o;
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

void launch_trim_scan(const uint32_t* ops, const RecInfo* recs, uint32_t n_rec, TrimScores sc, uint32_t* qp, long long* wp, TrimView* views,
                      cudaStream_t s) {
    if (n_rec) k_trim_scan<<<n_rec, TSCAN_THREADS, 0, s>>>(ops, recs, n_rec, sc, qp, wp, views);
}
void launch_trim_pairs(const void* sel, uint32_t n_sel, const uint32_t* ops, const RecInfo* recs, const uint32_t* qp, const long long* wp,
                       TrimScores sc, TrimView* views, unsigned long long* keys, void* out, cudaStream_t s) {
    static_assert(sizeof(TrimPairDev) == 24 && sizeof(TrimPairOut) == 40, "pair layouts");
    if (!n_sel) return;
    // enough slices per pair to fill the machine a few times over (148 SMs): few pairs -> many slices each
    uint32_t slices = (148u * 8u + n_sel - 1) / n_sel;
    if (slices > 64u) slices = 64u;
    if (slices < 1u) slices = 1u;
    cudaMemsetAsync(keys, 0, (size_t)n_sel * 8, s);
    k_trim_pairs<<<dim3(n_sel, slices), TPAIR_THREADS, 0, s>>>(reinterpret_cast<const TrimPairDev*>(sel), n_sel, ops, recs, qp, wp, sc, views, keys);
    k_trim_cut<<<(n_sel + 127) / 128, 128, 0, s>>>(reinterpret_cast<const TrimPairDev*>(sel), n_sel, ops, recs, qp, wp, sc, views, keys,
                                                   reinterpret_cast<TrimPairOut*>(out));
}
void launch_trim_rows(uint32_t n_rec, const RecInfo* recs, const TrimView* views, const uint32_t* ops, const Ctr* samples,
                      const uint8_t* dropped, PairRes* res, uint32_t* line_len, uint64_t* pair_off, LiftPlan* plans, cudaStream_t s) {
    k_trim_rows<<<n_rec / 128 + 1, 128, 0, s>>>(n_rec, recs, views, ops, samples, dropped, res, line_len, pair_off, plans);
}

}  // namespace rb
