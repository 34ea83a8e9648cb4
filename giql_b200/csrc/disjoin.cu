// disjoin.cu — DISJOIN(target[, reference]): split every target interval at the reference's breakpoints.
//
// Semantics (reference, abdenlab/giql src/giql/expanders/disjoin.py:102-202):
//   breakpoints = UNION (distinct) of the reference's starts and ends per chromosome            (:149-154)
//   cuts of a target row = {t.start, t.end} ∪ {bp : t.start < bp < t.end}                         (:159-172)
//   segments = consecutive cuts (LEAD over the cuts of the target key), kept when seg_end > seg_start (:173-178)
//   unless the reference is the target itself, a segment is kept only when some reference row covers its
//   start: r.start <= seg_start AND r.end > seg_start                                              (:184-190)
//   duplicate target rows share a key and each receives every segment                              (:191-199)
// Canonical coordinates in, canonical coordinates out (the host layer re-encodes, :139-141).
//
// Device formulation: the breakpoints are the sorted, de-duplicated linearised starts and ends of the
// reference index (one radix sort of 2n keys + a flag scan).  A target row's interior cuts are then the
// contiguous run bp[lo, hi), lo = upper_bound(bp, t.start), hi = lower_bound(bp, t.end): counting is two
// binary searches, and the coverage test of a segment start p is one more (rows with r.start <= p exist and
// their running max end exceeds p).  count -> exclusive scan -> emit, like the join.
#include "common.cuh"
#include "scan.cuh"

namespace {

using namespace giqlb_scan;

struct DisjoinView {
    const uint32_t* bp;
    int64_t n_bp;
    const uint32_t* rgs;    // reference index: sorted starts and running max of end (coverage test)
    const uint32_t* pmax;
    int64_t n_r;
    int filter;             // 1: keep only covered segments (reference is not the target itself)
    const uint8_t* cov;     // filter only: cov[j] = some reference row covers breakpoint bp[j] (computed once per call, so a
                            // segment that starts at a breakpoint needs one byte instead of a search)
    const uint32_t* cov_before;   // filter only: cov_before[j] = cov[0] + .. + cov[j - 1], j = 0 .. n_bp: the number of kept segments
                                  // among any run of a row's cuts is a difference of two entries (count and output slot in O(1))
    // bucket directory over bp (giqlb_bucket_dir_build): the two searches of every target row start inside one bucket instead of
    // walking ~21 dependent loads through the whole array (the count and emit kernels took 245 + 272 us for 1 M rows that way)
    const uint32_t* bp_dir;
    int bp_shift;
    uint32_t bp_last;
    const uint32_t* r_dir;        // the reference index's own directory over its starts (coverage test)
    int r_shift;
    uint32_t r_last;
};
__device__ __forceinline__ int64_t bp_upper_bound(const DisjoinView& V, uint32_t x) {
    return V.bp_dir ? dir_upper_bound(V.bp_dir, V.bp_shift, V.bp_last, V.bp, x) : upper_bound_dev<uint32_t, int64_t>(V.bp, 0, V.n_bp, x);
}
__device__ __forceinline__ int64_t bp_lower_bound(const DisjoinView& V, uint32_t x) {
    return V.bp_dir ? dir_lower_bound(V.bp_dir, V.bp_shift, V.bp_last, V.bp, x) : lower_bound_dev<uint32_t, int64_t>(V.bp, 0, V.n_bp, x);
}

__device__ __forceinline__ bool covered(const DisjoinView& V, uint32_t p) {
    const int64_t u = V.r_dir ? dir_upper_bound(V.r_dir, V.r_shift, V.r_last, V.rgs, p)
                              : upper_bound_dev<uint32_t, int64_t>(V.rgs, 0, V.n_r, p);   // rows with r.start <= p
    return u > 0 && __ldg(V.pmax + u - 1) > p;
}

// Visits the segments of target row i in order; f(seg_index, start_lin_or_raw...) receives canonical coordinates.
template <typename F>
__device__ __forceinline__ unsigned long long for_each_segment(const ProbeCols& T, const DisjoinView& V, int64_t i, F f) {
    const long long ts = (long long)T.start[i] + T.sadj, te = (long long)T.end[i] + T.eadj;
    if (ts == te) return 0ull;                       // a single cut: LEAD is NULL, no segment
    uint32_t gs, ge;
    const bool on_ref_chrom = V.n_r > 0 && probe_linearise(T, i, gs, ge);
    unsigned long long n = 0;
    if (te < ts) {                                   // cuts {end, start}: one segment (end, start), no interior cut
        if (!V.filter || (on_ref_chrom && covered(V, ge))) {
            f(n, te, ts);
            n++;
        }
        return n;
    }
    if (!on_ref_chrom) {                             // no breakpoint on this chromosome
        if (!V.filter) {
            f(n, ts, te);
            n++;
        }
        return n;
    }
    const int32_t c = T.code[i];
    const long long lo_c = __ldg(T.lo32 + c);
    const uint32_t b1 = __ldg(T.base + c) + 1u;
    const int64_t lo = bp_upper_bound(V, gs);   // first bp >  t.start
    int64_t hi = bp_lower_bound(V, ge);         // first bp >= t.end
    if (hi < lo) hi = lo;
    long long seg_s = ts;
    for (int64_t j = lo; j <= hi; j++) {
        const uint32_t cut_lin = j < hi ? __ldg(V.bp + j) : ge;
        const long long cut = j < hi ? lo_c + (long long)(cut_lin - b1) : te;
        if (!V.filter || (j == lo ? covered(V, gs) : __ldg(V.cov + j - 1) != 0)) {
            f(n, seg_s, cut);
            n++;
        }
        seg_s = cut;
    }
    return n;
}

// A target row cut by more than DJ_HEAVY breakpoints (a chromosome-spanning interval) is not walked by its own thread —
// ~10^5 cuts, each with a coverage search, serialised while the rest of the grid has finished — but by the lanes of its
// warp together: candidate segment j of the row is independent of the others (its start is breakpoint j - 1), so 32 of
// them are tested per step and written in order through a ballot prefix.
constexpr int64_t DJ_HEAVY = 64;

struct RowCuts {
    long long ts, te, lo_c;
    uint32_t gs, ge, b1;
    int64_t lo, hi;          // interior cuts bp[lo, hi): segments j = lo .. hi
};
// true when the row takes the breakpoint walk (proper interval on a chromosome that has reference rows)
__device__ __forceinline__ bool row_cuts(const ProbeCols& T, const DisjoinView& V, int64_t i, RowCuts& r) {
    r.ts = (long long)T.start[i] + T.sadj;
    r.te = (long long)T.end[i] + T.eadj;
    if (r.te <= r.ts || V.n_r <= 0 || !probe_linearise(T, i, r.gs, r.ge)) return false;
    const int32_t c = T.code[i];
    r.lo_c = __ldg(T.lo32 + c);
    r.b1 = __ldg(T.base + c) + 1u;
    r.lo = bp_upper_bound(V, r.gs);
    r.hi = bp_lower_bound(V, r.ge);
    if (r.hi < r.lo) r.hi = r.lo;
    return true;
}
// candidate segment j of a walked row: canonical (start, end) and whether it is kept
__device__ __forceinline__ bool cut_segment(const DisjoinView& V, const RowCuts& r, int64_t j, long long& s, long long& e) {
    const uint32_t s_lin = j == r.lo ? r.gs : __ldg(V.bp + j - 1);
    s = j == r.lo ? r.ts : r.lo_c + (long long)(s_lin - r.b1);
    e = j < r.hi ? r.lo_c + (long long)(__ldg(V.bp + j) - r.b1) : r.te;
    return !V.filter || (j == r.lo ? covered(V, s_lin) : __ldg(V.cov + j - 1) != 0);
}

__global__ void __launch_bounds__(256)
disjoin_cover_kernel(DisjoinView V, uint8_t* __restrict__ cov) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < V.n_bp; j += (int64_t)gridDim.x * blockDim.x)
        cov[j] = covered(V, V.bp[j]) ? 1 : 0;
}

// kept segments of a walked row among j = lo + 1 .. j_last (the ones that start at a breakpoint)
__device__ __forceinline__ unsigned long long kept_after_first(const DisjoinView& V, const RowCuts& r, int64_t j_last) {
    if (j_last <= r.lo) return 0ull;
    return V.filter ? (unsigned long long)(__ldg(V.cov_before + j_last) - __ldg(V.cov_before + r.lo)) : (unsigned long long)(j_last - r.lo);
}

__global__ void __launch_bounds__(256)
disjoin_count_kernel(ProbeCols T, DisjoinView V, unsigned long long* __restrict__ cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) {
        RowCuts r;
        unsigned long long n;
        if (row_cuts(T, V, i, r) && r.hi - r.lo + 1 > DJ_HEAVY) {
            // a heavily cut row is not walked to be counted: the first segment is tested, the others are a difference of prefix counts
            n = (!V.filter || covered(V, r.gs) ? 1ull : 0ull) + kept_after_first(V, r, r.hi);
        } else {
            n = for_each_segment(T, V, i, [](unsigned long long, long long, long long) {});
        }
        cnt[i] = n;
    }
}

struct InCounts {
    const unsigned long long* cnt;
    __device__ unsigned long long operator()(int64_t i) const { return cnt[i]; }
};
struct OutOffsets {
    unsigned long long* off;
    unsigned long long* total;
    int64_t n;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        off[i] = excl;
        if (i == n - 1) *total = excl + v;
    }
};

// Rows cut more than DJ_SUPER times (a chromosome-spanning target: ~10^5 cuts) are not even a warp's job: the emit kernel
// records them, and a second kernel deals their segments to the whole grid — the output slot of segment j is the row's offset
// plus a difference of prefix counts, so every thread can place its segment without knowing the others'.
constexpr int64_t DJ_SUPER = 4096;
constexpr uint32_t DJ_SUPER_MAX = 4096;   // list entries; rows beyond that take the warp path
struct SuperRow {
    RowCuts r;
    unsigned long long o;
    int32_t row;
    uint32_t keep0;
};

__global__ void __launch_bounds__(256)
disjoin_emit_kernel(ProbeCols T, DisjoinView V, const unsigned long long* __restrict__ off, int32_t* __restrict__ out_row,
                    int32_t* __restrict__ out_start, int32_t* __restrict__ out_end, SuperRow* __restrict__ super_rows,
                    uint32_t* __restrict__ n_super) {
    const uint32_t lt_mask = (1u << lane_id()) - 1u;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < T.n; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const bool live = i < T.n;
        RowCuts r;
        bool heavy = live && row_cuts(T, V, i, r) && r.hi - r.lo + 1 > DJ_HEAVY;
        const unsigned long long o = live ? off[i] : 0ull;
        if (heavy && r.hi - r.lo + 1 > DJ_SUPER) {
            const uint32_t slot = atomicAdd(n_super, 1u);
            if (slot < DJ_SUPER_MAX) {
                SuperRow sr;
                sr.r = r;
                sr.o = o;
                sr.row = (int32_t)i;
                sr.keep0 = !V.filter || covered(V, r.gs) ? 1u : 0u;
                super_rows[slot] = sr;
                heavy = false;             // (not walked here)
            }
        } else if (live && !heavy) {
            for_each_segment(T, V, i, [&](unsigned long long k, long long s, long long e) {
                out_row[o + k] = (int32_t)i;
                out_start[o + k] = (int32_t)s;
                out_end[o + k] = (int32_t)e;
            });
        }
        uint32_t todo = __ballot_sync(FULL_MASK, heavy);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            RowCuts b;
            b.ts = __shfl_sync(FULL_MASK, r.ts, src); b.te = __shfl_sync(FULL_MASK, r.te, src); b.lo_c = __shfl_sync(FULL_MASK, r.lo_c, src);
            b.gs = __shfl_sync(FULL_MASK, r.gs, src); b.ge = __shfl_sync(FULL_MASK, r.ge, src); b.b1 = __shfl_sync(FULL_MASK, r.b1, src);
            b.lo = __shfl_sync(FULL_MASK, r.lo, src); b.hi = __shfl_sync(FULL_MASK, r.hi, src);
            const int32_t row = (int32_t)__shfl_sync(FULL_MASK, i, src);
            unsigned long long bo = __shfl_sync(FULL_MASK, o, src);
            for (int64_t j0 = b.lo; j0 <= b.hi; j0 += 128) {   // four 32-segment steps per iteration: their loads are in flight together
                bool keep[4];
                long long sgs[4], sge[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int64_t j = j0 + 32 * u + lane_id();
                    sgs[u] = sge[u] = 0;
                    keep[u] = j <= b.hi && cut_segment(V, b, j, sgs[u], sge[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t m = __ballot_sync(FULL_MASK, keep[u]);
                    if (keep[u]) {
                        const unsigned long long pos = bo + __popc(m & lt_mask);
                        out_row[pos] = row;
                        out_start[pos] = (int32_t)sgs[u];
                        out_end[pos] = (int32_t)sge[u];
                    }
                    bo += __popc(m);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
disjoin_emit_super_kernel(DisjoinView V, const SuperRow* __restrict__ super_rows, const uint32_t* __restrict__ n_super,
                          int32_t* __restrict__ out_row, int32_t* __restrict__ out_start, int32_t* __restrict__ out_end) {
    const uint32_t n = min(*n_super, DJ_SUPER_MAX);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (uint32_t k = 0; k < n; k++) {
        const SuperRow sr = super_rows[k];
        for (int64_t j = sr.r.lo + tid; j <= sr.r.hi; j += stride) {
            long long s, e;
            const bool keep = j == sr.r.lo ? sr.keep0 != 0 : cut_segment(V, sr.r, j, s, e);
            if (!keep) continue;
            if (j == sr.r.lo) cut_segment(V, sr.r, j, s, e);     // (coordinates of the first segment; its coverage was tested once)
            const unsigned long long pos = sr.o + (j == sr.r.lo ? 0ull : sr.keep0 + kept_after_first(V, sr.r, j - 1));
            out_row[pos] = sr.row;
            out_start[pos] = (int32_t)s;
            out_end[pos] = (int32_t)e;
        }
    }
}

__global__ void __launch_bounds__(256)
concat_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n, uint32_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < n ? a[i] : b[i - n];
}

struct InCov {
    const uint8_t* cov;
    __device__ unsigned long long operator()(int64_t i) const { return cov[i]; }
};
struct OutCovBefore {
    uint32_t* before;
    int64_t n;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        before[i] = (uint32_t)excl;
        if (i == n - 1) before[n] = (uint32_t)(excl + v);
    }
};

struct InIsNew {
    const uint32_t* k;
    __device__ unsigned long long operator()(int64_t i) const { return (i == 0 || k[i] != k[i - 1]) ? 1ull : 0ull; }
};
struct OutUnique {
    const uint32_t* k;
    uint32_t* out;
    unsigned long long* total;
    int64_t n;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        if (v) out[excl] = k[i];
        if (i == n - 1) *total = excl + v;
    }
};

}  // namespace

extern "C" int giqlb_disjoin(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* t_code, const int32_t* t_start,
                             const int32_t* t_end, int64_t n_t, int32_t t_start_adj, int32_t t_end_adj, giqlb_index* REF,
                             int self_reference, int32_t* out_row, int32_t* out_start, int32_t* out_end, int64_t capacity,
                             int64_t* n_out_host) {
    if (!ctx || !REF || !n_out_host || n_t < 0 || capacity < 0) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "disjoin: bad argument");
    if (REF->layout != lay) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "disjoin: reference index was built on a different layout");
    if (REF->grouped_by_strand) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "disjoin: strand-grouped index not accepted");
    begin_call(ctx);
    *n_out_host = 0;
    if (n_t == 0) return end_call(ctx);
    ProbeCols T;
    GIQLB_TRY(probe_cols(ctx, lay, t_code, t_start, t_end, n_t, t_start_adj, t_end_adj, &T));
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, REF));
    const int64_t nr = REF->n;
    uint32_t *ka = nullptr, *kb = nullptr, *va = nullptr, *vb = nullptr, *bp = nullptr;
    unsigned long long* off = nullptr;
    unsigned long long* total = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 11);
    int st = GIQLB_OK;
    int64_t n_bp = 0;
    PhaseSpan ps(ctx, GIQLB_PH_DISJOIN);
    if (nr > 0) {
        st = dev_alloc(ctx, &ka, (size_t)2 * nr);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &kb, (size_t)2 * nr);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &va, (size_t)2 * nr);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &vb, (size_t)2 * nr);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &bp, (size_t)2 * nr);
        if (st == GIQLB_OK) {
            int64_t blocks = ceil_div(2 * nr, 1024), cap = (int64_t)ctx->sm_count * 8;
            concat_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(REF->gs, REF->ge, nr, ka);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            int in_b = 0;
            st = giqlb_radix_sort_pairs_u32(ctx, ka, va, kb, vb, 2 * nr, 0, lay->key_bits, 1, &in_b);
            if (st == GIQLB_OK) {
                const uint32_t* sorted = in_b ? kb : ka;
                cudaMemsetAsync(total, 0, 8, ctx->stream);
                st = launch_scan<OpAddU64>(ctx, 2 * nr, InIsNew{sorted}, OutUnique{sorted, bp, total, 2 * nr});
            }
            if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, total, 1);
            if (st == GIQLB_OK) n_bp = ctx->h_scalars[0];
        }
    }
    uint8_t* cov = nullptr;
    DisjoinView V{bp, n_bp, REF->gs, REF->pmax, nr, self_reference ? 0 : 1, nullptr, nullptr, nullptr, 0, 0u, nullptr, 0, 0u};
    if (st == GIQLB_OK && nr > 0 && !self_reference) {
        st = giqlb_index_ensure_dirs(ctx, REF, false, false);
        if (st == GIQLB_OK) {
            V.r_dir = REF->dir_gs;
            V.r_shift = REF->dir_shift;
            V.r_last = REF->dir_stride - 2;
        }
    }
    uint32_t* bp_dir = nullptr;
    if (st == GIQLB_OK && n_bp > 0) {
        uint32_t stride = 0;
        st = giqlb_bucket_dir_build(ctx, bp, n_bp, lay->total_extent, &V.bp_shift, &stride, &bp_dir);
        if (st == GIQLB_OK) {
            V.bp_dir = bp_dir;
            V.bp_last = stride - 2;
        }
    }
    if (st == GIQLB_OK && V.filter && n_bp > 0) {
        st = dev_alloc(ctx, &cov, (size_t)n_bp);
        if (st == GIQLB_OK) {
            int64_t blocks = ceil_div(n_bp, 256), cap = (int64_t)ctx->sm_count * 16;
            disjoin_cover_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(V, cov);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            V.cov = cov;
        }
    }
    uint32_t* cov_before = nullptr;
    if (st == GIQLB_OK && V.cov) {
        st = dev_alloc(ctx, &cov_before, (size_t)n_bp + 1);
        if (st == GIQLB_OK) st = launch_scan<OpAddU64>(ctx, n_bp, InCov{cov}, OutCovBefore{cov_before, n_bp});
        V.cov_before = cov_before;
    }
    SuperRow* super_rows = nullptr;
    uint32_t* n_super = reinterpret_cast<uint32_t*>(ctx->d_scalars + 12);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &super_rows, (size_t)DJ_SUPER_MAX);
    unsigned long long* cnt = nullptr;
    if (st == GIQLB_OK) st = dev_alloc(ctx, &off, (size_t)n_t);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &cnt, (size_t)n_t);
    if (st == GIQLB_OK) {
        cudaMemsetAsync(total, 0, 8, ctx->stream);
        int64_t blocks = ceil_div(n_t, 256), cap = (int64_t)ctx->sm_count * 16;
        disjoin_count_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(T, V, cnt);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        st = launch_scan<OpAddU64>(ctx, n_t, InCounts{cnt}, OutOffsets{off, total, n_t});
    }
    if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, total, 1);
    int64_t n_out = st == GIQLB_OK ? ctx->h_scalars[0] : 0;
    if (st == GIQLB_OK) {
        *n_out_host = n_out;
        if (n_out > capacity) st = giqlb_fail(ctx, GIQLB_ERR_CAPACITY, "disjoin: %lld segments exceed capacity %lld", (long long)n_out, (long long)capacity);
        else if (n_out > 0 && (!out_row || !out_start || !out_end)) st = giqlb_fail(ctx, GIQLB_ERR_INVALID, "disjoin: null output");
    }
    if (st == GIQLB_OK && n_out > 0) {
        int64_t blocks = ceil_div(n_t, 256), cap = (int64_t)ctx->sm_count * 16;
        cudaMemsetAsync(n_super, 0, 4, ctx->stream);
        disjoin_emit_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(T, V, off, out_row, out_start, out_end, super_rows, n_super);
        disjoin_emit_super_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(V, super_rows, n_super, out_row, out_start, out_end);
        ctx->launches_this_call += 2;
        ctx->stats.kernels_total += 2;
        if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "disjoin: emit launch failed");
    }
    dev_free(ctx, ka); dev_free(ctx, kb); dev_free(ctx, va); dev_free(ctx, vb); dev_free(ctx, bp); dev_free(ctx, off); dev_free(ctx, cnt); dev_free(ctx, cov); dev_free(ctx, cov_before); dev_free(ctx, super_rows); dev_free(ctx, bp_dir);
    if (st != GIQLB_OK) return st;
    ctx->stats.rows_in[0] = n_t;
    ctx->stats.rows_in[1] = nr;
    ctx->stats.rows_out = n_out;
    ctx->stats.algorithmic_bytes = 12 * (n_t + nr) + 12 * n_out;
    return end_call(ctx);
}
