// join.cu — column-to-column INTERSECTS: inner join (count -> exclusive scan -> emit), SEMI/ANTI
// existence flags and per-row overlap counts.
//
// Semantics (reference, abdenlab/giql): the pair predicate
//     l.chrom = r.chrom AND l.start < r.end AND l.end > r.start          src/giql/expanders/intersects.py:149-154
// evaluated per chromosome by the DuckDB IEJoin plan                      src/giql/expanders/intersects_duckdb.py:1283-1299
// with bag semantics (duplicate rows give duplicate pairs, tests/test_duckdb_iejoin.py:4514-4555), SEMI/ANTI
// as [NOT] EXISTS (intersects_duckdb.py:1254-1282) and count_overlaps as a per-left-row count
// (intersects_duckdb.py:432-548).  Pair order is unspecified (the reference compares sorted multisets,
// tests/integration/_oracle.py:77-89).
//
// Fast path (both tables hold only proper intervals, start < end).  Over the linearised coordinate g the
// overlapping pairs split into two disjoint classes, each a CONTIGUOUS run of a start-sorted array:
//     class A  (r starts inside l, or at l's start)   l.gs <= r.gs <  l.ge   -> for each l a run of R
//     class B  (l starts strictly inside r)           r.gs <  l.gs <  r.ge   -> for each r a run of L
// so counting is two binary searches per row, no candidate is ever filtered out, and emit work is exactly
// the output size.  Rows are grouped in item tiles of 1024; the count pass stores one total per tile, one
// CTA scans the totals, and the emit pass cuts the combined (items + outputs) work line into equal chunks
// (merge-path partitioning at tile granularity): a chunk recomputes the runs of the tiles it touches and
// writes its slice of the output with a per-slot search in shared memory, so a tile with a huge output is
// shared by many CTAs and a desert of empty tiles is shared out as well.
//
// General path (some row has end <= start): the literal predicate is scanned over the candidates
//     j in [first j with pmax[j] > l.gs,  lower_bound(R.gs, l.ge))        pmax = running max of R.ge
// which is exact for any rows, at the cost of visiting non-matching candidates.
#include "common.cuh"

namespace {

constexpr int JB = 256;             // threads per CTA
constexpr int J_IPT = 4;
constexpr int JT = JB * J_IPT;      // items per tile
constexpr int J_CHUNK = 8192;       // work units (items + outputs) per emit chunk

struct JoinSide {
    const uint32_t* gs;
    const uint32_t* ge;
    const int32_t* rid;
    int64_t n;
};

// first idx >= start with y[idx] >= x, galloping from `start` (runs are short: a few rows on average)
__device__ __forceinline__ int64_t lb_gallop(const uint32_t* __restrict__ y, int64_t start, int64_t n, uint32_t x) {
    if (start >= n || __ldg(y + start) >= x) return start;
    int64_t lo = start, step = 1;   // invariant: y[lo] < x
    while (lo + step < n && __ldg(y + lo + step) < x) {
        lo += step;
        step <<= 1;
    }
    int64_t hi = lo + step < n ? lo + step : n;
    return lower_bound_dev<uint32_t, int64_t>(y, lo + 1, hi, x);
}

// run [lo, hi) of Y matched by item x.  CLASS_B == false: y.gs in [x.gs, x.ge); true: y.gs in (x.gs, x.ge)
template <bool CLASS_B>
__device__ __forceinline__ void item_run(uint32_t xgs, uint32_t xge, const uint32_t* __restrict__ ygs, int64_t ny,
                                         int64_t& lo, int64_t& hi) {
    lo = CLASS_B ? upper_bound_dev<uint32_t, int64_t>(ygs, 0, ny, xgs) : lower_bound_dev<uint32_t, int64_t>(ygs, 0, ny, xgs);
    hi = lb_gallop(ygs, lo, ny, xge);
}

// ---- count pass: one total per item tile ---------------------------------------------------------
__global__ void __launch_bounds__(JB)
join_count_tiles_kernel(JoinSide L, JoinSide R, int64_t tiles_a, unsigned long long* __restrict__ tile_cnt) {
    __shared__ unsigned long long s_red[JB / 32];
    const int64_t tile = blockIdx.x;
    const bool class_b = tile >= tiles_a;
    const JoinSide& X = class_b ? R : L;
    const JoinSide& Y = class_b ? L : R;
    const int64_t base = (class_b ? tile - tiles_a : tile) * JT;
    unsigned long long cnt = 0;
#pragma unroll
    for (int j = 0; j < J_IPT; j++) {
        int64_t i = base + j * JB + threadIdx.x;
        if (i < X.n) {
            uint32_t xgs = X.gs[i], xge = X.ge[i];
            int64_t lo, hi;
            if (class_b) item_run<true>(xgs, xge, Y.gs, Y.n, lo, hi);
            else item_run<false>(xgs, xge, Y.gs, Y.n, lo, hi);
            cnt += (unsigned long long)(hi - lo);
        }
    }
    cnt = warp_reduce_add(cnt);
    if (lane_id() == 0) s_red[warp_id()] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < JB / 32; w++) t += s_red[w];
        tile_cnt[tile] = t;
    }
}

// ---- emit pass -----------------------------------------------------------------------------------
// work line: tile t occupies [t*JT + off[t], (t+1)*JT + off[t+1]); its first JT units stand for the cost of
// recomputing the tile, the remaining off[t+1]-off[t] units are its output slots.
__global__ void __launch_bounds__(JB)
join_emit_kernel(JoinSide L, JoinSide R, int64_t tiles_a, int64_t n_tiles,
                 const unsigned long long* __restrict__ tile_off, int32_t* __restrict__ out_l,
                 int32_t* __restrict__ out_r, int64_t capacity) {
    __shared__ unsigned long long s_off[JT + 1];
    __shared__ uint32_t s_lo[JT];
    __shared__ int32_t s_rid[JT];
    __shared__ unsigned long long s_scratch[JB / 32 + 2];

    const unsigned long long total = tile_off[n_tiles];
    if ((long long)total > capacity) return;   // caller's buffer too small: nothing is written
    const unsigned long long work = (unsigned long long)n_tiles * JT + total;
    const unsigned long long n_chunks = (work + J_CHUNK - 1) / J_CHUNK;

    for (unsigned long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const unsigned long long ws = chunk * J_CHUNK;
        const unsigned long long we = ws + J_CHUNK < work ? ws + J_CHUNK : work;
        // first tile whose work interval ends after ws
        int64_t tlo = 0, thi = n_tiles;
        while (tlo < thi) {
            int64_t mid = tlo + ((thi - tlo) >> 1);
            unsigned long long wend = (unsigned long long)(mid + 1) * JT + tile_off[mid + 1];
            if (wend <= ws) tlo = mid + 1; else thi = mid;
        }
        for (int64_t t = tlo; t < n_tiles; t++) {
            const unsigned long long off_t = tile_off[t], off_t1 = tile_off[t + 1];
            const unsigned long long w_begin = (unsigned long long)t * JT + off_t;
            if (w_begin >= we) break;
            const unsigned long long wo_s = w_begin + JT;                       // first output slot on the work line
            const unsigned long long wo_e = (unsigned long long)(t + 1) * JT + off_t1;
            const unsigned long long a_w = wo_s > ws ? wo_s : ws;
            const unsigned long long b_w = wo_e < we ? wo_e : we;
            if (a_w >= b_w) continue;
            const unsigned long long a = a_w - wo_s, b = b_w - wo_s;           // slice of the tile's local outputs

            const bool class_b = t >= tiles_a;
            const JoinSide& X = class_b ? R : L;
            const JoinSide& Y = class_b ? L : R;
            const int64_t base = (class_b ? t - tiles_a : t) * JT;
            // recompute the tile's runs and their local exclusive offsets
            unsigned long long carry = 0;
#pragma unroll
            for (int j = 0; j < J_IPT; j++) {
                const int li = j * JB + threadIdx.x;
                const int64_t i = base + li;
                int64_t lo = 0, hi = 0;
                int32_t rid = 0;
                if (i < X.n) {
                    uint32_t xgs = X.gs[i], xge = X.ge[i];
                    if (class_b) item_run<true>(xgs, xge, Y.gs, Y.n, lo, hi);
                    else item_run<false>(xgs, xge, Y.gs, Y.n, lo, hi);
                    rid = X.rid[i];
                }
                s_lo[li] = (uint32_t)lo;
                s_rid[li] = rid;
                unsigned long long tot;
                unsigned long long ex = block_excl_scan_add<unsigned long long>((unsigned long long)(hi - lo), s_scratch, &tot);
                s_off[li] = carry + ex;
                carry += tot;
                __syncthreads();
            }
            if (threadIdx.x == 0) s_off[JT] = carry;
            __syncthreads();
            // per-slot search: slot s belongs to the last item whose offset is <= s
            for (unsigned long long s = a + threadIdx.x; s < b; s += JB) {
                int lo = 0, hi = JT;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (s_off[mid] <= s) lo = mid + 1; else hi = mid;
                }
                const int item = lo - 1;
                const int64_t ypos = (int64_t)s_lo[item] + (int64_t)(s - s_off[item]);
                const int32_t xr = s_rid[item];
                const int32_t yr = Y.rid[ypos];
                const unsigned long long o = off_t + s;
                out_l[o] = class_b ? yr : xr;
                out_r[o] = class_b ? xr : yr;
            }
            __syncthreads();   // shared arrays are rewritten by the next tile
        }
    }
}

// ---- general path (degenerate rows present) --------------------------------------------------------
__device__ __forceinline__ void general_candidates(uint32_t lgs, uint32_t lge, const uint32_t* __restrict__ rgs,
                                                   const uint32_t* __restrict__ pmax, int64_t nr, int64_t& lo, int64_t& hi) {
    hi = lower_bound_dev<uint32_t, int64_t>(rgs, 0, nr, lge);      // r.gs < l.ge
    lo = upper_bound_dev<uint32_t, int64_t>(pmax, 0, hi, lgs);     // first j with running max(r.ge) > l.gs
}

__global__ void __launch_bounds__(JB)
join_count_tiles_general_kernel(JoinSide L, JoinSide R, const uint32_t* __restrict__ pmax,
                                unsigned long long* __restrict__ tile_cnt) {
    __shared__ unsigned long long s_red[JB / 32];
    const int64_t base = (int64_t)blockIdx.x * JT;
    unsigned long long cnt = 0;
#pragma unroll
    for (int j = 0; j < J_IPT; j++) {
        int64_t i = base + j * JB + threadIdx.x;
        if (i < L.n) {
            uint32_t lgs = L.gs[i], lge = L.ge[i];
            int64_t lo, hi;
            general_candidates(lgs, lge, R.gs, pmax, R.n, lo, hi);
            for (int64_t k = lo; k < hi; k++) cnt += __ldg(R.ge + k) > lgs;
        }
    }
    cnt = warp_reduce_add(cnt);
    if (lane_id() == 0) s_red[warp_id()] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < JB / 32; w++) t += s_red[w];
        tile_cnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(JB)
join_emit_general_kernel(JoinSide L, JoinSide R, const uint32_t* __restrict__ pmax, int64_t n_tiles,
                         const unsigned long long* __restrict__ tile_off, int32_t* __restrict__ out_l,
                         int32_t* __restrict__ out_r, int64_t capacity) {
    __shared__ unsigned long long s_scratch[JB / 32 + 2];
    if ((long long)tile_off[n_tiles] > capacity) return;
    const int64_t base = (int64_t)blockIdx.x * JT;
    unsigned long long carry = tile_off[blockIdx.x];
#pragma unroll 1
    for (int j = 0; j < J_IPT; j++) {
        int64_t i = base + j * JB + threadIdx.x;
        unsigned long long cnt = 0;
        int64_t lo = 0, hi = 0;
        uint32_t lgs = 0;
        if (i < L.n) {
            lgs = L.gs[i];
            general_candidates(lgs, L.ge[i], R.gs, pmax, R.n, lo, hi);
            for (int64_t k = lo; k < hi; k++) cnt += __ldg(R.ge + k) > lgs;
        }
        unsigned long long tot;
        unsigned long long o = carry + block_excl_scan_add<unsigned long long>(cnt, s_scratch, &tot);
        carry += tot;
        if (cnt) {
            int32_t lr = L.rid[i];
            for (int64_t k = lo; k < hi; k++) {
                if (__ldg(R.ge + k) > lgs) {
                    out_l[o] = lr;
                    out_r[o] = R.rid[k];
                    o++;
                }
            }
        }
        __syncthreads();
    }
}

// ---- SEMI / ANTI and per-row counts, probing from the raw (unsorted) left columns --------------------
__global__ void __launch_bounds__(256)
probe_exists_kernel(ProbeCols P, const uint32_t* __restrict__ rgs, const uint32_t* __restrict__ pmax, int64_t nr,
                    uint8_t* __restrict__ exists) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t gs, ge;
        uint8_t e = 0;
        if (nr > 0 && probe_linearise(P, i, gs, ge)) {
            int64_t ub = lower_bound_dev<uint32_t, int64_t>(rgs, 0, nr, ge);   // rows with r.gs < l.ge
            e = ub > 0 && __ldg(pmax + ub - 1) > gs;                           // ... one of which ends after l.gs
        }
        exists[i] = e;
    }
}

__global__ void __launch_bounds__(256)
probe_count_kernel(ProbeCols P, const uint32_t* __restrict__ rgs, const uint32_t* __restrict__ rge,
                   const uint32_t* __restrict__ pmax, const uint32_t* __restrict__ rge_sorted, int64_t nr,
                   int r_has_degenerate, long long* __restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t gs, ge;
        long long c = 0;
        if (nr > 0 && probe_linearise(P, i, gs, ge)) {
            int64_t ub = lower_bound_dev<uint32_t, int64_t>(rgs, 0, nr, ge);
            if (!r_has_degenerate && ge >= gs) {
                // #{r.gs < l.ge} - #{r.ge <= l.gs}: every r ending at or before l.gs also starts before l.ge
                c = ub - upper_bound_dev<uint32_t, int64_t>(rge_sorted, 0, nr, gs);
            } else {
                int64_t lb = upper_bound_dev<uint32_t, int64_t>(pmax, 0, ub, gs);
                for (int64_t k = lb; k < ub; k++) c += __ldg(rge + k) > gs;
            }
        }
        counts[i] = c;
    }
}

inline JoinSide side_of(const giqlb_index* ix) { return JoinSide{ix->gs, ix->ge, ix->rid, ix->n}; }

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------------------
extern "C" int giqlb_joinplan_free(giqlb_ctx* ctx, giqlb_joinplan* plan) {
    if (!plan) return GIQLB_OK;
    if (ctx) dev_free(ctx, plan->tile_off);
    delete plan;
    return GIQLB_OK;
}

static int plan_enqueue(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, giqlb_joinplan** out) {
    if (L->layout != R->layout) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects: indexes were built on different layouts");
    if (L->grouped_by_strand || R->grouped_by_strand)
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects: strand-grouped indexes are not accepted");
    giqlb_joinplan* p = new (std::nothrow) giqlb_joinplan();
    if (!p) return GIQLB_ERR_NOMEM;
    p->L = L;
    p->R = R;
    p->general = (L->n_degenerate + R->n_degenerate) > 0;
    p->tiles_a = ceil_div(L->n, JT);
    p->tiles_b = p->general ? 0 : ceil_div(R->n, JT);
    const int64_t n_tiles = p->tiles_a + p->tiles_b;
    unsigned long long* cnt = nullptr;
    int st = dev_alloc(ctx, &cnt, (size_t)n_tiles + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &p->tile_off, (size_t)n_tiles + 1);
    if (st != GIQLB_OK) {
        dev_free(ctx, cnt);
        giqlb_joinplan_free(ctx, p);
        return st;
    }
    if (n_tiles > 0) {
        if (p->general) {
            st = giqlb_index_ensure_pmax(ctx, R);
            if (st == GIQLB_OK) {
                PhaseSpan ps(ctx, GIQLB_PH_JOIN_COUNT);
                join_count_tiles_general_kernel<<<(unsigned)n_tiles, JB, 0, ctx->stream>>>(side_of(L), side_of(R), R->pmax, cnt);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
        } else {
            PhaseSpan ps(ctx, GIQLB_PH_JOIN_COUNT);
            join_count_tiles_kernel<<<(unsigned)n_tiles, JB, 0, ctx->stream>>>(side_of(L), side_of(R), p->tiles_a, cnt);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
        }
        if (st == GIQLB_OK && cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "intersects: count launch failed");
    }
    if (st == GIQLB_OK) st = giqlb_excl_scan_u64(ctx, cnt, p->tile_off, n_tiles, reinterpret_cast<unsigned long long*>(ctx->d_scalars));
    dev_free(ctx, cnt);
    if (st != GIQLB_OK) {
        giqlb_joinplan_free(ctx, p);
        return st;
    }
    *out = p;
    return GIQLB_OK;
}

static int emit_enqueue(giqlb_ctx* ctx, giqlb_joinplan* p, int32_t* out_l, int32_t* out_r, int64_t capacity) {
    const int64_t n_tiles = p->tiles_a + p->tiles_b;
    if (n_tiles == 0) return GIQLB_OK;
    PhaseSpan ps(ctx, GIQLB_PH_JOIN_EMIT);
    if (p->general) {
        join_emit_general_kernel<<<(unsigned)n_tiles, JB, 0, ctx->stream>>>(side_of(p->L), side_of(p->R), p->R->pmax, n_tiles,
                                                                            p->tile_off, out_l, out_r, capacity);
    } else {
        int grid = ctx->sm_count * 8;
        join_emit_kernel<<<grid, JB, 0, ctx->stream>>>(side_of(p->L), side_of(p->R), p->tiles_a, n_tiles, p->tile_off, out_l,
                                                       out_r, capacity);
    }
    LAUNCH_CHECK(ctx);
    return GIQLB_OK;
}

static void join_stats(giqlb_ctx* ctx, const giqlb_index* L, const giqlb_index* R, int64_t total) {
    ctx->stats.rows_in[0] = L->n;
    ctx->stats.rows_in[1] = R->n;
    ctx->stats.rows_out = total;
    ctx->stats.algorithmic_bytes = 12 * (L->n + R->n) + 8 * total;
}

extern "C" int giqlb_intersects_plan(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, giqlb_joinplan** out,
                                     int64_t* total_host) {
    if (!ctx || !L || !R || !out) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_plan: bad argument");
    begin_call(ctx);
    *out = nullptr;
    giqlb_joinplan* p = nullptr;
    GIQLB_TRY(plan_enqueue(ctx, L, R, &p));
    int st = fetch_scalars(ctx, ctx->d_scalars, 1);
    if (st != GIQLB_OK) {
        giqlb_joinplan_free(ctx, p);
        return st;
    }
    p->total = ctx->h_scalars[0];
    if (total_host) *total_host = p->total;
    join_stats(ctx, L, R, p->total);
    *out = p;
    return end_call(ctx);
}

extern "C" int giqlb_intersects_emit(giqlb_ctx* ctx, giqlb_joinplan* plan, int32_t* out_l, int32_t* out_r) {
    if (!ctx || !plan) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_emit: bad argument");
    if (plan->total > 0 && (!out_l || !out_r)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_emit: null output");
    begin_call(ctx);
    if (plan->total > 0) GIQLB_TRY(emit_enqueue(ctx, plan, out_l, out_r, plan->total));
    join_stats(ctx, plan->L, plan->R, plan->total);
    return end_call(ctx);
}

extern "C" int giqlb_intersects_join(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, int32_t* out_l, int32_t* out_r,
                                     int64_t capacity, int64_t* total_host) {
    if (!ctx || !L || !R || capacity < 0 || (capacity > 0 && (!out_l || !out_r)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_join: bad argument");
    begin_call(ctx);
    giqlb_joinplan* p = nullptr;
    GIQLB_TRY(plan_enqueue(ctx, L, R, &p));
    // the emit kernel reads the total on the device and writes nothing when it exceeds `capacity`
    int st = capacity > 0 ? emit_enqueue(ctx, p, out_l, out_r, capacity) : GIQLB_OK;
    if (st == GIQLB_OK) st = fetch_scalars(ctx, ctx->d_scalars, 1);
    int64_t total = st == GIQLB_OK ? ctx->h_scalars[0] : -1;
    giqlb_joinplan_free(ctx, p);
    if (st != GIQLB_OK) return st;
    if (total_host) *total_host = total;
    join_stats(ctx, L, R, total);
    end_call(ctx);
    if (total > capacity) return giqlb_fail(ctx, GIQLB_ERR_CAPACITY, "intersects_join: %lld pairs exceed capacity %lld", (long long)total, (long long)capacity);
    return GIQLB_OK;
}

extern "C" int giqlb_intersects_exists(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                       const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, giqlb_index* R,
                                       uint8_t* exists_by_row) {
    if (!ctx || !R || (n > 0 && !exists_by_row)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_exists: bad argument");
    if (R->layout != lay) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_exists: index was built on a different layout");
    if (R->grouped_by_strand) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_exists: strand-grouped index not accepted");
    begin_call(ctx);
    ProbeCols P;
    GIQLB_TRY(probe_cols(ctx, lay, code, start, end, n, start_adj, end_adj, &P));
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, R));
    if (n > 0) {
        int64_t blocks = ceil_div(n, 256 * 4);
        int64_t cap = (int64_t)ctx->sm_count * 16;
        PhaseSpan ps(ctx, GIQLB_PH_PROBE);
        probe_exists_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(P, R->gs, R->pmax, R->n, exists_by_row);
        LAUNCH_CHECK(ctx);
    }
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = R->n;
    ctx->stats.rows_out = n;
    ctx->stats.algorithmic_bytes = 12 * (n + R->n) + n;
    return end_call(ctx);
}

extern "C" int giqlb_intersects_count_rows(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                           const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, giqlb_index* R,
                                           int64_t* counts_by_row) {
    if (!ctx || !R || (n > 0 && !counts_by_row)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_count_rows: bad argument");
    if (R->layout != lay) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_count_rows: index was built on a different layout");
    if (R->grouped_by_strand) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_count_rows: strand-grouped index not accepted");
    begin_call(ctx);
    ProbeCols P;
    GIQLB_TRY(probe_cols(ctx, lay, code, start, end, n, start_adj, end_adj, &P));
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, R));
    const int r_deg = R->n_degenerate > 0;
    if (!r_deg) GIQLB_TRY(giqlb_index_ensure_end_sorted(ctx, R));
    if (n > 0) {
        int64_t blocks = ceil_div(n, 256 * 4);
        int64_t cap = (int64_t)ctx->sm_count * 16;
        PhaseSpan ps(ctx, GIQLB_PH_PROBE);
        probe_count_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
            P, R->gs, R->ge, R->pmax, R->ge_sorted, R->n, r_deg, (long long*)counts_by_row);
        LAUNCH_CHECK(ctx);
    }
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = R->n;
    ctx->stats.rows_out = n;
    ctx->stats.algorithmic_bytes = 12 * (n + R->n) + 8 * n;
    return end_call(ctx);
}
