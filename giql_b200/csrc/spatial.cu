// spatial.cu — column-to-column CONTAINS and WITHIN joins.
//
// Semantics (reference, abdenlab/giql src/giql/expanders/intersects.py:155-166):
//   CONTAINS   l.chrom = r.chrom AND l.start <= r.start AND l.end >= r.end
//   WITHIN     l.chrom = r.chrom AND l.start >= r.start AND l.end <= r.end        ( = CONTAINS with the roles swapped )
// The reference has no IEJoin override for these (they always run the naive plan).  Here: for every container row x
// the candidates are the rows of the other table whose start lies in [x.start, x.end] — one contiguous run of the
// start-sorted index — filtered by end <= x.end; count -> exclusive scan -> emit, one thread per container row.
#include "common.cuh"

namespace {

constexpr int SPB = 256;
constexpr int SP_IPT = 4;
constexpr int SPT = SPB * SP_IPT;

struct Side {
    const uint32_t* gs;
    const uint32_t* ge;
    const int32_t* rid;
    int64_t n;
};

// candidate run of container x in Y: rows with x.gs <= y.gs <= x.ge.  With negative-width rows in Y (end < start) a
// contained row may start beyond x.ge, so the run extends to the end of the chromosome's cell range.
__device__ __forceinline__ void candidates(uint32_t xgs, uint32_t xge, const Side& Y, const uint32_t* __restrict__ base, int32_t n_codes,
                                           int y_has_negative, int64_t& lo, int64_t& hi) {
    lo = lower_bound_dev<uint32_t, int64_t>(Y.gs, 0, Y.n, xgs);
    uint32_t top = xge;
    if (y_has_negative) {   // last cell of x's chromosome: first base above xgs, minus one
        int32_t a = 0, b = n_codes;
        while (b - a > 1) {
            int32_t mid = (a + b) >> 1;
            if (__ldg(base + mid) <= xgs) a = mid; else b = mid;
        }
        top = __ldg(base + a + 1) - 1u;
    }
    hi = upper_bound_dev<uint32_t, int64_t>(Y.gs, lo, Y.n, top);
}

// A container row with more than SP_HEAVY candidates (a chromosome-spanning interval) is not walked by its own thread:
// the lanes of its warp take its candidate run together, 32 rows per step, so one long row costs the warp a few thousand
// steps instead of serialising ~10^5 dependent loads on one thread while the rest of the grid has finished.
constexpr int64_t SP_HEAVY = 128;

__device__ __forceinline__ unsigned long long warp_count_contained(const uint32_t* __restrict__ yge, int64_t lo, int64_t hi, uint32_t xge) {
    unsigned long long c = 0;
#pragma unroll 8
    for (int64_t k = lo + lane_id(); k < hi; k += 32) c += __ldg(yge + k) <= xge;   // independent loads: eight in flight per lane
    return warp_reduce_add(c);
}

// Containers with more than SP_SUPER candidates (a chromosome-spanning interval: ~10^5) are the whole CTA's job: a warp alone
// walks them at one 128-row step per L2 round trip (1.1 ms for one such row among 1 M, three walks), 256 threads with eight
// loads each in flight do it in a tenth of that.  Up to SP_SUPER_SLOTS of them per 256-row step of the tile; more take the warp path.
constexpr int64_t SP_SUPER = 4096;
constexpr int SP_SUPER_SLOTS = 8;
struct SuperRows {
    int n;
    int64_t lo[SP_SUPER_SLOTS], hi[SP_SUPER_SLOTS];
    uint32_t xge[SP_SUPER_SLOTS];
    int32_t xr[SP_SUPER_SLOTS];
    unsigned long long cnt[SP_SUPER_SLOTS], out[SP_SUPER_SLOTS], cursor[SP_SUPER_SLOTS];
};

// per-thread count of row i's contained candidates; heavy runs are counted by the whole warp, super-heavy ones by the whole
// CTA (ALL threads of the CTA must call, with the same `sr`).  *slot = the row's entry in sr (-1: none).
__device__ __forceinline__ unsigned long long count_row(const Side& Y, bool live, int64_t lo, int64_t hi, uint32_t xge, SuperRows& sr, int* slot) {
    if (threadIdx.x == 0) sr.n = 0;
    __syncthreads();
    *slot = -1;
    if (live && hi - lo > SP_SUPER) {
        const int e = atomicAdd(&sr.n, 1);
        if (e < SP_SUPER_SLOTS) {
            sr.lo[e] = lo; sr.hi[e] = hi; sr.xge[e] = xge; sr.cnt[e] = 0ull; sr.cursor[e] = 0ull;
            *slot = e;
        }
    }
    const bool heavy = live && hi - lo > SP_HEAVY && *slot < 0;
    unsigned long long cnt = 0;
    if (live && hi - lo <= SP_HEAVY)
        for (int64_t k = lo; k < hi; k++) cnt += __ldg(Y.ge + k) <= xge;
    uint32_t todo = __ballot_sync(FULL_MASK, heavy);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned long long c = warp_count_contained(Y.ge, __shfl_sync(FULL_MASK, lo, src), __shfl_sync(FULL_MASK, hi, src),
                                                          __shfl_sync(FULL_MASK, xge, src));
        if ((int)lane_id() == src) cnt = c;
    }
    __syncthreads();
    const int n_super = min(sr.n, SP_SUPER_SLOTS);
    for (int e = 0; e < n_super; e++) {
        const int64_t blo = sr.lo[e], bhi = sr.hi[e];
        const uint32_t bx = sr.xge[e];
        unsigned long long c = 0;
#pragma unroll 8
        for (int64_t k = blo + threadIdx.x; k < bhi; k += SPB) c += __ldg(Y.ge + k) <= bx;
        c = warp_reduce_add(c);
        if (lane_id() == 0 && c) atomicAdd(&sr.cnt[e], c);
    }
    __syncthreads();
    if (*slot >= 0) cnt = sr.cnt[*slot];
    return cnt;
}

__global__ void __launch_bounds__(SPB)
contains_count_kernel(Side X, Side Y, const uint32_t* __restrict__ base, int32_t n_codes, int y_has_negative,
                      unsigned long long* __restrict__ tile_cnt) {
    __shared__ unsigned long long s_red[SPB / 32];
    __shared__ SuperRows s_super;
    unsigned long long cnt = 0;
    const int64_t t0 = (int64_t)blockIdx.x * SPT;
#pragma unroll 1
    for (int j = 0; j < SP_IPT; j++) {
        const int64_t i = t0 + j * SPB + threadIdx.x;
        const bool live = i < X.n;
        int64_t lo = 0, hi = 0;
        uint32_t xge = 0;
        if (live) {
            xge = X.ge[i];
            candidates(X.gs[i], xge, Y, base, n_codes, y_has_negative, lo, hi);
        }
        int slot;
        cnt += count_row(Y, live, lo, hi, xge, s_super, &slot);
    }
    cnt = warp_reduce_add(cnt);
    if (lane_id() == 0) s_red[warp_id()] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < SPB / 32; w++) t += s_red[w];
        tile_cnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(SPB)
contains_emit_kernel(Side X, Side Y, const uint32_t* __restrict__ base, int32_t n_codes, int y_has_negative,
                     const unsigned long long* __restrict__ tile_off, int32_t* __restrict__ out_x, int32_t* __restrict__ out_y) {
    __shared__ unsigned long long s_scratch[SPB / 32 + 2];
    __shared__ SuperRows s_super;
    const int64_t t0 = (int64_t)blockIdx.x * SPT;
    unsigned long long carry = tile_off[blockIdx.x];
    const uint32_t lt_mask = (1u << lane_id()) - 1u;
#pragma unroll 1
    for (int j = 0; j < SP_IPT; j++) {
        const int64_t i = t0 + j * SPB + threadIdx.x;
        const bool live = i < X.n;
        int64_t lo = 0, hi = 0;
        uint32_t xge = 0;
        if (live) {
            xge = X.ge[i];
            candidates(X.gs[i], xge, Y, base, n_codes, y_has_negative, lo, hi);
        }
        int slot;
        const unsigned long long cnt = count_row(Y, live, lo, hi, xge, s_super, &slot);
        unsigned long long tot;
        unsigned long long o = carry + block_excl_scan_add<unsigned long long>(cnt, s_scratch, &tot);
        carry += tot;
        const bool heavy = live && hi - lo > SP_HEAVY && slot < 0;
        const int32_t xr = live ? X.rid[i] : 0;
        if (slot >= 0) {
            s_super.out[slot] = o;
            s_super.xr[slot] = xr;
        }
        if (cnt && live && hi - lo <= SP_HEAVY) {
            for (int64_t k = lo; k < hi; k++)
                if (__ldg(Y.ge + k) <= xge) {
                    out_x[o] = xr;
                    out_y[o] = Y.rid[k];
                    o++;
                }
        }
        uint32_t todo = __ballot_sync(FULL_MASK, heavy && cnt);
        while (todo) {                               // heavy rows: the warp emits the run together, in candidate order
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t blo = __shfl_sync(FULL_MASK, lo, src), bhi = __shfl_sync(FULL_MASK, hi, src);
            const uint32_t bxge = __shfl_sync(FULL_MASK, xge, src);
            const int32_t bxr = __shfl_sync(FULL_MASK, xr, src);
            unsigned long long bo = __shfl_sync(FULL_MASK, o, src);
            for (int64_t k0 = blo; k0 < bhi; k0 += 128) {     // four 32-row steps per iteration: their loads are in flight together
                bool keep[4];
                int32_t yr[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int64_t k = k0 + 32 * u + lane_id();
                    keep[u] = k < bhi && __ldg(Y.ge + k) <= bxge;
                    yr[u] = k < bhi ? __ldg(Y.rid + k) : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t m = __ballot_sync(FULL_MASK, keep[u]);
                    if (keep[u]) {
                        const unsigned long long pos = bo + __popc(m & lt_mask);
                        out_x[pos] = bxr;
                        out_y[pos] = yr[u];
                    }
                    bo += __popc(m);
                }
            }
        }
        __syncthreads();
        // super-heavy rows: the whole CTA emits the run; slots are claimed per warp from a shared cursor, so the order of a
        // row's pairs is not the candidate order (pair order is unspecified anyway)
        const int n_super = min(s_super.n, SP_SUPER_SLOTS);
        for (int e = 0; e < n_super; e++) {
            const int64_t blo = s_super.lo[e], bhi = s_super.hi[e];
            const uint32_t bx = s_super.xge[e];
            const int32_t bxr = s_super.xr[e];
            const unsigned long long bo = s_super.out[e];
            const int64_t span = ((bhi - blo + 4 * SPB - 1) / (4 * SPB)) * (4 * SPB);     // whole warps stay in the loop together
            for (int64_t k0 = 0; k0 < span; k0 += 4 * SPB) {
                bool keep[4];
                int32_t yr[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int64_t k = blo + k0 + u * SPB + threadIdx.x;
                    keep[u] = k < bhi && __ldg(Y.ge + k) <= bx;
                    yr[u] = k < bhi ? __ldg(Y.rid + k) : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t m = __ballot_sync(FULL_MASK, keep[u]);
                    unsigned long long wbase = 0;
                    if (lane_id() == 0 && m) wbase = atomicAdd(&s_super.cursor[e], (unsigned long long)__popc(m));
                    wbase = __shfl_sync(FULL_MASK, wbase, 0);
                    if (keep[u]) {
                        const unsigned long long pos = bo + wbase + __popc(m & lt_mask);
                        out_x[pos] = bxr;
                        out_y[pos] = yr[u];
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int giqlb_spatial_join(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, int op, int32_t* out_l, int32_t* out_r,
                                  int64_t capacity, int64_t* total_host) {
    if (!ctx || !L || !R || !total_host || capacity < 0 || (op != GIQLB_OP_CONTAINS && op != GIQLB_OP_WITHIN))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "spatial_join: bad argument");
    if (L->layout != R->layout) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "spatial_join: indexes were built on different layouts");
    GIQLB_TRY(require_row_ids(ctx, L, "spatial_join"));
    GIQLB_TRY(require_row_ids(ctx, R, "spatial_join"));
    if (L->grouped_by_strand || R->grouped_by_strand) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "spatial_join: strand-grouped indexes are not accepted");
    begin_call(ctx);
    GIQLB_TRY(giqlb_index_resolve(ctx, L, R));
    *total_host = 0;
    // WITHIN(l, r) == CONTAINS(r, l): X is always the container
    giqlb_index* Xi = op == GIQLB_OP_CONTAINS ? L : R;
    giqlb_index* Yi = op == GIQLB_OP_CONTAINS ? R : L;
    int32_t* out_x = op == GIQLB_OP_CONTAINS ? out_l : out_r;
    int32_t* out_y = op == GIQLB_OP_CONTAINS ? out_r : out_l;
    if (Xi->n == 0 || Yi->n == 0) return end_call(ctx);
    const Side X{Xi->gs, Xi->ge, Xi->rid, Xi->n}, Y{Yi->gs, Yi->ge, Yi->rid, Yi->n};
    const int64_t n_tiles = ceil_div(X.n, SPT);
    unsigned long long *cnt = nullptr, *off = nullptr;
    unsigned long long* total = reinterpret_cast<unsigned long long*>(ctx->d_scalars);
    int st = dev_alloc(ctx, &cnt, (size_t)n_tiles + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &off, (size_t)n_tiles + 1);
    const giqlb_layout* lay = L->layout;
    const int y_neg = Yi->n_negative > 0;
    if (st == GIQLB_OK) {
        PhaseSpan ps(ctx, GIQLB_PH_JOIN_COUNT);
        contains_count_kernel<<<(unsigned)n_tiles, SPB, 0, ctx->stream>>>(X, Y, lay->d_base, lay->n_codes, y_neg, cnt);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
    }
    if (st == GIQLB_OK) st = giqlb_excl_scan_u64(ctx, cnt, off, n_tiles, total);
    if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, total, 1);
    const int64_t n_out = st == GIQLB_OK ? ctx->h_scalars[0] : 0;
    if (st == GIQLB_OK) {
        *total_host = n_out;
        if (n_out > capacity) st = giqlb_fail(ctx, GIQLB_ERR_CAPACITY, "spatial_join: %lld pairs exceed capacity %lld", (long long)n_out, (long long)capacity);
        else if (n_out > 0 && (!out_l || !out_r)) st = giqlb_fail(ctx, GIQLB_ERR_INVALID, "spatial_join: null output");
    }
    if (st == GIQLB_OK && n_out > 0) {
        PhaseSpan ps(ctx, GIQLB_PH_JOIN_EMIT);
        contains_emit_kernel<<<(unsigned)n_tiles, SPB, 0, ctx->stream>>>(X, Y, lay->d_base, lay->n_codes, y_neg, off, out_x, out_y);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "spatial_join: launch failed");
    }
    dev_free(ctx, cnt);
    dev_free(ctx, off);
    if (st != GIQLB_OK) return st;
    ctx->stats.rows_in[0] = L->n;
    ctx->stats.rows_in[1] = R->n;
    ctx->stats.rows_out = n_out;
    ctx->stats.algorithmic_bytes = 12 * (L->n + R->n) + 8 * n_out;
    return end_call(ctx);
}
