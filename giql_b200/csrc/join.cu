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
// so counting needs two searches per row, no candidate is ever filtered out, and emit work is exactly the
// output size.  Both classes are processed by ONE sweep over the merge of the two start arrays ("L before R on
// ties"): a merge-path partition kernel cuts the merged sequence into tiles of JT rows; a tile holds L[x0, x1)
// and R[y0, y1), and the merge rank of every row is the START of its run in the other table — inside the
// tile's own window by construction.  A CTA stages both windows (plus J_EXT rows beyond each, which covers
// the END of almost every run) in shared memory with coalesced loads, so every input row is read from HBM
// once per pass and all searches run in shared memory; the rare run that leaves the window continues with a
// galloping search in global memory.
// Two entry points share that tile machinery:
//   giqlb_intersects_plan + _emit (exact-size, two passes)
//     count pass : every thread merges 4 rows of the tile; writes one item code (run start, class, length) per
//                  row and the tile's pair count
//     scan       : one CTA turns tile counts into pair offsets and emit work units; a tile whose output exceeds
//                  J_PART pairs is split into several units, so a skewed tile is shared by many CTAs
//     emit pass  : persistent CTAs loop over the units, scan the item codes and write their slice; each thread
//                  produces 8 CONSECUTIVE output pairs, stored as 16-byte vectors
//   giqlb_intersects_join (caller brings the buffer; ONE pass — see "single-pass join" below): count, look-back
//     offset and emit of a tile inside one kernel, TMA bulk staging, no item codes through HBM.
//
// General path (some row has end <= start): the literal predicate is scanned over the candidates
//     j in [first j with pmax[j] > l.gs,  lower_bound(R.gs, l.ge))        pmax = running max of R.ge
// which is exact for any rows, at the cost of visiting non-matching candidates.
#include <cuda/ptx>

#include "common.cuh"
#include "scan.cuh"

namespace {

#ifndef GIQLB_JOIN_THREADS
#define GIQLB_JOIN_THREADS 256
#endif
constexpr int JB = GIQLB_JOIN_THREADS;   // threads per CTA
#ifndef GIQLB_JOIN_IPT
#define GIQLB_JOIN_IPT 4
#endif
constexpr int JT = JB * GIQLB_JOIN_IPT;   // merged rows per tile (fast path) / left rows per tile (general path)
constexpr int J_IPT = JT / JB;
#ifndef GIQLB_JOIN_EXT
#define GIQLB_JOIN_EXT 512
#endif
constexpr int J_EXT = GIQLB_JOIN_EXT;   // rows staged beyond the tile's own part of each table
constexpr int JW = JT + J_EXT;      // staged window capacity per table
constexpr int J_PART = 8192;        // output pairs per emit work unit
constexpr int J_OPT = 8;            // consecutive outputs per thread in the emit loop
#ifndef GIQLB_JOIN_RUN_WINDOW
#define GIQLB_JOIN_RUN_WINDOW 32    // rows of the other window covered by the fixed-trip run-end search (longer runs: binary search)
#endif
constexpr int J_RUNW = GIQLB_JOIN_RUN_WINDOW;
#ifndef GIQLB_JOIN_LOOKBACK_WORDS
#define GIQLB_JOIN_LOOKBACK_WORDS 1 // predecessor tiles fetched per lane and look-back step (32 lanes: 128 tiles per L2 round trip)
#endif
constexpr int J_LB = GIQLB_JOIN_LOOKBACK_WORDS;
#ifndef GIQLB_JOIN_TICKET
#define GIQLB_JOIN_TICKET 0           // 1: tiles take tickets from an atomic counter (a tile then only waits on tiles already scheduled)
#endif
#ifndef GIQLB_JOIN_BACKOFF_NS
#define GIQLB_JOIN_BACKOFF_NS 40     // pause between polls of predecessor tiles that have not published yet
#endif

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

// shared-memory searches over a[lo, hi)
__device__ __forceinline__ int s_lower_bound(const uint32_t* a, int lo, int hi, uint32_t x) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first idx in [lo, hi) with a[idx] >= x, galloping from lo (the run end is usually a few rows away)
__device__ __forceinline__ int s_gallop(const uint32_t* a, int lo, int hi, uint32_t x) {
    if (lo >= hi || a[lo] >= x) return lo;
    int step = 1;                    // invariant: a[lo] < x
    while (lo + step < hi && a[lo + step] < x) {
        lo += step;
        step <<= 1;
    }
    int top = lo + step < hi ? lo + step : hi;
    return s_lower_bound(a, lo + 1, top, x);
}

// number of entries of a[0, n) that are <= x
__device__ __forceinline__ int s_upper_bound_u64(const unsigned long long* a, int n, unsigned long long x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- merge-path partition: part[t] = number of L rows among the first t*JT rows of the merge -----------
// recs (nullable) additionally receives the finished geometry of every tile — first rows, row counts and staged
// window lengths, eight int32 — so that the fused kernel starts a tile with ONE 32-byte load instead of two
// dependent loads and 64-bit min arithmetic on its critical path.
struct __align__(16) TileRec {
    int32_t x0, y0, nl, nr, wl, wr, pad0, pad1;
};
__device__ __forceinline__ int64_t merge_diagonal(const uint32_t* __restrict__ lgs, int64_t nl, const uint32_t* __restrict__ rgs, int64_t nr, int64_t d) {
    int64_t lo = d > nr ? d - nr : 0, hi = d < nl ? d : nl;
    while (lo < hi) {                      // L[mid] precedes R[d-1-mid] in the merge  <=>  more than mid L rows are taken
        int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(lgs + mid) <= __ldg(rgs + (d - 1 - mid))) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__global__ void __launch_bounds__(256)
join_partition_kernel(const uint32_t* __restrict__ lgs, int64_t nl, const uint32_t* __restrict__ rgs, int64_t nr,
                      int64_t n_tiles, int32_t* __restrict__ part, TileRec* __restrict__ recs,
                      const unsigned long long* __restrict__ unfinished_l = nullptr, const unsigned long long* __restrict__ unfinished_r = nullptr) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if ((unfinished_l && *unfinished_l) || (unfinished_r && *unfinished_r)) {
        // an index built with GIQLB_INDEX_SORT_SHORT_WINDOWS is not in order: nothing may be derived from it.  nl < 0 tells
        // the fused kernel to leave at once; the host sees the counter with the pair total and rebuilds.
        part[t] = 0;
        if (recs && t < n_tiles) {
            TileRec r;
            r.x0 = r.y0 = r.wl = r.wr = r.pad0 = r.pad1 = 0;
            r.nl = r.nr = -1;
            recs[t] = r;
        }
        return;
    }
    const int64_t d0 = min(t * JT, nl + nr);
    const int64_t x0 = merge_diagonal(lgs, nl, rgs, nr, d0);
    part[t] = (int32_t)x0;
    if (recs && t < n_tiles) {
        const int64_t d1 = min((t + 1) * (int64_t)JT, nl + nr);
        const int64_t x1 = merge_diagonal(lgs, nl, rgs, nr, d1);
        const int64_t y0 = d0 - x0, y1 = d1 - x1;
        TileRec r;
        r.x0 = (int32_t)x0;
        r.y0 = (int32_t)y0;
        r.nl = (int32_t)(x1 - x0);
        r.nr = (int32_t)(y1 - y0);
        r.wl = (int32_t)min((int64_t)r.nl + J_EXT, nl - x0);
        r.wr = (int32_t)min((int64_t)r.nr + J_EXT, nr - y0);
        r.pad0 = r.pad1 = 0;
        recs[t] = r;
    }
}

struct TileGeom {
    int64_t x0, y0;      // first L / R row of the tile
    int nl, nr;          // rows of each table in the tile (nl + nr <= JT)
    int wl, wr;          // staged window lengths (tile part + extension, clipped to the table)
};

__device__ __forceinline__ TileGeom tile_geom(const int32_t* __restrict__ part, int64_t t, int64_t NL, int64_t NR) {
    TileGeom g;
    const int64_t d0 = t * JT, d1 = min((t + 1) * (int64_t)JT, NL + NR);
    g.x0 = part[t];
    const int64_t x1 = part[t + 1];
    g.y0 = d0 - g.x0;
    const int64_t y1 = d1 - x1;
    g.nl = (int)(x1 - g.x0);
    g.nr = (int)(y1 - g.y0);
    g.wl = (int)min((int64_t)g.nl + J_EXT, NL - g.x0);
    g.wr = (int)min((int64_t)g.nr + J_EXT, NR - g.y0);
    return g;
}

__device__ __forceinline__ void stage_window(uint32_t* dst, const uint32_t* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += JB) dst[i] = __ldg(src + i);
}

// Item code written by the count pass and consumed by the emit pass (one uint32 per merged row):
//   bits  0..14  run start, as an offset into the other table's window of the tile (0..JT)
//   bit   15     class: 0 = the item is an L row (run of R rows), 1 = an R row (run of L rows)
//   bits 16..31  run length, saturated at 0xFFFF (the emit pass then re-derives the exact end)
constexpr uint32_t J_CLASS_B = 0x8000u;
constexpr uint32_t J_CNT_SAT = 0xFFFFu;

// ---- count pass -------------------------------------------------------------------------------------------
// Thread t merges rows [4t, 4t+4) of the tile's merge sequence: one merge-path search for its diagonal, then
// four merge steps.  The merge position of a row IS the start of its run; the end is a short gallop.
__global__ void __launch_bounds__(JB)
join_count_tiles_kernel(JoinSide L, JoinSide R, const int32_t* __restrict__ part, uint32_t* __restrict__ item_code,
                        unsigned long long* __restrict__ tile_cnt) {
    __shared__ uint32_t s_l[JW];
    __shared__ uint32_t s_r[JW];
    __shared__ uint32_t s_lge[JT];
    __shared__ uint32_t s_rge[JT];
    __shared__ unsigned long long s_red[JB / 32];
    const TileGeom g = tile_geom(part, blockIdx.x, L.n, R.n);
    stage_window(s_l, L.gs + g.x0, g.wl);
    stage_window(s_r, R.gs + g.y0, g.wr);
    stage_window(s_lge, L.ge + g.x0, g.nl);
    stage_window(s_rge, R.ge + g.y0, g.nr);
    __syncthreads();
    const int n_items = g.nl + g.nr;
    const int m0 = threadIdx.x * J_IPT;
    int i = g.nl, j = g.nr;
    if (m0 < n_items) {   // merge-path diagonal m0: i = L rows among the first m0 merged rows ("L before R on ties")
        int lo = m0 > g.nr ? m0 - g.nr : 0, hi = m0 < g.nl ? m0 : g.nl;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_l[mid] <= s_r[m0 - 1 - mid]) lo = mid + 1; else hi = mid;
        }
        i = lo;
        j = m0 - lo;
    }
    uint32_t code[J_IPT];
    unsigned long long cnt = 0;
#pragma unroll
    for (int k = 0; k < J_IPT; k++) {
        code[k] = 0;
        if (m0 + k < n_items) {
            const bool take_l = j >= g.nr || (i < g.nl && s_l[i] <= s_r[j]);
            int64_t c;
            if (take_l) {   // class A: run of R rows [y0 + j, hi)
                const uint32_t xge = s_lge[i];
                const int h = s_gallop(s_r, j, g.wr, xge);
                int64_t hi_g = g.y0 + h;
                if (h == g.wr && hi_g < R.n) hi_g = lb_gallop(R.gs, hi_g, R.n, xge);
                c = hi_g - (g.y0 + j);
                code[k] = (uint32_t)j;
                i++;
            } else {        // class B: run of L rows [x0 + i, hi)
                const uint32_t xge = s_rge[j];
                const int h = s_gallop(s_l, i, g.wl, xge);
                int64_t hi_g = g.x0 + h;
                if (h == g.wl && hi_g < L.n) hi_g = lb_gallop(L.gs, hi_g, L.n, xge);
                c = hi_g - (g.x0 + i);
                code[k] = (uint32_t)i | J_CLASS_B;
                j++;
            }
            code[k] |= (uint32_t)(c < (int64_t)J_CNT_SAT ? c : (int64_t)J_CNT_SAT) << 16;
            cnt += (unsigned long long)c;
        }
    }
    static_assert(J_IPT % 4 == 0, "item codes are stored as uint4s, J_IPT / 4 per thread");
#pragma unroll
    for (int q = 0; q < J_IPT / 4; q++)
        reinterpret_cast<uint4*>(item_code + (size_t)blockIdx.x * JT)[threadIdx.x * (J_IPT / 4) + q] =
            make_uint4(code[4 * q], code[4 * q + 1], code[4 * q + 2], code[4 * q + 3]);
    cnt = warp_reduce_add(cnt);
    if (lane_id() == 0) s_red[warp_id()] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < JB / 32; w++) t += s_red[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// ---- scan: pair offsets and emit work units of every tile (one CTA) -------------------------------------
__global__ void __launch_bounds__(1024)
join_scan_tiles_kernel(const unsigned long long* __restrict__ tile_cnt, int64_t n_tiles, unsigned long long* __restrict__ tile_off,
                       unsigned long long* __restrict__ unit_off, unsigned long long* __restrict__ totals /* [0] pairs [1] units */) {
    __shared__ unsigned long long scratch[1024 / 32 + 2];
    // every thread owns a contiguous chunk: serial sums, ONE block scan per quantity, serial write-back
    const int64_t per = (n_tiles + 1023) / 1024;
    const int64_t lo = min((int64_t)threadIdx.x * per, n_tiles), hi = min(lo + per, n_tiles);
    unsigned long long sp = 0, su = 0;
    for (int64_t i = lo; i < hi; i++) {
        const unsigned long long c = tile_cnt[i];
        sp += c;
        su += c > J_PART ? (c + J_PART - 1) / J_PART : 1ull;
    }
    unsigned long long tot_p, tot_u;
    unsigned long long ex_p = block_excl_scan_add<unsigned long long>(sp, scratch, &tot_p);
    __syncthreads();
    unsigned long long ex_u = block_excl_scan_add<unsigned long long>(su, scratch, &tot_u);
    for (int64_t i = lo; i < hi; i++) {
        const unsigned long long c = tile_cnt[i];
        tile_off[i] = ex_p;
        unit_off[i] = ex_u;
        ex_p += c;
        ex_u += c > J_PART ? (c + J_PART - 1) / J_PART : 1ull;
    }
    if (threadIdx.x == 0) {
        tile_off[n_tiles] = tot_p;
        unit_off[n_tiles] = tot_u;
        totals[0] = tot_p;
        totals[1] = tot_u;
    }
}

// ---- emit pass ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JB)
join_emit_kernel(JoinSide L, JoinSide R, const int32_t* __restrict__ part, int64_t n_tiles, const uint32_t* __restrict__ item_code,
                 const unsigned long long* __restrict__ tile_off, const unsigned long long* __restrict__ unit_off,
                 int32_t* __restrict__ out_l, int32_t* __restrict__ out_r, int64_t capacity) {
    __shared__ int32_t s_lrid[JW];
    __shared__ int32_t s_rrid[JW];
    __shared__ unsigned long long s_off[JT + 1];   // local exclusive output offsets of the tile's items
    __shared__ uint32_t s_code[JT];
    __shared__ unsigned long long s_scratch[JB / 32 + 2];

    const unsigned long long total = tile_off[n_tiles];
    if ((long long)total > capacity) return;       // caller's buffer too small: nothing is written
    const unsigned long long n_units = unit_off[n_tiles];
    const bool split = n_units != (unsigned long long)n_tiles;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(out_l) | reinterpret_cast<uintptr_t>(out_r)) & 15) == 0;

    for (unsigned long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int64_t t = (int64_t)unit;
        if (split) {   // last tile whose first unit is <= unit
            int64_t lo = 0, hi = n_tiles;
            while (hi - lo > 1) {
                int64_t mid = lo + ((hi - lo) >> 1);
                if (unit_off[mid] <= unit) lo = mid; else hi = mid;
            }
            t = lo;
        }
        const unsigned long long off_t = tile_off[t], cnt_t = tile_off[t + 1] - off_t;
        if (cnt_t == 0) continue;
        const unsigned long long a = (unit - unit_off[t]) * (unsigned long long)J_PART;        // this unit's slice of the tile's outputs
        const unsigned long long b = min(a + (unsigned long long)J_PART, cnt_t);

        const TileGeom g = tile_geom(part, t, L.n, R.n);
        for (int i = threadIdx.x; i < g.wl; i += JB) s_lrid[i] = __ldg(L.rid + g.x0 + i);
        for (int i = threadIdx.x; i < g.wr; i += JB) s_rrid[i] = __ldg(R.rid + g.y0 + i);
        // the thread's four items: run lengths from the count pass (exact end re-derived when saturated)
        uint32_t code[J_IPT];
#pragma unroll
        for (int q = 0; q < J_IPT / 4; q++) {
            const uint4 cv = reinterpret_cast<const uint4*>(item_code + (size_t)t * JT)[threadIdx.x * (J_IPT / 4) + q];
            code[4 * q] = cv.x; code[4 * q + 1] = cv.y; code[4 * q + 2] = cv.z; code[4 * q + 3] = cv.w;
        }
        unsigned long long c[J_IPT], sum = 0;
#pragma unroll
        for (int k = 0; k < J_IPT; k++) {
            c[k] = code[k] >> 16;
            if (c[k] == J_CNT_SAT) {
                const int m = threadIdx.x * J_IPT + k;
                const int lo_local = (int)(code[k] & 0x7FFFu);
                if (code[k] & J_CLASS_B) {
                    const uint32_t xge = __ldg(R.ge + g.y0 + (m - lo_local));
                    c[k] = (unsigned long long)(lb_gallop(L.gs, g.x0 + lo_local, L.n, xge) - (g.x0 + lo_local));
                } else {
                    const uint32_t xge = __ldg(L.ge + g.x0 + (m - lo_local));
                    c[k] = (unsigned long long)(lb_gallop(R.gs, g.y0 + lo_local, R.n, xge) - (g.y0 + lo_local));
                }
            }
            sum += c[k];
        }
        unsigned long long tot;
        unsigned long long ex = block_excl_scan_add<unsigned long long>(sum, s_scratch, &tot);
#pragma unroll
        for (int k = 0; k < J_IPT; k++) {
            s_off[threadIdx.x * J_IPT + k] = ex;
            s_code[threadIdx.x * J_IPT + k] = code[k];
            ex += c[k];
        }
        if (threadIdx.x == 0) s_off[JT] = tot;
        __syncthreads();

        // each thread owns J_OPT consecutive GLOBAL output positions, aligned to J_OPT
        const unsigned long long g_a = off_t + a, g_b = off_t + b;
        const unsigned long long g_first = g_a & ~(unsigned long long)(J_OPT - 1);
        for (unsigned long long gp = g_first + (unsigned long long)threadIdx.x * J_OPT; gp < g_b; gp += (unsigned long long)JB * J_OPT) {
            const unsigned long long p0 = gp > g_a ? gp : g_a, p1 = min(gp + J_OPT, g_b);
            if (p0 >= p1) continue;
            int item = s_upper_bound_u64(s_off, JT + 1, p0 - off_t) - 1;    // last item whose offset is <= the first slot
            int32_t vl[J_OPT], vr[J_OPT];
#pragma unroll
            for (int q = 0; q < J_OPT; q++) {
                const unsigned long long pq = gp + q;
                vl[q] = 0;
                vr[q] = 0;
                if (pq >= p0 && pq < p1) {
                    const unsigned long long sl = pq - off_t;          // local output slot
                    while (s_off[item + 1] <= sl) item++;              // skip items with empty runs
                    const uint32_t cd = s_code[item];
                    const int lo_local = (int)(cd & 0x7FFFu);
                    const int64_t w = (int64_t)lo_local + (int64_t)(sl - s_off[item]);   // offset into the other table's window
                    if (!(cd & J_CLASS_B)) {                           // x is L row (item - lo_local), y an R row
                        vl[q] = s_lrid[item - lo_local];
                        vr[q] = w < g.wr ? s_rrid[w] : __ldg(R.rid + g.y0 + w);
                    } else {                                           // x is R row (item - lo_local), y an L row
                        vr[q] = s_rrid[item - lo_local];
                        vl[q] = w < g.wl ? s_lrid[w] : __ldg(L.rid + g.x0 + w);
                    }
                }
            }
            if (vec_ok && p0 == gp && p1 == gp + J_OPT) {
                st_stream_u32x4(reinterpret_cast<uint4*>(out_l + gp), make_uint4(vl[0], vl[1], vl[2], vl[3]));
                st_stream_u32x4(reinterpret_cast<uint4*>(out_l + gp) + 1, make_uint4(vl[4], vl[5], vl[6], vl[7]));
                st_stream_u32x4(reinterpret_cast<uint4*>(out_r + gp), make_uint4(vr[0], vr[1], vr[2], vr[3]));
                st_stream_u32x4(reinterpret_cast<uint4*>(out_r + gp) + 1, make_uint4(vr[4], vr[5], vr[6], vr[7]));
            } else {
#pragma unroll
                for (int q = 0; q < J_OPT; q++) {
                    const unsigned long long pq = gp + q;
                    if (pq >= p0 && pq < p1) {
                        out_l[pq] = vl[q];
                        out_r[pq] = vr[q];
                    }
                }
            }
        }
        __syncthreads();   // shared arrays are rewritten by the next unit
    }
}

// ---- single-pass join (giqlb_intersects_join): count, offset and emit of a tile in ONE kernel ------------------
// When the caller brings an output buffer, the count and emit passes collapse: a CTA stages its tile with bulk
// (TMA) copies, computes the runs, publishes the tile's pair count and obtains its output offset through a
// decoupled look-back over per-tile status words (tiles take tickets, so a tile only waits on tiles already
// scheduled), then emits straight from shared memory.  Every input row is read once, no item codes travel
// through HBM, and the capacity check is per tile (a tile that would overrun the buffer writes nothing; the
// grand total tells the host how much room a retry needs).  Tiles whose output exceeds J_BIG pairs only
// publish their offset and are emitted afterwards by a second kernel that shares each of them among many CTAs.
#ifndef GIQLB_JOIN_BIG
#define GIQLB_JOIN_BIG (8 * JT)
#endif
constexpr int J_BIG = GIQLB_JOIN_BIG;  // tiles with more pairs are shared among CTAs by join_emit_big_kernel
constexpr unsigned long long JF_FLAG_AGG = 1ull << 62;
constexpr unsigned long long JF_FLAG_INCL = 2ull << 62;
constexpr unsigned long long JF_VALUE_MASK = (1ull << 62) - 1;
constexpr uint32_t JF_SPIN_LIMIT = 1u << 26;

constexpr int J_STAGE = ((2 * (JW + 8) + 2 * (JT + 8)) / 2) & ~7;   // output pairs staged per window (aliases gs + ge)

struct __align__(16) TileSmem {
    uint32_t gs[2][JW + 8];                  // staged start windows: [0] L, [1] R (16-byte aligned source => up to 3 lead-in rows)
    uint32_t ge[2][JT + 8];                  // ends of the tile's own rows.  gs + ge are dead once the runs are known: the
                                             // fused kernel reuses them as the output staging window (stage_l / stage_r)
    int32_t rid[2][JW + 8];                  // row ids of the windows
    unsigned long long scratch[JB / 32 + 2];
    unsigned long long tile_off;
    unsigned long long bar[2];               // mbarriers: [0] starts + ends, [1] row ids
    uint32_t tile;
};

struct TileSlots {                           // fused kernel (tiles of <= J_BIG pairs): slot -> item lookup over the NON-EMPTY items
    uint2 ent[JT + 2];                       // merge order: .x = local exclusive output offset (bits 0..14) | class (bit 15) |
                                             // (index of the run's first row in the row-id windows - offset) as int16 (bits 16..31),
                                             // .y = the item's OWN row id (looked up once per item, not once per pair); then a sentinel
    uint16_t first_item[J_BIG / J_OPT + 2];  // non-empty item that owns the first slot of every J_OPT-slot chunk
};
constexpr uint32_t JE_OFF_MASK = 0x7FFFu, JE_CLASS_B = 0x8000u;
static_assert(J_BIG < (int)JE_OFF_MASK, "a tile's local offsets must fit 15 bits");
// Output staging window: one (left, right) row-id pair per slot, 8 bytes.  A thread of the expansion owns J_OPT = 8 consecutive
// slots, so at step q the lanes of a half-warp (which share one 64-bit shared-memory transaction) would all hit the same two
// banks.  XOR-ing the slot's low three bits with its 16-block number spreads every step over all banks, and keeps the flush's
// consecutive-slot reads conflict-free (ncu, round 1: 48.8 M bank-conflict wavefronts on 29.9 M shared-memory instructions;
// the shared-memory pipe was the kernel's limiter at 85 %).  For a chunk-aligned slot s0 the swizzled index of s0 + q is
// stage_swz(s0) ^ q: one instruction per slot.
__device__ __forceinline__ uint32_t stage_swz(uint32_t s) { return s ^ ((s >> 4) & 7u); }
struct TileOffsets {                         // big-tile kernel: per-item offsets for the slot -> item search
    unsigned long long off[JT + 1];          // local exclusive output offsets of the tile's items (merge order)
    uint32_t code[JT];                       // run start | class of every item
};

__device__ __forceinline__ void bar_wait(unsigned long long* bar, uint32_t parity) {
    // the hint lets the hardware park the warp instead of spinning on the barrier while the bulk copies land
    while (!cuda::ptx::mbarrier_try_wait_parity(reinterpret_cast<uint64_t*>(bar), parity, 2000u)) {
    }
}

// Issues the bulk copy of elements [first, first + count) of a 4-byte array into dst; the source is rounded down
// to a 16-byte boundary (the library pads its arrays), so dst[(first & 3) + i] = src[first + i].
__device__ __forceinline__ uint32_t window_bytes(int64_t first, int count) {
    return (uint32_t)((((int)(first & 3) + count + 3) >> 2) << 4);
}
__device__ __forceinline__ void window_copy(void* dst, const void* src_base, int64_t first, int count, unsigned long long* bar) {
    if (count <= 0) return;
    const char* src = reinterpret_cast<const char*>(src_base) + ((first & ~(int64_t)3) << 2);
    cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, dst, src, window_bytes(first, count),
                             reinterpret_cast<uint64_t*>(bar));
}

// Thread 0 arms both barriers and issues the six window copies of tile g.
__device__ __forceinline__ void tile_stage(TileSmem& sm, const TileGeom& g, const JoinSide& L, const JoinSide& R) {
    uint32_t b0 = 0, b1 = 0;
    if (g.wl > 0) b0 += window_bytes(g.x0, g.wl), b1 += window_bytes(g.x0, g.wl);
    if (g.wr > 0) b0 += window_bytes(g.y0, g.wr), b1 += window_bytes(g.y0, g.wr);
    if (g.nl > 0) b0 += window_bytes(g.x0, g.nl);
    if (g.nr > 0) b0 += window_bytes(g.y0, g.nr);
    cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared,
                                         reinterpret_cast<uint64_t*>(&sm.bar[0]), b0);
    window_copy(sm.gs[0], L.gs, g.x0, g.wl, &sm.bar[0]);
    window_copy(sm.gs[1], R.gs, g.y0, g.wr, &sm.bar[0]);
    window_copy(sm.ge[0], L.ge, g.x0, g.nl, &sm.bar[0]);
    window_copy(sm.ge[1], R.ge, g.y0, g.nr, &sm.bar[0]);
    cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared,
                                         reinterpret_cast<uint64_t*>(&sm.bar[1]), b1);
    window_copy(sm.rid[0], L.rid, g.x0, g.wl, &sm.bar[1]);
    window_copy(sm.rid[1], R.rid, g.y0, g.wr, &sm.bar[1]);
}

// The same by the first six lanes of warp 0, one bulk copy each (the whole warp must call): lane 0 arms both barriers
// first, so the six copies are issued side by side instead of one after the other behind ~500 scalar instructions.
__device__ __forceinline__ void tile_stage_warp(TileSmem& sm, const TileGeom& g, const JoinSide& L, const JoinSide& R) {
    const uint32_t lane = lane_id();
    if (lane == 0) {
        uint32_t b0 = 0, b1 = 0;
        if (g.wl > 0) b0 += window_bytes(g.x0, g.wl), b1 += window_bytes(g.x0, g.wl);
        if (g.wr > 0) b0 += window_bytes(g.y0, g.wr), b1 += window_bytes(g.y0, g.wr);
        if (g.nl > 0) b0 += window_bytes(g.x0, g.nl);
        if (g.nr > 0) b0 += window_bytes(g.y0, g.nr);
        cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared,
                                             reinterpret_cast<uint64_t*>(&sm.bar[0]), b0);
        cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared,
                                             reinterpret_cast<uint64_t*>(&sm.bar[1]), b1);
    }
    __syncwarp();
    if (lane < 6) {
        const bool right = lane & 1;                                    // lanes 0, 2, 4: L windows; 1, 3, 5: R windows
        const int what = (int)lane >> 1;                                // 0 starts, 1 ends (own rows only), 2 row ids
        const int64_t first = right ? g.y0 : g.x0;
        const int count = what == 1 ? (right ? g.nr : g.nl) : (right ? g.wr : g.wl);
        const JoinSide& S = right ? R : L;
        const void* src = what == 0 ? (const void*)S.gs : what == 1 ? (const void*)S.ge : (const void*)S.rid;
        void* dst = what == 0 ? (void*)sm.gs[right] : what == 1 ? (void*)sm.ge[right] : (void*)sm.rid[right];
        window_copy(dst, src, first, count, &sm.bar[what == 2 ? 1 : 0]);
    }
}

// The thread's J_IPT consecutive rows of the tile's merge sequence: run start/class codes and exact run lengths.
// Branch-free in the class: the row's own table and the other table are selected by index.
__device__ __forceinline__ void tile_runs(const TileSmem& sm, const TileGeom& g, const JoinSide& L, const JoinSide& R,
                                          uint32_t (&code)[J_IPT], uint32_t (&cnt)[J_IPT], bool* leaves_window = nullptr) {
    const uint32_t* gsl = sm.gs[0] + (g.x0 & 3);
    const uint32_t* gsr = sm.gs[1] + (g.y0 & 3);
    const uint32_t* gel = sm.ge[0] + (g.x0 & 3);
    const uint32_t* ger = sm.ge[1] + (g.y0 & 3);
    const int n_items = g.nl + g.nr;
    const int m0 = threadIdx.x * J_IPT;
    int i = g.nl, j = g.nr;
    if (m0 < n_items) {   // merge-path diagonal m0: i = L rows among the first m0 merged rows ("L before R on ties")
        int lo = m0 > g.nr ? m0 - g.nr : 0, hi = m0 < g.nl ? m0 : g.nl;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (gsl[mid] <= gsr[m0 - 1 - mid]) lo = mid + 1; else hi = mid;
        }
        i = lo;
        j = m0 - lo;
    }
    // Three passes over the thread's rows so that the J_IPT run searches — chains of dependent shared-memory loads —
    // are independent straight-line code the scheduler can interleave; the rare long-run paths come last.
    bool is_l[J_IPT];
    int start[J_IPT], h[J_IPT], lim[J_IPT];
    uint32_t xge[J_IPT];
    // the heads of both windows ride in registers: one shared-memory load per merge step (the head that advanced), and one
    // for the row's own end.  Reads at the window end stay inside the padded arrays and are guarded by the index tests.
    uint32_t cl = gsl[i], cr = gsr[j];
#pragma unroll
    for (int k = 0; k < J_IPT; k++) {
        const bool live = m0 + k < n_items;
        const bool take_l = live && (j >= g.nr || (i < g.nl && cl <= cr));
        is_l[k] = take_l;
        const uint32_t* ep = take_l ? gel + i : ger + j;
        xge[k] = live ? *ep : 0u;                             // the row's own end (0: no row of the other table starts below it)
        start[k] = take_l ? j : i;
        lim[k] = !live ? start[k] : min(start[k] + J_RUNW, take_l ? g.wr : g.wl);
        h[k] = start[k];
        i += (live && take_l) ? 1 : 0;
        j += (live && !take_l) ? 1 : 0;
        if (k + 1 < J_IPT) {
            const uint32_t nv = *(take_l ? gsl + i : gsr + j);
            cl = take_l ? nv : cl;
            cr = take_l ? cr : nv;
        }
    }
    // run end = first row of the other window whose start is >= the row's end.  Almost every run is a few rows long:
    // a fixed-trip (divergence-free) binary lifting over the next 32 rows
#pragma unroll
    for (int step = J_RUNW / 2; step > 0; step >>= 1) {
#pragma unroll
        for (int k = 0; k < J_IPT; k++) {
            const uint32_t* ys = is_l[k] ? gsr : gsl;         // the other table's window
            if (h[k] + step <= lim[k] && ys[h[k] + step - 1] < xge[k]) h[k] += step;
        }
    }
#pragma unroll
    for (int k = 0; k < J_IPT; k++) {
        const uint32_t* ys = is_l[k] ? gsr : gsl;
        if (h[k] < lim[k] && ys[h[k]] < xge[k]) h[k]++;
    }
#pragma unroll
    for (int k = 0; k < J_IPT; k++) {
        code[k] = 0;
        cnt[k] = 0;
        if (m0 + k < n_items) {
            const bool take_l = is_l[k];
            const uint32_t* ys = take_l ? gsr : gsl;
            const int wy = take_l ? g.wr : g.wl;
            int hh = h[k];
            if (hh == lim[k] && lim[k] < wy && ys[hh] < xge[k]) hh = s_lower_bound(ys, hh + 1, wy, xge[k]);   // the rare long run
            int64_t c = hh - start[k];
            if (hh == wy) {                                  // the run may continue beyond the staged window
                const uint32_t* yg = take_l ? R.gs : L.gs;
                const int64_t yn = take_l ? R.n : L.n, ybase = take_l ? g.y0 : g.x0;
                if (ybase + wy < yn) {
                    c = lb_gallop(yg, ybase + wy, yn, xge[k]) - (ybase + start[k]);
                    if (leaves_window && c > wy - start[k]) *leaves_window = true;
                }
            }
            code[k] = (uint32_t)start[k] | (take_l ? 0u : J_CLASS_B);
            cnt[k] = (uint32_t)c;
        }
    }
}

// Block scan of the run lengths into sm.off / sm.code; returns the tile's pair count (same value in every thread).
__device__ __forceinline__ unsigned long long tile_offsets(TileSmem& sm, TileOffsets& so, const uint32_t (&code)[J_IPT], const uint32_t (&cnt)[J_IPT]) {
    unsigned long long sum = 0;
#pragma unroll
    for (int k = 0; k < J_IPT; k++) sum += cnt[k];
    unsigned long long tot;
    unsigned long long ex = block_excl_scan_add<unsigned long long>(sum, sm.scratch, &tot);
#pragma unroll
    for (int k = 0; k < J_IPT; k++) {
        so.off[threadIdx.x * J_IPT + k] = ex;
        so.code[threadIdx.x * J_IPT + k] = code[k];
        ex += cnt[k];
    }
    if (threadIdx.x == 0) so.off[JT] = tot;
    return tot;
}

// Emits the tile's local output slots [a, b) at global positions off_t + slot; every thread writes J_OPT consecutive
// pairs aligned to J_OPT (16-byte vector stores).
__device__ __forceinline__ void tile_emit(const TileSmem& sm, const TileOffsets& so, const TileGeom& g, const JoinSide& L, const JoinSide& R,
                                          unsigned long long off_t, unsigned long long a, unsigned long long b,
                                          int32_t* __restrict__ out_l, int32_t* __restrict__ out_r, bool vec_ok) {
    const int32_t* lrid = sm.rid[0] + (g.x0 & 3);
    const int32_t* rrid = sm.rid[1] + (g.y0 & 3);
    const unsigned long long g_a = off_t + a, g_b = off_t + b;
    const unsigned long long g_first = g_a & ~(unsigned long long)(J_OPT - 1);
    for (unsigned long long gp = g_first + (unsigned long long)threadIdx.x * J_OPT; gp < g_b; gp += (unsigned long long)JB * J_OPT) {
        const unsigned long long p0 = gp > g_a ? gp : g_a, p1 = min(gp + J_OPT, g_b);
        if (p0 >= p1) continue;
        int item = s_upper_bound_u64(so.off, JT + 1, p0 - off_t) - 1;    // last item whose offset is <= the first slot
        int32_t vl[J_OPT], vr[J_OPT];
#pragma unroll
        for (int q = 0; q < J_OPT; q++) {
            const unsigned long long pq = gp + q;
            vl[q] = 0;
            vr[q] = 0;
            if (pq >= p0 && pq < p1) {
                const unsigned long long sl = pq - off_t;
                while (so.off[item + 1] <= sl) item++;              // skip items with empty runs
                const uint32_t cd = so.code[item];
                const int start = (int)(cd & 0x7FFFu);
                const int64_t w = (int64_t)start + (int64_t)(sl - so.off[item]);   // offset into the other table's window
                if (!(cd & J_CLASS_B)) {                            // item is L row (item - start); its run lies in R
                    vl[q] = lrid[item - start];
                    vr[q] = w < g.wr ? rrid[w] : __ldg(R.rid + g.y0 + w);
                } else {                                            // item is R row (item - start); its run lies in L
                    vr[q] = rrid[item - start];
                    vl[q] = w < g.wl ? lrid[w] : __ldg(L.rid + g.x0 + w);
                }
            }
        }
        if (vec_ok && p0 == gp && p1 == gp + J_OPT) {
            st_stream_u32x4(reinterpret_cast<uint4*>(out_l + gp), make_uint4(vl[0], vl[1], vl[2], vl[3]));
            st_stream_u32x4(reinterpret_cast<uint4*>(out_l + gp) + 1, make_uint4(vl[4], vl[5], vl[6], vl[7]));
            st_stream_u32x4(reinterpret_cast<uint4*>(out_r + gp), make_uint4(vr[0], vr[1], vr[2], vr[3]));
            st_stream_u32x4(reinterpret_cast<uint4*>(out_r + gp) + 1, make_uint4(vr[4], vr[5], vr[6], vr[7]));
        } else {
#pragma unroll
            for (int q = 0; q < J_OPT; q++) {
                const unsigned long long pq = gp + q;
                if (pq >= p0 && pq < p1) {
                    out_l[pq] = vl[q];
                    out_r[pq] = vr[q];
                }
            }
        }
    }
}

// Item table of the tile over its NON-EMPTY items (compacted with the item count that rides in the high bits of the
// block scan): one packed entry per item — offset, class, where its run starts in the row-id windows, and the item's own
// row id, read here once per item so that the expansion loads one row id per pair instead of two — and for every chunk of
// J_OPT output slots the entry owning its first slot (written by that item itself, so the expansion below needs no search
// and its walk never meets an empty item).  The row-id windows must have landed.
__device__ __forceinline__ void tile_slot_table(const TileSmem& sm, TileSlots& ts, const TileGeom& g, const uint32_t (&code)[J_IPT],
                                                const uint32_t (&cnt)[J_IPT], uint32_t ex, uint32_t ex_items, uint32_t tot, uint32_t tot_items) {
    const int32_t* rid_all = &sm.rid[0][0];
    const int l_off = (int)(g.x0 & 3), r_off = (JW + 8) + (int)(g.y0 & 3);     // first row of each window inside rid_all
    uint32_t o = ex, ci = ex_items;
#pragma unroll
    for (int k = 0; k < J_IPT; k++) {
        if (cnt[k] == 0) continue;
        const int item = threadIdx.x * J_IPT + k;
        const int start = (int)(code[k] & 0x7FFFu);
        const bool class_b = (code[k] & J_CLASS_B) != 0;                       // item is an R row: its run lies in the L window
        const int own = (class_b ? r_off : l_off) + (item - start);            // the item's own row
        const int wrel = (class_b ? l_off : r_off) + start - (int)o;           // rid_all[wrel + slot] = the run's row for `slot`
        ts.ent[ci] = make_uint2(o | (class_b ? JE_CLASS_B : 0u) | ((uint32_t)wrel << 16), (uint32_t)rid_all[own]);
        for (uint32_t q = (o + J_OPT - 1) / J_OPT; q * J_OPT < o + cnt[k]; q++) ts.first_item[q] = (uint16_t)ci;
        o += cnt[k];
        ci++;
    }
    if (threadIdx.x == JB - 1) {
        ts.ent[tot_items] = make_uint2(tot, 0u);         // tot <= J_BIG
        ts.ent[tot_items + 1] = make_uint2(JE_OFF_MASK, 0u); // sentinel: the walk below never leaves the table
    }
}

// Expands output slots [w0, w0 + n) of the tile into the staging window: a thread owns J_OPT consecutive slots.
// CHECKED = some run of the tile leaves the staged row-id windows (rare): those slots read global memory.
template <bool CHECKED>
__device__ __forceinline__ void tile_expand(TileSmem& sm, const TileSlots& ts, const TileGeom& g, const JoinSide& L, const JoinSide& R,
                                            uint32_t w0, uint32_t n, uint32_t n_threads) {
    uint2* stage = reinterpret_cast<uint2*>(&sm.gs[0][0]);
    const int32_t* rid_all = &sm.rid[0][0];
    const int l_off = (int)(g.x0 & 3), r_off = (JW + 8) + (int)(g.y0 & 3);
    for (uint32_t s0 = w0 + threadIdx.x * J_OPT; s0 < w0 + n; s0 += n_threads * J_OPT) {
        uint32_t ci = ts.first_item[s0 / J_OPT];
        uint2 e = ts.ent[ci], nx = ts.ent[ci + 1];
        const uint32_t pb = stage_swz(s0 - w0);                     // s0 - w0 is a multiple of J_OPT = 8
#pragma unroll
        for (int q = 0; q < J_OPT; q++) {
            const uint32_t sl = s0 + q;
            if (sl < w0 + n) {
                if (sl >= (nx.x & JE_OFF_MASK)) {                   // exactly one step: entries are non-empty
                    ci++;
                    e = nx;
                    nx = ts.ent[ci + 1];
                }
                const int idx = ((int)e.x >> 16) + (int)sl;         // the run's row for this slot, as an index into rid_all
                const bool in_l = (e.x & JE_CLASS_B) != 0;
                uint32_t other;
                if (!CHECKED || idx < (in_l ? l_off + g.wl : r_off + g.wr)) other = (uint32_t)rid_all[idx];
                else other = (uint32_t)(in_l ? __ldg(L.rid + g.x0 + (idx - l_off)) : __ldg(R.rid + g.y0 + (idx - r_off)));
                stage[pb ^ (uint32_t)q] = in_l ? make_uint2(other, e.y) : make_uint2(e.y, other);
            }
        }
    }
}

// staging window -> global memory: lane i of a warp moves slot base + i, so every store instruction of a warp writes 128
// contiguous bytes per column at any alignment (no vector-alignment cases, one 64-bit shared-memory load per pair)
__device__ __forceinline__ void tile_flush(const TileSmem& sm, unsigned long long g_w0, uint32_t n, int32_t* __restrict__ out_l,
                                           int32_t* __restrict__ out_r) {
    const uint2* stage = reinterpret_cast<const uint2*>(&sm.gs[0][0]);
    int32_t* __restrict__ ol = out_l + g_w0;
    int32_t* __restrict__ orr = out_r + g_w0;
    for (uint32_t i = threadIdx.x; i < n; i += JB) {
        const uint2 v = stage[stage_swz(i)];
        st_stream_u32(reinterpret_cast<uint32_t*>(ol + i), v.x);
        st_stream_u32(reinterpret_cast<uint32_t*>(orr + i), v.y);
    }
}

#ifdef GIQLB_JOIN_DEBUG_STATS
// Per-phase cycle sums of the fused kernel (debug builds only; read with giqlb_debug_join_stats).
// [0] tiles [1] ticket+geometry [2] bulk-copy wait [3] runs [4] scan [5] table [6] expand (thread 0) [7] barrier wait after expand (thread 0)
// [8] flush [9] look-back (lane 31 of the last warp) [10] look-back steps [11] look-back polls [12] whole tile
__device__ unsigned long long g_join_dbg[32];
#define JDBG_T(var) const long long var = clock64()
#define JDBG_ADD(slot, v) atomicAdd(&g_join_dbg[slot], (unsigned long long)(v))
#else
#define JDBG_T(var)
#define JDBG_ADD(slot, v)
#endif

// Decoupled look-back of tile t, run by ONE warp: sums the pair counts of the tiles before t.  Every step fetches
// 32 * J_LB predecessor status words in one L2 round trip (lane l owns the J_LB words nearest-first from base - l * J_LB)
// and uses them up to the nearest word that already carries an inclusive prefix; it only waits for words nearer than
// that one.  (Round 2, ncu: with 32 words per step the chain over the ~300 tiles in flight that had not yet published an
// inclusive prefix took longer than the expansion it overlaps — 14 % of the kernel's stall samples sat at the barrier
// behind it.)
__device__ __forceinline__ unsigned long long tile_lookback(volatile unsigned long long* st, int64_t t, int* __restrict__ err_flag) {
    constexpr uint32_t NONE = 1u << 20;
    const int lane = (int)lane_id();
    unsigned long long prefix = 0;
    int64_t base = t - 1;
    bool done = t == 0;
    uint32_t spins = 0;
    while (!done) {
        unsigned long long sv[J_LB];
#pragma unroll
        for (int j = 0; j < J_LB; j++) {
            const int64_t k = base - (int64_t)lane * J_LB - j;
            sv[j] = 2ull << 62;                                          // before tile 0: an inclusive zero (= JF_FLAG_INCL)
            if (k >= 0) sv[j] = st[k];
        }
        uint32_t first_incl;
        while (true) {
            uint32_t my_incl = NONE, my_unpub = NONE;
#pragma unroll
            for (int j = J_LB - 1; j >= 0; j--) {                        // descending: the nearest (smallest j) wins
                const unsigned long long flag = sv[j] & ~JF_VALUE_MASK;
                if (flag == JF_FLAG_INCL) my_incl = (uint32_t)(lane * J_LB + j);
                if (flag == 0) my_unpub = (uint32_t)(lane * J_LB + j);
            }
            first_incl = __reduce_min_sync(FULL_MASK, my_incl);
            const uint32_t first_unpub = __reduce_min_sync(FULL_MASK, my_unpub);
            if (first_unpub >= first_incl) break;                        // every word nearer than the first inclusive one is published
            if (lane == 0) JDBG_ADD(11, 1);
            if (++spins > JF_SPIN_LIMIT) {
                *err_flag = 1;
                return prefix;
            }
            // poll again ONLY the words still needed (740 resident tiles re-reading whole windows would by themselves
            // load the L2 with terabytes per second of status traffic)
            __nanosleep(GIQLB_JOIN_BACKOFF_NS);
#pragma unroll
            for (int j = 0; j < J_LB; j++) {
                const uint32_t pos = (uint32_t)(lane * J_LB + j);
                if ((sv[j] & ~JF_VALUE_MASK) == 0 && pos < first_incl) sv[j] = st[base - (int64_t)pos];
            }
        }
        unsigned long long part_sum = 0;
#pragma unroll
        for (int j = 0; j < J_LB; j++)
            if ((uint32_t)(lane * J_LB + j) <= first_incl) part_sum += sv[j] & JF_VALUE_MASK;
        prefix += warp_reduce_add<unsigned long long>(part_sum);
        if (lane == 0) JDBG_ADD(10, 1);
        if (first_incl != NONE) done = true;
        else base -= 32 * J_LB;
    }
    return prefix;
}

__global__ void __launch_bounds__(JB)
join_fused_kernel(JoinSide L, JoinSide R, const TileRec* __restrict__ recs, int64_t n_tiles,
                  unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket, int32_t* __restrict__ out_l,
                  int32_t* __restrict__ out_r, int64_t capacity, unsigned long long* __restrict__ totals /* [0] pairs */,
                  uint32_t* __restrict__ big_count, uint32_t* __restrict__ big_tile, unsigned long long* __restrict__ big_off,
                  unsigned long long* __restrict__ big_cnt, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];   // > 48 KB with 512-thread tiles: dynamic, opted in by the host
    TileSmem& sm = *reinterpret_cast<TileSmem*>(dyn_smem);
    TileSlots& ts = *reinterpret_cast<TileSlots*>(dyn_smem + sizeof(TileSmem));
    JDBG_T(c_begin);
#if GIQLB_JOIN_TICKET
    if (threadIdx.x == 0) {
        sm.tile = atomicAdd(ticket, 1u);
        cuda::ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&sm.bar[0]), 1);
        cuda::ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&sm.bar[1]), 1);
        cuda::ptx::fence_mbarrier_init(cuda::ptx::sem_release, cuda::ptx::scope_cluster);
    }
    __syncthreads();
    const int64_t t = sm.tile;
#else
    // tile = block index: relies on CTAs being dispatched in blockIdx order, as cub::DeviceScan's look-back does; saves
    // the ticket's L2 round trip and a block barrier at the head of every tile (the look-back's spin limit still bounds
    // a wait on a tile that was never scheduled)
    const int64_t t = blockIdx.x;
    if (threadIdx.x == 0) {
        cuda::ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&sm.bar[0]), 1);
        cuda::ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&sm.bar[1]), 1);
        cuda::ptx::fence_mbarrier_init(cuda::ptx::sem_release, cuda::ptx::scope_cluster);
    }
#endif
    TileGeom g;
    {
        const int4* rp = reinterpret_cast<const int4*>(recs + t);
        const int4 a = __ldg(rp), b = __ldg(rp + 1);
        g.x0 = a.x; g.y0 = a.y; g.nl = a.z; g.nr = a.w; g.wl = b.x; g.wr = b.y;
    }
    if (g.nl < 0) return;   // the partition kernel refused an index whose sort is unfinished (uniform over the grid)
    JDBG_T(c_geom);
    if (threadIdx.x < 32) {
        __syncwarp();                          // lane 0's barrier initialisation precedes the other lanes' copies
        tile_stage_warp(sm, g, L, R);
        if (threadIdx.x == 0) bar_wait(&sm.bar[0], 0);   // one thread polls the barrier; the others sleep at the hardware barrier below
    }
    __syncthreads();
    JDBG_T(c_loaded);
    uint32_t code[J_IPT], cnt[J_IPT];
    bool leaves = false;
    tile_runs(sm, g, L, R, code, cnt, &leaves);
    JDBG_T(c_runs);
    if (threadIdx.x == JB - 1) bar_wait(&sm.bar[1], 0);   // row ids have landed (also: never exit with bulk copies in flight); the
                                                          // scan's barriers publish them to the block before the item table reads them
    // one block scan carries both the pair count (low 40 bits: < 1024 * 2^30) and the non-empty item count (high bits)
    unsigned long long sum = 0;
#pragma unroll
    for (int k = 0; k < J_IPT; k++) sum += (unsigned long long)cnt[k] + (cnt[k] ? (1ull << 40) : 0ull);
    unsigned long long tot_packed;
    const unsigned long long ex_packed = block_excl_scan_add<unsigned long long>(sum, sm.scratch, &tot_packed);   // two barriers: gs / ge are dead now
    const unsigned long long tot = tot_packed & ((1ull << 40) - 1), ex64 = ex_packed & ((1ull << 40) - 1);
    const uint32_t tot_items = (uint32_t)(tot_packed >> 40), ex_items = (uint32_t)(ex_packed >> 40);
    const bool big = tot > (unsigned long long)J_BIG;
    JDBG_T(c_scan);
    volatile unsigned long long* st = status;
    if (threadIdx.x == JB - 1) {
        st[t] = (t == 0 ? JF_FLAG_INCL : JF_FLAG_AGG) | tot;   // publish the tile's count as early as possible
    }
    const uint32_t tot32 = (uint32_t)tot;
    if (!big && tot32) tile_slot_table(sm, ts, g, code, cnt, (uint32_t)ex64, ex_items, tot32, tot_items);
    const bool checked = __syncthreads_or(leaves);   // item table complete; does any run leave the staged windows?
    JDBG_T(c_table);
    constexpr uint32_t N_EXPAND = JB - 32;     // warps 0..6 expand the tile while the last warp looks back
    static_assert(J_STAGE % J_OPT == 0, "windows start on chunk boundaries");
    if (threadIdx.x >= N_EXPAND) {
        // The last warp sums the counts of the tiles before this one (decoupled look-back) while the other warps expand the tile.
        const int lane = (int)lane_id();
        const unsigned long long prefix = tile_lookback(st, t, err_flag);
#ifdef GIQLB_JOIN_DEBUG_STATS
        if (lane == 31) JDBG_ADD(9, clock64() - c_table);
#endif
        if (lane == 31) {
            if (t > 0) st[t] = JF_FLAG_INCL | (prefix + tot);
            sm.tile_off = prefix;
            if (t == n_tiles - 1) totals[0] = prefix + tot;
            if (big) {   // shared among many CTAs by join_emit_big_kernel
                const uint32_t slot = atomicAdd(big_count, 1u);
                big_tile[slot] = (uint32_t)t;
                big_off[slot] = prefix;
                big_cnt[slot] = tot;
            }
        }
    } else if (!big && tot32) {
        const uint32_t n = min((uint32_t)J_STAGE, tot32);
        if (checked) tile_expand<true>(sm, ts, g, L, R, 0, n, N_EXPAND);
        else tile_expand<false>(sm, ts, g, L, R, 0, n, N_EXPAND);
    }
    JDBG_T(c_expanded);
    __syncthreads();                           // first staging window complete, sm.tile_off published
    JDBG_T(c_synced);
    if (big || tot32 == 0) return;
    const unsigned long long off_t = sm.tile_off;
    if (off_t + tot > (unsigned long long)capacity) return;   // the caller's buffer is too small: write nothing
    for (uint32_t w0 = 0; w0 < tot32; w0 += J_STAGE) {
        const uint32_t n = min((uint32_t)J_STAGE, tot32 - w0);
        if (w0) {                              // further windows (rare): everyone expands
            __syncthreads();                   // the previous window has been flushed
            if (checked) tile_expand<true>(sm, ts, g, L, R, w0, n, JB);
            else tile_expand<false>(sm, ts, g, L, R, w0, n, JB);
            __syncthreads();
        }
        tile_flush(sm, off_t + w0, n, out_l, out_r);
    }
#ifdef GIQLB_JOIN_DEBUG_STATS
    if (threadIdx.x == 0) {
        const long long c_end = clock64();
        JDBG_ADD(0, 1);
        JDBG_ADD(1, c_geom - c_begin);
        JDBG_ADD(2, c_loaded - c_geom);
        JDBG_ADD(3, c_runs - c_loaded);
        JDBG_ADD(4, c_scan - c_runs);
        JDBG_ADD(5, c_table - c_scan);
        JDBG_ADD(6, c_expanded - c_table);
        JDBG_ADD(7, c_synced - c_expanded);
        JDBG_ADD(8, c_end - c_synced);
        JDBG_ADD(12, c_end - c_begin);
    }
#endif
}

// Tiles with more than J_BIG pairs.  Their output is cut into J_PART-pair units that are dealt round-robin to the
// CTAs; a CTA recomputes only the tiles it owns a unit of.
__global__ void __launch_bounds__(JB)
join_emit_big_kernel(JoinSide L, JoinSide R, const int32_t* __restrict__ part, const uint32_t* __restrict__ big_count,
                     const uint32_t* __restrict__ big_tile, const unsigned long long* __restrict__ big_off,
                     const unsigned long long* __restrict__ big_cnt, int32_t* __restrict__ out_l, int32_t* __restrict__ out_r,
                     int64_t capacity) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    TileSmem& sm = *reinterpret_cast<TileSmem*>(dyn_smem);
    TileOffsets& so = *reinterpret_cast<TileOffsets*>(dyn_smem + sizeof(TileSmem));
    const uint32_t n_big = *big_count;
    if (n_big == 0) return;
    if (threadIdx.x == 0) {
        cuda::ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&sm.bar[0]), 1);
        cuda::ptx::mbarrier_init(reinterpret_cast<uint64_t*>(&sm.bar[1]), 1);
        cuda::ptx::fence_mbarrier_init(cuda::ptx::sem_release, cuda::ptx::scope_cluster);
    }
    __syncthreads();
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(out_l) | reinterpret_cast<uintptr_t>(out_r)) & 15) == 0;
    uint32_t parity = 0;
    unsigned long long unit_base = 0;
    for (uint32_t bi = 0; bi < n_big; bi++) {
        const unsigned long long tot = big_cnt[bi], n_parts = (tot + J_PART - 1) / J_PART;
        // my first unit at or after unit_base
        const unsigned long long first = unit_base + ((unsigned long long)blockIdx.x + gridDim.x - unit_base % gridDim.x) % gridDim.x;
        const unsigned long long off_t = big_off[bi];
        if (first < unit_base + n_parts && off_t + tot <= (unsigned long long)capacity) {
            const int64_t t = big_tile[bi];
            const TileGeom g = tile_geom(part, t, L.n, R.n);
            if (threadIdx.x == 0) tile_stage(sm, g, L, R);
            bar_wait(&sm.bar[0], parity);
            uint32_t code[J_IPT], cnt[J_IPT];
            tile_runs(sm, g, L, R, code, cnt);
            tile_offsets(sm, so, code, cnt);
            __syncthreads();
            bar_wait(&sm.bar[1], parity);
            parity ^= 1;
            for (unsigned long long u = first; u < unit_base + n_parts; u += gridDim.x) {
                const unsigned long long p = u - unit_base;
                tile_emit(sm, so, g, L, R, off_t, p * J_PART, min((p + 1) * (unsigned long long)J_PART, tot), out_l, out_r, vec_ok);
            }
            __syncthreads();   // everyone is done with the staged tile before the next one overwrites it
        }
        unit_base += n_parts;
    }
}

// ---- general path (degenerate rows present) --------------------------------------------------------
__device__ __forceinline__ void general_candidates(uint32_t lgs, uint32_t lge, const uint32_t* __restrict__ rgs,
                                                   const uint32_t* __restrict__ pmax, int64_t nr, int64_t& lo, int64_t& hi) {
    hi = lower_bound_dev<uint32_t, int64_t>(rgs, 0, nr, lge);      // r.gs < l.ge
    lo = upper_bound_dev<uint32_t, int64_t>(pmax, 0, hi, lgs);     // first j with running max(r.ge) > l.gs
}

__global__ void __launch_bounds__(JB)
join_count_tiles_general_kernel(JoinSide L, JoinSide R, const uint32_t* __restrict__ pmax, const uint32_t* __restrict__ m1,
                                const uint32_t* __restrict__ m2, unsigned long long* __restrict__ tile_cnt) {
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
            // candidates are visited through the block maxima of R.ge, so a long right row early on the chromosome does
            // not make every later left row test all the rows in between
            for (int64_t k = next_end_above(R.ge, m1, m2, lo, hi, lgs); k < hi; k = next_end_above(R.ge, m1, m2, k + 1, hi, lgs)) cnt++;
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
join_emit_general_kernel(JoinSide L, JoinSide R, const uint32_t* __restrict__ pmax, const uint32_t* __restrict__ m1,
                         const uint32_t* __restrict__ m2, int64_t n_tiles, const unsigned long long* __restrict__ tile_off, int32_t* __restrict__ out_l,
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
            for (int64_t k = next_end_above(R.ge, m1, m2, lo, hi, lgs); k < hi; k = next_end_above(R.ge, m1, m2, k + 1, hi, lgs)) cnt++;
        }
        unsigned long long tot;
        unsigned long long o = carry + block_excl_scan_add<unsigned long long>(cnt, s_scratch, &tot);
        carry += tot;
        if (cnt) {
            int32_t lr = L.rid[i];
            for (int64_t k = next_end_above(R.ge, m1, m2, lo, hi, lgs); k < hi; k = next_end_above(R.ge, m1, m2, k + 1, hi, lgs)) {
                out_l[o] = lr;
                out_r[o] = R.rid[k];
                o++;
            }
        }
        __syncthreads();
    }
}

// ---- SEMI / ANTI and per-row counts, probing from the raw (unsorted) left columns --------------------
// The left rows arrive in input order, so neighbouring threads search unrelated places of the index: a plain binary
// search is ~20 dependent L2 round trips per row.  The searches go through the index's bucket directories instead (two
// adjacent table entries, then a handful of rows of one or two cache lines).
struct ProbeDirs {
    const uint32_t* gs;    // directory over rgs
    const uint32_t* ges;   // directory over rge_sorted (count probes of proper right tables)
    int shift;
    uint32_t last;         // last bucket of a directory (entries - 2)
};

// One probe row against dir_exists: exists <=> pmax[ub - 1] > l.gs with ub = number of rows with r.gs < l.ge.
// f0 = {p, gs[p], pmax[p - 1], pmax[p]}, f1 = {gs[p + 1], pmax[p + 1]}, p = first row of the bucket of l.ge or later.
__device__ __forceinline__ uint8_t exists_resolve(uint4 f0, uint4 f1, uint32_t b, uint32_t gs, uint32_t ge, const ProbeDirs& D,
                                                 const uint32_t* __restrict__ rgs, const uint32_t* __restrict__ pmax) {
    if (ge <= f0.y) return f0.z > gs;                     // ub = p (pmax[-1] = 0 is below every linear coordinate that can overlap)
    if (f0.w > gs) return 1;                              // row p starts below l.ge: ub > p, and pmax only grows
    if (ge <= f1.x) return 0;                             // ub = p + 1
    if (f1.y > gs) return 1;
    const int64_t p1 = (int64_t)__ldg(D.gs + b + 1);      // three or more rows of this bucket start below l.ge?
    if (p1 <= (int64_t)f0.x + 2) return 0;
    const int64_t ub = lower_bound_dev<uint32_t, int64_t>(rgs, (int64_t)f0.x + 2, p1, ge);
    return __ldg(pmax + ub - 1) > gs;
}

// (measured, 50 M probe rows against 1 M: two or four rows per thread with their directory loads in flight together run SLOWER —
//  0.48 / 0.67 ms against 0.46 — and so does nothing at 64 resident warps: the kernel sits at the rate at which the L2 serves
//  random sectors, about 120 G/s, which also bounds the plain gather kernels)
__global__ void __launch_bounds__(256)
probe_exists_kernel(ProbeCols P, const uint32_t* __restrict__ rgs, const uint32_t* __restrict__ pmax, int64_t nr, ProbeDirs D,
                    const uint4* __restrict__ fat /* giqlb_index::dir_exists */, uint8_t* __restrict__ exists) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t gs, ge;
        uint8_t e = 0;
        if (nr > 0 && probe_linearise(P, i, gs, ge)) {
            const uint32_t b = min(ge >> D.shift, D.last);
            const uint4* fp = fat + 2 * (size_t)b;
            e = exists_resolve(__ldg(fp), __ldg(fp + 1), b, gs, ge, D, rgs, pmax);
        }
        exists[i] = e;
    }
}

__global__ void __launch_bounds__(256)
probe_count_kernel(ProbeCols P, const uint32_t* __restrict__ rgs, const uint32_t* __restrict__ rge,
                   const uint32_t* __restrict__ pmax, const uint32_t* __restrict__ rge_sorted, int64_t nr,
                   int r_has_degenerate, ProbeDirs D, const uint32_t* __restrict__ m1, const uint32_t* __restrict__ m2,
                   long long* __restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t gs, ge;
        long long c = 0;
        if (nr > 0 && probe_linearise(P, i, gs, ge)) {
            const int64_t ub = dir_lower_bound(D.gs, D.shift, D.last, rgs, ge);
            if (!r_has_degenerate && ge >= gs) {
                // #{r.gs < l.ge} - #{r.ge <= l.gs}: every r ending at or before l.gs also starts before l.ge
                c = ub - dir_upper_bound(D.ges, D.shift, D.last, rge_sorted, gs);
            } else {
                int64_t lb = upper_bound_dev<uint32_t, int64_t>(pmax, 0, ub, gs);
                for (int64_t k = next_end_above(rge, m1, m2, lb, ub, gs); k < ub; k = next_end_above(rge, m1, m2, k + 1, ub, gs)) c++;
            }
        }
        counts[i] = c;
    }
}

// ---- proper / degenerate split of an index ---------------------------------------------------------------------------
// One stable compaction pass deals the rows of a sorted index into two sorted sub-indexes: proper rows (end > start) and
// degenerate ones.  The join then runs its fast path on the proper rows and the literal-predicate scan only where a
// degenerate row is involved, instead of sending the whole join down the general path for one zero-width row.
struct InIsProper {
    const uint32_t *gs, *ge;
    __device__ unsigned long long operator()(int64_t i) const { return ge[i] > gs[i] ? 1ull : 0ull; }
};
struct OutSplitRows {
    const uint32_t *gs, *ge;
    const int32_t* rid;
    uint32_t *pgs, *pge, *dgs, *dge;
    int32_t *prid, *drid;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        if (v) {
            pgs[excl] = gs[i]; pge[excl] = ge[i]; prid[excl] = rid[i];
        } else {
            const int64_t d = i - (int64_t)excl;
            dgs[d] = gs[i]; dge[d] = ge[i]; drid[d] = rid[i];
        }
    }
};

inline JoinSide side_of(const giqlb_index* ix) { return JoinSide{ix->gs, ix->ge, ix->rid, ix->n}; }

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------------------
#ifdef GIQLB_JOIN_DEBUG_STATS
extern "C" int giqlb_debug_join_stats(unsigned long long* out32, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, g_join_dbg, sizeof(g_join_dbg));
    if (reset) {
        unsigned long long z[32] = {0};
        cudaMemcpyToSymbol(g_join_dbg, z, sizeof(z));
    }
    return 0;
}
#endif
extern "C" int giqlb_joinplan_free(giqlb_ctx* ctx, giqlb_joinplan* plan) {
    if (!plan) return GIQLB_OK;
    if (ctx) {
        dev_free(ctx, plan->part);
        dev_free(ctx, plan->item_code);
        dev_free(ctx, plan->tile_off);
        dev_free(ctx, plan->unit_off);
    }
    delete plan;
    return GIQLB_OK;
}

// Enqueues partition + count + scan.  The pair total lands in ctx->d_scalars[0].
static int plan_enqueue(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, giqlb_joinplan** out) {
    if (L->layout != R->layout) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects: indexes were built on different layouts");
    GIQLB_TRY(require_row_ids(ctx, L, "intersects"));
    GIQLB_TRY(require_row_ids(ctx, R, "intersects"));
    if (L->grouped_by_strand || R->grouped_by_strand)
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects: strand-grouped indexes are not accepted");
    GIQLB_TRY(giqlb_index_resolve(ctx, L, R));
    giqlb_joinplan* p = new (std::nothrow) giqlb_joinplan();
    if (!p) return GIQLB_ERR_NOMEM;
    p->L = L;
    p->R = R;
    p->general = (L->n_degenerate + R->n_degenerate) > 0;
    const bool empty = L->n == 0 || R->n == 0;
    p->n_tiles = empty ? 0 : (p->general ? ceil_div(L->n, JT) : ceil_div(L->n + R->n, JT));
    const int64_t n_tiles = p->n_tiles;
    unsigned long long* cnt = nullptr;
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(ctx->d_scalars);
    int st = dev_alloc(ctx, &cnt, (size_t)n_tiles + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &p->tile_off, (size_t)n_tiles + 1);
    if (st == GIQLB_OK && !p->general) st = dev_alloc(ctx, &p->unit_off, (size_t)n_tiles + 1);
    if (st == GIQLB_OK && !p->general) st = dev_alloc(ctx, &p->part, (size_t)n_tiles + 1);
    if (st == GIQLB_OK && !p->general) st = dev_alloc(ctx, &p->item_code, (size_t)n_tiles * JT);
    if (st == GIQLB_OK && n_tiles == 0) {
        if (cudaMemsetAsync(totals, 0, 16, ctx->stream) != cudaSuccess || cudaMemsetAsync(p->tile_off, 0, 8, ctx->stream) != cudaSuccess)
            st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "intersects: memset failed");
    } else if (st == GIQLB_OK && p->general) {
        st = giqlb_index_ensure_pmax(ctx, R);
        if (st == GIQLB_OK) st = giqlb_index_ensure_block_max(ctx, R);
        if (st == GIQLB_OK) {
            PhaseSpan ps(ctx, GIQLB_PH_JOIN_COUNT);
            join_count_tiles_general_kernel<<<(unsigned)n_tiles, JB, 0, ctx->stream>>>(side_of(L), side_of(R), R->pmax, R->bmax1, R->bmax2, cnt);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
        }
        if (st == GIQLB_OK) st = giqlb_excl_scan_u64(ctx, cnt, p->tile_off, n_tiles, totals);
    } else if (st == GIQLB_OK) {
        {
            PhaseSpan ps(ctx, GIQLB_PH_JOIN_COUNT);
            join_partition_kernel<<<(unsigned)ceil_div(n_tiles + 1, 256), 256, 0, ctx->stream>>>(L->gs, L->n, R->gs, R->n, n_tiles, p->part, nullptr);
            join_count_tiles_kernel<<<(unsigned)n_tiles, JB, 0, ctx->stream>>>(side_of(L), side_of(R), p->part, p->item_code, cnt);
            ctx->launches_this_call += 2;
            ctx->stats.kernels_total += 2;
        }
        PhaseSpan ps(ctx, GIQLB_PH_JOIN_SCAN);
        join_scan_tiles_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, n_tiles, p->tile_off, p->unit_off, totals);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
    }
    if (st == GIQLB_OK && cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "intersects: plan launch failed");
    dev_free(ctx, cnt);
    if (st != GIQLB_OK) {
        giqlb_joinplan_free(ctx, p);
        return st;
    }
    *out = p;
    return GIQLB_OK;
}

static int emit_enqueue(giqlb_ctx* ctx, giqlb_joinplan* p, int32_t* out_l, int32_t* out_r, int64_t capacity) {
    const int64_t n_tiles = p->n_tiles;
    if (n_tiles == 0) return GIQLB_OK;
    PhaseSpan ps(ctx, GIQLB_PH_JOIN_EMIT);
    if (p->general) {
        join_emit_general_kernel<<<(unsigned)n_tiles, JB, 0, ctx->stream>>>(side_of(p->L), side_of(p->R), p->R->pmax, p->R->bmax1, p->R->bmax2, n_tiles,
                                                                            p->tile_off, out_l, out_r, capacity);
    } else {
        const int64_t cap = (int64_t)ctx->sm_count * 8;
        const int grid = (int)(n_tiles < cap ? n_tiles : cap);
        join_emit_kernel<<<grid, JB, 0, ctx->stream>>>(side_of(p->L), side_of(p->R), p->part, n_tiles, p->item_code, p->tile_off, p->unit_off,
                                                       out_l, out_r, capacity);
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

static giqlb_index* sub_index(const giqlb_index* ix, int64_t n, int64_t n_degenerate) {
    giqlb_index* s = new (std::nothrow) giqlb_index();
    if (!s) return nullptr;
    s->n = n;
    s->layout = ix->layout;
    s->tie_by_end = ix->tie_by_end;
    s->n_degenerate = n_degenerate;
    s->n_negative = 0;            // not tracked for the parts (only n_degenerate steers the join)
    s->resolved = true;
    s->block_off[1] = n;
    for (int i = 2; i <= GIQLB_MAX_STRAND_CODES; i++) s->block_off[i] = n;
    return s;
}

static int index_split(giqlb_ctx* ctx, const giqlb_index* ix, giqlb_index** proper, giqlb_index** degenerate) {
    using namespace giqlb_scan;
    *proper = sub_index(ix, ix->n - ix->n_degenerate, 0);
    *degenerate = sub_index(ix, ix->n_degenerate, ix->n_degenerate);
    if (!*proper || !*degenerate) return GIQLB_ERR_NOMEM;
    giqlb_index *P = *proper, *D = *degenerate;
    int st = dev_alloc(ctx, &P->gs, (size_t)P->n + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &P->ge, (size_t)P->n + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &P->rid, (size_t)P->n + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &D->gs, (size_t)D->n + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &D->ge, (size_t)D->n + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &D->rid, (size_t)D->n + 1);
    if (st == GIQLB_OK && ix->n > 0) {
        PhaseSpan ps(ctx, GIQLB_PH_SELECT);
        st = launch_scan<OpAddU64>(ctx, ix->n, InIsProper{ix->gs, ix->ge}, OutSplitRows{ix->gs, ix->ge, ix->rid, P->gs, P->ge, D->gs, D->ge, P->rid, D->rid});
    }
    return st;
}

// Join with degenerate rows (end <= start) on either side: the proper rows of both tables take the fused fast path; the
// literal predicate is scanned only for (degenerate L rows x all R rows) and (proper L rows x degenerate R rows) — the
// three parts are disjoint and cover every pair.  Pairs land back to back in out_l / out_r.
static int join_with_degenerate_rows(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, int32_t* out_l, int32_t* out_r, int64_t capacity,
                                     int64_t* total_out) {
    giqlb_index *lp = nullptr, *ld = nullptr, *rp = nullptr, *rd = nullptr;
    giqlb_joinplan *plan_a = nullptr, *plan_b = nullptr;
    int64_t t_fast = 0, t_a = 0, t_b = 0;
    int st = index_split(ctx, L, &lp, &ld);
    if (st == GIQLB_OK) st = index_split(ctx, R, &rp, &rd);
    if (st == GIQLB_OK) {   // 1. proper x proper: the fused single-pass join (reports the exact total even when it does not fit)
        st = giqlb_intersects_join_impl(ctx, lp, rp, out_l, out_r, capacity, &t_fast, false);
        if (st == GIQLB_ERR_CAPACITY) st = GIQLB_OK;
    }
    if (st == GIQLB_OK && ld->n > 0 && R->n > 0) {   // 2. degenerate L rows against the whole right table
        st = plan_enqueue(ctx, ld, R, &plan_a);
        if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
        if (st == GIQLB_OK) t_a = ctx->h_scalars[0];
    }
    if (st == GIQLB_OK && lp->n > 0 && rd->n > 0) {  // 3. proper L rows against the degenerate right rows
        st = plan_enqueue(ctx, lp, rd, &plan_b);
        if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
        if (st == GIQLB_OK) t_b = ctx->h_scalars[0];
    }
    const int64_t total = t_fast + t_a + t_b;
    if (st == GIQLB_OK && total <= capacity) {
        if (plan_a && t_a > 0) st = emit_enqueue(ctx, plan_a, out_l + t_fast, out_r + t_fast, t_a);
        if (st == GIQLB_OK && plan_b && t_b > 0) st = emit_enqueue(ctx, plan_b, out_l + t_fast + t_a, out_r + t_fast + t_a, t_b);
        if (st == GIQLB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "intersects_join: sync failed");
    }
    giqlb_joinplan_free(ctx, plan_a);
    giqlb_joinplan_free(ctx, plan_b);
    giqlb_index_free(ctx, lp);
    giqlb_index_free(ctx, ld);
    giqlb_index_free(ctx, rp);
    giqlb_index_free(ctx, rd);
    *total_out = st == GIQLB_OK ? total : -1;
    return st;
}

extern "C" int giqlb_intersects_plan(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, giqlb_joinplan** out,
                                     int64_t* total_host) {
    if (!ctx || !L || !R || !out) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_plan: bad argument");
    begin_call(ctx);
    *out = nullptr;
    giqlb_joinplan* p = nullptr;
    GIQLB_TRY(plan_enqueue(ctx, L, R, &p));
    int st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
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
    return giqlb_intersects_join_impl(ctx, L, R, out_l, out_r, capacity, total_host, false);
}

// speculative: when the degenerate-row counts of the indexes are still on the device, do not fetch them first —
// run the fused (proper-interval) path, bring the counts back with the pair total, and fall back to the general
// path only if a degenerate row turns up (the fused kernels are memory-safe on any input; their result is discarded).
int giqlb_intersects_join_impl(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, int32_t* out_l, int32_t* out_r, int64_t capacity,
                               int64_t* total_host, bool speculative) {
    if (!ctx || !L || !R || capacity < 0 || (capacity > 0 && (!out_l || !out_r)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_join: bad argument");
    begin_call(ctx);
    if (L->layout != R->layout) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects: indexes were built on different layouts");
    GIQLB_TRY(require_row_ids(ctx, L, "intersects"));
    GIQLB_TRY(require_row_ids(ctx, R, "intersects"));
    if (L->grouped_by_strand || R->grouped_by_strand)
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects: strand-grouped indexes are not accepted");
    const bool empty = L->n == 0 || R->n == 0;
    // maybe_unfinished: an index sorted with GIQLB_INDEX_SORT_SHORT_WINDOWS whose dense-window count has not come back yet.  The
    // fused path below checks it on the device and brings it back with the pair total; every other path looks first.
    const bool unchecked = L->maybe_unfinished || R->maybe_unfinished;
    const bool known = L->resolved && R->resolved;
    const bool speculate = !empty && ((speculative && !known) || (unchecked && known && L->n_degenerate + R->n_degenerate == 0));
    if (!speculate) {
        const int rs = giqlb_index_resolve(ctx, L, R);
        if (rs != GIQLB_OK) return rs;   // (GIQLB_RETRY_FULL_SORT included)
    }
    int64_t total = 0;
    int st = GIQLB_OK;
    if (!speculate && !empty && L->n_degenerate + R->n_degenerate > 0 && L->n_degenerate < L->n && R->n_degenerate < R->n) {
        // some degenerate rows: fast path for the proper rows, literal-predicate scan only where a degenerate row takes part
        st = join_with_degenerate_rows(ctx, L, R, out_l, out_r, capacity, &total);
    } else if (!speculate && (empty || L->n_degenerate + R->n_degenerate > 0)) {
        // a table of degenerate rows only: two-pass general path (the emit kernel writes nothing when the total exceeds `capacity`)
        giqlb_joinplan* p = nullptr;
        GIQLB_TRY(plan_enqueue(ctx, L, R, &p));
        st = capacity > 0 ? emit_enqueue(ctx, p, out_l, out_r, capacity) : GIQLB_OK;
        if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
        total = st == GIQLB_OK ? ctx->h_scalars[0] : -1;
        giqlb_joinplan_free(ctx, p);
    } else {
        const int64_t n_tiles = ceil_div(L->n + R->n, JT);
        int32_t* part = nullptr;
        unsigned long long* status = nullptr;   // [n_tiles] look-back words | [n_tiles] big offsets | ticket, big count | [n_tiles] big tiles
        unsigned long long* totals = reinterpret_cast<unsigned long long*>(ctx->d_scalars);
        TileRec* recs = nullptr;
        st = dev_alloc(ctx, &part, (size_t)n_tiles + 1);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &recs, (size_t)n_tiles + 1);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &status, (size_t)4 * n_tiles + 2);
        if (st == GIQLB_OK) {
            unsigned long long* big_off = status + n_tiles;
            unsigned long long* big_cnt = status + 2 * n_tiles;
            uint32_t* ticket = reinterpret_cast<uint32_t*>(status + 3 * n_tiles);
            uint32_t* big_count = ticket + 1;
            uint32_t* big_tile = reinterpret_cast<uint32_t*>(status + 3 * n_tiles + 1);
            cudaMemsetAsync(status, 0, ((size_t)n_tiles) * 8, ctx->stream);
            cudaMemsetAsync(ticket, 0, 8, ctx->stream);
            cudaMemsetAsync(totals, 0, 8, ctx->stream);
            {
                PhaseSpan ps(ctx, GIQLB_PH_JOIN_COUNT);
                join_partition_kernel<<<(unsigned)ceil_div(n_tiles + 1, 256), 256, 0, ctx->stream>>>(
                    L->gs, L->n, R->gs, R->n, n_tiles, part, recs, L->maybe_unfinished ? L->d_counters + 3 : nullptr, R->maybe_unfinished ? R->d_counters + 3 : nullptr);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
            {
                // the dynamic shared-memory limit is a per-device function attribute; a context is bound to one device
                if (!ctx->join_smem_opted_in) {
                    cudaFuncSetAttribute(join_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(TileSmem) + sizeof(TileSlots)));
                    cudaFuncSetAttribute(join_emit_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(TileSmem) + sizeof(TileOffsets)));
                    ctx->join_smem_opted_in = true;
                }
                PhaseSpan ps(ctx, GIQLB_PH_JOIN_EMIT);
                join_fused_kernel<<<(unsigned)n_tiles, JB, sizeof(TileSmem) + sizeof(TileSlots), ctx->stream>>>(side_of(L), side_of(R), recs, n_tiles, status, ticket, out_l, out_r,
                                                                                 capacity, totals, big_count, big_tile, big_off, big_cnt, ctx->d_err);
                join_emit_big_kernel<<<ctx->sm_count * 2, JB, sizeof(TileSmem) + sizeof(TileOffsets), ctx->stream>>>(side_of(L), side_of(R), part, big_count, big_tile, big_off,
                                                                                big_cnt, out_l, out_r, capacity);
                cudaMemcpyAsync(ctx->h_scalars + 8, big_count, 4, cudaMemcpyDeviceToHost, ctx->stream);
                ctx->launches_this_call += 2;
                ctx->stats.kernels_total += 2;
            }
            if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "intersects_join: launch failed");
        }
        giqlb_index* pending[2] = {!L->resolved || L->maybe_unfinished ? L : nullptr, (!R->resolved || R->maybe_unfinished) && R != L ? R : nullptr};
        for (int k = 0; k < 2 && st == GIQLB_OK; k++)   // the degenerate-row counts ride along with the pair total
            if (pending[k]) cudaMemcpyAsync(ctx->h_scalars + 32 + 4 * k, pending[k]->d_counters, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, totals, 1);
        total = st == GIQLB_OK ? ctx->h_scalars[0] : -1;
        bool degenerate = false, unfinished = false;
        for (int k = 0; k < 2 && st == GIQLB_OK; k++) {
            if (!pending[k]) continue;
            if (!pending[k]->resolved) {
                pending[k]->n_degenerate = ctx->h_scalars[32 + 4 * k];
                pending[k]->n_negative = ctx->h_scalars[33 + 4 * k];
                pending[k]->resolved = true;
            }
            if (pending[k]->maybe_unfinished) {
                pending[k]->maybe_unfinished = false;
                pending[k]->sort_unfinished = ctx->h_scalars[35 + 4 * k] != 0;
            }
            unfinished = unfinished || pending[k]->sort_unfinished;
            degenerate = degenerate || pending[k]->n_degenerate > 0;
        }
        if (st == GIQLB_OK && unfinished) {   // the kernels above did nothing; only the builder of the indexes can go on
            dev_free(ctx, part);
            dev_free(ctx, recs);
            dev_free(ctx, status);
            end_call(ctx);
            return GIQLB_RETRY_FULL_SORT;
        }
        if (st == GIQLB_OK && degenerate) {   // speculation failed: redo through the general path
            dev_free(ctx, part);
            dev_free(ctx, recs);
            dev_free(ctx, status);
            return giqlb_intersects_join_impl(ctx, L, R, out_l, out_r, capacity, total_host, false);
        }
        ctx->stats.reserved = (int32_t)(ctx->h_scalars[8] & 0xffffffff);   // tiles delegated to the big-tile kernel
        dev_free(ctx, part);
        dev_free(ctx, recs);
        dev_free(ctx, status);
    }
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
    if (n > 0 && R->n > 0) GIQLB_TRY(giqlb_index_ensure_exists_dir(ctx, R));
    if (n > 0) {
        int64_t blocks = ceil_div(n, 256 * 4);
        int64_t cap = (int64_t)ctx->sm_count * 16;
        PhaseSpan ps(ctx, GIQLB_PH_PROBE);
        probe_exists_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(P, R->gs, R->pmax, R->n, ProbeDirs{R->dir_gs, nullptr, R->dir_shift, R->dir_stride - 2},
                                                                                                R->dir_exists, exists_by_row);
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
    GIQLB_TRY(giqlb_index_resolve(ctx, R, nullptr));
    const int r_deg = R->n_degenerate > 0;
    if (!r_deg) GIQLB_TRY(giqlb_index_ensure_end_sorted(ctx, R));
    if (n > 0 && R->n > 0) GIQLB_TRY(giqlb_index_ensure_dirs(ctx, R, false, !r_deg));
    if (r_deg) GIQLB_TRY(giqlb_index_ensure_block_max(ctx, R));
    if (n > 0) {
        int64_t blocks = ceil_div(n, 256 * 4);
        int64_t cap = (int64_t)ctx->sm_count * 16;
        PhaseSpan ps(ctx, GIQLB_PH_PROBE);
        probe_count_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
            P, R->gs, R->ge, R->pmax, R->ge_sorted, R->n, r_deg, ProbeDirs{R->dir_gs, R->dir_ges, R->dir_shift, R->dir_stride - 2}, R->bmax1, R->bmax2,
            (long long*)counts_by_row);
        LAUNCH_CHECK(ctx);
    }
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = R->n;
    ctx->stats.rows_out = n;
    ctx->stats.algorithmic_bytes = 12 * (n + R->n) + 8 * n;
    return end_call(ctx);
}
