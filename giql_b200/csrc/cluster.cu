// cluster.cu — CLUSTER and MERGE as segmented scans over a start-sorted index.
//
// Semantics (reference, abdenlab/giql):
//   CLUSTER  src/giql/expanders/cluster.py:205-310, 430-434.  Per partition (chrom, + strand when
//            stranded := true, :243-245) ordered by start (:248):
//              flag_i = 0  iff  MAX(end) over the strictly preceding rows [+ distance] >= start_i   (:254-279)
//              id_i   = SUM(flag) over the default RANGE frame (:430-434): rows tied on start share the sum
//                       through their last peer; ids are 1-based and restart per partition.
//            Raw table coordinates, no canonicalisation (src/giql/expressions.py:223-226).
//   MERGE    src/giql/expanders/merge.py:183-439: GROUP BY chrom[, strand], id -> MIN(start), MAX(end)
//            (+ COUNT(*)), default ORDER BY chrom, start (:429-437).
//
// Device formulation over the linearised index (index.cu): the running max of end over ALL preceding rows
// (pmax, restarting per strand block) also serves as the per-chromosome running max, because every end of an
// earlier chromosome lies at or below the current chromosome's low guard cell: "has a predecessor in my
// partition" == pmax[i-1] > base[chrom(i)].  One decoupled look-back scan turns the flags into a running
// cluster count G; ids subtract G at the chromosome's first row; MERGE writes region G[i]-1 directly (the
// running count IS the output slot, so no compaction pass): the flagged row stores the start, the last row of
// the run stores the running max end.
#include "common.cuh"
#include "scan.cuh"

namespace {

using namespace giqlb_scan;

struct LayoutView {
    const uint32_t* base;   // [n_codes + 1]
    const int32_t* lo32;
    int32_t n_codes;
};

// One-entry memo for chrom_of: rows arrive sorted by linear coordinate, so consecutive lookups of a thread almost always fall
// into the same chromosome's cell range [lo, hi) and skip the binary search.
struct ChromMemo {
    uint32_t lo = 1u, hi = 0u;   // empty range
    int32_t c = 0;
};
__device__ __forceinline__ int32_t chrom_of(const LayoutView& V, uint32_t g);
__device__ __forceinline__ int32_t chrom_of(const LayoutView& V, uint32_t g, ChromMemo& m) {
    if (g >= m.lo && g < m.hi) return m.c;
    m.c = chrom_of(V, g);
    m.lo = __ldg(V.base + m.c);
    m.hi = m.c + 1 < V.n_codes ? __ldg(V.base + m.c + 1) : 0xffffffffu;
    return m.c;
}

// chromosome whose cell range holds linear coordinate g: last c with base[c] <= g
__device__ __forceinline__ int32_t chrom_of(const LayoutView& V, uint32_t g) {
    int32_t lo = 0, hi = V.n_codes;   // answer in [0, n_codes)
    while (hi - lo > 1) {
        int32_t mid = (lo + hi) >> 1;
        if (__ldg(V.base + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// Optional `predicate := cur <op> PREV(prev)` (cluster.py:281-299, 587-644): a row stays in the running cluster only when it
// is adjacent to the running-max edge AND the comparison holds against its IMMEDIATE sorted predecessor.  The two
// operand columns arrive as order-preserving int32 codes in sorted-row order (GIQLB_KEY_NULL = SQL NULL -> not TRUE).
struct ClusterPredicate {
    const int32_t* cur;    // code of the current row's operand, by sorted position (nullptr: no predicate)
    const int32_t* prev;   // code of the PREV(...) operand, by sorted position
    int op;
};
__device__ __forceinline__ bool predicate_holds(const ClusterPredicate& p, int64_t i) {
    if (!p.cur) return true;
    if (i == 0) return false;
    const int32_t a = p.cur[i], b = p.prev[i - 1];
    if (a == GIQLB_KEY_NULL || b == GIQLB_KEY_NULL) return false;
    switch (p.op) {
        case GIQLB_CMP_EQ: return a == b;
        case GIQLB_CMP_NE: return a != b;
        case GIQLB_CMP_LT: return a < b;
        case GIQLB_CMP_LE: return a <= b;
        case GIQLB_CMP_GT: return a > b;
        default: return a >= b;
    }
}

struct InClusterFlag {
    const uint32_t* gs;
    const uint32_t* pmax;
    LayoutView V;
    long long d;
    uint32_t carry;   // running max of end over the rows of EARLIER genome-range shards (0: none)
    ClusterPredicate pred;
    mutable ChromMemo memo;   // per-thread copy of the functor
    __device__ unsigned long long operator()(int64_t i) const {
        const uint32_t s = gs[i];
        uint32_t pm = carry;
        if (i > 0) pm = max(pm, pmax[i - 1]);
        chrom_of(V, s, memo);
        const uint32_t b = memo.lo;
        const bool joins = pm > b && (long long)pm + d >= (long long)s && predicate_holds(pred, i);
        return joins ? 0ull : 1ull;
    }
};

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ rid, int64_t n, int32_t* __restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = __ldg(src + rid[i]);
}
struct OutRunningCount {
    uint32_t* G;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const { G[i] = (uint32_t)(excl + v); }
};

// off[c] = running cluster count just before chromosome c's first row of this block
__global__ void cluster_chrom_offsets_kernel(const uint32_t* __restrict__ gs, const uint32_t* __restrict__ G, int64_t n,
                                             LayoutView V, uint32_t* __restrict__ off) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V.n_codes) return;
    int64_t first = lower_bound_dev<uint32_t, int64_t>(gs, 0, n, V.base[c]);
    off[c] = first > 0 ? G[first - 1] : 0u;
}

// last row of i's peer group (rows tied on start), galloping forward
__device__ __forceinline__ int64_t last_peer(const uint32_t* __restrict__ gs, int64_t i, int64_t n) {
    const uint32_t s = gs[i];
    if (i + 1 >= n || gs[i + 1] != s) return i;
    int64_t lo = i + 1, step = 1;   // invariant: gs[lo] == s
    while (lo + step < n && gs[lo + step] == s) {
        lo += step;
        step <<= 1;
    }
    int64_t hi = lo + step < n ? lo + step : n;   // gs[hi] != s or hi == n
    return upper_bound_dev<uint32_t, int64_t>(gs, lo, hi, s) - 1;
}

__global__ void __launch_bounds__(256)
cluster_ids_kernel(const uint32_t* __restrict__ gs, const int32_t* __restrict__ rid, const uint32_t* __restrict__ G,
                   const uint32_t* __restrict__ off, int64_t n, LayoutView V, int32_t* __restrict__ out_by_row) {
    ChromMemo memo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = last_peer(gs, i, n);
        const int32_t c = chrom_of(V, gs[i], memo);
        out_by_row[rid ? (int64_t)rid[i] : i] = (int32_t)(G[j] - off[c]);   // rid == nullptr: ids in sorted-row order
    }
}

__global__ void __launch_bounds__(256)
merge_emit_kernel(const uint32_t* __restrict__ gs, const uint32_t* __restrict__ pmax, const uint32_t* __restrict__ G,
                  int64_t n, LayoutView V, int64_t out_base, int64_t idx_base, int32_t* __restrict__ out_code,
                  int32_t* __restrict__ out_start, int32_t* __restrict__ out_end, long long* __restrict__ out_count) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t g = G[i];
        if (g == 0) continue;   // head fragment of a shard: continues a region owned by an earlier shard
        const bool first = i == 0 || G[i - 1] != g;
        const bool last = i + 1 == n || G[i + 1] != g;
        if (!first && !last) continue;
        const int64_t m = out_base + (int64_t)g - 1;
        const uint32_t s = gs[i];
        const int32_t c = chrom_of(V, s);
        const uint32_t b = __ldg(V.base + c) + 1u;
        const int32_t lo = __ldg(V.lo32 + c);
        if (first) {
            out_code[m] = c;
            out_start[m] = lo + (int32_t)(s - b);
            if (out_count) atomicAdd((unsigned long long*)&out_count[m], (unsigned long long)(-(long long)(idx_base + i)));
        }
        if (last) {
            out_end[m] = lo + (int32_t)(pmax[i] - b);
            if (out_count) atomicAdd((unsigned long long*)&out_count[m], (unsigned long long)(idx_base + i + 1));
        }
    }
}

// per-chromosome number of clusters STARTING in this shard: running count at the chromosome's last row minus the
// count before its first row
__global__ void cluster_chrom_counts_kernel(const uint32_t* __restrict__ gs, const uint32_t* __restrict__ G, int64_t n, LayoutView V,
                                            uint32_t* __restrict__ counts) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V.n_codes) return;
    const int64_t first = lower_bound_dev<uint32_t, int64_t>(gs, 0, n, V.base[c]);
    const int64_t last = lower_bound_dev<uint32_t, int64_t>(gs, 0, n, V.base[c + 1]);
    counts[c] = last > first ? G[last - 1] - (first > 0 ? G[first - 1] : 0u) : 0u;
}

__global__ void cluster_ids_offset_kernel(const uint32_t* __restrict__ gs, const int32_t* __restrict__ rid, const uint32_t* __restrict__ G,
                                          const uint32_t* __restrict__ off, const uint32_t* __restrict__ id_offset, int64_t n, LayoutView V,
                                          int32_t* __restrict__ out_by_row) {
    ChromMemo memo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = last_peer(gs, i, n);
        const int32_t c = chrom_of(V, gs[i], memo);
        out_by_row[rid ? (int64_t)rid[i] : i] = (int32_t)(G[j] - off[c] + id_offset[c]);
    }
}

// head fragment of a shard: rows before the first cluster start.  out[0] = number of head rows, out[1] = chromosome code,
// out[2] = canonical max end of the head rows
__global__ void merge_head_kernel(const uint32_t* __restrict__ gs, const uint32_t* __restrict__ pmax, const uint32_t* __restrict__ G, int64_t n,
                                  LayoutView V, long long* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int64_t lo = 0, hi = n;   // first row with G >= 1 (G is non-decreasing)
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (G[mid] < 1u) lo = mid + 1; else hi = mid;
    }
    out[0] = lo;
    out[1] = -1;
    out[2] = 0;
    if (lo > 0) {
        const int32_t c = chrom_of(V, gs[0]);
        out[1] = c;
        out[2] = (long long)__ldg(V.lo32 + c) + (long long)(pmax[lo - 1] - (__ldg(V.base + c) + 1u));
    }
}

__global__ void fill_u8_kernel(uint8_t* p, uint8_t v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

inline int grid_of(const giqlb_ctx* ctx, int64_t n, int per_thread) {
    int64_t want = ceil_div(n, 256 * per_thread), cap = (int64_t)ctx->sm_count * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// out[row[i]] = val[i] for a permutation `row` of [0, n).  Done directly, this is n random 4-byte writes: beyond the L2 (126 MB)
// every one of them costs a DRAM sector, and at 200 M rows the scatter alone took 10 ms (22 G writes/s).  Partitioning the pairs
// on the top 8 bits of the row id first, so that the writes fall into L2-resident 4 MB windows, hardly helps (7.1 ms: sub-sector
// writes are slow in L2 as well).  What works is to sort the (row, val) pairs on all but the lowest 4 bits of the row id with the
// radix sort (three passes of 1.3 ms at 200 M rows): neighbouring threads then write into the same two sectors and the scatter
// itself drops to 0.4 ms — 4.6 ms in all.
// (Row ids of a genome-range shard run over the whole table, not [0, n): the sort then sees only their lower bits and rows
// of several windows interleave — less locality, same result.)
#ifndef GIQLB_BINNED_SCATTER_MIN
#define GIQLB_BINNED_SCATTER_MIN (24ll << 20)   // rows: the output no longer fits L2 comfortably
#endif
__global__ void __launch_bounds__(256)
scatter_u32_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ val, int64_t n, uint32_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[row[i]] = val[i];
}

inline bool binned_scatter_wanted(int64_t n) { return n >= GIQLB_BINNED_SCATTER_MIN; }

// `val` is scratch of the caller (n entries, sorted-row order) and is left in an unspecified state
int scatter_by_row(giqlb_ctx* ctx, const int32_t* rid, uint32_t* val, int64_t n, int32_t* out) {
    uint32_t *ka = nullptr, *kb = nullptr, *vb = nullptr;
    GIQLB_TRY(dev_alloc(ctx, &kb, (size_t)n));
    int st = dev_alloc(ctx, &vb, (size_t)n);
    if (st == GIQLB_OK) {
        int bits = 0;
        while (bits < 32 && (1ll << bits) < n) bits++;
        const int shift = bits > 12 ? 4 : 0;
        int in_b = 0;
        // the passes ping-pong between (ka, val) and (kb, vb): the index's own row ids must stay intact, hence the copy
        st = dev_alloc(ctx, &ka, (size_t)n);
        if (st == GIQLB_OK && cudaMemcpyAsync(ka, rid, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
            st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "scatter_by_row: copy failed");
        if (st == GIQLB_OK) st = giqlb_radix_sort_pairs_u32(ctx, ka, val, kb, vb, n, shift, bits, 0, &in_b);
        if (st == GIQLB_OK) {
            scatter_u32_kernel<<<grid_of(ctx, n, 4), 256, 0, ctx->stream>>>(in_b ? kb : ka, in_b ? vb : val, n, reinterpret_cast<uint32_t*>(out));
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "scatter_by_row: launch failed");
        }
    }
    dev_free(ctx, ka);
    dev_free(ctx, kb);
    dev_free(ctx, vb);
    return st;
}

// CLUSTER ids by search instead of by scatter.  When the clusters are few (deep-coverage reads: 200 M rows fall into ~2 000
// clusters) an input row finds its id by itself: the number of cluster starts at or below its own linear start, minus the
// starts before its chromosome.  The rows are read and the ids written in INPUT order — coalesced — and nothing has to travel
// back through the sorted row ids (three radix passes + a scatter: 4.6 ms at 200 M rows against 1 ms here).
#ifndef GIQLB_CLUSTER_SEARCH_MAX
#define GIQLB_CLUSTER_SEARCH_MAX (4ll << 20)   // cluster starts: 16 MB of keys, resident in L2
#endif
__global__ void __launch_bounds__(256)
cluster_starts_kernel(const uint32_t* __restrict__ gs, const uint32_t* __restrict__ G, int64_t n, uint32_t* __restrict__ starts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t g = G[i];
        if (i == 0 ? g != 0u : g != G[i - 1]) starts[g - 1] = gs[i];
    }
}
// before[c] = cluster starts below chromosome c's cell range
__global__ void cluster_before_kernel(const uint32_t* __restrict__ base, int32_t n_codes, const uint32_t* __restrict__ starts, int64_t n_starts,
                                      uint32_t* __restrict__ before) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_codes) before[c] = (uint32_t)lower_bound_dev<uint32_t, int64_t>(starts, 0, n_starts, __ldg(base + c));
}
// first index in [0, n) with a[idx] > x, 32-bit (n <= GIQLB_CLUSTER_SEARCH_MAX): branch-free halving
__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t* __restrict__ a, uint32_t n, uint32_t x) {
    uint32_t lo = 0;
    while (n > 0) {
        const uint32_t half = n >> 1;
        const bool right = __ldg(a + lo + half) <= x;
        lo = right ? lo + half + 1 : lo;
        n = right ? n - half - 1 : half;
    }
    return lo;
}
__global__ void __launch_bounds__(256)
cluster_ids_search_kernel(const int32_t* __restrict__ code, const int32_t* __restrict__ start, int64_t n, int32_t start_adj,
                          const uint32_t* __restrict__ base, const int32_t* __restrict__ lo32, int32_t n_codes,
                          const uint32_t* __restrict__ starts, uint32_t n_starts, const uint32_t* __restrict__ before,
                          int32_t* __restrict__ out) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = code[r];
        if (c < 0 || c >= n_codes) {     // cannot happen for an observed table; keep the kernel memory-safe
            out[r] = 0;
            continue;
        }
        const uint32_t key = __ldg(base + c) + 1u + (uint32_t)(add_sat_s32(start[r], start_adj) - __ldg(lo32 + c));
        out[r] = (int32_t)(upper_bound_u32(starts, n_starts, key) - __ldg(before + c));
    }
}

inline LayoutView view_of(const giqlb_layout* lay) { return LayoutView{lay->d_base, lay->d_lo32, lay->n_codes}; }

// flags -> running count G for rows [lo, hi) of the index; returns the block's cluster count in *total_dev
int running_count(giqlb_ctx* ctx, const giqlb_index* X, int64_t lo, int64_t hi, long long d, uint32_t* G, uint32_t carry = 0,
                  ClusterPredicate pred = ClusterPredicate{nullptr, nullptr, 0}) {
    if (pred.cur) {   // the operand arrays cover the whole index: a block starts a fresh partition, its first row has no predecessor
        pred.cur += lo;
        pred.prev += lo;
    }
    return launch_scan<OpAddU32>(ctx, hi - lo, InClusterFlag{X->gs + lo, X->pmax + lo, view_of(X->layout), d, carry, pred},
                                 OutRunningCount{G + lo});
}

int check_index(giqlb_ctx* ctx, giqlb_index* X, const char* who) {
    if (!X->layout || !X->layout->finalized) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "%s: index has no layout", who);
    GIQLB_TRY(giqlb_index_resolve(ctx, X, nullptr));
    if (X->n_negative > 0)
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "%s: %lld rows have end < start; their grouping depends on the engine's tie order "
                          "in the reference itself and is not supported", who, (long long)X->n_negative);
    return GIQLB_OK;
}

}  // namespace

// G_ready (nullable): the running cluster count of an ungrouped index, already computed by the caller (ownership passes on)
static int cluster_impl(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, const int32_t* cur_key_by_row, const int32_t* prev_key_by_row, int op,
                        int32_t* cluster_id_by_input_row, uint32_t* G_ready = nullptr);

extern "C" int giqlb_cluster(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, int32_t* cluster_id_by_input_row) {
    return cluster_impl(ctx, X, distance, nullptr, nullptr, 0, cluster_id_by_input_row);
}

extern "C" int giqlb_cluster_predicate(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, const int32_t* cur_key_by_row,
                                       const int32_t* prev_key_by_row, int op, int32_t* cluster_id_by_input_row) {
    if (!ctx || !X || op < GIQLB_CMP_EQ || op > GIQLB_CMP_GE || (X->n > 0 && (!cur_key_by_row || !prev_key_by_row)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "cluster_predicate: bad argument");
    return cluster_impl(ctx, X, distance, cur_key_by_row, prev_key_by_row, op, cluster_id_by_input_row);
}

extern "C" int giqlb_cluster_table(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, const int32_t* code, const int32_t* start, int64_t n,
                                   int32_t start_adj, int32_t* cluster_id_by_input_row) {
    if (!ctx || !X || n != X->n || (n > 0 && (!code || !start || !cluster_id_by_input_row)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "cluster_table: bad argument (the index must cover all n rows of the table)");
    // small tables: the direct scatter of the ids is cheap, and the search path would add a round trip
    if (X->grouped_by_strand || !binned_scatter_wanted(n)) return giqlb_cluster(ctx, X, distance, cluster_id_by_input_row);
    begin_call(ctx);
    GIQLB_TRY(check_index(ctx, X, "cluster"));
    if (n == 0) return end_call(ctx);
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, X));
    const long long d = distance > 0 ? (long long)distance : 0;
    uint32_t *G = nullptr, *starts = nullptr, *before = nullptr;
    GIQLB_TRY(dev_alloc(ctx, &G, (size_t)n));
    const LayoutView V = view_of(X->layout);
    int st;
    int64_t n_starts = 0;
    {
        PhaseSpan ps(ctx, GIQLB_PH_CLUSTER);
        st = running_count(ctx, X, 0, n, d, G);
    }
    if (st == GIQLB_OK) {   // how many clusters?  (the round trip also checks the scan's sticky flags)
        cudaMemcpyAsync(ctx->h_scalars + 48, G + (n - 1), 4, cudaMemcpyDeviceToHost, ctx->stream);
        st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
        n_starts = (int64_t)(*reinterpret_cast<const uint32_t*>(ctx->h_scalars + 48));
    }
    if (st == GIQLB_OK && n_starts <= GIQLB_CLUSTER_SEARCH_MAX) {
        st = dev_alloc(ctx, &starts, (size_t)(n_starts > 0 ? n_starts : 1));
        if (st == GIQLB_OK) st = dev_alloc(ctx, &before, (size_t)V.n_codes + 1);
        if (st == GIQLB_OK) {
            PhaseSpan ps(ctx, GIQLB_PH_CLUSTER);
            cluster_starts_kernel<<<grid_of(ctx, n, 4), 256, 0, ctx->stream>>>(X->gs, G, n, starts);
            cluster_before_kernel<<<(V.n_codes + 127) / 128, 128, 0, ctx->stream>>>(V.base, V.n_codes, starts, n_starts, before);
            cluster_ids_search_kernel<<<grid_of(ctx, n, 4), 256, 0, ctx->stream>>>(code, start, n, start_adj, V.base, V.lo32, V.n_codes, starts,
                                                                                 (uint32_t)n_starts, before, cluster_id_by_input_row);
            ctx->launches_this_call += 3;
            ctx->stats.kernels_total += 3;
            if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "cluster_table: launch failed");
        }
        dev_free(ctx, G);
        dev_free(ctx, starts);
        dev_free(ctx, before);
        if (st != GIQLB_OK) return st;
        ctx->stats.rows_in[0] = n;
        ctx->stats.rows_in[1] = 0;
        ctx->stats.rows_out = n;
        ctx->stats.algorithmic_bytes = 12 * n + 4 * n;
        return end_call(ctx);
    }
    if (st != GIQLB_OK) {
        dev_free(ctx, G);
        return st;
    }
    // many clusters: the ids travel back through the sorted row ids (the running count is handed over, not recomputed)
    return cluster_impl(ctx, X, distance, nullptr, nullptr, 0, cluster_id_by_input_row, G);
}

static int cluster_impl(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, const int32_t* cur_key_by_row, const int32_t* prev_key_by_row, int op,
                        int32_t* cluster_id_by_input_row, uint32_t* G_ready) {
    if (!ctx || !X || (X->n > 0 && !cluster_id_by_input_row)) {
        dev_free(ctx, G_ready);
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "cluster: bad argument");
    }
    begin_call(ctx);
    const int64_t n = X->n;
    {
        int pre = check_index(ctx, X, "cluster");
        if (pre == GIQLB_OK) pre = require_row_ids(ctx, X, "cluster");
        if (pre == GIQLB_OK && n > 0) pre = giqlb_index_ensure_pmax(ctx, X);
        if (pre != GIQLB_OK || n == 0) {
            dev_free(ctx, G_ready);
            return pre != GIQLB_OK ? pre : end_call(ctx);
        }
    }
    const long long d = distance > 0 ? (long long)distance : 0;   // cluster.py:266-272: the term is added only when > 0
    uint32_t *G = G_ready, *off = nullptr;
    int32_t *kc = nullptr, *kp = nullptr;
    if (!G) GIQLB_TRY(dev_alloc(ctx, &G, (size_t)n));
    int st = dev_alloc(ctx, &off, (size_t)X->layout->n_codes);
    if (st == GIQLB_OK && cur_key_by_row) st = dev_alloc(ctx, &kc, (size_t)n);
    if (st == GIQLB_OK && cur_key_by_row) st = dev_alloc(ctx, &kp, (size_t)n);
    const int nb = X->grouped_by_strand ? X->n_blocks : 1;
    const LayoutView V = view_of(X->layout);
    if (st == GIQLB_OK) {
        PhaseSpan ps(ctx, GIQLB_PH_CLUSTER);
        if (cur_key_by_row) {   // operand codes into sorted-row order
            gather_i32_kernel<<<grid_of(ctx, n, 4), 256, 0, ctx->stream>>>(cur_key_by_row, X->rid, n, kc);
            gather_i32_kernel<<<grid_of(ctx, n, 4), 256, 0, ctx->stream>>>(prev_key_by_row, X->rid, n, kp);
            ctx->launches_this_call += 2;
            ctx->stats.kernels_total += 2;
        }
        for (int b = 0; b < nb && st == GIQLB_OK; b++) {
            const int64_t lo = X->grouped_by_strand ? X->block_off[b] : 0, hi = X->grouped_by_strand ? X->block_off[b + 1] : n;
            if (hi <= lo) continue;
            if (!G_ready) st = running_count(ctx, X, lo, hi, d, G, 0, ClusterPredicate{kc, kp, op});
            if (st != GIQLB_OK) break;
            cluster_chrom_offsets_kernel<<<(V.n_codes + 127) / 128, 128, 0, ctx->stream>>>(X->gs + lo, G + lo, hi - lo, V, off);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            if (binned_scatter_wanted(hi - lo)) {
                // ids in sorted-row order first (sequential writes, into their own array: G[j] of a later peer is still read), then
                // the partitioned scatter to input-row order
                uint32_t* ids = nullptr;
                st = dev_alloc(ctx, &ids, (size_t)(hi - lo));
                if (st == GIQLB_OK) {
                    cluster_ids_kernel<<<grid_of(ctx, hi - lo, 4), 256, 0, ctx->stream>>>(X->gs + lo, nullptr, G + lo, off, hi - lo, V,
                                                                                         reinterpret_cast<int32_t*>(ids));
                    ctx->launches_this_call++;
                    ctx->stats.kernels_total++;
                    st = scatter_by_row(ctx, X->rid + lo, ids, hi - lo, cluster_id_by_input_row);
                }
                dev_free(ctx, ids);
            } else {
                cluster_ids_kernel<<<grid_of(ctx, hi - lo, 4), 256, 0, ctx->stream>>>(X->gs + lo, X->rid + lo, G + lo, off, hi - lo, V,
                                                                                     cluster_id_by_input_row);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
            if (st == GIQLB_OK && cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "cluster: launch failed");
        }
    }
    dev_free(ctx, G);
    dev_free(ctx, off);
    dev_free(ctx, kc);
    dev_free(ctx, kp);
    if (st != GIQLB_OK) return st;
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = n;
    ctx->stats.algorithmic_bytes = 12 * n + (X->strand ? n : 0) + 4 * n + (cur_key_by_row ? 16 * n : 0);
    return end_call(ctx);
}

extern "C" int giqlb_merge(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, int32_t* out_code, int32_t* out_start,
                           int32_t* out_end, uint8_t* out_strand, int64_t* out_count, int64_t capacity, int64_t* n_out_host) {
    if (!ctx || !X || !n_out_host || capacity < 0 || (capacity > 0 && (!out_code || !out_start || !out_end)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "merge: bad argument");
    begin_call(ctx);
    GIQLB_TRY(check_index(ctx, X, "merge"));
    *n_out_host = 0;
    const int64_t n = X->n;
    if (n == 0) return end_call(ctx);
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, X));
    const long long d = distance > 0 ? (long long)distance : 0;
    uint32_t* G = nullptr;
    GIQLB_TRY(dev_alloc(ctx, &G, (size_t)n));
    const int nb = X->grouped_by_strand ? X->n_blocks : 1;
    const LayoutView V = view_of(X->layout);
    int st = GIQLB_OK;
    // running counts of every strand block, then their totals in one round trip
    std::vector<int64_t> blk_lo, blk_hi;
    std::vector<int> blk_code;
    {
        PhaseSpan ps(ctx, GIQLB_PH_CLUSTER);
        for (int b = 0; b < nb && st == GIQLB_OK; b++) {
            const int64_t lo = X->grouped_by_strand ? X->block_off[b] : 0, hi = X->grouped_by_strand ? X->block_off[b + 1] : n;
            if (hi <= lo) continue;
            st = running_count(ctx, X, lo, hi, d, G);
            blk_lo.push_back(lo);
            blk_hi.push_back(hi);
            blk_code.push_back(b);
        }
    }
    std::vector<uint32_t> totals(blk_lo.size(), 0u);
    for (size_t k = 0; k < blk_lo.size() && st == GIQLB_OK; k++)
        if (cudaMemcpyAsync(&totals[k], G + blk_hi[k] - 1, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "merge: copy failed");
    // the synchronisation also brings the sticky kernel-side flags back (a look-back that hit its spin limit must not
    // pass silently as wrong region counts)
    if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
    int64_t total = 0;
    for (uint32_t t : totals) total += t;
    *n_out_host = total;
    if (st == GIQLB_OK && total > capacity) {
        dev_free(ctx, G);
        return giqlb_fail(ctx, GIQLB_ERR_CAPACITY, "merge: %lld regions exceed capacity %lld", (long long)total, (long long)capacity);
    }
    if (st == GIQLB_OK && total > 0) {
        PhaseSpan ps(ctx, GIQLB_PH_MERGE);
        if (out_count) cudaMemsetAsync(out_count, 0, (size_t)total * 8, ctx->stream);
        int64_t out_base = 0;
        for (size_t k = 0; k < blk_lo.size(); k++) {
            const int64_t lo = blk_lo[k], hi = blk_hi[k];
            merge_emit_kernel<<<grid_of(ctx, hi - lo, 4), 256, 0, ctx->stream>>>(X->gs + lo, X->pmax + lo, G + lo, hi - lo, V, out_base, lo,
                                                                                out_code, out_start, out_end, (long long*)out_count);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            if (out_strand && totals[k] > 0) {
                // every row of a strand block carries the block's strand code
                const uint8_t code = (uint8_t)blk_code[k];
                fill_u8_kernel<<<grid_of(ctx, totals[k], 4), 256, 0, ctx->stream>>>(out_strand + out_base, code, totals[k]);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
            out_base += totals[k];
        }
        if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "merge: launch failed");
    }
    dev_free(ctx, G);
    if (st != GIQLB_OK) return st;
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = total;
    ctx->stats.algorithmic_bytes = 12 * n + (X->strand ? n : 0) + 12 * total;
    return end_call(ctx);
}

// ---- genome-range shards of CLUSTER / MERGE -------------------------------------------------------
// A shard owns the rows whose START lies in its genome range.  What it needs from the shards before it is one
// number — the running max of end over all earlier rows (`carry`, a linear coordinate; the guard cells make it
// chromosome-safe) — and, for cluster ids, how many clusters earlier shards started on each chromosome.
static int shard_index_ok(giqlb_ctx* ctx, giqlb_index* X, const char* who) {
    GIQLB_TRY(check_index(ctx, X, who));
    if (X->grouped_by_strand) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "%s: strand-grouped indexes are not sharded", who);
    return GIQLB_OK;
}

extern "C" int giqlb_shard_max_end(giqlb_ctx* ctx, giqlb_index* X, uint32_t* max_end_host) {
    if (!ctx || !X || !max_end_host) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "shard_max_end: bad argument");
    begin_call(ctx);
    *max_end_host = 0;
    if (X->n == 0) return end_call(ctx);
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, X));
    GIQLB_TRY(fetch_scalars(ctx, X->pmax + X->n - 1, 1) == GIQLB_OK ? GIQLB_OK : GIQLB_ERR_CUDA);
    *max_end_host = (uint32_t)(ctx->h_scalars[0] & 0xffffffffll);
    return end_call(ctx);
}

extern "C" int giqlb_cluster_shard(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, uint32_t carry, const uint32_t* id_offset_by_code_host,
                                   int32_t* cluster_id_by_input_row, uint32_t* counts_by_code_host) {
    if (!ctx || !X) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "cluster_shard: bad argument");
    begin_call(ctx);
    GIQLB_TRY(shard_index_ok(ctx, X, "cluster_shard"));
    if (cluster_id_by_input_row) GIQLB_TRY(require_row_ids(ctx, X, "cluster_shard"));
    const int32_t nc = X->layout->n_codes;
    if (counts_by_code_host) memset(counts_by_code_host, 0, (size_t)nc * 4);
    const int64_t n = X->n;
    if (n == 0) return end_call(ctx);
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, X));
    const long long d = distance > 0 ? (long long)distance : 0;
    const LayoutView V = view_of(X->layout);
    uint32_t *G = nullptr, *off = nullptr, *aux = nullptr;
    // the running count of the previous call on this index (same distance, same carry) is reused: counts first, ids second
    const bool reuse = X->shard_G && X->shard_G_distance == d && X->shard_G_carry == carry;
    int st = GIQLB_OK;
    if (reuse) {
        G = X->shard_G;
    } else {
        dev_free(ctx, X->shard_G);
        X->shard_G = nullptr;
        st = dev_alloc(ctx, &G, (size_t)n);
    }
    if (st == GIQLB_OK) st = dev_alloc(ctx, &off, (size_t)nc);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &aux, (size_t)nc);
    if (st == GIQLB_OK) {
        PhaseSpan ps(ctx, GIQLB_PH_CLUSTER);
        if (!reuse) st = running_count(ctx, X, 0, n, d, G, carry);
        if (st == GIQLB_OK && counts_by_code_host) {
            cluster_chrom_counts_kernel<<<(nc + 127) / 128, 128, 0, ctx->stream>>>(X->gs, G, n, V, aux);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            if (cudaMemcpyAsync(counts_by_code_host, aux, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "cluster_shard: count copy failed");
        }
        if (st == GIQLB_OK && cluster_id_by_input_row) {
            if (!id_offset_by_code_host) st = giqlb_fail(ctx, GIQLB_ERR_INVALID, "cluster_shard: ids need the per-chromosome offsets");
            if (st == GIQLB_OK && cudaMemcpyAsync(aux, id_offset_by_code_host, (size_t)nc * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
                st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "cluster_shard: offset copy failed");
            if (st == GIQLB_OK) {
                cluster_chrom_offsets_kernel<<<(nc + 127) / 128, 128, 0, ctx->stream>>>(X->gs, G, n, V, off);
                if (binned_scatter_wanted(n)) {
                    uint32_t* ids = nullptr;
                    st = dev_alloc(ctx, &ids, (size_t)n);
                    if (st == GIQLB_OK) {
                        cluster_ids_offset_kernel<<<grid_of(ctx, n, 4), 256, 0, ctx->stream>>>(X->gs, nullptr, G, off, aux, n, V, reinterpret_cast<int32_t*>(ids));
                        st = scatter_by_row(ctx, X->rid, ids, n, cluster_id_by_input_row);
                    }
                    dev_free(ctx, ids);
                } else {
                    cluster_ids_offset_kernel<<<grid_of(ctx, n, 4), 256, 0, ctx->stream>>>(X->gs, X->rid, G, off, aux, n, V, cluster_id_by_input_row);
                }
                ctx->launches_this_call += 2;
                ctx->stats.kernels_total += 2;
                // id_offset_by_code_host is pageable caller memory: wait before returning
                if (st == GIQLB_OK && (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess))
                    st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "cluster_shard: launch failed");
            }
        }
    }
    if (st == GIQLB_OK && !cluster_id_by_input_row) {   // a counts-only call: the ids call usually follows
        X->shard_G = G;
        X->shard_G_distance = d;
        X->shard_G_carry = carry;
    } else {
        dev_free(ctx, G);
        X->shard_G = nullptr;
    }
    dev_free(ctx, off);
    dev_free(ctx, aux);
    return st == GIQLB_OK ? end_call(ctx) : st;
}

extern "C" int giqlb_merge_shard(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, uint32_t carry, int32_t* out_code, int32_t* out_start,
                                 int32_t* out_end, int64_t* out_count, int64_t capacity, int64_t* n_out_host, int64_t* head_host) {
    if (!ctx || !X || !n_out_host || !head_host || capacity < 0) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "merge_shard: bad argument");
    begin_call(ctx);
    GIQLB_TRY(shard_index_ok(ctx, X, "merge_shard"));
    *n_out_host = 0;
    head_host[0] = 0; head_host[1] = -1; head_host[2] = 0;
    const int64_t n = X->n;
    if (n == 0) return end_call(ctx);
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, X));
    const long long d = distance > 0 ? (long long)distance : 0;
    const LayoutView V = view_of(X->layout);
    uint32_t* G = nullptr;
    long long* head = reinterpret_cast<long long*>(ctx->d_scalars + 12);   // [3] head rows, code, end; then [15] region count
    int st = dev_alloc(ctx, &G, (size_t)n);
    if (st == GIQLB_OK) {
        PhaseSpan ps(ctx, GIQLB_PH_CLUSTER);
        st = running_count(ctx, X, 0, n, d, G, carry);
        if (st == GIQLB_OK) {
            merge_head_kernel<<<1, 32, 0, ctx->stream>>>(X->gs, X->pmax, G, n, V, head);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            cudaMemcpyAsync(ctx->h_scalars + 4, G + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream);
            st = fetch_scalars_checked(ctx, head, 3);
        }
    }
    int64_t total = 0;
    if (st == GIQLB_OK) {
        head_host[0] = ctx->h_scalars[0];   // rows of the head fragment
        head_host[1] = ctx->h_scalars[1];   // its chromosome code (-1: no head)
        head_host[2] = ctx->h_scalars[2];   // its max end, table coordinates
        total = (int64_t)(ctx->h_scalars[4] & 0xffffffffll);
        *n_out_host = total;
        if (total > capacity) st = giqlb_fail(ctx, GIQLB_ERR_CAPACITY, "merge_shard: %lld regions exceed capacity %lld", (long long)total, (long long)capacity);
        else if (total > 0 && (!out_code || !out_start || !out_end)) st = giqlb_fail(ctx, GIQLB_ERR_INVALID, "merge_shard: null output");
    }
    if (st == GIQLB_OK && total > 0) {
        PhaseSpan ps(ctx, GIQLB_PH_MERGE);
        if (out_count) cudaMemsetAsync(out_count, 0, (size_t)total * 8, ctx->stream);
        merge_emit_kernel<<<grid_of(ctx, n, 4), 256, 0, ctx->stream>>>(X->gs, X->pmax, G, n, V, 0, 0, out_code, out_start, out_end, (long long*)out_count);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "merge_shard: launch failed");
    }
    dev_free(ctx, G);
    return st == GIQLB_OK ? end_call(ctx) : st;
}
