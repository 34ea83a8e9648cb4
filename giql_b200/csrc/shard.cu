// shard.cu — genome-range partitioning of a join across the GPUs of one box.
//
// The pair predicate never crosses chromosomes (abdenlab/giql src/giql/expanders/intersects.py:151) and the
// DuckDB plan already evaluates it per chromosome (src/giql/expanders/intersects_duckdb.py:1283-1299), so the
// join shards by contiguous ranges of the linearised genome.  Every rank holds both tables; it
//   1. histograms the LEFT table's linearised starts in 2^16-cell bins and cuts the genome into `world`
//      ranges holding equal numbers of left rows (giqlb_shard_cuts — identical on every rank, no exchange),
//   2. selects the left rows whose START lies in its range — each left row has exactly one owner, so no
//      pair is ever produced twice — and the right rows that can overlap that range; right rows straddling
//      a cut are selected by both neighbours (giqlb_shard_select),
//   3. builds its indexes over the selections (giqlb_index_build_rows) and joins them.
// Only the per-rank pair counts (one int64 each) and, on request, the pairs themselves cross NVLink.
#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int SH_BIN_SHIFT = 16;
constexpr int SH_THREADS = 256;

// exact linearisation of a row of an OBSERVED table (no clamping); false when the row lies outside the layout
__device__ __forceinline__ bool row_linearise(const ProbeCols& P, int64_t i, uint32_t& gs, uint32_t& ge) {
    int32_t c = P.code[i];
    if (c < 0 || c >= P.n_codes) return false;
    long long lo = __ldg(P.lo32 + c), hi = __ldg(P.hi32 + c);
    long long s = (long long)P.start[i] + P.sadj, e = (long long)P.end[i] + P.eadj;
    if (s < lo || s > hi || e < lo || e > hi) return false;
    uint32_t b = __ldg(P.base + c) + 1u;
    gs = b + (uint32_t)(s - lo);
    ge = b + (uint32_t)(e - lo);
    return true;
}

__global__ void __launch_bounds__(SH_THREADS)
shard_hist_kernel(ProbeCols P, uint32_t* __restrict__ hist, uint32_t* __restrict__ max_width, int* __restrict__ err) {
    uint32_t mw = 0;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * SH_THREADS + threadIdx.x; i < P.n; i += (int64_t)gridDim.x * SH_THREADS) {
        uint32_t gs, ge;
        if (!row_linearise(P, i, gs, ge)) {
            bad = true;
            continue;
        }
        atomicAdd(&hist[gs >> SH_BIN_SHIFT], 1u);
        if (ge > gs && ge - gs > mw) mw = ge - gs;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mw = max(mw, __shfl_xor_sync(FULL_MASK, mw, d));
    if (lane_id() == 0 && mw) atomicMax(max_width, mw);
    if (bad) err[2] = 1;
}

// one CTA: walks the bins with a running count and records, for g = 1..world-1, the first bin boundary at
// which at least ceil(g*n/world) left rows lie below.
__global__ void __launch_bounds__(1024)
shard_cuts_kernel(const uint32_t* __restrict__ hist, int64_t n_bins, int64_t n, int world,
                  unsigned long long* __restrict__ cuts /* [world + 1] */) {
    __shared__ unsigned long long scratch[1024 / 32 + 2];
    unsigned long long carry = 0;
    for (int64_t b0 = 0; b0 < n_bins; b0 += 1024) {
        int64_t b = b0 + threadIdx.x;
        unsigned long long v = b < n_bins ? hist[b] : 0ull, tot;
        unsigned long long excl = carry + block_excl_scan_add<unsigned long long>(v, scratch, &tot);
        unsigned long long incl = excl + v;
        if (v) {
            // targets g with excl < ceil(g*n/world) <= incl
            long long g_lo = (long long)((excl * (unsigned long long)world) / (unsigned long long)n);   // floor
            for (long long g = g_lo > 1 ? g_lo : 1; g < world; g++) {
                unsigned long long target = ((unsigned long long)g * (unsigned long long)n + world - 1) / world;
                if (target > incl) break;
                if (target > excl) cuts[g] = (unsigned long long)(b + 1) << SH_BIN_SHIFT;
            }
        }
        carry += tot;
        __syncthreads();
    }
}

struct InShardPred {
    ProbeCols P;
    uint32_t gs_lo, gs_hi, ge_gt;
    __device__ unsigned long long operator()(int64_t i) const {
        uint32_t gs, ge;
        if (!row_linearise(P, i, gs, ge)) return 0ull;
        return (gs >= gs_lo && gs < gs_hi && ge > ge_gt) ? 1ull : 0ull;
    }
};
struct OutShardRows {
    int32_t* rows;
    unsigned long long* total;
    int64_t n;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        if (v) rows[excl] = (int32_t)i;
        if (i == n - 1) *total = excl + v;
    }
};

}  // namespace

extern "C" int giqlb_shard_cuts(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, int world,
                                uint32_t* cuts_host, uint32_t* max_width_host) {
    if (!ctx || !cuts_host || world < 1 || world > 4096) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "shard_cuts: bad argument");
    begin_call(ctx);
    ProbeCols P;
    GIQLB_TRY(probe_cols(ctx, lay, code, start, end, n, start_adj, end_adj, &P));
    const int64_t n_bins = (int64_t)(lay->total_extent >> SH_BIN_SHIFT) + 1;
    uint32_t* hist = nullptr;           // [n_bins] then max_width
    unsigned long long* cuts = nullptr; // [world + 1]
    GIQLB_TRY(dev_alloc(ctx, &hist, (size_t)n_bins + 1));
    int st = dev_alloc(ctx, &cuts, (size_t)world + 1);
    if (st != GIQLB_OK) {
        dev_free(ctx, hist);
        return st;
    }
    std::vector<unsigned long long> h_cuts((size_t)world + 1, 0ull);
    uint32_t h_mw = 0;
    int h_err[4] = {0, 0, 0, 0};
    cudaMemsetAsync(hist, 0, ((size_t)n_bins + 1) * 4, ctx->stream);
    cudaMemsetAsync(cuts, 0, ((size_t)world + 1) * 8, ctx->stream);
    if (n > 0) {
        PhaseSpan ps(ctx, GIQLB_PH_SHARD);
        int64_t blocks = ceil_div(n, SH_THREADS * 8), cap = (int64_t)ctx->sm_count * 8;
        shard_hist_kernel<<<(unsigned)(blocks < cap ? blocks : cap), SH_THREADS, 0, ctx->stream>>>(P, hist, hist + n_bins, ctx->d_err);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        shard_cuts_kernel<<<1, 1024, 0, ctx->stream>>>(hist, n_bins, n, world, cuts);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
    }
    cudaMemcpyAsync(h_cuts.data(), cuts, ((size_t)world + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&h_mw, hist + n_bins, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(h_err, ctx->d_err, sizeof(h_err), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, hist);
    dev_free(ctx, cuts);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) return giqlb_fail(ctx, GIQLB_ERR_CUDA, "shard_cuts: %s", cudaGetErrorString(e));
    if (h_err[2]) {
        cudaMemsetAsync(ctx->d_err + 2, 0, sizeof(int), ctx->stream);
        return giqlb_fail(ctx, GIQLB_ERR_RANGE, "shard_cuts: a row lies outside the layout (observe the table before finalize)");
    }
    // unset cuts (empty tail ranges) inherit the next boundary; monotone by construction
    cuts_host[0] = 0u;
    cuts_host[world] = 0xffffffffu;
    for (int g = world - 1; g >= 1; g--) {
        unsigned long long c = h_cuts[g];
        if (c == 0 || c > 0xffffffffull) c = cuts_host[g + 1];
        cuts_host[g] = (uint32_t)c;
    }
    for (int g = 1; g < world; g++)
        if (cuts_host[g] < cuts_host[g - 1]) cuts_host[g] = cuts_host[g - 1];
    if (max_width_host) *max_width_host = h_mw;
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = world;
    ctx->stats.algorithmic_bytes = 12 * n;
    return end_call(ctx);
}

extern "C" int giqlb_shard_select(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                  const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, uint32_t gs_lo,
                                  uint32_t gs_hi, uint32_t ge_gt, int32_t* out_rows, int64_t* n_out_host) {
    using namespace giqlb_scan;
    if (!ctx || !n_out_host || (n > 0 && !out_rows)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "shard_select: bad argument");
    begin_call(ctx);
    ProbeCols P;
    GIQLB_TRY(probe_cols(ctx, lay, code, start, end, n, start_adj, end_adj, &P));
    unsigned long long* total = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 9);
    CUDA_TRY(ctx, cudaMemsetAsync(total, 0, 8, ctx->stream));
    {
        PhaseSpan ps(ctx, GIQLB_PH_SHARD);
        GIQLB_TRY((launch_scan<OpAddU64>(ctx, n, InShardPred{P, gs_lo, gs_hi, ge_gt}, OutShardRows{out_rows, total, n})));
    }
    GIQLB_TRY(fetch_scalars(ctx, total, 1));
    *n_out_host = ctx->h_scalars[0];
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = *n_out_host;
    ctx->stats.algorithmic_bytes = 12 * n + 4 * *n_out_host;
    return end_call(ctx);
}
