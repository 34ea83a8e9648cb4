// shard.cu — genome-range partitioning of a join across the GPUs of one box.
//
// The pair predicate never crosses chromosomes (abdenlab/giql src/giql/expanders/intersects.py:151) and the
// DuckDB plan already evaluates it per chromosome (src/giql/expanders/intersects_duckdb.py:1283-1299), so the
// join shards by contiguous ranges of the linearised genome.  Every rank holds both tables; it
//   1. histograms the LEFT table's linearised starts in 2^19-cell bins and cuts the genome into `world`
//      ranges holding equal numbers of left rows (giqlb_shard_cuts — identical on every rank, no exchange),
//   2. selects the left rows whose START lies in its range — each left row has exactly one owner, so no
//      pair is ever produced twice — and the right rows that can overlap that range; right rows straddling
//      a cut are selected by both neighbours (giqlb_shard_select),
//   3. builds its indexes over the selections (giqlb_index_build_rows) and joins them.
// Only the per-rank pair counts (one int64 each) and, on request, the pairs themselves cross NVLink.
#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int SH_BIN_SHIFT = 19;            // 2^19-cell bins: at most 8192 of them, a shared-memory histogram per CTA
constexpr int SH_MAX_BINS = 8192;
constexpr int SH_THREADS = 256;

// exact linearisation of a row of an OBSERVED table (no clamping); false when the row lies outside the layout
// (32-bit arithmetic as in probe_linearise_values: an offset that would carry the value past INT32_MAX / MIN saturates there,
//  outside every [lo, hi] the layout admits, so such a row is rejected either way)
__device__ __forceinline__ bool vals_linearise(const ProbeCols& P, int32_t c, int32_t s0, int32_t e0, uint32_t& gs, uint32_t& ge) {
    if (c < 0 || c >= P.n_codes) return false;
    const int32_t lo = __ldg(P.lo32 + c), hi = __ldg(P.hi32 + c);
    const int32_t s = add_sat_s32(s0, P.sadj), e = add_sat_s32(e0, P.eadj);
    if (s < lo || s > hi || e < lo || e > hi) return false;
    const uint32_t b = __ldg(P.base + c) + 1u;
    gs = b + (uint32_t)(s - lo);
    ge = b + (uint32_t)(e - lo);
    return true;
}
__device__ __forceinline__ bool row_linearise(const ProbeCols& P, int64_t i, uint32_t& gs, uint32_t& ge) {
    return vals_linearise(P, P.code[i], P.start[i], P.end[i], gs, ge);
}

// `stride` > 1 histograms every stride-th row only (a deterministic sample: every rank takes the same rows); the
// widest interval then comes from the layout's observe pass instead of this kernel.

__global__ void __launch_bounds__(SH_THREADS)
shard_hist_kernel(ProbeCols P, int64_t stride, uint32_t* __restrict__ hist, uint32_t* __restrict__ max_width, int* __restrict__ err) {
    __shared__ uint32_t s_hist[SH_MAX_BINS];
    for (int i = threadIdx.x; i < SH_MAX_BINS; i += SH_THREADS) s_hist[i] = 0;
    __syncthreads();
    uint32_t mw = 0;
    bool bad = false;
    for (int64_t k = (int64_t)blockIdx.x * SH_THREADS + threadIdx.x; k * stride < P.n; k += (int64_t)gridDim.x * SH_THREADS) {
        const int64_t i = k * stride;
        uint32_t gs, ge;
        if (!row_linearise(P, i, gs, ge)) {
            bad = true;
            continue;
        }
        atomicAdd(&s_hist[gs >> SH_BIN_SHIFT], 1u);
        if (ge > gs && ge - gs > mw) mw = ge - gs;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mw = max(mw, __shfl_xor_sync(FULL_MASK, mw, d));
    if (lane_id() == 0 && mw) atomicMax(max_width, mw);
    if (bad) err[2] = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < SH_MAX_BINS; i += SH_THREADS)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// one CTA: every thread owns a contiguous chunk of bins (serial sums, ONE block scan) and records, for
// g = 1..world-1, the first bin boundary at which at least ceil(g*n/world) left rows lie below.
__global__ void __launch_bounds__(1024)
shard_cuts_kernel(const uint32_t* __restrict__ hist, int64_t n_bins, int64_t n, int world,
                  unsigned long long* __restrict__ cuts /* [world + 1] */) {
    __shared__ unsigned long long scratch[1024 / 32 + 2];
    const int64_t per = (n_bins + 1023) / 1024;
    const int64_t lo = min((int64_t)threadIdx.x * per, n_bins), hi = min(lo + per, n_bins);
    unsigned long long sum = 0;
    for (int64_t b = lo; b < hi; b++) sum += hist[b];
    unsigned long long tot;
    unsigned long long excl = block_excl_scan_add<unsigned long long>(sum, scratch, &tot);
    for (int64_t b = lo; b < hi; b++) {
        const unsigned long long v = hist[b];
        if (v) {
            const unsigned long long incl = excl + v;
            // targets g with excl < ceil(g*n/world) <= incl
            long long g_lo = (long long)((excl * (unsigned long long)world) / (unsigned long long)n);   // floor
            for (long long g = g_lo > 1 ? g_lo : 1; g < world; g++) {
                const unsigned long long target = ((unsigned long long)g * (unsigned long long)n + world - 1) / world;
                if (target > incl) break;
                if (target > excl) cuts[g] = (unsigned long long)(b + 1) << SH_BIN_SHIFT;
            }
            excl = incl;
        }
    }
}

// Selection without a scan: a CTA takes 1024 rows, scans its predicate counts in shared memory and claims its
// slice of the output with ONE atomic; ids are ascending inside a slice, slices land in arbitrary order (the
// selection is sorted by the index build anyway).
__global__ void __launch_bounds__(SH_THREADS)
shard_select_kernel(ProbeCols P, uint32_t gs_lo, uint32_t gs_hi, uint32_t ge_gt, int32_t* __restrict__ out_rows,
                    unsigned long long* __restrict__ total) {
    __shared__ unsigned int scratch[SH_THREADS / 32 + 2];
    __shared__ unsigned long long s_base;
    for (int64_t base = (int64_t)blockIdx.x * (SH_THREADS * 4); base < P.n; base += (int64_t)gridDim.x * (SH_THREADS * 4)) {
        // thread t owns rows base + 4t .. base + 4t + 3 (ascending within the slice)
        bool keep[4];
        unsigned int cnt = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t i = base + (int64_t)threadIdx.x * 4 + k;
            uint32_t gs, ge;
            keep[k] = i < P.n && row_linearise(P, i, gs, ge) && gs >= gs_lo && gs < gs_hi && ge > ge_gt;
            cnt += keep[k];
        }
        unsigned int tot;
        unsigned int ex = block_excl_scan_add<unsigned int>(cnt, scratch, &tot);
        if (threadIdx.x == 0 && tot) s_base = atomicAdd(total, (unsigned long long)tot);
        __syncthreads();
        if (tot) {
            unsigned long long o = s_base + ex;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (keep[k]) out_rows[o++] = (int32_t)(base + (int64_t)threadIdx.x * 4 + k);
        }
        __syncthreads();   // scratch / s_base are reused by the next slice
    }
}

// Device-side finish of the cuts (what giqlb_shard_cuts does on the host): u32 cuts[world + 1] with cuts[0] = 0,
// cuts[world] = 2^32 - 1, unset cuts inherit the next boundary, monotone; cuts32[world + 1] = widest left interval.
__global__ void shard_cuts_finish_kernel(const unsigned long long* __restrict__ cuts64, const uint32_t* __restrict__ max_width,
                                         const unsigned long long* __restrict__ observed_width /* nullable: from layout observe */, int world,
                                         uint32_t* __restrict__ cuts32) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    cuts32[0] = 0u;
    cuts32[world] = 0xffffffffu;
    for (int g = world - 1; g >= 1; g--) {
        unsigned long long c = cuts64[g];
        if (c == 0 || c > 0xffffffffull) c = cuts32[g + 1];
        cuts32[g] = (uint32_t)c;
    }
    for (int g = 1; g < world; g++)
        if (cuts32[g] < cuts32[g - 1]) cuts32[g] = cuts32[g - 1];
    unsigned long long w = observed_width ? *observed_width : (unsigned long long)*max_width;
    cuts32[world + 1] = w > 0xffffffffull ? 0xffffffffu : (uint32_t)w;
}

// Fused selection + key generation for one rank: reads the table once, keeps the rows of the rank's genome range
// (side 0: owner-of-left-start; side 1: right rows that can overlap the range) and writes their linear keys and row
// ids compacted (CTA-aggregated: one atomic per 1024 rows).  The range comes from DEVICE memory, so the host never
// waits for the cuts.
__global__ void __launch_bounds__(SH_THREADS)
shard_select_keygen_kernel(ProbeCols P, const uint32_t* __restrict__ cuts32, int world, int rank, int side,
                           uint32_t* __restrict__ out_gs, uint32_t* __restrict__ out_ge, int32_t* __restrict__ out_rows,
                           unsigned long long* __restrict__ counters /* [0] end<=start [1] end<start [2] outside [3] selected */,
                           int* __restrict__ err) {
    __shared__ unsigned int scratch[SH_THREADS / 32 + 2];
    __shared__ unsigned long long s_base;
    const uint32_t lo = cuts32[rank], hi = cuts32[rank + 1], maxw = cuts32[world + 1];
    const uint32_t gs_lo = side == 0 ? lo : 0u;
    const uint32_t gs_hi = side == 0 ? hi : (hi > 0xffffffffu - maxw ? 0xffffffffu : hi + maxw);
    const uint32_t ge_gt = side == 0 ? 0u : lo;
    unsigned int degenerate = 0, negative = 0, outside = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(P.code) | reinterpret_cast<uintptr_t>(P.start) | reinterpret_cast<uintptr_t>(P.end)) & 15) == 0;
    for (int64_t base = (int64_t)blockIdx.x * (SH_THREADS * 4); base < P.n; base += (int64_t)gridDim.x * (SH_THREADS * 4)) {
        bool keep[4];
        uint32_t gs[4], ge[4];
        unsigned int cnt = 0;
        const int64_t i0 = base + (int64_t)threadIdx.x * 4;
        int32_t c4[4], s4[4], e4[4];
        if (vec && i0 + 3 < P.n) {   // one 16-byte load per column
            const int4 c = __ldg(reinterpret_cast<const int4*>(P.code + i0));
            const int4 s = __ldg(reinterpret_cast<const int4*>(P.start + i0));
            const int4 e = __ldg(reinterpret_cast<const int4*>(P.end + i0));
            c4[0] = c.x; c4[1] = c.y; c4[2] = c.z; c4[3] = c.w;
            s4[0] = s.x; s4[1] = s.y; s4[2] = s.z; s4[3] = s.w;
            e4[0] = e.x; e4[1] = e.y; e4[2] = e.z; e4[3] = e.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool in = i0 + k < P.n;
                c4[k] = in ? P.code[i0 + k] : -1;
                s4[k] = in ? P.start[i0 + k] : 0;
                e4[k] = in ? P.end[i0 + k] : 0;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            keep[k] = false;
            if (i0 + k < P.n) {
                if (!vals_linearise(P, c4[k], s4[k], e4[k], gs[k], ge[k])) outside++;
                else keep[k] = gs[k] >= gs_lo && gs[k] < gs_hi && ge[k] > ge_gt;
            }
            if (keep[k]) {
                degenerate += ge[k] <= gs[k];
                negative += ge[k] < gs[k];
            }
            cnt += keep[k];
        }
        unsigned int tot;
        unsigned int ex = block_excl_scan_add<unsigned int>(cnt, scratch, &tot);
        if (threadIdx.x == 0 && tot) s_base = atomicAdd(&counters[3], (unsigned long long)tot);
        __syncthreads();
        if (tot) {
            unsigned long long o = s_base + ex;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (keep[k]) {
                    out_gs[o] = gs[k];
                    out_ge[o] = ge[k];
                    out_rows[o] = (int32_t)(base + (int64_t)threadIdx.x * 4 + k);
                    o++;
                }
        }
        __syncthreads();
    }
    degenerate = warp_reduce_add(degenerate);
    negative = warp_reduce_add(negative);
    outside = warp_reduce_add(outside);
    if (lane_id() == 0) {
        if (degenerate) atomicAdd(&counters[0], (unsigned long long)degenerate);
        if (negative) atomicAdd(&counters[1], (unsigned long long)negative);
        if (outside) {
            atomicAdd(&counters[2], (unsigned long long)outside);
            err[3] = 1;
        }
    }
}

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
        // a caller that does not ask for the widest interval (CLUSTER / MERGE shards replicate nothing) gets its cuts from a
        // 1-in-16 sample of the rows, like the one-call sharded join: the same cuts on every rank, balance error ~ 1 / sqrt(n / 16)
        const int64_t stride = !max_width_host && n >= (1 << 20) ? 16 : 1, n_samp = ceil_div(n, stride);
        int64_t blocks = ceil_div(n_samp, SH_THREADS * 8), cap = (int64_t)ctx->sm_count * 6;
        shard_hist_kernel<<<(unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap)), SH_THREADS, 0, ctx->stream>>>(P, stride, hist, hist + n_bins, ctx->d_err);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        shard_cuts_kernel<<<1, 1024, 0, ctx->stream>>>(hist, n_bins, n_samp, world, cuts);
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
    if (!ctx || !n_out_host || (n > 0 && !out_rows)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "shard_select: bad argument");
    begin_call(ctx);
    ProbeCols P;
    GIQLB_TRY(probe_cols(ctx, lay, code, start, end, n, start_adj, end_adj, &P));
    unsigned long long* total = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 9);
    CUDA_TRY(ctx, cudaMemsetAsync(total, 0, 8, ctx->stream));
    if (n > 0) {
        PhaseSpan ps(ctx, GIQLB_PH_SHARD);
        int64_t blocks = ceil_div(n, SH_THREADS * 4), cap = (int64_t)ctx->sm_count * 8;
        shard_select_kernel<<<(unsigned)(blocks < cap ? blocks : cap), SH_THREADS, 0, ctx->stream>>>(P, gs_lo, gs_hi, ge_gt, out_rows, total);
        LAUNCH_CHECK(ctx);
    }
    GIQLB_TRY(fetch_scalars(ctx, total, 1));
    *n_out_host = ctx->h_scalars[0];
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = *n_out_host;
    ctx->stats.algorithmic_bytes = 12 * n + 4 * *n_out_host;
    return end_call(ctx);
}

// Index of the rows whose linear start lies in [gs_lo, gs_hi) (and whose end lies above ge_gt), in one go: the fused select +
// keygen pass (the table is read once, keys and row ids come out compacted), one round trip for the selection size, then
// the sort.  What giqlb_shard_select + giqlb_index_build_rows do in three passes over the rows (select, keygen through
// the row ids, remap of the sorted positions to row ids).
struct RangeCuts { uint32_t v[3]; };
__global__ void range_cuts_kernel(RangeCuts c, uint32_t* __restrict__ out) {
    if (threadIdx.x < 3) out[threadIdx.x] = c.v[threadIdx.x];
}

extern "C" int giqlb_index_build_range(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                       const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, uint32_t gs_lo,
                                       uint32_t gs_hi, int flags, giqlb_index** out) {
    if (!ctx || !out) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "index_build_range: bad argument");
    *out = nullptr;
    begin_call(ctx);
    ProbeCols P;
    GIQLB_TRY(probe_cols(ctx, lay, code, start, end, n, start_adj, end_adj, &P));
    giqlb_prekeyed pre;
    uint32_t* cuts32 = nullptr;
    int st = dev_alloc(ctx, &cuts32, 4);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &pre.gs, (size_t)(n > 0 ? n : 1));
    if (st == GIQLB_OK) st = dev_alloc(ctx, &pre.ge, (size_t)(n > 0 ? n : 1));
    if (st == GIQLB_OK) st = dev_alloc(ctx, &pre.rows, (size_t)(n > 0 ? n : 1));
    if (st == GIQLB_OK) st = dev_alloc(ctx, &pre.counters, 4 + 512);
    if (st == GIQLB_OK && cudaMemsetAsync(pre.counters, 0, (4 + 512) * 8, ctx->stream) != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "index_build_range: memset failed");
    int64_t n_sel = 0, deg = 0, neg = 0;
    if (st == GIQLB_OK) {
        PhaseSpan ps(ctx, GIQLB_PH_SHARD);
        RangeCuts rc;
        rc.v[0] = gs_lo; rc.v[1] = gs_hi; rc.v[2] = 0u;       // world = 1, rank = 0, side 0: starts in [lo, hi), no width bound needed
        range_cuts_kernel<<<1, 32, 0, ctx->stream>>>(rc, cuts32);
        if (n > 0) {
            const int64_t blocks = ceil_div(n, SH_THREADS * 4), cap = (int64_t)ctx->sm_count * 8;
            shard_select_keygen_kernel<<<(unsigned)(blocks < cap ? blocks : cap), SH_THREADS, 0, ctx->stream>>>(
                P, cuts32, 1, 0, 0, pre.gs, pre.ge, pre.rows, pre.counters, ctx->d_err);
        }
        ctx->launches_this_call += 2;
        ctx->stats.kernels_total += 2;
        if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "index_build_range: launch failed");
    }
    if (st == GIQLB_OK) {
        cudaMemcpyAsync(ctx->h_scalars + 40, pre.counters, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream);
        st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
        deg = ctx->h_scalars[40]; neg = ctx->h_scalars[41]; n_sel = ctx->h_scalars[43];
    }
    if (st == GIQLB_OK) st = giqlb_index_build_impl(ctx, lay, code, start, end, nullptr, pre.rows, n_sel, start_adj, end_adj, flags & 0xff, &pre, out);
    if (st == GIQLB_OK) {
        (*out)->n_degenerate = deg;
        (*out)->n_negative = neg;
        (*out)->resolved = true;
    }
    dev_free(ctx, pre.gs); dev_free(ctx, pre.ge); dev_free(ctx, pre.rows); dev_free(ctx, pre.counters);
    dev_free(ctx, cuts32);
    return st;
}

// ---- the rank-local hot path in one call --------------------------------------------------------------------------
// dictionary encode -> layout -> cuts (stay on the device) -> fused select + keygen of both tables -> ONE round trip
// for the two selection sizes -> sort / index of the selections -> single-pass join.  Host round trips: dictionary
// names, layout extent, selection sizes, pair total.
static int sharded_impl(giqlb_ctx* ctx, giqlb_table_t* L, giqlb_table_t* R, int32_t n_codes, int world, int rank, int32_t* out_l,
                        int32_t* out_r, int64_t capacity, int64_t* total_host, giqlb_dict** dict_out, bool async_dict) {
    const int32_t n_codes_in = n_codes;
    giqlb_dict* dict = nullptr;
    giqlb_dict_pending* pend = nullptr;
    const int32_t* slot2code = nullptr;
    const int* ok_flag = nullptr;
    giqlb_layout* lay = nullptr;
    giqlb_index *il = nullptr, *ir = nullptr;
    giqlb_prekeyed pre[2];
    uint32_t *hist = nullptr, *cuts32 = nullptr;
    unsigned long long* cuts64 = nullptr;
    int st = GIQLB_OK;
    const bool encode = !L->codes_valid;
    if (encode) {
        if (async_dict && giqlb_dict_encode_async(ctx, L->chrom_offsets, L->chrom_bytes, L->n, R->chrom_offsets, R->chrom_bytes, R->n,
                                                  L->codes, R->codes, &pend, true) == GIQLB_OK) {
            n_codes = GIQLB_ASYNC_DICT_MAX_CODES;   // upper bound until the names come back with the selection sizes
            giqlb_dict_pending_tables(pend, &slot2code, &ok_flag);   // the slot -> code translation rides in the observe pass
        } else {
            pend = nullptr;
            st = giqlb_dict_encode(ctx, L->chrom_offsets, L->chrom_bytes, L->n, R->chrom_offsets, R->chrom_bytes, R->n, L->codes, R->codes, &dict);
            if (st != GIQLB_OK) return st;
            n_codes = giqlb_dict_size(dict);
        }
    }
    if (st == GIQLB_OK) st = giqlb_layout_create(ctx, n_codes, 0, &lay);
    if (st == GIQLB_OK) st = giqlb_layout_observe_remap(ctx, lay, L->codes, L->start, L->end, L->n, L->start_adj, L->end_adj, slot2code, ok_flag);
    if (st == GIQLB_OK) st = giqlb_layout_observe_remap(ctx, lay, R->codes, R->start, R->end, R->n, R->start_adj, R->end_adj, slot2code, ok_flag);
    if (st == GIQLB_OK) st = giqlb_layout_finalize_async(ctx, lay);   // errors surface with the selection sizes
    giqlb_table_t* T[2] = {L, R};
    int64_t n_sel[2] = {0, 0};
    if (st == GIQLB_OK) {
        begin_call(ctx);
        ProbeCols P[2];
        for (int k = 0; k < 2 && st == GIQLB_OK; k++)
            st = probe_cols(ctx, lay, T[k]->codes, T[k]->start, T[k]->end, T[k]->n, T[k]->start_adj, T[k]->end_adj, &P[k]);
        const int64_t n_bins = (int64_t)(lay->total_extent >> SH_BIN_SHIFT) + 1;
        if (st == GIQLB_OK) st = dev_alloc(ctx, &hist, (size_t)n_bins + 1);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &cuts64, (size_t)world + 1);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &cuts32, (size_t)world + 2);
        for (int k = 0; k < 2 && st == GIQLB_OK; k++) {
            const size_t cap = (size_t)(T[k]->n > 0 ? T[k]->n : 1);
            st = dev_alloc(ctx, &pre[k].gs, cap);
            if (st == GIQLB_OK) st = dev_alloc(ctx, &pre[k].ge, cap);
            if (st == GIQLB_OK) st = dev_alloc(ctx, &pre[k].rows, cap);
            if (st == GIQLB_OK) st = dev_alloc(ctx, &pre[k].counters, 4 + 512);
            if (st == GIQLB_OK && cudaMemsetAsync(pre[k].counters, 0, (4 + 512) * 8, ctx->stream) != cudaSuccess) st = GIQLB_ERR_CUDA;
        }
        if (st == GIQLB_OK) {
            cudaMemsetAsync(hist, 0, ((size_t)n_bins + 1) * 4, ctx->stream);
            cudaMemsetAsync(cuts64, 0, ((size_t)world + 1) * 8, ctx->stream);
            PhaseSpan ps(ctx, GIQLB_PH_SHARD);
            const int64_t cap = (int64_t)ctx->sm_count * 8;
            if (L->n > 0) {
                // a 1-in-16 sample of the left rows places the cuts (balance error ~ 1/sqrt(n/16) of a shard); the widest
                // left interval, needed for the right-side replication bound, comes from the observe pass (slot 0 = L)
                const int64_t stride = L->n >= (1 << 20) ? 16 : 1, n_samp = ceil_div(L->n, stride);
                int64_t blocks = ceil_div(n_samp, SH_THREADS * 8), hcap = (int64_t)ctx->sm_count * 6;
                shard_hist_kernel<<<(unsigned)(blocks < 1 ? 1 : (blocks < hcap ? blocks : hcap)), SH_THREADS, 0, ctx->stream>>>(
                    P[0], stride, hist, hist + n_bins, ctx->d_err);
                shard_cuts_kernel<<<1, 1024, 0, ctx->stream>>>(hist, n_bins, n_samp, world, cuts64);
                ctx->launches_this_call += 2;
                ctx->stats.kernels_total += 2;
            }
            shard_cuts_finish_kernel<<<1, 32, 0, ctx->stream>>>(cuts64, hist + n_bins, lay->d_obs + 2 * GIQLB_MAX_OBSERVED, world, cuts32);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            for (int k = 0; k < 2; k++) {
                if (T[k]->n == 0) continue;
                int64_t blocks = ceil_div(T[k]->n, SH_THREADS * 4);
                shard_select_keygen_kernel<<<(unsigned)(blocks < cap ? blocks : cap), SH_THREADS, 0, ctx->stream>>>(
                    P[k], cuts32, world, rank, k, pre[k].gs, pre[k].ge, pre[k].rows, pre[k].counters, ctx->d_err);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
            if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "intersects_tables_sharded: launch failed");
        }
        // ONE round trip: both selection sizes (and the sticky error flags).  (Measured at 4 GPUs: bringing the two sizes back
        // behind separate events, so that the left index build is enqueued while the right table is still being selected,
        // changed nothing: 1.06 vs 1.03 ms median.)
        if (st == GIQLB_OK) {
            cudaMemcpyAsync(ctx->h_scalars + 40, pre[0].counters, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(ctx->h_scalars + 44, pre[1].counters, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream);
            st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
            n_sel[0] = ctx->h_scalars[43];
            n_sel[1] = ctx->h_scalars[47];
        }
    }
    if (pend) {   // the round trip above also landed the names (or the device's refusal to order them)
        int fallback = 0;
        if (st != GIQLB_OK) cudaStreamSynchronize(ctx->stream);
        const int st2 = giqlb_dict_finish(ctx, pend, &dict, &fallback);
        pend = nullptr;
        if (fallback) {   // codes were never written: redo everything with the host-ordered dictionary
            for (int k = 0; k < 2; k++) {
                dev_free(ctx, pre[k].gs); dev_free(ctx, pre[k].ge); dev_free(ctx, pre[k].rows); dev_free(ctx, pre[k].counters);
            }
            dev_free(ctx, hist); dev_free(ctx, cuts64); dev_free(ctx, cuts32);
            giqlb_layout_free(ctx, lay);
            cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream);
            return sharded_impl(ctx, L, R, n_codes, world, rank, out_l, out_r, capacity, total_host, dict_out, false);
        }
        if (st == GIQLB_OK) st = st2;
    }
    for (int k = 0; k < 2 && st == GIQLB_OK; k++) {
        giqlb_index** dst = k == 0 ? &il : &ir;
        const int64_t deg = ctx->h_scalars[40 + 4 * k], neg = ctx->h_scalars[41 + 4 * k];
        // private indexes that can be rebuilt here: three radix passes + the window-finishing pass (index.cu)
        st = giqlb_index_build_impl(ctx, lay, T[k]->codes, T[k]->start, T[k]->end, nullptr, pre[k].rows, n_sel[k], T[k]->start_adj,
                                    T[k]->end_adj, GIQLB_INDEX_SORT_SHORT_WINDOWS, &pre[k], dst);
        if (st == GIQLB_OK) {   // the counters came back with the selection sizes
            (*dst)->n_degenerate = deg;
            (*dst)->n_negative = neg;
            (*dst)->resolved = true;
        }
    }
    if (st == GIQLB_OK) st = giqlb_intersects_join_impl(ctx, il, ir, out_l, out_r, capacity, total_host, false);
    const bool resort = st == GIQLB_RETRY_FULL_SORT;   // a dense window: redo the call with fully sorted indexes (see pipeline.cu)
    if (resort) st = GIQLB_OK;
    char msg[sizeof(ctx->err)];
    memcpy(msg, ctx->err, sizeof(msg));
    for (int k = 0; k < 2; k++) {
        dev_free(ctx, pre[k].gs); dev_free(ctx, pre[k].ge); dev_free(ctx, pre[k].rows); dev_free(ctx, pre[k].counters);
    }
    dev_free(ctx, hist); dev_free(ctx, cuts64); dev_free(ctx, cuts32);
    giqlb_index_free(ctx, il);
    giqlb_index_free(ctx, ir);
    giqlb_layout_free(ctx, lay);
    if (resort && st == GIQLB_OK) {
        giqlb_dict_free(dict);
        ctx->full_sort_credit = 64;
        return sharded_impl(ctx, L, R, n_codes_in, world, rank, out_l, out_r, capacity, total_host, dict_out, async_dict);
    }
    if (encode && dict && (st == GIQLB_OK || st == GIQLB_ERR_CAPACITY)) L->codes_valid = R->codes_valid = 1;
    if (dict_out) *dict_out = dict;
    else giqlb_dict_free(dict);
    if (st != GIQLB_OK) memcpy(ctx->err, msg, sizeof(msg));
    return st;
}

extern "C" int giqlb_intersects_tables_sharded(giqlb_ctx* ctx, giqlb_table_t* L, giqlb_table_t* R, int32_t n_codes, int world,
                                               int rank, int32_t* out_l, int32_t* out_r, int64_t capacity, int64_t* total_host,
                                               giqlb_dict** dict_out) {
    if (!ctx || !L || !R || !total_host || world < 1 || rank < 0 || rank >= world || L == R)
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_tables_sharded: bad argument");
    if (dict_out) *dict_out = nullptr;
    if (L->codes_valid != R->codes_valid) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_tables_sharded: encode both tables or neither");
    return sharded_impl(ctx, L, R, n_codes, world, rank, out_l, out_r, capacity, total_host, dict_out, true);
}
