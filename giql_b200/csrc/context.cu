// context.cu — context lifetime, error/statistics plumbing and small generic device utilities
// (exclusive scan of tile counts, running max, flag compaction).
#include "common.cuh"
#include "scan.cuh"

extern "C" int giqlb_abi_version(void) { return GIQLB_ABI_VERSION; }

extern "C" int giqlb_create(int device, void* cuda_stream, giqlb_ctx** out) {
    if (!out) return GIQLB_ERR_INVALID;
    *out = nullptr;
    giqlb_ctx* ctx = new (std::nothrow) giqlb_ctx();
    if (!ctx) return GIQLB_ERR_NOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        delete ctx;
        return GIQLB_ERR_CUDA;
    }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) ctx->sm_count = sms;
    // NULL selects the legacy default stream: the caller's other work on that stream (PyTorch's default
    // stream has handle 0) stays ordered with ours.  The library never creates a stream of its own.
    ctx->stream = (cudaStream_t)cuda_stream;
    if (cudaMallocHost((void**)&ctx->h_scalars, 64 * sizeof(int64_t)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_stage, GIQLB_STAGE_BYTES) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_scalars, 64 * sizeof(int64_t)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_err, 4 * sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_err, 0, 4 * sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_scalars, 0, 64 * sizeof(int64_t)) != cudaSuccess) {
        giqlb_destroy(ctx);
        return GIQLB_ERR_NOMEM;
    }
    *out = ctx;
    return GIQLB_OK;
}

extern "C" int giqlb_destroy(giqlb_ctx* ctx) {
    if (!ctx) return GIQLB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& sp : ctx->prof_spans) { cudaEventDestroy(sp.start); cudaEventDestroy(sp.stop); }
    for (auto e : ctx->prof_free_events) cudaEventDestroy(e);
    for (auto& kv : ctx->cache_free) cudaFree(kv.second);
    for (auto& kv : ctx->cache_live) cudaFree(kv.first);   // objects the caller never freed
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_scalars) cudaFree(ctx->d_scalars);
    if (ctx->d_err) cudaFree(ctx->d_err);
    delete ctx;
    return GIQLB_OK;
}

extern "C" const char* giqlb_last_error(const giqlb_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int giqlb_sync(giqlb_ctx* ctx) {
    if (!ctx) return GIQLB_ERR_INVALID;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int h_err[4];
    CUDA_TRY(ctx, cudaMemcpy(h_err, ctx->d_err, sizeof(h_err), cudaMemcpyDeviceToHost));
    if (h_err[0] || h_err[3]) {
        cudaMemset(ctx->d_err, 0, sizeof(h_err));
        if (h_err[3]) return giqlb_fail(ctx, GIQLB_ERR_RANGE, "index: rows lie outside the layout (observe the table before finalize)");
        return giqlb_fail(ctx, GIQLB_ERR_INTERNAL, "look-back spin limit exceeded (kernel-side)");
    }
    return GIQLB_OK;
}

extern "C" int giqlb_trim(giqlb_ctx* ctx) {
    if (!ctx) return GIQLB_ERR_INVALID;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& kv : ctx->cache_free) {
        cudaFree(kv.second);
        ctx->cache_bytes -= kv.first;
    }
    ctx->cache_free.clear();
    return GIQLB_OK;
}

extern "C" int giqlb_stats(const giqlb_ctx* ctx, giqlb_stats_t* out_host) {
    if (!ctx || !out_host) return GIQLB_ERR_INVALID;
    *out_host = ctx->stats;
    return GIQLB_OK;
}

extern "C" int giqlb_profile_enable(giqlb_ctx* ctx, int on) {
    if (!ctx) return GIQLB_ERR_INVALID;
    ctx->prof_enabled = on != 0;
    return GIQLB_OK;
}

extern "C" int giqlb_profile_read(giqlb_ctx* ctx, double* ms_host, int64_t* launches_host) {
    if (!ctx) return GIQLB_ERR_INVALID;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& sp : ctx->prof_spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.start, sp.stop) == cudaSuccess) ctx->prof_ms[sp.phase] += ms;
        ctx->prof_free_events.push_back(sp.start);
        ctx->prof_free_events.push_back(sp.stop);
    }
    ctx->prof_spans.clear();
    for (int i = 0; i < GIQLB_PH_COUNT; i++) {
        if (ms_host) ms_host[i] += ctx->prof_ms[i];
        if (launches_host) launches_host[i] += ctx->prof_launches[i];
        ctx->prof_ms[i] = 0;
        ctx->prof_launches[i] = 0;
    }
    return GIQLB_OK;
}

// ---- exclusive scan of a short u64 array (tile counts) by one CTA ----------------------------------
namespace {
constexpr int XS_THREADS = 1024;

__global__ void __launch_bounds__(XS_THREADS)
excl_scan_u64_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, int64_t n,
                     unsigned long long* __restrict__ total) {
    __shared__ unsigned long long scratch[XS_THREADS / 32 + 2];
    unsigned long long carry = 0;
    for (int64_t base = 0; base < n; base += XS_THREADS) {
        int64_t i = base + threadIdx.x;
        unsigned long long v = i < n ? in[i] : 0ull;
        unsigned long long tot;
        unsigned long long ex = block_excl_scan_add<unsigned long long>(v, scratch, &tot);
        if (i < n) out[i] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[n] = carry;
        if (total) *total = carry;
    }
}

struct InU32 {
    const uint32_t* p;
    __device__ uint32_t operator()(int64_t i) const { return p[i]; }
};
struct OutInclMax {
    uint32_t* p;
    __device__ void operator()(int64_t i, uint32_t excl, uint32_t v) const { p[i] = excl > v ? excl : v; }
};

struct InFlagEq {
    const uint8_t* f;
    int want;
    __device__ unsigned long long operator()(int64_t i) const { return (f[i] != 0) == (want != 0) ? 1ull : 0ull; }
};
struct OutCompact {
    int32_t* rows;
    unsigned long long* total;
    int64_t n;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        if (v) rows[excl] = (int32_t)i;
        if (i == n - 1) *total = excl + v;
    }
};
}  // namespace

// out has n + 1 entries; out[n] (and *total_dev when given) receive the grand total.
int giqlb_excl_scan_u64(giqlb_ctx* ctx, const unsigned long long* in, unsigned long long* out, int64_t n,
                        unsigned long long* total_dev) {
    PhaseSpan ps(ctx, GIQLB_PH_JOIN_SCAN);
    excl_scan_u64_kernel<<<1, XS_THREADS, 0, ctx->stream>>>(in, out, n, total_dev);
    LAUNCH_CHECK(ctx);
    return GIQLB_OK;
}

// out[i] = max(in[blk_lo .. i]) where blk_lo is the start of i's strand block (block_off may be null).
int giqlb_incl_scan_max_u32(giqlb_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n,
                            const int64_t* h_block_off, int n_blocks) {
    using namespace giqlb_scan;
    PhaseSpan ps(ctx, GIQLB_PH_SCAN_MAX);
    if (!h_block_off || n_blocks <= 1) return launch_scan<OpMaxU32>(ctx, n, InU32{in}, OutInclMax{out});
    for (int b = 0; b < n_blocks; b++) {
        int64_t lo = h_block_off[b], hi = h_block_off[b + 1];
        if (hi > lo) GIQLB_TRY((launch_scan<OpMaxU32>(ctx, hi - lo, InU32{in + lo}, OutInclMax{out + lo})));
    }
    return GIQLB_OK;
}

extern "C" int giqlb_select_rows(giqlb_ctx* ctx, const uint8_t* flags, int64_t n, int want, int32_t* out_rows,
                                 int64_t* n_out_host) {
    using namespace giqlb_scan;
    if (!ctx || n < 0 || (n > 0 && (!flags || !out_rows))) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "select_rows: bad argument");
    begin_call(ctx);
    unsigned long long* total = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 8);
    CUDA_TRY(ctx, cudaMemsetAsync(total, 0, 8, ctx->stream));
    {
        PhaseSpan ps(ctx, GIQLB_PH_SELECT);
        GIQLB_TRY((launch_scan<OpAddU64>(ctx, n, InFlagEq{flags, want}, OutCompact{out_rows, total, n})));
    }
    if (n_out_host) {
        GIQLB_TRY(fetch_scalars_checked(ctx, total, 1));
        *n_out_host = ctx->h_scalars[0];
    }
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = n_out_host ? *n_out_host : -1;
    ctx->stats.algorithmic_bytes = n + 4 * (n_out_host ? *n_out_host : 0);
    return end_call(ctx);
}
