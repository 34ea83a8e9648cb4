// common.cuh — context, error plumbing and device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <iterator>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/giqlb.h"

#define GIQLB_SM_COUNT_DEFAULT 148
#define GIQLB_MAX_STRAND_CODES 256
#define GIQLB_STAGE_BYTES (256u * 1024u)

struct giqlb_ctx {
    int device = 0;
    int sm_count = GIQLB_SM_COUNT_DEFAULT;
    cudaStream_t stream = nullptr;   // caller's stream; nullptr = legacy default stream
    char err[512] = {0};
    // small pinned host + device scratch for scalars returned to the host
    int64_t* h_scalars = nullptr;   // pinned, 64 slots
    int64_t* d_scalars = nullptr;   // device, 64 slots
    void* h_stage = nullptr;        // pinned staging buffer for small host -> device tables (GIQLB_STAGE_BYTES)
    int* d_err = nullptr;           // device int[4]: sticky kernel-side error flags ([0] look-back spin limit)
    giqlb_stats_t stats = {};
    bool join_smem_opted_in = false;   // cudaFuncSetAttribute is per device: remembered per context, not per process
    bool dict_smem_opted_in = false;
    int full_sort_credit = 0;          // > 0: the next index builds skip the short-window sort (a dense table was seen; see index.cu)
    int32_t launches_this_call = 0;
    // Scratch cache.  Every allocation of a context is used on ITS stream only, so a freed block can be handed
    // out again immediately (stream order protects it) without a driver call: dev_free parks the block here,
    // dev_alloc takes the smallest parked block that fits (within 2x) before falling back to cudaMalloc.
    std::multimap<size_t, void*> cache_free;
    std::unordered_map<void*, size_t> cache_live;
    size_t cache_bytes = 0;   // bytes owned (parked + live)
    size_t park_limit = 4ull << 30;   // see cache_trim_to; set from the device's memory size when the context is created
    // per-phase CUDA-event profiling (giqlb_profile_*): pairs of events recorded around kernel launches
    bool prof_enabled = false;
    struct ProfSpan { int phase; cudaEvent_t start, stop; };
    std::vector<ProfSpan> prof_spans;
    std::vector<cudaEvent_t> prof_free_events;
    double prof_ms[GIQLB_PH_COUNT] = {0};
    int64_t prof_launches[GIQLB_PH_COUNT] = {0};
};

// RAII span: records an event pair on the context's stream around the launches made while it is alive.
struct PhaseSpan {
    giqlb_ctx* ctx;
    int phase;
    cudaEvent_t start = nullptr, stop = nullptr;
    int launches_before;
    static cudaEvent_t get_event(giqlb_ctx* c) {
        if (!c->prof_free_events.empty()) {
            cudaEvent_t e = c->prof_free_events.back();
            c->prof_free_events.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    PhaseSpan(giqlb_ctx* c, int ph) : ctx(c), phase(ph), launches_before(c->launches_this_call) {
        if (!c->prof_enabled) return;
        start = get_event(c);
        stop = get_event(c);
        cudaEventRecord(start, c->stream);
    }
    ~PhaseSpan() {
        if (!start) return;
        cudaEventRecord(stop, ctx->stream);
        ctx->prof_spans.push_back({phase, start, stop});
        ctx->prof_launches[phase] += ctx->launches_this_call - launches_before;
    }
};

struct giqlb_dict {
    std::vector<std::string> names;   // sorted bytewise; code = position
};

// Genome linearisation.  For chromosome code c with observed canonical coordinate range [lo, hi]:
//   g(pos) = base[c] + 1 + (pos - lo[c])            (uint32)
//   base[c+1] = base[c] + (hi - lo) + 3 + pad       (one guard cell either side, `pad` extra)
// so coordinates of different chromosomes never compare equal or interleave, and a probe coordinate
// clamped to [lo-1, hi+1] stays inside its own chromosome's cell range.
struct giqlb_layout {
    int32_t n_codes = 0;
    int64_t pad = 0;
    int64_t* d_lohi = nullptr;   // [2*n_codes]: lo (init INT64_MAX) then hi (init INT64_MIN)
    uint32_t* d_base = nullptr;  // [n_codes + 1] linear base per code; base[n_codes] = total extent
    int32_t* d_lo32 = nullptr;   // [n_codes] lo as int32 (valid after finalize; 0 for empty chromosomes)
    int32_t* d_hi32 = nullptr;   // [n_codes] hi (lo - 1 for empty chromosomes)
    bool finalized = false;
    uint64_t total_extent = 0;   // host copy after finalize
    int key_bits = 32;
    std::vector<int64_t> h_lo, h_hi;
    std::vector<uint32_t> h_base;
    // tables folded in by observe(): their degenerate-row counts are gathered by the same pass and fetched by
    // finalize()'s round trip, so an index over an observed table needs no round trip of its own
    struct Observed {
        const int32_t* code;
        int64_t n;
        int32_t sadj, eadj;
        int64_t n_degenerate, n_negative;
    };
    std::vector<Observed> observed;
    unsigned long long* d_obs = nullptr;   // [2 * GIQLB_MAX_OBSERVED]: end <= start, end < start per observe() call;
                                           // then [GIQLB_MAX_OBSERVED]: widest interval of each observed table
};
#define GIQLB_MAX_OBSERVED 16

struct giqlb_index {
    int64_t n = 0;
    uint32_t* gs = nullptr;      // linearised start, ascending (within each strand block when grouped)
    uint32_t* ge = nullptr;      // linearised end of the same row
    int32_t* rid = nullptr;      // input row id
    uint8_t* strand = nullptr;   // strand code of the same row (nullptr when table has none)
    uint32_t* pmax = nullptr;    // lazily built inclusive running max of ge (restarts per strand block)
    uint32_t* ge_sorted = nullptr;   // lazily built: ge ascending (per strand block) ...
    int32_t* pos_by_end = nullptr;   // ... and the sorted position (into gs/ge/rid) of each
    int64_t n_degenerate = 0;    // rows with end <= start (host copy; valid when `resolved`)
    int64_t n_negative = 0;      // rows with end <  start
    bool resolved = true;        // false: the counts still sit in d_counters (giqlb_index_resolve fetches them)
    bool maybe_unfinished = false;  // built with GIQLB_INDEX_SORT_SHORT_WINDOWS and d_counters[3] not looked at yet: consumers must check first
    bool sort_unfinished = false;   // the short-window sort met a dense window: gs/ge/rid are NOT in order, rebuild (host copy of d_counters[3])
    unsigned long long* d_counters = nullptr;   // [4] end<=start, end<start, rows outside the layout, rows in windows too dense for
                                                // the short-window sort (GIQLB_INDEX_SORT_SHORT_WINDOWS builds only)
    // lazily built bucket directories (giqlb_index_ensure_dirs): for every strand block with rows, dir[slot][b] = first
    // block-relative position whose value is >= b << dir_shift, over gs / pmax / ge_sorted.  A search then starts inside one
    // bucket (about one row) instead of walking log2(n) dependent loads from the whole array.
    // lazily built block maxima of ge (giqlb_index_ensure_block_max): bmax1[b] = max ge over rows [32 b, 32 b + 32),
    // bmax2[b] over rows [1024 b, 1024 b + 1024).  "Next row at or after j whose end exceeds x" skips whole blocks with them, so
    // one chromosome-spanning row does not turn every later probe into a scan of the chromosome.
    uint32_t* bmax1 = nullptr;
    uint32_t* bmax2 = nullptr;
    uint32_t* dir_gs = nullptr;
    // giqlb_cluster_shard is called twice per shard (cluster counts, then — after the ranks exchanged them — the ids): the running
    // cluster count of the first call is kept for the second (valid for the same distance and carry)
    uint32_t* shard_G = nullptr;
    long long shard_G_distance = -1;
    uint32_t shard_G_carry = 0;
    uint4* dir_exists = nullptr;     // lazily built (giqlb_index_ensure_exists_dir): per bucket of dir_gs one 32-byte sector with the first two
                                     // rows' starts and running maxima — it answers most SEMI / ANTI probes alone (plain indexes only)
    uint32_t* dir_pmax = nullptr;
    uint32_t* dir_ges = nullptr;
    int32_t* d_dir_slot = nullptr;   // device [GIQLB_MAX_STRAND_CODES]: strand block -> directory slot, -1 for empty blocks
    int dir_shift = 0;
    uint32_t dir_stride = 0;         // entries per slot = (layout extent >> dir_shift) + 2
    bool tie_by_end = false;     // equal starts are ordered by end
    bool grouped_by_strand = false;
    int32_t n_blocks = 1;                                   // strand blocks (1 when not grouped)
    int64_t block_off[GIQLB_MAX_STRAND_CODES + 1] = {0};    // host: row offsets of each strand-code block
    int64_t* d_block_off = nullptr;                         // device copy [257]
    const giqlb_layout* layout = nullptr;
};

struct giqlb_joinplan {
    const giqlb_index* L = nullptr;
    const giqlb_index* R = nullptr;
    int64_t n_tiles = 0;
    int32_t* part = nullptr;                  // [n_tiles + 1] merge-path partition: L rows before each tile (fast path)
    uint32_t* item_code = nullptr;            // [n_tiles * 1024] run start/class/length of every merged row (fast path)
    unsigned long long* tile_off = nullptr;   // [n_tiles + 1] exclusive pair offsets (device)
    unsigned long long* unit_off = nullptr;   // [n_tiles + 1] exclusive emit-work-unit offsets (fast path)
    int64_t total = 0;
    bool general = false;                     // degenerate-interval path (literal predicate scan)
};

// ---- host-side helpers ---------------------------------------------------------------------------
static inline int giqlb_fail(giqlb_ctx* ctx, int code, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
static inline int giqlb_fail(giqlb_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return giqlb_fail((ctx), _e == cudaErrorMemoryAllocation ? GIQLB_ERR_NOMEM : GIQLB_ERR_CUDA, \
                              "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
    } while (0)

#define GIQLB_TRY(expr)              \
    do {                             \
        int _s = (expr);             \
        if (_s != GIQLB_OK) return _s; \
    } while (0)

#define LAUNCH_CHECK(ctx)                 \
    do {                                  \
        (ctx)->launches_this_call++;      \
        (ctx)->stats.kernels_total++;     \
        CUDA_TRY((ctx), cudaGetLastError()); \
    } while (0)

static inline int dev_alloc_bytes(giqlb_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    bytes = (bytes + 64 + 511) & ~(size_t)511;   // >= 64 bytes of slack: vector loads / 16-byte bulk copies of a tail stay in bounds
    auto it = ctx->cache_free.lower_bound(bytes);
    if (it != ctx->cache_free.end() && it->first <= 2 * bytes + 4096) {
        *p = it->second;
        ctx->cache_live.emplace(it->second, it->first);
        ctx->cache_free.erase(it);
        return GIQLB_OK;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {   // give the parked blocks back to the driver and retry once
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        for (auto& kv : ctx->cache_free) {
            cudaFree(kv.second);
            ctx->cache_bytes -= kv.first;
        }
        ctx->cache_free.clear();
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return giqlb_fail(ctx, e == cudaErrorMemoryAllocation ? GIQLB_ERR_NOMEM : GIQLB_ERR_CUDA, "cudaMalloc(%zu) -> %s", bytes,
                          cudaGetErrorString(e));
    }
    ctx->cache_live.emplace(*p, bytes);
    ctx->cache_bytes += bytes;
    return GIQLB_OK;
}
template <typename T>
static inline int dev_alloc(giqlb_ctx* ctx, T** p, size_t count) {
    return dev_alloc_bytes(ctx, reinterpret_cast<void**>(p), count * sizeof(T));
}
template <typename T>
static inline void dev_free(giqlb_ctx* ctx, T* p) {
    if (!p) return;
    auto it = ctx->cache_live.find((void*)p);
    if (it == ctx->cache_live.end()) return;   // not ours (or freed twice): never hand it to the driver
    ctx->cache_free.emplace(it->second, it->first);
    ctx->cache_live.erase(it);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

static inline void begin_call(giqlb_ctx* ctx) {
    ctx->launches_this_call = 0;
    ctx->err[0] = 0;
}
// Parked scratch above ctx->park_limit bytes is handed back to the driver when a call ends (largest blocks first): PyTorch's
// allocator, which owns every input and output buffer, cannot see or reclaim what the library parks.  The limit is the larger
// of GIQLB_SCRATCH_PARK_LIMIT and one eighth of the device's memory (22 GB of a B200's 180): an operator over 200 M rows
// works in 5 GB of scratch, and with a 4 GB limit every call paid cudaMalloc / cudaFree for it (measured: +30 ms a call).
#ifndef GIQLB_SCRATCH_PARK_LIMIT
#define GIQLB_SCRATCH_PARK_LIMIT (4ull << 30)
#endif
static inline void cache_trim_to(giqlb_ctx* ctx, size_t limit) {
    size_t parked = 0;
    for (auto& kv : ctx->cache_free) parked += kv.first;
    if (parked <= limit) return;
    cudaStreamSynchronize(ctx->stream);   // a parked block may still be in use by work enqueued on the stream
    while (parked > limit && !ctx->cache_free.empty()) {
        auto it = std::prev(ctx->cache_free.end());
        cudaFree(it->second);
        parked -= it->first;
        ctx->cache_bytes -= it->first;
        ctx->cache_free.erase(it);
    }
}
static inline int end_call(giqlb_ctx* ctx, int status = GIQLB_OK) {
    ctx->stats.kernels_launched = ctx->launches_this_call;
    cache_trim_to(ctx, ctx->park_limit);
    return status;
}

// copies n int64 scalars from device to the pinned host scratch and waits for them
static inline int fetch_scalars(giqlb_ctx* ctx, const void* d_src, int n_words) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars, d_src, (size_t)n_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return GIQLB_OK;
}

// Raw columns of one table plus the layout tables needed to linearise them on the fly.
struct ProbeCols {
    const int32_t* code;
    const int32_t* start;
    const int32_t* end;
    int64_t n;
    int32_t sadj, eadj;
    int32_t n_codes;
    const uint32_t* base;
    const int32_t* lo32;
    const int32_t* hi32;
};


static inline int probe_cols(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                             const int32_t* end, int64_t n, int32_t sadj, int32_t eadj, ProbeCols* out) {
    if (!lay || !lay->finalized) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "layout not finalized");
    if (n < 0 || (n > 0 && (!code || !start || !end))) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "bad column pointers");
    *out = ProbeCols{code, start, end, n, sadj, eadj, lay->n_codes, lay->d_base, lay->d_lo32, lay->d_hi32};
    return GIQLB_OK;
}

// fetch_scalars + the sticky kernel-side error flags in the same round trip
// the sticky kernel-side error flags as copied to h_scalars + 60 -> status (and the flags are cleared)
static inline int decode_err_flags(giqlb_ctx* ctx) {
    const int* e = reinterpret_cast<const int*>(ctx->h_scalars + 60);
    if (e[0] || e[1] || e[3]) {
        cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream);
        if (e[1] == 2) return giqlb_fail(ctx, GIQLB_ERR_RANGE, "layout: the summed chromosome extent does not fit the 32-bit linear coordinate");
        if (e[1] == 3) return giqlb_fail(ctx, GIQLB_ERR_RANGE, "nearest: a distance does not fit the int32 output column");
        if (e[1] == 4) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "distance_rows: a row id lies outside its table");
        if (e[1]) return giqlb_fail(ctx, GIQLB_ERR_RANGE, "layout: a row has a chromosome code outside the dictionary or a coordinate outside [-2^30, 2^31-4]");
        if (e[3]) return giqlb_fail(ctx, GIQLB_ERR_RANGE, "index: rows lie outside the layout (observe the table before finalize)");
        return giqlb_fail(ctx, GIQLB_ERR_INTERNAL, "look-back spin limit exceeded (kernel-side)");
    }
    return GIQLB_OK;
}
static inline int fetch_scalars_checked(giqlb_ctx* ctx, const void* d_src, int n_words) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars, d_src, (size_t)n_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars + 60, ctx->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return decode_err_flags(ctx);
}

// ---- device helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

// first index in [lo, hi) with a[idx] >= x (a ascending)
template <typename T, typename I>
__device__ __forceinline__ I lower_bound_dev(const T* __restrict__ a, I lo, I hi, T x) {
    while (lo < hi) {
        I mid = lo + ((hi - lo) >> 1);
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index in [lo, hi) with a[idx] > x
template <typename T, typename I>
__device__ __forceinline__ I upper_bound_dev(const T* __restrict__ a, I lo, I hi, T x) {
    while (lo < hi) {
        I mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Searches through a bucket directory (see giqlb_index::dir_*): tab[b] = first index with a[idx] >= b << shift, so the
// answer for x lies in [tab[b], tab[b + 1]] with b = x >> shift.  The directory covers the layout's extent only: larger
// x are clamped into the last bucket `last` (= entries - 2; tab[last + 1] = n), which stays exact.
__device__ __forceinline__ int64_t dir_lower_bound(const uint32_t* __restrict__ tab, int shift, uint32_t last, const uint32_t* __restrict__ a, uint32_t x) {
    const uint32_t b = min(x >> shift, last);
    return lower_bound_dev<uint32_t, int64_t>(a, (int64_t)__ldg(tab + b), (int64_t)__ldg(tab + b + 1), x);
}
__device__ __forceinline__ int64_t dir_upper_bound(const uint32_t* __restrict__ tab, int shift, uint32_t last, const uint32_t* __restrict__ a, uint32_t x) {
    const uint32_t b = min(x >> shift, last);
    return upper_bound_dev<uint32_t, int64_t>(a, (int64_t)__ldg(tab + b), (int64_t)__ldg(tab + b + 1), x);
}

// First position in [j, hi) whose ge exceeds thr (hi when there is none), skipping 32- and 1024-row blocks whose maximum
// does not (giqlb_index::bmax1 / bmax2; a block straddling hi is judged by its full maximum, which only skips less).
__device__ __forceinline__ int64_t next_end_above(const uint32_t* __restrict__ ge, const uint32_t* __restrict__ m1,
                                                  const uint32_t* __restrict__ m2, int64_t j, int64_t hi, uint32_t thr) {
    while (j < hi) {
        const int64_t bend = min(hi, (j | 31) + 1);
        if ((j & 31) != 0 || __ldg(m1 + (j >> 5)) > thr) {
            // four rows per load once the address is 16-byte aligned (this scan was 52 % of NEAREST k = 5's instructions when it
            // went row by row: behind one long target the running maximum admits whole blocks of rows that end before the query)
            for (; j < bend && (reinterpret_cast<uintptr_t>(ge + j) & 15) != 0; j++)
                if (__ldg(ge + j) > thr) return j;
            for (; j + 4 <= bend; j += 4) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(ge + j));
                const unsigned m = (v.x > thr ? 1u : 0u) | (v.y > thr ? 2u : 0u) | (v.z > thr ? 4u : 0u) | (v.w > thr ? 8u : 0u);
                if (m) return j + (__ffs(m) - 1);
            }
            for (; j < bend; j++)
                if (__ldg(ge + j) > thr) return j;
        } else {
            j = bend;
        }
        while (j < hi) {                                        // j is 32-aligned here
            if ((j & 1023) == 0 && __ldg(m2 + (j >> 10)) <= thr) { j += 1024; continue; }
            if (__ldg(m1 + (j >> 5)) <= thr) { j += 32; continue; }
            break;
        }
    }
    return hi;
}

// streaming (read-once) loads / stores that do not pollute L1
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream_u32x4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ void st_stream_u32x4(uint4* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w));
}

// linearise one probe row; coordinates are clamped into the chromosome's guard cells, which preserves
// every comparison against indexed rows of that chromosome.  Returns false when the chromosome holds no
// indexed row at all (or the code is outside the dictionary): nothing can match.
__device__ __forceinline__ int32_t add_sat_s32(int32_t a, int32_t b) {
    int32_t r;
    asm("add.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// Linear coordinates of a probe interval: canonical offsets applied, both ends clamped to [lo - 1, hi + 1] of the row's
// chromosome.  All in 32 bits (this sits on the path of every probe row; the 64-bit version cost ~60 instructions): the layout
// admits coordinates in [-2^30, 2^31 - 4] only, so lo - 1 and hi + 1 cannot wrap, and an offset that would carry the column
// value past INT32_MIN / MAX saturates there — beyond the clamp on either side, so the result is the same.
__device__ __forceinline__ bool probe_linearise_values(const ProbeCols& P, int32_t c, int32_t start, int32_t end, uint32_t& gs, uint32_t& ge) {
    if (c < 0 || c >= P.n_codes) return false;
    const int32_t lo = __ldg(P.lo32 + c), hi = __ldg(P.hi32 + c);
    if (hi < lo) return false;
    const int32_t s = max(min(add_sat_s32(start, P.sadj), hi + 1), lo - 1);
    const int32_t e = max(min(add_sat_s32(end, P.eadj), hi + 1), lo - 1);
    const uint32_t b = __ldg(P.base + c) + 1u;
    gs = b + (uint32_t)(s - lo);   // lo - 1 maps to the lower guard cell (wraps to b - 1)
    ge = b + (uint32_t)(e - lo);
    return true;
}
__device__ __forceinline__ bool probe_linearise(const ProbeCols& P, int64_t i, uint32_t& gs, uint32_t& ge) {
    return probe_linearise_values(P, P.code[i], P.start[i], P.end[i], gs, ge);
}

template <typename T>
__device__ __forceinline__ T warp_incl_scan_add(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(FULL_MASK, v, d);
        if (lane_id() >= (uint32_t)d) v += o;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_reduce_add(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}

// Block-wide exclusive scan of one value per thread.  `scratch` holds >= (blockDim/32 + 1) T's.
// Returns the exclusive prefix; *total receives the block sum.  Contains two __syncthreads(); callers
// must __syncthreads() before reusing `scratch`.
template <typename T>
__device__ __forceinline__ T block_excl_scan_add(T v, T* scratch, T* total) {
    T incl = warp_incl_scan_add(v);
    const uint32_t w = warp_id(), l = lane_id(), nw = (blockDim.x + 31) >> 5;
    if (l == 31) scratch[w] = incl;
    __syncthreads();
    if (w == 0) {
        T x = l < nw ? scratch[l] : T(0);
        T xi = warp_incl_scan_add(x);
        if (l < nw) scratch[l] = xi - x;
        if (l == nw - 1) scratch[nw] = xi;
    }
    __syncthreads();
    T res = scratch[w] + incl - v;
    *total = scratch[nw];
    return res;
}

#endif  // __CUDACC__

// ---- internal cross-file entry points ---------------------------------------------------------------
int giqlb_radix_sort_pairs_u32(giqlb_ctx* ctx, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b,
                               uint32_t* vals_b, int64_t n, int begin_bit, int end_bit, int identity_vals,
                               int* result_in_b, const uint32_t* hist_dev = nullptr, const uint32_t* gather_src = nullptr,
                               uint32_t* gather_dst = nullptr, uint32_t* extra_a = nullptr, uint32_t* extra_b = nullptr);
int giqlb_excl_scan_u64(giqlb_ctx* ctx, const unsigned long long* in, unsigned long long* out, int64_t n,
                        unsigned long long* total_dev);
int giqlb_incl_scan_max_u32(giqlb_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n,
                            const int64_t* d_block_off, int n_blocks);
// keys produced by a fused select + keygen pass (shard.cu), handed to the index builder
struct giqlb_prekeyed {
    uint32_t* gs = nullptr;               // [n] linear starts of the selected rows (unsorted)
    uint32_t* ge = nullptr;               // [n] their linear ends
    int32_t* rows = nullptr;              // [n] their row ids
    unsigned long long* counters = nullptr;   // [4 + 512] as giqlb_index::d_counters
};
int giqlb_index_build_impl(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                           const int32_t* end, const uint8_t* strand, const int32_t* rows, int64_t n, int32_t start_adj,
                           int32_t end_adj, int flags, giqlb_prekeyed* pre, giqlb_index** out);
struct giqlb_dict_pending;
#define GIQLB_ASYNC_DICT_MAX_CODES 1024   // upper bound of the code space while the dictionary is still on the device
// defer_remap: leave the parked table slots in codes_l / codes_r; the caller translates them with
// giqlb_layout_observe_remap (tables given in slot2code / ok_flag of giqlb_dict_pending_tables)
int giqlb_dict_encode_async(giqlb_ctx* ctx, const int32_t* offs_l, const uint8_t* bytes_l, int64_t n_l, const int32_t* offs_r,
                            const uint8_t* bytes_r, int64_t n_r, int32_t* codes_l, int32_t* codes_r, giqlb_dict_pending** out,
                            bool defer_remap = false);
void giqlb_dict_pending_tables(const giqlb_dict_pending* p, const int32_t** slot2code, const int** ok_flag);
int giqlb_layout_observe_remap(giqlb_ctx* ctx, giqlb_layout* lay, int32_t* code, const int32_t* start, const int32_t* end, int64_t n,
                               int32_t start_adj, int32_t end_adj, const int32_t* slot2code, const int* ok_flag);
int giqlb_dict_finish(giqlb_ctx* ctx, giqlb_dict_pending* p, giqlb_dict** dict_out, int* fallback);
int giqlb_layout_finalize_async(giqlb_ctx* ctx, giqlb_layout* lay);
int giqlb_intersects_join_impl(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, int32_t* out_l, int32_t* out_r, int64_t capacity,
                               int64_t* total_host, bool speculative);
int giqlb_index_ensure_pmax(giqlb_ctx* ctx, giqlb_index* ix);
// operators that report input rows cannot take an index built with GIQLB_INDEX_NO_ROW_IDS
static inline int require_row_ids(giqlb_ctx* ctx, const giqlb_index* ix, const char* who) {
    if (ix && ix->n > 0 && !ix->rid) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "%s: the index was built with GIQLB_INDEX_NO_ROW_IDS", who);
    return GIQLB_OK;
}
int giqlb_index_resolve(giqlb_ctx* ctx, giqlb_index* a, giqlb_index* b);
// Internal giqlb_index_build flag (never part of the C-ABI: callers must be able to rebuild the index).  Sort the start keys on
// bits [8, key_bits) only — one radix pass fewer — and let one finishing pass order every run of equal upper bits ("window" of
// 256 linear positions) by its low byte while it gathers the ends.  Windows holding more than 32 rows are not finished: the
// build then counts their rows in d_counters[3], the index is NOT sorted, and the consumer (pipeline.cu) must skip it, rebuild
// without the flag and go on with ctx->full_sort_credit set.
#define GIQLB_INDEX_SORT_SHORT_WINDOWS 0x100
#define GIQLB_RETRY_FULL_SORT 1   // internal status of giqlb_intersects_join_impl: an index came back with sort_unfinished
int giqlb_index_ensure_end_sorted(giqlb_ctx* ctx, giqlb_index* ix);
// bucket directories over gs (always), pmax and ge_sorted (on request; they must exist already)
int giqlb_index_ensure_dirs(giqlb_ctx* ctx, giqlb_index* ix, bool want_pmax, bool want_ges);
int giqlb_index_ensure_block_max(giqlb_ctx* ctx, giqlb_index* ix);
int giqlb_index_ensure_exists_dir(giqlb_ctx* ctx, giqlb_index* ix);
int giqlb_bucket_dir_build(giqlb_ctx* ctx, const uint32_t* arr, int64_t n, uint64_t extent, int* shift, uint32_t* stride, uint32_t** tab);
