// radix_sort.cu — device LSD radix sort of uint32 keys with a carried uint32 payload (row ids).
//
// Replaces the per-chromosome sorts DuckDB's IE_JOIN performs for the SQL emitted at
// src/giql/expanders/intersects_duckdb.py:1283-1299 (reference) with ONE sort of the linearised
// (chrom, start) key.  Design ("onesweep"): one histogram kernel counts every digit of every pass in a
// single read of the keys; then one kernel per 8-bit digit ranks a 4096-key tile with warp match-any
// multi-split, obtains its global digit bases through a decoupled look-back over per-tile status words
// (tiles take tickets from an atomic counter, so a tile only ever waits on tiles already scheduled),
// stages the tile in shared memory in digit order and writes it out in coalesced runs.
// HBM traffic: 4 B/key (histogram) + 16 B/key/pass (keys+payload in and out).
#include "common.cuh"
#include <type_traits>

namespace {

// threads per CTA of the digit passes (the first 256 of them also own one digit each); the tile stays 4096 keys
// (measured: 512 threads x 8 keys at three CTAs per SM — 48 resident warps — 77 us per pass against 69 us at 256 x 16)
#ifndef GIQLB_RS_THREADS
#define GIQLB_RS_THREADS 256
#endif
constexpr int RS_THREADS = GIQLB_RS_THREADS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_HIST_THREADS = 256;
#ifndef GIQLB_RS_IPT
#define GIQLB_RS_IPT (4096 / GIQLB_RS_THREADS)
#endif
// per-warp digit counters: a warp holds at most 32 * IPT keys, so 16 bits do when there are many warps
typedef std::conditional<(RS_WARPS > 8), uint16_t, uint32_t>::type rs_hist_t;
constexpr int RS_IPT = GIQLB_RS_IPT;
#ifndef GIQLB_RS_LOOKBACK
#define GIQLB_RS_LOOKBACK 4      // predecessor tiles fetched per look-back step
#endif
#ifndef GIQLB_RS_LOOKBACK_WIDE
#define GIQLB_RS_LOOKBACK_WIDE 4 // ... per step after the first (prefetched) one
#endif
#ifndef GIQLB_RS_BACKOFF_NS
#define GIQLB_RS_BACKOFF_NS 20   // pause between polls of a predecessor tile that has not published yet
#endif
#ifndef GIQLB_RS_CTAS_PER_SM
#define GIQLB_RS_CTAS_PER_SM 4   // measured: 5 (48 registers, small spills) is slower
#endif
#ifndef GIQLB_RS_USE_MATCH
#define GIQLB_RS_USE_MATCH 0
#endif
constexpr int RS_TILE = RS_THREADS * RS_IPT;   // 4096 keys per tile
constexpr int RS_RADIX = 256;
constexpr int RS_MAX_PASSES = 4;
constexpr int RS_HIST_COPIES = 1;   // 4 lane-private copies measured no faster: ATOMS issue rate, not bank conflicts, bounds the kernel

constexpr uint32_t ST_FLAG_AGG = 1u << 30;
constexpr uint32_t ST_FLAG_INCL = 2u << 30;
constexpr uint32_t ST_VALUE_MASK = (1u << 30) - 1;
constexpr uint32_t RS_SPIN_LIMIT = 1u << 26;

// One look-back step over W predecessor status words sv[0] (nearest) .. sv[W-1]: adds what is published to `prefix`, moves
// t, sets `done` at the first inclusive word; returns false when an unpublished word stopped the step early (retry).
// W is a multiple of 4: far from the frontier all W words are plain aggregates, and since W flag fields of 1 << 30 sum to
// a multiple of 2^32, adding the raw words adds exactly the W counts.
template <int W>
__device__ __forceinline__ bool lookback_step(const uint32_t (&sv)[W], int64_t& t, uint32_t& prefix, bool& done) {
    static_assert(W % 4 == 0, "the raw-sum shortcut needs W * 2^30 == 0 mod 2^32");
    uint32_t all_and = sv[0], all_or = sv[0], sum = sv[0];
#pragma unroll
    for (int j = 1; j < W; j++) {
        all_and &= sv[j];
        all_or |= sv[j];
        sum += sv[j];
    }
    if ((all_and & ~ST_VALUE_MASK) == ST_FLAG_AGG && (all_or & ~ST_VALUE_MASK) == ST_FLAG_AGG) {
        prefix += sum;
        t -= W;
        return true;
    }
    int used = 0;
#pragma unroll
    for (int j = 0; j < W; j++) {
        if (done || used != j) continue;
        const uint32_t flag = sv[j] & ~ST_VALUE_MASK;
        if (flag == 0) continue;              // not published yet: retry from this tile
        prefix += sv[j] & ST_VALUE_MASK;
        used = j + 1;
        if (flag == ST_FLAG_INCL) done = true;
    }
    t -= used;
    return used == W;
}

// ---- histogram of all digits in one read ---------------------------------------------------------
__global__ void __launch_bounds__(RS_HIST_THREADS)
rs_histogram_kernel(const uint32_t* __restrict__ keys, int64_t n, int begin_bit, int n_passes,
                    uint32_t* __restrict__ ghist /* [n_passes][256] */) {
    // RS_HIST_COPIES private copies, chosen by lane: the shared-memory atomics are the bottleneck of this kernel and
    // lanes of one warp that hit the same bin serialise (the high digits of genome coordinates are far from uniform)
    __shared__ uint32_t sh_all[RS_HIST_COPIES][RS_MAX_PASSES][RS_RADIX];
    for (int i = threadIdx.x; i < RS_HIST_COPIES * RS_MAX_PASSES * RS_RADIX; i += RS_HIST_THREADS) (&sh_all[0][0][0])[i] = 0;
    uint32_t (*sh)[RS_RADIX] = sh_all[threadIdx.x & (RS_HIST_COPIES - 1)];
    __syncthreads();
    // scalar head up to the first 16-byte boundary (sub-ranges of a larger array need not be aligned), vector body, scalar tail
    const int64_t head = min(n, (int64_t)((16 - (reinterpret_cast<uintptr_t>(keys) & 15)) & 15) >> 2);
    const int64_t n_vec = (n - head) >> 2;
    const uint4* kv = reinterpret_cast<const uint4*>(keys + head);
    for (int64_t i = (int64_t)blockIdx.x * RS_HIST_THREADS + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * RS_HIST_THREADS) {
        uint4 k = ld_stream_u32x4(kv + i);
        uint32_t kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int p = 0; p < RS_MAX_PASSES; p++)
                if (p < n_passes) atomicAdd(&sh[p][(kk[j] >> (begin_bit + 8 * p)) & 255u], 1u);
        }
    }
    if (blockIdx.x == 0) {   // at most 3 head keys and 3 tail keys
        const int64_t tail0 = head + (n_vec << 2);
        const int64_t i = threadIdx.x < head ? (int64_t)threadIdx.x : tail0 + ((int64_t)threadIdx.x - head);
        if (i < n && threadIdx.x < 8) {
            uint32_t k = keys[i];
            for (int p = 0; p < n_passes; p++) atomicAdd(&sh[p][(k >> (begin_bit + 8 * p)) & 255u], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_passes * RS_RADIX; i += RS_HIST_THREADS) {
        uint32_t c = 0;
#pragma unroll
        for (int q = 0; q < RS_HIST_COPIES; q++) c += (&sh_all[q][0][0])[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// ---- one digit pass ----------------------------------------------------------------------------------
// IDENTITY_VALS: payload of key i is i itself (first pass; no payload array is read).
// GATHER: the last pass of an index build also fetches gather_src[payload] (the row's linear end) into
// gather_dst at the row's sorted position, which saves the separate gather pass.
// IPT keys per thread: 16 (4096-key tiles) for large inputs; 8 (2048-key tiles) when the input has fewer tiles than the
// GPU has resident CTA slots, where the pass time is one tile's lifetime and shorter tiles finish sooner.
// EXTRA: a second 32-bit payload (extra_in -> extra_out) travels with every row.  Used for the linear ends of tables too large
// for the L2: gathering them afterwards through the sorted row ids is one random DRAM sector per row (5 ms at 200 M rows).
template <bool IDENTITY_VALS, bool GATHER, int IPT, bool EXTRA = false>
__global__ void __launch_bounds__(RS_THREADS, RS_THREADS * IPT >= 4096 ? GIQLB_RS_CTAS_PER_SM : 6)
rs_onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                   const uint32_t* __restrict__ gather_src /* EXTRA: extra_in */, uint32_t* __restrict__ gather_dst /* EXTRA: extra_out */,
                   const uint32_t* __restrict__ ghist /* this pass, [256] */,
                   uint32_t* __restrict__ status /* this pass, [tiles][256] */,
                   uint32_t* __restrict__ ticket, int* __restrict__ err_flag) {
    __shared__ rs_hist_t s_warp_hist[RS_WARPS][RS_RADIX];   // per-warp digit counters -> warp-exclusive offsets
    __shared__ uint32_t s_keys[(RS_THREADS * IPT)];
    __shared__ uint32_t s_vals[(RS_THREADS * IPT)];
    __shared__ uint32_t s_digit_excl[RS_RADIX];            // tile-local exclusive digit offsets
    __shared__ int32_t s_global_base[RS_RADIX];            // global position of local sorted slot 0 of each digit (n < 2^30)
    __shared__ uint32_t s_scan[RS_WARPS + 2];
    __shared__ uint32_t s_tile;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * (RS_THREADS * IPT);
    const int tile_n = (int)min((int64_t)(RS_THREADS * IPT), n - tile_base);
    const uint32_t w = warp_id(), l = lane_id();
    const uint32_t lt_mask = (1u << l) - 1u;

    // load: warp w owns keys [w*512, w*512+512) of the tile; item i of lane l is w*512 + i*32 + l
    uint32_t key[IPT];
    uint32_t rank[IPT];
    const int warp_off = w * (32 * IPT);
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        int li = warp_off + i * 32 + l;
        key[i] = li < tile_n ? ld_stream_u32(keys_in + tile_base + li) : 0xffffffffu;
    }
    // rank within the warp, stable in (item, lane) order.  All match instructions are issued first (they are
    // independent), so only the shared-memory counter update remains on the serial chain.
#pragma unroll
    for (int i = 0; i < IPT; i++) {
#if GIQLB_RS_USE_MATCH
        rank[i] = __match_any_sync(FULL_MASK, (key[i] >> shift) & 255u);
#else
        // lanes holding the same digit, from eight ballots: the hardware MATCH instruction has a long latency and
        // a low issue rate, the votes pipeline freely
        const uint32_t d = key[i] >> shift;
        uint32_t peers = FULL_MASK;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            // four instructions per bit (LOP3 -> predicate, VOTE, SEL, LOP3); written in PTX because the compiler
            // otherwise canonicalises the bit test into shift + and + compare (six per bit)
            uint32_t m, own;
            asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                "and.b32 t, %2, %3;\n\t"
                "setp.ne.u32 p, t, 0;\n\t"
                "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
                "selp.b32 %1, 0xffffffff, 0, p;\n\t}"
                : "=r"(m), "=r"(own) : "r"(d), "r"(1u << b));
            peers &= ~(m ^ own);
        }
        rank[i] = peers;
#endif
    }
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t peers = rank[i];
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)l == leader) {
            base = s_warp_hist[w][d];
            s_warp_hist[w][d] = (rs_hist_t)(base + __popc(peers));
        }
        base = __shfl_sync(FULL_MASK, base, leader);
        rank[i] = base + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // thread d (the first 256 threads): turn per-warp counts into warp-exclusive offsets, get the tile's count of digit d
    const bool owns_digit = RS_THREADS == RS_RADIX || threadIdx.x < RS_RADIX;
    const uint32_t d_own = owns_digit ? threadIdx.x : 0u;
    uint32_t count = 0;
    uint32_t* my_status = status + (size_t)tile * RS_RADIX + d_own;
    uint32_t pre[GIQLB_RS_LOOKBACK];
    if (owns_digit) {
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            uint32_t c = s_warp_hist[ww][d_own];
            s_warp_hist[ww][d_own] = (rs_hist_t)count;
            count += c;
        }
        if (d_own == 255u) count -= (uint32_t)((RS_THREADS * IPT) - tile_n);   // padding keys are all-ones -> digit 255
        if (tile == 0) {
            *(volatile uint32_t*)my_status = ST_FLAG_INCL | count;
        } else {
            *(volatile uint32_t*)my_status = ST_FLAG_AGG | count;
        }
        // first look-back step issued now: its L2 round trip overlaps the two block scans and the shared-memory scatter
#pragma unroll
        for (int j = 0; j < GIQLB_RS_LOOKBACK; j++)
            pre[j] = (int64_t)tile - 1 - j >= 0 ? *(volatile uint32_t*)(status + (size_t)(tile - 1 - j) * RS_RADIX + d_own) : (2u << 30);
    }
    // tile-local exclusive offsets over digits, and global start of each digit bin (threads without a digit add nothing)
    uint32_t tot;
    uint32_t dex = block_excl_scan_add<uint32_t>(count, s_scan, &tot);
    __syncthreads();
    uint32_t gtot;
    uint32_t gstart = block_excl_scan_add<uint32_t>(owns_digit ? ghist[d_own] : 0u, s_scan, &gtot);
    if (owns_digit) s_digit_excl[d_own] = dex;

    __syncthreads();   // s_digit_excl and the warp-exclusive offsets are complete

    // scatter the tile into shared memory in digit order FIRST: it needs only tile-local offsets, and it gives the
    // preceding tiles time to publish their counts before this tile looks back
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        uint32_t d = (key[i] >> shift) & 255u;
        uint32_t pos = s_digit_excl[d] + s_warp_hist[w][d] + rank[i];
        int li = warp_off + i * 32 + l;
        uint32_t v;
        if (IDENTITY_VALS) v = (uint32_t)(tile_base + li);
        else v = li < tile_n ? ld_stream_u32(vals_in + tile_base + li) : 0u;
        s_keys[pos] = key[i];
        s_vals[pos] = v;
    }

    // decoupled look-back over earlier tiles for digit d_own; four predecessors are fetched per step so that the
    // chain of dependent L2 round trips is four times shorter
    uint32_t prefix = 0;
    if (tile > 0 && owns_digit) {
        int64_t t = (int64_t)tile - 1;
        uint32_t spins = 0;
        bool done = false;
        // first step: the four words prefetched above; then GIQLB_RS_LOOKBACK_WIDE predecessors per round trip
        if (!lookback_step<GIQLB_RS_LOOKBACK>(pre, t, prefix, done) && !done) __nanosleep(GIQLB_RS_BACKOFF_NS);
        while (!done) {
            uint32_t sv[GIQLB_RS_LOOKBACK_WIDE];
#pragma unroll
            for (int j = 0; j < GIQLB_RS_LOOKBACK_WIDE; j++)
                sv[j] = t - j >= 0 ? *(volatile uint32_t*)(status + (size_t)(t - j) * RS_RADIX + d_own) : (2u << 30);   // = ST_FLAG_INCL | 0
            if (!lookback_step<GIQLB_RS_LOOKBACK_WIDE>(sv, t, prefix, done) && !done) {
                if (++spins > RS_SPIN_LIMIT) { *err_flag = 1; break; }
                __nanosleep(GIQLB_RS_BACKOFF_NS);
            }
        }
        *(volatile uint32_t*)my_status = ST_FLAG_INCL | (prefix + count);
    }
    if (owns_digit) s_global_base[d_own] = (int32_t)(gstart + prefix - dex);
    __syncthreads();
    // write out: consecutive slots of one digit are consecutive in global memory
    for (int j = threadIdx.x; j < tile_n; j += RS_THREADS) {
        uint32_t k = s_keys[j];
        uint32_t d = (k >> shift) & 255u;
        const int64_t out = (int64_t)(s_global_base[d] + j);
        const uint32_t v = s_vals[j];
        keys_out[out] = k;
        vals_out[out] = v;
        if (GATHER) gather_dst[out] = __ldg(gather_src + v);
    }
    if (EXTRA) {
        // the second payload takes the same route through s_vals once the first has left it
        __syncthreads();
#pragma unroll
        for (int i = 0; i < IPT; i++) {
            const uint32_t d = (key[i] >> shift) & 255u;
            const uint32_t pos = s_digit_excl[d] + s_warp_hist[w][d] + rank[i];
            const int li = warp_off + i * 32 + l;
            s_vals[pos] = li < tile_n ? ld_stream_u32(gather_src + tile_base + li) : 0u;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < tile_n; j += RS_THREADS) {
            const uint32_t d = (s_keys[j] >> shift) & 255u;
            gather_dst[(int64_t)(s_global_base[d] + j)] = s_vals[j];
        }
    }
}

}  // namespace

// Sorts n (key, value) pairs on bits [begin_bit, end_bit).  keys_a/vals_a hold the input; the two buffer
// pairs ping-pong and *result_in_b says where the output landed.  identity_vals: the payload of key i is
// i itself (vals_a is scratch and is not read by the first pass).
int giqlb_radix_sort_pairs_u32(giqlb_ctx* ctx, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b,
                               uint32_t* vals_b, int64_t n, int begin_bit, int end_bit, int identity_vals,
                               int* result_in_b, const uint32_t* hist_dev, const uint32_t* gather_src, uint32_t* gather_dst,
                               uint32_t* extra_a, uint32_t* extra_b) {
    *result_in_b = 0;
    if (n >= (int64_t)ST_VALUE_MASK) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "radix sort: n=%lld too large", (long long)n);
    int n_passes = (end_bit - begin_bit + 7) / 8;
    if (n_passes < 1) n_passes = 1;   // always run one pass so the payload is materialised
    if (n_passes > RS_MAX_PASSES) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "radix sort: too many passes");
    if (n == 0) return GIQLB_OK;
    // fewer 4096-key tiles than SMs: halve the tile (measured: 8 passes over 250 k keys 101 -> 89 us; at 1.25 M keys, i.e.
    // between one tile per SM and one per resident CTA slot, the halved tiles are SLOWER — 20.6 vs 18.7 us per pass)
    const bool small = ceil_div(n, RS_TILE) < (int64_t)ctx->sm_count;
    const int64_t tiles = ceil_div(n, small ? RS_TILE / 2 : RS_TILE);

    uint32_t* ghist = nullptr;    // [n_passes][256] | tickets[RS_MAX_PASSES] | status [n_passes][tiles][256]
    const size_t head_words = (size_t)n_passes * RS_RADIX + RS_MAX_PASSES;
    const size_t words = head_words + (size_t)n_passes * tiles * RS_RADIX;
    GIQLB_TRY(dev_alloc(ctx, &ghist, words));
    CUDA_TRY(ctx, cudaMemsetAsync(ghist, 0, words * sizeof(uint32_t), ctx->stream));
    uint32_t* tickets = ghist + (size_t)n_passes * RS_RADIX;
    uint32_t* status = ghist + head_words;

    if (hist_dev && begin_bit % 8 == 0 && begin_bit / 8 + n_passes <= RS_MAX_PASSES) {
        // the producer of the keys already counted every digit (index_keygen_kernel): [4][256] for bits 0..31
        CUDA_TRY(ctx, cudaMemcpyAsync(ghist, hist_dev + (size_t)(begin_bit / 8) * RS_RADIX, (size_t)n_passes * RS_RADIX * sizeof(uint32_t),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        int64_t hb_want = ceil_div(n, RS_HIST_THREADS * 16), hb_cap = (int64_t)ctx->sm_count * 8;
        int hist_blocks = (int)(hb_want < hb_cap ? hb_want : hb_cap);
        PhaseSpan ps(ctx, GIQLB_PH_SORT_HIST);
        rs_histogram_kernel<<<hist_blocks, RS_HIST_THREADS, 0, ctx->stream>>>(keys_a, n, begin_bit, n_passes, ghist);
        LAUNCH_CHECK(ctx);
    }

    // extra_a / extra_b: a second payload that ping-pongs with the keys (its result lies beside them: extra_b iff *result_in_b)
    const bool extra = extra_a && extra_b;
    if (extra && gather_src) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "radix sort: a carried payload and a gathered one exclude each other");
    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b, *ein = extra_a, *eout = extra_b;
    PhaseSpan ps_passes(ctx, GIQLB_PH_SORT_PASS);
    for (int p = 0; p < n_passes; p++) {
        const int shift = begin_bit + 8 * p;
        const uint32_t* gh = ghist + (size_t)p * RS_RADIX;
        uint32_t* st = status + (size_t)p * tiles * RS_RADIX;
        const bool gather = gather_src && gather_dst && p == n_passes - 1;
#define RS_LAUNCH(IDV, GTH)                                                                                                                     \
    do {                                                                                                                                        \
        if (small)                                                                                                                              \
            rs_onesweep_kernel<IDV, GTH, RS_IPT / 2><<<(unsigned)tiles, RS_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, gather_src,  \
                                                                                                      gather_dst, gh, st, tickets + p, ctx->d_err); \
        else                                                                                                                                    \
            rs_onesweep_kernel<IDV, GTH, RS_IPT><<<(unsigned)tiles, RS_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, gather_src,      \
                                                                                                  gather_dst, gh, st, tickets + p, ctx->d_err);    \
    } while (0)
#define RS_LAUNCH_EXTRA(IDV)                                                                                                                    \
    do {                                                                                                                                        \
        if (small)                                                                                                                              \
            rs_onesweep_kernel<IDV, false, RS_IPT / 2, true><<<(unsigned)tiles, RS_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, ein, eout, \
                                                                                                              gh, st, tickets + p, ctx->d_err);  \
        else                                                                                                                                    \
            rs_onesweep_kernel<IDV, false, RS_IPT, true><<<(unsigned)tiles, RS_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, ein, eout,     \
                                                                                                          gh, st, tickets + p, ctx->d_err);      \
    } while (0)
        if (extra) {
            if (p == 0 && identity_vals) RS_LAUNCH_EXTRA(true); else RS_LAUNCH_EXTRA(false);
        } else if (p == 0 && identity_vals) {
            if (gather) RS_LAUNCH(true, true); else RS_LAUNCH(true, false);
        } else {
            if (gather) RS_LAUNCH(false, true); else RS_LAUNCH(false, false);
        }
#undef RS_LAUNCH
#undef RS_LAUNCH_EXTRA
        LAUNCH_CHECK(ctx);
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
        t = ein; ein = eout; eout = t;
        *result_in_b ^= 1;
    }
    dev_free(ctx, ghist);
    return GIQLB_OK;
}
