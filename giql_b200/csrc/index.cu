// index.cu — genome layout (linearisation of (chrom code, position) into one uint32 coordinate) and the
// sorted interval index built with the LSD radix sort.
//
// Replaces, for the SQL emitted at src/giql/expanders/intersects_duckdb.py:1283-1299 (reference), the
// per-chromosome `WHERE chrom = 'c'` re-scans and per-chromosome sorts of DuckDB's IE_JOIN: folding the
// chromosome into the coordinate makes ONE sort and ONE search space serve all chromosomes, while the
// guard cells between chromosomes keep `l.chrom = r.chrom` (src/giql/expanders/intersects.py:151) exact.
// Canonical offsets (src/giql/canonical.py:16-52) are applied here as integers.
#include "common.cuh"
#include "scan.cuh"

namespace {

#ifndef GIQLB_SORT_GATHERS_END
#define GIQLB_SORT_GATHERS_END 0
#endif
// 1: the key generation kernel also counts the four radix digit histograms of the start keys (shared-memory
// atomics), which saves the sort's own histogram pass over the keys (0.061 ms at 2 x 10M rows against +0.013 ms in
// keygen).  Round 1 measured this slower with the scalar keygen; with the 4-rows-per-thread keygen it wins.
#ifndef GIQLB_SORT_CARRY_END_MIN
#define GIQLB_SORT_CARRY_END_MIN (24ll << 20)   // rows: three 4-byte columns no longer fit the 126 MB L2
#endif
#ifndef GIQLB_KEYGEN_HIST
#define GIQLB_KEYGEN_HIST 1
#endif
constexpr int LY_THREADS = 256;
constexpr int LY_SMEM_CODES = 2048;

// per-chromosome min/max of canonical coordinates.  Shared-memory staging with a read-before-atomic
// test: after the first few rows almost no row improves a bound, so almost no atomics are issued.
// REMAP: the code column still holds the dictionary encoder's parked table slots; they are translated through
// slot2code here and written back, which saves the separate remap pass over the column (the kernel does nothing
// when the device-side ranking declined — the caller then redoes the encode on the host-ordered path).
template <bool REMAP>
__global__ void __launch_bounds__(LY_THREADS)
layout_observe_kernel(int32_t* code, const int32_t* __restrict__ start,
                      const int32_t* __restrict__ end, int64_t n, int32_t n_codes, int32_t sadj, int32_t eadj,
                      long long* __restrict__ lohi, unsigned long long* __restrict__ obs /* [2] degenerate, negative */,
                      unsigned long long* __restrict__ max_width /* widest interval of the table (cells) */, int* __restrict__ err,
                      const int32_t* __restrict__ slot2code, const int* __restrict__ ok_flag) {
    if (REMAP && !*ok_flag) return;
    __shared__ int32_t s_lo[LY_SMEM_CODES];
    __shared__ int32_t s_hi[LY_SMEM_CODES];
    const bool use_smem = n_codes <= LY_SMEM_CODES;
    if (use_smem) {
        for (int i = threadIdx.x; i < n_codes; i += LY_THREADS) {
            s_lo[i] = INT32_MAX;
            s_hi[i] = INT32_MIN;
        }
        __syncthreads();
    }
    bool bad = false;
    unsigned int degenerate = 0, negative = 0;
    long long widest = 0;
    auto fold = [&](int32_t c, int32_t s0, int32_t e0) {
        long long s = (long long)s0 + sadj, e = (long long)e0 + eadj;
        if (e - s > widest) widest = e - s;
        long long lo = s < e ? s : e, hi = s < e ? e : s;
        if (c < 0 || c >= n_codes || lo < -(1ll << 30) || hi > (1ll << 31) - 4) {
            bad = true;
            return;
        }
        degenerate += e <= s;
        negative += e < s;
        if (use_smem) {
            if ((int32_t)lo < s_lo[c]) atomicMin(&s_lo[c], (int32_t)lo);
            if ((int32_t)hi > s_hi[c]) atomicMax(&s_hi[c], (int32_t)hi);
        } else {
            if (lo < *(volatile long long*)&lohi[c]) atomicMin(&lohi[c], lo);
            if (hi > *(volatile long long*)&lohi[n_codes + c]) atomicMax(&lohi[n_codes + c], hi);
        }
    };
    // 16-byte vector loads when the three columns are aligned (they are when they come from one allocation each)
    const bool vec = ((reinterpret_cast<uintptr_t>(code) | reinterpret_cast<uintptr_t>(start) | reinterpret_cast<uintptr_t>(end)) & 15) == 0;
    const int64_t n_vec = vec ? n >> 2 : 0;
    for (int64_t i = (int64_t)blockIdx.x * LY_THREADS + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * LY_THREADS) {
        int4 c = reinterpret_cast<const int4*>(code)[i];
        if (REMAP) {
            c = make_int4(__ldg(slot2code + c.x), __ldg(slot2code + c.y), __ldg(slot2code + c.z), __ldg(slot2code + c.w));
            reinterpret_cast<int4*>(code)[i] = c;
        }
        const int4 s = __ldg(reinterpret_cast<const int4*>(start) + i);
        const int4 e = __ldg(reinterpret_cast<const int4*>(end) + i);
        fold(c.x, s.x, e.x);
        fold(c.y, s.y, e.y);
        fold(c.z, s.z, e.z);
        fold(c.w, s.w, e.w);
    }
    for (int64_t i = (n_vec << 2) + (int64_t)blockIdx.x * LY_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LY_THREADS) {
        int32_t c = code[i];
        if (REMAP) code[i] = c = __ldg(slot2code + c);
        fold(c, start[i], end[i]);
    }
    if (bad) err[1] = 1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) widest = max(widest, __shfl_xor_sync(FULL_MASK, widest, d));
    if (lane_id() == 0 && widest > 0 && (unsigned long long)widest > *(volatile unsigned long long*)max_width)
        atomicMax(max_width, (unsigned long long)widest);
    if (__any_sync(FULL_MASK, degenerate != 0)) {
        degenerate = warp_reduce_add(degenerate);
        negative = warp_reduce_add(negative);
        if (lane_id() == 0) {
            atomicAdd(&obs[0], (unsigned long long)degenerate);
            if (negative) atomicAdd(&obs[1], (unsigned long long)negative);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_codes; i += LY_THREADS) {
            if (s_lo[i] != INT32_MAX) atomicMin(&lohi[i], (long long)s_lo[i]);
            if (s_hi[i] != INT32_MIN) atomicMax(&lohi[n_codes + i], (long long)s_hi[i]);
        }
    }
}

__global__ void layout_init_kernel(long long* lohi, int32_t n_codes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_codes) {
        lohi[i] = INT64_MAX;
        lohi[n_codes + i] = INT64_MIN;
    }
}

// base[] = exclusive scan of per-chromosome cell counts; one CTA, chunked with a carry
__global__ void __launch_bounds__(1024)
layout_finalize_kernel(const long long* __restrict__ lohi, int32_t n_codes, long long pad,
                       uint32_t* __restrict__ base, int32_t* __restrict__ lo32, int32_t* __restrict__ hi32,
                       unsigned long long* __restrict__ total, int* __restrict__ err) {
    __shared__ unsigned long long scratch[1024 / 32 + 2];
    unsigned long long carry = 0;
    for (int b0 = 0; b0 < n_codes; b0 += 1024) {
        int i = b0 + threadIdx.x;
        unsigned long long cells = 0;
        if (i < n_codes) {
            long long lo = lohi[i], hi = lohi[n_codes + i];
            if (lo > hi) {   // chromosome never observed
                lo32[i] = 0;
                hi32[i] = -1;
                cells = 0;
            } else {
                lo32[i] = (int32_t)lo;
                hi32[i] = (int32_t)hi;
                cells = (unsigned long long)(hi - lo) + 3ull + (unsigned long long)pad;
            }
        }
        unsigned long long tot;
        unsigned long long ex = block_excl_scan_add<unsigned long long>(cells, scratch, &tot);
        if (i < n_codes) {
            unsigned long long b = carry + ex;
            base[i] = b > 0xffffffffull ? 0xffffffffu : (uint32_t)b;
        }
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        base[n_codes] = carry > 0xffffffffull ? 0xffffffffu : (uint32_t)carry;
        *total = carry;
        if (carry >= 0xffffffffull) err[1] = 2;   // the summed extent does not fit the 32-bit linear coordinate
    }
}

// linearised start key + linearised end (unsorted), degenerate-row counters
template <bool HIST>
__global__ void __launch_bounds__(256)
index_keygen_kernel(const int32_t* __restrict__ code, const int32_t* __restrict__ start,
                    const int32_t* __restrict__ end, const int32_t* __restrict__ rows /* nullable: row selection */,
                    int64_t n, int32_t n_codes, int32_t sadj, int32_t eadj,
                    const uint32_t* __restrict__ base, const int32_t* __restrict__ lo32,
                    const int32_t* __restrict__ hi32, uint32_t* __restrict__ gs, uint32_t* __restrict__ ge,
                    unsigned long long* __restrict__ counters /* [0] end<=start [1] end<start [2] outside layout */,
                    uint32_t* __restrict__ ghist /* nullable: [4][256] digit counts of the start keys, for the sort */,
                    int* __restrict__ err) {
    __shared__ uint32_t s_hist[HIST ? 4 : 1][256];
    if (HIST) {
        for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&s_hist[0][0])[i] = 0;
        __syncthreads();
    }
    unsigned int degenerate = 0, negative = 0, outside = 0;
    auto one = [&](int32_t c, int32_t s0, int32_t e0, uint32_t& ks, uint32_t& ke) {
        const long long s = (long long)s0 + sadj, e = (long long)e0 + eadj;
        ks = 0xffffffffu;
        ke = 0xffffffffu;
        if (c < 0 || c >= n_codes) {
            outside++;
            return;
        }
        const long long lo = __ldg(lo32 + c), hi = __ldg(hi32 + c);
        if (s < lo || s > hi || e < lo || e > hi) {
            outside++;
            return;
        }
        const uint32_t b = __ldg(base + c) + 1u;
        ks = b + (uint32_t)(s - lo);
        ke = b + (uint32_t)(e - lo);
        degenerate += e <= s;
        negative += e < s;
        if (HIST) {
            atomicAdd(&s_hist[0][ks & 255u], 1u);
            atomicAdd(&s_hist[HIST ? 1 : 0][(ks >> 8) & 255u], 1u);
            atomicAdd(&s_hist[HIST ? 2 : 0][(ks >> 16) & 255u], 1u);
            atomicAdd(&s_hist[HIST ? 3 : 0][ks >> 24], 1u);
        }
    };
    // whole tables (no row selection) with aligned columns: four rows per thread and iteration, 16-byte loads and stores
    const bool vec = !rows && ((reinterpret_cast<uintptr_t>(code) | reinterpret_cast<uintptr_t>(start) | reinterpret_cast<uintptr_t>(end) |
                                reinterpret_cast<uintptr_t>(gs) | reinterpret_cast<uintptr_t>(ge)) & 15) == 0;
    const int64_t n_vec = vec ? n >> 2 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 c = __ldg(reinterpret_cast<const int4*>(code) + i), s4 = __ldg(reinterpret_cast<const int4*>(start) + i),
                   e4 = __ldg(reinterpret_cast<const int4*>(end) + i);
        uint4 ks, ke;
        one(c.x, s4.x, e4.x, ks.x, ke.x);
        one(c.y, s4.y, e4.y, ks.y, ke.y);
        one(c.z, s4.z, e4.z, ks.z, ke.z);
        one(c.w, s4.w, e4.w, ks.w, ke.w);
        reinterpret_cast<uint4*>(gs)[i] = ks;
        reinterpret_cast<uint4*>(ge)[i] = ke;
    }
    for (int64_t i = (n_vec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = rows ? (int64_t)rows[i] : i;
        uint32_t ks, ke;
        one(code[r], start[r], end[r], ks, ke);
        gs[i] = ks;
        ge[i] = ke;
    }
    if (HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
            const uint32_t c = (&s_hist[0][0])[i];
            if (c) atomicAdd(&ghist[i], c);
        }
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

template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, T* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}

// Finishing pass of the short-window sort (GIQLB_INDEX_SORT_SHORT_WINDOWS): k3 / v3 are ordered (stably) by key >> 8.  A row's
// final position inside its window of equal upper bits is the number of window rows with a smaller key plus the earlier rows
// with the same key, which is exactly where the fourth stable radix pass would have put it.  Windows are short (2 rows on
// average at 10 M rows over a human genome), so a warp ranks 32 consecutive rows at once with votes: one ballot marks the window
// heads, eight ballots over the bits of the low byte give every lane the lanes of its window that hold a smaller / an equal
// byte.  The warp then moves on to the last head it saw, so no window is ever cut by a group boundary and nothing loops per row
// (measured: rows walking left and right through shared memory cost 223 instructions per row).  Windows of 32 rows or more are
// left as they are and counted in counters[3]: the build is then unfinished and the caller rebuilds with the full sort.  The same
// kernel fetches every row's linear end, which replaces the separate gather over the sorted row ids.
constexpr int SF_THREADS = 256;
constexpr int SF_WARPS = SF_THREADS / 32;
constexpr int SF_IPT = 8;
constexpr int SF_TILE = SF_THREADS * SF_IPT;
constexpr int SF_HALO = 32;          // rows after the tile that a window begun inside the tile may reach
constexpr int SF_CHUNK = SF_TILE / SF_WARPS;
constexpr int SF_LEAD = 4;           // sk[SF_LEAD + e] = key of tile row e (16-byte aligned for vector staging); sk[SF_LEAD - 1] = the row before
constexpr uint32_t SF_NONE = 0xffffu;

// one step of the bit-serial comparison of the low bytes inside a window: lanes whose byte has bit `mask` set vote; lanes of my
// window equal to me so far (eq) that lack a bit I have are smaller (lt), those that differ from me in this bit leave eq
__device__ __forceinline__ void sf_vote_bit(uint32_t k, uint32_t mask, uint32_t& eq, uint32_t& lt) {
    const bool bit = (k & mask) != 0u;
    const uint32_t m = __ballot_sync(FULL_MASK, bit);
    const uint32_t own = bit ? 0xffffffffu : 0u;
    lt |= eq & ~m & own;
    eq &= ~(m ^ own);
}

__global__ void __launch_bounds__(SF_THREADS)
sort_finish_kernel(const uint32_t* __restrict__ k3, const uint32_t* __restrict__ v3, int64_t n, const uint32_t* __restrict__ ge_rows,
                   uint32_t* __restrict__ gs, uint32_t* __restrict__ rid, uint32_t* __restrict__ ge,
                   unsigned long long* __restrict__ counters) {
    __shared__ __align__(16) uint32_t sk[SF_LEAD + SF_TILE + SF_HALO + 32];   // + 32: a group may read (and ignore) past the halo
    __shared__ __align__(4) uint16_t dst[SF_TILE + SF_HALO];   // tile-relative destination of row e, SF_NONE: the row belongs to the previous tile
    const int64_t base = (int64_t)blockIdx.x * SF_TILE;
    const int rows_here = (int)min((int64_t)(SF_TILE + SF_HALO), n - base);   // rows of the array staged in sk
    const uint32_t* const k_t = k3 + base;
    if (rows_here == SF_TILE + SF_HALO) {
        for (int j = threadIdx.x; j < (SF_TILE + SF_HALO) / 4; j += SF_THREADS)
            reinterpret_cast<uint4*>(sk + SF_LEAD)[j] = __ldg(reinterpret_cast<const uint4*>(k_t) + j);
    } else {
        for (int j = threadIdx.x; j < rows_here; j += SF_THREADS) sk[SF_LEAD + j] = __ldg(k_t + j);
    }
    // the row before the tile decides whether tile row 0 begins a window; before row 0 of the array: anything with other upper bits
    if (threadIdx.x == 0) sk[SF_LEAD - 1] = base > 0 ? __ldg(k_t - 1) : ~__ldg(k_t);
    for (int j = threadIdx.x; j < (SF_TILE + SF_HALO) / 2; j += SF_THREADS) reinterpret_cast<uint32_t*>(dst)[j] = SF_NONE | (SF_NONE << 16);
    // the row ids and (random gathers) the ends of the tile's rows are requested now and arrive while the ranks are worked out
    const uint32_t* const v_t = v3 + base;
    uint32_t val[SF_IPT], end_of[SF_IPT];
#pragma unroll
    for (int r = 0; r < SF_IPT; r++) {
        const int e = r * SF_THREADS + threadIdx.x;
        val[r] = e < rows_here ? ld_stream_u32(v_t + e) : 0u;
    }
#pragma unroll
    for (int r = 0; r < SF_IPT; r++) end_of[r] = __ldg(ge_rows + val[r]);   // (row 0's end for slots past the array: never stored)
    __syncthreads();

    // ---- ranks: warp w owns the windows whose first row lies in its chunk of the tile
    const int lane = (int)lane_id();
    const int chunk_end = ((int)warp_id() + 1) * SF_CHUNK;
    const uint32_t below = (1u << lane) - 1u, upto = below | (1u << lane);
    const uint32_t* const skp = sk + SF_LEAD + lane;
    uint16_t* const dstp = dst + lane;
    int p = chunk_end - SF_CHUNK;
    bool aligned = false;      // lane 0 of the group is the first row of a window
    bool own = false;          // rows passed over while not aligned are this warp's (a window of 32 rows or more) or an earlier warp's
    unsigned int dense = 0;
    while (p < rows_here && (p < chunk_end || (own && !aligned))) {
        const uint32_t k = skp[p];
        const bool head = p + lane >= rows_here || (k ^ skp[p - 1]) > 255u;    // a slot past the array closes the last window
        const uint32_t heads = __ballot_sync(FULL_MASK, head);
        if (!aligned) {
            const int skip = heads ? __ffs(heads) - 1 : 32;
            if (own && lane < skip) {      // the tail of a window too long to order here: the rows stay where they are
                dstp[p] = (uint16_t)(p + lane);
                dense++;
            }
            p += skip;
            aligned = skip < 32;
            continue;
        }
        const int q = 31 - __clz(heads);   // rows [0, q) of the group are whole windows (lane 0 is a head)
        if (q == 0) {
            dstp[p] = (uint16_t)(p + lane);   // all 32 lanes are rows of one window (a slot past the array would be a head)
            dense++;
            p += 32;
            aligned = false;
            own = true;
            continue;
        }
        const int ws = 31 - __clz(heads & upto);                 // head of this lane's window
        const int we = lane + __ffs((heads >> 1) >> lane);       // head of the next window (lanes < q have one)
        uint32_t eq = ((1u << we) - 1u) & ~((1u << ws) - 1u);    // lanes of the window with the same byte so far (junk for lanes >= q)
        uint32_t lt = 0;                                         // ... with a smaller byte
        sf_vote_bit(k, 0x80u, eq, lt);
        sf_vote_bit(k, 0x40u, eq, lt);
        sf_vote_bit(k, 0x20u, eq, lt);
        sf_vote_bit(k, 0x10u, eq, lt);
        sf_vote_bit(k, 0x08u, eq, lt);
        sf_vote_bit(k, 0x04u, eq, lt);
        sf_vote_bit(k, 0x02u, eq, lt);
        sf_vote_bit(k, 0x01u, eq, lt);
        if (lane < q && p + lane < rows_here) dstp[p] = (uint16_t)(p + ws + __popc(lt) + __popc(eq & below));   // (slots past the array are heads too)
        p += q;
    }
    dense = warp_reduce_add(dense);
    if (lane == 0 && dense) atomicAdd(&counters[3], (unsigned long long)dense);
    __syncthreads();

    // ---- move: row e of the tile (and of the rows after it that this tile ranked) goes to dst[e], with its row id and its end
    uint32_t* const gs_t = gs + base;
    uint32_t* const rid_t = rid + base;
    uint32_t* const ge_t = ge + base;
#pragma unroll
    for (int r = 0; r < SF_IPT; r++) {
        const int e = r * SF_THREADS + threadIdx.x;
        const uint32_t to = dst[e];
        if (to == SF_NONE) continue;
        gs_t[to] = sk[SF_LEAD + e];
        rid_t[to] = val[r];
        ge_t[to] = end_of[r];
    }
    if (threadIdx.x < SF_HALO) {       // rows after the tile that belong to a window begun inside it
        const int e = SF_TILE + threadIdx.x;
        const uint32_t t = dst[e];
        if (t != SF_NONE) {
            const uint32_t v = v_t[e];
            gs_t[t] = sk[SF_LEAD + e];
            rid_t[t] = v;
            ge_t[t] = __ldg(ge_rows + v);
        }
    }
}

__global__ void __launch_bounds__(256)
remap_rows_kernel(int32_t* __restrict__ rid, const int32_t* __restrict__ rows, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        rid[i] = rows[rid[i]];
}

__global__ void __launch_bounds__(256)
strand_histogram_kernel(const uint8_t* __restrict__ strand, int64_t n, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[strand[i]], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// stable partition of the sorted rows by strand code: pos = block_off[code] + rank among equal codes.
// The rank is an exclusive scan of a one-hot per code; with few distinct codes in practice we scan one
// code at a time (launch count = number of codes present).
struct InIsCode {
    const uint8_t* s;
    uint8_t code;
    __device__ unsigned long long operator()(int64_t i) const { return s[i] == code ? 1ull : 0ull; }
};
struct OutScatterRow {
    const uint32_t *gs, *ge;
    const int32_t* rid;
    const uint8_t* strand;
    uint32_t *gs_o, *ge_o;
    int32_t* rid_o;
    uint8_t* strand_o;
    int64_t dst0;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        if (v) {
            int64_t d = dst0 + (int64_t)excl;
            gs_o[d] = gs[i];
            ge_o[d] = ge[i];
            rid_o[d] = rid[i];
            strand_o[d] = strand[i];
        }
    }
};

// Bucket directory of an ascending array in one coalesced pass: row j, the first row of its bucket run, fills every
// entry from the bucket after its predecessor's up to its own (tab[b] = first position with a value >= b << shift);
// the thread of j == n fills the tail with n.  Longer gaps of empty buckets are filled by the whole warp, 32 entries per step,
// so one long gap does not serialise on a single thread (and short ones do not serialise the warp: a warp-cooperative turn
// for every lane made the kernel 32 steps per 32 rows).  Values beyond the directory fall into its last bucket.
__global__ void __launch_bounds__(256)
bucket_dir_kernel(const uint32_t* __restrict__ a, int64_t n, int shift, uint32_t stride, uint32_t* __restrict__ tab) {
    const long long last_entry = (long long)stride - 1, last_bucket = (long long)stride - 2;
    const int64_t span = ((n + 1 + 31) / 32) * 32;                     // whole warps stay in the loop together
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < span; j += (int64_t)gridDim.x * blockDim.x) {
        long long first = 1, last = 0;                                // empty range for the padding lanes
        if (j <= n) {
            first = j == 0 ? 0 : min((long long)(a[j - 1] >> shift), last_bucket) + 1;
            last = j == n ? last_entry : min((long long)(a[j] >> shift), last_bucket);
        }
        // a few entries (the usual case at about one row per bucket): the lane writes them itself
        const long long len = last - first + 1;
        if (len > 0 && len <= 4)
            for (long long b = first; b <= last; b++) tab[b] = (uint32_t)j;
        uint32_t todo = __ballot_sync(FULL_MASK, len > 4);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const long long f = __shfl_sync(FULL_MASK, first, src), l = __shfl_sync(FULL_MASK, last, src);
            const uint32_t v = (uint32_t)__shfl_sync(FULL_MASK, (long long)j, src);
            for (long long b = f + lane_id(); b <= l; b += 32) tab[b] = v;
        }
    }
}

// "Is there a row starting below x whose end lies above y" needs, after the directory, the rows of the bucket and the running
// maximum before the answer: three dependent loads at unrelated addresses for every probe row.  Almost every bucket holds at
// most two rows, so what the probe needs is stored beside the bucket's first position p = tab[b], one 32-byte sector per bucket:
//   [2 b]     = {p, gs[p], pmax[p - 1], pmax[p]}        [2 b + 1] = {gs[p + 1], pmax[p + 1], 0, 0}
// (gs = 0xffffffff and pmax = the last running maximum past the last row, pmax[-1] = 0).
__global__ void __launch_bounds__(256)
exists_dir_kernel(const uint32_t* __restrict__ tab, uint32_t entries, const uint32_t* __restrict__ gs, const uint32_t* __restrict__ pmax,
                  int64_t n, uint4* __restrict__ out) {
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < entries; b += gridDim.x * blockDim.x) {
        const uint32_t p = tab[b];
        uint4 e, f;
        e.x = p;
        e.y = p < n ? gs[p] : 0xffffffffu;
        e.z = p > 0 ? pmax[p - 1] : 0u;
        e.w = p < n ? pmax[p] : e.z;
        f.x = (int64_t)p + 1 < n ? gs[p + 1] : 0xffffffffu;
        f.y = (int64_t)p + 1 < n ? pmax[p + 1] : e.w;
        f.z = f.w = 0u;
        out[2 * (size_t)b] = e;
        out[2 * (size_t)b + 1] = f;
    }
}

struct DirSlots { int32_t slot[GIQLB_MAX_STRAND_CODES]; };
__global__ void dir_slots_kernel(DirSlots s, int32_t* __restrict__ out) {
    if (threadIdx.x < GIQLB_MAX_STRAND_CODES) out[threadIdx.x] = s.slot[threadIdx.x];
}

// block maxima of ge: one warp per 32-row block (bmax1), then one warp per 32 entries of bmax1 (bmax2)
__global__ void __launch_bounds__(256)
block_max_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out, int64_t n_out) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < n_out; b += n_warps) {
        const int64_t i = b * 32 + lane_id();
        uint32_t v = i < n ? in[i] : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(FULL_MASK, v, d));
        if (lane_id() == 0) out[b] = v;
    }
}

inline int grid_for(const giqlb_ctx* ctx, int64_t n, int threads, int per_sm) {
    int64_t want = ceil_div(n, threads);
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

// ---- layout ---------------------------------------------------------------------------------------------
extern "C" int giqlb_layout_create(giqlb_ctx* ctx, int32_t n_codes, int64_t pad, giqlb_layout** out) {
    if (!ctx || !out || n_codes < 0 || pad < 0) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "layout_create: bad argument");
    begin_call(ctx);
    giqlb_layout* lay = new (std::nothrow) giqlb_layout();
    if (!lay) return GIQLB_ERR_NOMEM;
    lay->n_codes = n_codes;
    lay->pad = pad;
    int st = dev_alloc(ctx, &lay->d_lohi, (size_t)2 * n_codes);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &lay->d_base, (size_t)n_codes + 1);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &lay->d_lo32, (size_t)n_codes);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &lay->d_hi32, (size_t)n_codes);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &lay->d_obs, (size_t)3 * GIQLB_MAX_OBSERVED);
    if (st == GIQLB_OK && cudaMemsetAsync(lay->d_obs, 0, 3 * GIQLB_MAX_OBSERVED * 8, ctx->stream) != cudaSuccess) st = GIQLB_ERR_CUDA;
    if (st != GIQLB_OK) {
        giqlb_layout_free(ctx, lay);
        return st;
    }
    if (n_codes > 0) {
        PhaseSpan ps(ctx, GIQLB_PH_LAYOUT);
        layout_init_kernel<<<(n_codes + 255) / 256, 256, 0, ctx->stream>>>((long long*)lay->d_lohi, n_codes);
        LAUNCH_CHECK(ctx);
    }
    *out = lay;
    return end_call(ctx);
}

extern "C" int giqlb_layout_observe(giqlb_ctx* ctx, giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                    const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj) {
    return giqlb_layout_observe_remap(ctx, lay, const_cast<int32_t*>(code), start, end, n, start_adj, end_adj, nullptr, nullptr);
}

// observe() that first translates the parked dictionary slots of `code` into codes, in place (slot2code == nullptr: plain observe)
int giqlb_layout_observe_remap(giqlb_ctx* ctx, giqlb_layout* lay, int32_t* code, const int32_t* start, const int32_t* end, int64_t n,
                               int32_t start_adj, int32_t end_adj, const int32_t* slot2code, const int* ok_flag) {
    if (!ctx || !lay || n < 0 || (n > 0 && (!code || !start || !end))) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "layout_observe: bad argument");
    if (lay->finalized) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "layout_observe: layout already finalized");
    begin_call(ctx);
    if (n > 0) {
        if (lay->n_codes == 0) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "layout_observe: rows given but the dictionary is empty");
        // the last slot collects every observe() beyond GIQLB_MAX_OBSERVED - 1 (those tables resolve lazily)
        const size_t slot = lay->observed.size() < GIQLB_MAX_OBSERVED - 1 ? lay->observed.size() : GIQLB_MAX_OBSERVED - 1;
        PhaseSpan ps(ctx, GIQLB_PH_LAYOUT);
        if (slot2code)
            layout_observe_kernel<true><<<grid_for(ctx, n, LY_THREADS * 8, 8), LY_THREADS, 0, ctx->stream>>>(
                code, start, end, n, lay->n_codes, start_adj, end_adj, (long long*)lay->d_lohi, lay->d_obs + 2 * slot,
                lay->d_obs + 2 * GIQLB_MAX_OBSERVED + slot, ctx->d_err, slot2code, ok_flag);
        else
            layout_observe_kernel<false><<<grid_for(ctx, n, LY_THREADS * 8, 8), LY_THREADS, 0, ctx->stream>>>(
                code, start, end, n, lay->n_codes, start_adj, end_adj, (long long*)lay->d_lohi, lay->d_obs + 2 * slot,
                lay->d_obs + 2 * GIQLB_MAX_OBSERVED + slot, ctx->d_err, nullptr, nullptr);
        LAUNCH_CHECK(ctx);
        if (lay->observed.size() < GIQLB_MAX_OBSERVED - 1) lay->observed.push_back({code, n, start_adj, end_adj, -1, -1});
    }
    return end_call(ctx);
}

// finalize without a host round trip: the bases are computed on the device; the host assumes the widest key
// (32 bits) and an overflowing extent or a bad row raises the sticky error flag that the next checked fetch reports.
// The observed tables' degenerate-row counts stay on the device (indexes resolve them lazily).
int giqlb_layout_finalize_async(giqlb_ctx* ctx, giqlb_layout* lay) {
    if (!ctx || !lay) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "layout_finalize: bad argument");
    if (lay->finalized) return GIQLB_OK;
    begin_call(ctx);
    {
        PhaseSpan ps(ctx, GIQLB_PH_LAYOUT);
        layout_finalize_kernel<<<1, 1024, 0, ctx->stream>>>((const long long*)lay->d_lohi, lay->n_codes, (long long)lay->pad, lay->d_base,
                                                            lay->d_lo32, lay->d_hi32, reinterpret_cast<unsigned long long*>(ctx->d_scalars + 2),
                                                            ctx->d_err);
        LAUNCH_CHECK(ctx);
    }
    lay->total_extent = 0xfffffffeull;   // upper bound
    lay->key_bits = 32;
    lay->finalized = true;
    return end_call(ctx);
}

extern "C" int giqlb_layout_finalize(giqlb_ctx* ctx, giqlb_layout* lay) {
    if (!ctx || !lay) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "layout_finalize: bad argument");
    if (lay->finalized) return GIQLB_OK;
    begin_call(ctx);
    unsigned long long* total = reinterpret_cast<unsigned long long*>(ctx->d_scalars);
    {
        PhaseSpan ps(ctx, GIQLB_PH_LAYOUT);
        layout_finalize_kernel<<<1, 1024, 0, ctx->stream>>>((const long long*)lay->d_lohi, lay->n_codes, (long long)lay->pad,
                                                            lay->d_base, lay->d_lo32, lay->d_hi32, total, ctx->d_err);
        LAUNCH_CHECK(ctx);
    }
    int h_err[4] = {0, 0, 0, 0};
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars, total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    lay->h_lo.assign(lay->n_codes, 0);
    lay->h_hi.assign(lay->n_codes, -1);
    lay->h_base.assign((size_t)lay->n_codes + 1, 0);
    std::vector<int32_t> lo32(lay->n_codes), hi32(lay->n_codes);
    if (lay->n_codes > 0) {
        CUDA_TRY(ctx, cudaMemcpyAsync(lo32.data(), lay->d_lo32, (size_t)lay->n_codes * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(hi32.data(), lay->d_hi32, (size_t)lay->n_codes * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(lay->h_base.data(), lay->d_base, ((size_t)lay->n_codes + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_err, ctx->d_err, sizeof(h_err), cudaMemcpyDeviceToHost, ctx->stream));
    unsigned long long h_obs[2 * GIQLB_MAX_OBSERVED];
    CUDA_TRY(ctx, cudaMemcpyAsync(h_obs, lay->d_obs, sizeof(h_obs), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < lay->observed.size(); k++) {
        lay->observed[k].n_degenerate = (int64_t)h_obs[2 * k];
        lay->observed[k].n_negative = (int64_t)h_obs[2 * k + 1];
    }
    for (int i = 0; i < lay->n_codes; i++) {
        lay->h_lo[i] = lo32[i];
        lay->h_hi[i] = hi32[i];
    }
    if (h_err[1]) {
        cudaMemsetAsync(ctx->d_err + 1, 0, sizeof(int), ctx->stream);
        return giqlb_fail(ctx, GIQLB_ERR_RANGE, "layout: a row has a chromosome code outside the dictionary or a coordinate outside [-2^30, 2^31-4]");
    }
    lay->total_extent = (uint64_t)ctx->h_scalars[0];
    if (lay->total_extent >= 0xffffffffull)
        return giqlb_fail(ctx, GIQLB_ERR_RANGE, "layout: summed chromosome extent %llu does not fit the 32-bit linear coordinate",
                          (unsigned long long)lay->total_extent);
    int bits = 1;
    while (bits < 32 && (lay->total_extent >> bits) != 0) bits++;
    lay->key_bits = bits;
    lay->finalized = true;
    return end_call(ctx);
}

extern "C" int giqlb_layout_free(giqlb_ctx* ctx, giqlb_layout* lay) {
    if (!lay) return GIQLB_OK;
    if (ctx) {
        dev_free(ctx, lay->d_lohi);
        dev_free(ctx, lay->d_base);
        dev_free(ctx, lay->d_lo32);
        dev_free(ctx, lay->d_hi32);
        dev_free(ctx, lay->d_obs);
    }
    delete lay;
    return GIQLB_OK;
}

extern "C" int giqlb_layout_info(const giqlb_layout* lay, int32_t* n_codes, uint64_t* total_extent, int32_t* key_bits) {
    if (!lay || !lay->finalized) return GIQLB_ERR_INVALID;
    if (n_codes) *n_codes = lay->n_codes;
    if (total_extent) *total_extent = lay->total_extent;
    if (key_bits) *key_bits = lay->key_bits;
    return GIQLB_OK;
}

extern "C" int giqlb_layout_tables(const giqlb_layout* lay, const uint32_t** base_dev, const int32_t** lo_dev,
                                   const int32_t** hi_dev) {
    if (!lay || !lay->finalized) return GIQLB_ERR_INVALID;
    if (base_dev) *base_dev = lay->d_base;
    if (lo_dev) *lo_dev = lay->d_lo32;
    if (hi_dev) *hi_dev = lay->d_hi32;
    return GIQLB_OK;
}

// ---- index ----------------------------------------------------------------------------------------------
extern "C" int giqlb_index_free(giqlb_ctx* ctx, giqlb_index* ix) {
    if (!ix) return GIQLB_OK;
    if (ctx) {
        dev_free(ctx, ix->gs);
        dev_free(ctx, ix->ge);
        dev_free(ctx, ix->rid);
        dev_free(ctx, ix->strand);
        dev_free(ctx, ix->pmax);
        dev_free(ctx, ix->ge_sorted);
        dev_free(ctx, ix->pos_by_end);
        dev_free(ctx, ix->d_block_off);
        dev_free(ctx, ix->d_counters);
        dev_free(ctx, ix->bmax1);
        dev_free(ctx, ix->bmax2);
        dev_free(ctx, ix->dir_gs);
        dev_free(ctx, ix->dir_exists);
        dev_free(ctx, ix->shard_G);
        dev_free(ctx, ix->dir_pmax);
        dev_free(ctx, ix->dir_ges);
        dev_free(ctx, ix->d_dir_slot);
    }
    delete ix;
    return GIQLB_OK;
}

extern "C" int64_t giqlb_index_rows(const giqlb_index* ix) { return ix ? ix->n : -1; }

extern "C" int giqlb_index_columns(const giqlb_index* ix, const uint32_t** gstart, const uint32_t** gend,
                                   const int32_t** row_id, const uint8_t** strand) {
    if (!ix) return GIQLB_ERR_INVALID;
    if (gstart) *gstart = ix->gs;
    if (gend) *gend = ix->ge;
    if (row_id) *row_id = ix->rid;
    if (strand) *strand = ix->strand;
    return GIQLB_OK;
}

extern "C" int giqlb_index_info(giqlb_ctx* ctx, giqlb_index* ix, int64_t* n_degenerate, int64_t* n_negative, int32_t* n_blocks,
                                int64_t* block_off_host /* [257] or NULL */) {
    if (!ctx || !ix) return GIQLB_ERR_INVALID;
    GIQLB_TRY(giqlb_index_resolve(ctx, ix, nullptr));
    if (n_degenerate) *n_degenerate = ix->n_degenerate;
    if (n_negative) *n_negative = ix->n_negative;
    if (n_blocks) *n_blocks = ix->n_blocks;
    if (block_off_host) memcpy(block_off_host, ix->block_off, sizeof(ix->block_off));
    return GIQLB_OK;
}

extern "C" int giqlb_index_build(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                 const int32_t* end, const uint8_t* strand, int64_t n, int32_t start_adj,
                                 int32_t end_adj, int flags, giqlb_index** out) {
    return giqlb_index_build_impl(ctx, lay, code, start, end, strand, nullptr, n, start_adj, end_adj, flags & 0xff, nullptr, out);
}

// `rows` (nullable) selects and orders the rows that are indexed; the index then carries rows[i] as row id.
extern "C" int giqlb_index_build_rows(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                      const int32_t* end, const uint8_t* strand, const int32_t* rows, int64_t n,
                                      int32_t start_adj, int32_t end_adj, int flags, giqlb_index** out) {
    return giqlb_index_build_impl(ctx, lay, code, start, end, strand, rows, n, start_adj, end_adj, flags & 0xff, nullptr, out);
}

// `pre` (nullable): the keys were already generated by a fused select + keygen pass (shard.cu); the builder takes
// ownership of its buffers.
int giqlb_index_build_impl(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                           const int32_t* end, const uint8_t* strand, const int32_t* rows, int64_t n, int32_t start_adj,
                           int32_t end_adj, int flags, giqlb_prekeyed* pre, giqlb_index** out) {
    if (!ctx || !lay || !out || n < 0 || (n > 0 && (!code || !start || !end)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "index_build: bad argument");
    if (!lay->finalized) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "index_build: layout not finalized");
    if (n >= (1ll << 30)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "index_build: n=%lld exceeds 2^30 rows per table", (long long)n);
    begin_call(ctx);
    *out = nullptr;
    giqlb_index* ix = new (std::nothrow) giqlb_index();
    if (!ix) return GIQLB_ERR_NOMEM;
    ix->n = n;
    ix->layout = lay;
    ix->tie_by_end = (flags & GIQLB_INDEX_TIE_BY_END) != 0;
    // one radix pass fewer when the caller can rebuild (see GIQLB_INDEX_SORT_SHORT_WINDOWS); pointless below two passes
    bool short_windows = (flags & GIQLB_INDEX_SORT_SHORT_WINDOWS) && !ix->tie_by_end && lay->key_bits > 16 && n > 0 && n < GIQLB_SORT_CARRY_END_MIN;
    if (short_windows && ctx->full_sort_credit > 0) {
        ctx->full_sort_credit--;
        short_windows = false;
    }
    ix->block_off[1] = n;
    for (int i = 2; i <= GIQLB_MAX_STRAND_CODES; i++) ix->block_off[i] = n;

    uint32_t *ka = nullptr, *kb = nullptr, *va = nullptr, *vb = nullptr, *ge_u = nullptr, *gs_u = nullptr;
    int32_t* pre_rows = nullptr;   // selection buffer handed over by a fused select + keygen pass
    bool ge_gathered = false;      // the last sort pass already wrote ix->ge
    bool rows_are_payload = false; // the sort carried the caller's row ids, not positions in the selection
    bool no_row_ids = false;       // GIQLB_INDEX_NO_ROW_IDS took effect: the index has no rid column
    unsigned long long* counters = nullptr;
    uint32_t* key_hist = nullptr;
    int st = GIQLB_OK;
#define IB_TRY(expr)                      \
    do {                                  \
        st = (expr);                      \
        if (st != GIQLB_OK) goto fail;    \
    } while (0)
#define IB_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            goto fail;                                                                                  \
        }                                                                                               \
    } while (0)

    if (pre) {
        ka = pre->gs; ge_u = pre->ge; rows = pre->rows; ix->d_counters = pre->counters;
        pre_rows = pre->rows;
        pre->gs = pre->ge = nullptr; pre->rows = nullptr; pre->counters = nullptr;
    } else {
        IB_TRY(dev_alloc(ctx, &ix->d_counters, 4 + 512));   // [0..3] counters, then uint32 [4][256] digit counts of the start keys
        IB_TRY(dev_alloc(ctx, &ka, (size_t)n));
        IB_TRY(dev_alloc(ctx, &ge_u, (size_t)n));
    }
    counters = ix->d_counters;
    // the key generation pass counts the sort's digit histograms as well (GIQLB_KEYGEN_HIST; not for pre-keyed or end-tied builds)
    key_hist = GIQLB_KEYGEN_HIST && !ix->tie_by_end && !pre ? reinterpret_cast<uint32_t*>(ix->d_counters + 4) : nullptr;
    IB_TRY(dev_alloc(ctx, &kb, (size_t)n));
    IB_TRY(dev_alloc(ctx, &va, (size_t)n));
    IB_TRY(dev_alloc(ctx, &vb, (size_t)n));
    IB_TRY(dev_alloc(ctx, &ix->ge, (size_t)n));
    if (!pre) IB_CUDA(cudaMemsetAsync(counters, 0, (4 + 512) * 8, ctx->stream));
    else if (short_windows) IB_CUDA(cudaMemsetAsync(counters + 3, 0, 8, ctx->stream));   // the selection size (known to the caller by now) makes room for the dense-window count
    ix->maybe_unfinished = short_windows;
    if (n > 0) {
        int result_in_b = 0;
        if (!pre) {
        PhaseSpan ps(ctx, GIQLB_PH_KEYGEN);
        if (key_hist)
            index_keygen_kernel<true><<<grid_for(ctx, n, 256 * 4, 8), 256, 0, ctx->stream>>>(
                code, start, end, rows, n, lay->n_codes, start_adj, end_adj, lay->d_base, lay->d_lo32, lay->d_hi32, ka, ge_u, counters, key_hist, ctx->d_err);
        else
            index_keygen_kernel<false><<<grid_for(ctx, n, 256 * 4, 8), 256, 0, ctx->stream>>>(
                code, start, end, rows, n, lay->n_codes, start_adj, end_adj, lay->d_base, lay->d_lo32, lay->d_hi32, ka, ge_u, counters, nullptr, ctx->d_err);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        }
        IB_CUDA(cudaGetLastError());
        if (ix->tie_by_end) {
            // LSD over the compound key (start, end): sort by end first, then stably by start
            IB_TRY(dev_alloc(ctx, &gs_u, (size_t)n));
            IB_CUDA(cudaMemcpyAsync(gs_u, ka, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            IB_CUDA(cudaMemcpyAsync(ka, ge_u, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            IB_TRY(giqlb_radix_sort_pairs_u32(ctx, ka, va, kb, vb, n, 0, lay->key_bits, 1, &result_in_b));
            uint32_t* rid1 = result_in_b ? vb : va;
            uint32_t* k2 = result_in_b ? ka : kb;   // the buffer pair NOT holding the result becomes the next input
            uint32_t* v2 = result_in_b ? va : vb;
            PhaseSpan ps(ctx, GIQLB_PH_GATHER);
            gather_kernel<uint32_t><<<grid_for(ctx, n, 256 * 4, 8), 256, 0, ctx->stream>>>(gs_u, rid1, k2, n);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            IB_CUDA(cudaGetLastError());
            IB_CUDA(cudaMemcpyAsync(v2, rid1, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            uint32_t* k3 = result_in_b ? kb : ka;
            uint32_t* v3 = result_in_b ? vb : va;
            int r2 = 0;
            IB_TRY(giqlb_radix_sort_pairs_u32(ctx, k2, v2, k3, v3, n, 0, lay->key_bits, 0, &r2));
            // result is in (k3, v3) when r2 else (k2, v2); normalise so that `result_in_b` describes (kb, vb)
            uint32_t* rk = r2 ? k3 : k2;
            result_in_b = (rk == kb) ? 1 : 0;
        } else {
            // (measured twice: letting the last radix pass gather the linear ends costs 72 us per 10 M-row pass and saves a 56 us gather
            //  kernel, so GIQLB_SORT_GATHERS_END stays off)
            // tables beyond the L2: the ends travel through the sort as a second payload (1.5 x the bytes of every pass) instead of
            // being gathered afterwards through the sorted row ids, one random DRAM sector each (200 M rows: +2.6 ms against 5.0 ms)
            const bool carry_end = n >= GIQLB_SORT_CARRY_END_MIN && !short_windows;
            const bool fused_gather = GIQLB_SORT_GATHERS_END && !short_windows && !carry_end;
            // with the ends carried nothing needs a row's position in the selection any more, so the sort can carry the caller's
            // row ids themselves instead of positions that are translated afterwards (a random gather: 2.2 ms at 100 M rows)
            rows_are_payload = carry_end && rows != nullptr;
            // GIQLB_INDEX_NO_ROW_IDS: the ends are the ONLY payload (no row ids: 16 instead of 24 bytes per row and pass)
            const bool ends_only = carry_end && (flags & GIQLB_INDEX_NO_ROW_IDS) && !strand;
            if (ends_only) {
                dev_free(ctx, va);                 // the value buffers of the sort are the end buffers
                dev_free(ctx, vb);
                va = vb = nullptr;
                IB_TRY(giqlb_radix_sort_pairs_u32(ctx, ka, ge_u, kb, ix->ge, n, 0, lay->key_bits, 0, &result_in_b, key_hist));
            } else {
                if (rows_are_payload) IB_CUDA(cudaMemcpyAsync(va, rows, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
                IB_TRY(giqlb_radix_sort_pairs_u32(ctx, ka, va, kb, vb, n, short_windows ? 8 : 0, lay->key_bits, rows_are_payload ? 0 : 1, &result_in_b, key_hist,
                                                  fused_gather ? ge_u : nullptr, fused_gather ? ix->ge : nullptr,
                                                  carry_end ? ge_u : nullptr, carry_end ? ix->ge : nullptr));
            }
            if (carry_end && !result_in_b) {   // the sorted ends lie beside the sorted keys: in the first buffer after an even number of passes
                uint32_t* t = ge_u;
                ge_u = ix->ge;
                ix->ge = t;
            }
            no_row_ids = ends_only;
            ge_gathered = fused_gather || short_windows || carry_end;
            if (short_windows) {
                // (k3, v3) -> the other buffer pair, in final order, with the ends
                uint32_t *k3 = result_in_b ? kb : ka, *v3 = result_in_b ? vb : va;
                uint32_t *k4 = result_in_b ? ka : kb, *v4 = result_in_b ? va : vb;
                PhaseSpan ps(ctx, GIQLB_PH_GATHER);
                sort_finish_kernel<<<(unsigned)ceil_div(n, SF_TILE), SF_THREADS, 0, ctx->stream>>>(k3, v3, n, ge_u, k4, v4, ix->ge, counters);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
                IB_CUDA(cudaGetLastError());
                result_in_b ^= 1;
            }
        }
        if (result_in_b) {
            ix->gs = kb; ix->rid = reinterpret_cast<int32_t*>(vb);
            kb = nullptr; vb = nullptr;
        } else {
            ix->gs = ka; ix->rid = reinterpret_cast<int32_t*>(va);
            ka = nullptr; va = nullptr;
        }
        if (no_row_ids) ix->rid = nullptr;   // (va / vb were released before the sort)
        PhaseSpan ps_g(ctx, GIQLB_PH_GATHER);
        if (!ge_gathered) {
            gather_kernel<uint32_t><<<grid_for(ctx, n, 256 * 4, 8), 256, 0, ctx->stream>>>(
                ge_u, reinterpret_cast<const uint32_t*>(ix->rid), ix->ge, n);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            IB_CUDA(cudaGetLastError());
        }
        if (rows && !rows_are_payload && !no_row_ids) {   // positions in the selection -> caller's row ids
            remap_rows_kernel<<<grid_for(ctx, n, 256 * 4, 8), 256, 0, ctx->stream>>>(ix->rid, rows, n);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            IB_CUDA(cudaGetLastError());
        }
        if (strand) {
            IB_TRY(dev_alloc(ctx, &ix->strand, (size_t)n));
            gather_kernel<uint8_t><<<grid_for(ctx, n, 256 * 4, 8), 256, 0, ctx->stream>>>(
                strand, reinterpret_cast<const uint32_t*>(ix->rid), ix->strand, n);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
            IB_CUDA(cudaGetLastError());
        }
    } else {
        ix->gs = ka; ix->rid = reinterpret_cast<int32_t*>(va);
        ka = nullptr; va = nullptr;
        if (strand) IB_TRY(dev_alloc(ctx, &ix->strand, 1));
    }
    dev_free(ctx, ka); dev_free(ctx, kb); dev_free(ctx, va); dev_free(ctx, vb); dev_free(ctx, ge_u); dev_free(ctx, gs_u);
    dev_free(ctx, pre_rows);
    ka = kb = va = vb = ge_u = gs_u = nullptr;
    pre_rows = nullptr;
    // degenerate-row counts: known already when the table was observed on this layout with the same offsets
    // (finalize() fetched them); otherwise they stay on the device until giqlb_index_resolve needs them.
    // Rows outside the layout raise the sticky flag checked at the next host round trip.
    ix->resolved = false;
    for (const auto& o : lay->observed) {
        if (o.code != code || o.sadj != start_adj || o.eadj != end_adj || o.n_degenerate < 0) continue;
        if (o.n_degenerate == 0 || (!rows && o.n == n)) {
            ix->n_degenerate = o.n_degenerate;
            ix->n_negative = o.n_negative;
            ix->resolved = true;
            break;
        }
    }
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = n;
    ctx->stats.algorithmic_bytes = 12 * n + 12 * n;
    *out = ix;
    return end_call(ctx);
fail:
    dev_free(ctx, ka); dev_free(ctx, kb); dev_free(ctx, va); dev_free(ctx, vb); dev_free(ctx, ge_u); dev_free(ctx, gs_u);
    dev_free(ctx, pre_rows);
    giqlb_index_free(ctx, ix);
    return st;
#undef IB_TRY
#undef IB_CUDA
}

// Reorders the index into blocks of equal strand code (ascending code), each block still sorted by start.
extern "C" int giqlb_index_group_by_strand(giqlb_ctx* ctx, giqlb_index* ix) {
    using namespace giqlb_scan;
    if (!ctx || !ix) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "group_by_strand: bad argument");
    if (!ix->strand) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "group_by_strand: index was built without a strand column");
    GIQLB_TRY(require_row_ids(ctx, ix, "group_by_strand"));
    if (ix->grouped_by_strand) return GIQLB_OK;
    begin_call(ctx);
    const int64_t n = ix->n;
    unsigned long long* hist = nullptr;
    GIQLB_TRY(dev_alloc(ctx, &hist, 256));
    CUDA_TRY(ctx, cudaMemsetAsync(hist, 0, 256 * 8, ctx->stream));
    if (n > 0) {
        strand_histogram_kernel<<<grid_for(ctx, n, 256 * 16, 4), 256, 0, ctx->stream>>>(ix->strand, n, hist);
        LAUNCH_CHECK(ctx);
    }
    std::vector<unsigned long long> h(256);
    CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), hist, 256 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, hist);
    ix->block_off[0] = 0;
    for (int c = 0; c < 256; c++) ix->block_off[c + 1] = ix->block_off[c] + (int64_t)h[c];
    ix->n_blocks = 256;
    if (n > 0) {
        uint32_t *gs_o = nullptr, *ge_o = nullptr;
        int32_t* rid_o = nullptr;
        uint8_t* st_o = nullptr;
        GIQLB_TRY(dev_alloc(ctx, &gs_o, (size_t)n));
        GIQLB_TRY(dev_alloc(ctx, &ge_o, (size_t)n));
        GIQLB_TRY(dev_alloc(ctx, &rid_o, (size_t)n));
        GIQLB_TRY(dev_alloc(ctx, &st_o, (size_t)n));
        for (int c = 0; c < 256; c++) {
            if (!h[c]) continue;
            GIQLB_TRY((launch_scan<OpAddU64>(ctx, n, InIsCode{ix->strand, (uint8_t)c},
                                             OutScatterRow{ix->gs, ix->ge, ix->rid, ix->strand, gs_o, ge_o, rid_o, st_o, ix->block_off[c]})));
        }
        dev_free(ctx, ix->gs); dev_free(ctx, ix->ge); dev_free(ctx, ix->rid); dev_free(ctx, ix->strand);
        ix->gs = gs_o; ix->ge = ge_o; ix->rid = rid_o; ix->strand = st_o;
        // derived arrays are stale now
        dev_free(ctx, ix->pmax); dev_free(ctx, ix->ge_sorted); dev_free(ctx, ix->pos_by_end);
        ix->pmax = nullptr; ix->ge_sorted = nullptr; ix->pos_by_end = nullptr;
        dev_free(ctx, ix->dir_gs); dev_free(ctx, ix->dir_pmax); dev_free(ctx, ix->dir_ges); dev_free(ctx, ix->d_dir_slot);
        ix->dir_gs = ix->dir_pmax = ix->dir_ges = nullptr; ix->d_dir_slot = nullptr;
        dev_free(ctx, ix->dir_exists);
        ix->dir_exists = nullptr;
        dev_free(ctx, ix->shard_G);
        ix->shard_G = nullptr;
        dev_free(ctx, ix->bmax1); dev_free(ctx, ix->bmax2);
        ix->bmax1 = ix->bmax2 = nullptr;
    }
    GIQLB_TRY(dev_alloc(ctx, &ix->d_block_off, (size_t)GIQLB_MAX_STRAND_CODES + 1));
    CUDA_TRY(ctx, cudaMemcpyAsync(ix->d_block_off, ix->block_off, sizeof(ix->block_off), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // block_off is pageable host memory inside *ix
    ix->grouped_by_strand = true;
    return end_call(ctx);
}

int giqlb_index_ensure_pmax(giqlb_ctx* ctx, giqlb_index* ix) {
    if (ix->pmax) return GIQLB_OK;
    GIQLB_TRY(dev_alloc(ctx, &ix->pmax, (size_t)ix->n));
    return giqlb_incl_scan_max_u32(ctx, ix->ge, ix->pmax, ix->n, ix->grouped_by_strand ? ix->block_off : nullptr,
                                   ix->grouped_by_strand ? ix->n_blocks : 1);
}

// ge ascending with the sorted position of each (pos_by_end), independently inside every strand block;
// positions are block-relative.  Stable LSD passes keep rows tied on end in index (start) order.
int giqlb_index_ensure_end_sorted(giqlb_ctx* ctx, giqlb_index* ix) {
    if (ix->ge_sorted) return GIQLB_OK;
    const int64_t n = ix->n;
    GIQLB_TRY(dev_alloc(ctx, &ix->ge_sorted, (size_t)n));
    int st = dev_alloc(ctx, &ix->pos_by_end, (size_t)n);
    uint32_t *ka = nullptr, *kb = nullptr, *va = nullptr, *vb = nullptr;
    if (st == GIQLB_OK) st = dev_alloc(ctx, &ka, (size_t)n);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &kb, (size_t)n);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &va, (size_t)n);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &vb, (size_t)n);
    const int nb = ix->grouped_by_strand ? ix->n_blocks : 1;
    for (int b = 0; b < nb && st == GIQLB_OK; b++) {
        const int64_t lo = ix->grouped_by_strand ? ix->block_off[b] : 0, hi = ix->grouped_by_strand ? ix->block_off[b + 1] : n;
        if (hi <= lo) continue;
        const size_t bytes = (size_t)(hi - lo) * 4;
        if (cudaMemcpyAsync(ka + lo, ix->ge + lo, bytes, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) {
            st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "end-sorted view: copy failed");
            break;
        }
        int in_b = 0;
        st = giqlb_radix_sort_pairs_u32(ctx, ka + lo, va + lo, kb + lo, vb + lo, hi - lo, 0, ix->layout->key_bits, 1, &in_b);
        if (st != GIQLB_OK) break;
        cudaMemcpyAsync(ix->ge_sorted + lo, (in_b ? kb : ka) + lo, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
        cudaMemcpyAsync(ix->pos_by_end + lo, (in_b ? vb : va) + lo, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    dev_free(ctx, ka); dev_free(ctx, kb); dev_free(ctx, va); dev_free(ctx, vb);
    if (st != GIQLB_OK) {
        dev_free(ctx, ix->ge_sorted); dev_free(ctx, ix->pos_by_end);
        ix->ge_sorted = nullptr; ix->pos_by_end = nullptr;
    }
    return st;
}

int giqlb_index_ensure_block_max(giqlb_ctx* ctx, giqlb_index* ix) {
    if (ix->bmax1 || ix->n == 0) return GIQLB_OK;
    const int64_t n1 = ceil_div(ix->n, 32), n2 = ceil_div(n1, 32);
    GIQLB_TRY(dev_alloc(ctx, &ix->bmax1, (size_t)n1 + 32));   // padded: bmax2's last warp reads a full 32 entries
    int st = dev_alloc(ctx, &ix->bmax2, (size_t)n2 + 1);
    if (st != GIQLB_OK) {
        dev_free(ctx, ix->bmax1);
        ix->bmax1 = nullptr;
        return st;
    }
    PhaseSpan ps(ctx, GIQLB_PH_SCAN_MAX);
    block_max_kernel<<<grid_for(ctx, n1 * 32, 256, 16), 256, 0, ctx->stream>>>(ix->ge, ix->n, ix->bmax1, n1);
    LAUNCH_CHECK(ctx);
    block_max_kernel<<<grid_for(ctx, n2 * 32, 256, 16), 256, 0, ctx->stream>>>(ix->bmax1, n1, ix->bmax2, n2);
    LAUNCH_CHECK(ctx);
    return GIQLB_OK;
}

// Bucket directories of the index's ascending arrays, one slot per strand block that holds rows.  The bucket width is
// chosen so that there are about as many buckets as rows (at most 2^24): a search through the directory reads two
// adjacent table entries and then a handful of rows, all in one or two cache lines.
int giqlb_index_ensure_dirs(giqlb_ctx* ctx, giqlb_index* ix, bool want_pmax, bool want_ges) {
    if ((want_pmax && !ix->pmax) || (want_ges && !ix->ge_sorted)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "index dirs: derived array missing");
    const int nb = ix->grouped_by_strand ? ix->n_blocks : 1;
    int slots = 0;
    int32_t h_slot[GIQLB_MAX_STRAND_CODES];
    for (int b = 0; b < GIQLB_MAX_STRAND_CODES; b++) {
        const bool has = b < nb && (ix->grouped_by_strand ? ix->block_off[b + 1] > ix->block_off[b] : ix->n > 0);
        h_slot[b] = has ? slots++ : -1;
    }
    if (!ix->d_dir_slot) {
        int bits = 1;
        while (bits < 24 && (1ll << bits) < ix->n) bits++;
        ix->dir_shift = 32 - bits;
        const uint64_t extent = ix->layout && ix->layout->total_extent ? ix->layout->total_extent : 0xffffffffull;
        ix->dir_stride = (uint32_t)((extent > 0xffffffffull ? 0xffffffffull : extent) >> ix->dir_shift) + 2u;
        GIQLB_TRY(dev_alloc(ctx, &ix->d_dir_slot, (size_t)GIQLB_MAX_STRAND_CODES));
        DirSlots ds;
        memcpy(ds.slot, h_slot, sizeof(h_slot));
        dir_slots_kernel<<<1, GIQLB_MAX_STRAND_CODES, 0, ctx->stream>>>(ds, ix->d_dir_slot);   // the table rides in the kernel arguments: no staging buffer, no wait
        LAUNCH_CHECK(ctx);
    }
    struct Want { bool on; uint32_t** tab; const uint32_t* arr; } wants[3] = {
        {true, &ix->dir_gs, ix->gs}, {want_pmax, &ix->dir_pmax, ix->pmax}, {want_ges, &ix->dir_ges, ix->ge_sorted}};
    for (auto& w : wants) {
        if (!w.on || *w.tab || slots == 0) continue;
        GIQLB_TRY(dev_alloc(ctx, w.tab, (size_t)slots * ix->dir_stride));
        PhaseSpan ps(ctx, GIQLB_PH_PROBE);
        for (int b = 0; b < nb; b++) {
            if (h_slot[b] < 0) continue;
            const int64_t lo = ix->grouped_by_strand ? ix->block_off[b] : 0, hi = ix->grouped_by_strand ? ix->block_off[b + 1] : ix->n;
            bucket_dir_kernel<<<grid_for(ctx, hi - lo + 1, 256, 16), 256, 0, ctx->stream>>>(w.arr + lo, hi - lo, ix->dir_shift, ix->dir_stride,
                                                                                              *w.tab + (size_t)h_slot[b] * ix->dir_stride);
            LAUNCH_CHECK(ctx);
        }
    }
    return GIQLB_OK;
}

// Bucket directory of any ascending device array (not only an index's own): tab[b] = first position with a value >= b << shift,
// about one bucket per element (at most 2^24), values beyond `extent` fall into the last bucket.  The caller frees *tab.
int giqlb_bucket_dir_build(giqlb_ctx* ctx, const uint32_t* arr, int64_t n, uint64_t extent, int* shift, uint32_t* stride, uint32_t** tab) {
    int bits = 1;
    while (bits < 24 && (1ll << bits) < n) bits++;
    *shift = 32 - bits;
    const uint64_t ext = extent ? extent : 0xffffffffull;
    *stride = (uint32_t)((ext > 0xffffffffull ? 0xffffffffull : ext) >> *shift) + 2u;
    GIQLB_TRY(dev_alloc(ctx, tab, (size_t)*stride));
    bucket_dir_kernel<<<grid_for(ctx, n + 1, 256, 16), 256, 0, ctx->stream>>>(arr, n, *shift, *stride, *tab);
    LAUNCH_CHECK(ctx);
    return GIQLB_OK;
}

// dir_exists of a plain index (requires pmax; builds dir_gs when missing)
int giqlb_index_ensure_exists_dir(giqlb_ctx* ctx, giqlb_index* ix) {
    if (ix->dir_exists || ix->n == 0) return GIQLB_OK;
    if (ix->grouped_by_strand || !ix->pmax) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "index exists-dir: plain index with pmax required");
    GIQLB_TRY(giqlb_index_ensure_dirs(ctx, ix, false, false));
    GIQLB_TRY(dev_alloc(ctx, &ix->dir_exists, (size_t)2 * ix->dir_stride));
    PhaseSpan ps(ctx, GIQLB_PH_PROBE);
    exists_dir_kernel<<<grid_for(ctx, ix->dir_stride, 256, 16), 256, 0, ctx->stream>>>(ix->dir_gs, ix->dir_stride, ix->gs, ix->pmax, ix->n, ix->dir_exists);
    LAUNCH_CHECK(ctx);
    return GIQLB_OK;
}

// Fetches the degenerate-row counts of up to two indexes in ONE round trip (no-op when already known).
int giqlb_index_resolve(giqlb_ctx* ctx, giqlb_index* a, giqlb_index* b) {
    giqlb_index* todo[2] = {a && (!a->resolved || a->maybe_unfinished) ? a : nullptr,
                            b && (!b->resolved || b->maybe_unfinished) && b != a ? b : nullptr};
    if (!todo[0] && !todo[1]) return GIQLB_OK;
    for (int k = 0; k < 2; k++)
        if (todo[k]) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars + 32 + 4 * k, todo[k]->d_counters, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 2; k++) {
        if (!todo[k]) continue;
        const int64_t* c = ctx->h_scalars + 32 + 4 * k;
        if (!todo[k]->resolved) {
            todo[k]->n_degenerate = c[0];
            todo[k]->n_negative = c[1];
            todo[k]->resolved = true;
        }
        if (todo[k]->maybe_unfinished) {
            todo[k]->maybe_unfinished = false;
            todo[k]->sort_unfinished = c[3] != 0;
        }
        if (todo[k]->sort_unfinished) {   // only the builder of the index can go on (pipeline.cu, shard.cu): internal status
            giqlb_fail(ctx, GIQLB_ERR_INTERNAL, "index: short-window sort left rows unordered and the caller did not rebuild");
            return GIQLB_RETRY_FULL_SORT;
        }
        if (c[2] != 0) {
            cudaMemsetAsync(ctx->d_err + 3, 0, sizeof(int), ctx->stream);
            return giqlb_fail(ctx, GIQLB_ERR_RANGE, "index: %lld rows lie outside the layout (observe the table before finalize)", (long long)c[2]);
        }
    }
    return GIQLB_OK;
}
