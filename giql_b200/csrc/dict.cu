// dict.cu — chromosome dictionary encoding on the device.
//
// The reference compares chromosomes as VARCHAR (`l.chrom = r.chrom`, abdenlab/giql
// src/giql/expanders/intersects.py:151) and the DuckDB plan re-scans both tables once per distinct name
// (src/giql/expanders/_per_chrom.py:27,46-74).  Here both tables are encoded ONCE against one shared
// dictionary whose codes preserve bytewise string order (MERGE's default ORDER BY chrom sorts VARCHAR,
// src/giql/expanders/merge.py:429-437), so every later kernel works on int32 codes.
//
// Pass 1 (dict_insert_kernel): a CTA takes 1024 consecutive rows, stages their offsets and their (contiguous)
// name bytes in shared memory with coalesced loads, and resolves every row to a slot of a global
// open-addressing table whose slots hold (representative row + 1).  Names of up to 16 bytes — all real
// chromosome names — are compared as two 64-bit words against a small per-CTA cache of names already
// resolved (256 16-byte entries; the home entry and a second, independent position are loaded and tested
// together, so a lookup is two shared-memory loads whatever the other names hash to), so after the first
// few rows a lookup touches shared memory only; a cache miss (or a longer name) probes the global table: one CAS claims an empty slot, otherwise the bytes are compared with the
// representative's — exact, no reliance on hash uniqueness.  The row's slot is parked in the output array.
// Pass 2 gathers the distinct names into a small buffer; the host sorts them (one round trip for the whole
// call) and uploads a slot -> code table.  Pass 3 rewrites the parked slots into codes.
// Traffic: offsets + bytes once, 8 B/row of codes.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int DI_THREADS = 256;
constexpr int DI_ROWS = 1024;          // rows per CTA chunk
#ifndef GIQLB_DICT_V4
#define GIQLB_DICT_V4 1                // 16-byte cache entries probed two at a time, hit test first (0: the round-1 cache; A/B switch)
#endif
#ifndef GIQLB_DICT_TWO_CHOICE
#define GIQLB_DICT_TWO_CHOICE 1        // V4: the second probed entry is an independent hash position instead of the neighbour
#endif
#ifndef GIQLB_DICT_MAX_PROBES
#define GIQLB_DICT_MAX_PROBES 32       // cache entries examined before a lookup goes to the global table (a FULL cache — more
                                       // names than entries — would otherwise be scanned end to end by every miss)
#endif
#ifndef GIQLB_DICT_WIDE_SLOTS
#define GIQLB_DICT_WIDE_SLOTS 0        // 1: global table slots carry the (<= 16-byte) name beside the representative row (measured:
                                       // 3 000 names 8.5 -> 7.0 ms, but 24 names 0.130 -> 0.140 ms: three dependent L2 round trips
                                       // where the byte comparison through the representative row runs out of L1)
#endif
#ifndef GIQLB_DICT_LOAD_CAP
#define GIQLB_DICT_LOAD_CAP 0          // 1: the per-CTA cache stops inserting at half full
#endif
#ifndef GIQLB_DICT_STAGE_BYTES
#define GIQLB_DICT_STAGE_BYTES 14336   // + the 6 KB cache = 43.1 KB: five CTAs per SM (names average <= 14 bytes per row, else the
                                       // chunk is staged synchronously)
#endif
constexpr int DI_BYTES = GIQLB_DICT_STAGE_BYTES;   // staged name bytes per chunk (longer chunks read global memory directly)
// Measured, 2 x 10 M rows (scripts/dict_probe.py, bench.py's dict phase): the round-1 cache (128 entries of separate key /
// packed-word arrays, one probe loop) 0.161 ms for the 24 UCSC names, and its time moved with how those 24 names happened to hash
// (8 / 9 bits 0.19 ms, 10 bits 0.152 ms); 194 names (alt contigs) overflowed it and took 11.7 ms, 3 000 scaffolds 11.6 ms.  Now:
// 256 16-byte entries, the home entry and a second independent position tested together before any loop, at most 32 further
// probes: 0.130 ms, 1.5 ms and 8.5 ms; the same with 128 entries 0.22 ms (the 24 names collide), with 512 entries and 8 KB of
// staged bytes 0.128 ms but RefSeq-style 12-byte names then miss the staging buffer (0.65 ms against 0.30).
#ifndef GIQLB_DICT_CACHE_BITS
#define GIQLB_DICT_CACHE_BITS (GIQLB_DICT_V4 ? 8 : 7)
#endif
constexpr int DI_CACHE = 1 << GIQLB_DICT_CACHE_BITS;   // per-CTA cache entries
constexpr int DC_MAX_ENTRIES = 4096;   // distinct names gathered in the first round trip
constexpr int DC_NAME_CAP = 65536;     // their bytes

struct StrTable {
    const int32_t* offs;
    const uint8_t* bytes;
    int64_t n;
};

__device__ __forceinline__ uint64_t mix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ull;
    return h ^ (h >> 33);
}
// hash of a name: names of <= 16 bytes hash their two zero-padded words, longer names all their bytes
__device__ __forceinline__ uint64_t hash_short(uint64_t w0, uint64_t w1, int len) {
    return (w0 ^ (w1 * 0x9E3779B97F4A7C15ull) ^ (uint64_t)len) * 0xff51afd7ed558ccdull;   // high bits are the well-mixed ones
}
__device__ __forceinline__ uint64_t hash_long(const uint8_t* __restrict__ p, int len) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < len; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return mix64(h);
}
__device__ __forceinline__ bool same_bytes(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int len) {
    for (int i = 0; i < len; i++)
        if (a[i] != b[i]) return false;
    return true;
}

// Global table: returns the slot of the name (s, len), inserting representative row g when it is new.  A slot is DS_WORDS
// 64-bit words — one 32-byte sector: [0] representative row + 1 (0 = free; claimed with one CAS), [1] [2] the zero-padded
// words of a name of <= 16 bytes, [3] its length + 1, published LAST with release semantics (0 = not published, or a longer
// name).  A short name is therefore compared against a published entry in registers, from the one sector the probe loads;
// only unpublished entries and names longer than 16 bytes are compared byte by byte through the representative row.  (Round
// 2, scripts/dict_probe.py: with 194 contig names — more than the per-CTA cache held — every row took the byte-by-byte
// route, ~20 dependent global loads, and 2 x 10 M rows encoded in 11.7 ms against 0.29 ms for 24 names.)
constexpr int DS_WORDS = GIQLB_DICT_WIDE_SLOTS ? 4 : 1;

__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t global_resolve(const StrTable& A, const StrTable& B, unsigned long long* __restrict__ slots,
                                                   uint32_t cap_mask, const uint8_t* __restrict__ s, int len, uint64_t h, int64_t g,
                                                   unsigned int* __restrict__ n_distinct, uint64_t w0, uint64_t w1) {
    uint32_t slot = (uint32_t)(h >> 24) & cap_mask;
    const bool is_short = GIQLB_DICT_WIDE_SLOTS && len <= 16;   // the caller passes (w0, w1) for these
    for (uint32_t probes = 0; probes <= cap_mask; probes++) {
        unsigned long long* e = slots + (size_t)slot * DS_WORDS;
        unsigned long long w = *(volatile unsigned long long*)&e[0];
        if (w == 0) {
            unsigned long long prev = atomicCAS(&e[0], 0ull, (unsigned long long)g + 1ull);
            if (prev == 0) {
                if (is_short) {
                    *(volatile unsigned long long*)&e[1] = w0;
                    *(volatile unsigned long long*)&e[2] = w1;
                    st_release_u64(&e[3], (unsigned long long)len + 1ull);
                }
                atomicAdd(n_distinct, 1u);
                return slot;
            }
            w = prev;
        }
        if (is_short) {
            // The words are loaded only after the tag has been seen non-zero (a control dependency: the loads below are issued
            // after this one has returned; all of them are served by the L2, the coherence point), and the writer released the
            // tag after the words.  An acquiring load here put a fence on every lookup: +9 % on the 24-name encode.
            const unsigned long long tag = *(volatile unsigned long long*)&e[3];
            if (tag != 0) {                          // a published short name: exact comparison in registers
                if (tag == (unsigned long long)len + 1ull && *(volatile unsigned long long*)&e[1] == w0 &&
                    *(volatile unsigned long long*)&e[2] == w1)
                    return slot;
                slot = (slot + 1) & cap_mask;
                continue;
            }
        }
        const int64_t rg = (int64_t)w - 1;
        const StrTable& RT = rg >= A.n ? B : A;
        const int64_t ri = rg >= A.n ? rg - A.n : rg;
        const int32_t r0 = RT.offs[ri], rlen = RT.offs[ri + 1] - r0;
        if (rlen == len && same_bytes(s, RT.bytes + r0, len)) return slot;
        slot = (slot + 1) & cap_mask;
    }
    return cap_mask;   // table full: cannot happen, the host keeps the load factor <= 1/2 and re-runs otherwise
}

// A name read straight from global memory (chunks that were not staged, names longer than 16 bytes).
__device__ __forceinline__ uint32_t resolve_unstaged(const StrTable& A, const StrTable& B, unsigned long long* __restrict__ slots,
                                                     uint32_t cap_mask, const uint8_t* __restrict__ p, int len, int64_t g,
                                                     unsigned int* __restrict__ n_distinct) {
    if (len > 16) return global_resolve(A, B, slots, cap_mask, p, len, hash_long(p, len), g, n_distinct, 0, 0);
    uint64_t w[2] = {0, 0};
    for (int i = 0; i < len; i++) w[i >> 3] |= (uint64_t)p[i] << (8 * (i & 7));
    return global_resolve(A, B, slots, cap_mask, p, len, hash_short(w[0], w[1], len), g, n_distinct, w[0], w[1]);
}

constexpr uint32_t DI_EMPTY = 0xFFFFFFFFu, DI_PENDING = 0xFFFFFFFEu;
constexpr uint32_t DI_SLOT_BITS = 27, DI_SLOT_MASK = (1u << DI_SLOT_BITS) - 1u;   // cache word = slot | len << 27 (len <= 16)

#if GIQLB_DICT_V4
// Per-CTA cache of resolved names, open addressing over DI_CACHE entries.  An entry is ONE 16-byte word — the name's first
// eight bytes and the packed (global slot | len << 27) — so a probe is one shared-memory load, and a lookup loads the home
// entry and a second, independently hashed entry together and tests them for a hit before anything else: the 24 names of a
// genome spread over the table with the odd collision, and a warp pays for the longest probe chain among its lanes.  Names
// of 9..16 bytes keep their second word in k1.  (ncu, round 2: the probe loop was 37 % of the kernel's 162 instructions per
// row; with this cache the kernel executes 75.7 M warp instructions instead of 98.5 M, profiles/r02s_kernels.csv.)
struct __align__(16) DictEnt {
    unsigned long long k0;
    uint32_t meta;             // DI_EMPTY / DI_PENDING / (global slot | len << 27)
    uint32_t pad;
};
struct DictCache {
    DictEnt e[DI_CACHE];
    unsigned long long k1[DI_CACHE];
    uint32_t count;            // entries filled; with GIQLB_DICT_LOAD_CAP inserts stop at half the table (measured: no gain, the
                               // bound on the probe sequence below is what keeps a full cache cheap)
};

__device__ __forceinline__ uint4 lds_volatile_v4(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p))
                 : "memory");
    return v;
}

__device__ __forceinline__ void dict_cache_init(DictCache& c) {
    // k0 = all ones: an entry whose packed word has been published always carries its own key (the key is written first,
    // then a block-level fence, then the packed word), and an unpublished one can never pass the length test
    for (int i = threadIdx.x; i < DI_CACHE; i += DI_THREADS) {
        c.e[i].k0 = ~0ull;
        c.e[i].meta = DI_EMPTY;
        c.e[i].pad = 0;
    }
    if (threadIdx.x == 0) c.count = 0;
}

__device__ __forceinline__ uint32_t cached_resolve(DictCache& c, const StrTable& A, const StrTable& B, unsigned long long* __restrict__ slots,
                                                   uint32_t cap_mask, const uint8_t* __restrict__ name, uint64_t w0, uint64_t w1, int len,
                                                   int64_t g, unsigned int* __restrict__ n_distinct) {
    const uint32_t lo = (uint32_t)w0, hi = (uint32_t)(w0 >> 32);
    const uint32_t h = (lo ^ (hi * 0x9E3779B1u) ^ ((uint32_t)w1 * 0x85EBCA6Bu) ^ (uint32_t)(w1 >> 32) ^ (uint32_t)len) * 0xC2B2AE35u;
    int ci = (int)(h >> (32 - GIQLB_DICT_CACHE_BITS));
#if GIQLB_DICT_TWO_CHOICE
    // second, independent position: a name whose home entry is taken goes THERE (and on linearly from there), so two names
    // that share a home entry do not also push each other along one chain
    int cj = (int)((h * 0x9E3779B1u) >> (32 - GIQLB_DICT_CACHE_BITS));
    cj = cj == ci ? (ci + 1) & (DI_CACHE - 1) : cj;
#else
    const int cj = (ci + 1) & (DI_CACHE - 1);
#endif
    const uint32_t want_len = (uint32_t)len << DI_SLOT_BITS;
    {
        const uint4 a = lds_volatile_v4(&c.e[ci]), b = lds_volatile_v4(&c.e[cj]);
        // EMPTY and PENDING carry 31 in the length bits: they never match a length <= 16
        const bool hit_a = a.x == lo && a.y == hi && (a.z & ~DI_SLOT_MASK) == want_len;
        const bool hit_b = b.x == lo && b.y == hi && (b.z & ~DI_SLOT_MASK) == want_len;
        if (len <= 8) {
            if (hit_a) return a.z & DI_SLOT_MASK;
            if (hit_b) return b.z & DI_SLOT_MASK;       // an entry that carries this name and length IS this name's entry
        } else {
            if (hit_a && *(volatile unsigned long long*)&c.k1[ci] == w1) return a.z & DI_SLOT_MASK;
            if (hit_b && *(volatile unsigned long long*)&c.k1[cj] == w1) return b.z & DI_SLOT_MASK;
        }
    }
    int free_at = -1;
#pragma unroll 1
    for (int p = 0; p < (DI_CACHE < GIQLB_DICT_MAX_PROBES ? DI_CACHE : GIQLB_DICT_MAX_PROBES); p++) {
        const uint4 a = lds_volatile_v4(&c.e[ci]);
        if (a.z == DI_EMPTY) {
            free_at = ci;
            break;
        }
        if (a.z == DI_PENDING) break;                      // another thread is filling this entry (almost surely with this very name):
                                                           // resolve globally and do NOT insert, or the cache fills up with duplicates
        if ((a.z & ~DI_SLOT_MASK) == want_len && a.x == lo && a.y == hi && (len <= 8 || *(volatile unsigned long long*)&c.k1[ci] == w1))
            return a.z & DI_SLOT_MASK;
        ci = p == 0 ? cj : (ci + 1) & (DI_CACHE - 1);      // probe sequence: home, second position, then linearly from there
    }
    const uint32_t slot = global_resolve(A, B, slots, cap_mask, name, len, hash_short(w0, w1, len), g, n_distinct, w0, w1);
    if (free_at >= 0 && slot <= DI_SLOT_MASK && (!GIQLB_DICT_LOAD_CAP || *(volatile uint32_t*)&c.count < (uint32_t)(DI_CACHE / 2)) &&
        atomicCAS(&c.e[free_at].meta, DI_EMPTY, DI_PENDING) == DI_EMPTY) {
        atomicAdd(&c.count, 1u);
        *(volatile unsigned long long*)&c.e[free_at].k0 = w0;
        *(volatile unsigned long long*)&c.k1[free_at] = w1;
        __threadfence_block();
        *(volatile uint32_t*)&c.e[free_at].meta = slot | want_len;
    }
    return slot;
}
#else
// Per-CTA cache of resolved names (open addressing over DI_CACHE entries): key words + one packed word.
struct DictCache {
    unsigned long long k0[DI_CACHE], k1[DI_CACHE];
    uint32_t meta[DI_CACHE];   // DI_EMPTY / DI_PENDING / (global slot | len << 27)
};

// Resolves the name (w0, w1, len <= 16; zero padded) of global row g to its slot of the global table, through the cache.
__device__ __forceinline__ uint32_t cached_resolve(DictCache& c, const StrTable& A, const StrTable& B, unsigned long long* __restrict__ slots,
                                                   uint32_t cap_mask, const uint8_t* __restrict__ name, uint64_t w0, uint64_t w1, int len,
                                                   int64_t g, unsigned int* __restrict__ n_distinct) {
    // 32-bit mix for the cache index (the 64-bit hash is only needed on a miss)
    const uint32_t lo = (uint32_t)w0, hi = (uint32_t)(w0 >> 32);
    uint32_t h = (lo ^ (hi * 0x9E3779B1u) ^ ((uint32_t)w1 * 0x85EBCA6Bu) ^ (uint32_t)(w1 >> 32) ^ (uint32_t)len) * 0xC2B2AE35u;
    int ci = (int)(h >> (32 - GIQLB_DICT_CACHE_BITS));
    const uint32_t want_len = (uint32_t)len << DI_SLOT_BITS;
    int free_at = -1;
#pragma unroll 1
    for (int p = 0; p < (DI_CACHE < GIQLB_DICT_MAX_PROBES ? DI_CACHE : GIQLB_DICT_MAX_PROBES); p++) {
        const uint32_t v = *(volatile uint32_t*)&c.meta[ci];
        const unsigned long long k = c.k0[ci];             // issued together with the meta load
        if (v == DI_EMPTY) {
            free_at = ci;
            break;
        }
        if (v == DI_PENDING) break;                        // another thread is filling this entry (almost surely with this very name):
                                                           // resolve globally and do NOT insert, or the cache fills up with duplicates
        if ((v & ~DI_SLOT_MASK) == want_len && k == w0 && (len <= 8 || c.k1[ci] == w1)) return v & DI_SLOT_MASK;   // PENDING never matches: its top bits are all ones
        ci = (ci + 1) & (DI_CACHE - 1);
    }
    const uint32_t slot = global_resolve(A, B, slots, cap_mask, name, len, hash_short(w0, w1, len), g, n_distinct, w0, w1);
    if (free_at >= 0 && slot <= DI_SLOT_MASK && atomicCAS(&c.meta[free_at], DI_EMPTY, DI_PENDING) == DI_EMPTY) {
        c.k0[free_at] = w0;
        c.k1[free_at] = w1;
        __threadfence_block();
        *(volatile uint32_t*)&c.meta[free_at] = slot | want_len;
    }
    return slot;
}

__device__ __forceinline__ void dict_cache_init(DictCache& c) {
    for (int i = threadIdx.x; i < DI_CACHE; i += DI_THREADS) c.meta[i] = DI_EMPTY;
}
#endif

// name of <= 16 bytes at byte offset so of the staged buffer, as two zero-padded little-endian words
__device__ __forceinline__ void staged_name(const uint8_t* s_bytes, int so, int len, uint64_t& w0, uint64_t& w1) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_bytes) + (so >> 2);
    const int sh = (so & 3) * 8;
    const uint32_t t0 = wp[0], t1 = wp[1], t2 = wp[2];
    w0 = (uint64_t)__funnelshift_r(t0, t1, sh) | ((uint64_t)__funnelshift_r(t1, t2, sh) << 32);
    w1 = 0;
    if (len > 8) {
        const uint32_t t3 = wp[3], t4 = wp[4];
        w1 = (uint64_t)__funnelshift_r(t2, t3, sh) | ((uint64_t)__funnelshift_r(t3, t4, sh) << 32);
        if (len < 16) w1 &= ~0ull >> (128 - 8 * len);
    } else {
#if GIQLB_DICT_V4
        // zero padding in 32-bit halves (a 64-bit variable shift is several instructions): nb bits of the name are kept
        const uint32_t nb = 8u * (uint32_t)len;                                      // 0 .. 64
        const uint32_t mlo = nb >= 32u ? 0xFFFFFFFFu : ((1u << nb) - 1u);
        const uint32_t mhi = nb >= 64u ? 0xFFFFFFFFu : (nb <= 32u ? 0u : ((1u << (nb - 32u)) - 1u));
        w0 = (uint64_t)((uint32_t)w0 & mlo) | ((uint64_t)((uint32_t)(w0 >> 32) & mhi) << 32);
#else
        w0 = len ? w0 & (~0ull >> (64 - 8 * len)) : 0ull;
#endif
    }
}

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 1024-row chunk of table A or B, and what is needed to stage it before its offsets are in shared memory:
// the byte range comes from two scalar loads of the offsets column.
struct DictChunk {
    bool valid, in_b, async;       // async: full chunk, aligned columns, bytes fit -> staged with cp.async one iteration ahead
    int rows;
    int64_t r0;
    int32_t b0, b1;
};

__device__ __forceinline__ DictChunk dict_chunk(const StrTable& A, const StrTable& B, int chunk, int chunks_a, int total) {
    DictChunk c;
    c.valid = chunk < total;
    c.in_b = chunk >= chunks_a;
    c.r0 = 0; c.rows = 0; c.b0 = c.b1 = 0; c.async = false;
    if (!c.valid) return c;
    const StrTable& T = c.in_b ? B : A;
    c.r0 = (int64_t)(c.in_b ? chunk - chunks_a : chunk) * DI_ROWS;
    c.rows = (int)min((int64_t)DI_ROWS, T.n - c.r0);
    c.b0 = __ldg(T.offs + c.r0);
    c.b1 = __ldg(T.offs + c.r0 + c.rows);
    c.async = c.rows == DI_ROWS && ((reinterpret_cast<uintptr_t>(T.offs) | reinterpret_cast<uintptr_t>(T.bytes)) & 15) == 0 &&
              c.b1 - (c.b0 & ~15) <= DI_BYTES;
    return c;
}

struct DictStage {
    __align__(16) int32_t offs[DI_ROWS + 4];
    __align__(16) uint8_t bytes[DI_BYTES + 32];
};

__device__ __forceinline__ void dict_stage_async(const StrTable& A, const StrTable& B, const DictChunk& c, DictStage& st) {
    if (!c.valid || !c.async) return;
    const StrTable& T = c.in_b ? B : A;
    cp_async_16(reinterpret_cast<int4*>(st.offs) + threadIdx.x, reinterpret_cast<const int4*>(T.offs + c.r0) + threadIdx.x);
    if (threadIdx.x == 0) cp_async_4(&st.offs[DI_ROWS], T.offs + c.r0 + DI_ROWS);
    const int32_t a0 = c.b0 & ~15;
    const uint4* src = reinterpret_cast<const uint4*>(T.bytes + a0);
    const int n_vec = (c.b1 - a0 + 15) >> 4;               // the tail vector stays inside the (padded) allocation
    for (int i = threadIdx.x; i < n_vec; i += DI_THREADS) cp_async_16(reinterpret_cast<uint4*>(st.bytes) + i, src + i);
}

__global__ void __launch_bounds__(DI_THREADS)
dict_insert_kernel(StrTable A, StrTable B, unsigned long long* __restrict__ slots, uint32_t cap_mask,
                   int32_t* __restrict__ slot_a, int32_t* __restrict__ slot_b, unsigned int* __restrict__ n_distinct) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    DictStage* stage = reinterpret_cast<DictStage*>(s_dyn);                 // two stages: one being read, one in flight
    DictCache& cache = *reinterpret_cast<DictCache*>(s_dyn + 2 * sizeof(DictStage));

    dict_cache_init(cache);
    const int chunks_a = (int)((A.n + DI_ROWS - 1) / DI_ROWS), chunks_b = (int)((B.n + DI_ROWS - 1) / DI_ROWS);
    const int total = chunks_a + chunks_b, stride = gridDim.x;
    int chunk = blockIdx.x;
    DictChunk cur = dict_chunk(A, B, chunk, chunks_a, total);
    dict_stage_async(A, B, cur, stage[0]);
    cp_async_commit();
    DictChunk nxt = dict_chunk(A, B, chunk + stride, chunks_a, total);
    for (int it = 0; cur.valid; it++, chunk += stride) {
        DictStage& st = stage[it & 1];
        dict_stage_async(A, B, nxt, stage[(it & 1) ^ 1]);       // that stage was released by the barrier ending the previous iteration
        cp_async_commit();
        const DictChunk after = dict_chunk(A, B, chunk + 2 * stride, chunks_a, total);   // its two scalar loads overlap this chunk's work
        cp_async_wait<1>();                                      // everything but the group just committed has landed
        const StrTable& T = cur.in_b ? B : A;
        const int rows = cur.rows;
        const int64_t r0 = cur.r0, g0 = (cur.in_b ? A.n : 0) + r0;
        const int32_t a0 = cur.b0 & ~15;
        bool staged = cur.async;
        if (!cur.async) {                                        // tail chunk / unaligned columns / long names: staged synchronously
            for (int i = threadIdx.x; i <= rows; i += DI_THREADS) st.offs[i] = T.offs[r0 + i];
            staged = (cur.b1 - a0) <= DI_BYTES && (reinterpret_cast<uintptr_t>(T.bytes) & 15) == 0;
            if (staged) {
                const uint4* src = reinterpret_cast<const uint4*>(T.bytes + a0);
                const int n_vec = (cur.b1 - a0 + 15) >> 4;
                for (int i = threadIdx.x; i < n_vec; i += DI_THREADS) reinterpret_cast<uint4*>(st.bytes)[i] = ld_stream_u32x4(src + i);
            }
        }
        __syncthreads();
        int32_t* out = (cur.in_b ? slot_b : slot_a) + r0;
        if (cur.async && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            // the common case: a thread owns four consecutive rows — one 16-byte load of their offsets, one 16-byte store
            // of their slots
            const int k = threadIdx.x * 4;
            const int4 o4 = reinterpret_cast<const int4*>(st.offs)[threadIdx.x];
            const int32_t o[5] = {o4.x, o4.y, o4.z, o4.w, st.offs[k + 4]};
            int32_t res[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int len = o[q + 1] - o[q];
                if (len <= 16) {
                    uint64_t w0, w1;
                    staged_name(st.bytes, o[q] - a0, len, w0, w1);
                    res[q] = (int32_t)cached_resolve(cache, A, B, slots, cap_mask, T.bytes + o[q], w0, w1, len, g0 + k + q, n_distinct);
                } else {
                    const uint8_t* sp = T.bytes + o[q];
                    res[q] = (int32_t)resolve_unstaged(A, B, slots, cap_mask, sp, len, g0 + k + q, n_distinct);
                }
            }
            reinterpret_cast<int4*>(out)[threadIdx.x] = make_int4(res[0], res[1], res[2], res[3]);
        } else {
            for (int k = threadIdx.x; k < rows; k += DI_THREADS) {
                const int32_t o = st.offs[k], len = st.offs[k + 1] - o;
                uint32_t slot;
                if (staged && len <= 16) {
                    uint64_t w0, w1;
                    staged_name(st.bytes, o - a0, len, w0, w1);
                    slot = cached_resolve(cache, A, B, slots, cap_mask, T.bytes + o, w0, w1, len, g0 + k, n_distinct);
                } else {
                    const uint8_t* sp = T.bytes + o;
                    slot = resolve_unstaged(A, B, slots, cap_mask, sp, len, g0 + k, n_distinct);
                }
                out[k] = (int32_t)slot;
            }
        }
        __syncthreads();   // this stage may be refilled from the next iteration on
        cur = nxt;
        nxt = after;
    }
    cp_async_wait<0>();
}

constexpr size_t DI_SMEM = 2 * sizeof(DictStage) + sizeof(DictCache);

// above 48 KB the dynamic shared-memory size is a per-device opt-in (remembered per context)
inline void dict_smem_opt_in(giqlb_ctx* ctx) {
    if (DI_SMEM <= 48 * 1024 || ctx->dict_smem_opted_in) return;
    cudaFuncSetAttribute(dict_insert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DI_SMEM);
    ctx->dict_smem_opted_in = true;
}

// one thread per slot: copy the representative's bytes into the name buffer
__global__ void dict_collect_kernel(StrTable A, StrTable B, const unsigned long long* __restrict__ slots, uint32_t cap,
                                    uint8_t* __restrict__ name_buf, unsigned long long name_cap,
                                    unsigned long long* __restrict__ cursor /* [0] bytes, [1] entries */,
                                    int32_t* __restrict__ ent /* [3][max_entries]: slot, offset, length */, uint32_t max_entries) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    unsigned long long w = slots[(size_t)slot * DS_WORDS];
    if (w == 0) return;
    const int64_t rg = (int64_t)w - 1;
    const StrTable& RT = rg >= A.n ? B : A;
    const int64_t ri = rg >= A.n ? rg - A.n : rg;
    const int32_t r0 = RT.offs[ri], len = RT.offs[ri + 1] - r0;
    unsigned long long off = atomicAdd(&cursor[0], (unsigned long long)len);
    unsigned long long e = atomicAdd(&cursor[1], 1ull);
    if (e < max_entries) {
        ent[e] = (int32_t)slot;
        ent[max_entries + e] = (int32_t)off;
        ent[2 * max_entries + e] = len;
    }
    if (off + len <= name_cap)
        for (int i = 0; i < len; i++) name_buf[off + i] = RT.bytes[r0 + i];
}

__global__ void __launch_bounds__(256)
dict_remap_kernel(int32_t* __restrict__ codes, int64_t n, const int32_t* __restrict__ slot2code, const int* __restrict__ ok_flag) {
    if (ok_flag && !*ok_flag) return;   // the device-side ranking declined: the host path will redo the encode
    const int64_t n_vec = n >> 2;
    int4* cv = reinterpret_cast<int4*>(codes);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        int4 v = cv[i];
        v.x = __ldg(slot2code + v.x);
        v.y = __ldg(slot2code + v.y);
        v.z = __ldg(slot2code + v.z);
        v.w = __ldg(slot2code + v.w);
        cv[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n_vec << 2) + threadIdx.x;
        codes[i] = __ldg(slot2code + codes[i]);
    }
}

// Orders the gathered names bytewise ON THE DEVICE (at most DR_MAX of them): entry e's code is the number of names
// that compare below it.  Declines (ok = 0) when there are more names, or the first-round buffers overflowed, or the
// table is too full — the host path then redoes the encode.
constexpr int DR_MAX = 1024;

__device__ __forceinline__ bool name_less(const uint8_t* __restrict__ a, int la, const uint8_t* __restrict__ b, int lb) {
    const int m = la < lb ? la : lb;
    for (int i = 0; i < m; i++)
        if (a[i] != b[i]) return a[i] < b[i];
    return la < lb;
}

__global__ void __launch_bounds__(DR_MAX)
dict_rank_kernel(const unsigned long long* __restrict__ cursor /* [0] bytes [1] entries [2] distinct */, const uint8_t* __restrict__ name_buf,
                 uint32_t name_cap, const int32_t* __restrict__ ent, uint32_t max_entries, uint32_t cap, int32_t* __restrict__ slot2code,
                 int32_t* __restrict__ ent_rank, int* __restrict__ ok_flag) {
    const unsigned long long n = cursor[1], bytes = cursor[0], distinct = cursor[2] & 0xffffffffull;
    const bool ok = n <= (unsigned long long)DR_MAX && n <= max_entries && bytes <= name_cap && distinct * 2 <= cap && n == distinct;
    if (threadIdx.x == 0) *ok_flag = ok ? 1 : 0;
    if (!ok || threadIdx.x >= n) return;
    const int e = threadIdx.x;
    const uint8_t* me = name_buf + ent[max_entries + e];
    const int len = ent[2 * max_entries + e];
    int rank = 0;
    for (int f = 0; f < (int)n; f++)
        if (f != e && name_less(name_buf + ent[max_entries + f], ent[2 * max_entries + f], me, len)) rank++;
    slot2code[ent[e]] = rank;
    ent_rank[e] = rank;
}

}  // namespace

// ---- asynchronous encode: no host round trip ---------------------------------------------------------------------------
// Enqueues insert, collect, device-side ranking and remap, plus the copies that bring the names back; the caller
// synchronises later (with whatever it waits for anyway) and then calls giqlb_dict_finish.
struct giqlb_dict_pending {
    unsigned long long* slots = nullptr;
    uint8_t* name_buf = nullptr;
    int32_t *ent = nullptr, *slot2code = nullptr, *ent_rank = nullptr;
    int* ok_flag = nullptr;
    uint32_t cap = 0, max_entries = 0, name_cap = 0;
};

int giqlb_dict_encode_async(giqlb_ctx* ctx, const int32_t* offs_l, const uint8_t* bytes_l, int64_t n_l, const int32_t* offs_r,
                            const uint8_t* bytes_r, int64_t n_r, int32_t* codes_l, int32_t* codes_r, giqlb_dict_pending** out,
                            bool defer_remap) {
    *out = nullptr;
    if (n_l + n_r == 0 || !ctx->h_stage) return GIQLB_ERR_INVALID;   // caller uses the synchronous path
    if (((reinterpret_cast<uintptr_t>(codes_l) | reinterpret_cast<uintptr_t>(codes_r)) & 15) != 0) return GIQLB_ERR_INVALID;
    begin_call(ctx);
    giqlb_dict_pending* p = new (std::nothrow) giqlb_dict_pending();
    if (!p) return GIQLB_ERR_NOMEM;
    p->cap = 1u << 14;
    p->max_entries = DR_MAX;
    p->name_cap = 32768;
    StrTable A{offs_l, bytes_l, n_l}, B{offs_r, bytes_r, n_r};
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 24);
    unsigned int* n_distinct = reinterpret_cast<unsigned int*>(ctx->d_scalars + 26);
    int st = dev_alloc(ctx, &p->slots, (size_t)p->cap * DS_WORDS);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &p->name_buf, p->name_cap);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &p->ent, 3 * (size_t)p->max_entries);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &p->slot2code, (size_t)p->cap);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &p->ent_rank, (size_t)p->max_entries);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &p->ok_flag, 4);
    if (st != GIQLB_OK) {
        giqlb_dict_finish(ctx, p, nullptr, nullptr);
        return st;
    }
    cudaMemsetAsync(p->slots, 0, (size_t)p->cap * DS_WORDS * 8, ctx->stream);
    cudaMemsetAsync(p->slot2code, 0xFF, (size_t)p->cap * 4, ctx->stream);
    cudaMemsetAsync(cursor, 0, 3 * 8, ctx->stream);
    const int64_t chunks = ceil_div(n_l, DI_ROWS) + ceil_div(n_r, DI_ROWS), cap_blocks = (int64_t)ctx->sm_count * 5;   // 5 CTAs of 43.1 KB shared memory are resident per SM
    const unsigned grid = (unsigned)(chunks < cap_blocks ? chunks : cap_blocks);
    {
        PhaseSpan ps(ctx, GIQLB_PH_DICT);
        dict_smem_opt_in(ctx);
        dict_insert_kernel<<<grid, DI_THREADS, DI_SMEM, ctx->stream>>>(A, B, p->slots, p->cap - 1, codes_l, codes_r, n_distinct);
        dict_collect_kernel<<<(p->cap + 255) / 256, 256, 0, ctx->stream>>>(A, B, p->slots, p->cap, p->name_buf, p->name_cap, cursor, p->ent,
                                                                           p->max_entries);
        dict_rank_kernel<<<1, DR_MAX, 0, ctx->stream>>>(cursor, p->name_buf, p->name_cap, p->ent, p->max_entries, p->cap, p->slot2code,
                                                        p->ent_rank, p->ok_flag);
        const int64_t rb = ceil_div(n_l > n_r ? n_l : n_r, 256 * 16), rcap = (int64_t)ctx->sm_count * 8;
        const unsigned rgrid = (unsigned)(rb < 1 ? 1 : (rb < rcap ? rb : rcap));
        const int remaps = defer_remap ? 0 : (n_l > 0) + (n_r > 0);
        if (!defer_remap && n_l > 0) dict_remap_kernel<<<rgrid, 256, 0, ctx->stream>>>(codes_l, n_l, p->slot2code, p->ok_flag);
        if (!defer_remap && n_r > 0) dict_remap_kernel<<<rgrid, 256, 0, ctx->stream>>>(codes_r, n_r, p->slot2code, p->ok_flag);
        ctx->launches_this_call += 3 + remaps;
        ctx->stats.kernels_total += 3 + remaps;
    }
    // names, entries, ranks and the verdict land in the context's PINNED staging buffer whenever the caller next waits
    unsigned char* h = reinterpret_cast<unsigned char*>(ctx->h_stage);
    cudaMemcpyAsync(h, cursor, 24, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(h + 32, p->ok_flag, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(h + 64, p->ent, 3 * (size_t)p->max_entries * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(h + 64 + 12 * (size_t)p->max_entries, p->ent_rank, (size_t)p->max_entries * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(h + 64 + 16 * (size_t)p->max_entries, p->name_buf, p->name_cap, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaGetLastError() != cudaSuccess) {
        giqlb_dict_finish(ctx, p, nullptr, nullptr);
        return giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode_async: launch failed");
    }
    static_assert(64 + 16 * DR_MAX + 32768 <= GIQLB_STAGE_BYTES, "staging buffer too small");
    ctx->stats.rows_in[0] = n_l;
    ctx->stats.rows_in[1] = n_r;
    ctx->stats.rows_out = n_l + n_r;
    ctx->stats.algorithmic_bytes = 8 * (n_l + n_r);
    *out = p;
    return end_call(ctx);
}

// After the stream has been synchronised: builds the dictionary from the staged names (or reports that the device
// declined: *fallback = 1, *dict_out = nullptr) and releases the scratch.  dict_out == nullptr just releases.
void giqlb_dict_pending_tables(const giqlb_dict_pending* p, const int32_t** slot2code, const int** ok_flag) {
    *slot2code = p->slot2code;
    *ok_flag = p->ok_flag;
}

int giqlb_dict_finish(giqlb_ctx* ctx, giqlb_dict_pending* p, giqlb_dict** dict_out, int* fallback) {
    if (!p) return GIQLB_OK;
    int st = GIQLB_OK;
    if (dict_out) {
        *dict_out = nullptr;
        const unsigned char* h = reinterpret_cast<const unsigned char*>(ctx->h_stage);
        const unsigned long long n = reinterpret_cast<const unsigned long long*>(h)[1];
        const int ok = *reinterpret_cast<const int*>(h + 32);
        if (fallback) *fallback = ok ? 0 : 1;
        if (ok) {
            const int32_t* ent = reinterpret_cast<const int32_t*>(h + 64);
            const int32_t* rank = reinterpret_cast<const int32_t*>(h + 64 + 12 * (size_t)p->max_entries);
            const char* names = reinterpret_cast<const char*>(h + 64 + 16 * (size_t)p->max_entries);
            giqlb_dict* d = new (std::nothrow) giqlb_dict();
            if (!d) st = GIQLB_ERR_NOMEM;
            else {
                d->names.resize((size_t)n);
                for (unsigned long long e = 0; e < n; e++)
                    d->names[(size_t)rank[e]] = std::string(names + ent[p->max_entries + e], (size_t)ent[2 * p->max_entries + e]);
                *dict_out = d;
            }
        }
    }
    dev_free(ctx, p->slots); dev_free(ctx, p->name_buf); dev_free(ctx, p->ent); dev_free(ctx, p->slot2code);
    dev_free(ctx, p->ent_rank); dev_free(ctx, p->ok_flag);
    delete p;
    return st;
}

extern "C" int giqlb_dict_encode(giqlb_ctx* ctx, const int32_t* offs_l, const uint8_t* bytes_l, int64_t n_l,
                                 const int32_t* offs_r, const uint8_t* bytes_r, int64_t n_r, int32_t* codes_l,
                                 int32_t* codes_r, giqlb_dict** out) {
    if (!ctx || !out || n_l < 0 || n_r < 0 || (n_l > 0 && (!offs_l || !codes_l)) || (n_r > 0 && (!offs_r || !codes_r)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "dict_encode: bad argument");
    if (((reinterpret_cast<uintptr_t>(codes_l) | reinterpret_cast<uintptr_t>(codes_r)) & 15) != 0)
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "dict_encode: code arrays must be 16-byte aligned");
    begin_call(ctx);
    *out = nullptr;
    giqlb_dict* d = new (std::nothrow) giqlb_dict();
    if (!d) return GIQLB_ERR_NOMEM;
    const int64_t n = n_l + n_r;
    if (n == 0) {
        *out = d;
        return end_call(ctx);
    }
    StrTable A{offs_l, bytes_l, n_l}, B{offs_r, bytes_r, n_r};
    unsigned long long* slots = nullptr;
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 24);   // [0] bytes [1] entries [2] distinct
    unsigned int* n_distinct = reinterpret_cast<unsigned int*>(ctx->d_scalars + 26);
    uint8_t* name_buf = nullptr;
    int32_t *ent = nullptr, *slot2code = nullptr;
    int st = GIQLB_OK;
    uint32_t cap = 1u << 14;
    size_t max_entries = DC_MAX_ENTRIES, name_cap = DC_NAME_CAP;
    const int64_t chunks = ceil_div(n_l, DI_ROWS) + ceil_div(n_r, DI_ROWS), cap_blocks = (int64_t)ctx->sm_count * 5;   // 5 CTAs of 43.1 KB shared memory are resident per SM
    const unsigned grid = (unsigned)(chunks < cap_blocks ? chunks : cap_blocks);
    std::vector<uint8_t> h_names;
    std::vector<int32_t> h_ent;
    uint64_t distinct = 0;
    // One round trip in the common case: insert + collect with generous fixed buffers, then ONE synchronisation
    // fetching the counts, the entry table and the name bytes.  A table that turns out too full, or names that
    // overflow the buffers, repeat the round with exact sizes.
    for (int round = 0; round < 4; round++) {
        dev_free(ctx, slots); dev_free(ctx, name_buf); dev_free(ctx, ent);
        slots = nullptr; name_buf = nullptr; ent = nullptr;
        st = dev_alloc(ctx, &slots, (size_t)cap * DS_WORDS);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &name_buf, name_cap);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &ent, 3 * max_entries);
        if (st != GIQLB_OK) break;
        cudaMemsetAsync(slots, 0, (size_t)cap * DS_WORDS * 8, ctx->stream);
        cudaMemsetAsync(cursor, 0, 3 * 8, ctx->stream);
        {
            PhaseSpan ps(ctx, GIQLB_PH_DICT);
            dict_smem_opt_in(ctx);
            dict_insert_kernel<<<grid, DI_THREADS, DI_SMEM, ctx->stream>>>(A, B, slots, cap - 1, codes_l, codes_r, n_distinct);
            dict_collect_kernel<<<(cap + 255) / 256, 256, 0, ctx->stream>>>(A, B, slots, cap, name_buf, name_cap, cursor, ent,
                                                                            (uint32_t)max_entries);
            ctx->launches_this_call += 2;
            ctx->stats.kernels_total += 2;
        }
        if (cudaGetLastError() != cudaSuccess) { st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode: launch failed"); break; }
        h_names.resize(name_cap);
        h_ent.resize(3 * max_entries);
        // small tables ride along with the counters (pageable destinations: the copies complete at the sync below)
        const bool speculative = max_entries * 12 + name_cap <= (size_t)256 * 1024;
        if (speculative) {
            cudaMemcpyAsync(h_ent.data(), ent, 3 * max_entries * 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(h_names.data(), name_buf, name_cap, cudaMemcpyDeviceToHost, ctx->stream);
        }
        st = fetch_scalars(ctx, cursor, 3);
        if (st != GIQLB_OK) break;
        const size_t need_bytes = (size_t)ctx->h_scalars[0];
        distinct = (uint64_t)(ctx->h_scalars[2] & 0xffffffffll);
        bool redo = false;
        if (distinct * 2 > cap) {   // too full for cheap probing: grow and redo
            if (cap >= (1u << 30)) { st = giqlb_fail(ctx, GIQLB_ERR_RANGE, "dict_encode: more than 2^29 distinct chromosome names"); break; }
            while (distinct * 2 > cap && cap < (1u << 30)) cap <<= 2;
            redo = true;
        }
        if (distinct > max_entries || need_bytes > name_cap) {
            max_entries = std::max<size_t>(max_entries, (size_t)distinct * 2);
            name_cap = std::max<size_t>(name_cap, need_bytes * 2);
            redo = true;
        }
        if (!redo) {
            if (!speculative) {
                cudaMemcpyAsync(h_ent.data(), ent, 3 * max_entries * 4, cudaMemcpyDeviceToHost, ctx->stream);
                cudaMemcpyAsync(h_names.data(), name_buf, name_cap, cudaMemcpyDeviceToHost, ctx->stream);
                if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode: name copy failed");
            }
            break;
        }
        if (round == 3) st = giqlb_fail(ctx, GIQLB_ERR_INTERNAL, "dict_encode: sizing did not converge");
    }
    if (st == GIQLB_OK) {
        const int32_t *h_slot = h_ent.data(), *h_off = h_ent.data() + max_entries, *h_len = h_ent.data() + 2 * max_entries;
        std::vector<uint32_t> order(distinct);
        for (uint32_t i = 0; i < distinct; i++) order[i] = i;
        std::vector<std::string> names(distinct);
        for (uint32_t i = 0; i < distinct; i++)
            names[i] = std::string(reinterpret_cast<const char*>(h_names.data()) + h_off[i], (size_t)h_len[i]);
        // bytewise (unsigned char) order, the order VARCHAR comparison gives for UTF-8
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            const std::string &x = names[a], &y = names[b];
            size_t m = x.size() < y.size() ? x.size() : y.size();
            int c = memcmp(x.data(), y.data(), m);
            return c != 0 ? c < 0 : x.size() < y.size();
        });
        // slot -> code table through the context's pinned staging buffer: the upload and the remap kernels are
        // enqueued without waiting (the next use of the staging buffer comes after a synchronisation)
        int32_t* h_s2c = ctx->h_stage && (size_t)cap * 4 <= GIQLB_STAGE_BYTES ? reinterpret_cast<int32_t*>(ctx->h_stage) : nullptr;
        std::vector<int32_t> pageable;
        if (!h_s2c) {
            pageable.assign(cap, -1);
            h_s2c = pageable.data();
        } else {
            std::fill(h_s2c, h_s2c + cap, -1);
        }
        d->names.resize(distinct);
        for (uint32_t code = 0; code < distinct; code++) {
            d->names[code] = names[order[code]];
            h_s2c[h_slot[order[code]]] = (int32_t)code;
        }
        st = dev_alloc(ctx, &slot2code, (size_t)cap);
        if (st == GIQLB_OK) {
            cudaMemcpyAsync(slot2code, h_s2c, (size_t)cap * 4, cudaMemcpyHostToDevice, ctx->stream);
            PhaseSpan ps(ctx, GIQLB_PH_DICT);
            const int64_t rb = ceil_div(n_l > n_r ? n_l : n_r, 256 * 16), rcap = (int64_t)ctx->sm_count * 8;
            const unsigned rgrid = (unsigned)(rb < 1 ? 1 : (rb < rcap ? rb : rcap));
            if (n_l > 0) {
                dict_remap_kernel<<<rgrid, 256, 0, ctx->stream>>>(codes_l, n_l, slot2code, nullptr);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
            if (n_r > 0) {
                dict_remap_kernel<<<rgrid, 256, 0, ctx->stream>>>(codes_r, n_r, slot2code, nullptr);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
            if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode: remap failed");
            if (!pageable.empty() && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode: remap failed");
        }
    }
    dev_free(ctx, slots);
    dev_free(ctx, name_buf);
    dev_free(ctx, ent);
    dev_free(ctx, slot2code);
    if (st != GIQLB_OK) {
        delete d;
        return st;
    }
    ctx->stats.rows_in[0] = n_l;
    ctx->stats.rows_in[1] = n_r;
    ctx->stats.rows_out = n;
    ctx->stats.algorithmic_bytes = 8 * n;
    *out = d;
    return end_call(ctx);
}

extern "C" int giqlb_dict_size(const giqlb_dict* d) { return d ? (int)d->names.size() : -1; }

extern "C" int giqlb_dict_get(const giqlb_dict* d, int32_t code, const char** bytes_host, int32_t* len_host) {
    if (!d || code < 0 || (size_t)code >= d->names.size()) return GIQLB_ERR_INVALID;
    if (bytes_host) *bytes_host = d->names[code].data();
    if (len_host) *len_host = (int32_t)d->names[code].size();
    return GIQLB_OK;
}

extern "C" int giqlb_dict_free(giqlb_dict* d) {
    delete d;
    return GIQLB_OK;
}
