// dict.cu — chromosome dictionary encoding on the device.
//
// The reference compares chromosomes as VARCHAR (`l.chrom = r.chrom`, abdenlab/giql
// src/giql/expanders/intersects.py:151) and the DuckDB plan re-scans both tables once per distinct name
// (src/giql/expanders/_per_chrom.py:27,46-74).  Here both tables are encoded ONCE against one shared
// dictionary whose codes preserve bytewise string order (MERGE's default ORDER BY chrom sorts VARCHAR,
// src/giql/expanders/merge.py:429-437), so every later kernel works on int32 codes.
//
// Pass 1: every row probes an open-addressing table whose slots hold (representative row + 1); a row either
// claims an empty slot with one CAS or compares its bytes with the representative's — exact, no reliance on
// hash uniqueness.  The row's slot is parked in the output code array.  Pass 2 gathers the distinct names
// into a small buffer for the host, which sorts them and uploads a slot -> code table.  Pass 3 rewrites the
// parked slots into codes.  Traffic: offsets + bytes once, 8 B/row of codes.
#include <algorithm>

#include "common.cuh"

namespace {

struct StrTable {
    const int32_t* offs;
    const uint8_t* bytes;
    int64_t n;
};

__device__ __forceinline__ uint64_t fnv1a(const uint8_t* __restrict__ p, int len) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < len; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h ^ (h >> 29);
}

__device__ __forceinline__ bool same_bytes(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int len) {
    for (int i = 0; i < len; i++)
        if (a[i] != b[i]) return false;
    return true;
}

__global__ void __launch_bounds__(256)
dict_insert_kernel(StrTable A, StrTable B, unsigned long long* __restrict__ slots, uint32_t cap_mask,
                   int32_t* __restrict__ slot_a, int32_t* __restrict__ slot_b, unsigned int* __restrict__ n_distinct) {
    const int64_t n = A.n + B.n;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x) {
        const bool in_b = g >= A.n;
        const StrTable& T = in_b ? B : A;
        const int64_t i = in_b ? g - A.n : g;
        const int32_t o0 = T.offs[i], len = T.offs[i + 1] - o0;
        const uint8_t* s = T.bytes + o0;
        uint32_t slot = (uint32_t)fnv1a(s, len) & cap_mask;
        for (uint32_t probes = 0; probes <= cap_mask; probes++) {
            unsigned long long w = *(volatile unsigned long long*)&slots[slot];
            if (w == 0) {
                unsigned long long prev = atomicCAS(&slots[slot], 0ull, (unsigned long long)g + 1ull);
                if (prev == 0) {
                    atomicAdd(n_distinct, 1u);
                    break;
                }
                w = prev;
            }
            const int64_t rg = (int64_t)w - 1;
            const StrTable& RT = rg >= A.n ? B : A;
            const int64_t ri = rg >= A.n ? rg - A.n : rg;
            const int32_t r0 = RT.offs[ri], rlen = RT.offs[ri + 1] - r0;
            if (rlen == len && same_bytes(s, RT.bytes + r0, len)) break;
            slot = (slot + 1) & cap_mask;
        }
        (in_b ? slot_b : slot_a)[i] = (int32_t)slot;
    }
}

// one thread per slot: copy the representative's bytes into the name buffer
__global__ void dict_collect_kernel(StrTable A, StrTable B, const unsigned long long* __restrict__ slots, uint32_t cap,
                                    uint8_t* __restrict__ name_buf, unsigned long long name_cap,
                                    unsigned long long* __restrict__ cursor /* [0] bytes, [1] entries */,
                                    int32_t* __restrict__ ent_slot, int32_t* __restrict__ ent_off, int32_t* __restrict__ ent_len,
                                    uint32_t max_entries) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    unsigned long long w = slots[slot];
    if (w == 0) return;
    const int64_t rg = (int64_t)w - 1;
    const StrTable& RT = rg >= A.n ? B : A;
    const int64_t ri = rg >= A.n ? rg - A.n : rg;
    const int32_t r0 = RT.offs[ri], len = RT.offs[ri + 1] - r0;
    unsigned long long off = atomicAdd(&cursor[0], (unsigned long long)len);
    unsigned long long e = atomicAdd(&cursor[1], 1ull);
    if (e < max_entries) {
        ent_slot[e] = (int32_t)slot;
        ent_off[e] = (int32_t)off;
        ent_len[e] = len;
    }
    if (off + len <= name_cap)
        for (int i = 0; i < len; i++) name_buf[off + i] = RT.bytes[r0 + i];
}

__global__ void __launch_bounds__(256)
dict_remap_kernel(int32_t* __restrict__ codes, int64_t n, const int32_t* __restrict__ slot2code) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        codes[i] = __ldg(slot2code + codes[i]);
}

}  // namespace

extern "C" int giqlb_dict_encode(giqlb_ctx* ctx, const int32_t* offs_l, const uint8_t* bytes_l, int64_t n_l,
                                 const int32_t* offs_r, const uint8_t* bytes_r, int64_t n_r, int32_t* codes_l,
                                 int32_t* codes_r, giqlb_dict** out) {
    if (!ctx || !out || n_l < 0 || n_r < 0 || (n_l > 0 && (!offs_l || !codes_l)) || (n_r > 0 && (!offs_r || !codes_r)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "dict_encode: bad argument");
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
    int st = GIQLB_OK;
    uint32_t cap = 1u << 14;
    int64_t blocks = ceil_div(n, 256 * 4), cap_blocks = (int64_t)ctx->sm_count * 16;
    const unsigned grid = (unsigned)(blocks < cap_blocks ? blocks : cap_blocks);
    uint64_t distinct = 0;
    while (true) {
        st = dev_alloc(ctx, &slots, (size_t)cap);
        if (st != GIQLB_OK) break;
        cudaMemsetAsync(slots, 0, (size_t)cap * 8, ctx->stream);
        cudaMemsetAsync(cursor, 0, 3 * 8, ctx->stream);
        PhaseSpan ps(ctx, GIQLB_PH_DICT);
        dict_insert_kernel<<<grid, 256, 0, ctx->stream>>>(A, B, slots, cap - 1, codes_l, codes_r, n_distinct);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        if (cudaGetLastError() != cudaSuccess) { st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode: insert launch failed"); break; }
        st = fetch_scalars(ctx, cursor, 3);
        if (st != GIQLB_OK) break;
        distinct = (uint64_t)(ctx->h_scalars[2] & 0xffffffffll);
        if (distinct * 2 <= cap) break;
        dev_free(ctx, slots);   // too full for cheap probing: grow and redo
        slots = nullptr;
        if (cap >= (1u << 30)) { st = giqlb_fail(ctx, GIQLB_ERR_RANGE, "dict_encode: more than 2^29 distinct chromosome names"); break; }
        cap <<= 2;
    }
    if (st != GIQLB_OK) {
        dev_free(ctx, slots);
        delete d;
        return st;
    }
    // gather distinct names; retry once with the exact byte size if the first buffer was too small
    size_t name_cap = 1 << 16;
    std::vector<uint8_t> h_names;
    std::vector<int32_t> h_slot(distinct), h_off(distinct), h_len(distinct);
    uint8_t* name_buf = nullptr;
    int32_t *ent_slot = nullptr, *ent_off = nullptr, *ent_len = nullptr, *slot2code = nullptr;
    for (int attempt = 0; attempt < 2 && st == GIQLB_OK; attempt++) {
        st = dev_alloc(ctx, &name_buf, name_cap);
        if (st == GIQLB_OK && !ent_slot) st = dev_alloc(ctx, &ent_slot, (size_t)distinct);
        if (st == GIQLB_OK && !ent_off) st = dev_alloc(ctx, &ent_off, (size_t)distinct);
        if (st == GIQLB_OK && !ent_len) st = dev_alloc(ctx, &ent_len, (size_t)distinct);
        if (st != GIQLB_OK) break;
        cudaMemsetAsync(cursor, 0, 2 * 8, ctx->stream);
        PhaseSpan ps(ctx, GIQLB_PH_DICT);
        dict_collect_kernel<<<(cap + 255) / 256, 256, 0, ctx->stream>>>(A, B, slots, cap, name_buf, name_cap, cursor, ent_slot,
                                                                        ent_off, ent_len, (uint32_t)distinct);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        if (cudaGetLastError() != cudaSuccess) { st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode: collect launch failed"); break; }
        st = fetch_scalars(ctx, cursor, 2);
        if (st != GIQLB_OK) break;
        size_t need = (size_t)ctx->h_scalars[0];
        if (need <= name_cap) {
            h_names.resize(need ? need : 1);
            if (need) cudaMemcpyAsync(h_names.data(), name_buf, need, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(h_slot.data(), ent_slot, distinct * 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(h_off.data(), ent_off, distinct * 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(h_len.data(), ent_len, distinct * 4, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode: name copy failed");
            dev_free(ctx, name_buf);
            name_buf = nullptr;
            break;
        }
        dev_free(ctx, name_buf);
        name_buf = nullptr;
        name_cap = need;
        if (attempt == 1) st = giqlb_fail(ctx, GIQLB_ERR_INTERNAL, "dict_encode: name buffer sizing did not converge");
    }
    if (st == GIQLB_OK) {
        std::vector<uint32_t> order(distinct);
        for (uint32_t i = 0; i < distinct; i++) order[i] = i;
        auto name_of = [&](uint32_t e) { return std::string(reinterpret_cast<const char*>(h_names.data()) + h_off[e], (size_t)h_len[e]); };
        std::vector<std::string> names(distinct);
        for (uint32_t i = 0; i < distinct; i++) names[i] = name_of(i);
        // bytewise (unsigned char) order == std::string::compare with char_traits<char> only for ASCII; compare as unsigned
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            const std::string &x = names[a], &y = names[b];
            size_t m = x.size() < y.size() ? x.size() : y.size();
            int c = memcmp(x.data(), y.data(), m);
            return c != 0 ? c < 0 : x.size() < y.size();
        });
        std::vector<int32_t> h_s2c(cap, -1);
        d->names.resize(distinct);
        for (uint32_t code = 0; code < distinct; code++) {
            d->names[code] = names[order[code]];
            h_s2c[h_slot[order[code]]] = (int32_t)code;
        }
        st = dev_alloc(ctx, &slot2code, (size_t)cap);
        if (st == GIQLB_OK) {
            cudaMemcpyAsync(slot2code, h_s2c.data(), (size_t)cap * 4, cudaMemcpyHostToDevice, ctx->stream);
            PhaseSpan ps(ctx, GIQLB_PH_DICT);
            if (n_l > 0) {
                dict_remap_kernel<<<grid, 256, 0, ctx->stream>>>(codes_l, n_l, slot2code);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
            if (n_r > 0) {
                dict_remap_kernel<<<grid, 256, 0, ctx->stream>>>(codes_r, n_r, slot2code);
                ctx->launches_this_call++;
                ctx->stats.kernels_total++;
            }
            // h_s2c is pageable: wait before it goes out of scope
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
                st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "dict_encode: remap failed");
        }
    }
    dev_free(ctx, slots);
    dev_free(ctx, name_buf);
    dev_free(ctx, ent_slot);
    dev_free(ctx, ent_off);
    dev_free(ctx, ent_len);
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
