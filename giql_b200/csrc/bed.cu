// bed.cu — BED text ingest on the device (the data format on the input side of the path).
//
// The reference never reads files itself: its users load BED into DuckDB with read_csv and GIQL sees VARCHAR / INTEGER
// columns (abdenlab/giql docs/recipes, demo.ipynb).  This is that loader for the engine: the raw text is copied to the
// device once and parsed there into the columns every other entry point takes — chromosome names in Arrow utf8 layout,
// int32 start / end, optional name (utf8), score (int32) and strand codes.
//
//   1. bed_count_kernel        newlines per 16-byte word, grid-stride                       (read the text once)
//   2. launch_scan + OutLines  exclusive scan of the per-word newline counts -> line starts  (read the text once more)
//   3. bed_parse_kernel        a CTA stages the bytes of 256 consecutive lines in shared memory (coalesced 16-byte
//                              loads) and one thread walks one line: tab-separated fields, decimal integers
//   4. launch_scan x2          name-length scans -> utf8 offsets;  5. bed_copy_kernel: the name bytes
// Traffic: 3 reads of the text + the columns written.  Lines end in \n or \r\n; fields are TAB separated (bedtools'
// rule); header lines (track / browser / #) are stripped by the host before upload; blank or malformed lines are an
// error that names the first offending line.
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int BP_THREADS = 256;
constexpr int BP_STAGE = 24576;   // staged bytes per CTA (256 lines of <= 96 bytes on average; longer chunks read global memory)

__device__ __forceinline__ uint32_t newline_mask(uint32_t w) { return __vcmpeq4(w, 0x0A0A0A0Au); }   // 0xFF per '\n' byte

// 16 text bytes at word index i (bytes beyond n read as 0); the text base is 16-byte aligned
__device__ __forceinline__ uint4 load_word(const uint8_t* __restrict__ text, int64_t n, int64_t i) {
    const int64_t p = i * 16;
    if (p + 16 <= n) return __ldg(reinterpret_cast<const uint4*>(text + p));
    uint32_t w[4] = {0, 0, 0, 0};
    for (int64_t k = p; k < n; k++) w[(k - p) >> 2] |= (uint32_t)text[k] << (8 * ((k - p) & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ int word_newlines(uint4 v) {
    return (__popc(newline_mask(v.x)) + __popc(newline_mask(v.y)) + __popc(newline_mask(v.z)) + __popc(newline_mask(v.w))) >> 3;
}

__global__ void __launch_bounds__(256)
bed_count_kernel(const uint8_t* __restrict__ text, int64_t n, unsigned long long* __restrict__ total) {
    const int64_t words = (n + 15) >> 4;
    unsigned long long cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x)
        cnt += word_newlines(load_word(text, n, i));
    cnt = warp_reduce_add(cnt);
    if (lane_id() == 0 && cnt) atomicAdd(total, cnt);
}

struct InNewlines {
    const uint8_t* text;
    int64_t n;
    __device__ unsigned long long operator()(int64_t i) const { return (unsigned long long)word_newlines(load_word(text, n, i)); }
};
struct OutLines {
    const uint8_t* text;
    int64_t n;
    int32_t* line_start;   // [n_lines + 1]; entry 0 is written by the parse call
    int64_t n_lines;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        if (!v) return;
        const uint4 q = load_word(text, n, i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        int64_t k = (int64_t)excl;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t m = newline_mask(w[j]);
            while (m) {
                const int b = (__ffs(m) - 1) >> 3;
                m &= ~(0xFFu << (8 * b));
                k++;
                if (k <= n_lines) line_start[k] = (int32_t)(i * 16 + j * 4 + b + 1);
            }
        }
    }
};

struct BedOut {
    int32_t* start;
    int32_t* end;
    int32_t* chrom_len;     // scratch [n + 1]
    int32_t* name_pos;      // scratch [n] or null
    int32_t* name_len;      // scratch [n + 1] or null
    int32_t* score;         // or null
    uint8_t* strand;        // or null
};

struct Cursor {
    const uint8_t* p;       // bytes of the line (shared or global)
    int len;
    int at;
    __device__ __forceinline__ int field_end() const {
        int k = at;
        while (k < len && p[k] != '\t') k++;
        return k;
    }
};

// decimal int32 in [a, b); false on an empty field, a non-digit or overflow
__device__ __forceinline__ bool parse_int(const uint8_t* p, int a, int b, int32_t* out) {
    bool neg = false;
    if (a < b && (p[a] == '-' || p[a] == '+')) {
        neg = p[a] == '-';
        a++;
    }
    if (a >= b || b - a > 10) return false;
    long long v = 0;
    for (int k = a; k < b; k++) {
        const int d = (int)p[k] - '0';
        if (d < 0 || d > 9) return false;
        v = v * 10 + d;
    }
    if (neg) v = -v;
    if (v > 2147483647ll || v < -2147483648ll) return false;
    *out = (int32_t)v;
    return true;
}

__global__ void __launch_bounds__(BP_THREADS)
bed_parse_kernel(const uint8_t* __restrict__ text, int64_t n_bytes, const int32_t* __restrict__ line_start, int64_t n_lines, int n_fields,
                 BedOut o, unsigned long long* __restrict__ first_bad) {
    __shared__ __align__(16) uint8_t s_text[BP_STAGE + 16];
    const int64_t r0 = (int64_t)blockIdx.x * BP_THREADS;
    const int rows = (int)min((int64_t)BP_THREADS, n_lines - r0);
    const int64_t b0 = line_start[r0], b1 = min((int64_t)line_start[r0 + rows], n_bytes);
    const int64_t a0 = b0 & ~15ll;
    const bool staged = b1 - a0 <= BP_STAGE;
    if (staged) {
        const int n_vec = (int)((b1 - a0 + 15) >> 4);
        for (int i = threadIdx.x; i < n_vec; i += BP_THREADS)
            reinterpret_cast<uint4*>(s_text)[i] = load_word(text, n_bytes, (a0 >> 4) + i);
    }
    __syncthreads();
    if ((int)threadIdx.x >= rows) return;
    const int64_t row = r0 + threadIdx.x;
    const int64_t ls = line_start[row];
    int len = (int)(min((int64_t)line_start[row + 1] - 1, n_bytes) - ls);      // without the '\n'
    Cursor c{staged ? s_text + (ls - a0) : text + ls, len, 0};
    if (c.len > 0 && c.p[c.len - 1] == '\r') c.len--;
    bool ok = c.len > 0;
    int32_t v_start = 0, v_end = 0, v_score = 0, npos = 0, nlen = 0;
    uint8_t v_strand = GIQLB_STRAND_NULL;
    // chrom
    int e = c.field_end();
    const int clen = e;
    ok = ok && clen > 0 && e < c.len;
    c.at = e + 1;
    // start
    e = c.field_end();
    ok = ok && parse_int(c.p, c.at, e, &v_start) && e < c.len;
    c.at = e + 1;
    // end
    e = c.field_end();
    ok = ok && parse_int(c.p, c.at, e, &v_end) && (n_fields == 3 || e < c.len);
    c.at = e + 1;
    if (n_fields >= 4 && ok) {   // name
        e = c.field_end();
        npos = (int32_t)(ls + c.at);
        nlen = e - c.at;
        ok = n_fields == 4 || e < c.len;
        c.at = e + 1;
    }
    if (n_fields >= 5 && ok) {   // score: an integer, or '.' for "no score" (-> 0)
        e = c.field_end();
        if (e - c.at == 1 && c.p[c.at] == '.') v_score = 0;
        else ok = parse_int(c.p, c.at, e, &v_score);
        ok = ok && (n_fields == 5 || e < c.len);
        c.at = e + 1;
    }
    if (n_fields >= 6 && ok) {   // strand
        e = c.field_end();
        ok = e - c.at == 1;
        if (ok) {
            const uint8_t ch = c.p[c.at];
            v_strand = ch == '+' ? 0 : ch == '-' ? 1 : ch == '.' ? 2 : GIQLB_STRAND_NULL;
            ok = v_strand != GIQLB_STRAND_NULL;
        }
    }
    if (!ok) atomicMin(first_bad, (unsigned long long)row);
    o.start[row] = v_start;
    o.end[row] = v_end;
    o.chrom_len[row] = ok ? clen : 0;
    if (o.name_len) {
        o.name_pos[row] = npos;
        o.name_len[row] = ok ? nlen : 0;
    }
    if (o.score) o.score[row] = v_score;
    if (o.strand) o.strand[row] = v_strand;
}

struct InLen {
    const int32_t* len;
    __device__ unsigned long long operator()(int64_t i) const { return (unsigned long long)len[i]; }
};
struct OutOffsets32 {
    int32_t* offs;          // [n + 1]
    unsigned long long* total;
    int64_t n;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        offs[i] = (int32_t)min(excl, 0x7fffffffull);
        if (i == n - 1) {
            offs[n] = (int32_t)min(excl + v, 0x7fffffffull);
            *total = excl + v;
        }
    }
};

// one thread per row: bytes [pos, pos + len) of the text -> out[offs[row] ...]
__global__ void bed_copy_kernel(const uint8_t* __restrict__ text, const int32_t* __restrict__ pos, const int32_t* __restrict__ offs,
                                int64_t n, uint8_t* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int32_t o = offs[row], len = offs[row + 1] - o;
    const uint8_t* src = text + pos[row];
    for (int k = 0; k < len; k++) out[o + k] = src[k];
}

}  // namespace

static int bed_check_text(giqlb_ctx* ctx, const uint8_t* text, int64_t n_bytes, const char* who) {
    if (!ctx || n_bytes < 0 || (n_bytes > 0 && !text)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "%s: bad argument", who);
    if (n_bytes >= (1ll << 31) - 16) return giqlb_fail(ctx, GIQLB_ERR_RANGE, "%s: at most 2 GiB of text per call (split the file at a line boundary)", who);
    if (reinterpret_cast<uintptr_t>(text) & 15) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "%s: the text must be 16-byte aligned", who);
    return GIQLB_OK;
}

extern "C" int giqlb_bed_count(giqlb_ctx* ctx, const uint8_t* text, int64_t n_bytes, int64_t* n_lines_host) {
    GIQLB_TRY(bed_check_text(ctx, text, n_bytes, "bed_count"));
    if (!n_lines_host) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "bed_count: null output");
    begin_call(ctx);
    *n_lines_host = 0;
    if (n_bytes == 0) return end_call(ctx);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 8);
    CUDA_TRY(ctx, cudaMemsetAsync(d, 0, 16, ctx->stream));
    {
        PhaseSpan ps(ctx, GIQLB_PH_OTHER);
        const int64_t words = (n_bytes + 15) >> 4;
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(words, 256), 148 * 16);
        bed_count_kernel<<<grid, 256, 0, ctx->stream>>>(text, n_bytes, d);
        LAUNCH_CHECK(ctx);
        // the last byte decides whether an unterminated final line exists
        CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(d + 1), text + n_bytes - 1, 1, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    GIQLB_TRY(fetch_scalars(ctx, d, 2));
    *n_lines_host = ctx->h_scalars[0] + ((ctx->h_scalars[1] & 0xFF) != '\n' ? 1 : 0);
    ctx->stats.rows_in[0] = n_bytes;
    ctx->stats.rows_out = *n_lines_host;
    ctx->stats.algorithmic_bytes = n_bytes;
    return end_call(ctx);
}

extern "C" int giqlb_bed_parse(giqlb_ctx* ctx, const uint8_t* text, int64_t n_bytes, int64_t n_lines, int n_fields, int32_t* chrom_offsets,
                               uint8_t* chrom_bytes, int64_t chrom_capacity, int32_t* start, int32_t* end, int32_t* name_offsets,
                               uint8_t* name_bytes, int64_t name_capacity, int32_t* score, uint8_t* strand, int64_t* chrom_total_host,
                               int64_t* name_total_host) {
    using namespace giqlb_scan;
    GIQLB_TRY(bed_check_text(ctx, text, n_bytes, "bed_parse"));
    if (n_lines < 0 || n_fields < 3 || n_fields > 6 || !chrom_total_host || !name_total_host)
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "bed_parse: bad argument");
    if (n_lines > 0 && (!chrom_offsets || !start || !end)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "bed_parse: null column");
    const bool want_name = n_fields >= 4 && name_offsets != nullptr;
    if (n_fields < 5) score = nullptr;
    if (n_fields < 6) strand = nullptr;
    begin_call(ctx);
    *chrom_total_host = *name_total_host = 0;
    if (n_lines == 0) return end_call(ctx);
    int32_t *line_start = nullptr, *chrom_len = nullptr, *name_pos = nullptr, *name_len = nullptr;
    unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 8);   // [0] chrom bytes, [1] name bytes, [2] first bad line
    int st = dev_alloc(ctx, &line_start, (size_t)n_lines + 2);
    if (st == GIQLB_OK) st = dev_alloc(ctx, &chrom_len, (size_t)n_lines + 1);
    if (st == GIQLB_OK && want_name) st = dev_alloc(ctx, &name_pos, (size_t)n_lines + 1);
    if (st == GIQLB_OK && want_name) st = dev_alloc(ctx, &name_len, (size_t)n_lines + 1);
    auto cleanup = [&](int code) {
        dev_free(ctx, line_start);
        dev_free(ctx, chrom_len);
        dev_free(ctx, name_pos);
        dev_free(ctx, name_len);
        return code;
    };
    if (st != GIQLB_OK) return cleanup(st);
    {
        PhaseSpan ps(ctx, GIQLB_PH_OTHER);
        const unsigned long long init[3] = {0ull, 0ull, ~0ull};
        const int32_t ends[2] = {0, (int32_t)(n_bytes + 1)};   // line_start[0]; line_start[n_lines] when the last line is unterminated
        if (cudaMemcpyAsync(d, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(line_start, &ends[0], 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(line_start + n_lines, &ends[1], 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
            return cleanup(giqlb_fail(ctx, GIQLB_ERR_CUDA, "bed_parse: copy failed"));
        st = launch_scan<OpAddU64>(ctx, (n_bytes + 15) >> 4, InNewlines{text, n_bytes}, OutLines{text, n_bytes, line_start, n_lines});
        if (st != GIQLB_OK) return cleanup(st);
        const BedOut o{start, end, chrom_len, name_pos, want_name ? name_len : nullptr, score, strand};
        bed_parse_kernel<<<(unsigned)ceil_div(n_lines, BP_THREADS), BP_THREADS, 0, ctx->stream>>>(text, n_bytes, line_start, n_lines, n_fields, o, d + 2);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        st = launch_scan<OpAddU64>(ctx, n_lines, InLen{chrom_len}, OutOffsets32{chrom_offsets, d, n_lines});
        if (st == GIQLB_OK && want_name) st = launch_scan<OpAddU64>(ctx, n_lines, InLen{name_len}, OutOffsets32{name_offsets, d + 1, n_lines});
        if (st != GIQLB_OK) return cleanup(st);
    }
    st = fetch_scalars_checked(ctx, d, 3);
    if (st != GIQLB_OK) return cleanup(st);
    const int64_t chrom_total = ctx->h_scalars[0], name_total = ctx->h_scalars[1];
    const unsigned long long bad = (unsigned long long)ctx->h_scalars[2];
    *chrom_total_host = chrom_total;
    *name_total_host = name_total;
    if (bad != ~0ull)
        return cleanup(giqlb_fail(ctx, GIQLB_ERR_INVALID, "bed_parse: line %llu is blank or malformed (expected %d TAB-separated fields, integer coordinates%s)",
                                  bad + 1, n_fields, n_fields >= 6 ? ", strand + - or ." : ""));
    if (chrom_total > chrom_capacity || (want_name && name_total > name_capacity))
        return cleanup(giqlb_fail(ctx, GIQLB_ERR_CAPACITY, "bed_parse: name bytes exceed the buffers (chrom %lld of %lld, name %lld of %lld)",
                                  (long long)chrom_total, (long long)chrom_capacity, (long long)name_total, (long long)name_capacity));
    if (chrom_total >= (1ll << 31) || name_total >= (1ll << 31)) return cleanup(giqlb_fail(ctx, GIQLB_ERR_RANGE, "bed_parse: name bytes exceed int32 offsets"));
    {
        PhaseSpan ps(ctx, GIQLB_PH_OTHER);
        const unsigned grid = (unsigned)ceil_div(n_lines, 256);
        if (chrom_total > 0) {
            if (!chrom_bytes) return cleanup(giqlb_fail(ctx, GIQLB_ERR_INVALID, "bed_parse: null chrom_bytes"));
            bed_copy_kernel<<<grid, 256, 0, ctx->stream>>>(text, line_start, chrom_offsets, n_lines, chrom_bytes);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
        }
        if (want_name && name_total > 0) {
            if (!name_bytes) return cleanup(giqlb_fail(ctx, GIQLB_ERR_INVALID, "bed_parse: null name_bytes"));
            bed_copy_kernel<<<grid, 256, 0, ctx->stream>>>(text, name_pos, name_offsets, n_lines, name_bytes);
            ctx->launches_this_call++;
            ctx->stats.kernels_total++;
        }
        if (cudaGetLastError() != cudaSuccess) return cleanup(giqlb_fail(ctx, GIQLB_ERR_CUDA, "bed_parse: launch failed"));
    }
    ctx->stats.rows_in[0] = n_bytes;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = n_lines;
    ctx->stats.algorithmic_bytes = 3 * n_bytes + 12 * n_lines + chrom_total + name_total;
    return end_call(ctx, cleanup(GIQLB_OK));
}
