// distance.cu — DISTANCE(a, b[, stranded, signed]) over already-paired rows.
//
// Follows the CASE expression the reference emits (abdenlab/giql src/giql/expanders/_distance.py:67-117,
// identical to src/giql/expanders/distance.py:126-239):
//   NULL when the chromosomes differ; 0 when a.start < b.end AND a.end > b.start; otherwise the
//   bedtools `closest -d` gap (+1): downstream b.start - a.end + 1 when a.end <= b.start, else upstream
//   a.start - b.end + 1.  signed: upstream is negated.  stranded (both strand columns given): NULL when
//   either strand is NULL/'.'/'?'; downstream is negated on an a.strand of '-'; upstream is negated on '-'
//   (unsigned) or made positive on '-' and negative otherwise (signed).
// The same device function is reused by nearest.cu.
#include "distance.cuh"

namespace {

// a_rows / b_rows (nullable): pair i is (a row a_rows[i], b row b_rows[i]) — e.g. the two columns of a join result — instead of
// (row i, row i); the columns are gathered here, so no gathered copies of six columns are made first.  Row ids outside
// [0, a_n) / [0, b_n) raise err[1] = 4 and yield NULL.
__global__ void __launch_bounds__(256)
distance_kernel(const int32_t* __restrict__ a_code, const int32_t* __restrict__ a_start, const int32_t* __restrict__ a_end,
                const uint8_t* __restrict__ a_strand, const int32_t* __restrict__ b_code, const int32_t* __restrict__ b_start,
                const int32_t* __restrict__ b_end, const uint8_t* __restrict__ b_strand, const int32_t* __restrict__ a_rows,
                const int32_t* __restrict__ b_rows, int64_t a_n, int64_t b_n, int64_t n, int stranded, int signed_,
                int32_t a_sadj, int32_t a_eadj, int32_t b_sadj, int32_t b_eadj, int32_t* __restrict__ out,
                uint8_t* __restrict__ valid, int* __restrict__ err) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ia = a_rows ? (int64_t)a_rows[i] : i, ib = b_rows ? (int64_t)b_rows[i] : i;
        long long d = 0;
        bool ok = ia >= 0 && ia < a_n && ib >= 0 && ib < b_n;
        if (!ok) err[1] = 4;
        ok = ok && a_code[ia] == b_code[ib] && a_code[ia] >= 0;
        if (ok) {
            uint8_t sa = stranded ? a_strand[ia] : (uint8_t)GIQLB_STRAND_PLUS;
            uint8_t sb = stranded ? b_strand[ib] : (uint8_t)GIQLB_STRAND_PLUS;
            ok = giqlb_distance_dev((long long)a_start[ia] + a_sadj, (long long)a_end[ia] + a_eadj, sa,
                                    (long long)b_start[ib] + b_sadj, (long long)b_end[ib] + b_eadj, sb, stranded != 0,
                                    signed_ != 0, &d);
        }
        out[i] = ok ? (int32_t)d : 0;
        valid[i] = ok ? 1 : 0;
    }
}

}  // namespace

static int distance_impl(giqlb_ctx* ctx, const int32_t* a_code, const int32_t* a_start, const int32_t* a_end, const uint8_t* a_strand,
                         int64_t a_n, const int32_t* b_code, const int32_t* b_start, const int32_t* b_end, const uint8_t* b_strand,
                         int64_t b_n, const int32_t* a_rows, const int32_t* b_rows, int64_t n, int stranded, int signed_,
                         int32_t a_start_adj, int32_t a_end_adj, int32_t b_start_adj, int32_t b_end_adj, int32_t* out, uint8_t* valid) {
    if (!ctx || n < 0 || (n > 0 && (!a_code || !a_start || !a_end || !b_code || !b_start || !b_end || !out || !valid)))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "distance: bad argument");
    if (stranded && n > 0 && (!a_strand || !b_strand))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "distance: stranded requires both strand columns");
    begin_call(ctx);
    const bool by_rows = a_rows || b_rows;
    if (n > 0) {
        int64_t blocks = ceil_div(n, 256 * 4);
        int64_t cap = (int64_t)ctx->sm_count * 16;
        PhaseSpan ps(ctx, GIQLB_PH_DISTANCE);
        distance_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
            a_code, a_start, a_end, a_strand, b_code, b_start, b_end, b_strand, a_rows, b_rows, a_n, b_n, n, stranded, signed_,
            a_start_adj, a_end_adj, b_start_adj, b_end_adj, out, valid, ctx->d_err);
        LAUNCH_CHECK(ctx);
    }
    ctx->stats.rows_in[0] = n;
    ctx->stats.rows_in[1] = n;
    ctx->stats.rows_out = n;
    ctx->stats.algorithmic_bytes = n * (24 + (stranded ? 2 : 0) + 5 + (a_rows ? 4 : 0) + (b_rows ? 4 : 0));
    if (by_rows && n > 0) {   // the only way to learn about a row id outside its table
        int st = fetch_scalars_checked(ctx, ctx->d_scalars, 1);
        if (st != GIQLB_OK) return st;
    }
    return end_call(ctx);
}

extern "C" int giqlb_distance(giqlb_ctx* ctx, const int32_t* a_code, const int32_t* a_start, const int32_t* a_end,
                              const uint8_t* a_strand, const int32_t* b_code, const int32_t* b_start, const int32_t* b_end,
                              const uint8_t* b_strand, int64_t n, int stranded, int signed_, int32_t a_start_adj,
                              int32_t a_end_adj, int32_t b_start_adj, int32_t b_end_adj, int32_t* out, uint8_t* valid) {
    return distance_impl(ctx, a_code, a_start, a_end, a_strand, n, b_code, b_start, b_end, b_strand, n, nullptr, nullptr, n, stranded, signed_,
                         a_start_adj, a_end_adj, b_start_adj, b_end_adj, out, valid);
}

extern "C" int giqlb_distance_rows(giqlb_ctx* ctx, const int32_t* a_code, const int32_t* a_start, const int32_t* a_end,
                                   const uint8_t* a_strand, int64_t a_n, const int32_t* b_code, const int32_t* b_start,
                                   const int32_t* b_end, const uint8_t* b_strand, int64_t b_n, const int32_t* a_rows,
                                   const int32_t* b_rows, int64_t n, int stranded, int signed_, int32_t a_start_adj,
                                   int32_t a_end_adj, int32_t b_start_adj, int32_t b_end_adj, int32_t* out, uint8_t* valid) {
    if (a_n < 0 || b_n < 0 || (!a_rows && n > a_n) || (!b_rows && n > b_n)) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "distance_rows: bad argument");
    return distance_impl(ctx, a_code, a_start, a_end, a_strand, a_n, b_code, b_start, b_end, b_strand, b_n, a_rows, b_rows, n, stranded, signed_,
                         a_start_adj, a_end_adj, b_start_adj, b_end_adj, out, valid);
}
