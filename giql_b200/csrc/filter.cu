// filter.cu — residual predicates over join pairs.
//
// The reference's IEJoin override keeps `a.interval INTERSECTS b.interval AND <residual>` as an extra conjunct of every
// per-chromosome ON clause (abdenlab/giql src/giql/expanders/intersects_duckdb.py:1239-1243), and bedtools' -s / -S
// (same / opposite strand) are written that way in GIQL: `AND a.strand = b.strand` / `AND a.strand != b.strand`
// (docs/recipes/bedtools-migration.rst:271-307).  Here the residual is a pass over the pair list the join produced:
// pair i survives iff  left_col[left_rows[i]]  <op>  right_col[right_rows[i]]  with both operands non-NULL (SQL
// three-valued logic: a comparison with NULL is not TRUE).  One single-pass stream compaction (scan.cuh): read 8 B per
// pair + two gathered operands, write 8 B per surviving pair; order of the survivors is the order of the input.
// Floating-point operands follow DuckDB's total order (NaN equals NaN and is greater than every number).
#include "common.cuh"
#include "scan.cuh"

namespace {

template <typename T>
struct Cmp {
    __device__ static int order(T a, T b) { return a < b ? -1 : (a > b ? 1 : 0); }
};
template <>
struct Cmp<float> {
    __device__ static int order(float a, float b) {
        const bool na = a != a, nb = b != b;
        if (na || nb) return na == nb ? 0 : (na ? 1 : -1);
        return a < b ? -1 : (a > b ? 1 : 0);
    }
};
template <>
struct Cmp<double> {
    __device__ static int order(double a, double b) {
        const bool na = a != a, nb = b != b;
        if (na || nb) return na == nb ? 0 : (na ? 1 : -1);
        return a < b ? -1 : (a > b ? 1 : 0);
    }
};

__device__ __forceinline__ bool holds(int c, int op) {
    switch (op) {
        case GIQLB_CMP_EQ: return c == 0;
        case GIQLB_CMP_NE: return c != 0;
        case GIQLB_CMP_LT: return c < 0;
        case GIQLB_CMP_LE: return c <= 0;
        case GIQLB_CMP_GT: return c > 0;
        default: return c >= 0;
    }
}

template <typename T, bool STRAND>
struct InKeep {
    const int32_t* lrow;
    const int32_t* rrow;
    const T* lcol;
    const T* rcol;
    const uint8_t* lvalid;
    const uint8_t* rvalid;
    int op;
    __device__ unsigned long long operator()(int64_t i) const {
        const int32_t l = __ldg(lrow + i), r = __ldg(rrow + i);
        if (lvalid && !__ldg(lvalid + l)) return 0ull;
        if (rvalid && !__ldg(rvalid + r)) return 0ull;
        const T a = __ldg(lcol + l), b = __ldg(rcol + r);
        if (STRAND && (a == (T)GIQLB_STRAND_NULL || b == (T)GIQLB_STRAND_NULL)) return 0ull;
        return holds(Cmp<T>::order(a, b), op) ? 1ull : 0ull;
    }
};

struct OutPairs {
    const int32_t* lrow;
    const int32_t* rrow;
    int32_t* out_l;
    int32_t* out_r;
    unsigned long long* total;
    int64_t n;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        if (v) {
            out_l[excl] = __ldg(lrow + i);
            out_r[excl] = __ldg(rrow + i);
        }
        if (i == n - 1) *total = excl + v;
    }
};

template <typename T, bool STRAND>
int run(giqlb_ctx* ctx, const int32_t* lrow, const int32_t* rrow, int64_t n, const void* lcol, const void* rcol, const uint8_t* lvalid,
        const uint8_t* rvalid, int op, int32_t* out_l, int32_t* out_r, unsigned long long* total) {
    using namespace giqlb_scan;
    return launch_scan<OpAddU64>(ctx, n, InKeep<T, STRAND>{lrow, rrow, (const T*)lcol, (const T*)rcol, lvalid, rvalid, op},
                                 OutPairs{lrow, rrow, out_l, out_r, total, n});
}

}  // namespace

extern "C" int giqlb_pairs_filter(giqlb_ctx* ctx, const int32_t* left_rows, const int32_t* right_rows, int64_t n_pairs, const void* left_col,
                                  const void* right_col, int dtype, const uint8_t* left_valid, const uint8_t* right_valid, int op,
                                  int32_t* out_left, int32_t* out_right, int64_t* n_out_host) {
    if (!ctx || n_pairs < 0 || !n_out_host || op < GIQLB_CMP_EQ || op > GIQLB_CMP_GE || dtype < GIQLB_T_STRAND || dtype > GIQLB_T_F64)
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "pairs_filter: bad argument");
    if (n_pairs > 0 && (!left_rows || !right_rows || !left_col || !right_col || !out_left || !out_right))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "pairs_filter: null pointer");
    if (n_pairs > 0 && (out_left == left_rows || out_right == right_rows || out_left == right_rows || out_right == left_rows))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "pairs_filter: the outputs must not alias the inputs");
    begin_call(ctx);
    *n_out_host = 0;
    if (n_pairs == 0) return end_call(ctx);
    unsigned long long* total = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 8);
    CUDA_TRY(ctx, cudaMemsetAsync(total, 0, 8, ctx->stream));
    int st;
    {
        PhaseSpan ps(ctx, GIQLB_PH_SELECT);
        switch (dtype) {
            case GIQLB_T_STRAND: st = run<uint8_t, true>(ctx, left_rows, right_rows, n_pairs, left_col, right_col, left_valid, right_valid, op, out_left, out_right, total); break;
            case GIQLB_T_U8: st = run<uint8_t, false>(ctx, left_rows, right_rows, n_pairs, left_col, right_col, left_valid, right_valid, op, out_left, out_right, total); break;
            case GIQLB_T_I32: st = run<int32_t, false>(ctx, left_rows, right_rows, n_pairs, left_col, right_col, left_valid, right_valid, op, out_left, out_right, total); break;
            case GIQLB_T_I64: st = run<long long, false>(ctx, left_rows, right_rows, n_pairs, left_col, right_col, left_valid, right_valid, op, out_left, out_right, total); break;
            case GIQLB_T_F32: st = run<float, false>(ctx, left_rows, right_rows, n_pairs, left_col, right_col, left_valid, right_valid, op, out_left, out_right, total); break;
            default: st = run<double, false>(ctx, left_rows, right_rows, n_pairs, left_col, right_col, left_valid, right_valid, op, out_left, out_right, total); break;
        }
    }
    if (st != GIQLB_OK) return st;
    GIQLB_TRY(fetch_scalars_checked(ctx, total, 1));
    *n_out_host = ctx->h_scalars[0];
    ctx->stats.rows_in[0] = n_pairs;
    ctx->stats.rows_in[1] = 0;
    ctx->stats.rows_out = *n_out_host;
    ctx->stats.algorithmic_bytes = 8 * n_pairs + 8 * *n_out_host;
    return end_call(ctx);
}
