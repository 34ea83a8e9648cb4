/*
 * CPU restatement of GIQL's interval-join / cluster semantics in plain C + OpenMP.
 * TEST INFRASTRUCTURE ONLY: linked by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  Never linked or called by giql_b200/.
 *
 * The reference (abdenlab/giql) emits SQL; DuckDB (duckdb>=1.4.0, not vendored, not
 * installed here) executes it.  This file restates what that SQL computes:
 *
 *   INTERSECTS   l.chrom = r.chrom AND l.start < r.end AND l.end > r.start
 *                (src/giql/expanders/intersects.py:149-154), evaluated per chromosome
 *                (src/giql/expanders/intersects_duckdb.py:1283-1299).
 *   SEMI / ANTI  [NOT] EXISTS of the same predicate (intersects_duckdb.py:1254-1282).
 *   CLUSTER      flag = !(running MAX(end) over preceding rows [+ d] >= start), id = SUM(flag)
 *                per chrom[,strand] partition ordered by start (cluster.py:242-310, 430-434).
 *   MERGE        GROUP BY cluster id -> MIN(start), MAX(end), COUNT(*) (merge.py:293-308).
 *
 * The algorithm is the CPU analogue of an IEJoin/sort-sweep: per chromosome sort the right
 * table by start, keep a running prefix-max of end, and for each left row binary-search
 * r.start < l.end then walk back while the prefix-max still exceeds l.start, testing the
 * literal predicate.  It is deliberately a different algorithm from the GPU engine's
 * two-sided rank decomposition so the two check each other.
 *
 * Parity status: validated against the pure-Python restatement and the reference's golden
 * vectors in tests/test_oracle.py; unpinned against DuckDB itself (absent from this image).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t n;
    int32_t n_codes;
    int64_t *seg;     /* n_codes + 1 segment offsets into the sorted arrays */
    int32_t *start;   /* sorted by (code, start) */
    int32_t *end;
    int32_t *pmax;    /* running max of end within the segment, inclusive */
    int32_t *rid;
} ridx_t;

typedef struct { int32_t start, end, rid; } row_t;

static int cmp_row(const void *a, const void *b) {
    const row_t *x = (const row_t *)a, *y = (const row_t *)b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    return x->rid < y->rid ? -1 : (x->rid > y->rid);
}

static void ridx_free(ridx_t *ix) {
    free(ix->seg); free(ix->start); free(ix->end); free(ix->pmax); free(ix->rid);
}

/* bucket by chrom code, sort each bucket by start, prefix-max the ends */
static int ridx_build(ridx_t *ix, int64_t n, int32_t n_codes, const int32_t *code,
                      const int32_t *start, const int32_t *end) {
    ix->n = n; ix->n_codes = n_codes;
    ix->seg = (int64_t *)calloc((size_t)n_codes + 2, sizeof(int64_t));
    ix->start = (int32_t *)malloc((size_t)(n ? n : 1) * 4);
    ix->end = (int32_t *)malloc((size_t)(n ? n : 1) * 4);
    ix->pmax = (int32_t *)malloc((size_t)(n ? n : 1) * 4);
    ix->rid = (int32_t *)malloc((size_t)(n ? n : 1) * 4);
    row_t *rows = (row_t *)malloc((size_t)(n ? n : 1) * sizeof(row_t));
    if (!ix->seg || !ix->start || !ix->end || !ix->pmax || !ix->rid || !rows) return -1;
    for (int64_t i = 0; i < n; i++) ix->seg[code[i] + 1]++;
    for (int32_t c = 0; c < n_codes; c++) ix->seg[c + 1] += ix->seg[c];
    int64_t *cur = (int64_t *)malloc((size_t)(n_codes + 1) * sizeof(int64_t));
    memcpy(cur, ix->seg, (size_t)(n_codes + 1) * sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        int64_t p = cur[code[i]]++;
        rows[p].start = start[i]; rows[p].end = end[i]; rows[p].rid = (int32_t)i;
    }
    free(cur);
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t c = 0; c < n_codes; c++) {
        int64_t lo = ix->seg[c], hi = ix->seg[c + 1];
        qsort(rows + lo, (size_t)(hi - lo), sizeof(row_t), cmp_row);
        int32_t m = INT32_MIN;
        for (int64_t j = lo; j < hi; j++) {
            ix->start[j] = rows[j].start; ix->end[j] = rows[j].end; ix->rid[j] = rows[j].rid;
            if (rows[j].end > m) m = rows[j].end;
            ix->pmax[j] = m;
        }
    }
    free(rows);
    return 0;
}

static inline int64_t lower_bound_i32(const int32_t *a, int64_t lo, int64_t hi, int32_t x) {
    while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

/* visit every right row matching left row (c, s, e); returns count; optionally writes rids */
static inline int64_t probe(const ridx_t *ix, int32_t c, int32_t s, int32_t e, int32_t *out_r, int first_only) {
    if (c < 0 || c >= ix->n_codes) return 0;
    int64_t lo = ix->seg[c], hi = ix->seg[c + 1];
    int64_t ub = lower_bound_i32(ix->start, lo, hi, e);      /* r.start < l.end */
    int64_t cnt = 0;
    for (int64_t j = ub - 1; j >= lo && ix->pmax[j] > s; j--) {
        if (ix->end[j] > s) {                                 /* l.start < r.end */
            if (out_r) out_r[cnt] = ix->rid[j];
            cnt++;
            if (first_only) break;
        }
    }
    return cnt;
}

/* counts[n_l] <- matches per left row; returns total, or -1 on allocation failure */
int64_t giql_oracle_intersects_count(int64_t n_l, const int32_t *lc, const int32_t *ls, const int32_t *le,
                                     int64_t n_r, const int32_t *rc, const int32_t *rs, const int32_t *re,
                                     int32_t n_codes, int64_t *counts, int nthreads) {
    ridx_t ix;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (ridx_build(&ix, n_r, n_codes, rc, rs, re)) return -1;
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < n_l; i++) {
        int64_t c = probe(&ix, lc[i], ls[i], le[i], NULL, 0);
        counts[i] = c; total += c;
    }
    ridx_free(&ix);
    return total;
}

/* offsets = exclusive scan of counts; writes (left_row, right_row) pairs */
int64_t giql_oracle_intersects_emit(int64_t n_l, const int32_t *lc, const int32_t *ls, const int32_t *le,
                                    int64_t n_r, const int32_t *rc, const int32_t *rs, const int32_t *re,
                                    int32_t n_codes, const int64_t *offsets, int32_t *out_l, int32_t *out_r,
                                    int nthreads) {
    ridx_t ix;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (ridx_build(&ix, n_r, n_codes, rc, rs, re)) return -1;
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < n_l; i++) {
        int64_t c = probe(&ix, lc[i], ls[i], le[i], out_r + offsets[i], 0);
        for (int64_t k = 0; k < c; k++) out_l[offsets[i] + k] = (int32_t)i;
        total += c;
    }
    ridx_free(&ix);
    return total;
}

/* whole join in one call (count + scan + emit): the timed CPU baseline.  Caller passes
 * capacity; returns total pairs (may exceed capacity, in which case nothing is written). */
int64_t giql_oracle_intersects_join(int64_t n_l, const int32_t *lc, const int32_t *ls, const int32_t *le,
                                    int64_t n_r, const int32_t *rc, const int32_t *rs, const int32_t *re,
                                    int32_t n_codes, int64_t capacity, int32_t *out_l, int32_t *out_r,
                                    int nthreads) {
    ridx_t ix;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (ridx_build(&ix, n_r, n_codes, rc, rs, re)) return -1;
    int64_t *offs = (int64_t *)malloc((size_t)(n_l + 1) * sizeof(int64_t));
    if (!offs) { ridx_free(&ix); return -1; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_l; i++) offs[i + 1] = probe(&ix, lc[i], ls[i], le[i], NULL, 0);
    offs[0] = 0;
    for (int64_t i = 0; i < n_l; i++) offs[i + 1] += offs[i];
    int64_t total = offs[n_l];
    if (total <= capacity) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_l; i++) {
            int64_t c = probe(&ix, lc[i], ls[i], le[i], out_r + offs[i], 0);
            for (int64_t k = 0; k < c; k++) out_l[offs[i] + k] = (int32_t)i;
        }
    }
    free(offs);
    ridx_free(&ix);
    return total;
}

/* exists[n_l] <- 1 when the left row has at least one match (SEMI); ANTI is the complement */
int64_t giql_oracle_intersects_exists(int64_t n_l, const int32_t *lc, const int32_t *ls, const int32_t *le,
                                      int64_t n_r, const int32_t *rc, const int32_t *rs, const int32_t *re,
                                      int32_t n_codes, uint8_t *exists, int nthreads) {
    ridx_t ix;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (ridx_build(&ix, n_r, n_codes, rc, rs, re)) return -1;
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < n_l; i++) {
        int64_t c = probe(&ix, lc[i], ls[i], le[i], NULL, 1);
        exists[i] = c > 0; total += c > 0;
    }
    ridx_free(&ix);
    return total;
}

/* CLUSTER: part[i] = partition code (chrom, or chrom*S+strand); ids aligned to input rows.
 * Returns number of clusters summed over partitions. */
int64_t giql_oracle_cluster(int64_t n, const int32_t *part, const int32_t *start, const int32_t *end,
                            int32_t n_parts, int64_t distance, int32_t *cluster_id, int nthreads) {
    ridx_t ix;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (ridx_build(&ix, n, n_parts, part, start, end)) return -1;
    int64_t d = distance > 0 ? distance : 0, total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int32_t c = 0; c < n_parts; c++) {
        int64_t lo = ix.seg[c], hi = ix.seg[c + 1];
        int32_t id = 0;
        for (int64_t j = lo; j < hi; j++) {
            int is_new = (j == lo) || !((int64_t)ix.pmax[j - 1] + d >= (int64_t)ix.start[j]);
            id += is_new;
            cluster_id[ix.rid[j]] = id;
        }
        total += id;
    }
    ridx_free(&ix);
    return total;
}

/* MERGE: emits (part, min start, max end, count) per cluster ordered by (part, start).
 * out_* must hold n rows; returns number of merged rows. */
int64_t giql_oracle_merge(int64_t n, const int32_t *part, const int32_t *start, const int32_t *end,
                          int32_t n_parts, int64_t distance, int32_t *out_part, int32_t *out_start,
                          int32_t *out_end, int64_t *out_count, int nthreads) {
    ridx_t ix;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (ridx_build(&ix, n, n_parts, part, start, end)) return -1;
    int64_t d = distance > 0 ? distance : 0, m = 0;
    for (int32_t c = 0; c < n_parts; c++) {
        int64_t lo = ix.seg[c], hi = ix.seg[c + 1];
        for (int64_t j = lo; j < hi; j++) {
            int is_new = (j == lo) || !((int64_t)ix.pmax[j - 1] + d >= (int64_t)ix.start[j]);
            if (is_new) {
                out_part[m] = c; out_start[m] = ix.start[j]; out_end[m] = ix.end[j]; out_count[m] = 1; m++;
            } else {
                if (ix.end[j] > out_end[m - 1]) out_end[m - 1] = ix.end[j];
                out_count[m - 1]++;
            }
        }
    }
    ridx_free(&ix);
    return m;
}

int giql_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
