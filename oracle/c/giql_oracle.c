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
 *   CONTAINS / WITHIN joins, NEAREST (LATERAL form) and DISJOIN: see the section comments further down.
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

/* ---------------------------------------------------------------------------------------------------------
 * CONTAINS / WITHIN column joins  (src/giql/expanders/intersects.py:155-166)
 *   contains:  l.start <= r.start AND l.end >= r.end   (same chrom);   within = contains with the roles swapped.
 * Requires end >= start on the right table (the wrapper checks): the candidates of a left row are then the right rows
 * whose start lies in [l.start, l.end], scanned with the end test.  counts -> scan -> emit like the overlap join.
 * --------------------------------------------------------------------------------------------------------- */
static inline int64_t probe_contains(const ridx_t *ix, int32_t c, int32_t s, int32_t e, int32_t *out_r) {
    if (c < 0 || c >= ix->n_codes) return 0;
    int64_t lo = ix->seg[c], hi = ix->seg[c + 1], cnt = 0;
    for (int64_t j = lower_bound_i32(ix->start, lo, hi, s); j < hi && ix->start[j] <= e; j++) {
        if (ix->end[j] <= e) {
            if (out_r) out_r[cnt] = ix->rid[j];
            cnt++;
        }
    }
    return cnt;
}

int64_t giql_oracle_contains_join(int64_t n_l, const int32_t *lc, const int32_t *ls, const int32_t *le,
                                  int64_t n_r, const int32_t *rc, const int32_t *rs, const int32_t *re,
                                  int32_t n_codes, int64_t capacity, int32_t *out_l, int32_t *out_r, int nthreads) {
    ridx_t ix;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (ridx_build(&ix, n_r, n_codes, rc, rs, re)) return -1;
    int64_t *offs = (int64_t *)malloc((size_t)(n_l + 1) * sizeof(int64_t));
    if (!offs) { ridx_free(&ix); return -1; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_l; i++) offs[i + 1] = probe_contains(&ix, lc[i], ls[i], le[i], NULL);
    offs[0] = 0;
    for (int64_t i = 0; i < n_l; i++) offs[i + 1] += offs[i];
    int64_t total = offs[n_l];
    if (total <= capacity) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_l; i++) {
            int64_t c = probe_contains(&ix, lc[i], ls[i], le[i], out_r + offs[i]);
            for (int64_t k = 0; k < c; k++) out_l[offs[i] + k] = (int32_t)i;
        }
    }
    free(offs);
    ridx_free(&ix);
    return total;
}

/* ---------------------------------------------------------------------------------------------------------
 * NEAREST, LATERAL form  (src/giql/expanders/nearest.py:255-397; distance CASE src/giql/expanders/_distance.py:67-117)
 *   per query row: targets of the same chrom (and, when stranded, the same non-NULL strand literal), optionally
 *   ABS(distance) <= max_distance, ORDER BY ABS(distance) NULLS FIRST, start, end (ties here: target row id), LIMIT k.
 * Proper intervals only (start < end on both sides; the wrapper checks): a target then either overlaps the query
 * (distance 0), lies downstream (t.start >= q.end, distance t.start - q.end + 1) or upstream (t.end <= q.start).
 * Algorithm (not the GPU's three-stream merge): per partition keep the targets sorted by (start, end, row) and a second
 * view sorted by (end, -start, -row); candidates of a query = every overlapping row (prefix-max walk) + the first k
 * downstream rows + the last k upstream rows; they are keyed, sorted and cut to k.
 * strand codes (giql_b200's): 0 '+', 1 '-', 2 '.', 3 '?', >= 4 other literals, 255 SQL NULL.  '.', '?' and NULL give a
 * NULL distance in the stranded CASE (_distance.py:93-95).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct { int32_t start, end, rid; } nrow_t;
static int cmp_nrow_start(const void *a, const void *b) {
    const nrow_t *x = (const nrow_t *)a, *y = (const nrow_t *)b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    return x->rid < y->rid ? -1 : (x->rid > y->rid);
}
static int cmp_nrow_end(const void *a, const void *b) {   /* end asc, start desc, row desc: walked backwards */
    const nrow_t *x = (const nrow_t *)a, *y = (const nrow_t *)b;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    if (x->start != y->start) return x->start > y->start ? -1 : 1;
    return x->rid > y->rid ? -1 : (x->rid < y->rid);
}
typedef struct { int64_t absd; int32_t start, end, rid; int64_t d; int null_d; } ncand_t;
static int cmp_ncand(const void *a, const void *b) {
    const ncand_t *x = (const ncand_t *)a, *y = (const ncand_t *)b;
    if (x->null_d != y->null_d) return x->null_d ? -1 : 1;          /* NULLS FIRST */
    if (x->absd != y->absd) return x->absd < y->absd ? -1 : 1;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    return x->rid < y->rid ? -1 : (x->rid > y->rid);
}

/* out_*: capacity k * n_q rows; returns rows written (query-major, key order inside a query), or -1 */
int64_t giql_oracle_nearest(int64_t n_q, const int32_t *qc, const int32_t *qs, const int32_t *qe, const uint8_t *qstrand,
                            int64_t n_t, const int32_t *tc, const int32_t *ts, const int32_t *te, const uint8_t *tstrand,
                            int32_t n_codes, int32_t k, int64_t max_distance /* < 0: none */, int stranded, int is_signed,
                            int32_t *out_q, int32_t *out_t, int64_t *out_d, uint8_t *out_valid, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const int use_strand = stranded && qstrand && tstrand;
    const int32_t n_parts = use_strand ? n_codes * 256 : n_codes;
    int64_t *seg = (int64_t *)calloc((size_t)n_parts + 2, sizeof(int64_t));
    nrow_t *by_start = (nrow_t *)malloc((size_t)(n_t ? n_t : 1) * sizeof(nrow_t));
    nrow_t *by_end = (nrow_t *)malloc((size_t)(n_t ? n_t : 1) * sizeof(nrow_t));
    int32_t *pmax = (int32_t *)malloc((size_t)(n_t ? n_t : 1) * 4);
    int64_t *per_q = (int64_t *)calloc((size_t)n_q + 1, sizeof(int64_t));
    if (!seg || !by_start || !by_end || !pmax || !per_q) return -1;
#define NPART(c, st) (use_strand ? (c) * 256 + (int32_t)(st) : (c))
    for (int64_t i = 0; i < n_t; i++) seg[NPART(tc[i], use_strand ? tstrand[i] : 0) + 1]++;
    for (int32_t p = 0; p < n_parts; p++) seg[p + 1] += seg[p];
    {
        int64_t *cur = (int64_t *)malloc((size_t)(n_parts + 1) * sizeof(int64_t));
        memcpy(cur, seg, (size_t)(n_parts + 1) * sizeof(int64_t));
        for (int64_t i = 0; i < n_t; i++) {
            int64_t p = cur[NPART(tc[i], use_strand ? tstrand[i] : 0)]++;
            by_start[p].start = ts[i]; by_start[p].end = te[i]; by_start[p].rid = (int32_t)i;
        }
        free(cur);
    }
    memcpy(by_end, by_start, (size_t)n_t * sizeof(nrow_t));
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t p = 0; p < n_parts; p++) {
        int64_t lo = seg[p], hi = seg[p + 1];
        qsort(by_start + lo, (size_t)(hi - lo), sizeof(nrow_t), cmp_nrow_start);
        qsort(by_end + lo, (size_t)(hi - lo), sizeof(nrow_t), cmp_nrow_end);
        int32_t m = INT32_MIN;
        for (int64_t j = lo; j < hi; j++) { if (by_start[j].end > m) m = by_start[j].end; pmax[j] = m; }
    }
    /* every query writes into its own k-slot window; compacted afterwards */
#pragma omp parallel
    {
        int64_t cap = 64;
        ncand_t *cand = (ncand_t *)malloc((size_t)cap * sizeof(ncand_t));
#pragma omp for schedule(dynamic, 1024)
        for (int64_t qi = 0; qi < n_q; qi++) {
            const int32_t c = qc[qi], s = qs[qi], e = qe[qi];
            const uint8_t qst = use_strand ? qstrand[qi] : 0;
            if (c < 0 || c >= n_codes || (use_strand && qst == 255)) continue;        /* NULL = NULL is not TRUE */
            const int32_t p = NPART(c, qst);
            const int64_t lo = seg[p], hi = seg[p + 1];
            if (hi <= lo) continue;
            const int null_d = use_strand && (qst == 2 || qst == 3);                  /* '.', '?': the stranded CASE yields NULL */
            int64_t nc = 0;
#define NPUSH(row, dist)                                                                                   \
    do {                                                                                                   \
        if (nc == cap) { cap *= 2; cand = (ncand_t *)realloc(cand, (size_t)cap * sizeof(ncand_t)); }       \
        cand[nc].start = (row).start; cand[nc].end = (row).end; cand[nc].rid = (row).rid;                  \
        cand[nc].d = (dist); cand[nc].absd = (dist) < 0 ? -(dist) : (dist); cand[nc].null_d = null_d; nc++; \
    } while (0)
            if (null_d) {
                if (max_distance >= 0) continue;                                      /* NULL <= x is not TRUE */
                for (int64_t j = lo; j < hi && j < lo + k; j++) NPUSH(by_start[j], 0);
            } else {
                /* overlapping: t.start < q.end, walk back while the running max end still exceeds q.start */
                int64_t ub = lo, top = hi;
                while (ub < top) { int64_t mid = ub + ((top - ub) >> 1); if (by_start[mid].start < e) ub = mid + 1; else top = mid; }
                for (int64_t j = ub - 1; j >= lo && pmax[j] > s; j--)
                    if (by_start[j].end > s) NPUSH(by_start[j], 0);
                /* downstream: t.start >= q.end, nearest first */
                for (int64_t j = ub, taken = 0; j < hi && taken < k; j++, taken++) {
                    const int64_t gap = (int64_t)by_start[j].start - e + 1;
                    if (max_distance >= 0 && gap > max_distance) break;
                    NPUSH(by_start[j], (use_strand && qst == 1) ? -gap : gap);                        /* _distance.py:75, 98, 112 */
                }
                /* upstream: t.end <= q.start, largest end first */
                int64_t ue = lo; top = hi;
                while (ue < top) { int64_t mid = ue + ((top - ue) >> 1); if (by_end[mid].end <= s) ue = mid + 1; else top = mid; }
                for (int64_t j = ue - 1, taken = 0; j >= lo && taken < k; j--, taken++) {
                    const int64_t gap = (int64_t)s - by_end[j].end + 1;
                    if (max_distance >= 0 && gap > max_distance) break;
                    int64_t d;
                    if (!use_strand) d = is_signed ? -gap : gap;                                      /* _distance.py:76 / :83 */
                    else if (is_signed) d = qst == 1 ? gap : -gap;                                    /* :101-102 */
                    else d = qst == 1 ? -gap : gap;                                                   /* :115-116 */
                    NPUSH(by_end[j], d);
                }
            }
            qsort(cand, (size_t)nc, sizeof(ncand_t), cmp_ncand);
            const int64_t take = nc < k ? nc : k, base = qi * (int64_t)k;
            for (int64_t j = 0; j < take; j++) {
                out_q[base + j] = (int32_t)qi; out_t[base + j] = cand[j].rid; out_d[base + j] = cand[j].d; out_valid[base + j] = !null_d;
            }
            per_q[qi] = take;
        }
        free(cand);
    }
#undef NPUSH
#undef NPART
    int64_t m = 0;
    for (int64_t qi = 0; qi < n_q; qi++) {
        const int64_t base = qi * (int64_t)k;
        for (int64_t j = 0; j < per_q[qi]; j++, m++) {
            if (m != base + j) { out_q[m] = out_q[base + j]; out_t[m] = out_t[base + j]; out_d[m] = out_d[base + j]; out_valid[m] = out_valid[base + j]; }
        }
    }
    free(seg); free(by_start); free(by_end); free(pmax); free(per_q);
    return m;
}

/* ---------------------------------------------------------------------------------------------------------
 * DISJOIN  (src/giql/expanders/disjoin.py:102-202)
 *   breakpoints = distinct reference starts / ends per chrom (:149-154); a target (chrom, start, end) is cut at every
 *   breakpoint strictly inside it (:159-172); consecutive cuts form segments (:173-178); unless the reference is the
 *   target itself, a segment is kept when some reference row has r.start <= seg_start < r.end (:184-190).
 * Proper reference intervals only (start <= end; the wrapper checks): coverage of a point a is then
 *   #(r.start <= a) - #(r.end <= a) > 0   over the chromosome's sorted starts and sorted ends.
 * Two calls: counts per target (out_* NULL), then emit at the scanned offsets.  Returns the total.
 * --------------------------------------------------------------------------------------------------------- */
static int cmp_i32(const void *a, const void *b) {
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return x < y ? -1 : (x > y);
}
static inline int64_t upper_bound_i32(const int32_t *a, int64_t lo, int64_t hi, int32_t x) {
    while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (a[mid] <= x) lo = mid + 1; else hi = mid; }
    return lo;
}
int64_t giql_oracle_disjoin(int64_t n_t, const int32_t *tc, const int32_t *ts, const int32_t *te,
                            int64_t n_r, const int32_t *rc, const int32_t *rs, const int32_t *re, int32_t n_codes,
                            int self_reference, int64_t *counts /* [n_t], written */, const int64_t *offsets /* NULL: count only */,
                            int32_t *out_row, int32_t *out_start, int32_t *out_end, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    int64_t *seg = (int64_t *)calloc((size_t)n_codes + 2, sizeof(int64_t));
    int64_t *bseg = (int64_t *)calloc((size_t)n_codes + 2, sizeof(int64_t));
    int32_t *S = (int32_t *)malloc((size_t)(n_r ? n_r : 1) * 4), *E = (int32_t *)malloc((size_t)(n_r ? n_r : 1) * 4);
    int32_t *B = (int32_t *)malloc((size_t)(n_r ? 2 * n_r : 1) * 4);
    if (!seg || !bseg || !S || !E || !B) return -1;
    for (int64_t i = 0; i < n_r; i++) seg[rc[i] + 1]++;
    for (int32_t c = 0; c < n_codes; c++) seg[c + 1] += seg[c];
    {
        int64_t *cur = (int64_t *)malloc((size_t)(n_codes + 1) * sizeof(int64_t));
        memcpy(cur, seg, (size_t)(n_codes + 1) * sizeof(int64_t));
        for (int64_t i = 0; i < n_r; i++) { int64_t p = cur[rc[i]]++; S[p] = rs[i]; E[p] = re[i]; }
        free(cur);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t c = 0; c < n_codes; c++) {
        const int64_t lo = seg[c], hi = seg[c + 1], n = hi - lo;
        int32_t *b = B + 2 * lo;
        memcpy(b, S + lo, (size_t)n * 4);
        memcpy(b + n, E + lo, (size_t)n * 4);
        qsort(b, (size_t)(2 * n), 4, cmp_i32);
        int64_t u = 0;
        for (int64_t j = 0; j < 2 * n; j++) if (j == 0 || b[j] != b[j - 1]) b[u++] = b[j];
        bseg[c + 1] = u;                                     /* per-chrom count, offsets are 2 * seg[c] */
        qsort(S + lo, (size_t)n, 4, cmp_i32);
        qsort(E + lo, (size_t)n, 4, cmp_i32);
    }
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < n_t; i++) {
        const int32_t c = tc[i], s = ts[i], e = te[i];
        int64_t cnt = 0;
        if (c >= 0 && c < n_codes && s != e) {
            const int64_t lo = seg[c], hi = seg[c + 1];
            const int32_t *b = B + 2 * lo;
            const int64_t nb = bseg[c + 1];
            const int32_t first = s < e ? s : e, last = s < e ? e : s;
            /* interior breakpoints only exist for a proper target (start < bp < end) */
            int64_t j0 = 0, j1 = 0;
            if (s < e) { j0 = upper_bound_i32(b, 0, nb, s); j1 = lower_bound_i32(b, 0, nb, e); if (j1 < j0) j1 = j0; }
            int32_t a = first;
            for (int64_t j = j0; j <= j1; j++) {
                const int32_t nxt = j < j1 ? b[j] : last;
                int keep = nxt > a;
                if (keep && !self_reference)
                    keep = upper_bound_i32(S, lo, hi, a) - lo > upper_bound_i32(E, lo, hi, a) - lo;
                if (keep) {
                    if (offsets) { const int64_t o = offsets[i] + cnt; out_row[o] = (int32_t)i; out_start[o] = a; out_end[o] = nxt; }
                    cnt++;
                }
                a = nxt;
            }
        }
        counts[i] = cnt;
        total += cnt;
    }
    free(seg); free(bseg); free(S); free(E); free(B);
    return total;
}

int giql_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
