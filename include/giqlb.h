/*
 * giqlb.h — C ABI of libgiqlb200.so, the B200 (sm_100a) execution engine for GIQL's interval
 * operators.  This is the drop-in boundary: the Python host layer (giql_b200/) binds exactly these
 * symbols through ctypes, and a maintainer of the reference would bind the same symbols from an
 * expander registered with @giql.register(B200Target, <Operator>).
 *
 * What each entry point replaces in the reference (abdenlab/giql; paths relative to its root).  The
 * reference has no FFI of its own — it hands a SQL string to DuckDB — so "replaces" means "computes
 * what DuckDB computes for the SQL emitted at":
 *
 *   giqlb_dict_encode          chrom VARCHAR equality / per-chrom partitioning
 *                              src/giql/expanders/_per_chrom.py:27,46-74; intersects.py:151
 *   giqlb_layout_* / giqlb_index_build
 *                              the per-chromosome sort DuckDB's IE_JOIN performs for
 *                              src/giql/expanders/intersects_duckdb.py:1283-1299; canonical ±1 offsets of
 *                              src/giql/canonical.py:16-52 are applied here as integers
 *   giqlb_intersects_plan/emit/join
 *                              INNER branch  intersects_duckdb.py:1217-1238,1283-1299,1317-1334
 *                              (= naive predicate src/giql/expanders/intersects.py:149-154)
 *   giqlb_intersects_exists    SEMI / ANTI branch intersects_duckdb.py:1254-1282, partition rule :1317-1330
 *   giqlb_intersects_count_rows  count_overlaps fast path intersects_duckdb.py:432-548, 806-854
 *   giqlb_distance             src/giql/expanders/_distance.py:22-117 (== distance.py:126-239)
 *   giqlb_nearest              src/giql/expanders/nearest.py:255-397 (LATERAL form)
 *   giqlb_cluster              src/giql/expanders/cluster.py:205-310, 430-434
 *   giqlb_merge                src/giql/expanders/merge.py:183-439
 *   giqlb_disjoin_*            src/giql/expanders/disjoin.py:102-202
 *
 * Conventions
 *   - Every function returns 0 (GIQLB_OK) or a negative giqlb_status.  giqlb_last_error(ctx) returns a
 *     library-owned, context-local message.  The library never aborts and never frees caller memory.
 *   - Pointers are DEVICE pointers owned by the caller (PyTorch tensors) unless the parameter name ends
 *     in _host.  Work is enqueued on the context's stream; functions that return a host scalar
 *     synchronise that stream, all others are asynchronous.
 *   - One giqlb_ctx per (thread, device).  No global state; re-entrant across contexts.
 *   - Row ids are int32 positions in the caller's input arrays (n < 2^31 per table); pair counts and
 *     offsets are int64.  Coordinates are int32; canonical 0-based half-open after the *_adj offsets.
 *   - NULL chrom/start/end are rejected by the host layer before upload (BED never has them).
 */
#ifndef GIQLB_H
#define GIQLB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIQLB_ABI_VERSION 1

typedef enum giqlb_status {
    GIQLB_OK = 0,
    GIQLB_ERR_INVALID = -1,     /* bad argument */
    GIQLB_ERR_CUDA = -2,        /* CUDA runtime error, see giqlb_last_error */
    GIQLB_ERR_NOMEM = -3,
    GIQLB_ERR_CAPACITY = -4,    /* caller's output buffer too small; required size returned */
    GIQLB_ERR_RANGE = -5,       /* coordinates outside the layout / genome extent >= 2^32 */
    GIQLB_ERR_INTERNAL = -6
} giqlb_status;

typedef struct giqlb_ctx giqlb_ctx;
typedef struct giqlb_dict giqlb_dict;       /* shared chromosome dictionary (host strings, sorted bytewise) */
typedef struct giqlb_layout giqlb_layout;   /* genome linearisation shared by the tables of one query */
typedef struct giqlb_index giqlb_index;     /* one table sorted by linearised start, with carried row ids */
typedef struct giqlb_joinplan giqlb_joinplan;

/* strand codes used on the device (host layer maps VARCHAR -> code; SQL NULL never equals anything) */
#define GIQLB_STRAND_PLUS 0
#define GIQLB_STRAND_MINUS 1
#define GIQLB_STRAND_DOT 2       /* '.'  -> stranded distance NULL */
#define GIQLB_STRAND_QMARK 3     /* '?'  -> stranded distance NULL */
#define GIQLB_STRAND_OTHER0 4    /* 4..254: other literal values, equality-comparable */
#define GIQLB_STRAND_NULL 255

/* per-call statistics of the most recent operation on a context */
typedef struct giqlb_stats_t {
    int64_t rows_in[2];
    int64_t rows_out;
    int64_t algorithmic_bytes;   /* SURVEY.md §8(d) formula for the operation */
    int32_t kernels_launched;    /* kernels enqueued by the last API call */
    int32_t reserved;
    int64_t kernels_total;       /* kernels enqueued since giqlb_create */
} giqlb_stats_t;

/* phases of the per-kernel profile (giqlb_profile_read) */
enum {
    GIQLB_PH_DICT = 0,        /* dictionary encoding kernels */
    GIQLB_PH_LAYOUT = 1,      /* per-chromosome extent reduction + base scan */
    GIQLB_PH_KEYGEN = 2,      /* linearised key generation */
    GIQLB_PH_SORT_HIST = 3,   /* radix sort: digit histogram */
    GIQLB_PH_SORT_PASS = 4,   /* radix sort: one-sweep digit passes */
    GIQLB_PH_GATHER = 5,      /* gathers of end / strand into sorted order */
    GIQLB_PH_SCAN_MAX = 6,    /* running max of end */
    GIQLB_PH_JOIN_COUNT = 7,  /* join count pass */
    GIQLB_PH_JOIN_SCAN = 8,   /* exclusive scan of tile counts */
    GIQLB_PH_JOIN_EMIT = 9,   /* join emit pass */
    GIQLB_PH_PROBE = 10,      /* SEMI/ANTI existence and per-row count probes */
    GIQLB_PH_SELECT = 11,     /* flag compaction */
    GIQLB_PH_DISTANCE = 12,
    GIQLB_PH_CLUSTER = 13,
    GIQLB_PH_MERGE = 14,
    GIQLB_PH_NEAREST = 15,
    GIQLB_PH_OTHER = 16,
    GIQLB_PH_SHARD = 17,      /* genome-range partitioning (histogram, cuts, selection) */
    GIQLB_PH_DISJOIN = 18,
    GIQLB_PH_COUNT = 19
};

int giqlb_abi_version(void);
/* cuda_stream: the cudaStream_t every call of this context enqueues on; NULL = the legacy default stream
 * (PyTorch's default stream), so the caller's copies and kernels on that stream stay ordered with ours. */
int giqlb_create(int device, void* cuda_stream, giqlb_ctx** out);
int giqlb_destroy(giqlb_ctx* ctx);
const char* giqlb_last_error(const giqlb_ctx* ctx);
int giqlb_sync(giqlb_ctx* ctx);
/* Scratch and index memory is cached per context (freed blocks are reused without a driver call).
 * trim() synchronises and returns every parked block to the driver. */
int giqlb_trim(giqlb_ctx* ctx);
int giqlb_stats(const giqlb_ctx* ctx, giqlb_stats_t* out_host);
/* Per-phase device time.  While enabled, CUDA events are recorded on the context's stream around each
 * kernel group; read() synchronises, ADDS the elapsed milliseconds and launch counts since the last read
 * into ms_host[GIQLB_PH_COUNT] / launches_host[GIQLB_PH_COUNT] and clears the internal spans. */
int giqlb_profile_enable(giqlb_ctx* ctx, int on);
int giqlb_profile_read(giqlb_ctx* ctx, double* ms_host, int64_t* launches_host);

/* ---- chromosome dictionary --------------------------------------------------------------------
 * Arrow utf8 layout per table: int32 offsets[n+1], uint8 bytes[].  Builds ONE dictionary over both
 * tables (n_r may be 0), order-preserving w.r.t. bytewise string order (merge.py:429-437 sorts VARCHAR),
 * case-sensitive, any byte allowed (tests/test_duckdb_iejoin.py:3780-3810).  Writes int32 codes. */
int giqlb_dict_encode(giqlb_ctx* ctx,
                      const int32_t* offs_l, const uint8_t* bytes_l, int64_t n_l,
                      const int32_t* offs_r, const uint8_t* bytes_r, int64_t n_r,
                      int32_t* codes_l, int32_t* codes_r, giqlb_dict** out);
int giqlb_dict_size(const giqlb_dict* d);
int giqlb_dict_get(const giqlb_dict* d, int32_t code, const char** bytes_host, int32_t* len_host);
int giqlb_dict_free(giqlb_dict* d);

/* ---- genome layout ----------------------------------------------------------------------------
 * Linearises (chrom code, position) into one uint32 coordinate: g = base[code] + 1 + (pos - lo[code]),
 * with one guard cell either side of every chromosome (+ `pad` extra cells between chromosomes).
 * observe() folds a table's per-chrom extent in (device reduction); finalize() fixes the bases and
 * synchronises (returns GIQLB_ERR_RANGE when the summed extent does not fit 32 bits).  Every table that
 * will be INDEXED on the layout must be observed first; tables that only PROBE an index
 * (giqlb_intersects_exists / _count_rows / giqlb_nearest queries) need not be. */
int giqlb_layout_create(giqlb_ctx* ctx, int32_t n_codes, int64_t pad, giqlb_layout** out);
int giqlb_layout_observe(giqlb_ctx* ctx, giqlb_layout* lay, const int32_t* code, const int32_t* start,
                         const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj);
int giqlb_layout_finalize(giqlb_ctx* ctx, giqlb_layout* lay);
int giqlb_layout_free(giqlb_ctx* ctx, giqlb_layout* lay);
int giqlb_layout_info(const giqlb_layout* lay, int32_t* n_codes, uint64_t* total_extent, int32_t* key_bits);
/* device views of base[n_codes+1] (uint32), lo[n_codes], hi[n_codes] (int32; hi < lo for unseen codes) */
int giqlb_layout_tables(const giqlb_layout* lay, const uint32_t** base_dev, const int32_t** lo_dev,
                        const int32_t** hi_dev);

/* index_build is asynchronous.  Rows that lie outside the layout (a table that was not observed) raise a sticky
 * error reported as GIQLB_ERR_RANGE by the next call that returns a host value (plan/join, index_info, sync).
 * observe() and index_build() of one table must use the same offsets. */
/* ---- sorted index -----------------------------------------------------------------------------
 * LSD radix sort of the linearised start with carried row ids; then gathers the linearised end (tables of 24 M rows and
 * more carry the end through the sort as a second payload instead: a gather beyond the L2 costs a DRAM sector per row).
 * strand may be NULL.  start_adj/end_adj are the canonical.py:16-52 offsets (raw + adj = canonical).
 * flags: GIQLB_INDEX_TIE_BY_END orders equal starts by end (NEAREST's ORDER BY start, end). */
#define GIQLB_INDEX_TIE_BY_END 1
/* GIQLB_INDEX_NO_ROW_IDS: the caller will not ask which input row a sorted position came from (MERGE returns regions, not rows:
 * src/giql/expanders/merge.py:183-439).  Tables too large for the L2 are then sorted with the linear end as the only payload
 * (16 instead of 24 bytes per row and pass); giqlb_index_columns reports a NULL row-id column.  A hint: smaller tables and
 * strand-carrying indexes keep their row ids. */
#define GIQLB_INDEX_NO_ROW_IDS 2
int giqlb_index_build(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                      const int32_t* end, const uint8_t* strand, int64_t n, int32_t start_adj,
                      int32_t end_adj, int flags, giqlb_index** out);
/* Same, over a selection: `rows` (device int32[n], nullable) lists the rows of the table that are indexed
 * (a genome-range shard, giqlb_shard_select); the index carries rows[i] as row id. */
int giqlb_index_build_rows(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                           const int32_t* end, const uint8_t* strand, const int32_t* rows, int64_t n,
                           int32_t start_adj, int32_t end_adj, int flags, giqlb_index** out);
int giqlb_index_free(giqlb_ctx* ctx, giqlb_index* ix);
int64_t giqlb_index_rows(const giqlb_index* ix);
/* device views of the sorted columns (library-owned, valid until giqlb_index_free) */
int giqlb_index_columns(const giqlb_index* ix, const uint32_t** gstart, const uint32_t** gend,
                        const int32_t** row_id, const uint8_t** strand);
/* rows with end <= start / end < start; strand blocks after giqlb_index_group_by_strand */
int giqlb_index_info(giqlb_ctx* ctx, giqlb_index* ix, int64_t* n_degenerate, int64_t* n_negative, int32_t* n_blocks,
                     int64_t* block_off_host /* [257] or NULL */);
/* reorders the index into blocks of equal strand code, each still sorted by start (CLUSTER/MERGE/NEAREST
 * with stranded := true) */
int giqlb_index_group_by_strand(giqlb_ctx* ctx, giqlb_index* ix);

/* ---- INTERSECTS inner join --------------------------------------------------------------------
 * plan(): count pass; returns the exact number of pairs (host).  emit(): fills out_l/out_r (int32 row
 * ids, total entries).  join(): plan + emit without a host round trip when capacity suffices; returns
 * GIQLB_ERR_CAPACITY (and the required total) otherwise.  Pair order is unspecified (multiset).
 * The indexes are not const: derived arrays (running max of end) are cached on them on demand. */
int giqlb_intersects_plan(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, giqlb_joinplan** out,
                          int64_t* total_host);
int giqlb_intersects_emit(giqlb_ctx* ctx, giqlb_joinplan* plan, int32_t* out_l, int32_t* out_r);
int giqlb_joinplan_free(giqlb_ctx* ctx, giqlb_joinplan* plan);
int giqlb_intersects_join(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, int32_t* out_l, int32_t* out_r,
                          int64_t capacity, int64_t* total_host);

/* The whole path in one call: dictionary encode (skipped when both tables arrive with codes_valid = 1 and a shared
 * dictionary of n_codes names), layout, index of both tables, single-pass join into out_l/out_r.  Returns
 * GIQLB_ERR_CAPACITY and the exact total when `capacity` is too small.  On success with encoding done here,
 * *dict_out (nullable) receives the dictionary (caller frees) and L->codes / R->codes hold the codes. */
typedef struct giqlb_table_t {
    const int32_t* chrom_offsets;   /* Arrow utf8 offsets[n + 1] (device); unused when codes_valid */
    const uint8_t* chrom_bytes;
    int32_t* codes;                 /* int32[n] (device): written by the encode, or supplied with codes_valid = 1 */
    int32_t codes_valid;
    const int32_t* start;
    const int32_t* end;
    int64_t n;
    int32_t start_adj, end_adj;     /* canonical.py:16-52 offsets */
} giqlb_table_t;
int giqlb_intersects_tables(giqlb_ctx* ctx, giqlb_table_t* L, giqlb_table_t* R, int32_t n_codes, int32_t* out_l,
                            int32_t* out_r, int64_t capacity, int64_t* total_host, giqlb_dict** dict_out);

/* The same for rank `rank` of `world` genome-range shards: both tables are given in full; the cuts
 * are derived on the device from the left table, the rank's rows are selected and keyed in one pass per table, and
 * only the rank's pairs are produced (row ids refer to the full tables).  The caller all-gathers the totals. */
int giqlb_intersects_tables_sharded(giqlb_ctx* ctx, giqlb_table_t* L, giqlb_table_t* R, int32_t n_codes, int world, int rank,
                                    int32_t* out_l, int32_t* out_r, int64_t capacity, int64_t* total_host,
                                    giqlb_dict** dict_out);

/* ---- CONTAINS / WITHIN column-to-column joins (intersects.py:155-166) -----------------------------
 * op GIQLB_OP_CONTAINS: l.start <= r.start AND l.end >= r.end;  GIQLB_OP_WITHIN: the mirror.  Same chromosome, bag
 * semantics, pair order unspecified.  GIQLB_ERR_CAPACITY + the exact total when the outputs are too small. */
#define GIQLB_OP_CONTAINS 1
#define GIQLB_OP_WITHIN 2
int giqlb_spatial_join(giqlb_ctx* ctx, giqlb_index* L, giqlb_index* R, int op, int32_t* out_l, int32_t* out_r,
                       int64_t capacity, int64_t* total_host);

/* ---- BED text ingest (the reference's users load BED with DuckDB's read_csv) -----------------------
 * text: the file's bytes on the device (16-byte aligned, < 2 GiB, header lines already stripped), lines ending in \n or
 * \r\n, TAB-separated fields.  giqlb_bed_count returns the number of lines (a final unterminated line counts).
 * giqlb_bed_parse fills, for n_fields in 3..6: chromosome names in Arrow utf8 layout (chrom_offsets int32[n_lines + 1],
 * chrom_bytes), start / end (int32), and when requested (non-NULL, n_fields large enough) the name column (utf8),
 * score (int32; '.' reads as 0) and strand codes (GIQLB strand coding: '+' 0, '-' 1, '.' 2).  The byte totals are
 * returned either way; GIQLB_ERR_CAPACITY when a byte buffer is too small (call again with the totals);
 * GIQLB_ERR_INVALID naming the first blank or malformed line. */
int giqlb_bed_count(giqlb_ctx* ctx, const uint8_t* text, int64_t n_bytes, int64_t* n_lines_host);
int giqlb_bed_parse(giqlb_ctx* ctx, const uint8_t* text, int64_t n_bytes, int64_t n_lines, int n_fields,
                    int32_t* chrom_offsets, uint8_t* chrom_bytes, int64_t chrom_capacity, int32_t* start, int32_t* end,
                    int32_t* name_offsets, uint8_t* name_bytes, int64_t name_capacity, int32_t* score, uint8_t* strand,
                    int64_t* chrom_total_host, int64_t* name_total_host);

/* ---- residual predicates over join pairs (intersects_duckdb.py:1239-1243; bedtools -s / -S as
 * `AND a.strand = b.strand` / `AND a.strand != b.strand`, docs/recipes/bedtools-migration.rst:271-307) ------------
 * Keeps pair i iff  left_col[left_rows[i]] <op> right_col[right_rows[i]]  with both operands non-NULL; survivors keep
 * their order.  dtype names the element type of both columns; GIQLB_T_STRAND is uint8 strand codes where
 * GIQLB_STRAND_NULL is SQL NULL.  left_valid / right_valid: optional uint8[n_rows] masks (1 = value present) or NULL.
 * Floating point follows DuckDB's total order (NaN = NaN, NaN > every number).  Outputs must not alias the inputs
 * and hold up to n_pairs entries; *n_out_host receives the number kept (one host round trip). */
#define GIQLB_T_STRAND 0
#define GIQLB_T_U8 1
#define GIQLB_T_I32 2
#define GIQLB_T_I64 3
#define GIQLB_T_F32 4
#define GIQLB_T_F64 5
#define GIQLB_CMP_EQ 0
#define GIQLB_CMP_NE 1
#define GIQLB_CMP_LT 2
#define GIQLB_CMP_LE 3
#define GIQLB_CMP_GT 4
#define GIQLB_CMP_GE 5
int giqlb_pairs_filter(giqlb_ctx* ctx, const int32_t* left_rows, const int32_t* right_rows, int64_t n_pairs,
                       const void* left_col, const void* right_col, int dtype, const uint8_t* left_valid,
                       const uint8_t* right_valid, int op, int32_t* out_left, int32_t* out_right, int64_t* n_out_host);

/* SEMI/ANTI: exists_by_row[i] = 1 iff left input row i has >= 1 match (uint8[n]).  The left table is
 * probed from its raw columns in input order (no index, no sort).  giqlb_select_rows compacts the row ids
 * with flag == want (ascending) and returns how many. */
int giqlb_intersects_exists(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                            const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, giqlb_index* R,
                            uint8_t* exists_by_row);
int giqlb_select_rows(giqlb_ctx* ctx, const uint8_t* flags, int64_t n, int want, int32_t* out_rows,
                      int64_t* n_out_host);
/* count_overlaps: counts_by_row[i] = number of right rows overlapping left input row i (int64[n]) */
int giqlb_intersects_count_rows(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                                const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, giqlb_index* R,
                                int64_t* counts_by_row);

/* ---- NEAREST ----------------------------------------------------------------------------------------
 * For every query row (raw columns, input order; canonical offsets q_*_adj) the k nearest rows of the target
 * index T by ORDER BY ABS(distance), start, end (nearest.py:391-395), restricted to the query's chromosome,
 * to its strand when `stranded` (both tables carry a strand column, nearest.py:304-307) and to
 * ABS(distance) <= max_distance when max_distance >= 0 (nearest.py:330-331).  T must be built with
 * GIQLB_INDEX_TIE_BY_END on `lay` (only the TARGET table needs to be observed) and, for stranded, grouped by
 * strand.  Emits up to k rows per query: (query row, target row, distance, valid = 0 for SQL NULL distance);
 * the outputs must hold k * n_q rows.  Rows tied on (|distance|, start, end) are order-ambiguous in the
 * reference itself (nearest.py:366-372); here they come in index order (input order among full ties). */
int giqlb_nearest(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* q_code, const int32_t* q_start,
                  const int32_t* q_end, const uint8_t* q_strand, int64_t n_q, int32_t q_start_adj, int32_t q_end_adj,
                  giqlb_index* T, int k, int64_t max_distance, int stranded, int signed_, int32_t* out_q,
                  int32_t* out_t, int32_t* out_dist, uint8_t* out_valid, int64_t capacity, int64_t* n_out_host);

/* ---- CLUSTER / MERGE ------------------------------------------------------------------------------
 * X: an index built with RAW coordinates (start_adj = end_adj = 0: cluster.py works on the table's own
 * columns, expressions.py:223-226); for stranded := true build it with a strand column and call
 * giqlb_index_group_by_strand first.  distance <= 0 means "no bridging" (cluster.py:266-272).
 * cluster(): 1-based ids restarting per (chrom[, strand]) partition, written at the caller's row ids.
 * merge():   regions ordered by ([strand block,] chrom, start); returns GIQLB_ERR_CAPACITY and the required
 *            count in *n_out_host when capacity is too small (capacity = rows always suffices).
 *            out_strand (nullable) receives the strand code of each region, out_count (nullable) COUNT(*).
 * Rows with end < start are rejected (their grouping depends on the tie order of the executing engine). */
int giqlb_cluster(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, int32_t* cluster_id_by_input_row);
/* The same for an index that covers ALL n rows of a table, given that table's raw chromosome-code and start columns (the
 * offsets the index was built with): when the clusters are few, every input row finds its id by searching its own start among
 * the cluster starts, in input order, instead of the ids travelling back through the sorted row ids.  Identical results;
 * strand-grouped indexes and many-cluster tables take giqlb_cluster's path. */
int giqlb_cluster_table(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, const int32_t* code, const int32_t* start, int64_t n,
                        int32_t start_adj, int32_t* cluster_id_by_input_row);
/* CLUSTER(interval[, distance], predicate := <cur column> <op> PREV(<prev column>))  (cluster.py:281-299, 587-644): a row
 * stays in the running cluster only when it is adjacent to the running-max edge AND the comparison holds against its
 * immediate sorted predecessor (rows tied on start keep their input order).  The two operand columns are given as
 * order-preserving int32 codes per input row (the host ranks the values of both columns together); GIQLB_KEY_NULL is
 * SQL NULL, which makes the comparison not TRUE.  op: GIQLB_CMP_*. */
#define GIQLB_KEY_NULL (-2147483647 - 1)
int giqlb_cluster_predicate(giqlb_ctx* ctx, giqlb_index* ix, int64_t distance, const int32_t* cur_key_by_row,
                            const int32_t* prev_key_by_row, int op, int32_t* cluster_id_by_input_row);
int giqlb_merge(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, int32_t* out_code, int32_t* out_start,
                int32_t* out_end, uint8_t* out_strand, int64_t* out_count, int64_t capacity, int64_t* n_out_host);

/* ---- DISJOIN ----------------------------------------------------------------------------------------
 * Splits every target row (raw columns, canonical offsets t_*_adj) at the breakpoints (distinct starts and
 * ends) of the reference index REF; with self_reference = 0 only segments whose start is covered by a
 * reference row are kept (disjoin.py:184-190).  Emits (target row, segment start, segment end) in
 * CANONICAL coordinates, grouped by target row in input order, segments ascending.  `lay` needs only the
 * reference table observed.  GIQLB_ERR_CAPACITY + required count when the outputs are too small. */
int giqlb_disjoin(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* t_code, const int32_t* t_start,
                  const int32_t* t_end, int64_t n_t, int32_t t_start_adj, int32_t t_end_adj, giqlb_index* REF,
                  int self_reference, int32_t* out_row, int32_t* out_start, int32_t* out_end, int64_t capacity,
                  int64_t* n_out_host);

/* Genome-range shards of CLUSTER / MERGE.  X indexes the rows whose start lies in the shard's range
 * (giqlb_shard_select + giqlb_index_build_rows, raw coordinates, unstranded).  `carry` = max over the EARLIER shards
 * of giqlb_shard_max_end() (0 for the first shard): the running max of end of everything before this shard.
 *   cluster_shard(): counts_by_code_host (nullable, [n_codes]) <- clusters this shard STARTS per chromosome; with
 *     id_offset_by_code_host ([n_codes]: clusters started by earlier shards per chromosome) it writes the global
 *     1-based ids of the shard's rows into cluster_id_by_input_row (rows of other shards are not touched).
 *   merge_shard(): the regions that START in this shard (their end / count may still grow by the head fragments of
 *     later shards) and, in head_host[3], the shard's head fragment — rows that continue a region of an earlier
 *     shard: {row count, chromosome code or -1, max end in table coordinates}.  The host stitches heads onto the
 *     last region of the nearest earlier shard that starts one (giql_b200/sharding.py). */
int giqlb_shard_max_end(giqlb_ctx* ctx, giqlb_index* X, uint32_t* max_end_host);
int giqlb_cluster_shard(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, uint32_t carry, const uint32_t* id_offset_by_code_host,
                        int32_t* cluster_id_by_input_row, uint32_t* counts_by_code_host);
int giqlb_merge_shard(giqlb_ctx* ctx, giqlb_index* X, int64_t distance, uint32_t carry, int32_t* out_code, int32_t* out_start,
                      int32_t* out_end, int64_t* out_count, int64_t capacity, int64_t* n_out_host, int64_t* head_host);

/* ---- genome-range sharding of a join across GPUs ------------------------------
 * Every rank holds both tables and the same finalized layout.  shard_cuts(): histogram of the LEFT
 * table's linearised starts (2^19-cell bins) -> cuts_host[world + 1] (cuts[0] = 0, cuts[world] = 2^32 - 1) such that every
 * range [cuts[g], cuts[g+1]) holds ~n/world left rows; max_width_host = widest left interval (cells).  max_width_host may be
 * NULL (a caller that replicates no rows, e.g. CLUSTER / MERGE shards): tables of 2^20 rows or more are then histogrammed
 * from every 16th row only.
 * Deterministic: every rank computes identical cuts without any exchange.
 * shard_select(): ids of the rows with  gs_lo <= linear start < gs_hi  AND  linear end > ge_gt, in unspecified order
 *   (slices of ascending ids; the index build sorts the selection anyway).
 *   owner-of-left-start:  left  rows of rank g:  (cuts[g], cuts[g+1], 0)
 *   replicated straddlers: right rows of rank g:  (0, cuts[g+1] + max_width, cuts[g])
 * out_rows must hold n entries. */
int giqlb_shard_cuts(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                     const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, int world,
                     uint32_t* cuts_host, uint32_t* max_width_host);
int giqlb_shard_select(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                       const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, uint32_t gs_lo,
                       uint32_t gs_hi, uint32_t ge_gt, int32_t* out_rows, int64_t* n_out_host);

/* The sorted index of the rows of a table whose linear start lies in [gs_lo, gs_hi) — a genome-range shard of CLUSTER / MERGE
 * (same partition rule as above; src/giql/expanders/cluster.py:205-310, merge.py:183-439 run per partition) — built in one
 * call: fused select + key generation (one read of the columns), one round trip for the shard's size, sort.  Row ids of the
 * index are row ids of the TABLE.  Equivalent to giqlb_shard_select(ge_gt = 0) + giqlb_index_build_rows. */
int giqlb_index_build_range(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* code, const int32_t* start,
                            const int32_t* end, int64_t n, int32_t start_adj, int32_t end_adj, uint32_t gs_lo,
                            uint32_t gs_hi, int flags, giqlb_index** out);

/* ---- DISTANCE (elementwise over already-paired rows) ------------------------------------------
 * out[i] = distance, valid[i] = 0 for SQL NULL.  strands may be NULL when !stranded.  *_adj are the
 * canonical offsets of table a and table b. */
int giqlb_distance(giqlb_ctx* ctx, const int32_t* a_code, const int32_t* a_start, const int32_t* a_end,
                   const uint8_t* a_strand, const int32_t* b_code, const int32_t* b_start,
                   const int32_t* b_end, const uint8_t* b_strand, int64_t n, int stranded, int signed_,
                   int32_t a_start_adj, int32_t a_end_adj, int32_t b_start_adj, int32_t b_end_adj,
                   int32_t* out, uint8_t* valid);
/* The same over SELECTED pairs: pair i is (row a_rows[i] of table a, row b_rows[i] of table b), e.g. the two columns of a join
 * result (DISTANCE in the select list of an INTERSECTS / NEAREST query, src/giql/expanders/_distance.py:22-117 applied to the
 * joined rows).  a_rows / b_rows may each be NULL (= row i).  The columns are gathered inside the kernel.  a_n / b_n are the
 * table sizes: a row id outside its table is GIQLB_ERR_INVALID, never a wild read. */
int giqlb_distance_rows(giqlb_ctx* ctx, const int32_t* a_code, const int32_t* a_start, const int32_t* a_end,
                        const uint8_t* a_strand, int64_t a_n, const int32_t* b_code, const int32_t* b_start,
                        const int32_t* b_end, const uint8_t* b_strand, int64_t b_n, const int32_t* a_rows,
                        const int32_t* b_rows, int64_t n, int stranded, int signed_, int32_t a_start_adj,
                        int32_t a_end_adj, int32_t b_start_adj, int32_t b_end_adj, int32_t* out, uint8_t* valid);

#ifdef __cplusplus
}
#endif
#endif /* GIQLB_H */
