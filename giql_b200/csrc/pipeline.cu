// pipeline.cu — the whole INTERSECTS hot path behind ONE C call.
//
// giqlb_intersects_tables() runs, for two tables given as raw columns, exactly the sequence the host layer would
// issue call by call — dictionary encode (unless the caller passes codes), genome layout, sorted index of both
// tables, single-pass join — so a host written in any language pays one FFI crossing, and the GPU never idles
// waiting for the host: the dictionary is ordered on the device, the layout is finalised without a round trip, the
// join runs speculatively on the proper-interval path, and everything the host needs (pair total, dictionary names,
// degenerate-row counts, error flags) comes back with ONE synchronisation at the end.  When the device declines
// (more than 1024 distinct chromosome names, a degenerate row) the affected stage is redone on the synchronous path.
// Replaces, in one call, what DuckDB does for the statement emitted at
// src/giql/expanders/intersects_duckdb.py:1283-1334 (reference).
#include "common.cuh"

#ifndef GIQLB_PIPELINE_SHORT_WINDOW_SORT
#define GIQLB_PIPELINE_SHORT_WINDOW_SORT 1   // 0: always the full four-pass sort (A/B switch)
#endif

static int tables_impl(giqlb_ctx* ctx, giqlb_table_t* L, giqlb_table_t* R, int32_t n_codes, int32_t* out_l, int32_t* out_r,
                       int64_t capacity, int64_t* total_host, giqlb_dict** dict_out, bool async_dict) {
    const int32_t n_codes_in = n_codes;
    giqlb_dict* dict = nullptr;
    giqlb_dict_pending* pend = nullptr;
    const int32_t* slot2code = nullptr;
    const int* ok_flag = nullptr;
    giqlb_layout* lay = nullptr;
    giqlb_index *il = nullptr, *ir = nullptr;
    int st = GIQLB_OK;
    const bool same = L == R;
    const bool encode = !L->codes_valid;
    if (encode) {
        const int32_t* ro = same ? nullptr : R->chrom_offsets;
        const uint8_t* rb = same ? nullptr : R->chrom_bytes;
        const int64_t rn = same ? 0 : R->n;
        int32_t* rc = same ? nullptr : R->codes;
        if (async_dict && giqlb_dict_encode_async(ctx, L->chrom_offsets, L->chrom_bytes, L->n, ro, rb, rn, L->codes, rc, &pend, true) == GIQLB_OK) {
            n_codes = GIQLB_ASYNC_DICT_MAX_CODES;   // upper bound: the real size is known when the names come back
            giqlb_dict_pending_tables(pend, &slot2code, &ok_flag);   // the slot -> code translation rides in the observe pass
        } else {
            pend = nullptr;
            st = giqlb_dict_encode(ctx, L->chrom_offsets, L->chrom_bytes, L->n, ro, rb, rn, L->codes, rc, &dict);
            if (st != GIQLB_OK) return st;
            n_codes = giqlb_dict_size(dict);
        }
    }
    if (st == GIQLB_OK) st = giqlb_layout_create(ctx, n_codes, 0, &lay);
    if (st == GIQLB_OK) st = giqlb_layout_observe_remap(ctx, lay, L->codes, L->start, L->end, L->n, L->start_adj, L->end_adj, slot2code, ok_flag);
    if (st == GIQLB_OK && !same) st = giqlb_layout_observe_remap(ctx, lay, R->codes, R->start, R->end, R->n, R->start_adj, R->end_adj, slot2code, ok_flag);
    if (st == GIQLB_OK) st = giqlb_layout_finalize_async(ctx, lay);   // no round trip: errors surface with the pair total
    // the indexes are private to this call and can be rebuilt here, so they may take the sort with one radix pass fewer
    const int ix_flags = GIQLB_PIPELINE_SHORT_WINDOW_SORT ? GIQLB_INDEX_SORT_SHORT_WINDOWS : 0;
    if (st == GIQLB_OK) st = giqlb_index_build_impl(ctx, lay, L->codes, L->start, L->end, nullptr, nullptr, L->n, L->start_adj, L->end_adj, ix_flags, nullptr, &il);
    if (st == GIQLB_OK && !same) st = giqlb_index_build_impl(ctx, lay, R->codes, R->start, R->end, nullptr, nullptr, R->n, R->start_adj, R->end_adj, ix_flags, nullptr, &ir);
    if (st == GIQLB_OK) st = giqlb_intersects_join_impl(ctx, il, same ? il : ir, out_l, out_r, capacity, total_host, true);
    const bool resort = st == GIQLB_RETRY_FULL_SORT;   // a window of 256 positions holds more rows than the finishing pass orders
    if (resort) st = GIQLB_OK;
    // keep the failing call's message across the clean-up calls
    char msg[sizeof(ctx->err)];
    memcpy(msg, ctx->err, sizeof(msg));
    giqlb_index_free(ctx, il);
    if (!same) giqlb_index_free(ctx, ir);
    giqlb_layout_free(ctx, lay);
    if (pend) {
        // the join's round trip also landed the names (or the device's refusal to order them)
        int fallback = 0;
        if (st != GIQLB_OK && st != GIQLB_ERR_CAPACITY) cudaStreamSynchronize(ctx->stream);
        const int st2 = giqlb_dict_finish(ctx, pend, &dict, &fallback);
        if (fallback) {   // codes were never written: redo everything with the host-ordered dictionary
            cudaMemsetAsync(ctx->d_err, 0, 4 * sizeof(int), ctx->stream);
            return tables_impl(ctx, L, R, n_codes, out_l, out_r, capacity, total_host, dict_out, false);
        }
        if (st == GIQLB_OK) st = st2;
    }
    if (resort && st == GIQLB_OK) {
        // redo the call with fully sorted indexes, and keep sorting fully for a while: tables like this one (deep coverage,
        // amplicons) tend to come in series.  The chromosome codes are encoded again from the names when they were encoded here.
        giqlb_dict_free(dict);
        ctx->full_sort_credit = 64;
        return tables_impl(ctx, L, R, n_codes_in, out_l, out_r, capacity, total_host, dict_out, async_dict);
    }
    if (encode && dict && (st == GIQLB_OK || st == GIQLB_ERR_CAPACITY)) L->codes_valid = R->codes_valid = 1;
    if (dict_out) *dict_out = dict;
    else giqlb_dict_free(dict);
    if (st != GIQLB_OK) memcpy(ctx->err, msg, sizeof(msg));
    return st;
}

extern "C" int giqlb_intersects_tables(giqlb_ctx* ctx, giqlb_table_t* L, giqlb_table_t* R, int32_t n_codes, int32_t* out_l,
                                       int32_t* out_r, int64_t capacity, int64_t* total_host, giqlb_dict** dict_out) {
    if (!ctx || !L || !R || !total_host) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_tables: bad argument");
    if (dict_out) *dict_out = nullptr;
    if (L->codes_valid != R->codes_valid) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "intersects_tables: encode both tables or neither");
    return tables_impl(ctx, L, R, n_codes, out_l, out_r, capacity, total_host, dict_out, true);
}
