// nearest.cu — NEAREST(target, reference := q.interval, k, max_distance, stranded, signed) as a sorted-neighbour
// search: three candidate streams per query row, merged in the reference's ORDER BY.
//
// Semantics (reference, abdenlab/giql src/giql/expanders/nearest.py:255-397, LATERAL form): for every query row
//   candidates = target rows with t.chrom = q.chrom [AND q.strand = t.strand when stranded and both tables carry
//   a strand column, :304-307, :328-329] [AND ABS(distance) <= max_distance, :330-331];
//   distance   = DISTANCE(q, t) with the same stranded/signed flags (:313-324; _distance.py:67-117);
//   ORDER BY ABS(distance), t.start, t.end LIMIT k (:391-395).  Fewer than k candidates -> fewer rows; a query
//   row without candidates produces nothing (CROSS JOIN LATERAL).
//
// Device formulation.  The target index is sorted by (start, end) [GIQLB_INDEX_TIE_BY_END], optionally grouped
// into strand blocks, with two derived arrays: the running max of end (pmax) and the ends in ascending order
// (ge_sorted / pos_by_end; ties keep start order).  With the CASE of _distance.py every target row falls in
// exactly one class, each an ordered stream:
//   overlap     q.start < t.end AND q.end > t.start   |d| = 0      index order = (start, end) order
//   downstream  q.end <= t.start                       |d| = t.start - q.end + 1, ascending along the index
//   upstream    otherwise (t.end <= q.start)           |d| = q.start - t.end + 1, ascending walking ge_sorted
//               backwards; rows tied on end are visited in start order (forward inside the tie group)
// At equal |d| upstream rows start before downstream rows, so the answer is: the first overlaps, then a 2-way merge
// of the two streams that prefers upstream on ties.  Query rows are probed from their raw columns; their
// coordinates are clamped into the chromosome's guard cells for the searches (this preserves every comparison
// with indexed rows) while distances use the unclamped canonical coordinates.
#include "common.cuh"
#include "distance.cuh"
#include "scan.cuh"

namespace {

struct TargetView {
    const uint32_t* gs;
    const uint32_t* ge;
    const int32_t* rid;
    const uint8_t* strand;       // nullable
    const uint32_t* pmax;
    const uint32_t* ge_sorted;   // per strand block
    const int32_t* pos_by_end;   // block-relative position of each sorted end
    const int64_t* block_off;    // device [257]; nullptr when not grouped
    int64_t n;
    // bucket directories over gs / pmax / ge_sorted of every strand block (giqlb_index_ensure_dirs)
    const uint32_t* bmax1;       // block maxima of ge over 32 / 1024 rows (giqlb_index_ensure_block_max)
    const uint32_t* bmax2;
    const uint32_t* dir_gs;
    const uint32_t* dir_pmax;
    const uint32_t* dir_ges;
    const int32_t* dir_slot;
    int dir_shift;
    uint32_t dir_stride;
};

struct NearestArgs {
    int k;
    long long max_distance;   // < 0: none
    int stranded;             // both tables carry a strand column and stranded := true
    int signed_;
};

// `order` (nullable): the query rows sorted by linear start.  Thread x then handles row order[x], so neighbouring threads
// search neighbouring places of the target index (their loads share cache lines) instead of ~80 unrelated L2 round trips
// each; its k result slots are x * k .. x * k + k - 1, written straight into the output arrays (with every query row
// yielding exactly k rows — the common case — they need no compaction afterwards).
__global__ void __launch_bounds__(128)
nearest_kernel(ProbeCols Q, const uint8_t* __restrict__ q_strand, const uint32_t* __restrict__ order, const int4* __restrict__ packed,
               TargetView T, NearestArgs A,
               int32_t* __restrict__ tmp_q, int32_t* __restrict__ tmp_t, int32_t* __restrict__ tmp_d, uint8_t* __restrict__ tmp_v,
               uint32_t* __restrict__ cnt, unsigned long long* __restrict__ total, int* __restrict__ err) {
    unsigned long long emitted = 0;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < Q.n; x += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = order ? (int64_t)order[x] : x;
        uint32_t n_out = 0;
        const int64_t slot0 = x * A.k;
        uint32_t qs_lin, qe_lin;
        // the query row: from its packed record when the rows are visited in key order, from the columns otherwise
        int32_t q_code, q_start, q_end;
        uint8_t sq;
        if (packed) {
            const int4 r = __ldg(packed + i);
            q_code = r.x; q_start = r.y; q_end = r.z;
            sq = A.stranded ? (uint8_t)r.w : (uint8_t)GIQLB_STRAND_PLUS;
        } else {
            q_code = Q.code[i]; q_start = Q.start[i]; q_end = Q.end[i];
            sq = A.stranded ? q_strand[i] : (uint8_t)GIQLB_STRAND_PLUS;
        }
        int64_t blo = 0, bhi = T.n;
        bool any = T.n > 0 && probe_linearise_values(Q, q_code, q_start, q_end, qs_lin, qe_lin);
        int slot = 0;
        if (any && A.stranded) {
            if (sq == GIQLB_STRAND_NULL) any = false;   // NULL = anything is never true
            else {
                blo = T.block_off[sq];
                bhi = T.block_off[sq + 1];
                any = bhi > blo;
                slot = any ? T.dir_slot[sq] : 0;
            }
        }
        if (any) {
            const int32_t c = q_code;
            const uint32_t cb_lo = __ldg(Q.base + c), cb_hi = __ldg(Q.base + c + 1);
            const long long lo_c = __ldg(Q.lo32 + c);
            const long long qs = (long long)q_start + Q.sadj, qe = (long long)q_end + Q.eadj;
            const uint32_t* gs = T.gs + blo;
            const uint32_t* ge = T.ge + blo;
            const int32_t* rid = T.rid + blo;
            const int64_t nb = bhi - blo;
            const uint32_t* dgs = T.dir_gs + (size_t)slot * T.dir_stride;
            const uint32_t dlast = T.dir_stride - 2;
            // canonical coordinate of a linear one on chromosome c
            auto canon = [&](uint32_t g) { return lo_c + (long long)g - (long long)cb_lo - 1; };
            const bool null_dist = A.stranded && (sq == GIQLB_STRAND_DOT || sq == GIQLB_STRAND_QMARK);
            auto emit = [&](int64_t pos) {
                long long d = 0;
                bool ok = !null_dist && giqlb_distance_dev(qs, qe, sq, canon(gs[pos]), canon(ge[pos]), sq, A.stranded != 0,
                                                           A.signed_ != 0, &d);
                tmp_q[slot0 + n_out] = (int32_t)i;
                tmp_t[slot0 + n_out] = rid[pos];
                if (ok && (d > INT32_MAX || d < INT32_MIN)) err[1] = 3;   // reported as GIQLB_ERR_RANGE by the checked fetch
                tmp_d[slot0 + n_out] = ok ? (int32_t)d : 0;
                tmp_v[slot0 + n_out] = ok ? 1 : 0;
                n_out++;
            };
            const int64_t ub = dir_lower_bound(dgs, T.dir_shift, dlast, gs, qe_lin);   // first row with t.start >= q.end
            if (null_dist) {
                // every distance is NULL: ORDER BY degenerates to (start, end) over the chromosome's rows of the block,
                // and ABS(NULL) <= max_distance keeps nothing (nearest.py:330-331)
                if (A.max_distance < 0) {
                    int64_t j = dir_upper_bound(dgs, T.dir_shift, dlast, gs, cb_lo);
                    for (; j < nb && n_out < (uint32_t)A.k && gs[j] < cb_hi; j++) emit(j);
                }
            } else {
                // 1. overlaps, in index order
                const uint32_t* pmax = T.pmax + blo;
                int64_t j = min(ub, dir_upper_bound(T.dir_pmax + (size_t)slot * T.dir_stride, T.dir_shift, dlast, pmax, qs_lin));   // first row whose running max end > q.start
                // (rows are visited through the block maxima of ge: behind one long target row the running max stays high
                //  for the rest of the chromosome, and a plain scan would test every row before the query)
                while (n_out < (uint32_t)A.k) {
                    j = next_end_above(T.ge, T.bmax1, T.bmax2, blo + j, blo + ub, qs_lin) - blo;
                    if (j >= ub) break;
                    emit(j);
                    j++;
                }
                // 2. merge of the upstream and downstream streams
                if (n_out < (uint32_t)A.k) {
                    const uint32_t* ges = T.ge_sorted + blo;
                    const int32_t* pbe = T.pos_by_end + blo;
                    int64_t dn = ub;                                                                  // downstream cursor
                    const uint32_t* dges = T.dir_ges + (size_t)slot * T.dir_stride;
                    int64_t up_hi = dir_upper_bound(dges, T.dir_shift, dlast, ges, qs_lin);                  // rows with t.end <= q.start: [.., up_hi)
                    int64_t grp_lo = up_hi, grp_cur = up_hi;                                          // current tie group [grp_cur, up_hi) is empty
                    const long long maxd = A.max_distance;
                    while (n_out < (uint32_t)A.k) {
                        // next upstream candidate: inside the current tie group, else open the next group below
                        long long m_up = -1, m_dn = -1;
                        int64_t up_pos = -1;
                        while (true) {
                            if (grp_cur < up_hi) {
                                int64_t p = pbe[grp_cur];
                                if (gs[p] >= qe_lin) {   // q.end <= t.start: that row is DOWNSTREAM by the CASE order
                                    grp_cur++;
                                    continue;
                                }
                                up_pos = p;
                                m_up = qs - canon(ges[grp_cur]) + 1;
                                break;
                            }
                            if (grp_lo <= 0) break;
                            const uint32_t v = ges[grp_lo - 1];
                            if (v <= cb_lo) break;   // left the chromosome
                            up_hi = grp_lo;
                            grp_lo = dir_lower_bound(dges, T.dir_shift, dlast, ges, v);
                            grp_cur = grp_lo;
                        }
                        if (dn < nb && gs[dn] < cb_hi) m_dn = canon(gs[dn]) - qe + 1;
                        if (maxd >= 0) {
                            if (m_up > maxd) m_up = -1;
                            if (m_dn > maxd) m_dn = -1;
                        }
                        if (m_up < 0 && m_dn < 0) break;
                        if (m_up >= 0 && (m_dn < 0 || m_up <= m_dn)) {
                            emit(up_pos);
                            grp_cur++;
                        } else {
                            emit(dn);
                            dn++;
                        }
                    }
                }
            }
        }
        cnt[x] = n_out;
        emitted += n_out;
    }
    emitted = warp_reduce_add(emitted);
    if (lane_id() == 0 && emitted) atomicAdd(total, emitted);
}

struct InCnt {
    const uint32_t* cnt;
    __device__ unsigned long long operator()(int64_t i) const { return cnt[i]; }
};
// linear start of every query row (clamped into its chromosome's guard cells; rows that cannot match sort last), and the row
// itself as one 16-byte record {code, start, end, strand}: the search kernel visits the rows in key order, i.e. at random, and
// fetches one sector per query instead of one per column (ncu: 2.0 GB of DRAM reads for 10 M queries before, four sectors each)
__global__ void __launch_bounds__(256)
nearest_query_keys_kernel(ProbeCols Q, const uint8_t* __restrict__ q_strand, uint32_t* __restrict__ keys, int4* __restrict__ packed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Q.n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = Q.code[i], s = Q.start[i], e = Q.end[i];
        uint32_t gs, ge;
        keys[i] = probe_linearise_values(Q, c, s, e, gs, ge) ? gs : 0xffffffffu;
        packed[i] = make_int4(c, s, e, q_strand ? (int)q_strand[i] : (int)GIQLB_STRAND_PLUS);
    }
}

struct OutCompactNearest {
    const int32_t* tmp_q;
    const int32_t* tmp_t;
    const int32_t* tmp_d;
    const uint8_t* tmp_v;
    int k;
    int32_t *out_q, *out_t, *out_d;
    uint8_t* out_v;
    unsigned long long* total;
    int64_t n;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long v) const {
        for (unsigned long long j = 0; j < v; j++) {
            out_q[excl + j] = tmp_q[i * k + j];
            out_t[excl + j] = tmp_t[i * k + j];
            out_d[excl + j] = tmp_d[i * k + j];
            out_v[excl + j] = tmp_v[i * k + j];
        }
        if (i == n - 1) *total = excl + v;
    }
};

}  // namespace

extern "C" int giqlb_nearest(giqlb_ctx* ctx, const giqlb_layout* lay, const int32_t* q_code, const int32_t* q_start,
                             const int32_t* q_end, const uint8_t* q_strand, int64_t n_q, int32_t q_start_adj, int32_t q_end_adj,
                             giqlb_index* T, int k, int64_t max_distance, int stranded, int signed_, int32_t* out_q,
                             int32_t* out_t, int32_t* out_dist, uint8_t* out_valid, int64_t capacity, int64_t* n_out_host) {
    using namespace giqlb_scan;
    if (!ctx || !T || !n_out_host || k < 1 || n_q < 0) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "nearest: bad argument");
    if (T->layout != lay) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "nearest: target index was built on a different layout");
    if (!T->tie_by_end) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "nearest: target index must be built with GIQLB_INDEX_TIE_BY_END");
    GIQLB_TRY(require_row_ids(ctx, T, "nearest"));
    if (stranded && (!q_strand || !T->strand || !T->grouped_by_strand))
        return giqlb_fail(ctx, GIQLB_ERR_INVALID, "nearest: stranded needs a query strand column and a strand-grouped target index");
    if (!stranded && T->grouped_by_strand) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "nearest: unstranded search over a strand-grouped index");
    if ((double)k * (double)n_q > (double)capacity) {
        *n_out_host = (int64_t)k * n_q;
        return giqlb_fail(ctx, GIQLB_ERR_CAPACITY, "nearest: outputs must hold k * n_q = %lld rows", (long long)k * (long long)n_q);
    }
    begin_call(ctx);
    *n_out_host = 0;
    if (n_q == 0) return end_call(ctx);
    if (!out_q || !out_t || !out_dist || !out_valid) return giqlb_fail(ctx, GIQLB_ERR_INVALID, "nearest: null output");
    ProbeCols Q;
    GIQLB_TRY(probe_cols(ctx, lay, q_code, q_start, q_end, n_q, q_start_adj, q_end_adj, &Q));
    GIQLB_TRY(giqlb_index_ensure_pmax(ctx, T));
    GIQLB_TRY(giqlb_index_ensure_end_sorted(ctx, T));
    if (T->n > 0) GIQLB_TRY(giqlb_index_ensure_dirs(ctx, T, true, true));
    GIQLB_TRY(giqlb_index_ensure_block_max(ctx, T));
    uint32_t* cnt = nullptr;
    uint32_t *ka = nullptr, *kb = nullptr, *va = nullptr, *vb = nullptr;
    int4* packed = nullptr;
    const uint32_t* order = nullptr;
    unsigned long long* total = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 10);
    const size_t slots = (size_t)k * (size_t)n_q;
    int st = dev_alloc(ctx, &cnt, (size_t)n_q);
    // query rows in the order of their linear start (one key pass + the radix sort: ~0.4 ms per 10 M rows, against
    // searches that are several times faster once neighbouring threads share their cache lines)
    if (st == GIQLB_OK && n_q >= 4096 && n_q < (1ll << 30) && T->n > 0) {
        st = dev_alloc(ctx, &ka, (size_t)n_q);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &kb, (size_t)n_q);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &va, (size_t)n_q);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &vb, (size_t)n_q);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &packed, (size_t)n_q);
        if (st == GIQLB_OK) {
            {
                PhaseSpan ps(ctx, GIQLB_PH_KEYGEN);
                int64_t blocks = ceil_div(n_q, 256 * 4), cap = (int64_t)ctx->sm_count * 8;
                nearest_query_keys_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(Q, stranded ? q_strand : nullptr, ka, packed);
                LAUNCH_CHECK(ctx);
            }
            int in_b = 0;
            // the order only has to bring neighbouring queries to neighbouring threads: for k = 1, windows of 2^16 linear positions
            // (a few dozen target rows, one or two cache lines of the index) do that with two digit passes instead of four
            // (10 M queries: sort -0.14 ms, search +0.015 ms); k > 1 walks further from its starting point and keeps 2^8 windows
            // (three passes; with 2^16 the k = 5 search lost the 0.14 ms the sort had saved)
            const int coarse = k == 1 ? 16 : 8;
            const int begin_bit = lay->key_bits > coarse + 8 ? coarse : 0;
            st = giqlb_radix_sort_pairs_u32(ctx, ka, va, kb, vb, n_q, begin_bit, lay->key_bits, 1, &in_b);
            order = in_b ? vb : va;
        }
    }
    if (st == GIQLB_OK) {
        PhaseSpan ps(ctx, GIQLB_PH_NEAREST);
        cudaMemsetAsync(total, 0, 8, ctx->stream);
        TargetView V{T->gs, T->ge, T->rid, T->strand, T->pmax, T->ge_sorted, T->pos_by_end,
                     T->grouped_by_strand ? T->d_block_off : nullptr, T->n, T->bmax1, T->bmax2,
                     T->dir_gs, T->dir_pmax, T->dir_ges, T->d_dir_slot, T->dir_shift, T->dir_stride};
        NearestArgs A{k, max_distance < 0 ? -1ll : (long long)max_distance, stranded != 0, signed_ != 0};
        int64_t blocks = ceil_div(n_q, 128), cap = (int64_t)ctx->sm_count * 32;
        // results land in the output arrays themselves, k slots per query row
        nearest_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, ctx->stream>>>(Q, q_strand, order, order ? packed : nullptr, V, A, out_q, out_t, out_dist, out_valid, cnt, total, ctx->d_err);
        ctx->launches_this_call++;
        ctx->stats.kernels_total++;
        if (cudaGetLastError() != cudaSuccess) st = giqlb_fail(ctx, GIQLB_ERR_CUDA, "nearest: launch failed");
    }
    if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, total, 1);
    if (st == GIQLB_OK && (unsigned long long)ctx->h_scalars[0] != (unsigned long long)slots) {
        // some query rows found fewer than k candidates: close the gaps (copy the slotted rows aside, compact them back)
        int32_t *tmp_q = nullptr, *tmp_t = nullptr, *tmp_d = nullptr;
        uint8_t* tmp_v = nullptr;
        st = dev_alloc(ctx, &tmp_q, slots);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &tmp_t, slots);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &tmp_d, slots);
        if (st == GIQLB_OK) st = dev_alloc(ctx, &tmp_v, slots);
        if (st == GIQLB_OK) {
            cudaMemcpyAsync(tmp_q, out_q, slots * 4, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(tmp_t, out_t, slots * 4, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(tmp_d, out_dist, slots * 4, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(tmp_v, out_valid, slots, cudaMemcpyDeviceToDevice, ctx->stream);
            PhaseSpan ps(ctx, GIQLB_PH_NEAREST);
            st = launch_scan<OpAddU64>(ctx, n_q, InCnt{cnt},
                                       OutCompactNearest{tmp_q, tmp_t, tmp_d, tmp_v, k, out_q, out_t, out_dist, out_valid, total, n_q});
        }
        if (st == GIQLB_OK) st = fetch_scalars_checked(ctx, total, 1);
        dev_free(ctx, tmp_q);
        dev_free(ctx, tmp_t);
        dev_free(ctx, tmp_d);
        dev_free(ctx, tmp_v);
    }
    dev_free(ctx, cnt);
    dev_free(ctx, packed);
    dev_free(ctx, ka);
    dev_free(ctx, kb);
    dev_free(ctx, va);
    dev_free(ctx, vb);
    if (st != GIQLB_OK) return st;
    *n_out_host = ctx->h_scalars[0];
    ctx->stats.rows_in[0] = n_q;
    ctx->stats.rows_in[1] = T->n;
    ctx->stats.rows_out = *n_out_host;
    ctx->stats.algorithmic_bytes = 12 * n_q + 12 * T->n + (stranded ? n_q + T->n : 0) + 12 * *n_out_host;
    return end_call(ctx);
}
