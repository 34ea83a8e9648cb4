// scan.cuh — single-pass prefix scan (decoupled look-back) over n elements, generic in the operator.
//
// One CTA per 2048-element tile.  Tiles take a ticket so a tile only waits on tiles already running;
// each tile publishes (flag | value) in ONE 64-bit status word, so no fence is needed between flag and
// payload.  Elements are loaded striped (coalesced), transposed through padded shared memory to a
// blocked arrangement for the serial part of the scan, and transposed back for the coalesced store.
// HBM traffic: whatever In() reads + whatever Out() writes + 8 B per tile.
//
//   struct Op  { typedef <u32|u64> T;  static T identity();  static T combine(T earlier, T later); }
//              T must fit in 62 bits.
//   struct In  { T operator()(int64_t i) const; }                      element i
//   struct Out { void operator()(int64_t i, T excl, T v) const; }      excl = scan of elements [0, i)
#pragma once
#include "common.cuh"

namespace giqlb_scan {

constexpr int SC_THREADS = 256;
// items per thread (measured on 200 M rows: running max 1.27 / 0.97 / 0.68 ms at 4 / 8 / 16; small inputs unchanged)
#ifndef GIQLB_SCAN_IPT
#define GIQLB_SCAN_IPT 16
#endif
constexpr int SC_IPT = GIQLB_SCAN_IPT;
constexpr int SC_TILE = SC_THREADS * SC_IPT;
constexpr unsigned long long SC_FLAG_AGG = 1ull << 62;
constexpr unsigned long long SC_FLAG_INCL = 2ull << 62;
constexpr unsigned long long SC_VALUE_MASK = (1ull << 62) - 1;
constexpr uint32_t SC_SPIN_LIMIT = 1u << 26;

__device__ __forceinline__ int sc_pad(int i) { return i + (i >> 5); }

template <typename Op, typename In, typename Out>
__global__ void __launch_bounds__(SC_THREADS)
scan_lookback_kernel(int64_t n, In in, Out out, unsigned long long* __restrict__ status,
                     uint32_t* __restrict__ ticket, int* __restrict__ err_flag) {
    typedef typename Op::T T;
    __shared__ T s_val[SC_TILE + SC_TILE / 32 + 1];
    __shared__ T s_warp[SC_THREADS / 32];
    __shared__ T s_prefix;
    __shared__ uint32_t s_tile;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * SC_TILE;
    const int tid = threadIdx.x;
    const uint32_t w = warp_id(), l = lane_id();

    T v[SC_IPT];
    In in_regs = in;   // a functor with per-thread state (a `mutable` memo) must not live in the kernel's parameter space
#pragma unroll
    for (int j = 0; j < SC_IPT; j++) {
        int64_t i = base + j * SC_THREADS + tid;
        v[j] = i < n ? in_regs(i) : Op::identity();
        s_val[sc_pad(j * SC_THREADS + tid)] = v[j];
    }
    __syncthreads();
    // blocked: thread owns elements [tid*IPT, tid*IPT + IPT)
    T b[SC_IPT];
    T run = Op::identity();
#pragma unroll
    for (int j = 0; j < SC_IPT; j++) {
        b[j] = s_val[sc_pad(tid * SC_IPT + j)];
        run = Op::combine(run, b[j]);
    }
    // warp inclusive scan of thread totals (order-preserving: earlier lanes combine on the left)
    T incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(FULL_MASK, incl, d);
        if (l >= (uint32_t)d) incl = Op::combine(o, incl);
    }
    T warp_excl = __shfl_up_sync(FULL_MASK, incl, 1);
    if (l == 0) warp_excl = Op::identity();
    if (l == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        // lane 0: serial scan over the 8 warp totals, publish the tile aggregate.  Then the whole warp looks back,
        // 32 predecessor tiles per step (lane j reads tile - 1 - j), combining in tile order.
        volatile unsigned long long* st = status;
        T acc = Op::identity();
        if (l == 0) {
            for (int ww = 0; ww < SC_THREADS / 32; ww++) {
                T t = s_warp[ww];
                s_warp[ww] = acc;
                acc = Op::combine(acc, t);
            }
            st[tile] = (tile == 0 ? SC_FLAG_INCL : SC_FLAG_AGG) | (unsigned long long)acc;
        }
        T prefix = Op::identity();
        if (tile > 0) {
            int64_t base = (int64_t)tile - 1;
            uint32_t spins = 0;
            bool done = false;
            while (!done) {
                const int64_t k = base - (int64_t)l;
                unsigned long long sv = SC_FLAG_INCL | (unsigned long long)Op::identity();   // before tile 0
                if (k >= 0) sv = st[k];
                while (__any_sync(FULL_MASK, (sv & ~SC_VALUE_MASK) == 0)) {
                    if ((sv & ~SC_VALUE_MASK) == 0) {
                        __nanosleep(20);
                        sv = st[k];
                    }
                    if (++spins > SC_SPIN_LIMIT) {
                        *err_flag = 1;
                        sv |= SC_FLAG_INCL;
                    }
                }
                const uint32_t incl = __ballot_sync(FULL_MASK, (sv & ~SC_VALUE_MASK) == SC_FLAG_INCL);
                const int first_incl = incl ? __ffs(incl) - 1 : 32;
                T v = (int)l <= first_incl ? (T)(sv & SC_VALUE_MASK) : Op::identity();
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {   // lane j ends with combine(tile-1-31 .. tile-1-j): earlier tiles on the left
                    T o = __shfl_down_sync(FULL_MASK, v, d);
                    if (l + d < 32) v = Op::combine(o, v);
                }
                prefix = Op::combine(__shfl_sync(FULL_MASK, v, 0), prefix);
                done = incl != 0;
                base -= 32;
            }
            if (l == 0) st[tile] = SC_FLAG_INCL | (unsigned long long)Op::combine(prefix, acc);
        }
        if (l == 0) s_prefix = prefix;
    }
    __syncthreads();
    T excl = Op::combine(Op::combine(s_prefix, s_warp[w]), warp_excl);
#pragma unroll
    for (int j = 0; j < SC_IPT; j++) {
        s_val[sc_pad(tid * SC_IPT + j)] = excl;
        excl = Op::combine(excl, b[j]);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SC_IPT; j++) {
        int64_t i = base + j * SC_THREADS + tid;
        if (i < n) out(i, s_val[sc_pad(j * SC_THREADS + tid)], v[j]);
    }
}

// Enqueues the scan on ctx->stream.  Allocates and frees its own status words (stream-ordered).
template <typename Op, typename In, typename Out>
int launch_scan(giqlb_ctx* ctx, int64_t n, In in, Out out) {
    if (n <= 0) return GIQLB_OK;
    const int64_t tiles = ceil_div(n, SC_TILE);
    unsigned long long* status = nullptr;
    GIQLB_TRY(dev_alloc(ctx, &status, (size_t)tiles + 1));
    CUDA_TRY(ctx, cudaMemsetAsync(status, 0, ((size_t)tiles + 1) * sizeof(unsigned long long), ctx->stream));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(status + tiles);
    scan_lookback_kernel<Op, In, Out><<<(unsigned)tiles, SC_THREADS, 0, ctx->stream>>>(n, in, out, status, ticket, ctx->d_err);
    LAUNCH_CHECK(ctx);
    dev_free(ctx, status);
    return GIQLB_OK;
}

struct OpMaxU32 {
    typedef uint32_t T;
    __device__ static T identity() { return 0u; }
    __device__ static T combine(T a, T b) { return a > b ? a : b; }
};
struct OpAddU32 {   // sums known to stay below 2^32 (row counts: tables hold fewer than 2^30 rows); half the registers and shared memory
    typedef uint32_t T;
    __device__ static T identity() { return 0u; }
    __device__ static T combine(T a, T b) { return a + b; }
};
struct OpAddU64 {
    typedef unsigned long long T;
    __device__ static T identity() { return 0ull; }
    __device__ static T combine(T a, T b) { return a + b; }
};

}  // namespace giqlb_scan
