// Comparison bar for the hand-written LSD radix sort: cub::DeviceRadixSort::SortPairs on
// the same keys (uint32 linearised starts of the C2 workload shape) and uint32 payloads.  NOT part of the
// product; built and run by scripts/cub_bar/run.sh on the GPU box only.
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 10000000ull;
    std::vector<uint32_t> h(n);
    std::mt19937_64 rng(7);
    for (size_t i = 0; i < n; i++) h[i] = (uint32_t)(rng() % 3088269832ull);
    uint32_t *k0, *k1, *v0, *v1;
    cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
    uint32_t* src; cudaMalloc(&src, n * 4);
    cudaMemcpy(src, h.data(), n * 4, cudaMemcpyHostToDevice);
    void* tmp = nullptr; size_t tmp_bytes = 0;
    cub::DoubleBuffer<uint32_t> dk(k0, k1), dv(v0, v1);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)n, 0, 32);
    cudaMalloc(&tmp, tmp_bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int it = 0; it < 13; it++) {
        cudaMemcpy(k0, src, n * 4, cudaMemcpyDeviceToDevice);
        cub::DoubleBuffer<uint32_t> dk2(k0, k1), dv2(v0, v1);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk2, dv2, (int)n, 0, 32);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it >= 3 && ms < best) best = ms;
    }
    printf("{\"cub_sort_pairs_u32\": {\"n\": %zu, \"bits\": 32, \"best_ms\": %.4f, \"gbs_at_16B_per_key_per_pass_x4\": %.1f}}\n", n, best,
           (double)n * 16 * 4 / (best * 1e6));
    return 0;
}
