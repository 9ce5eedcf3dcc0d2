// Micro-benchmark: cost of __match_any_sync on (mostly distinct) 9-bit digits vs a ballot-per-bit emulation.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE, int BITS>
__global__ void __launch_bounds__(512) k(uint32_t *out, int iters) {
    uint32_t x = hash(blockIdx.x * 512 + threadIdx.x + 1), acc = 0;
    const uint32_t lt = (1u << (threadIdx.x & 31)) - 1u;
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        uint32_t d = (x >> 13) & ((1u << BITS) - 1u);
        uint32_t peers;
        if (MODE == 0) {
            peers = __match_any_sync(0xffffffffu, d);
        } else {
            peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < BITS; ++b) {
                uint32_t v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? v : ~v;
            }
        }
        acc += __popc(peers & lt) + (peers >> 31);
    }
    out[blockIdx.x * 512 + threadIdx.x] = acc;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    uint32_t *out; cudaMalloc(&out, 296 * 512 * 4);
    const int iters = 4096; const double ops = 296.0 * 16 * iters;  // warp-level operations
    printf("match_any 9-bit : %.3f ms  (%.1f cycles per warp-op per SM-quadrant-ish)\n", timeit([&] { k<0, 9><<<296, 512>>>(out, iters); }), 0.0);
    printf("ballot x9 9-bit : %.3f ms\n", timeit([&] { k<1, 9><<<296, 512>>>(out, iters); }));
    printf("match_any 4-bit : %.3f ms\n", timeit([&] { k<0, 4><<<296, 512>>>(out, iters); }));
    printf("ballot x4 4-bit : %.3f ms\n", timeit([&] { k<1, 4><<<296, 512>>>(out, iters); }));
    printf("warp-ops per launch: %.0f\n", ops);
    return 0;
}
