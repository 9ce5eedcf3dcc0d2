// Micro-benchmark: how fast can a pure streaming read (sum) go on this GPU, by load flavour and launch shape.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float4 ld_nc_na(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void ld_v8(const float *p, float4 &a, float4 &b) {
    asm volatile("ld.global.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) rd(const float *p, int64_t n4, float *out) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const float *q = p + 4 * (i + u * stride);
            if (MODE == 0) v[u] = *reinterpret_cast<const float4 *>(q);
            else if (MODE == 1) v[u] = ld_nc_na(q);
            else if (MODE == 2) v[u] = __ldcs(reinterpret_cast<const float4 *>(q));
            else v[u] = __ldg(reinterpret_cast<const float4 *>(q));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { float4 v = *reinterpret_cast<const float4 *>(p + 4 * i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

// contiguous chunk per CTA iteration (tile) like the tree kernel: CTA reads 16 KB tiles
template <int UNROLL>
__global__ void __launch_bounds__(256) rd_tile(const float *p, int64_t n4, float *out) {
    float acc = 0.f;
    const int64_t tiles = n4 / (256 * UNROLL);
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const float *q = p + 4 * (t * 256 * UNROLL);
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ld_nc_na(q + 4 * (u * 256 + threadIdx.x));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int UNROLL>
__global__ void __launch_bounds__(256) rd_v8(const float *p, int64_t n8, float *out) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n8; i += UNROLL * stride) {
        float4 a[UNROLL], b[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) ld_v8(p + 8 * (i + u * stride), a[u], b[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += a[u].x + a[u].y + a[u].z + a[u].w + b[u].x + b[u].y + b[u].z + b[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(256) cp(const float4 *a, float4 *b, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) b[i] = a[i];
}

template <class F>
float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    float best = 1e9;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int64_t n = 1ll << 30;
    float *p, *q, *out;
    CK(cudaMalloc(&p, n * 4)); CK(cudaMalloc(&q, n * 4)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(p, 0, n * 4)); CK(cudaMemset(q, 0, n * 4));
    const double gb = n * 4 / 1e9;
    int grids[] = {148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 32};
    for (int g : grids) {
        printf("grid %5d: ", g);
        printf("plain/u1 %6.0f  ", gb / timeit([&] { rd<0, 1><<<g, 256>>>(p, n / 4, out); }) * 1e3);
        printf("plain/u4 %6.0f  ", gb / timeit([&] { rd<0, 4><<<g, 256>>>(p, n / 4, out); }) * 1e3);
        printf("ncna/u4 %6.0f  ", gb / timeit([&] { rd<1, 4><<<g, 256>>>(p, n / 4, out); }) * 1e3);
        printf("ldcs/u4 %6.0f  ", gb / timeit([&] { rd<2, 4><<<g, 256>>>(p, n / 4, out); }) * 1e3);
        printf("ldg/u8 %6.0f  ", gb / timeit([&] { rd<3, 8><<<g, 256>>>(p, n / 4, out); }) * 1e3);
        printf("tile/u4 %6.0f  ", gb / timeit([&] { rd_tile<4><<<g, 256>>>(p, n / 4, out); }) * 1e3);
        printf("v8/u2 %6.0f  ", gb / timeit([&] { rd_v8<2><<<g, 256>>>(p, n / 8, out); }) * 1e3);
        printf("v8/u4 %6.0f  ", gb / timeit([&] { rd_v8<4><<<g, 256>>>(p, n / 8, out); }) * 1e3);
        printf("copy(r+w) %6.0f GB/s\n", 2 * gb / timeit([&] { cp<<<g, 256>>>((const float4 *)p, (float4 *)q, n / 4); }) * 1e3);
    }
    CK(cudaDeviceSynchronize());
    float ms = timeit([&] { cudaMemcpyAsync(q, p, n * 4, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D (r+w) %.0f GB/s\n", 2 * gb / ms * 1e3);
    return 0;
}
