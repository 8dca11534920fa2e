// FFMA vs packed FFMA2 (fma.rn.f32x2) throughput per SM on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_ffma2.bin scripts/ubench_ffma2.cu
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b, float2 c) {
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
}
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float2 a[8], b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
            else ffma2(a[i], a[i], b, c);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* o; long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 148 * 8);
    for (int warps : {4, 8, 16, 32}) {
        for (int mode = 0; mode < 2; ++mode) {
            int iters = 10000;
            if (mode == 0) k<0><<<148, warps * 32>>>(o, iters, c); else k<1><<<148, warps * 32>>>(o, iters, c);
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            double fma_lane_ops = (double)warps * 32 * iters * 16;
            printf("mode %s warps %2d: %lld clk, %.1f fp32 FMA lanes/clk/SM\n", mode ? "FFMA2" : "FFMA ", warps, h, fma_lane_ops / h);
        }
    }
    return 0;
}
