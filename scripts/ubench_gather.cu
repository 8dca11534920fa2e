// TMEM gather micro-benchmark for the fused kernel's epilogue: per "vertex" a warp gathers 4 joints x 12
// columns (+ 4 accumulator columns) from tensor memory at warp-uniform, data-dependent columns and runs
// the 28 FFMA2 of the skinning math.  Variants isolate load shape (x8+x4, 3 x x4, x16) and column pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_gather.bin scripts/ubench_gather.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../poserisk_release_b200/csrc/prk_tc.cuh"
using namespace prk::tc;

__device__ __forceinline__ void ld_x4(uint32_t t, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t) : "memory");
}
__device__ __forceinline__ void ld_x8(uint32_t t, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t) : "memory");
}
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// SHAPE 0: x8+x4 per joint, 1: 3 x x4 per joint, 2: x16 per joint (16-column pitch)
// PATTERN 0: random joints (table in smem), 1: sequential joints, 2: all warps same joint sequence
// STORE 0: none, 1: smem transpose + scattered global stores as in the fused kernel, 2: smem transpose only
template <int SHAPE, int PATTERN, bool MATH, int STORE = 0>
__global__ void __launch_bounds__(576, 1) gather_bench(int iters, const uint32_t* __restrict__ table, long long* out_cycles, uint32_t* sink,
                                                       float* __restrict__ verts = nullptr) {
    __shared__ uint32_t tmem_holder;
    __shared__ uint4 s_cols[256];
    __shared__ float4 s_w[512];
    __shared__ __align__(16) float s_out[18][32 * 12];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_w[i] = make_float4(0.25f, 0.25f, 0.125f, 0.125f);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint4 c;
        uint32_t* cc = reinterpret_cast<uint32_t*>(&c);
        for (int q = 0; q < 4; ++q) {
            uint32_t j = PATTERN == 0 ? table[(i * 4 + q) & 4095] % 24u : (uint32_t)((i * 4 + q) % 24);
            cc[q] = j * (SHAPE == 2 ? 16u : 12u);
        }
        s_cols[i] = c;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_holder, 0);
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const long long t0 = clock64();
    uint64_t accxy = 0, accz = 0;
    int idx = PATTERN == 2 ? 0 : warp * 61;
    float res[12];
    int rb_smem[3], rb_glob[3];
    for (int j = 0; j < 3; ++j) { const int q = j * 32 + lane; const int row = q / 6, c2 = (q - row * 6) * 2; rb_smem[j] = row * 12 + c2; rb_glob[j] = row * 20670 + c2; }
    float* my_out = s_out[warp];
    for (int it4 = 0; it4 < iters; it4 += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int it = it4 + kk;
        const uint4 cj = s_cols[idx & 255];
        const float4 wa = s_w[(idx * 2) & 511], wb = s_w[(idx * 2 + 1) & 511];
        const uint64_t ww[4] = {pack2(__float_as_uint(wa.x), __float_as_uint(wa.y)), pack2(__float_as_uint(wa.z), __float_as_uint(wa.w)),
                                pack2(__float_as_uint(wb.x), __float_as_uint(wb.y)), pack2(__float_as_uint(wb.z), __float_as_uint(wb.w))};
        idx += 1;
        if (STORE) { accxy = 0; accz = 0; }
        uint32_t p[4], r[SHAPE == 2 ? 64 : 48];
        ld_x4(t_lane + 400 + (it & 31) * 3, p);
        const uint32_t cols[4] = {cj.x, cj.y, cj.z, cj.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (SHAPE == 0) { ld_x8(t_lane + cols[q], r + q * 12); ld_x4(t_lane + cols[q] + 8, r + q * 12 + 8); }
            else if (SHAPE == 1) { ld_x4(t_lane + cols[q], r + q * 12); ld_x4(t_lane + cols[q] + 4, r + q * 12 + 4); ld_x4(t_lane + cols[q] + 8, r + q * 12 + 8); }
            else { tmem_ld_32x16(t_lane + cols[q], r + q * 16); }
        }
        tmem_ld_wait();
        const uint64_t pxx = pack2(p[0], p[0]), pyy = pack2(p[1], p[1]), pzz = pack2(p[2], p[2]), pxy = pack2(p[0], p[1]), pz1 = pack2(p[2], 0x3f800000u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t* a = r + q * (SHAPE == 2 ? 16 : 12);
            if (MATH) {
                uint64_t xy = fma2(pack2(a[0], a[1]), pxx, pack2(a[6], a[7]));
                xy = fma2(pack2(a[2], a[3]), pyy, xy);
                xy = fma2(pack2(a[4], a[5]), pzz, xy);
                uint64_t zz = fma2(pack2(a[8], a[9]), pxy, pack2(0u, 0u));
                zz = fma2(pack2(a[10], a[11]), pz1, zz);
                accxy = fma2(STORE ? ww[q] : pxx, xy, accxy);
                accz = fma2(STORE ? ww[q] : pyy, zz, accz);
            } else {
                accxy ^= pack2(a[0] ^ a[5], a[11]) ^ pxx;
            }
        }
        if (STORE) {
            const int k = kk;
            res[k * 3 + 0] = __uint_as_float((uint32_t)accxy); res[k * 3 + 1] = __uint_as_float((uint32_t)(accxy >> 32));
            res[k * 3 + 2] = __uint_as_float((uint32_t)accz) + __uint_as_float((uint32_t)(accz >> 32));
            if (k == 3) {
                float4* dst = reinterpret_cast<float4*>(my_out + lane * 12);
                dst[0] = make_float4(res[0], res[1], res[2], res[3]);
                dst[1] = make_float4(res[4], res[5], res[6], res[7]);
                dst[2] = make_float4(res[8], res[9], res[10], res[11]);
                __syncwarp();
                // rows: this CTA's 128 "frames" (warp & 3 selects the quarter), columns advance with the iteration
                float* vhalf = verts + ((size_t)(blockIdx.x * 128 + (warp & 3) * 32)) * 20670 + (size_t)(((it >> 2) * 4 + (warp >> 2)) % 1700) * 12;
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    const int j = t % 3, up = (t / 3) * 16;
                    const float2 val = *reinterpret_cast<const float2*>(my_out + rb_smem[j] + up * 12);
                    if (STORE == 1) *reinterpret_cast<float2*>(vhalf + rb_glob[j] + up * 20670) = val;
                    else if (val.x == 123.456f) sink[1] = 2;
                }
                __syncwarp();
            }
        }
      }
    }
    const long long t1 = clock64();
    tcgen05_fence_before();
    __syncthreads();
    if (lane == 0) out_cycles[blockIdx.x * 32 + warp] = t1 - t0;
    if (accxy == 0x12345678ull && accz == 1) sink[0] = 1;
    if (warp == 0) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int SHAPE, int PATTERN, bool MATH, int STORE = 0>
static void run(const char* name, int warps, const uint32_t* d_table, float* d_verts = nullptr) {
    const int grid = 148, iters = 3000;
    long long* d_c; uint32_t* d_s;
    cudaMalloc(&d_c, sizeof(long long) * 32 * grid); cudaMalloc(&d_s, 4);
    for (int rep = 0; rep < 2; ++rep) gather_bench<SHAPE, PATTERN, MATH, STORE><<<grid, warps * 32>>>(iters, d_table, d_c, d_s, d_verts);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> c(32 * grid);
    cudaMemcpy(c.data(), d_c, sizeof(long long) * 32 * grid, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) if (c[w] > mx) mx = c[w];
    const double bytes = (double)warps * iters * (SHAPE == 2 ? 68 : 52) * 128.0;
    printf("%-44s warps=%2d  %7.1f clk per vertex-warp, %6.1f clk per vertex per SM, %6.1f B/clk/SM\n", name, warps, (double)mx / iters,
           (double)mx / iters / warps, bytes / mx);
    cudaFree(d_c); cudaFree(d_s);
}

int main() {
    std::vector<uint32_t> tab(4096);
    uint32_t s = 12345;
    for (auto& t : tab) { s = s * 1664525u + 1013904223u; t = s >> 8; }
    uint32_t* d_table; cudaMalloc(&d_table, 4096 * 4);
    cudaMemcpy(d_table, tab.data(), 4096 * 4, cudaMemcpyHostToDevice);
    float* d_verts; cudaMalloc(&d_verts, (size_t)148 * 128 * 20670 * 4);
    for (int w : {16}) {
        run<0, 0, true, 0>("x8+x4, random, math (weights from smem)", w, d_table, d_verts);
        run<0, 0, true, 2>("x8+x4, random, math + smem transpose", w, d_table, d_verts);
        run<0, 0, true, 1>("x8+x4, random, math + transpose + STG", w, d_table, d_verts);
    }
    for (int w : {16}) {
        run<0, 0, true>("x8+x4, random joints, FFMA2 math", w, d_table);
        run<0, 1, true>("x8+x4, sequential joints, FFMA2 math", w, d_table);
        run<0, 2, true>("x8+x4, same joints in all warps, math", w, d_table);
        run<0, 0, false>("x8+x4, random joints, no math", w, d_table);
        run<1, 0, true>("3 x x4, random joints, FFMA2 math", w, d_table);
        run<2, 0, true>("x16 (16-col pitch), random joints, math", w, d_table);
        run<2, 0, false>("x16 (16-col pitch), random joints, no math", w, d_table);
    }
    return 0;
}
