// Throughput of bulk tensor stores of narrow boxes into a [frames][pitch] fp32 output whose row pitch is a multiple of 16 bytes:
// 148 CTAs x 16 warps, every warp stores boxes of `bw` floats x 32 rows from a private shared-memory tile, walking the output the
// way the vertex kernel does (CTA -> 128-frame tile, warp -> 32-frame quarter and column group).  Prints GB/s.
// nvcc -arch=sm_100a -o ubench_tmastore2.bin ubench_tmastore2.cu -lcuda ; ./ubench_tmastore2.bin <box_cols 12|48> <frames>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define PITCH 20672
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int BW>
__global__ void __launch_bounds__(512) k(const __grid_constant__ CUtensorMap m, int n_ft, int sts) {
    extern __shared__ __align__(128) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = sm + warp * (32 * BW);
    const int quarter = warp & 3, oct = warp >> 2;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    // unit u = (frame tile, 48-column block): CTA b takes units b, b + grid, ...
    const int n_cb = 20670 / 48 + 1;     // 431 column blocks of 48 floats
    for (int u = blockIdx.x; u < n_ft * n_cb; u += gridDim.x) {
        const int ft = u / n_cb, cb = u % n_cb;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (sts) {
            for (int c = 0; c < BW; c += 4)
                *reinterpret_cast<float4*>(tile + lane * BW + c) = make_float4(u, c, lane, 1.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
        }
        if (lane == 0 && (BW == 12 || oct == 0)) {
            const int col = cb * 48 + (BW == 12 ? oct * 12 : 0), row = ft * 128 + quarter * 32;
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                         ::"l"((unsigned long long)&m), "r"(smem_u32(tile)), "r"(col), "r"(row), "l"(pol) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int bw = argc > 1 ? atoi(argv[1]) : 12, B = argc > 2 ? atoi(argv[2]) : 4096, sts = argc > 3 ? atoi(argv[3]) : 1;
    float* d;
    cudaMalloc(&d, (size_t)B * PITCH * 4);
    void* fp; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t gdim[2] = {20670, (cuuint64_t)B}, gstr[1] = {(cuuint64_t)PITCH * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, 32}, es[2] = {1, 1};
    CUresult r = ((Enc)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    const int smem = 16 * 32 * bw * 4;
    cudaFuncSetAttribute(k<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    cudaFuncSetAttribute(k<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (bw == 12) k<12><<<148, 512, smem>>>(m, B / 128, sts); else k<48><<<148, 512, smem>>>(m, B / 128, sts);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s  %.1f us  %.0f GB/s\n", cudaGetErrorString(e), ms * 1e3, (double)B * 20670 * 4 / ms / 1e6);
    }
    return 0;
}
