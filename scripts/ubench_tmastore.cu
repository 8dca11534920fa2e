// Does a bulk tensor store reach frame rows that are only 8-byte aligned?  Two maps over rows of equal parity (165,360 B apart),
// the odd one based 8 bytes early and addressed two columns to the right.  nvcc -arch=sm_100a -o ubench_tmastore.bin ubench_tmastore.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define NVC 20670
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, int s0, int s1, int col, int row, int hint, int parmask) {
    __shared__ __align__(128) float tile[2][16][12];
    const int lane = threadIdx.x;
    for (int c = 0; c < 12; ++c) tile[lane & 1][lane >> 1][c] = 1000.f * lane + c;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        for (int par = 0; par < 2; ++par) {
            if (!((parmask >> par) & 1)) continue;
            const CUtensorMap* m = par ? &m1 : &m0;
            const int c0 = (par ? s1 : s0) + col;
            if (hint)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                             ::"l"((unsigned long long)m), "r"(smem_u32(&tile[par][0][0])), "r"(c0), "r"(row), "l"(pol) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                             ::"l"((unsigned long long)m), "r"(smem_u32(&tile[par][0][0])), "r"(c0), "r"(row) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 100, col = argc > 2 ? atoi(argv[2]) : 96, row = argc > 3 ? atoi(argv[3]) : 2;
    const int hint = argc > 4 ? atoi(argv[4]) : 1, extra = argc > 5 ? atoi(argv[5]) : 0;
    const int parmask = argc > 6 ? atoi(argv[6]) : 3, promo = argc > 7 ? atoi(argv[7]) : 0;   // extra: byte offset of the buffer (0 or 8)
    float* base;
    cudaMalloc(&base, (size_t)B * NVC * 4 + 64);
    cudaMemset(base, 0, (size_t)B * NVC * 4 + 64);
    float* d = (float*)((char*)base + extra);
    void* fp; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    CUtensorMap m[2]; int sh[2];
    for (int par = 0; par < 2; ++par) {
        uintptr_t start = (uintptr_t)d + (uintptr_t)par * NVC * 4, mis = start & 15;
        sh[par] = (int)(mis / 4);
        cuuint64_t gdim[2] = {(cuuint64_t)(NVC + sh[par]), (cuuint64_t)((B + 1 - par) / 2)}, gstr[1] = {2ull * NVC * 4};
        cuuint32_t box[2] = {12, 16}, es[2] = {1, 1};
        CUresult r = enc(&m[par], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(start - mis), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode par %d: %d (shift %d)\n", par, (int)r, sh[par]);
    }
    k<<<1, 32>>>(m[0], m[1], sh[0], sh[1], col, row, hint, parmask);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> h((size_t)B * NVC);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0, written = 0;
    const int parmask_h = parmask;
    for (int f = 0; f < B; ++f)
        for (int c = 0; c < NVC; ++c) {
            const float v = h[(size_t)f * NVC + c];
            const int lane = f - 2 * row;           // frame 2*row + lane <-> lane
            const bool in_box = lane >= 0 && lane < 32 && c >= col && c < col + 12 && ((parmask_h >> (f & 1)) & 1);
            const float want = in_box ? 1000.f * lane + (c - col) : 0.f;
            if (v != want) { if (bad < 5) printf("mismatch frame %d col %d: %g want %g\n", f, c, v, want); ++bad; }
            written += v != 0.f;
        }
    printf("non-zero elements %ld, mismatches %ld\n", written, bad);
    return bad != 0;
}
