// K1: shape + pose blend shapes as one tcgen05 / TMA GEMM.
//
// Reference behaviour (lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py:93-99):
//   v_shaped = v_template + shapedirs @ betas
//   v_posed  = v_shaped  + posedirs  @ (R_1..23 - I)
// restated as  v_posed[frame][vc] = sum_k A'[frame][k] * B'[vc][k]  with K = 704 bf16
// columns that carry hi/lo splits of every factor (prk_internal.h), so the fp32
// accumulator in tensor memory reproduces the fp32 reference to ~1e-7 relative.
//
// Kernel shape (persistent, one CTA per SM, 192 threads):
//   warp 0      TMA producer: A' tile 128x64 and B' tile 256x64 (128B swizzle) per k-block,
//               4-stage mbarrier ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=256, K=16 x4 per
//               k-block), accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> 128B-swizzled smem
//               staging (per warp, double buffered) -> TMA store of 32x32 fp32 boxes, so every
//               v_posed write is a full 128-byte line (row pitch 20736 floats)
// Roofline: tensor pipe.  Algorithmic work 2*217*20670 = 8.97 MFLOP/frame; executed MMA
// work 2*704*20736 = 29.2 MFLOP/frame (split precision x padding).
#include "prk_internal.h"
#include "prk_tc.cuh"

#include <cuda_bf16.h>

namespace prk {

namespace {

constexpr int kStages = 3;   // 3 x 48 KB + 32 KB store staging = 177 KB: leaves room for 4 skinning blocks per SM
constexpr int kATileBytes = GEMM_BM * GEMM_BK * 2;   // 16 KB
constexpr int kBTileBytes = GEMM_BN * GEMM_BK * 2;   // 32 KB
constexpr int kStageBytes = kATileBytes + kBTileBytes;
constexpr int kStoreBufBytes = 32 * 32 * 4;            // 32 rows x 32 fp32 columns (128-byte rows, 128B swizzle)
constexpr int kStoreBytes = 4 * 2 * kStoreBufBytes;    // 4 epilogue warps x 2 buffers
constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;
constexpr int kUmmaK = 16;

using namespace tc;

__global__ void __launch_bounds__(kThreads, 1)
blend_gemm_kernel(const __grid_constant__ CUtensorMap tmap_A, const __grid_constant__ CUtensorMap tmap_B,
                  const __grid_constant__ CUtensorMap tmap_D, int num_m_blocks, int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* store_smem = smem + kStages * kStageBytes;   // [4 warps][2][32 rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(store_smem + kStoreBytes);
    uint64_t* full_bar = bars;                 // [kStages]
    uint64_t* empty_bar = bars + kStages;      // [kStages]
    uint64_t* tfull_bar = bars + 2 * kStages;  // [2]
    uint64_t* tempty_bar = tfull_bar + 2;      // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_A)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_B)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_D)) : "memory");
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation (whole warp), 512 columns = two 128x256 fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % num_m_blocks, n_blk = tile / num_m_blocks;
                for (int kb = 0; kb < GEMM_KBLOCKS; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStageBytes;
                    mbar_expect_tx(&full_bar[stage], kStageBytes);
                    tma_load_2d(&tmap_A, &full_bar[stage], sa, kb * GEMM_BK, m_blk * GEMM_BM);
                    tma_load_2d(&tmap_B, &full_bar[stage], sa + kATileBytes, kb * GEMM_BK, n_blk * GEMM_BN);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(GEMM_BM, GEMM_BN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);     // epilogue drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * GEMM_BN;
                for (int kb = 0; kb < GEMM_KBLOCKS; ++kb) {
                    mbar_wait(&full_bar[stage], phase);          // TMA bytes landed
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t adesc = make_smem_desc(sa);
                    const uint64_t bdesc = make_smem_desc(sa + kATileBytes);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / kUmmaK; ++k) {
                        // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the encoded address
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                  (uint32_t)((kb | k) != 0));
                    }
                    tcgen05_commit(&empty_bar[stage]);           // frees the smem slot when MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(&tfull_bar[acc]);                 // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> swizzled smem -> TMA store (warps 2..5 own lane quarters 2,3,0,1) =====
        const int quarter = warp & 3;
        uint8_t* my_store = store_smem + quarter * 2 * kStoreBufBytes;
        int acc = 0; uint32_t acc_phase = 0;
        int sbuf = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % num_m_blocks, n_blk = tile / num_m_blocks;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tcgen05_fence_after();
            const int row0 = m_blk * GEMM_BM + quarter * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * GEMM_BN;
#pragma unroll 1
            for (int c = 0; c < GEMM_BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + (uint32_t)c * 32, r);
                // the TMA store that last read this staging buffer (two chunks ago) must be done with it
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
                tmem_ld_wait();
                uint8_t* buf = my_store + sbuf * kStoreBufBytes;
                // row = lane (128 B); 16-byte chunk q lands at q ^ (row & 7): the 128B-swizzle pattern of tmap_D
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4* d = reinterpret_cast<float4*>(buf + lane * 128 + ((q ^ (lane & 7)) << 4));
                    *d = make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                     __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_D, buf, n_blk * GEMM_BN + c * 32, row0);
                    tma_store_commit();
                }
                sbuf ^= 1;
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_read<0>();   // smem must stay valid until the last store has read it
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// Verification only: same bf16 operands, plain FFMA accumulation (prk_debug_blend use_simt=1).
__global__ void __launch_bounds__(256)
blend_simt_kernel(const uint16_t* __restrict__ Arows, const uint16_t* __restrict__ Bmat, int64_t rows,
                  float* __restrict__ vposed) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;   // vertex coordinate
    const int64_t f = blockIdx.y;
    if (n >= GEMM_N || f >= rows) return;
    const uint16_t* a = Arows + f * GEMM_K;
    const uint16_t* b = Bmat + (size_t)n * GEMM_K;
    float acc = 0.f;
    for (int k = 0; k < GEMM_K; ++k)
        acc = fmaf(__uint_as_float((uint32_t)a[k] << 16), __uint_as_float((uint32_t)b[k] << 16), acc);
    vposed[f * VPOSED_PITCH + n] = acc;
}

}  // namespace

cudaError_t launch_blend_gemm(const Model& m, const CUtensorMap& tmap_A, int64_t rows_pad, float* d_vposed,
                              cudaStream_t s) {
    if (rows_pad == 0) return cudaSuccess;
    static bool attr_set[64] = {};
    if (m.device >= 0 && m.device < 64 && !attr_set[m.device]) {
        cudaError_t e = cudaFuncSetAttribute(blend_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set[m.device] = true;
    }
    const int num_m_blocks = (int)(rows_pad / GEMM_BM);
    const int num_tiles = num_m_blocks * GEMM_NBLOCKS;
    int grid = m.sm_count > 0 ? m.sm_count : 148;
    if (grid > num_tiles) grid = num_tiles;
    CUtensorMap tmap_D;   // v_posed [rows_pad][20736] fp32, box 32 rows x 32 columns, 128B swizzle
    if (encode_tmap_2d(&tmap_D, d_vposed, (uint64_t)rows_pad, VPOSED_PITCH, 32, 32, 4) != PRK_OK) return cudaErrorInvalidValue;
    blend_gemm_kernel<<<grid, kThreads, kSmemBytes, s>>>(tmap_A, m.tmap_B, tmap_D, num_m_blocks, num_tiles);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_blend_simt(const Model& m, const uint16_t* d_Arows, int64_t rows, float* d_vposed,
                              cudaStream_t s) {
    if (rows == 0) return cudaSuccess;
    dim3 grid((GEMM_N + 255) / 256, (unsigned)rows);
    blend_simt_kernel<<<grid, 256, 0, s>>>(d_Arows, m.d_Bmat, rows, d_vposed);
    count_launch();
    return cudaGetLastError();
}

}  // namespace prk
