// K2b (tensor-core variant): linear-blend skinning with the per-vertex transform blend on
// tcgen05.
//
// Reference behaviour (lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py:134-155):
//   T_v  = sum_j weights[v][j] * A_j                  (th_T = th_results2 @ weights^T)
//   vert = (T_v @ [v_posed; 1])[:3]  (+ trans | - centre joint)
//
// The SIMT kernel (prk_skin.cu) is bound by shared-memory bandwidth: every (frame, vertex)
// pair pulls 4 x 12 floats of A_j through LDS.  Here the blend is what the reference wrote
// it as -- a dense GEMM over the 24 joints -- and runs on the otherwise idle tensor cores:
//
//   T[v][(f,e)] = sum_{j,p} Wsplit[v][8j+p] * Asplit[(f,e)][8j+p]         (M=128, N=192, K=192)
//
// with fp32-exact operands from three-way bf16 splits (six products per joint:
// Wh*Ah + Wh*Am + Wm*Ah + Wh*Al + Wm*Am + Wl*Ah, two zero slots pad each joint to one
// 16-byte chunk).  Both operands are BUILT IN SHARED MEMORY in the UMMA 128B-swizzled
// K-major layout, so nothing but the compact inputs crosses L2:
//   * Asplit (16 frames x 12 rows) once per CTA frame tile from fp32 A_j,
//   * Wsplit (128 vertices) per vertex tile from the compacted (joint, weight) pairs: only
//     the non-zero chunks are written and later cleared again, double buffered.
// The epilogue (thread = vertex = TMEM lane) reads T with tcgen05.ld, v_posed from TMA-staged
// shared memory, applies the 3x4 transform and writes coalesced vertex rows.
// HBM traffic per (frame, vertex): 12 B v_posed in, 12 B out -> HBM roofline.
//
// Warp roles (704 threads): warps 0-15 epilogue (TMEM lane quarter = warp & 3, frame group =
// warp >> 2, 4 frames each), 16-19 Wsplit builders, 20 MMA issuer (+TMEM alloc), 21 TMA
// producer for v_posed.  All waits are time-bounded (prk_tc.cuh).
#include "prk_internal.h"
#include "prk_tc.cuh"

#include <cuda_bf16.h>
#include <cstdlib>

namespace prk {

namespace {

using namespace tc;

constexpr int NF = 16;                         // frames per tile
constexpr int SK_N = NF * 12;                  // 192 accumulator columns (frame, matrix entry)
constexpr int SK_M = 128;                      // vertices per tile
constexpr int SK_KB = 3;                       // k-blocks of 64 bf16 (8 joints x 8 slots)
constexpr int SK_VT = (NV + SK_M - 1) / SK_M;  // 54 vertex tiles
constexpr int kAsplitKb = SK_N * 128;          // bytes per k-block of Asplit (24,576)
constexpr int kAsplitBytes = SK_KB * kAsplitKb;
constexpr int kWKb = SK_M * 128;               // bytes per k-block of Wsplit (16,384)
constexpr int kWBytes = SK_KB * kWKb;          // 49,152 per buffer
constexpr int kVpRows = 8;                     // frames per v_posed ring slot
constexpr int kVpSlots = 4;
constexpr int kVpBoxFloats = 192;              // TMA box: 8 rows x 192 floats, two boxes per slot
constexpr int kVpSlotBytes = kVpRows * 384 * 4;   // 12,288
constexpr int kSkSmem = kAsplitBytes + 2 * kWBytes + kVpSlots * kVpSlotBytes + NF * 3 * 4 + 320 + 1024;
constexpr int kEpiWarps = 16;                  // 4 TMEM lane quarters x 4 frame groups: 4 warps per scheduler hide latency
constexpr int kSkThreads = (kEpiWarps + 6) * 32;   // + 4 builder warps + MMA + TMA = 704
constexpr uint32_t kSkTmemCols = 512;

__device__ __forceinline__ uint32_t bf16_bits(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16_val(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ void split3(float v, uint32_t& h, uint32_t& m, uint32_t& l) {
    h = bf16_bits(v);
    const float r1 = v - bf16_val(h);
    m = bf16_bits(r1);
    l = bf16_bits(r1 - bf16_val(m));
}
// byte offset of joint j's 16-byte chunk in row `row` of a [rows][64 bf16] 128B-swizzled k-block set
__device__ __forceinline__ uint32_t chunk_off(int row, int j, int kb_bytes) {
    return (uint32_t)((j >> 3) * kb_bytes + row * 128 + (((j & 7) ^ (row & 7)) << 4));
}

#ifdef PRK_SKIN_DEBUG
__device__ unsigned long long g_skin_dbg[16];
#define DBG_T0() const long long _t0 = clock64()
#define DBG_ACC(k) do { if (lane == 0) atomicAdd(&g_skin_dbg[k], (unsigned long long)(clock64() - _t0)); } while (0)
#else
#define DBG_T0()
#define DBG_ACC(k)
#endif

__global__ void __launch_bounds__(kSkThreads, 1)
skin_mma_kernel(const __grid_constant__ CUtensorMap tmap_vp, const float* __restrict__ Askin,
                const float* __restrict__ off, const float4* __restrict__ wval, const uint32_t* __restrict__ widx,
                int nnz_groups, int64_t B, float* __restrict__ verts, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                   // Asplit  [3][192][128 B]
    uint8_t* sW = sA + kAsplitBytes;                      // Wsplit  [2][3][128][128 B]
    float* sVp = reinterpret_cast<float*>(sW + 2 * kWBytes);   // [4 slots][2 boxes][8 rows][192]
    float* sOff = sVp + kVpSlots * kVpSlotBytes / 4;      // [16][3]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOff + NF * 3);
    // Few waiters per mbarrier (wake-ups of warps parked on one barrier serialise): the accumulator
    // "full" signal is committed once per frame group, each v_posed box has its own barrier.
    uint64_t* w_full = bars;            // [2]
    uint64_t* w_empty = bars + 2;       // [2]
    uint64_t* t_full = bars + 4;        // [2 accumulators][4 frame groups]
    uint64_t* t_empty = bars + 12;      // [2]
    uint64_t* vp_full = bars + 14;      // [4 slots][2 boxes]
    uint64_t* vp_empty = bars + 22;     // [4]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 26);

#ifdef PRK_SKIN_DEBUG
    __shared__ long long s_ts[8];   // 0,1: commit time per acc; 2,3: last t_empty arrive per acc; 4,5: w_full arrive; 
#endif
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_vp)) : "memory");
        for (int i = 0; i < 2; ++i) {
            mbar_init(&w_full[i], 4); mbar_init(&w_empty[i], 1);
            mbar_init(&t_empty[i], kEpiWarps);
        }
        for (int i = 0; i < 8; ++i) { mbar_init(&t_full[i], 1); mbar_init(&vp_full[i], 1); }
        for (int i = 0; i < kVpSlots; ++i) mbar_init(&vp_empty[i], kEpiWarps / 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kEpiWarps + 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                     "r"(kSkTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Wsplit buffers start all-zero; afterwards only the non-zero chunks are written and cleared
    for (int i = tid; i < 2 * kWBytes / 16; i += kSkThreads) reinterpret_cast<uint4*>(sW)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_holder;

#ifdef PRK_SKIN_DEBUG
    const long long _k0 = clock64();
#endif
    const int64_t n_ftiles = (B + NF - 1) / NF;
    uint32_t tiles_done = 0;    // frame tiles this CTA has finished: all ring/phase counters derive from it
    // builders: group-0 pairs currently stored in this thread's row of W buffer 0 / 1
    float4 w_old0 = make_float4(0.f, 0.f, 0.f, 0.f), w_old1 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t id_old0 = 0, id_old1 = 0;

    for (int64_t ft = blockIdx.x; ft < n_ftiles; ft += gridDim.x, ++tiles_done) {
        const int64_t f0 = ft * NF;
        const int nf = (int)((B - f0) < NF ? (B - f0) : NF);
        const uint32_t item0 = tiles_done * SK_VT;          // SK_VT is even: buffer index == vt & 1
        const uint32_t vp0 = tiles_done * 2 * SK_VT;

        // ---- Asplit for this frame tile (all threads) ----
        for (int idx = tid; idx < SK_N * NJ; idx += kSkThreads) {
            const int n = idx / NJ, j = idx - n * NJ;
            const int f = n / 12, e = n - f * 12;
            const float v = f < nf ? Askin[((f0 + f) * NJ + j) * 12 + e] : 0.0f;
            uint32_t h, m, l;
            split3(v, h, m, l);
            // slots: Ah Am Ah Al Am Ah 0 0   (pairs with Wh Wh Wm Wh Wm Wl)
            *reinterpret_cast<uint4*>(sA + chunk_off(n, j, kAsplitKb)) = make_uint4(h | (m << 16), h | (l << 16), m | (h << 16), 0u);
        }
        if (tid < NF * 3) sOff[tid] = (tid / 3) < nf ? off[f0 * 3 + tid] : 0.0f;
        fence_proxy_async_smem();
        __syncthreads();

        if (warp < kEpiWarps) {
            // ===== epilogue: thread = vertex row = TMEM lane; each warp owns 4 of the tile's 16 frames =====
            const int quarter = warp & 3, fg = warp >> 2;       // fg: frames 4*fg .. 4*fg+3
            const int r = quarter * 32 + lane;
            const int hf = fg >> 1;                             // which 8-frame v_posed slot of the item
            // column 3r of the 384-float row lives in box (3r)/192 at float (3r)%192
            const int rb = (3 * r) / kVpBoxFloats, rc = (3 * r) - rb * kVpBoxFloats;
            for (int vt = 0; vt < SK_VT; ++vt) {
                const uint32_t item = item0 + vt;
                const int a = item & 1;
                const int v = vt * SK_M + r;
                const uint32_t vi = vp0 + 2 * vt + hf;
                const int slot = vi & (kVpSlots - 1);
                { DBG_T0(); mbar_wait(&t_full[a * 4 + fg], (item >> 1) & 1); DBG_ACC(0); }
#ifdef PRK_SKIN_DEBUG
                if (lane == 0 && warp == 0) { atomicAdd(&g_skin_dbg[9], (unsigned long long)(clock64() - s_ts[a])); atomicAdd(&g_skin_dbg[12], 1ull); }
#endif
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * SK_N + fg * 48);
                uint32_t T[48];
                if (dbg & 32) {
#pragma unroll
                    for (int q = 0; q < 48; ++q) T[q] = 0x3f800000u + q;
                } else {
                    tmem_ld_32x16(taddr + 0, T);
                    tmem_ld_32x16(taddr + 16, T + 16);
                    tmem_ld_32x16(taddr + 32, T + 32);
                }
                if (!(dbg & 64)) { DBG_T0(); mbar_wait(&vp_full[slot * 2 + rb], (vi / kVpSlots) & 1); DBG_ACC(1); }
                const float* vp = sVp + slot * (kVpSlotBytes / 4) + rb * (kVpRows * kVpBoxFloats) + ((fg & 1) * 4) * kVpBoxFloats + rc;
                float p[4][3];
#pragma unroll
                for (int fi = 0; fi < 4; ++fi) {
                    if (dbg & 4) { p[fi][0] = p[fi][1] = p[fi][2] = 1.0f; continue; }
                    p[fi][0] = vp[fi * kVpBoxFloats + 0]; p[fi][1] = vp[fi * kVpBoxFloats + 1]; p[fi][2] = vp[fi * kVpBoxFloats + 2];
                }
                tmem_ld_wait();
                // TMEM and the v_posed slot are in registers now: release them before the stores
                tcgen05_fence_before();
                __syncwarp();
#ifdef PRK_SKIN_DEBUG
                if (lane == 0) s_ts[2 + a] = clock64();
#endif
                if (lane == 0) { mbar_arrive(&t_empty[a]); if (!(dbg & 64)) mbar_arrive(&vp_empty[slot]); }
#pragma unroll
                for (int fi = 0; fi < ((dbg & 128) ? 0 : 4); ++fi) {
                    const int f = fg * 4 + fi;
                    if (f < nf && v < NV && !(dbg & 1)) {
                        const uint32_t* t = T + fi * 12;
                        float* o = verts + ((size_t)(f0 + f) * NV + v) * 3;
                        o[0] = __uint_as_float(t[0]) * p[fi][0] + __uint_as_float(t[1]) * p[fi][1] +
                               __uint_as_float(t[2]) * p[fi][2] + __uint_as_float(t[3]) + sOff[f * 3 + 0];
                        o[1] = __uint_as_float(t[4]) * p[fi][0] + __uint_as_float(t[5]) * p[fi][1] +
                               __uint_as_float(t[6]) * p[fi][2] + __uint_as_float(t[7]) + sOff[f * 3 + 1];
                        o[2] = __uint_as_float(t[8]) * p[fi][0] + __uint_as_float(t[9]) * p[fi][1] +
                               __uint_as_float(t[10]) * p[fi][2] + __uint_as_float(t[11]) + sOff[f * 3 + 2];
                    }
                }
            }
        } else if (warp < kEpiWarps + 4) {
            // ===== Wsplit builders: thread = vertex row =====
            // The (joint, weight) pairs of the next vertex tile are fetched before waiting for the
            // buffer, and the pairs written two items ago (the row's current occupant) are kept in
            // registers for the clean-up, so no global-load latency sits on the item's critical path.
            const int r = tid - kEpiWarps * 32;
            float4 w_nx = make_float4(0.f, 0.f, 0.f, 0.f);   // group 0 of vertex tile vt (prefetched)
            uint32_t id_nx = 0;
            if (r < NV) { w_nx = wval[r]; id_nx = widx[r]; }
            for (int vt = 0; vt < SK_VT; ++vt) {
                const uint32_t item = item0 + vt;
                const int s = item & 1;
                const int v = vt * SK_M + r;
                uint8_t* wb = sW + s * kWBytes;
                const float4 w_cur = w_nx;
                const uint32_t id_cur = id_nx;
                const int vn = v + SK_M;                            // prefetch the next tile's pairs
                if (vt + 1 < SK_VT && vn < NV) { w_nx = wval[vn]; id_nx = widx[vn]; }
                else { w_nx = make_float4(0.f, 0.f, 0.f, 0.f); id_nx = 0; }
                { DBG_T0(); mbar_wait(&w_empty[s], ((item >> 1) & 1) ^ 1); DBG_ACC(2); }    // MMAs that read this buffer have retired
#ifdef PRK_SKIN_DEBUG
                if (tid == kEpiWarps * 32) { atomicAdd(&g_skin_dbg[13], (unsigned long long)(clock64() - s_ts[s])); s_ts[4 + s] = clock64(); }
#endif
                if (dbg & 2) { fence_proxy_async_smem(); __syncwarp(); if (lane == 0) mbar_arrive(&w_full[s]); continue; }
                if (dbg & 16) { __syncwarp(); if (lane == 0) mbar_arrive(&w_full[s]); continue; }
                // clear the chunks of the vertex that occupied this row of the buffer (two items ago)
                const float4 w_old = s ? w_old1 : w_old0;
                const uint32_t id_old = s ? id_old1 : id_old0;
                if (w_old.x != 0.f) *reinterpret_cast<uint4*>(wb + chunk_off(r, id_old & 0xFF, kWKb)) = make_uint4(0, 0, 0, 0);
                if (w_old.y != 0.f) *reinterpret_cast<uint4*>(wb + chunk_off(r, (id_old >> 8) & 0xFF, kWKb)) = make_uint4(0, 0, 0, 0);
                if (w_old.z != 0.f) *reinterpret_cast<uint4*>(wb + chunk_off(r, (id_old >> 16) & 0xFF, kWKb)) = make_uint4(0, 0, 0, 0);
                if (w_old.w != 0.f) *reinterpret_cast<uint4*>(wb + chunk_off(r, id_old >> 24, kWKb)) = make_uint4(0, 0, 0, 0);
                if (nnz_groups > 1) {                               // dense rows: further groups straight from global
                    const int pv = vt >= 2 ? (vt - 2) * SK_M + r : (tiles_done > 0 ? (SK_VT - 2 + vt) * SK_M + r : -1);
                    for (int g = 1; g < nnz_groups && pv >= 0 && pv < NV; ++g) {
                        const float4 w = wval[(size_t)g * NV + pv];
                        const uint32_t id = widx[(size_t)g * NV + pv];
                        if (w.x != 0.f) *reinterpret_cast<uint4*>(wb + chunk_off(r, id & 0xFF, kWKb)) = make_uint4(0, 0, 0, 0);
                        if (w.y != 0.f) *reinterpret_cast<uint4*>(wb + chunk_off(r, (id >> 8) & 0xFF, kWKb)) = make_uint4(0, 0, 0, 0);
                        if (w.z != 0.f) *reinterpret_cast<uint4*>(wb + chunk_off(r, (id >> 16) & 0xFF, kWKb)) = make_uint4(0, 0, 0, 0);
                        if (w.w != 0.f) *reinterpret_cast<uint4*>(wb + chunk_off(r, id >> 24, kWKb)) = make_uint4(0, 0, 0, 0);
                    }
                }
                for (int g = 0; g < nnz_groups; ++g) {
                    float4 w = w_cur;
                    uint32_t id = id_cur;
                    if (g > 0) {
                        if (v >= NV) break;
                        w = wval[(size_t)g * NV + v];
                        id = widx[(size_t)g * NV + v];
                    }
                    const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (ws[k] != 0.f) {
                            uint32_t h, m, l;
                            split3(ws[k], h, m, l);
                            // slots: Wh Wh Wm Wh Wm Wl 0 0
                            *reinterpret_cast<uint4*>(wb + chunk_off(r, (id >> (8 * k)) & 0xFF, kWKb)) =
                                make_uint4(h | (h << 16), m | (h << 16), m | (l << 16), 0u);
                        }
                    }
                }
                if (s) { w_old1 = w_cur; id_old1 = id_cur; } else { w_old0 = w_cur; id_old0 = id_cur; }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&w_full[s]);     // one arrival per builder warp
            }
        } else if (warp == kEpiWarps + 4) {
            // ===== MMA issuer =====
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc(SK_M, SK_N);
                for (int vt = 0; vt < SK_VT; ++vt) {
                    const uint32_t item = item0 + vt;
                    const int s = item & 1;
                    { DBG_T0(); mbar_wait(&w_full[s], (item >> 1) & 1); DBG_ACC(3); }
                    { DBG_T0(); mbar_wait(&t_empty[s], ((item >> 1) & 1) ^ 1); DBG_ACC(4); }
#ifdef PRK_SKIN_DEBUG
                    if (item >= 2) { atomicAdd(&g_skin_dbg[10], (unsigned long long)(clock64() - s_ts[2 + s])); atomicAdd(&g_skin_dbg[11], (unsigned long long)(clock64() - s_ts[4 + s])); }
#endif
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)s * SK_N;
                    const uint32_t wa = smem_u32(sW + s * kWBytes), aa = smem_u32(sA);
#pragma unroll
                    for (int kb = 0; kb < ((dbg & 8) ? 0 : SK_KB); ++kb) {
                        const uint64_t adesc = make_smem_desc(wa + kb * kWKb);
                        const uint64_t bdesc = make_smem_desc(aa + kb * kAsplitKb);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                    }
#ifdef PRK_SKIN_DEBUG
                    s_ts[s] = clock64();
#endif
                    tcgen05_commit(&w_empty[s]);
#pragma unroll
                    for (int g = 0; g < 4; ++g) tcgen05_commit(&t_full[s * 4 + g]);
                }
            }
        } else {
            // ===== TMA producer: v_posed rows of the tile's frames, 8 frames x 128 vertices per slot =====
            if (lane == 0 && !(dbg & 64)) {
                // v_posed streams from HBM (2-3 us away): the 4-slot ring alone cannot cover that
                // latency, so every box is first pulled into L2 kPrefetch items ahead.
                constexpr int kPrefetch = 8;
                const int64_t ft_next = ft + gridDim.x;
                if (tiles_done == 0)
                    for (int pv = 0; pv < kPrefetch; ++pv)
                        for (int hf = 0; hf < 2; ++hf) {
                            tma_prefetch_l2_2d(&tmap_vp, pv * 384, (int)(f0 + hf * kVpRows));
                            tma_prefetch_l2_2d(&tmap_vp, pv * 384 + kVpBoxFloats, (int)(f0 + hf * kVpRows));
                        }
                for (int vt = 0; vt < SK_VT; ++vt) {
                    {   // prefetch item vt + kPrefetch (possibly the first items of this CTA's next frame tile)
                        int pvt = vt + kPrefetch;
                        int64_t pf0 = f0;
                        if (pvt >= SK_VT) { pvt -= SK_VT; pf0 = ft_next * NF; }
                        if (pf0 == f0 || ft_next < n_ftiles)
                            for (int hf = 0; hf < 2; ++hf) {
                                tma_prefetch_l2_2d(&tmap_vp, pvt * 384, (int)(pf0 + hf * kVpRows));
                                tma_prefetch_l2_2d(&tmap_vp, pvt * 384 + kVpBoxFloats, (int)(pf0 + hf * kVpRows));
                            }
                    }
                    for (int hf = 0; hf < 2; ++hf) {
                        const uint32_t vi = vp0 + 2 * vt + hf;
                        const int slot = vi & (kVpSlots - 1);
                        { DBG_T0(); mbar_wait(&vp_empty[slot], ((vi / kVpSlots) & 1) ^ 1); DBG_ACC(5); }
                        mbar_expect_tx(&vp_full[slot * 2 + 0], kVpSlotBytes / 2);
                        mbar_expect_tx(&vp_full[slot * 2 + 1], kVpSlotBytes / 2);
                        float* dst = sVp + slot * (kVpSlotBytes / 4);
                        const int row0 = (int)(f0 + hf * kVpRows);
                        tma_load_2d(&tmap_vp, &vp_full[slot * 2 + 0], dst, vt * 384, row0);
                        tma_load_2d(&tmap_vp, &vp_full[slot * 2 + 1], dst + kVpRows * kVpBoxFloats, vt * 384 + kVpBoxFloats, row0);
                    }
                }
            }
        }
        // every role has finished its 54 items: the accumulators of this frame tile are drained,
        // so Asplit may be rebuilt
        tcgen05_fence_before();
        __syncthreads();
        tcgen05_fence_after();
    }

#ifdef PRK_SKIN_DEBUG
    if (tid == 0) atomicAdd(&g_skin_dbg[8], (unsigned long long)(clock64() - _k0));
#endif
    if (warp == kEpiWarps + 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kSkTmemCols) : "memory");
    }
}

}  // namespace

#ifdef PRK_SKIN_DEBUG
extern "C" __attribute__((visibility("default"))) int prk_skin_debug_read(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_skin_dbg, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_skin_dbg, z, sizeof z); }
    return 0;
}
#endif

cudaError_t launch_skin_mma(const Model& m, const float* d_vposed, int64_t rows_pad, const float* d_Askin,
                            const float* d_off, int64_t B, float* d_verts, cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    static bool attr_set[64] = {};
    if (m.device >= 0 && m.device < 64 && !attr_set[m.device]) {
        cudaError_t e = cudaFuncSetAttribute(skin_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkSmem);
        if (e != cudaSuccess) return e;
        attr_set[m.device] = true;
    }
    CUtensorMap tmap_vp;   // v_posed [rows_pad][20736] fp32, box 8 rows x 192 floats, no swizzle
    if (encode_tmap_2d_ex(&tmap_vp, d_vposed, (uint64_t)rows_pad, VPOSED_PITCH, kVpRows, kVpBoxFloats, 4, 0) != PRK_OK)
        return cudaErrorInvalidValue;
    const int64_t n_ftiles = (B + NF - 1) / NF;
    int grid = m.sm_count > 0 ? m.sm_count : 148;
    if (grid > n_ftiles) grid = (int)n_ftiles;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("PRK_SKIN_DBGFLAGS"); dbg = e ? atoi(e) : 0; }
    skin_mma_kernel<<<grid, kSkThreads, kSkSmem, s>>>(tmap_vp, d_Askin, d_off, m.d_wval, m.d_widx, m.nnz_groups, B, d_verts, dbg);
    count_launch();
    return cudaGetLastError();
}

}  // namespace prk
