// K12: blend-shape GEMM and linear-blend skinning in ONE kernel -- v_posed never leaves the SM.
//
// Reference behaviour (lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py):
//   :93-99    v_posed = v_template + shapedirs @ betas + posedirs @ (R_1..23 - I)
//   :134-144  T_v = sum_j weights[v][j] * A_j ;  vert = (T_v @ [v_posed; 1])[:3]
//   :148-155  + trans | - centre joint
//
// Shape of the computation (one CTA per SM, persistent over a contiguous range of work units;
// a unit = 128 frames x 32 vertices):
//   * frames are the TMEM lanes (M = 128), vertex coordinates the accumulator columns
//     (N = 96 = 32 vertices x 3), so after the MMA thread `lane` of an epilogue warp owns ONE
//     FRAME and walks over vertices: every per-vertex quantity (weights, joint ids) is
//     warp-uniform and every per-frame quantity is thread-private.
//   * the per-frame skinning transforms A_j (24 x 12 fp32) live in TENSOR MEMORY, columns
//     0..287 of the frame's lane, written once per frame tile with tcgen05.st and gathered per
//     vertex with tcgen05.ld at a warp-uniform column 12*joint.  TMEM reads run at ~260 B/clk/SM
//     (scripts/ubench_tmem.cu), twice what shared memory gives the same gather.
//   * the blend GEMM runs on tcgen05 (kind::f16, bf16 x bf16 -> fp32) with fp32-class accuracy
//     from split precision, but the hi/lo parts are stored ONCE (prk_internal.h "K12 operand
//     layout"): the A' tile of the 128 frames stays resident in shared
//     memory for the whole frame tile (28 k-steps, 112 KB), only the 12 KB B' chunks stream through a TMA ring, and
//     each B' k-step is multiplied with every A' k-step it pairs with (44 MMAs per unit).
//   * B' is stored as pre-swizzled chunk images, so every 12 KB chunk arrives with one linear bulk copy.
//   * accumulators are double buffered in TMEM (columns 288..383 / 384..479), so the MMAs of
//     unit i+1 run under the skinning of unit i.
//   * vertices leave through a staging tile shared by the four warps of a TMEM lane quarter, so every
//     frame row is written as one 192-byte run, with evict-first stores; the read-back + store of a half
//     unit is issued in the middle of the next half's gathers, and the tile is handed over through
//     split-phase mbarriers, so the store phase hides under the tensor-memory gathers.
//   * per (vertex, joint) item the epilogue issues 2 gathers + 6 packed FMAs: x, y are accumulated per
//     joint, the z row of the blended transform is accumulated first and applied once per vertex; the
//     per-tile column table holds absolute tensor-memory addresses per lane quarter.
// HBM traffic per frame: 82,680 B of vertices out, ~2.2 KB of operands in (B' is L2 resident)
// -> HBM roofline; executed MMA work 216 tiles x 44 MMAs x 2*96*16 = 29.2 MFLOP/frame.
//
// Warp roles (576 threads): warps 0-15 epilogue (TMEM lane quarter = warp & 3, vertex octet =
// warp >> 2), warp 16 TMA producer, warp 17 TMEM allocator + MMA issuer.  All mbarrier waits
// are time-bounded (prk_tc.cuh).
#include "prk_internal.h"
#include "prk_tc.cuh"
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <atomic>
#include <cstdlib>

namespace prk {

namespace {

using namespace tc;

#ifdef PRK_FUSED_SPIN
#define MBAR_WAIT mbar_wait_spin
#else
#define MBAR_WAIT mbar_wait
#endif

constexpr int kAChunkBytes = FUSED_BM * 128;                     // 128 frames x 64 bf16
constexpr int kABytes = FUSED_A_CHUNKS * kAChunkBytes;           // 114,688: resident A' tile
constexpr int kBChunkBytes = FUSED_BN * 128;                     // 96 vertex coords x 64 bf16 = 12,288
constexpr int kEpiWarps = 16;
constexpr int kOutPitch = 52;                                    // floats per staged frame row: 16 vertices x 3 (+4: conflict-free float4)
constexpr int kOutBytesPerQuarter = 32 * kOutPitch * 4;          // 32 frames of one TMEM lane quarter
constexpr int kOutBytes = 4 * kOutBytesPerQuarter;               // 26,624
constexpr int kWSlotsMax = 8;
// weight slots of the ring: a slot is reused kWS units later, which must be more than the producer's lead over the epilogue
// (ring length in units + the two accumulator buffers + the unit in progress): 4 with the 6-chunk ring (< 1 unit),
// 5 with the 10 .. 12 half-chunk rings of the CTA pairs (1.25 .. 1.5 units: the MMA of unit i-2 has started when the weights of
// unit i are loaded, so the epilogue is past the middle of unit i-4 and units i-4 .. i-1 may still read their slots)
#ifndef PRK_WSLOTS_PAIR
#define PRK_WSLOTS_PAIR 5
#endif
constexpr int w_slots(int pair) { return pair == 2 ? PRK_WSLOTS_PAIR : 4; }
constexpr int kMaxStages = 12;                                   // 6 x 12 KB chunks, or 12 x 6 KB half chunks (CTA pairs)
constexpr int kNumBars = 2 * kMaxStages + 2 * kEpiWarps + 2 + kWSlotsMax + 2 + 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;                   // 576
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol0 = 288;                               // accumulators behind the 24 x 12 A_j columns
constexpr int kSmemLimit = 232448;                               // 227 KB opt-in maximum per CTA

// Vertex stores by TMA (kTma): possible when the caller's vertex rows are 16-byte aligned (row pitch a multiple of 4 floats,
// e.g. 20,672 instead of 20,670 floats: prk_smpl_forward_pitched).  Every epilogue warp then owns 8 CONSECUTIVE vertices of the
// tile, stages its 32 frames x 24 floats in a private tile and one lane issues ONE bulk tensor store per unit: no read-back,
// no store instruction, no hand-over between the warps of a lane quarter, ragged edges clipped by the TMA unit.
// (A bulk tensor store needs the global address of the box start 16-byte aligned -- scripts/ubench_tmastore.cu -- which is
// why the dense 82,680-byte rows of the reference layout cannot use it: every second row starts 8 bytes off.)
constexpr int kWarpTileBytes = 32 * 24 * 4;                      // [32 frames][8 vertices x 3] = 3,072 B per epilogue warp
constexpr int kOutBytesTma = kEpiWarps * kWarpTileBytes;         // 49,152
constexpr int fused_smem_bytes(int stages, int groups, int pair = 1, bool tma = false) {
    return kABytes + stages * (kBChunkBytes / pair) + (tma ? kOutBytesTma : kOutBytes) + w_slots(pair) * groups * FUSED_WGROUP_BYTES + kNumBars * 8 + 16 +
           1024 /*alignment slack*/;
}

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 339 MB of vertices stream through the L2 that also holds the 21 MB blend matrix every frame tile re-reads:
// vertex stores are evict-first (st.global.cs).  (An evict_last hint on the B' loads measured no gain.)
__device__ __forceinline__ void store_vertex_pair(float* dst, float2 v) {
    __stcs(reinterpret_cast<float2*>(dst), v);                  // st.global.cs: evict-first
}
// ---- CTA pairs (cta_group::2): two CTAs of a cluster walk the same vertex tiles for two consecutive frame tiles; ONE
// tcgen05.mma (M = 256) issued by the pair's leader multiplies both A' tiles with a B' k-step of which each CTA holds
// half (48 of the 96 rows): half the shared-memory fill per SM and chunk, half the B operand reads, half the MMA issues.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {   // same barrier, CTA `cta` of the cluster
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
constexpr uint32_t kLeaderMask = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared-memory address: the pair's CTA 0
// tensor-map load into THIS CTA's shared memory whose bytes are counted on the LEADER's mbarrier (same offset)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kLeaderMask), "r"(c0), "r"(c1)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once every MMA issued so far has retired
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// bulk tensor store shared -> global of one box; evict-first like the plain vertex stores
__device__ __forceinline__ void tma_store_box(const CUtensorMap* map, const void* src, int c0, int c1, uint64_t policy) {
#ifdef PRK_TMA_NOHINT
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
    return;
#endif
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void quarter_barrier(int quarter) {   // the 4 epilogue warps of one TMEM lane quarter
    asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// descriptor offset (units of 16 B) of A' k-step `a` inside the resident tile: chunk a/4, 32 B per k-step
__host__ __device__ constexpr uint32_t a_step_off(int a) { return (uint32_t)((a >> 2) * (kAChunkBytes >> 4) + (a & 3) * 2); }
// kind::f16 MMA with the two shared-memory descriptors given as (low word, common high word):
// all tiles here share SBO / version / swizzle, only the 14-bit start address differs.
// One MMA of a K12 k-step: D[tmem] (+)= A[smem] * B[smem]^T over 32 bytes of K per row.  kF8: kind::f8f6f4 (32 e4m3 elements),
// else kind::f16 (16 fp16 or bf16 elements, chosen by the instruction descriptor).  The two shared-memory descriptors are given as
// (low word, common high word): all tiles share SBO / version / swizzle, only the 14-bit start address differs.
template <int kPair, bool kF8>
__device__ __forceinline__ void umma_step(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                          bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
#define PRK_UMMA(GROUP, KIND)                                                                       \
    asm volatile(                                                                                   \
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"                                              \
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"                                        \
        "setp.ne.b32 p, %5, 0;\n\t"                                                                 \
        "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%0], da, db, %4, p;\n\t}"                  \
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory")
    if (kPair == 2) { if (kF8) PRK_UMMA("2", "f8f6f4"); else PRK_UMMA("2", "f16"); }
    else            { if (kF8) PRK_UMMA("1", "f8f6f4"); else PRK_UMMA("1", "f16"); }
#undef PRK_UMMA
}
// instruction descriptors (D = fp32, K-major operands): kind::f16 with fp16 inputs, kind::f8f6f4 with e4m3
__host__ __device__ constexpr uint32_t idesc_common(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__host__ __device__ constexpr uint32_t idesc_fp16(int M, int N) { return idesc_common(M, N); }                          // a/b format 0 = F16
__host__ __device__ constexpr uint32_t idesc_e4m3(int M, int N) { return idesc_common(M, N); }                          // 0 = E4M3

// Knock-out builds for finding out where the time goes: -DPRK_KNOCK=<mask> removes one side of the pipeline at
// COMPILE time (run-time switches perturb the unrolled epilogue too much to be trusted; profiles/r2_ab_runs.txt):
//   1 = no MMAs issued, 2 = epilogue skips gather + math + stores, 4 = no global stores (staging and read-back kept),
//   8 = no TMEM gather of A_j (math on stale registers), 16 = no B' loads (producer only signals),
//   32 = no skinning FMAs (gathers kept), 64 = constant weights / columns, 128 = no staging, read-back or stores,
//   256 = no read-back and no stores (staging and hand-shakes kept).
// -DPRK_FUSED_DEBUG adds clock64 phase timers to the epilogue (they cost ~50 % themselves).
#ifndef PRK_KNOCK
#define PRK_KNOCK 0
#endif
#define DBG(bit) ((PRK_KNOCK) & (bit))
#ifdef PRK_FUSED_DEBUG
__device__ unsigned long long g_fdbg[8];
#define TCLK(var) const long long var = clock64()
#define TACC(k, a, b) t_sum[k] += (b) - (a)
#else
#define TCLK(var)
#define TACC(k, a, b)
#endif

// Packed fp32 pairs (Blackwell FFMA2): one issue slot per two FMAs -- the epilogue is issue-bound.
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) { return pack2(__float_as_uint(lo), __float_as_uint(hi)); }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    uint32_t a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
    lo = __uint_as_float(a); hi = __uint_as_float(b);
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// kGroups: weight groups of 4 per vertex known at compile time (1 = SMPL), 0 = run-time `groups`
// the table holds absolute tensor-memory addresses (4 copies, one per lane quarter) or plain columns (1 copy)
#define GATHER_ADDR(col) (FUSED_WCOL_COPIES > 1 ? (col) : (col) + t_quarter)

// kPair = 2: CTA pairs (see above); n_units then counts pair units: (two consecutive frame tiles) x vertex tile
template <int kGroups, int kPair, bool kTma>
__global__ void __launch_bounds__(kThreads, 1)
fused_blend_skin_kernel(const __grid_constant__ CUtensorMap tmap_A, const __grid_constant__ CUtensorMap tmap_B,
                        const __grid_constant__ CUtensorMap tmap_V, const uint16_t* __restrict__ B2img,
                        const float* __restrict__ AskinT, const float* __restrict__ off,
                        const uint8_t* __restrict__ wpack, int groups_rt, int stages, int64_t B, int64_t n_units,
                        float* __restrict__ verts, int dbg, int switch_cost16) {
    const int groups = kGroups > 0 ? kGroups : groups_rt;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by OFFSET on the __shared__ array, so every derived pointer stays in the
    // shared address space (LDS/STS instead of generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                                           // [8 chunks][128 rows][128 B], 128B swizzle
    uint8_t* sB = sA + kABytes;                                   // [stages][96 rows][128 B]
    constexpr int kBSlotBytes = kBChunkBytes / kPair;             // a CTA of a pair holds 48 of a chunk's 96 rows
    uint8_t* sOut = sB + stages * kBSlotBytes;                   // [16 warps][32 frames][12 floats]
    uint8_t* sW = sOut + (kTma ? kOutBytesTma : kOutBytes);                               // [4 slots][groups][32 x float4 weights | 32 x uint4 columns]
    constexpr int kWS = w_slots(kPair);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + kWS * groups * FUSED_WGROUP_BYTES);
    uint64_t* full_bar = bars;                                    // [kMaxStages]
    uint64_t* empty_bar = bars + kMaxStages;                      // [kMaxStages]
    // one "accumulator full" barrier PER EPILOGUE WARP: warps parked on a common mbarrier are woken
    // one after the other (~2.5 us per unit for 16 waiters), a private barrier wakes at once
    uint64_t* tfull_bar = bars + 2 * kMaxStages;                  // [2 accumulators][kEpiWarps]
    uint64_t* tempty_bar = tfull_bar + 2 * kEpiWarps;             // [2]
    uint64_t* wfull_bar = tempty_bar + 2;                         // [kWSlotsMax]
    uint64_t* afull_bar = wfull_bar + kWSlotsMax;
    uint64_t* aempty_bar = afull_bar + 1;
    // staging tile of a lane quarter: "staged" (all four warps wrote half h) / "flushed" (all four read it back);
    // arrive and wait are far apart in the instruction stream, so the four warps need not run in lock-step
    uint64_t* qstaged_bar = aempty_bar + 1;                       // [4 quarters]
    uint64_t* qflushed_bar = qstaged_bar + 4;                     // [4 quarters]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(qflushed_bar + 4);

    // The warp index is broadcast from lane 0 so the compiler knows it is warp-uniform: role
    // branches stay convergent and the MMA / TMA warps compute their operands in uniform registers.
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    uint32_t lane_opaque = threadIdx.x & 31;
    asm volatile("" : "+r"(lane_opaque));                         // never re-derived from S2R inside the loops
    const int lane = (int)lane_opaque;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_A)) : "memory");
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&tfull_bar[i], 1);
        for (int i = 0; i < 2; ++i) mbar_init(&tempty_bar[i], kEpiWarps * kPair);   // pair: both CTAs' epilogue warps arrive on the leader's
        for (int i = 0; i < kWSlotsMax; ++i) mbar_init(&wfull_bar[i], 1);
        mbar_init(afull_bar, 1);
        mbar_init(aempty_bar, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&qstaged_bar[i], 4); mbar_init(&qflushed_bar[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kEpiWarps + 1) {
        if (kPair == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                         "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                         "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (kPair == 2) cluster_sync_all();         // both CTAs' barriers exist before the peer's copies / arrives reach them
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_holder, 0);
    // everything above ran while the previous kernel of the stream (the pose chain) was still draining; its
    // outputs (A' rows, A_j tiles, offsets) are visible from here on (no-op without a programmatic dependency)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tmem_base != 0) __trap();              // all 512 columns are ours: the allocation can only start at 0

    // contiguous unit range of this CTA (pair); unit u = frame tile (pair of frame tiles) u / 216, vertex tile u % 216.
    // CTA `crank` of a pair takes frame tile 2 * (u / 216) + crank; a tile beyond the batch computes on zero rows (the
    // tensor map fills them) and stores nothing
    const int crank = kPair == 2 ? (int)cluster_ctarank() : 0;
    const int64_t cid = blockIdx.x / kPair, ncl = gridDim.x / kPair;
    // Balanced by cost, not by count (prk_internal.h fused_first_unit): a frame-tile switch inside a range (drain the MMAs of the
    // old A' tile, load 112 KB of A' and 147 KB of A_j into shared / tensor memory) costs `switch_cost16` sixteenths of a unit,
    // and the ranges that contain one would otherwise finish last.
    const int64_t u0 = fused_first_unit(cid, ncl, n_units, switch_cost16);
    const int64_t u1 = fused_first_unit(cid + 1, ncl, n_units, switch_cost16);
    const int n_my = (int)(u1 - u0);
    const int64_t ft0 = (u0 / FUSED_NT) * kPair + crank;
    const int vt0 = (int)(u0 - (u0 / FUSED_NT) * FUSED_NT);

    if (warp == kEpiWarps) {
        // ===== TMA producer (whole warp converged, one elected lane issues) =====
        int stage = 0; uint32_t phase = 0;
        int64_t ft = ft0; int vt = vt0;
        uint32_t n_ft = 0;
        for (int i = 0; i < n_my; ++i) {
            if (i == 0 || vt == 0) {
                // the MMAs of the previous frame tile still read the resident A' tile
                if (n_ft > 0) MBAR_WAIT(aempty_bar, (n_ft - 1) & 1);
                if (elect_one()) {
                    if (kPair == 2) {       // both CTAs' tiles are counted on the leader's barrier (the leader issues the MMAs)
                        if (crank == 0) mbar_expect_tx(afull_bar, 2 * kABytes);
                        for (int c = 0; c < FUSED_A_CHUNKS; ++c)
                            tma_load_2d_pair(&tmap_A, afull_bar, sA + c * kAChunkBytes, c * 64, (int)(ft * FUSED_BM));
                    } else {
                        mbar_expect_tx(afull_bar, kABytes);
                        for (int c = 0; c < FUSED_A_CHUNKS; ++c)
                            tma_load_2d(&tmap_A, afull_bar, sA + c * kAChunkBytes, c * 64, (int)(ft * FUSED_BM));
                    }
                }
                ++n_ft;
            }
            // skinning weights of the tile's 32 vertices.  Slot i % kWS was last read by unit i - kWS.  kWS = 4: the ring is
            // shorter than one unit, so the chunk loads of unit i-1 already issued imply the MMA of unit i-1 has started,
            // i.e. every epilogue warp finished unit i-3.  CTA pairs: the ring holds 1.25 units, the MMA of unit i-2 has
            // started, every epilogue warp has passed the middle of unit i-4: hence kWS = 5 there (4 slots were a race).
            const uint32_t wbytes = (uint32_t)groups * FUSED_WGROUP_BYTES;
            if (elect_one()) {
                uint64_t* wb = &wfull_bar[i % kWS];
                mbar_expect_tx(wb, wbytes);
                bulk_load_1d(sW + (i % kWS) * wbytes, wpack + (size_t)vt * wbytes, wbytes, wb);
            }
#pragma unroll 1
            for (int c = 0; c < FUSED_B_CHUNKS; ++c) {
                MBAR_WAIT(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (kPair == 2) {
                        // this CTA's 48 rows of the chunk image (the swizzle phase of a row only depends on row & 7, and
                        // 48 is a multiple of 8); both halves are counted on the leader's barrier
                        if (crank == 0) mbar_expect_tx(&full_bar[stage], kBChunkBytes);
                        tma_load_2d_pair(&tmap_B, &full_bar[stage], sB + stage * kBSlotBytes, 0,
                                         (vt * FUSED_B_CHUNKS + c) * FUSED_BN + crank * (FUSED_BN / 2));
                    } else if (DBG(16)) { mbar_arrive(&full_bar[stage]); }
                    else {
                        // one contiguous 12 KB block of the pre-swizzled B' image (prk_internal.h fused_b2_index)
                        mbar_expect_tx(&full_bar[stage], kBChunkBytes);
                        bulk_load_1d(sB + stage * kBChunkBytes,
                                     B2img + ((size_t)vt * FUSED_B_CHUNKS + c) * (kBChunkBytes / 2), kBChunkBytes, &full_bar[stage]);
                    }
                }
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (++vt == FUSED_NT) { vt = 0; ft += kPair; }
        }
    } else if (warp == kEpiWarps + 1 && kPair == 2 && crank != 0) {
        // (the pair's MMAs are issued by the leader CTA)
    } else if (warp == kEpiWarps + 1) {
        // ===== MMA issuer (whole warp converged, one elected lane issues) =====
        // Everything about a k-step is a compile-time constant (both loops fully unrolled) and every
        // run-time operand is warp-uniform, so an MMA costs a couple of uniform adds plus the issue.
        constexpr uint32_t id16 = idesc_fp16(FUSED_BM * kPair, FUSED_BN), id8 = idesc_e4m3(FUSED_BM * kPair, FUSED_BN);
        const uint64_t adesc0 = make_smem_desc(smem_u32(sA));
        const uint32_t a_lo = (uint32_t)adesc0, d_hi = (uint32_t)(adesc0 >> 32);
        const uint32_t b_lo0 = (uint32_t)make_smem_desc(smem_u32(sB));
        int stage = 0; uint32_t phase = 0;
        int vt = vt0;
        uint32_t n_ft = 0;
        for (int i = 0; i < n_my; ++i) {
            if (i == 0 || vt == 0) {
                if (n_ft > 0 && elect_one()) { if (kPair == 2) tcgen05_commit_pair(aempty_bar); else tcgen05_commit(aempty_bar); }   // all MMAs on the old A' tile retire first
                MBAR_WAIT(afull_bar, n_ft & 1);
                tcgen05_fence_after();
                ++n_ft;
            }
            const int acc = i & 1;
            // the tile's skinning weights: waited for here (single waiter) so the epilogue warps,
            // which see them through their tfull barrier, never park on the TMA barrier
            MBAR_WAIT(&wfull_bar[i % kWS], (i / kWS) & 1);
            MBAR_WAIT(&tempty_bar[acc], ((i >> 1) & 1) ^ 1);       // epilogue drained this accumulator
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + kAccCol0 + (uint32_t)acc * FUSED_BN;
#pragma unroll
            for (int c = 0; c < FUSED_B_CHUNKS; ++c) {
                MBAR_WAIT(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t b_lo = b_lo0 + (uint32_t)stage * (kBSlotBytes >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int s = 0; s < (DBG(1) ? 0 : 4); ++s) {
                        const int b = c * 4 + s;                  // B' k-step (compile-time)
                        const uint32_t bl = b_lo + 2 * s;
                        if (b < FUSED_POSE_STEPS) {               // fp16 main part: Fh . Ph
                            umma_step<kPair, false>(d_tmem, a_lo + a_step_off(b), bl, d_hi, id16, b != 0);
                        } else if (b < 2 * FUSED_POSE_STEPS) {    // e4m3 cross terms: (F - Fh) . P + F . (P - Ph), K = 32
                            umma_step<kPair, true>(d_tmem, a_lo + a_step_off(b), bl, d_hi, id8, true);
                        } else {                                  // betas x shapedirs + template, fp16: bh.sh + T, bl.sh | bh.sl
                            constexpr int A26 = 2 * FUSED_POSE_STEPS, A27 = A26 + 1;
                            umma_step<kPair, false>(d_tmem, a_lo + a_step_off(A26), bl, d_hi, id16, true);
                            if (b == A26) umma_step<kPair, false>(d_tmem, a_lo + a_step_off(A27), bl, d_hi, id16, true);
                        }
                    }
                    if (kPair == 2) {
                        tcgen05_commit_pair(&empty_bar[stage]);   // frees the ring slot in both CTAs when the MMAs retire
                        if (c == FUSED_B_CHUNKS - 1) {
#pragma unroll
                            for (int w = 0; w < kEpiWarps; ++w) tcgen05_commit_pair(&tfull_bar[acc * kEpiWarps + w]);
                        }
                    } else {
                        tcgen05_commit(&empty_bar[stage]);        // frees the ring slot when the MMAs retire
                        if (c == FUSED_B_CHUNKS - 1) {
#pragma unroll
                            for (int w = 0; w < kEpiWarps; ++w) tcgen05_commit(&tfull_bar[acc * kEpiWarps + w]);
                        }
                    }
                }
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (++vt == FUSED_NT) vt = 0;
        }
    } else {
        // ===== epilogue: thread = frame (TMEM lane), loop over the warp's 8 vertices of each unit =====
        const int quarter = warp & 3, oct = warp >> 2;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t t_quarter = (uint32_t)(quarter * 32) << 16;   // (only used with a single column table)
        // Vertices leave through a staging tile shared by the four warps of a lane quarter:
        // [32 frames][16 vertices x 3 floats] per half unit, so every frame row is written to HBM as one
        // 192-byte run (the L1->L2 store path is paid per 128-byte line touched, not per byte).
        // In half h warp `oct` computes vertices 16h + 4oct .. +3 and stores frame rows 8oct .. 8oct+7.
        float* q_out = reinterpret_cast<float*>(sOut + quarter * kOutBytesPerQuarter);
        // read-back: float2 index q = it*32 + lane over [8 rows][24 float2]; iterations it+3 hit the same
        // column 4 rows further down
        int rb_smem[3], rb_glob[3], rb_row[3], rb_c2[3];
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int q = it * 32 + lane;
            rb_row[it] = oct * 8 + q / 24; rb_c2[it] = (q % 24) * 2;
            rb_smem[it] = rb_row[it] * kOutPitch + rb_c2[it];     // float index in the staging tile
            rb_glob[it] = rb_row[it] * NVC + rb_c2[it];           // float offset from the tile's first frame row
        }

        // Deferred read-back (kGroups == 1): the global stores of a staged half are issued in the middle of the NEXT
        // half's gather + math, so the store phase of a lane quarter runs under its tensor-memory gathers instead of
        // after them.  `pend` = first element of the pending half's frame rows (nullptr: nothing pending).  The tile
        // hand-over between the four warps uses two mbarriers per quarter ("staged", "flushed") whose arrive and wait
        // sit half a half apart, so a warp rarely blocks and the four warps drift apart instead of marching in lock-step.
        float* pend = nullptr;
        uint32_t n_staged = 0;                                        // halves this warp has staged so far
        uint8_t* const wtile = sOut + warp * kWarpTileBytes;          // kTma: this warp's private tile
        uint64_t store_policy = 0;
        if (kTma) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(store_policy));
        auto flush_pending = [&]() {
            if (kTma) return;                                         // nothing is deferred: the TMA unit does the stores
            if (pend == nullptr) return;                              // warp-uniform
            MBAR_WAIT(&qstaged_bar[quarter], (n_staged - 1) & 1);    // all four warps staged the pending half
#pragma unroll
            for (int it = 0; it < (DBG(256) ? 0 : 6); ++it) {
                const int j = it % 3, up = (it / 3) * 4;
                const float2 val = *reinterpret_cast<const float2*>(q_out + rb_smem[j] + up * kOutPitch);
                if (!DBG(4) || val.x == 123.456f) store_vertex_pair(pend + rb_glob[j] + up * NVC, val);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&qflushed_bar[quarter]);      // my rows of that half are in registers / on their way
            pend = nullptr;
        };
        int64_t ft = ft0; int vt = vt0;
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#ifdef PRK_FUSED_DEBUG
        long long t_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long t_begin = clock64();
#endif
        for (int i = 0; i < n_my; ++i) {
            if (i == 0 || vt == 0) {
                // ---- new frame tile: A_j of my 32 frames -> TMEM columns [72*oct, 72*oct+72) ----
                epi_barrier();                                    // nobody still gathers the old A_j
                const int64_t f = ft * FUSED_BM + quarter * 32 + lane;
                const float* src = AskinT + ((size_t)(ft * 4 + quarter) * FUSED_ASKIN_COLS + oct * 72) * 32 + lane;
                // 72 coalesced loads in PRK_AJ_BATCHES batches: enough loads in flight to cover the L2 latency
#ifndef PRK_AJ_BATCHES
#define PRK_AJ_BATCHES 3
#endif
                constexpr int kAjPer = 72 / PRK_AJ_BATCHES;
                static_assert(kAjPer * PRK_AJ_BATCHES == 72 && kAjPer % 8 == 0, "A_j batches of whole tcgen05.st.x8 groups");
#pragma unroll 1
                for (int b = 0; b < PRK_AJ_BATCHES; ++b) {
                    uint32_t v[kAjPer];
#pragma unroll
                    for (int k = 0; k < kAjPer; ++k)
                        v[k] = (kPair == 1 || ft * FUSED_BM < B) ? __float_as_uint(__ldg(src + (b * kAjPer + k) * 32)) : 0u;   // no A_j tile beyond the batch
#pragma unroll
                    for (int c = 0; c < kAjPer / 8; ++c) tmem_st_x8(t_lane + (uint32_t)(oct * 72 + b * kAjPer + c * 8), v + c * 8);
                }
                tmem_st_wait();
                if (f < B) { o0 = off[f * 3 + 0]; o1 = off[f * 3 + 1]; o2 = off[f * 3 + 2]; }
                tcgen05_fence_before();
                epi_barrier();
                tcgen05_fence_after();
            }
            const int acc = i & 1;
            const uint32_t wbytes = (uint32_t)groups * FUSED_WGROUP_BYTES;
            const uint8_t* wslot = sW + (i % kWS) * wbytes;
            TCLK(tw0);
            MBAR_WAIT(&tfull_bar[acc * kEpiWarps + warp], (i >> 1) & 1);
            tcgen05_fence_after();
            TCLK(tw1);
            MBAR_WAIT(&wfull_bar[i % kWS], (i / kWS) & 1);   // completed before the MMAs were issued: never parks
            TCLK(tw2);
            TACC(0, tw0, tw1); TACC(1, tw1, tw2);
            if (DBG(2)) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) { if (kPair == 2 && crank != 0) mbar_arrive_remote(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
                if (++vt == FUSED_NT) { vt = 0; ft += kPair; }
                continue;
            }
            // the warp's accumulator columns: 2 x 12 (vertices 4 oct.. of each half) or, kTma, 24 consecutive ones (vertices 8 oct..)
            const uint32_t t_acc = t_lane + kAccCol0 + (uint32_t)acc * FUSED_BN + (uint32_t)(oct * (kTma ? 24 : 12));
            const int c_unit = vt * FUSED_VT * 3;                   // first vertex coordinate of the tile
            float* vrow = verts + (size_t)(ft * FUSED_BM + quarter * 32) * NVC + c_unit;
            // recomputed where it is used (two integer ops) instead of living in a register across the unit body
            auto rows_valid = [&]() {
                const int left = (int)B - ((int)ft * FUSED_BM + quarter * 32);
                return left < 32 ? left : 32;
            };

            // transposed store of the quarter's 16 finished vertices (48 floats per frame) of half unit `half`
            auto store_half = [&](const float* res, int half) {
                if (DBG(128)) { if (res[0] == 123.456f) vrow[0] = res[1]; return; }
                TCLK(ts0);
                float4* dst = reinterpret_cast<float4*>(q_out + lane * kOutPitch + oct * 12);   // pitch 208 B: conflict-free float4
                dst[0] = make_float4(res[0], res[1], res[2], res[3]);
                dst[1] = make_float4(res[4], res[5], res[6], res[7]);
                dst[2] = make_float4(res[8], res[9], res[10], res[11]);
                quarter_barrier(quarter);                           // all four warps' columns are staged
                const int c_first = c_unit + half * 48;             // first vertex coordinate of this half
                float* vhalf = vrow + half * 48;
                if (rows_valid() == 32 && c_first + 48 <= NVC && !DBG(4)) {   // warp-uniform: all but the edge tiles
#pragma unroll
                    for (int it = 0; it < 6; ++it) {
                        const int j = it % 3, up = (it / 3) * 4;
                        store_vertex_pair(vhalf + rb_glob[j] + up * NVC,
                                          *reinterpret_cast<const float2*>(q_out + rb_smem[j] + up * kOutPitch));
                    }
                } else {
#pragma unroll
                    for (int it = 0; it < 6; ++it) {
                        const int j = it % 3, up = (it / 3) * 4;
                        const float2 val = *reinterpret_cast<const float2*>(q_out + rb_smem[j] + up * kOutPitch);
                        if (rb_row[j] + up < rows_valid() && c_first + rb_c2[j] < NVC && !DBG(4))
                            store_vertex_pair(vhalf + rb_glob[j] + up * NVC, val);
                    }
                }
                quarter_barrier(quarter);                           // tile may be overwritten by the next half
                TCLK(ts1);
                TACC(3, ts0, ts1);
            };
            // one joint's contribution: TMEM columns R00 R10 R01 R11 R02 R12 t0 t1 | R20 R21 R22 t2
            auto joint_math = [&](const uint32_t* a, uint64_t ww, uint64_t pxx, uint64_t pyy, uint64_t pzz, uint64_t pxy,
                                  uint64_t pz1, uint64_t& accxy, uint64_t& accz) {
                uint64_t xy = fma2(pack2(a[0], a[1]), pxx, pack2(a[6], a[7]));
                xy = fma2(pack2(a[2], a[3]), pyy, xy);
                xy = fma2(pack2(a[4], a[5]), pzz, xy);
                uint64_t zz = mul2(pack2(a[8], a[9]), pxy);     // (R20 px, R21 py)
                zz = fma2(pack2(a[10], a[11]), pz1, zz);        // (+ R22 pz, + t2): z = lo + hi
                accxy = fma2(ww, xy, accxy);
                accz = fma2(ww, zz, accz);
            };

            if (kGroups == 1) {
                // Software pipeline over the 32 (vertex, joint) items of the unit: the gather of item
                // n+1 is in flight while item n is multiplied, with two 12-register buffers -- TMEM
                // read bandwidth is the bound of this kernel, so it must never sit idle.
                // the warp's k-th vertex is tile vertex 16 (k >> 2) + 4 oct + (k & 3)
                const uint4* cols = reinterpret_cast<const uint4*>(wslot + 512 + (FUSED_WCOL_COPIES > 1 ? quarter * 512 : 0)) + oct * (kTma ? 8 : 4);
                const float4* wgt = reinterpret_cast<const float4*>(wslot) + oct * (kTma ? 8 : 4);
                constexpr int kTileV[8] = {0, 1, 2, 3, kTma ? 4 : 16, kTma ? 5 : 17, kTma ? 6 : 18, kTma ? 7 : 19};
                uint32_t buf[2][12], p[12];
                uint4 cj = DBG(64) ? make_uint4(12, 36, 120, 240) : cols[0];
                TCLK(tg0);
                // v_posed of the warp's 4 vertices of a half (12 accumulator columns) at a time: 12 registers instead
                // of 24 (the kernel sits at the 96-register cap of an 18-warp CTA).  The accumulator goes back to the
                // MMA warp once the second half's columns are in registers; the other buffer serves unit i+1 meanwhile.
                tmem_ld_x8(t_acc, p); tmem_ld_x4(t_acc + 8, p + 8);
                tmem_ld_x8(GATHER_ADDR(cj.x), buf[0]); tmem_ld_x4(GATHER_ADDR(cj.x) + 8, buf[0] + 8);
                tmem_ld_wait();
                float res[12];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if ((k & 3) == 1) flush_pending();              // the previous half leaves under this half's gathers
                    const float4 w4 = DBG(64) ? make_float4(.25f, .25f, .25f, .25f) : wgt[kTileV[k]];
                    const uint64_t ww[4] = {pack2f(w4.x, w4.x), pack2f(w4.y, w4.y), pack2f(w4.z, w4.z), pack2f(w4.w, w4.w)};
                    const uint4 cj_next = DBG(64) ? make_uint4(24, 48, 132, 252) : cols[kTileV[k < 7 ? k + 1 : 7]];
                    uint64_t accxy = pack2f(o0, o1);
                    uint64_t pxx = 0, pyy = 0, pzz = 0, pxy = 0, pz1 = 0;
                    uint64_t tz01 = 0, tz23 = pack2f(0.f, o2);       // sum_j w_j (R20, R21) and (R22, t2) [+ offset z]
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int n = k * 4 + q;
                        if (n > 0) tmem_ld_wait();                    // item n has landed
                        if (n == 16) {                                // ... and so have the second half's v_posed columns
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) { if (kPair == 2 && crank != 0) mbar_arrive_remote(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
                        }
                        if (q == 0) {
                            const uint32_t* pk = p + (k & 3) * 3;
                            pxx = pack2(pk[0], pk[0]); pyy = pack2(pk[1], pk[1]); pzz = pack2(pk[2], pk[2]);
                            pxy = pack2(pk[0], pk[1]); pz1 = pack2(pk[2], 0x3f800000u);
                        }
                        uint32_t* nb = buf[(n + 1) & 1];
                        if (q < 3) {
                            const uint32_t col = q == 0 ? cj.y : (q == 1 ? cj.z : cj.w);
                            if (!DBG(8)) { tmem_ld_x8(GATHER_ADDR(col), nb); tmem_ld_x4(GATHER_ADDR(col) + 8, nb + 8); }
                        } else if (k < 7) {
                            if (!DBG(8)) { tmem_ld_x8(GATHER_ADDR(cj_next.x), nb); tmem_ld_x4(GATHER_ADDR(cj_next.x) + 8, nb + 8); }
                        }
                        if (DBG(32)) {
                            const uint32_t* a = buf[n & 1];
                            accxy ^= pack2(a[0] ^ a[11], a[5]);
                        } else {   // x, y per joint; the z row is blended first and applied once per vertex:
                            // 6 packed FMAs per item, every operand straight out of the gather registers
                            const uint32_t* a = buf[n & 1];
                            uint64_t xy = fma2(pack2(a[0], a[1]), pxx, pack2(a[6], a[7]));
                            xy = fma2(pack2(a[2], a[3]), pyy, xy);
                            xy = fma2(pack2(a[4], a[5]), pzz, xy);
                            accxy = fma2(ww[q], xy, accxy);
                            tz01 = q == 0 ? mul2(ww[q], pack2(a[8], a[9])) : fma2(ww[q], pack2(a[8], a[9]), tz01);
                            tz23 = fma2(ww[q], pack2(a[10], a[11]), tz23);
                        }
                    }
                    cj = cj_next;
                    const uint64_t accz = fma2(tz23, pz1, mul2(tz01, pxy));   // (T20 px + T22 pz, T21 py + t2): z = lo + hi
                    float zl, zh;
                    unpack2(accxy, res[(k & 3) * 3 + 0], res[(k & 3) * 3 + 1]);
                    unpack2(accz, zl, zh);
                    res[(k & 3) * 3 + 2] = zl + zh;
                    if ((k & 3) == 3) {
                        TCLK(tg1);
                        TACC(2, tg0, tg1);
                        const int half = k >> 2;
                        if (k == 3) { tmem_ld_x8(t_acc + (kTma ? 12 : 48), p); tmem_ld_x4(t_acc + (kTma ? 20 : 56), p + 8); }   // next half's v_posed
                        if (DBG(128)) {                              // knock-out: no staging, no read-back, no stores
                            float sum = 0.f;                         // every result stays live: nothing of the math may be dropped
#pragma unroll
                            for (int r = 0; r < 12; ++r) sum += res[r];
                            if (sum == 123.456f) vrow[half] = sum;
                            continue;
                        }
                        if (kTma) {
                            if (half == 0) {    // the previous unit's box has been read out of the tile (issued a unit ago)
                                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                                __syncwarp();
                            }
                            float4* dst = reinterpret_cast<float4*>(wtile + lane * 96 + half * 48);
                            dst[0] = make_float4(res[0], res[1], res[2], res[3]);
                            dst[1] = make_float4(res[4], res[5], res[6], res[7]);
                            dst[2] = make_float4(res[8], res[9], res[10], res[11]);
                            if (half == 1) {
                                fence_proxy_async_smem();               // generic-proxy writes -> visible to the async proxy
                                __syncwarp();
                                if (lane == 0 && !DBG(4)) {
                                    tma_store_box(&tmap_V, wtile, c_unit + oct * 24, (int)ft * FUSED_BM + quarter * 32, store_policy);
                                    tma_store_commit();
                                }
                            }
                            continue;
                        }
                        if (n_staged > 0) MBAR_WAIT(&qflushed_bar[quarter], (n_staged - 1) & 1);   // tile is free
                        float4* dst = reinterpret_cast<float4*>(q_out + lane * kOutPitch + oct * 12);
                        dst[0] = make_float4(res[0], res[1], res[2], res[3]);
                        dst[1] = make_float4(res[4], res[5], res[6], res[7]);
                        dst[2] = make_float4(res[8], res[9], res[10], res[11]);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&qstaged_bar[quarter]);
                        ++n_staged;
                        const int c_first = c_unit + half * 48;
                        float* vhalf = vrow + half * 48;
                        if (rows_valid() == 32 && c_first + 48 <= NVC) {
                            pend = vhalf;                           // interior tile: stored during the next half
                        } else {                                    // edge tile: predicated stores right away
                            MBAR_WAIT(&qstaged_bar[quarter], (n_staged - 1) & 1);
#pragma unroll
                            for (int it = 0; it < 6; ++it) {
                                const int j = it % 3, up = (it / 3) * 4;
                                const float2 val = *reinterpret_cast<const float2*>(q_out + rb_smem[j] + up * kOutPitch);
                                if (rb_row[j] + up < rows_valid() && c_first + rb_c2[j] < NVC)
                                    store_vertex_pair(vhalf + rb_glob[j] + up * NVC, val);
                            }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&qflushed_bar[quarter]);
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    float res[12];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int vl = half * 16 + oct * 4 + k;       // vertex within the tile
                        uint32_t p[4];
                        tmem_ld_x4(t_acc + (uint32_t)((half * 16 + k) * 3), p);
                        uint64_t accxy = pack2f(o0, o1), accz = pack2f(o2, 0.f);
                        uint64_t pxx = 0, pyy = 0, pzz = 0, pxy = 0, pz1 = 0;
#pragma unroll 1
                        for (int g = 0; g < groups; ++g) {
                            const uint8_t* wg = wslot + g * FUSED_WGROUP_BYTES;
                            const uint4 cj = reinterpret_cast<const uint4*>(wg + 512 + (FUSED_WCOL_COPIES > 1 ? quarter * 512 : 0))[vl];   // 12 * joint, 4 joints
                            uint32_t r[48];
                            tmem_ld_x8(GATHER_ADDR(cj.x), r +  0); tmem_ld_x4(GATHER_ADDR(cj.x) + 8, r +  8);
                            tmem_ld_x8(GATHER_ADDR(cj.y), r + 12); tmem_ld_x4(GATHER_ADDR(cj.y) + 8, r + 20);
                            tmem_ld_x8(GATHER_ADDR(cj.z), r + 24); tmem_ld_x4(GATHER_ADDR(cj.z) + 8, r + 32);
                            tmem_ld_x8(GATHER_ADDR(cj.w), r + 36); tmem_ld_x4(GATHER_ADDR(cj.w) + 8, r + 44);
                            const float4 w4 = reinterpret_cast<const float4*>(wg)[vl];
                            tmem_ld_wait();
                            if (g == 0) {
                                pxx = pack2(p[0], p[0]); pyy = pack2(p[1], p[1]); pzz = pack2(p[2], p[2]);
                                pxy = pack2(p[0], p[1]); pz1 = pack2(p[2], 0x3f800000u);
                            }
                            const uint64_t ww[4] = {pack2f(w4.x, w4.x), pack2f(w4.y, w4.y), pack2f(w4.z, w4.z), pack2f(w4.w, w4.w)};
#pragma unroll
                            for (int q = 0; q < 4; ++q) joint_math(r + q * 12, ww[q], pxx, pyy, pzz, pxy, pz1, accxy, accz);
                        }
                        float zl, zh;
                        unpack2(accxy, res[k * 3 + 0], res[k * 3 + 1]);
                        unpack2(accz, zl, zh);
                        res[k * 3 + 2] = zl + zh;
                    }
                    if (half == 1) {
                        // both halves' accumulator columns are in registers: hand the buffer back to the MMA warp
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (kPair == 2 && crank != 0) mbar_arrive_remote(&tempty_bar[acc], 0); else mbar_arrive(&tempty_bar[acc]); }
                    }
                    store_half(res, half);
                }
            }
            if (++vt == FUSED_NT) { vt = 0; ft += kPair; }
        }
        flush_pending();
        if (kTma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores done before the CTA exits
#ifdef PRK_FUSED_DEBUG
        if (lane == 0) {
            t_sum[4] = clock64() - t_begin;
            for (int k = 0; k < 5; ++k) atomicAdd(&g_fdbg[k], (unsigned long long)t_sum[k]);
            atomicAdd(&g_fdbg[7], 1ull);
        }
#endif
    }

    tcgen05_fence_before();
    __syncthreads();
    if (kPair == 2) cluster_sync_all();         // neither CTA leaves (or frees tensor memory) while the pair's MMAs, copies or arrives may still touch it
    if (warp == kEpiWarps + 1) {
        tcgen05_fence_after();
        if (kPair == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- verification only (prk_debug_blend) ------------------------------------------------
// The same operand bytes and the same 31 k-step products as the tensor-core path (prk_internal.h "K12 operand layout"), plain FFMA.
__global__ void __launch_bounds__(256)
blend_simt_kernel(const uint16_t* __restrict__ Arows, const uint16_t* __restrict__ B2, int64_t rows,
                  float* __restrict__ vposed, float rot_scale) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;   // vertex coordinate
    const int64_t f = blockIdx.y;
    if (n >= NVC || f >= rows) return;
    const uint16_t* a = Arows + f * FUSED_K;
    const uint8_t* ab = reinterpret_cast<const uint8_t*>(a);
    const uint8_t* bb = reinterpret_cast<const uint8_t*>(B2);
    auto hf = [](uint16_t v) { return __half2float(__ushort_as_half(v)); };
    auto f8 = [](uint8_t v) { return __half2float(__half(__nv_cvt_fp8_to_halfraw(v, __NV_E4M3))); };
    float acc = 0.f;
    for (int k = 0; k < 16 * FUSED_POSE_STEPS; ++k)                   // fp16 main part
        acc = fmaf(hf(a[k]), hf(B2[fused_b2_index(n, k)]), acc);
    for (int k = FUSED_X_BYTE0; k < FUSED_X_BYTE0 + 2 * FUSED_COL_LO; ++k)   // e4m3 cross terms, byte meets byte
        acc = fmaf(f8(ab[k]), f8(bb[fused_b2_byte_index(n, k)]), acc);
    auto step = [&](int sa, int sb, float acc) {                      // fp16 beta / template products
        for (int k = 0; k < 16; ++k) acc = fmaf(hf(a[sa * 16 + k]), hf(B2[fused_b2_index(n, sb * 16 + k)]), acc);
        return acc;
    };
    const int A26 = 2 * FUSED_POSE_STEPS, A27 = A26 + 1, B26 = A26;
    acc = step(A26, B26, acc); acc = step(A27, B26, acc); acc = step(A26, B26 + 1, acc);
    vposed[f * NVC + n] = acc * rot_scale;
}

// A_j = 2^-S identity for every joint (columns R00 = 0, R11 = 3, R22 = 10 of each group of 12), off = 0
__global__ void __launch_bounds__(256)
identity_askin_kernel(float* __restrict__ AskinT, float* __restrict__ off, int64_t rows_pad, float rot_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows_pad * FUSED_ASKIN_COLS) {
        const int col = (int)((i >> 5) % FUSED_ASKIN_COLS) % 12;
        AskinT[i] = (col == 0 || col == 3 || col == 10) ? rot_scale : 0.0f;
    }
    if (i < rows_pad * 3) off[i] = 0.0f;
}

}  // namespace

cudaError_t launch_blend_simt(const Model& m, const uint16_t* d_Arows, int64_t rows, float* d_vposed, cudaStream_t s) {
    if (rows == 0) return cudaSuccess;
    dim3 grid((NVC + 255) / 256, (unsigned)rows);
    blend_simt_kernel<<<grid, 256, 0, s>>>(d_Arows, m.d_B2, rows, d_vposed, m.pc.rot_scale);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_identity_askin(const Model& m, float* d_AskinT, float* d_off, int64_t rows_pad, cudaStream_t s) {
    const int64_t n = rows_pad * FUSED_ASKIN_COLS;
    identity_askin_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_AskinT, d_off, rows_pad, m.pc.rot_scale);
    count_launch();
    return cudaGetLastError();
}

#ifdef PRK_FUSED_DEBUG
extern "C" __attribute__((visibility("default"))) int prk_fused_debug_read(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_fdbg, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_fdbg, z, sizeof z); }
    return 0;
}
#endif

static thread_local bool t_exchange_hint = false;   // this thread's current call issues a multi-GPU exchange beside the vertex kernel
int fused_stages(int groups, int pair, bool tma) {
#ifndef PRK_PAIR_STAGES
#define PRK_PAIR_STAGES 12
#endif
    int stages = pair == 2 ? PRK_PAIR_STAGES : 6;   // 12 half chunks (72 KB) + 5 weight slots, or 6 chunks (72 KB) + 4 slots
#ifndef PRK_TMA_STAGES
#define PRK_TMA_STAGES 10
#endif
    if (tma) stages = pair == 2 ? PRK_TMA_STAGES : 4;   // the private tiles take 22.5 KB more
#ifdef PRK_STAGE_CAP
    stages = PRK_STAGE_CAP * pair;
#endif
    while (stages > 2 && fused_smem_bytes(stages, groups, pair, tma) > kSmemLimit) --stages;
    // One ring slot fewer when the call takes part in a multi-GPU exchange (set_fused_exchange_hint; PRK_STAGE_ROOM=n overrides):
    // a CTA with all 10 slots owns the SM's whole shared memory, so no block of another kernel -- the exchange's push kernel,
    // the next call's scoring -- can run beside it and everything queues up at the kernel boundary.  Measured on the host path:
    // +11 us per step with the exchange and all slots, +4 us with 9 slots; 8 GPUs end to end 162 - 177 -> 198 M frames/s
    // together with the wait for the peers moved off the scoring stream (prk_comm.cu).
    static const int room_env = [] { const char* e = getenv("PRK_STAGE_ROOM"); return e ? atoi(e) : -1; }();
    const int room = room_env >= 0 ? room_env : (t_exchange_hint ? 1 : 0);
    if (room > 0 && stages - room >= 2) stages -= room;
    return stages;
}
void set_fused_exchange_hint(bool on) { t_exchange_hint = on; }

// cost of a frame-tile switch inside a CTA's unit range, in sixteenths of a unit (see the kernel's range computation)
#ifndef PRK_SWITCH_COST16_DEFAULT
#define PRK_SWITCH_COST16_DEFAULT 48     // measured (profiles/r2_ab_runs.txt ab_r2_33): 0 / 16 / 32 / 48 / 64 -> 106.8 / 105.5 / 104.5 / 104.3 / 104.2 us
#endif
// CTA pairs: PRK_PAIR=1|2 overrides; default PRK_PAIR_DEFAULT when the batch has at least two frame tiles
#ifndef PRK_PAIR_DEFAULT
#define PRK_PAIR_DEFAULT 2
#endif
static int fused_pair(int64_t n_ft) {
    static const int forced = [] { const char* e = getenv("PRK_PAIR"); return e ? atoi(e) : 0; }();
    const int p = forced > 0 ? forced : PRK_PAIR_DEFAULT;
    return (p == 2 && n_ft >= 2) ? 2 : 1;
}

template <int kGroups, int kPair, bool kTma>
static cudaError_t launch_fused_t(const Model& m, const CUtensorMap& tmap_A, int64_t rows_pad, const float* d_AskinT,
                                  const float* d_off, int64_t B, float* d_verts, int64_t vpitch, cudaStream_t s, bool pdl) {
    auto kern = fused_blend_skin_kernel<kGroups, kPair, kTma>;
    const int groups = m.nnz_groups;
    const int stages = fused_stages(groups, kPair, kTma);
    const int smem = fused_smem_bytes(stages, groups, kPair, kTma);
    if (smem > kSmemLimit) return cudaErrorInvalidConfiguration;
    static std::atomic<int> attr_set[64];       // per device: dynamic shared memory this instantiation was opted in for
    if (m.device >= 0 && m.device < 64 && attr_set[m.device].load(std::memory_order_acquire) < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_set[m.device].store(smem, std::memory_order_release);
    }
    const int64_t n_ft = rows_pad / FUSED_BM;
    const int64_t n_units = ((n_ft + kPair - 1) / kPair) * FUSED_NT;          // (pair) units
    int grid = m.sm_count > 0 ? m.sm_count : 148;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (kPair == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
        static std::atomic<int> max_clusters[64];     // one CTA per SM: as many pairs as the GPU holds at once
        int mc = (m.device >= 0 && m.device < 64) ? max_clusters[m.device].load(std::memory_order_acquire) : 0;
        if (mc == 0) {
            cfg.gridDim = dim3((unsigned)(grid / 2 * 2));
            cfg.attrs = attr; cfg.numAttrs = na;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&mc, kern, &cfg);
            if (e != cudaSuccess) return e;
            if (mc < 1) return cudaErrorInvalidConfiguration;
            if (m.device >= 0 && m.device < 64) max_clusters[m.device].store(mc, std::memory_order_release);
        }
        if (grid > mc * 2) grid = mc * 2;
        grid = grid / 2 * 2;
        if ((int64_t)grid / 2 > n_units) grid = (int)n_units * 2;
    } else if (grid > n_units) {
        grid = (int)n_units;
    }
    // Programmatic dependent launch: the CTAs become resident and run their prologue (barrier init, tensor-memory
    // allocation, descriptor prefetch) while the pose-chain kernel in front of them drains; every thread passes
    // griddepcontrol.wait before it touches that kernel's outputs.  PRK_PDL=0 launches without the attribute.
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.gridDim = dim3((unsigned)grid);
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const uint8_t* wpack = m.d_wpack;
    const uint16_t* b2img = m.d_B2;
    CUtensorMap tmV = m.tmap_B2;                // (a valid placeholder for the instantiations that store with STG)
    if (kTma) {                                 // [B][vpitch] fp32, boxes of 24 floats x 32 frames
        const int rc = encode_tmap_2d_f32_strided(&tmV, d_verts, (uint64_t)B, (uint64_t)NVC, (uint64_t)vpitch * 4, 32, 24);
        if (rc != PRK_OK) return cudaErrorInvalidValue;
    }
    static const int switch_cost16 = [] { const char* e = getenv("PRK_SWITCH_COST16"); return e ? atoi(e) : PRK_SWITCH_COST16_DEFAULT; }();
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmap_A, m.tmap_B2, tmV, b2img, d_AskinT, d_off, wpack, groups, stages, B, n_units,
                                       d_verts, 0, n_units >= 8 * (int64_t)(grid / kPair) ? switch_cost16 : 0);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_fused(const Model& m, const CUtensorMap& tmap_A, int64_t rows_pad, const float* d_AskinT,
                         const float* d_off, int64_t B, float* d_verts, int64_t vpitch, cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    static const bool pdl = [] { const char* e = getenv("PRK_PDL"); return !e || atoi(e) != 0; }();
    static const bool tma_ok = [] { const char* e = getenv("PRK_TMA_STORE"); return !e || atoi(e) != 0; }();
    const int pair = fused_pair(rows_pad / FUSED_BM);
    // TMA vertex stores need 16-byte aligned rows: base and pitch; the dense reference layout (pitch 20,670) has neither
    const bool tma = tma_ok && m.nnz_groups == 1 && (vpitch % 4) == 0 && (reinterpret_cast<uintptr_t>(d_verts) & 15) == 0;
    if (vpitch != NVC && !tma) return cudaErrorNotSupported;      // only the TMA path knows about padded rows (PRK_TMA_STORE=0 with a pitch)
#define PRK_GO(G, P, T) launch_fused_t<G, P, T>(m, tmap_A, rows_pad, d_AskinT, d_off, B, d_verts, vpitch, s, pdl)
    if (m.nnz_groups == 1) {
        if (tma) return pair == 2 ? PRK_GO(1, 2, true) : PRK_GO(1, 1, true);
        return pair == 2 ? PRK_GO(1, 2, false) : PRK_GO(1, 1, false);
    }
    return pair == 2 ? PRK_GO(0, 2, false) : PRK_GO(0, 1, false);
#undef PRK_GO
}

}  // namespace prk
