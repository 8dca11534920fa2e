// K2a: per-frame batch Rodrigues + 24-joint kinematic chain, one thread per frame.
//
// Reference behaviour (paths relative to the reference root):
//   th_posemap_axisang / batch_rodrigues / quat2mat   lib/smplpytorch/smplpytorch/pytorch/tensutils.py:6-19,
//                                                     rodrigues_layer.py:13-52
//   subtract_flat_id (pose_map = R - I)               tensutils.py:41-48
//   th_j = J_regressor @ v_shaped                     smpl_layer.py:87-95   (folded: J = J0 + Jdirs*beta)
//   kinematic chain, rest-pose removal, joints        smpl_layer.py:103-132,145
//   centring / translation                            smpl_layer.py:148-155
//
// Outputs per frame:
//   joints [24][3]                                     (always)
//   A' row [448 of 512] bf16: split-precision operand of the fused kernel's blend GEMM, K12 layout
//                      (prk_internal.h)                                     (full mesh only)
//   AskinT [frame/32][288][frame%32] fp32: A_j = [R_g | t_g - R_g j_rest] in the column order
//                      the fused kernel keeps in tensor memory              (full mesh only)
//   off    [3]: +trans or -centre joint, applied to vertices by K12        (full mesh only)
//
// The chain is walked depth-first with every index a compile-time constant, so the at
// most three live 3x4 transforms stay in registers; HBM traffic is 288 B pose (+40 betas,
// +12 trans) in and 288 B joints (+896 A' +1152 A_j +12 off) out per frame.
#include "prk_internal.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cstdlib>

namespace prk {

// SMPL kintree parents (smpl_layer.py:60-62) and a depth-first visiting order.
__host__ __device__ constexpr int smpl_parent(int j) {
    constexpr int t[24] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
    return t[j];
}
__host__ __device__ constexpr int smpl_dfs(int pos) {
    constexpr int t[24] = {0, 1, 4, 7, 10, 2, 5, 8, 11, 3, 6, 9, 12, 15, 13, 16, 18, 20, 22, 14, 17, 19, 21, 23};
    return t[pos];
}

// the three parts of a pose feature F in the K12 operand row (prk_internal.h): fp16(F), e4m3((F - fp16(F)) 2^12), e4m3(F)
__device__ __forceinline__ void feature_parts(float v, uint16_t& hi, uint8_t& lo8, uint8_t& hi8) {
    const __half h = __float2half_rn(v);
    hi = __half_as_ushort(h);
    lo8 = (uint8_t)__nv_cvt_float_to_fp8((v - __half2float(h)) * 4096.0f, __NV_SATFINITE, __NV_E4M3);
    hi8 = (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
}

// Streams bf16 values into a global row, 8 at a time (one 16-byte store).
struct RowWriter {
    uint4* dst;
    uint32_t buf[4];
    int n;
    bool live;      // false: lanes beyond the batch run the same code and store nothing
    __device__ __forceinline__ void push(uint16_t v) {
        const int w = (n >> 1) & 3;
        if (n & 1) buf[w] |= (uint32_t)v << 16; else buf[w] = v;
        ++n;
        if ((n & 7) == 0) { if (live) *dst = make_uint4(buf[0], buf[1], buf[2], buf[3]); ++dst; }
    }
    // the same stream in bytes (the e4m3 blocks): n counts 16-bit slots, nb the bytes of the slot being filled
    int nb;
    uint32_t half_word;
    __device__ __forceinline__ void push8(uint8_t v) {
        if (nb == 0) { half_word = v; nb = 1; }
        else { nb = 0; push((uint16_t)(half_word | ((uint32_t)v << 8))); }
    }
};

// batch_rodrigues + quat2mat in fp32 (rodrigues_layer.py:41-52, 13-38), row-major R[9].
__device__ __forceinline__ void smpl_rodrigues(float ax, float ay, float az, float* R) {
    // every product/sum is an explicit _rn op or fmaf: identical bits in every kernel variant
    const float x = __fadd_rn(ax, 1e-8f), y = __fadd_rn(ay, 1e-8f), z = __fadd_rn(az, 1e-8f);
    const float angle = sqrtf(fmaf(z, z, fmaf(y, y, __fmul_rn(x, x))));
    const float nx = __fdiv_rn(ax, angle), ny = __fdiv_rn(ay, angle), nz = __fdiv_rn(az, angle);   // :45
    float s, c;
    sincosf(__fmul_rn(angle, 0.5f), &s, &c);
    float qw = c, qx = __fmul_rn(s, nx), qy = __fmul_rn(s, ny), qz = __fmul_rn(s, nz);
    const float qn = sqrtf(fmaf(qz, qz, fmaf(qy, qy, fmaf(qx, qx, __fmul_rn(qw, qw)))));
    // (measured: one reciprocal + multiplies instead of these seven divisions saves 13 % of the joints-only kernel and doubles
    // the joint error against the float64 CPU restatement, 2.1e-7 -> 4.2e-7: not taken)
    qw = __fdiv_rn(qw, qn); qx = __fdiv_rn(qx, qn); qy = __fdiv_rn(qy, qn); qz = __fdiv_rn(qz, qn);           // :21
    const float w2 = __fmul_rn(qw, qw), x2 = __fmul_rn(qx, qx), y2 = __fmul_rn(qy, qy), z2 = __fmul_rn(qz, qz);
    const float wx = __fmul_rn(qw, qx), wy = __fmul_rn(qw, qy), wz = __fmul_rn(qw, qz);
    const float xy = __fmul_rn(qx, qy), xz = __fmul_rn(qx, qz), yz = __fmul_rn(qy, qz);
    R[0] = __fsub_rn(__fsub_rn(__fadd_rn(w2, x2), y2), z2);
    R[1] = __fsub_rn(__fmul_rn(2.f, xy), __fmul_rn(2.f, wz));
    R[2] = __fadd_rn(__fmul_rn(2.f, wy), __fmul_rn(2.f, xz));
    R[3] = __fadd_rn(__fmul_rn(2.f, wz), __fmul_rn(2.f, xy));
    R[4] = __fsub_rn(__fadd_rn(__fsub_rn(w2, x2), y2), z2);
    R[5] = __fsub_rn(__fmul_rn(2.f, yz), __fmul_rn(2.f, wx));
    R[6] = __fsub_rn(__fmul_rn(2.f, xz), __fmul_rn(2.f, wy));
    R[7] = __fadd_rn(__fmul_rn(2.f, wx), __fmul_rn(2.f, yz));
    R[8] = __fadd_rn(__fsub_rn(__fsub_rn(w2, x2), y2), z2);
}

// One out-of-line copy for the thread-per-frame kernel, which evaluates 24 joints per thread with compile-time joint indices:
// inlined 24 times (sincosf with its large-argument path, seven IEEE divisions with theirs) the joints-only kernel was 13,900
// instructions = 222 KB of code and ran out of the instruction caches.
__device__ __noinline__ void smpl_rodrigues_call(float ax, float ay, float az, float* R) { smpl_rodrigues(ax, ay, az, R); }

// Shared arithmetic of both kernel variants (explicit fmaf / _rn ops so the thread-per-frame
// and the lane-per-joint kernels produce bit-identical results).
__device__ __forceinline__ float rest_joint(const float jt, const float* __restrict__ jd, const float* beta) {
    float acc = jt;                                   // J = J_template + Jdirs * beta
#pragma unroll
    for (int k = 0; k < NBETA; ++k) acc = fmaf(jd[k], beta[k], acc);
    return acc;
}
__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a1, b1, __fmul_rn(a0, b0)));
}
// G = Gp * [R | t]   (smpl_layer.py:109-119), rows of 4
__device__ __forceinline__ void compose(const float* Gp, const float* R, float t0, float t1, float t2, float* G) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float g0 = Gp[r * 4 + 0], g1 = Gp[r * 4 + 1], g2 = Gp[r * 4 + 2];
        G[r * 4 + 0] = dot3(g0, g1, g2, R[0], R[3], R[6]);
        G[r * 4 + 1] = dot3(g0, g1, g2, R[1], R[4], R[7]);
        G[r * 4 + 2] = dot3(g0, g1, g2, R[2], R[5], R[8]);
        G[r * 4 + 3] = __fadd_rn(dot3(g0, g1, g2, t0, t1, t2), Gp[r * 4 + 3]);
    }
}
// translation column of A_j = G_j - pack(G_j @ [j_rest; 0])  (smpl_layer.py:126-132)
__device__ __forceinline__ float skin_t(const float* G, int r, float j0, float j1, float j2) {
    return __fsub_rn(G[r * 4 + 3], dot3(G[r * 4 + 0], G[r * 4 + 1], G[r * 4 + 2], j0, j1, j2));
}

// K12 TMEM column of entry (row r, column c) of a 3x4 skinning transform: rows 0 and 1 are interleaved
// so the fused kernel's packed FFMA2 sees (R0c, R1c) register pairs:  R00 R10 R01 R11 R02 R12 t0 t1 | R20 R21 R22 t2
__host__ __device__ constexpr int askin_col(int r, int c) { return r == 2 ? 8 + c : 2 * c + r; }

constexpr int kChainPitch = 73;    // floats per staged pose / joint row of the thread-per-frame kernel (72 + 1)

// mode bits
constexpr uint32_t MODE_FRAME_BETAS_ALWAYS = 1u;  // betas given and model betas are zero
constexpr uint32_t MODE_FRAME_BETAS_FLAG = 2u;    // betas given, honour flags->betas_nonzero
constexpr uint32_t MODE_TRANS_ALWAYS = 4u;        // trans given, no centre joint configured
constexpr uint32_t MODE_TRANS_FLAG = 8u;          // trans given and centre joint configured

#ifndef PRK_CHAIN_MINBLOCKS
#define PRK_CHAIN_MINBLOCKS 6      // 80 registers, 6 x 4 warps per SM: 7 % faster than 4 x 4 at 128 (measured, 1M frames)
#endif
template <bool kStd, bool kMesh>
__global__ void __launch_bounds__(128, PRK_CHAIN_MINBLOCKS)
pose_chain_kernel(const __grid_constant__ PoseConsts pc, const float* __restrict__ pose,
                  const float* __restrict__ betas, const float* __restrict__ trans,
                  const BatchFlags* __restrict__ flags, uint32_t mode, int center_idx, int64_t B,
                  uint16_t* __restrict__ Arows, float* __restrict__ Askin,
                  float* __restrict__ off, float* __restrict__ joints) {
    // Memory path: a frame's pose row is 288 B, so thread-private row accesses would touch 32 different lines per
    // warp-wide load / store.  Every warp copies the 32 consecutive pose rows of its frames into a shared-memory
    // tile with fully coalesced loads (odd pitch: conflict-free per-thread row accesses), the joints are written
    // into the same rows (a joint's pose entries are read before its position is written) and leave the same way.
    __shared__ float s_tile[4][32 * kChainPitch];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t f0 = (int64_t)blockIdx.x * blockDim.x + warp * 32;     // first frame of this warp
    if (f0 >= B) return;
    const int nf = (int)((B - f0) < 32 ? (B - f0) : 32);                 // frames of this warp (warp-uniform)
    float* tile = s_tile[warp];
    {
        const float* src = pose + f0 * 72;
        const int n = nf * 72;
        int r = 0, c = lane;
        for (int q = lane; q < 32 * 72; q += 32) {                       // rows beyond the batch: zeros
            tile[r * kChainPitch + c] = q < n ? src[q] : 0.0f;
            c += 32;
            if (c >= 72) { c -= 72; ++r; }
        }
    }
    __syncwarp();
    const bool live = lane < nf;
    const int64_t f = live ? f0 + lane : f0;      // lanes beyond the batch read frame f0's betas / trans and store nothing

    bool frame_betas = (mode & MODE_FRAME_BETAS_ALWAYS) != 0;
    bool add_trans = (mode & MODE_TRANS_ALWAYS) != 0;
    if (mode & MODE_FRAME_BETAS_FLAG) frame_betas = flags->betas_nonzero != 0;
    if (mode & MODE_TRANS_FLAG) add_trans = flags->trans_nonzero != 0;
    const bool centre = (center_idx >= 0) && !add_trans;

    float beta[NBETA];
#pragma unroll
    for (int k = 0; k < NBETA; ++k) beta[k] = frame_betas ? betas[f * NBETA + k] : pc.model_betas[k];

    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    if (add_trans) { o0 = trans[f * 3 + 0]; o1 = trans[f * 3 + 1]; o2 = trans[f * 3 + 2]; }

    // K12 operand row: a hi stream (column 0) and a lo stream (column 208)
    RowWriter rw, rw_lo;
    rw.dst = kMesh ? reinterpret_cast<uint4*>(Arows + f * FUSED_K) : nullptr;
    rw.n = 0; rw.nb = 0; rw.half_word = 0;
    rw.live = live;
    rw.buf[0] = rw.buf[1] = rw.buf[2] = rw.buf[3] = 0;
    rw_lo = rw;
    RowWriter rw_hi8 = rw;                       // three streams: fp16 at byte 0, e4m3 residuals at byte 416, e4m3 features at byte 624
    if (kMesh) {
        rw_lo.dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(Arows + f * FUSED_K) + FUSED_X_BYTE0);
        rw_hi8.dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(Arows + f * FUSED_K) + FUSED_X_BYTE1);
    }

    const float* p = tile + lane * kChainPitch;
    float* jout = tile + lane * kChainPitch;
    float G[NJ][12];     // global transforms, rows [R | t]
    float J[NJ][3];      // rest joints
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;   // centre joint translation

#pragma unroll
    for (int pos = 0; pos < NJ; ++pos) {
        const int j = kStd ? smpl_dfs(pos) : pos;
        const int par = kStd ? smpl_parent(j) : (pos == 0 ? -1 : pc.parents[j]);

        float R[9];
        smpl_rodrigues_call(p[j * 3 + 0], p[j * 3 + 1], p[j * 3 + 2], R); // (read before jout[j * 3 ..] is written)

        if (kMesh && pos > 0) {   // pose_map = R - I (tensutils.py:41-48) in the three parts of the K12 operand row
#pragma unroll
            for (int e = 0; e < 9; ++e) {
                const float v = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
                uint16_t hi; uint8_t lo8, hi8;
                feature_parts(v, hi, lo8, hi8);
                rw.push(hi); rw_lo.push8(lo8); rw_hi8.push8(hi8);
            }
        }

        // rest joint: J = J_template + Jdirs * beta
#pragma unroll
        for (int c = 0; c < 3; ++c) J[j][c] = rest_joint(pc.J_template[j * 3 + c], &pc.Jdirs[(j * 3 + c) * NBETA], beta);

        if (pos == 0) {          // smpl_layer.py:105-106
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                G[j][r * 4 + 0] = R[r * 3 + 0]; G[j][r * 4 + 1] = R[r * 3 + 1];
                G[j][r * 4 + 2] = R[r * 3 + 2]; G[j][r * 4 + 3] = J[j][r];
            }
        } else {                 // smpl_layer.py:109-119
            compose(G[par], R, __fsub_rn(J[j][0], J[par][0]), __fsub_rn(J[j][1], J[par][1]),
                    __fsub_rn(J[j][2], J[par][2]), G[j]);
        }
        // joints (smpl_layer.py:145), translation added now, centring applied below
        jout[j * 3 + 0] = __fadd_rn(G[j][3], o0);
        jout[j * 3 + 1] = __fadd_rn(G[j][7], o1);
        jout[j * 3 + 2] = __fadd_rn(G[j][11], o2);
        if (j == center_idx) { c0 = G[j][3]; c1 = G[j][7]; c2 = G[j][11]; }

        if (kMesh && live) {   // A_j = G_j - pack(G_j @ [j_rest; 0])  (smpl_layer.py:126-132), TMEM-tile layout
            float* dst = Askin + ((f >> 5) * FUSED_ASKIN_COLS + j * 12) * 32 + (f & 31);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                // rotations times 2^-S (exact): the blend accumulator holds 2^S v_posed
                dst[askin_col(r, 0) * 32] = G[j][r * 4 + 0] * pc.rot_scale; dst[askin_col(r, 1) * 32] = G[j][r * 4 + 1] * pc.rot_scale;
                dst[askin_col(r, 2) * 32] = G[j][r * 4 + 2] * pc.rot_scale; dst[askin_col(r, 3) * 32] = skin_t(G[j], r, J[j][0], J[j][1], J[j][2]);
            }
        }
    }

    if (centre) {   // smpl_layer.py:149-152
        o0 = -c0; o1 = -c1; o2 = -c2;
#pragma unroll
        for (int k = 0; k < 72; k += 3) {
            jout[k + 0] = __fadd_rn(jout[k + 0], o0); jout[k + 1] = __fadd_rn(jout[k + 1], o1);
            jout[k + 2] = __fadd_rn(jout[k + 2], o2);
        }
    }

    {   // coalesced copy of the nf joint rows
        __syncwarp();
        float* dst = joints + f0 * 72;
        const int n = nf * 72;
        int r = 0, c = lane;
        for (int q = lane; q < n; q += 32) {
            dst[q] = tile[r * kChainPitch + c];
            c += 32;
            if (c >= 72) { c -= 72; ++r; }
        }
    }

    if (kMesh) {
        if (live) { off[f * 3 + 0] = o0; off[f * 3 + 1] = o1; off[f * 3 + 2] = o2; }
        rw.push(0); rw_lo.push8(0); rw_hi8.push8(0);   // element 207 closes the fp16 block and both e4m3 blocks
        rw.dst = reinterpret_cast<uint4*>(Arows + f * FUSED_K + FUSED_COL_BETA);
        // k-step 26: bh | 2^15 2^15 2^15 | 0 0 0      k-step 27: bl | 0 x 6      (fp16; prk_internal.h "K12 operand layout")
        uint16_t bl[NBETA];
#pragma unroll
        for (int k = 0; k < NBETA; ++k) {
            const __half h = __float2half_rn(beta[k]);
            rw.push(__half_as_ushort(h));
            bl[k] = __half_as_ushort(__float2half_rn(beta[k] - __half2float(h)));
        }
        rw.push(0x7800); rw.push(0x7800); rw.push(0x7800); rw.push(0); rw.push(0); rw.push(0);
#pragma unroll
        for (int k = 0; k < NBETA; ++k) rw.push(bl[k]);
#pragma unroll
        for (int k = NBETA; k < 16; ++k) rw.push(0);
    }
}


// ---- latency-optimised variant: one warp per frame, lane = joint ---------------------
// For small batches the thread-per-frame kernel is bound by one thread's ~7,000 dependent
// instructions.  Here the 24 joints' Rodrigues / rest joints run in parallel across lanes and
// the kinematic chain is a level-synchronous scan in warp registers (__shfl_sync, tree depth
// 8), cutting the critical path ~8x.  Same arithmetic helpers => bit-identical outputs.
__device__ __constant__ int8_t c_parent[32] = {0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21,
                                               0, 0, 0, 0, 0, 0, 0, 0};
__device__ __constant__ int8_t c_depth[32] = {0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8,
                                              99, 99, 99, 99, 99, 99, 99, 99};
// position of joint j in the depth-first order used for the A' column blocks
__device__ __constant__ int8_t c_dfs_pos[32] = {0, 1, 5, 9, 2, 6, 10, 3, 7, 11, 4, 8, 12, 14, 19, 13, 15, 20, 16, 21, 17, 22, 18, 23,
                                                0, 0, 0, 0, 0, 0, 0, 0};
constexpr int kWarpsPerBlock = 8;        // joints-only variant
#ifndef PRK_CHAIN_WPB
#define PRK_CHAIN_WPB 16
#endif
constexpr int kWarpsPerBlockMesh = PRK_CHAIN_WPB;   // full-mesh variant: 16 consecutive frames per block, so that the block
                                         // writes its AskinT columns as 64-byte runs (two full sectors) instead of
                                         // 288 scattered 4-byte stores per frame
// Above these frame counts per launch the thread-per-frame kernel is faster (scripts/chain_threshold.py).  Full mesh, round 2c: the
// lane-per-joint kernel wins up to the 65,536 frames a launch can have (139 vs 142 us there, 80 vs 103 us at 32,768; before the
// beta-major Jdirs copy it lost from 49,152 on), so mesh launches of standard-tree models always take it.
constexpr int64_t kWarpVariantMaxFrames = 65536;
constexpr int64_t kWarpVariantMaxFramesJointsOnly = 16384;

template <bool kMesh>
__global__ void __launch_bounds__((kMesh ? kWarpsPerBlockMesh : kWarpsPerBlock) * 32)
pose_chain_warp_kernel(const float* __restrict__ Jc /* J_template[72] | Jdirs^T[10][72] | model_betas[10] | rot_scale */,
                       const float* __restrict__ pose, const float* __restrict__ betas,
                       const float* __restrict__ trans, const BatchFlags* __restrict__ flags, uint32_t mode,
                       int center_idx, int64_t B, uint16_t* __restrict__ Arows, float* __restrict__ Askin,
                       float* __restrict__ off, float* __restrict__ joints) {
    constexpr int kW = kMesh ? kWarpsPerBlockMesh : kWarpsPerBlock;
    __shared__ __align__(16) uint16_t s_row[kMesh ? kW : 1][FUSED_K];
    __shared__ float s_askin[kMesh ? FUSED_ASKIN_COLS : 1][kMesh ? kW + 1 : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t f_own = (int64_t)blockIdx.x * kW + warp;
    const bool valid = f_own < B;                 // warp-uniform; invalid warps compute on the last frame and store nothing
    if (!kMesh && !valid) return;
    const int64_t f = valid ? f_own : B - 1;
    const unsigned FULL = 0xffffffffu;
    // the vertex kernel behind this one is launched with a programmatic dependency: let its CTAs become resident and
    // run their prologue as soon as SMs free up (it waits with griddepcontrol.wait before reading our outputs)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    bool frame_betas = (mode & MODE_FRAME_BETAS_ALWAYS) != 0;
    bool add_trans = (mode & MODE_TRANS_ALWAYS) != 0;
    if (mode & MODE_FRAME_BETAS_FLAG) frame_betas = flags->betas_nonzero != 0;
    if (mode & MODE_TRANS_FLAG) add_trans = flags->trans_nonzero != 0;
    const bool centre = (center_idx >= 0) && !add_trans;

    const bool active = lane < NJ;
    const int j = active ? lane : 0;
    float beta[NBETA];
#pragma unroll
    for (int k = 0; k < NBETA; ++k) beta[k] = frame_betas ? betas[f * NBETA + k] : Jc[72 + 720 + k];

    float R[9];
    smpl_rodrigues(pose[f * 72 + j * 3 + 0], pose[f * 72 + j * 3 + 1], pose[f * 72 + j * 3 + 2], R);
    float J[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {   // rest_joint() on the beta-major copy of Jdirs: a warp-wide load touches 288 contiguous bytes
        float acc = Jc[j * 3 + c];
#pragma unroll
        for (int k = 0; k < NBETA; ++k) acc = fmaf(Jc[72 + k * 72 + j * 3 + c], beta[k], acc);
        J[c] = acc;
    }

    const int par = c_parent[lane], depth = c_depth[lane];
    float G[12];
#pragma unroll
    for (int r = 0; r < 3; ++r) {   // root transform (only lane 0's value is used)
        G[r * 4 + 0] = R[r * 3 + 0]; G[r * 4 + 1] = R[r * 3 + 1]; G[r * 4 + 2] = R[r * 3 + 2]; G[r * 4 + 3] = J[r];
    }
    const float pj0 = __shfl_sync(FULL, J[0], par), pj1 = __shfl_sync(FULL, J[1], par), pj2 = __shfl_sync(FULL, J[2], par);
    const float t0 = __fsub_rn(J[0], pj0), t1 = __fsub_rn(J[1], pj1), t2 = __fsub_rn(J[2], pj2);
#pragma unroll 1
    for (int d = 1; d <= 8; ++d) {
        float Gp[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) Gp[e] = __shfl_sync(FULL, G[e], par);
        if (depth == d) compose(Gp, R, t0, t1, t2, G);
    }

    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    if (add_trans) { o0 = trans[f * 3 + 0]; o1 = trans[f * 3 + 1]; o2 = trans[f * 3 + 2]; }
    float x0 = __fadd_rn(G[3], o0), x1 = __fadd_rn(G[7], o1), x2 = __fadd_rn(G[11], o2);
    if (centre) {
        o0 = -__shfl_sync(FULL, G[3], center_idx); o1 = -__shfl_sync(FULL, G[7], center_idx);
        o2 = -__shfl_sync(FULL, G[11], center_idx);
        x0 = __fadd_rn(x0, o0); x1 = __fadd_rn(x1, o1); x2 = __fadd_rn(x2, o2);
    }
    if (active && valid) {
        float* jo = joints + f * 72 + j * 3;
        jo[0] = x0; jo[1] = x1; jo[2] = x2;
    }
    if (!kMesh) return;

    if (active) {   // A_j columns of this frame into the block's staging tile [12 * joint + e][frame in block]
        const float rs = Jc[72 + 720 + NBETA];   // 2^-S on the rotations (exact): the blend accumulator holds 2^S v_posed
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            s_askin[j * 12 + askin_col(r, 0)][warp] = G[r * 4 + 0] * rs; s_askin[j * 12 + askin_col(r, 1)][warp] = G[r * 4 + 1] * rs;
            s_askin[j * 12 + askin_col(r, 2)][warp] = G[r * 4 + 2] * rs;
            s_askin[j * 12 + askin_col(r, 3)][warp] = skin_t(G, r, J[0], J[1], J[2]);
        }
    }
    __syncthreads();
    {   // TMEM-tile layout [frame / 32][288][frame % 32]: the block's 16 frames are one half of every 128-byte line
        const int64_t f0 = (int64_t)blockIdx.x * kW;
        float* dst = Askin + (f0 >> 5) * FUSED_ASKIN_COLS * 32 + (f0 & 31);
        for (int idx = threadIdx.x; idx < FUSED_ASKIN_COLS * kW; idx += kW * 32) {
            const int col = idx / kW, fl = idx - col * kW;
            dst[col * 32 + fl] = s_askin[col][fl];
        }
    }
    if (!valid) return;
    if (lane == 0) { off[f * 3 + 0] = o0; off[f * 3 + 1] = o1; off[f * 3 + 2] = o2; }

    // A' row assembled in shared memory, then written with 16-byte stores
    uint16_t* row = s_row[warp];
    uint8_t* rowb = reinterpret_cast<uint8_t*>(row);
    if (active && j > 0) {   // the three parts of every pose feature (prk_internal.h "K12 operand layout")
        const int base = 9 * (c_dfs_pos[j] - 1);
#pragma unroll
        for (int e = 0; e < 9; ++e) {
            const float v = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
            uint16_t hi; uint8_t lo8, hi8;
            feature_parts(v, hi, lo8, hi8);
            row[base + e] = hi; rowb[FUSED_X_BYTE0 + base + e] = lo8; rowb[FUSED_X_BYTE1 + base + e] = hi8;
        }
    }
    if (lane < 16) {   // k-step 26: bh | 2^15 2^15 2^15 | 0 0 0      k-step 27: bl | 0 x 6      (fp16)
        uint16_t h16 = lane < NBETA + 3 ? 0x7800 : 0, l16 = 0;
        if (lane < NBETA) {
            const __half h = __float2half_rn(beta[lane]);
            h16 = __half_as_ushort(h);
            l16 = __half_as_ushort(__float2half_rn(beta[lane] - __half2float(h)));
        }
        row[FUSED_COL_BETA + lane] = h16; row[FUSED_COL_BETA + 16 + lane] = l16;
    }
    if (lane == 31) { row[FUSED_COL_LO - 1] = 0; rowb[FUSED_X_BYTE0 + 207] = 0; rowb[FUSED_X_BYTE1 + 207] = 0; }
    __syncwarp();
    const uint4* src = reinterpret_cast<const uint4*>(row);
    uint4* dst = reinterpret_cast<uint4*>(Arows + f * FUSED_K);
    for (int i = lane; i < (FUSED_A_STEPS * 16) / 8; i += 32) dst[i] = src[i];
}

// Whole-batch tests `torch.norm(x) == 0` (smpl_layer.py:87,148): true iff every x*x is 0
// in fp32 (NaN counts as non-zero).  d_flags must be zeroed before launch.
__global__ void __launch_bounds__(256)
batch_flags_kernel(const float* __restrict__ betas, const float* __restrict__ trans, int64_t B,
                   BatchFlags* __restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bnz = 0, tnz = 0;
    if (betas)
        for (int64_t i = t0; i < B * NBETA; i += stride) { const float v = betas[i]; bnz |= !(v * v == 0.0f); }
    if (trans)
        for (int64_t i = t0; i < B * 3; i += stride) { const float v = trans[i]; tnz |= !(v * v == 0.0f); }
    bnz = __syncthreads_or(bnz);
    tnz = __syncthreads_or(tnz);
    if (threadIdx.x == 0) {
        if (bnz) atomicOr(&flags->betas_nonzero, 1);
        if (tnz) atomicOr(&flags->trans_nonzero, 1);
    }
}

cudaError_t launch_batch_flags(const float* d_betas, const float* d_trans, int64_t B,
                               BatchFlags* d_flags, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_flags, 0, sizeof(BatchFlags), s);
    if (e != cudaSuccess) return e;
    int64_t n = B * NBETA;
    unsigned g = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
    if (g < 1) g = 1;
    if (g > 148u * 4u) g = 148u * 4u;
    batch_flags_kernel<<<g, 256, 0, s>>>(d_betas, d_trans, B, d_flags);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_pose_chain(const Model& m, const float* d_pose, const float* d_betas,
                              const float* d_trans, const BatchFlags* d_flags, int center_idx,
                              int64_t B, bool full_mesh, uint16_t* d_Arows, float* d_Askin,
                              float* d_off, float* d_joints, cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    bool model_betas_zero = true;
    for (int k = 0; k < NBETA; ++k) model_betas_zero &= (m.pc.model_betas[k] == 0.0f);
    uint32_t mode = 0;
    if (d_betas) mode |= model_betas_zero ? MODE_FRAME_BETAS_ALWAYS : MODE_FRAME_BETAS_FLAG;
    if (d_trans) mode |= (center_idx < 0) ? MODE_TRANS_ALWAYS : MODE_TRANS_FLAG;
    const unsigned grid = (unsigned)((B + 127) / 128);
    const bool std_tree = m.pc.standard_tree != 0;
    static const int64_t warp_max = [] { const char* e = getenv("PRK_CHAIN_WARP_MAX"); return e ? (int64_t)atoll(e) : kWarpVariantMaxFrames; }();
    static const int64_t warp_max_jo = [] { const char* e = getenv("PRK_CHAIN_WARP_MAX_JO"); return e ? (int64_t)atoll(e) : kWarpVariantMaxFramesJointsOnly; }();
    if (std_tree && B <= (full_mesh ? warp_max : warp_max_jo)) {   // latency-bound regime: one warp per frame
        const int wpb = full_mesh ? kWarpsPerBlockMesh : kWarpsPerBlock;
        const unsigned g = (unsigned)((B + wpb - 1) / wpb);
#define PRK_LAUNCH_W(MESH)                                                                                \
    pose_chain_warp_kernel<MESH><<<g, wpb * 32, 0, s>>>(m.d_Jc, d_pose, d_betas, d_trans, d_flags, mode, \
                                                        center_idx, B, d_Arows, d_Askin, d_off, d_joints)
        if (full_mesh) PRK_LAUNCH_W(true); else PRK_LAUNCH_W(false);
#undef PRK_LAUNCH_W
        count_launch();
        return cudaGetLastError();
    }
#define PRK_LAUNCH(STD, MESH)                                                                     \
    pose_chain_kernel<STD, MESH><<<grid, 128, 0, s>>>(m.pc, d_pose, d_betas, d_trans, d_flags, mode, \
                                                      center_idx, B, d_Arows, d_Askin, d_off, d_joints)
    if (std_tree) { if (full_mesh) PRK_LAUNCH(true, true); else PRK_LAUNCH(true, false); }
    else          { if (full_mesh) PRK_LAUNCH(false, true); else PRK_LAUNCH(false, false); }
#undef PRK_LAUNCH
    count_launch();
    return cudaGetLastError();
}

// true when launch_pose_chain will read d_flags for this configuration
bool pose_chain_needs_flags(const Model& m, const float* d_betas, const float* d_trans, int center_idx) {
    bool model_betas_zero = true;
    for (int k = 0; k < NBETA; ++k) model_betas_zero &= (m.pc.model_betas[k] == 0.0f);
    return (d_betas && !model_betas_zero) || (d_trans && center_idx >= 0);
}

}  // namespace prk
