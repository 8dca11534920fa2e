// Internal declarations shared by the translation units of libposerisk_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

#include "../../include/poserisk_b200.h"

namespace prk {

constexpr int NV = PRK_NUM_VERTS;       // 6890
constexpr int NJ = PRK_NUM_JOINTS;      // 24
constexpr int NBETA = PRK_NUM_BETAS;    // 10
constexpr int NPOSE = PRK_NUM_POSE_FEATS;  // 207
constexpr int NVC = NV * 3;             // 20670 vertex coordinates

constexpr int GEMM_N = 20736;           // 20670 vertex coordinates padded to 216 tiles of 96

// ---- fused blend + skinning geometry (DESIGN.md "K12", prk_fused.cu) ---------
// K12 operand layout: rows of k-steps of 32 bytes (16 x 16-bit or 32 x 8-bit elements), each part stored ONCE.
// Everything on the B' side is multiplied by 2^S (S = prk_model::blend_scale_log2, chosen so that max|posedirs| 2^S <= 2^14 and
// max|shapedirs| 2^S <= 2^15); the accumulator holds 2^S v_posed and the pose kernel hands the skinning rotations over as
// 2^-S R, so nothing is rescaled in the epilogue.  With F = a pose feature (R - I entry), P = a posedirs entry 2^S,
// b = bh + bl a beta and s = sh + sl a shapedirs entry 2^S as two fp16 parts each, and a v_template entry 2^S = 2^15 (u1 + u2 + u3)
// in three fp16 parts:
//   A' row (per frame, 28 k-steps = 7 chunks of 128 bytes):
//     steps  0..12  fp16   Fh = fp16(F)                          (207 + 1 zero; feature p = 9*(pos-1)+e, pos = DFS position)
//     steps 13..25  e4m3   208 bytes fp8((F - Fh) 2^12)  |  208 bytes fp8(F)
//     step  26      fp16   bh[0..9] | 2^15 2^15 2^15 | 0 0 0
//     step  27      fp16   bl[0..9] | 0 x 6
//   B' row (per vertex coordinate, 28 k-steps = 7 chunks):
//     steps  0..12  fp16   Ph = fp16(P)
//     steps 13..25  e4m3   208 bytes fp8(P 2^-12)        |  208 bytes fp8(P - Ph)
//     step  26      fp16   sh[0..9] | u1 u2 u3 | 0 0 0        (x A26: bh.sh + template;  x A27: bl.sh)
//     step  27      fp16   sl[0..9] | 0 x 6                   (x A26: bh.sl)
// MMAs per tile: A'[b] x B'[b] for b = 0..12 (kind::f16, fp16: Fh.Ph), for b = 13..25 (kind::f8f6f4, K = 32: the two cross terms
// (F - Fh).P + F.(P - Ph), byte i of one row meeting byte i of the other) and the three beta products: 29 MMAs and 7 streamed
// chunks instead of the 44 MMAs / 8 chunks of the bf16 hi.hi + lo.hi + hi.lo scheme of round 1.  The pose cross terms are 2^-12
// of the main one, so the 4 significant bits of e4m3 leave 2^-15.5 of the pose blend shapes as error: 2.3e-7 .. 4.7e-7 of the
// vertex range on the config-2 / pose x 0.6 / pose x 1.2 / large-beta cases (numpy model) against 0.8e-7 .. 2.0e-7 before and
// a budget of 2e-6.  Shape part: 22 bits per factor, bl.sl dropped: 0.5e-7 (betas ~ N(0,1)) .. 1.3e-7 (betas x 10).
constexpr int FUSED_K = 512;                             // bf16 columns per stored row (A' rows use 448 of them)
constexpr int FUSED_POSE_STEPS = 13;
constexpr int FUSED_A_STEPS = 2 * FUSED_POSE_STEPS + 2;  // 28
constexpr int FUSED_B_STEPS = 2 * FUSED_POSE_STEPS + 2;  // 28
constexpr int FUSED_A_CHUNKS = FUSED_A_STEPS / 4;        // 7 resident chunks of 64 columns
constexpr int FUSED_B_CHUNKS = FUSED_B_STEPS / 4;        // 7 streamed chunks
constexpr int FUSED_COL_LO = 16 * FUSED_POSE_STEPS;      // 208: first 16-bit column of the e4m3 block (byte 416 of a row)
constexpr int FUSED_COL_BETA = 2 * FUSED_COL_LO;         // 416
constexpr int FUSED_X_BYTE0 = 2 * FUSED_COL_LO;          // byte offset of the first e4m3 sub-block (208 bytes) in a row
constexpr int FUSED_X_BYTE1 = FUSED_X_BYTE0 + FUSED_COL_LO;   // ... and of the second one
constexpr int FUSED_BM = 128;            // frames per tile (TMEM lanes)
constexpr int FUSED_VT = 32;             // vertices per tile
constexpr int FUSED_BN = FUSED_VT * 3;   // 96 accumulator columns
constexpr int FUSED_NT = GEMM_N / FUSED_BN;   // 216 vertex tiles (6912 vertex slots)
constexpr int FUSED_ASKIN_COLS = NJ * 12;     // 288 TMEM columns of A_j per frame
// per tile and group of 4 weights: 32 x float4 weights | 32 x uint4 TMEM columns (12 * joint)
// (round 1 repeated the column table per TMEM lane quarter as absolute tensor-memory addresses, (32 * quarter) << 16 | 12 * joint,
// to save one add per gather (-2 %); round 2 keeps ONE table and spends the 1.5 KB per slot on two more ring slots (-2 %, measured))
#ifndef PRK_WCOL_COPIES
#define PRK_WCOL_COPIES 1
#endif
constexpr int FUSED_WCOL_COPIES = PRK_WCOL_COPIES;   // 1: one table of plain columns 12 * joint; the kernel adds its lane quarter
constexpr int FUSED_WGROUP_BYTES = FUSED_VT * 16 * (1 + FUSED_WCOL_COPIES);
// element index of B'[n][k] (vertex coordinate n, K12 column k) inside the chunk-image layout of prk_model::d_B2
__host__ __device__ constexpr size_t fused_b2_index(int n, int k) {
    const int tile = n / FUSED_BN, r = n % FUSED_BN, chunk = k / 64, kk = k % 64;
    return (((size_t)tile * FUSED_B_CHUNKS + chunk) * FUSED_BN + r) * 64 + (size_t)((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7));
}

// byte index of byte `byte` (0..1023) of B' row n inside the same chunk images (16-byte units swizzled as above)
__host__ __device__ constexpr size_t fused_b2_byte_index(int n, int byte) {
    return fused_b2_index(n, byte >> 1) * 2 + (size_t)(byte & 1);
}

// Unit ranges of the vertex kernel's CTAs (pairs), balanced by cost: unit u = (frame tile (pair) u / 216, vertex tile u % 216)
// sits at 16 u + switch_cost16 * (u / 216) on a cost axis on which entering a new frame tile inside a range costs switch_cost16
// sixteenths of a unit; range k of n_ranges starts at the first unit whose position is >= k * total / n_ranges.
// Contiguous, covers [0, n_units) exactly, fused_first_unit(n_ranges) = n_units.
__host__ __device__ inline int64_t fused_first_unit(int64_t k, int64_t n_ranges, int64_t n_units, int switch_cost16) {
    if (k >= n_ranges) return n_units;
    const int64_t tile_cost = 16 * (int64_t)FUSED_NT + switch_cost16;
    const int64_t total_cost = 16 * n_units + switch_cost16 * (n_units / FUSED_NT - 1);
    const int64_t t = k * total_cost / n_ranges;
    const int64_t f = t / tile_cost, r = t - f * tile_cost;
    int64_t v = (r + 15) / 16;
    if (v > FUSED_NT) v = FUSED_NT;
    const int64_t u = f * FUSED_NT + v;
    return u < n_units ? u : n_units;
}

// Rest joints as an affine function of betas: J = J_template + Jdirs * beta
// (folds J_regressor @ (v_template + shapedirs beta), smpl_layer.py:91,95).
struct PoseConsts {
    float J_template[NJ * 3];
    float Jdirs[NJ * 3 * NBETA];   // [joint*3+c][beta]
    float model_betas[NBETA];
    int32_t parents[NJ];
    int32_t standard_tree;         // 1: parents == SMPL kintree (static unroll path)
    float rot_scale;               // 2^-S: the skinning rotations handed to the vertex kernel are scaled by it (the accumulator holds 2^S v_posed)
};

}  // namespace prk

// The opaque handle of the C ABI.
struct prk_model {
    int device = -1;
    int sm_count = 0;
    int nnz_groups = 1;            // ceil(max non-zero weights per vertex / 4)
    int max_weights = 0;
    int blend_scale_log2 = 0;      // S: every B' operand is multiplied by 2^S
    prk::PoseConsts pc;                 // host copy, passed by value to the pose kernel
    // device buffers
    float* d_Jc = nullptr;         // J_template[72] | Jdirs^T[10][72] | model_betas[10] | rot_scale (lane-per-joint pose kernel)
    // B' as the shared-memory IMAGES of its TMA chunks: [vertex tile][chunk][96 rows][64 bf16], every chunk one
    // contiguous 12 KB block with the 128-byte swizzle already applied (16-byte unit j of row r sits at j ^ (r & 7)),
    // so a chunk is fetched with ONE linear bulk copy instead of a 96-row tensor box (fused_b2_index)
    uint16_t* d_B2 = nullptr;
    CUtensorMap tmap_B2;           // the same image as [tiles x chunks x 96 rows][64 bf16], boxes of 48 rows (CTA pairs load halves)
    uint8_t* d_wpack = nullptr;    // [FUSED_NT][nnz_groups][FUSED_WGROUP_BYTES] per-tile skinning weights
    // scoring only reads the pose: it runs on its own stream beside the mesh path
    cudaStream_t s_score = nullptr;
    cudaEvent_t ev_in = nullptr, ev_score = nullptr;
    // host-buffer pipeline (prk_pipeline_host): copy-in / copy-out streams beside the kernels, two input sets
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_joints = nullptr, ev_out = nullptr, ev_set_free[2] = {nullptr, nullptr};
    uint64_t host_calls = 0;
};

namespace prk {
using Model = ::prk_model;

// Per-batch device flags (whole-batch tests of smpl_layer.py:87,148)
struct BatchFlags {
    int32_t betas_nonzero;
    int32_t trans_nonzero;
};

// ---- launchers (each returns cudaError_t from the launch) ------------------
cudaError_t launch_batch_flags(const float* d_betas, const float* d_trans, int64_t B,
                               BatchFlags* d_flags, cudaStream_t s);

// K2a: Rodrigues + kinematic chain (+ the fused kernel's per-frame operands when full_mesh)
cudaError_t launch_pose_chain(const Model& m, const float* d_pose, const float* d_betas,
                              const float* d_trans, const BatchFlags* d_flags, int center_idx,
                              int64_t B, bool full_mesh, uint16_t* d_Arows, float* d_Askin,
                              float* d_off, float* d_joints, cudaStream_t s);

bool pose_chain_needs_flags(const Model& m, const float* d_betas, const float* d_trans, int center_idx);

// K12: fused blend GEMM + skinning (prk_fused.cu); A' rows / AskinT in the K12 layouts, rows_pad = multiple of 128
// vpitch: floats per vertex row (NVC = dense); 16-byte aligned rows leave through TMA stores (models with <= 4 weights)
cudaError_t launch_fused(const Model& m, const CUtensorMap& tmap_A, int64_t rows_pad, const float* d_AskinT,
                         const float* d_off, int64_t B, float* d_verts, int64_t vpitch, cudaStream_t s);

// the calling thread's next vertex-kernel launches run beside a multi-GPU exchange: leave shared memory for its blocks
void set_fused_exchange_hint(bool on);

// verification hooks of prk_debug_blend: plain FFMA evaluation of the same bf16 operands, identity A_j tiles
cudaError_t launch_blend_simt(const Model& m, const uint16_t* d_Arows, int64_t rows, float* d_vposed, cudaStream_t s);
cudaError_t launch_identity_askin(const Model& m, float* d_AskinT, float* d_off, int64_t rows_pad, cudaStream_t s);

// multi-GPU exchange (prk_comm.cu): prk_allgather_rows with / without joining `stream` with the wait for the peers' rows
extern "C" int prk_allgather_rows_impl(prk_comm* c, const void* d_local, int64_t n_local, int64_t row_offset, int64_t row_bytes,
                                       void** d_gathered_out, void* stream, bool join);

// K3: Euler angles + REBA/RULA
// --debug_joints list as a by-value kernel parameter: bit j of `mask` = joint j is listed, slot[j] = its position
struct DebugSlots {
    uint32_t mask = 0;
    int8_t slot[NJ] = {};
};
cudaError_t launch_score_pose(const void* d_pose, int pose_dtype, const prk_addinfo* d_info, int32_t n_tracks,
                              const int32_t* d_track, int64_t B, uint32_t which,
                              prk_score_rec* d_out, double* d_euler_out, const DebugSlots& dbg, int n_debug,
                              cudaStream_t s);
cudaError_t launch_score_euler(const double* d_euler, const prk_addinfo* d_info, int32_t n_tracks,
                               const int32_t* d_track, int64_t B, uint32_t which,
                               prk_score_rec* d_out, cudaStream_t s);
cudaError_t launch_euler(const void* d_pose, int pose_dtype, int64_t n_rot, double* d_euler,
                         uint8_t* d_bad, cudaStream_t s);
cudaError_t launch_rot_to_angle(const void* d_rotmat, int dtype, int64_t n_rot, void* d_rvec, uint8_t* d_bad,
                                cudaStream_t s);
cudaError_t launch_score_hist(const prk_score_rec* d_scores, int64_t B, uint32_t which,
                              unsigned long long* d_hist, cudaStream_t s);

// TMA descriptor helper (driver entry point fetched at run time; no libcuda link)
int encode_tmap_2d_ex(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint32_t box_cols, int elem_bytes, int swizzle128);
int encode_tmap_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                   uint32_t box_cols, int elem_bytes);
int encode_tmap_2d_f32_strided(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                               uint32_t box_rows, uint32_t box_cols);
int encode_tmap_2d_bf16(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols,
                        uint32_t box_rows, uint32_t box_cols);

void count_launch(int n = 1);

}  // namespace prk
