// K2b: linear-blend skinning, HBM-bound.
//
// Reference behaviour (lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py:134-155):
//   T_v   = sum_j weights[v][j] * A_j            (dense (B,4,4,24)@(24,6890) there)
//   vert  = (T_v @ [v_posed; 1])[:3]  (+ trans | - centre joint)
// Here the weights are compacted to groups of 4 (joint id, weight) pairs per vertex
// (exactly-zero weights contribute exactly 0 in the reference and are dropped).
//
// Mapping: one thread = one vertex, a block = 256 consecutive vertices x kFramesPerBlock
// frames.  The block stages the frames' 24 A_j matrices in shared memory as [joint][13]
// (13-word pitch: the 32 lanes of a warp hit up to 24 distinct joints -> 24 distinct
// banks, same joint -> broadcast, so every LDS is conflict free), each thread keeps its
// weights/ids in registers across the frame loop, v_posed reads and vertex writes are
// coalesced along the vertex axis.
// Algorithmic HBM bytes per frame: 82,680 (v_posed) + 82,680 (verts) + 1,152 (A) + 12.
#include "prk_internal.h"

namespace prk {

namespace {

constexpr int kVertsPerBlock = 256;
constexpr int kFramesPerBlock = 8;
constexpr int kJointPitch = 13;
constexpr int kFramePitch = NJ * kJointPitch;   // 312 floats

__global__ void __launch_bounds__(kVertsPerBlock)
skin_kernel(const float* __restrict__ vposed, const float* __restrict__ Askin, const float* __restrict__ off,
            const float4* __restrict__ wval, const uint32_t* __restrict__ widx, int nnz_groups, int64_t B,
            float* __restrict__ verts) {
    __shared__ float sA[kFramesPerBlock * kFramePitch];
    __shared__ float sOff[kFramesPerBlock * 3];

    const int64_t f0 = (int64_t)blockIdx.y * kFramesPerBlock;
    const int nf = (int)((B - f0) < kFramesPerBlock ? (B - f0) : kFramesPerBlock);

    // stage A_j of nf frames: global [frame][24][12] -> shared [frame][24][13]
    for (int i = threadIdx.x; i < nf * NJ * 12; i += kVertsPerBlock) {
        const int fr = i / (NJ * 12), r = i - fr * (NJ * 12);
        const int j = r / 12, e = r - j * 12;
        sA[fr * kFramePitch + j * kJointPitch + e] = Askin[f0 * (NJ * 12) + i];
    }
    if (threadIdx.x < nf * 3) sOff[threadIdx.x] = off[f0 * 3 + threadIdx.x];
    __syncthreads();

    const int v = blockIdx.x * kVertsPerBlock + threadIdx.x;
    if (v >= NV) return;

    if (nnz_groups == 1) {
        const float4 w = wval[v];
        const uint32_t id = widx[v];
        const int j0 = (id & 0xFF) * kJointPitch, j1 = ((id >> 8) & 0xFF) * kJointPitch;
        const int j2 = ((id >> 16) & 0xFF) * kJointPitch, j3 = (id >> 24) * kJointPitch;
#pragma unroll
        for (int fr = 0; fr < kFramesPerBlock; ++fr) {
            if (fr < nf) {
                const float* vp = vposed + (size_t)(f0 + fr) * VPOSED_PITCH + (size_t)v * 3;
                const float px = vp[0], py = vp[1], pz = vp[2];
                const float* a = sA + fr * kFramePitch;
                float T[12];
#pragma unroll
                for (int e = 0; e < 12; ++e)
                    T[e] = a[j0 + e] * w.x + a[j1 + e] * w.y + a[j2 + e] * w.z + a[j3 + e] * w.w;
                float* o = verts + ((size_t)(f0 + fr) * NV + v) * 3;
                o[0] = T[0] * px + T[1] * py + T[2] * pz + T[3] + sOff[fr * 3 + 0];
                o[1] = T[4] * px + T[5] * py + T[6] * pz + T[7] + sOff[fr * 3 + 1];
                o[2] = T[8] * px + T[9] * py + T[10] * pz + T[11] + sOff[fr * 3 + 2];
            }
        }
    } else {   // dense-ish rows: more than 4 non-zero weights per vertex
        for (int fr = 0; fr < nf; ++fr) {
            const float* vp = vposed + (size_t)(f0 + fr) * VPOSED_PITCH + (size_t)v * 3;
            const float px = vp[0], py = vp[1], pz = vp[2];
            const float* a = sA + fr * kFramePitch;
            float T[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) T[e] = 0.f;
            for (int g = 0; g < nnz_groups; ++g) {
                const float4 w = wval[(size_t)g * NV + v];
                const uint32_t id = widx[(size_t)g * NV + v];
                const int j0 = (id & 0xFF) * kJointPitch, j1 = ((id >> 8) & 0xFF) * kJointPitch;
                const int j2 = ((id >> 16) & 0xFF) * kJointPitch, j3 = (id >> 24) * kJointPitch;
#pragma unroll
                for (int e = 0; e < 12; ++e)
                    T[e] += a[j0 + e] * w.x + a[j1 + e] * w.y + a[j2 + e] * w.z + a[j3 + e] * w.w;
            }
            float* o = verts + ((size_t)(f0 + fr) * NV + v) * 3;
            o[0] = T[0] * px + T[1] * py + T[2] * pz + T[3] + sOff[fr * 3 + 0];
            o[1] = T[4] * px + T[5] * py + T[6] * pz + T[7] + sOff[fr * 3 + 1];
            o[2] = T[8] * px + T[9] * py + T[10] * pz + T[11] + sOff[fr * 3 + 2];
        }
    }
}

}  // namespace

cudaError_t launch_skin(const Model& m, const float* d_vposed, const float* d_Askin, const float* d_off,
                        int64_t B, float* d_verts, cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    const int64_t fy_total = (B + kFramesPerBlock - 1) / kFramesPerBlock;
    // gridDim.y is limited to 65535: walk the batch in slabs
    const int64_t kMaxY = 65535;
    for (int64_t y0 = 0; y0 < fy_total; y0 += kMaxY) {
        const int64_t ny = (fy_total - y0) < kMaxY ? (fy_total - y0) : kMaxY;
        const int64_t fbase = y0 * kFramesPerBlock;
        dim3 grid((NV + kVertsPerBlock - 1) / kVertsPerBlock, (unsigned)ny);
        skin_kernel<<<grid, kVertsPerBlock, 0, s>>>(d_vposed + (size_t)fbase * VPOSED_PITCH,
                                                    d_Askin + (size_t)fbase * NJ * 12, d_off + (size_t)fbase * 3,
                                                    m.d_wval, m.d_widx, m.nnz_groups, B - fbase,
                                                    d_verts + (size_t)fbase * NV * 3);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace prk
