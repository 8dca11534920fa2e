// C ABI of libposerisk_b200.so (include/poserisk_b200.h): model packing, workspace
// layout, chunking and stream ordering of the three kernel stages.
#include "prk_internal.h"

#include <atomic>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace prk {

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- optional per-stage CUDA-event timing (prk_profile_begin / prk_profile_end) --------
struct StageTimer {
    bool on = false;
    uint32_t mask = 0xF;                // stages that get event pairs (bit = stage id)
    std::vector<cudaEvent_t> pool;      // start/stop pairs
    std::vector<int> stage;             // stage id of pair i
    size_t used = 0;                    // pairs in use
};
static StageTimer g_timer;
static std::mutex g_timer_mu;                 // stage timing may be switched on while other threads launch
static std::atomic<bool> g_timer_on{false};   // the only thing the hot path looks at when timing is off
struct StageScope {
    cudaStream_t s;
    long idx = -1;
    StageScope(int stage_id, cudaStream_t stream) : s(stream) {
        if (!g_timer_on.load(std::memory_order_relaxed)) return;
        std::lock_guard<std::mutex> lk(g_timer_mu);
        if (!g_timer.on || g_timer.used >= 65536 || !((g_timer.mask >> stage_id) & 1u)) return;
        if (g_timer.used * 2 + 2 > g_timer.pool.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            g_timer.pool.push_back(a); g_timer.pool.push_back(b);
            g_timer.stage.push_back(stage_id);
        } else {
            g_timer.stage[g_timer.used] = stage_id;
        }
        idx = (long)g_timer.used++;
        cudaEventRecord(g_timer.pool[idx * 2], s);
    }
    ~StageScope() {
        if (idx < 0) return;
        std::lock_guard<std::mutex> lk(g_timer_mu);
        cudaEventRecord(g_timer.pool[idx * 2 + 1], s);
    }
};

static thread_local char t_detail[256] = "";
static void set_detail(const char* what, const char* msg) { snprintf(t_detail, sizeof t_detail, "%s: %s", what, msg); }

void comm_set_detail(const char* what, const char* msg) { set_detail(what, msg); }

static int cuda_fail(cudaError_t e, const char* what) {
    set_detail(what, cudaGetErrorString(e));
    return PRK_ERR_CUDA;
}
#define PRK_CUDA(expr)                                          \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);     \
    } while (0)

// ---- TMA descriptor ---------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows][cols] row-major tensor of bf16 (elem_bytes 2) or fp32 (elem_bytes 4), box = box_rows x box_cols with
// box_cols * elem_bytes == 128 (SWIZZLE_128B)
int encode_tmap_2d_ex(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint32_t box_cols, int elem_bytes, int swizzle128) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_detail("cuTensorMapEncodeTiled", "driver entry point not found"); return PRK_ERR_DRIVER; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(gptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[64];
        snprintf(msg, sizeof msg, "CUresult %d", (int)r);
        set_detail("cuTensorMapEncodeTiled", msg);
        return PRK_ERR_DRIVER;
    }
    return PRK_OK;
}
// fp32 [rows][cols] with an explicit row pitch in bytes (a multiple of 16), no swizzle: the vertex-store map
int encode_tmap_2d_f32_strided(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                               uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_detail("cuTensorMapEncodeTiled", "driver entry point not found"); return PRK_ERR_DRIVER; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[64];
        snprintf(msg, sizeof msg, "CUresult %d (vertex-store map)", (int)r);
        set_detail("cuTensorMapEncodeTiled", msg);
        return PRK_ERR_DRIVER;
    }
    return PRK_OK;
}
int encode_tmap_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                   uint32_t box_cols, int elem_bytes) {
    return encode_tmap_2d_ex(out, gptr, rows, cols, box_rows, box_cols, elem_bytes, 1);
}
int encode_tmap_2d_bf16(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                        uint32_t box_cols) {
    return encode_tmap_2d(out, gptr, rows, cols, box_rows, box_cols, 2);
}

// ---- host-side bf16 helpers --------------------------------------------------

static const int kSmplParents[NJ] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
static const int kSmplDfs[NJ] = {0, 1, 4, 7, 10, 2, 5, 8, 11, 3, 6, 9, 12, 15, 13, 16, 18, 20, 22, 14, 17, 19, 21, 23};

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- workspace layout ---------------------------------------------------------
constexpr int64_t kMaxSuper = 65536;   // frames per pose-chain + fused-kernel launch pair (A' + A_j scratch: 2.2 KB per frame)

struct Layout {
    int64_t S = 0;   // frames per launch pair (multiple of 128)
    size_t off_flags = 0, off_arows = 0, off_askin = 0, off_off = 0, total = 0;
};
static size_t align_up(size_t x) { return (x + 1023) & ~(size_t)1023; }

static Layout make_layout(int64_t S, bool mesh) {
    Layout L;
    L.S = S;
    size_t o = 0;
    L.off_flags = o; o += 1024;
    if (mesh) {   // v_posed never exists in memory: only the per-frame operands of the fused kernel
        L.off_arows = o;  o += align_up((size_t)S * FUSED_K * 2);
        L.off_askin = o;  o += align_up((size_t)S * NJ * 12 * 4);
        L.off_off = o;    o += align_up((size_t)S * 3 * 4);
    }
    L.total = o;
    return L;
}
static Layout want_layout(int64_t B, bool mesh) {
    int64_t S = round_up(B < 1 ? 1 : B, FUSED_BM);
    if (S > kMaxSuper) S = kMaxSuper;
    return make_layout(S, mesh);
}
// largest layout that fits in `bytes`
static bool fit_layout(int64_t B, bool mesh, size_t bytes, Layout& L) {
    L = want_layout(B, mesh);
    int64_t S = L.S;
    while (make_layout(S, mesh).total > bytes) {
        if (S <= FUSED_BM) return false;
        S = round_up(S / 2, FUSED_BM);
    }
    L = make_layout(S, mesh);
    return true;
}

// ---- which prk_pipeline_host call last used a workspace ------------------------------------
// The copy-in of a host call may run ahead of the caller's stream only when the PREVIOUS user of the same workspace
// was a host call of the same model, batch size and stream: then both calls agree on the staging layout and the
// model's own events order the two input sets.  Anything else (another batch size => other offsets, another model
// handle => other events, a device-API call on the same memory) makes the copy-in wait for the caller's stream.
struct HostChain { const void* ws = nullptr; const Model* m = nullptr; int64_t B = 0; cudaStream_t s = nullptr; };
static std::mutex g_chain_mu;
static HostChain g_chain[16];
static unsigned g_chain_next = 0;
// true when this call continues such a chain; records the call as the workspace's last user
static bool chain_continues(const void* ws, const Model* m, int64_t B, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_chain_mu);
    for (HostChain& c : g_chain)
        if (c.ws == ws) {
            const bool same = c.m == m && c.B == B && c.s == s;
            c.m = m; c.B = B; c.s = s;
            return same;
        }
    g_chain[g_chain_next++ % 16] = HostChain{ws, m, B, s};
    return false;
}
static void chain_break(const void* ws, const Model* m) {   // a device-API call touched `ws` / model `m` goes away
    std::lock_guard<std::mutex> lk(g_chain_mu);
    for (HostChain& c : g_chain)
        if ((ws && c.ws == ws) || (m && c.m == m)) c = HostChain{};
}

// ev_joints (optional) is recorded on `s` once every frame's joints are final, i.e. after the last
// pose-chain launch and long before the vertex kernels finish
// argument / workspace checks of a forward call, done before anything is launched
static int forward_check(const char* who, const Model* m, const float* d_pose, int center_idx, int64_t B,
                         const float* d_verts, int64_t vpitch, const float* d_joints, const void* ws, size_t ws_bytes, Layout& L) {
    if (!m || B < 0 || (B > 0 && (!d_pose || !d_joints)) || center_idx >= NJ) {
        set_detail(who, "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    if (B == 0) return PRK_OK;
    const bool mesh = d_verts != nullptr;
    if (mesh && (reinterpret_cast<uintptr_t>(d_verts) & 7)) {   // frame rows are written in 8- and 16-byte pieces
        set_detail(who, "d_verts must be 8-byte aligned");
        return PRK_ERR_INVALID_ARG;
    }
    if (mesh && vpitch != 0 && vpitch != NVC) {
        // padded vertex rows exist for the TMA store path only: 16-byte aligned rows, models with <= 4 weights per vertex
        if (vpitch < NVC || (vpitch & 3) || (reinterpret_cast<uintptr_t>(d_verts) & 15)) {
            set_detail(who, "verts_pitch must be 0 / 20670 or a multiple of 4 floats >= 20670 with a 16-byte aligned d_verts");
            return PRK_ERR_INVALID_ARG;
        }
        if (m->nnz_groups != 1) {
            set_detail(who, "a padded verts_pitch needs a model with at most 4 skinning weights per vertex");
            return PRK_ERR_UNSUPPORTED;
        }
    }
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 1023)) {
        set_detail(who, "workspace missing or not 1024-byte aligned");
        return PRK_ERR_WORKSPACE;
    }
    if (!fit_layout(B, mesh, ws_bytes, L)) { set_detail(who, "workspace too small"); return PRK_ERR_WORKSPACE; }
    return PRK_OK;
}

static int forward_impl(Model* m, const float* d_pose, const float* d_betas, const float* d_trans, int center_idx,
                        int64_t B, float* d_verts, int64_t vpitch, float* d_joints, void* ws, size_t ws_bytes, cudaStream_t s,
                        cudaEvent_t ev_joints = nullptr) {
    Layout L;
    if (vpitch == 0) vpitch = NVC;
    const int rc0 = forward_check("prk_smpl_forward", m, d_pose, center_idx, B, d_verts, vpitch, d_joints, ws, ws_bytes, L);
    if (rc0 != PRK_OK) return rc0;
    if (B == 0) return PRK_OK;
    const bool mesh = d_verts != nullptr;
    uint8_t* w = static_cast<uint8_t*>(ws);
    BatchFlags* d_flags = reinterpret_cast<BatchFlags*>(w + L.off_flags);
    uint16_t* d_arows = reinterpret_cast<uint16_t*>(w + L.off_arows);
    float* d_askin = reinterpret_cast<float*>(w + L.off_askin);
    float* d_off = reinterpret_cast<float*>(w + L.off_off);

    if (pose_chain_needs_flags(*m, d_betas, d_trans, center_idx)) PRK_CUDA(launch_batch_flags(d_betas, d_trans, B, d_flags, s));

    const int64_t S = mesh ? L.S : B;   // the joints-only path needs no scratch: one launch
    for (int64_t s0 = 0; s0 < B; s0 += S) {
        const int64_t ns = (B - s0) < S ? (B - s0) : S;
        {
            StageScope sc(0, s);
            PRK_CUDA(launch_pose_chain(*m, d_pose + s0 * 72, d_betas ? d_betas + s0 * NBETA : nullptr,
                                       d_trans ? d_trans + s0 * 3 : nullptr, d_flags, center_idx, ns, mesh, d_arows,
                                       d_askin, d_off, d_joints + s0 * 72, s));
        }
        if (ev_joints && s0 + S >= B) PRK_CUDA(cudaEventRecord(ev_joints, s));
        if (!mesh) continue;
        const int64_t rows_pad = round_up(ns, FUSED_BM);
        CUtensorMap tmA;
        int rc = encode_tmap_2d_bf16(&tmA, d_arows, (uint64_t)rows_pad, FUSED_K, FUSED_BM, 64);
        if (rc != PRK_OK) return rc;
        StageScope sc(1, s);
        PRK_CUDA(launch_fused(*m, tmA, rows_pad, d_askin, d_off, ns, d_verts + (size_t)s0 * vpitch, vpitch, s));
    }
    return PRK_OK;
}

}  // namespace prk

using namespace prk;

extern "C" {

int prk_abi_version(void) { return PRK_ABI_VERSION; }

const char* prk_strerror(int status) {
    switch (status) {
        case PRK_OK: return "ok";
        case PRK_ERR_INVALID_ARG: return "invalid argument";
        case PRK_ERR_CUDA: return "CUDA error";
        case PRK_ERR_WORKSPACE: return "workspace too small or misaligned";
        case PRK_ERR_UNSUPPORTED: return "unsupported configuration";
        case PRK_ERR_DRIVER: return "CUDA driver entry point unavailable or failed";
        case PRK_ERR_PEER: return "peer-memory exchange failed (a peer did not arrive, or its memory could not be mapped)";
        default: return "unknown status";
    }
}
const char* prk_last_error_detail(void) { return t_detail; }
uint64_t prk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int64_t prk_vposed_pitch(void) { return NVC; }

int prk_model_create(prk_model** out, int device, const float* vt, const float* sd, const float* pd, const float* jr,
                     const float* wt, const int32_t* parents, const float* betas) {
    if (!out || !vt || !sd || !pd || !jr || !wt || !parents) { set_detail("prk_model_create", "null argument"); return PRK_ERR_INVALID_ARG; }
    *out = nullptr;
    PRK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PRK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_detail("prk_model_create", "this library is built for sm_100a (B200) only");
        return PRK_ERR_UNSUPPORTED;
    }
    Model* m = new (std::nothrow) Model();
    if (!m) return PRK_ERR_INVALID_ARG;
    m->device = device;
    m->sm_count = prop.multiProcessorCount;

    // kinematic tree
    bool std_tree = true;
    m->pc.parents[0] = -1;
    for (int j = 1; j < NJ; ++j) {
        if (parents[j] < 0 || parents[j] >= j) {   // parents must precede children (smpl_layer.py:109-119)
            delete m;
            set_detail("prk_model_create", "kintree parent must have a smaller index than its child");
            return PRK_ERR_INVALID_ARG;
        }
        m->pc.parents[j] = parents[j];
        std_tree &= (parents[j] == kSmplParents[j]);
    }
    m->pc.standard_tree = std_tree ? 1 : 0;
    for (int k = 0; k < NBETA; ++k) m->pc.model_betas[k] = betas ? betas[k] : 0.0f;

    // folded joint regressor (double accumulation)
    for (int j = 0; j < NJ; ++j)
        for (int c = 0; c < 3; ++c) {
            double a = 0.0, d[NBETA] = {0};
            for (int v = 0; v < NV; ++v) {
                const double r = jr[(size_t)j * NV + v];
                if (r == 0.0) continue;
                a += r * vt[v * 3 + c];
                for (int k = 0; k < NBETA; ++k) d[k] += r * sd[((size_t)v * 3 + c) * NBETA + k];
            }
            m->pc.J_template[j * 3 + c] = (float)a;
            for (int k = 0; k < NBETA; ++k) m->pc.Jdirs[(j * 3 + c) * NBETA + k] = (float)d[k];
        }

    // K12 operand (prk_internal.h "K12 operand layout"): fp16 main part, e4m3 cross-term parts, fp16 shape / template parts,
    // everything times 2^S
    auto max_abs = [](const float* a, size_t n) {
        float mx = 0.f;
        for (size_t i = 0; i < n; ++i) { const float v = fabsf(a[i]); if (v > mx && std::isfinite(v)) mx = v; }
        return mx;
    };
    const float pd_max = max_abs(pd, (size_t)NVC * NPOSE), sd_max = max_abs(sd, (size_t)NVC * NBETA), vt_max = max_abs(vt, NVC);
    int S = 0;
    {   // max|posedirs| 2^S <= 2^14 (e4m3 range of the cross terms), max|shapedirs| 2^S <= 2^15 and max|v_template| 2^(S-15) <= 2^15
        // (fp16 range of sh and of u1)
        float lim = 1e30f;
        if (pd_max > 0.f) lim = fminf(lim, 16384.0f / pd_max);
        if (sd_max > 0.f) lim = fminf(lim, 32768.0f / sd_max);
        if (vt_max > 0.f) lim = fminf(lim, 1073741824.0f / vt_max);
        S = lim < 1e30f ? (int)floorf(log2f(lim)) : 0;
        if (S < 0) S = 0;
        if (S > 30) S = 30;
    }
    m->blend_scale_log2 = S;
    m->pc.rot_scale = ldexpf(1.0f, -S);
    auto f16 = [](float x) { const __half hv = __float2half_rn(x); uint16_t u; memcpy(&u, &hv, 2); return u; };
    auto f16f = [](uint16_t u) { __half hv; memcpy(&hv, &u, 2); return __half2float(hv); };
    auto e4m3 = [](float x) { return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3); };
    std::vector<uint16_t> B2((size_t)GEMM_N * FUSED_K, 0), B2row(FUSED_K);
    uint8_t* B2b = reinterpret_cast<uint8_t*>(B2.data());
    for (int n = 0; n < NVC; ++n) {
        uint16_t* row = B2row.data();
        uint8_t* rowb = reinterpret_cast<uint8_t*>(row);
        std::fill(B2row.begin(), B2row.end(), (uint16_t)0);
        for (int pos = 1; pos < NJ; ++pos) {
            const int j = std_tree ? kSmplDfs[pos] : pos;
            for (int e = 0; e < 9; ++e) {
                const int k = 9 * (pos - 1) + e;
                const float v = ldexpf(pd[(size_t)n * NPOSE + (j - 1) * 9 + e], S);
                const uint16_t hi = f16(v);
                row[k] = hi;
                rowb[FUSED_X_BYTE0 + k] = e4m3(ldexpf(v, -12));          // meets fp8((F - Fh) 2^12)
                rowb[FUSED_X_BYTE1 + k] = e4m3(v - f16f(hi));           // meets fp8(F)
            }
        }
        uint16_t* x = row + FUSED_COL_BETA;          // k-steps 26, 27 (prk_internal.h "K12 operand layout"): sh | u1 u2 u3 | 0 0 0,  sl | 0 x 6
        for (int bq = 0; bq < NBETA; ++bq) {
            const float v = ldexpf(sd[(size_t)n * NBETA + bq], S);
            x[bq] = f16(v);
            x[16 + bq] = f16(v - f16f(x[bq]));
        }
        float u = ldexpf(vt[n], S - 15);              // v_template 2^S = 2^15 (u1 + u2 + u3)
        for (int q = 0; q < 3; ++q) { x[NBETA + q] = f16(u); u -= f16f(x[NBETA + q]); }
        for (int k = 0; k < FUSED_K * 2; ++k) B2b[fused_b2_byte_index(n, k)] = rowb[k];   // into the pre-swizzled chunk images
    }

    // compacted skinning weights
    int mx = 0;
    for (int v = 0; v < NV; ++v) {
        int c = 0;
        for (int j = 0; j < NJ; ++j) c += wt[(size_t)v * NJ + j] != 0.0f;
        if (c > mx) mx = c;
    }
    m->max_weights = mx;
    m->nnz_groups = mx <= 4 ? 1 : (mx + 3) / 4;
    std::vector<float4> wv((size_t)m->nnz_groups * NV, make_float4(0, 0, 0, 0));
    std::vector<uint32_t> wi((size_t)m->nnz_groups * NV, 0);
    for (int v = 0; v < NV; ++v) {
        int c = 0;
        for (int j = 0; j < NJ; ++j) {
            const float x = wt[(size_t)v * NJ + j];
            if (x == 0.0f) continue;
            const int g = c >> 2, k = c & 3;
            float* f = reinterpret_cast<float*>(&wv[(size_t)g * NV + v]);
            f[k] = x;
            wi[(size_t)g * NV + v] |= (uint32_t)j << (8 * k);
            ++c;
        }
    }

    // the same pairs per 32-vertex tile for K12: [tile][group][32 x float4 weights | 32 x uint4 TMEM columns]
    std::vector<uint8_t> wp((size_t)FUSED_NT * m->nnz_groups * FUSED_WGROUP_BYTES, 0);
    for (int v = 0; v < NV; ++v) {
        const int tile = v / FUSED_VT, vl = v % FUSED_VT;
        for (int g = 0; g < m->nnz_groups; ++g) {
            uint8_t* base = &wp[((size_t)tile * m->nnz_groups + g) * FUSED_WGROUP_BYTES];
            memcpy(base + vl * 16, &wv[(size_t)g * NV + v], 16);
            const uint32_t id = wi[(size_t)g * NV + v];
            uint32_t cols[4];
            for (int qc = 0; qc < FUSED_WCOL_COPIES; ++qc) {
                for (int k = 0; k < 4; ++k) cols[k] = ((id >> (8 * k)) & 0xFFu) * 12u + ((uint32_t)(32 * qc) << 16);
                memcpy(base + FUSED_VT * 16 * (1 + qc) + vl * 16, cols, 16);
            }
        }
    }

    cudaError_t e;
#define PRK_M(expr) do { e = (expr); if (e != cudaSuccess) { prk_model_destroy(m); return cuda_fail(e, #expr); } } while (0)
    PRK_M(cudaMalloc(&m->d_B2, B2.size() * 2));
    PRK_M(cudaMemcpy(m->d_B2, B2.data(), B2.size() * 2, cudaMemcpyHostToDevice));
    {   // raw (un-swizzled) view of the pre-swizzled chunk images: a box = 48 rows of 128 bytes = one CTA's half of a chunk
        const int rc = encode_tmap_2d_ex(&m->tmap_B2, m->d_B2, (uint64_t)FUSED_NT * FUSED_B_CHUNKS * FUSED_BN, 64, FUSED_BN / 2, 64, 2, 0);
        if (rc != PRK_OK) { prk_model_destroy(m); return rc; }
    }
    PRK_M(cudaMalloc(&m->d_wpack, wp.size()));
    PRK_M(cudaMemcpy(m->d_wpack, wp.data(), wp.size(), cudaMemcpyHostToDevice));
    {
        std::vector<float> jc(72 + 720 + NBETA + 1);
        memcpy(jc.data(), m->pc.J_template, 72 * 4);
        for (int i = 0; i < 72; ++i)
            for (int k = 0; k < NBETA; ++k) jc[72 + k * 72 + i] = m->pc.Jdirs[i * NBETA + k];   // beta-major for the lane-per-joint kernel
        memcpy(jc.data() + 792, m->pc.model_betas, NBETA * 4);
        jc[802] = m->pc.rot_scale;
        PRK_M(cudaMalloc(&m->d_Jc, jc.size() * 4));
        PRK_M(cudaMemcpy(m->d_Jc, jc.data(), jc.size() * 4, cudaMemcpyHostToDevice));
    }
    PRK_M(cudaStreamCreateWithFlags(&m->s_score, cudaStreamNonBlocking));
    PRK_M(cudaEventCreateWithFlags(&m->ev_in, cudaEventDisableTiming));
    PRK_M(cudaEventCreateWithFlags(&m->ev_score, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) PRK_M(cudaEventCreateWithFlags(&m->ev_set_free[i], cudaEventDisableTiming));
    PRK_M(cudaStreamCreateWithFlags(&m->s_in, cudaStreamNonBlocking));
    PRK_M(cudaStreamCreateWithFlags(&m->s_out, cudaStreamNonBlocking));
    PRK_M(cudaEventCreateWithFlags(&m->ev_h2d, cudaEventDisableTiming));
    PRK_M(cudaEventCreateWithFlags(&m->ev_joints, cudaEventDisableTiming));
    PRK_M(cudaEventCreateWithFlags(&m->ev_out, cudaEventDisableTiming));
#undef PRK_M
    *out = m;
    return PRK_OK;
}

void prk_model_destroy(prk_model* model) {
    Model* m = model;
    if (!m) return;
    chain_break(nullptr, m);
    if (m->device >= 0) cudaSetDevice(m->device);
    cudaFree(m->d_B2); cudaFree(m->d_wpack);
    cudaFree(m->d_Jc);
    if (m->s_score) { cudaStreamSynchronize(m->s_score); cudaStreamDestroy(m->s_score); }
    if (m->ev_in) cudaEventDestroy(m->ev_in);
    if (m->ev_score) cudaEventDestroy(m->ev_score);
    if (m->s_in) { cudaStreamSynchronize(m->s_in); cudaStreamDestroy(m->s_in); }
    if (m->s_out) { cudaStreamSynchronize(m->s_out); cudaStreamDestroy(m->s_out); }
    if (m->ev_h2d) cudaEventDestroy(m->ev_h2d);
    if (m->ev_joints) cudaEventDestroy(m->ev_joints);
    if (m->ev_out) cudaEventDestroy(m->ev_out);
    for (int i = 0; i < 2; ++i) if (m->ev_set_free[i]) cudaEventDestroy(m->ev_set_free[i]);
    delete m;
}
int prk_model_device(const prk_model* model) { return model ? model->device : -1; }
int prk_model_max_weights(const prk_model* model) { return model ? model->max_weights : 0; }

size_t prk_workspace_bytes(const prk_model*, int64_t B, uint32_t flags) {
    return want_layout(B, !(flags & PRK_FLAG_JOINTS_ONLY)).total;
}

int prk_smpl_forward(prk_model* model, const float* d_pose, const float* d_betas, const float* d_trans,
                     int center_idx, int64_t B, float* d_verts, int64_t verts_pitch, float* d_joints, void* ws, size_t ws_bytes,
                     void* stream) {
    Model* m = model;
    if (!m) { set_detail("prk_smpl_forward", "null model"); return PRK_ERR_INVALID_ARG; }
    chain_break(ws, nullptr);
    PRK_CUDA(cudaSetDevice(m->device));
    return forward_impl(m, d_pose, d_betas, d_trans, center_idx, B, d_verts, verts_pitch, d_joints, ws, ws_bytes,
                        static_cast<cudaStream_t>(stream));
}

// --debug_joints list -> by-value kernel parameter (no device table, no allocation, no synchronisation)
static bool make_debug_slots(const int32_t* h_ids, int n_debug, DebugSlots& d) {
    d.mask = 0;
    for (int j = 0; j < NJ; ++j) d.slot[j] = -1;
    for (int k = 0; k < n_debug; ++k) {
        const int j = h_ids[k];
        if (j < 0 || j >= NJ || (d.mask & (1u << j))) return false;
        d.mask |= 1u << j;
        d.slot[j] = (int8_t)k;
    }
    return true;
}

int prk_score_pose(const void* d_pose, int pose_dtype, const prk_addinfo* d_info, int32_t n_tracks, const int32_t* d_track,
                   int64_t B, uint32_t which, prk_score_rec* d_out, double* d_euler_out, const int32_t* h_debug_joint_ids,
                   int n_debug, void* stream) {
    if (B < 0 || (B > 0 && (!d_pose || !d_info || !d_out)) || n_tracks < 1 ||
        (pose_dtype != PRK_DTYPE_F32 && pose_dtype != PRK_DTYPE_F64) ||
        !(which & 3u) || n_debug < 0 || n_debug > NJ || (n_debug > 0 && (!d_euler_out || !h_debug_joint_ids))) {
        set_detail("prk_score_pose", "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    DebugSlots dbg;
    if (!make_debug_slots(h_debug_joint_ids, n_debug, dbg)) {
        set_detail("prk_score_pose", "bad or duplicate debug joint id");
        return PRK_ERR_INVALID_ARG;
    }
    PRK_CUDA(launch_score_pose(d_pose, pose_dtype, d_info, n_tracks, d_track, B, which, d_out,
                               n_debug > 0 ? d_euler_out : nullptr, dbg, n_debug, static_cast<cudaStream_t>(stream)));
    return PRK_OK;
}

int prk_score_euler(const double* d_euler, const prk_addinfo* d_info, int32_t n_tracks, const int32_t* d_track, int64_t B,
                    uint32_t which, prk_score_rec* d_out, void* stream) {
    if (B < 0 || (B > 0 && (!d_euler || !d_info || !d_out)) || n_tracks < 1 || !(which & 3u)) {
        set_detail("prk_score_euler", "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    PRK_CUDA(launch_score_euler(d_euler, d_info, n_tracks, d_track, B, which, d_out, static_cast<cudaStream_t>(stream)));
    return PRK_OK;
}

int prk_euler(const void* d_pose, int pose_dtype, int64_t n_rot, double* d_euler, uint8_t* d_bad, void* stream) {
    if (n_rot < 0 || (n_rot > 0 && (!d_pose || !d_euler)) || (pose_dtype != PRK_DTYPE_F32 && pose_dtype != PRK_DTYPE_F64)) {
        set_detail("prk_euler", "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    PRK_CUDA(launch_euler(d_pose, pose_dtype, n_rot, d_euler, d_bad, static_cast<cudaStream_t>(stream)));
    return PRK_OK;
}

int prk_rot_to_angle(const void* d_rotmat, int dtype, int64_t n_rot, void* d_rvec, uint8_t* d_bad, void* stream) {
    if (n_rot < 0 || (n_rot > 0 && (!d_rotmat || !d_rvec)) || (dtype != PRK_DTYPE_F32 && dtype != PRK_DTYPE_F64)) {
        set_detail("prk_rot_to_angle", "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    PRK_CUDA(launch_rot_to_angle(d_rotmat, dtype, n_rot, d_rvec, d_bad, static_cast<cudaStream_t>(stream)));
    return PRK_OK;
}

int prk_pipeline(prk_model* model, const float* d_pose, const float* d_betas, const float* d_trans, int center_idx,
                 const prk_addinfo* d_info, int32_t n_tracks, const int32_t* d_track, int64_t B, float* d_verts,
                 int64_t verts_pitch, float* d_joints, prk_score_rec* d_scores, double* d_euler_out, const int32_t* h_debug_joint_ids,
                 int n_debug, prk_comm* comm_scores, prk_comm* comm_euler, int64_t frame_offset, void* ws, size_t ws_bytes,
                 void* stream) {
    Model* m = model;
    // every check happens before the first launch: a refused call leaves the stream and the outputs untouched
    if (!m || !d_info || n_tracks < 1 || (B > 0 && !d_scores) || n_debug < 0 || n_debug > NJ ||
        (n_debug > 0 && (!d_euler_out || !h_debug_joint_ids))) {
        set_detail("prk_pipeline", "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    DebugSlots dbg;
    if (!make_debug_slots(h_debug_joint_ids, n_debug, dbg)) {
        set_detail("prk_pipeline", "bad or duplicate debug joint id");
        return PRK_ERR_INVALID_ARG;
    }
    {
        Layout L;
        const int rc = forward_check("prk_pipeline", m, d_pose, center_idx, B, d_verts, verts_pitch, d_joints, ws, ws_bytes, L);
        if (rc != PRK_OK) return rc;
    }
    if (comm_euler && n_debug == 0) comm_euler = nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (B == 0) {   // an empty shard still takes part in the exchange
        int rc = PRK_OK;
        if (comm_scores) rc = prk_allgather_rows(comm_scores, nullptr, 0, frame_offset, sizeof(prk_score_rec), nullptr, s);
        if (rc == PRK_OK && comm_euler) rc = prk_allgather_rows(comm_euler, nullptr, 0, frame_offset, n_debug * 24, nullptr, s);
        return rc;
    }
    chain_break(ws, nullptr);
    PRK_CUDA(cudaSetDevice(m->device));
    // scoring only reads the pose: it runs on the model's scoring stream, beside the mesh path
    cudaStream_t ss = m->s_score;
    PRK_CUDA(cudaEventRecord(m->ev_in, s));
    PRK_CUDA(cudaStreamWaitEvent(ss, m->ev_in, 0));
    cudaError_t e;
    {
        StageScope sc(3, ss);
        e = launch_score_pose(d_pose, PRK_DTYPE_F32, d_info, n_tracks, d_track, B, PRK_SCORE_REBA | PRK_SCORE_RULA,
                              d_scores, n_debug > 0 ? d_euler_out : nullptr, dbg, n_debug, ss);
    }
    // the caller's stream joins the scoring stream where the RECORDS are final, whatever happens next: nothing of this
    // call's own outputs is still being written once `stream` has passed it, also when the call reports an error
    cudaError_t e2 = cudaEventRecord(m->ev_score, ss);
    // multi-GPU: the records (and debug Euler rows) go to every rank's gather buffer straight from the scoring stream,
    // i.e. underneath the vertex kernel launched on `s` below.  The exchange waits for the peers, so it is NOT joined
    // into `stream`: prk_comm_wait does that where the gathered rows are needed (a rank may run one call ahead).
    int rc = PRK_OK;
    if (e == cudaSuccess && comm_scores)
        rc = prk_allgather_rows_impl(comm_scores, d_scores, B, frame_offset, sizeof(prk_score_rec), nullptr, ss, false);
    if (e == cudaSuccess && rc == PRK_OK && comm_euler)
        rc = prk_allgather_rows_impl(comm_euler, d_euler_out, B, frame_offset, n_debug * 24, nullptr, ss, false);
    if (e == cudaSuccess && e2 == cudaSuccess && rc == PRK_OK) {
        set_fused_exchange_hint(comm_scores != nullptr || comm_euler != nullptr);
        rc = forward_impl(m, d_pose, d_betas, d_trans, center_idx, B, d_verts, verts_pitch, d_joints, ws, ws_bytes, s);
        set_fused_exchange_hint(false);
    }
    if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(s, m->ev_score, 0);
    if (e != cudaSuccess) return cuda_fail(e, "score_pose_kernel");
    if (e2 != cudaSuccess) return cuda_fail(e2, "join of the scoring stream");
    return rc;
}

// host-staging layout in front of the device workspace: two input sets (the copy-in of call i+1
// runs under the kernels of call i), one output set (joints and scores are copied out while the
// vertex kernel of the same call is still running)
struct HostStage { size_t in_set, pose, betas, trans, info, track, joints, scores, inner, total; };
static HostStage host_stage(int64_t B, int32_t n_tracks, size_t inner_bytes) {
    HostStage h;
    size_t o = 0;
    h.pose = o;   o += align_up((size_t)B * 72 * 4);
    h.betas = o;  o += align_up((size_t)B * NBETA * 4);
    h.trans = o;  o += align_up((size_t)B * 3 * 4);
    h.info = o;   o += align_up((size_t)(n_tracks < 1 ? 1 : n_tracks) * sizeof(prk_addinfo));
    h.track = o;  o += align_up((size_t)B * 4);
    h.in_set = o;                      // bytes of one input set; the second set follows
    o *= 2;
    h.joints = o; o += align_up((size_t)B * 72 * 4);
    h.scores = o; o += align_up((size_t)B * sizeof(prk_score_rec));
    h.inner = o;  o += inner_bytes;
    h.total = o;
    return h;
}

size_t prk_host_workspace_bytes(const prk_model* model, int64_t B, uint32_t flags) {
    // room for up to 4096 tracks of add_info
    return host_stage(B, 4096, prk_workspace_bytes(model, B, flags)).total;
}

size_t prk_host_scores_offset(const prk_model*, int64_t B) { return host_stage(B, 4096, 0).scores; }

int prk_pipeline_host(prk_model* model, const float* h_pose, const float* h_betas, const float* h_trans,
                      int center_idx, const prk_addinfo* h_info, int32_t n_tracks, const int32_t* h_track, int64_t B,
                      float* d_verts, int64_t verts_pitch, float* h_joints, prk_score_rec* h_scores, prk_comm* comm_scores,
                      int64_t frame_offset, void* ws, size_t ws_bytes, void* stream) {
    Model* m = model;
    if (!m || B < 0 || !h_info || n_tracks < 1 || n_tracks > 4096 || (B > 0 && (!h_pose || !h_scores))) {
        set_detail("prk_pipeline_host", "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    if (B == 0)
        return comm_scores ? prk_allgather_rows(comm_scores, nullptr, 0, frame_offset, sizeof(prk_score_rec), nullptr, stream) : PRK_OK;
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 1023)) { set_detail("prk_pipeline_host", "workspace missing or misaligned"); return PRK_ERR_WORKSPACE; }
    PRK_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const HostStage h = host_stage(B, 4096, 0);
    if (ws_bytes <= h.inner) { set_detail("prk_pipeline_host", "workspace too small"); return PRK_ERR_WORKSPACE; }
    {
        Layout L;
        if (center_idx >= NJ || (d_verts && (reinterpret_cast<uintptr_t>(d_verts) & 7)) ||
            (d_verts && verts_pitch != 0 && verts_pitch != NVC &&
             (verts_pitch < NVC || (verts_pitch & 3) || (reinterpret_cast<uintptr_t>(d_verts) & 15) || m->nnz_groups != 1))) {
            set_detail("prk_pipeline_host", "invalid argument");
            return PRK_ERR_INVALID_ARG;
        }
        if (!fit_layout(B, d_verts != nullptr, ws_bytes - h.inner, L)) { set_detail("prk_pipeline_host", "workspace too small"); return PRK_ERR_WORKSPACE; }
    }
    const int set = (int)(m->host_calls++ & 1);
    uint8_t* w = static_cast<uint8_t*>(ws);
    uint8_t* wi = w + (size_t)set * h.in_set;
    float* d_pose = reinterpret_cast<float*>(wi + h.pose);
    float* d_betas = h_betas ? reinterpret_cast<float*>(wi + h.betas) : nullptr;
    float* d_trans = h_trans ? reinterpret_cast<float*>(wi + h.trans) : nullptr;
    prk_addinfo* d_info = reinterpret_cast<prk_addinfo*>(wi + h.info);
    int32_t* d_track = h_track ? reinterpret_cast<int32_t*>(wi + h.track) : nullptr;
    float* d_joints = reinterpret_cast<float*>(w + h.joints);
    prk_score_rec* d_scores = reinterpret_cast<prk_score_rec*>(w + h.scores);

    // ---- copy-in stream: inputs of this call land while the kernels of the previous call run (the
    // host buffers must be ready when the call is made, as for any host argument).  The set was
    // last read by the call before the previous one (ev_set_free).  Only a chain of host calls of
    // the same model, batch size and stream on the same workspace shares one staging layout and one
    // set of events (chain_continues): after anything else the copy waits for the work already
    // queued on the caller's stream.
    PRK_CUDA(cudaEventRecord(m->ev_in, s));
    if (!chain_continues(ws, m, B, s)) PRK_CUDA(cudaStreamWaitEvent(m->s_in, m->ev_in, 0));
    PRK_CUDA(cudaStreamWaitEvent(m->s_in, m->ev_set_free[set], 0));
    PRK_CUDA(cudaMemcpyAsync(d_pose, h_pose, (size_t)B * 72 * 4, cudaMemcpyHostToDevice, m->s_in));
    if (h_betas) PRK_CUDA(cudaMemcpyAsync(d_betas, h_betas, (size_t)B * NBETA * 4, cudaMemcpyHostToDevice, m->s_in));
    if (h_trans) PRK_CUDA(cudaMemcpyAsync(d_trans, h_trans, (size_t)B * 3 * 4, cudaMemcpyHostToDevice, m->s_in));
    PRK_CUDA(cudaMemcpyAsync(d_info, h_info, (size_t)n_tracks * sizeof(prk_addinfo), cudaMemcpyHostToDevice, m->s_in));
    if (h_track) PRK_CUDA(cudaMemcpyAsync(d_track, h_track, (size_t)B * 4, cudaMemcpyHostToDevice, m->s_in));
    PRK_CUDA(cudaEventRecord(m->ev_h2d, m->s_in));

    // ---- scoring on its own stream, pose chain + vertex kernels on the caller's stream
    cudaStream_t ss = m->s_score;
    PRK_CUDA(cudaStreamWaitEvent(ss, m->ev_h2d, 0));
    // the staged score records of the previous call may still be read by work the caller queued
    // after it (prk_host_scores_offset): scoring starts where the caller's stream stands now
    PRK_CUDA(cudaStreamWaitEvent(ss, m->ev_in, 0));
    {
        StageScope sc(3, ss);
        PRK_CUDA(launch_score_pose(d_pose, PRK_DTYPE_F32, d_info, n_tracks, d_track, B, PRK_SCORE_REBA | PRK_SCORE_RULA,
                                   d_scores, nullptr, DebugSlots{}, 0, ss));
    }
    PRK_CUDA(cudaEventRecord(m->ev_score, ss));
    if (comm_scores) {   // multi-GPU: all-gather of the records underneath the vertex kernel; joined by prk_comm_wait
        const int rcg = prk_allgather_rows_impl(comm_scores, d_scores, B, frame_offset, sizeof(prk_score_rec), nullptr, ss, false);
        if (rcg != PRK_OK) { cudaStreamWaitEvent(s, m->ev_score, 0); chain_break(ws, nullptr); return rcg; }
    }
    PRK_CUDA(cudaStreamWaitEvent(s, m->ev_h2d, 0));
    PRK_CUDA(cudaStreamWaitEvent(s, m->ev_out, 0));         // d_joints of the previous call has been copied out
    set_fused_exchange_hint(comm_scores != nullptr);
    int rc = forward_impl(m, d_pose, d_betas, d_trans, center_idx, B, d_verts, verts_pitch, d_joints, w + h.inner, ws_bytes - h.inner, s,
                          m->ev_joints);
    set_fused_exchange_hint(false);
    if (rc != PRK_OK) {                                     // the caller's stream still joins what was launched
        cudaStreamWaitEvent(s, m->ev_score, 0);
        chain_break(ws, nullptr);
        return rc;
    }

    // ---- copy-out stream: joints and scores leave under the vertex kernel of this call
    PRK_CUDA(cudaStreamWaitEvent(m->s_out, m->ev_joints, 0));
    PRK_CUDA(cudaStreamWaitEvent(m->s_out, m->ev_score, 0));
    // The input set is free from HERE: its only readers are the flags / pose-chain kernels (done at ev_joints) and the scoring
    // kernel (ev_score); the vertex kernel reads the inner workspace.  Recording the event here and not at the end of the call
    // lets the copy-in of call i+2 start under the vertex kernel of call i -- a whole step earlier -- which is what the host
    // path needs when 8 ranks share the host's links (copies alone: 120 us of a 147 us step, scripts/pcie_scaling.py).
    PRK_CUDA(cudaEventRecord(m->ev_set_free[set], m->s_out));
    if (h_joints) PRK_CUDA(cudaMemcpyAsync(h_joints, d_joints, (size_t)B * 72 * 4, cudaMemcpyDeviceToHost, m->s_out));
    PRK_CUDA(cudaMemcpyAsync(h_scores, d_scores, (size_t)B * sizeof(prk_score_rec), cudaMemcpyDeviceToHost, m->s_out));
    PRK_CUDA(cudaEventRecord(m->ev_out, m->s_out));
    // completion stays ordered on the caller's stream: it joins the copy-out and the scoring stream
    PRK_CUDA(cudaStreamWaitEvent(s, m->ev_out, 0));
    return PRK_OK;
}

int prk_debug_unit_range(int64_t n_units, int64_t n_ranges, int switch_cost16, int64_t k, int64_t* u0, int64_t* u1) {
    if (n_units < 0 || n_ranges < 1 || switch_cost16 < 0 || k < 0 || k >= n_ranges || !u0 || !u1 || (n_units % FUSED_NT) != 0) {
        set_detail("prk_debug_unit_range", "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    *u0 = fused_first_unit(k, n_ranges, n_units, switch_cost16);
    *u1 = fused_first_unit(k + 1, n_ranges, n_units, switch_cost16);
    return PRK_OK;
}

int prk_score_histogram(const prk_score_rec* d_scores, int64_t B, uint32_t which, unsigned long long* d_hist,
                        void* stream) {
    if (B < 0 || !d_hist || (B > 0 && !d_scores) || (which != PRK_SCORE_REBA && which != PRK_SCORE_RULA)) {
        set_detail("prk_score_histogram", "invalid argument");
        return PRK_ERR_INVALID_ARG;
    }
    PRK_CUDA(launch_score_hist(d_scores, B, which, d_hist, static_cast<cudaStream_t>(stream)));
    return PRK_OK;
}

int prk_profile_begin_stages(uint32_t stage_mask) {
    std::lock_guard<std::mutex> lk(g_timer_mu);
    g_timer.on = true;
    g_timer.mask = stage_mask & 0xFu;
    g_timer.used = 0;
    g_timer_on.store(true);
    return PRK_OK;
}
int prk_profile_begin(void) { return prk_profile_begin_stages(0xFu); }

int prk_profile_end(double* ms_out, int64_t* launches_out) {
    std::lock_guard<std::mutex> lk(g_timer_mu);
    g_timer_on.store(false);
    g_timer.on = false;
    for (int k = 0; k < 4; ++k) { if (ms_out) ms_out[k] = 0.0; if (launches_out) launches_out[k] = 0; }
    for (size_t i = 0; i < g_timer.used; ++i) {
        PRK_CUDA(cudaEventSynchronize(g_timer.pool[i * 2 + 1]));
        float ms = 0.f;
        PRK_CUDA(cudaEventElapsedTime(&ms, g_timer.pool[i * 2], g_timer.pool[i * 2 + 1]));
        const int st = g_timer.stage[i];
        if (st >= 0 && st < 4) { if (ms_out) ms_out[st] += ms; if (launches_out) launches_out[st] += 1; }
    }
    g_timer.used = 0;
    return PRK_OK;
}

int prk_debug_blend(prk_model* model, const float* d_pose, const float* d_betas, int64_t B, float* d_vposed,
                    int use_simt, void* ws, size_t ws_bytes, void* stream) {
    Model* m = model;
    if (!m || B <= 0 || !d_pose || !d_vposed || !ws) { set_detail("prk_debug_blend", "invalid argument"); return PRK_ERR_INVALID_ARG; }
    PRK_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t rows_pad = round_up(B, FUSED_BM);
    // scratch: A' rows, A_j tiles, off, joints
    size_t o = 0;
    const size_t o_arows = o; o += align_up((size_t)rows_pad * FUSED_K * 2);
    const size_t o_askin = o; o += align_up((size_t)rows_pad * NJ * 12 * 4);
    const size_t o_off = o;   o += align_up((size_t)rows_pad * 3 * 4);
    const size_t o_j = o;     o += align_up((size_t)rows_pad * 72 * 4);
    const size_t o_flags = o; o += 1024;
    if (o > ws_bytes || (reinterpret_cast<uintptr_t>(ws) & 1023)) { set_detail("prk_debug_blend", "workspace too small or misaligned"); return PRK_ERR_WORKSPACE; }
    uint8_t* w = static_cast<uint8_t*>(ws);
    uint16_t* d_arows = reinterpret_cast<uint16_t*>(w + o_arows);
    float* d_askin = reinterpret_cast<float*>(w + o_askin);
    float* d_off = reinterpret_cast<float*>(w + o_off);
    BatchFlags* d_flags = reinterpret_cast<BatchFlags*>(w + o_flags);
    PRK_CUDA(cudaMemsetAsync(d_arows, 0, (size_t)rows_pad * FUSED_K * 2, s));
    if (pose_chain_needs_flags(*m, d_betas, nullptr, -1)) PRK_CUDA(launch_batch_flags(d_betas, nullptr, B, d_flags, s));
    PRK_CUDA(launch_pose_chain(*m, d_pose, d_betas, nullptr, d_flags, -1, B, true, d_arows, d_askin, d_off,
                               reinterpret_cast<float*>(w + o_j), s));
    if (use_simt) {
        PRK_CUDA(launch_blend_simt(*m, d_arows, B, d_vposed, s));
    } else {
        // the product kernel with identity skinning transforms and no offset: vertices = v_posed
        PRK_CUDA(launch_identity_askin(*m, d_askin, d_off, rows_pad, s));
        CUtensorMap tmA;
        int rc = encode_tmap_2d_bf16(&tmA, d_arows, (uint64_t)rows_pad, FUSED_K, FUSED_BM, 64);
        if (rc != PRK_OK) return rc;
        PRK_CUDA(launch_fused(*m, tmA, rows_pad, d_askin, d_off, B, d_vposed, NVC, s));
    }
    return PRK_OK;
}

}  // extern "C"
