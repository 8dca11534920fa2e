/*
 * poserisk_b200.h -- C ABI of the B200-native PoseRisk body-model -> risk-score path.
 *
 * The reference (hygenie1228/PoseRisk_RELEASE) has no FFI: its boundary for this
 * path is a Python call surface.  Each entry point below names the reference
 * interface it stands behind (paths relative to the reference root).  The Python
 * drop-ins in poserisk_release_b200/ (SMPL_Layer, REBA, RULA, get_joint_cam,
 * axis_angle_to_euler_angle) bind these symbols with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; every `d_` pointer is caller-owned DEVICE memory on the
 *    model's device, every `h_` pointer is HOST memory;
 *  - nothing is allocated inside a hot call: scratch comes from the caller's
 *    workspace (prk_workspace_bytes);
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *  - no exceptions cross the ABI: int status, 0 = PRK_OK, text via prk_strerror;
 *  - a handle belongs to one device; use one host thread per handle or lock.
 */
#ifndef POSERISK_B200_H
#define POSERISK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define PRK_ABI_VERSION 2

#define PRK_NUM_VERTS 6890
#define PRK_NUM_JOINTS 24
#define PRK_NUM_BETAS 10
#define PRK_NUM_POSE_FEATS 207

enum prk_status {
    PRK_OK = 0,
    PRK_ERR_INVALID_ARG = 1,
    PRK_ERR_CUDA = 2,
    PRK_ERR_WORKSPACE = 3,   /* workspace too small or misaligned */
    PRK_ERR_UNSUPPORTED = 4,
    PRK_ERR_DRIVER = 5,      /* cuTensorMapEncodeTiled unavailable / failed */
    PRK_ERR_PEER = 6         /* peer-memory exchange: a peer did not arrive in time / handle could not be opened */
};

/* additional_information.json (example/additional_information.json:1-25), one per track.
 * reba[7]: Legs_bilateral_weight_bearing/walking, Sitting, Load/Force Score,
 *          Arm_supported_leaning_L, Arm_supported_leaning_R, Coupling, Activity_Score
 *          (the keys lib/utils/reba.py:59,64,69,113,190,221,240 read)
 * rula[9]: Arm_supported_leaning_L, Arm_supported_leaning_R, A_Muscle_use_L,
 *          A_Muscle_use_R, A_Load/Force_L, A_Load/Force_R, Legs_bilateral_weight_bearing,
 *          B_Muscle_use, B_Load/Force   (lib/utils/rula.py:75-76,81,151,177,196) */
typedef struct prk_addinfo {
    int32_t reba[7];
    int32_t rula[9];
} prk_addinfo;

/* One scored frame, 32 bytes.  Holds what REBA.__call__ / RULA.__call__ return per
 * frame (`score`, flattened `log_score`; lib/utils/reba.py:71-75, rula.py:88-92).
 * reba_parts: trunk, neck, leg, upper_arm L,R, lower_arm L,R, wrist L,R
 * rula_parts: upper_arm L,R, lower_arm L,R, wrist L,R, wrist_twist L,R, neck, trunk, leg
 * flags bit0: a scored joint's rotation matrix is not finite, i.e. the reference
 *             would stop at assert(isRotationMatrix(R)) (lib/utils/coord_utils.py:70).
 * flags bit1: the frame's track id was outside [0, n_tracks): the reference would raise an
 *             IndexError picking that person's additional information; the frame is scored
 *             with track 0's values so that no out-of-bounds read happens. */
typedef struct prk_score_rec {
    int16_t reba_score;
    int16_t rula_score;
    uint8_t reba_parts[9];
    uint8_t rula_parts[11];
    uint8_t flags;
    uint8_t pad[7];
} prk_score_rec;

#define PRK_SCORE_REBA 1u
#define PRK_SCORE_RULA 2u

/* pose element type of the scoring front-end: cv2.Rodrigues returns a matrix of the
 * input's dtype (lib/utils/coord_utils.py:86), so float32 poses give float32-rounded R. */
#define PRK_DTYPE_F32 0
#define PRK_DTYPE_F64 1

/* prk_workspace_bytes / prk_smpl_forward flags */
#define PRK_FLAG_JOINTS_ONLY 1u   /* no vertex output: blend GEMM and skinning are skipped */

typedef struct prk_model prk_model;
struct prk_comm;   /* multi-GPU exchange handle, see below */

int prk_abi_version(void);
const char* prk_strerror(int status);
/* text of the last CUDA/driver error seen by this thread's calls ("" if none) */
const char* prk_last_error_detail(void);

/* SMPL_Layer.__init__ (lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py:15-63): takes the
 * arrays that constructor registers as buffers (HOST pointers, float32, C order) and builds
 * the device-side operands: split-precision (fp16 + e4m3, scaled by 2^S) blend matrix for the tcgen05 GEMM,
 * compacted skinning weights, folded joint regressor.
 *   h_v_template [6890*3]      th_v_template   (:46-48)
 *   h_shapedirs  [6890*3*10]   th_shapedirs    (:42-43)
 *   h_posedirs   [6890*3*207]  th_posedirs     (:44-45)
 *   h_J_regressor[24*6890]     th_J_regressor  (:49-51, dense)
 *   h_weights    [6890*24]     th_weights      (:52-53)
 *   h_parents    [24]          kintree_table[0] (:60-62); entry 0 is ignored
 *   h_betas      [10] or NULL  th_betas        (:40-41); NULL = zeros */
int prk_model_create(prk_model** out, int device, const float* h_v_template,
                     const float* h_shapedirs, const float* h_posedirs,
                     const float* h_J_regressor, const float* h_weights,
                     const int32_t* h_parents, const float* h_betas);
void prk_model_destroy(prk_model* model);
int prk_model_device(const prk_model* model);
/* max non-zero skinning weights per vertex found at create time (4 for SMPL) */
int prk_model_max_weights(const prk_model* model);

/* Scratch needed by prk_smpl_forward / prk_pipeline for a batch of B frames.  Any size
 * >= prk_workspace_bytes(model, 1, flags) works: frames are processed in chunks that fit. */
size_t prk_workspace_bytes(const prk_model* model, int64_t B, uint32_t flags);

/* SMPL_Layer.forward(th_pose_axisang, th_betas, th_trans) (smpl_layer.py:65-158).
 *   d_pose  [B*72] float32 axis-angle
 *   d_betas [B*10] or NULL; NULL or an all-zero batch selects the model's betas (:87-91)
 *   d_trans [B*3]  or NULL; NULL or an all-zero batch: no translation, and the
 *           `center_idx` joint is subtracted when center_idx >= 0 (:148-152)
 *   center_idx  -1 = None
 *   d_verts [B*6890*3] float32, or NULL for the joints-only path
 *   verts_pitch  floats from one frame's vertices to the next: 0 or 20670 = the reference's dense (B, 6890, 3)
 *           layout; a multiple of 4 >= 20670 (e.g. PRK_VERTS_PITCH_ALIGNED = 20672) with a 16-byte aligned d_verts
 *           = 16-byte aligned rows, which leave through bulk tensor stores (faster; models with <= 4 weights per vertex;
 *           the two floats that complete a row's 82,680 bytes to a multiple of 16 may be overwritten)
 *   d_joints[B*24*3]  float32 (chain translations, :145)
 * Both whole-batch tests are evaluated on the device (no host sync). */
#define PRK_VERTS_PITCH_ALIGNED 20672
int prk_smpl_forward(prk_model* model, const float* d_pose, const float* d_betas,
                     const float* d_trans, int center_idx, int64_t B, float* d_verts, int64_t verts_pitch,
                     float* d_joints, void* d_workspace, size_t workspace_bytes, void* stream);

/* axis_angle_to_euler_angle (lib/utils/coord_utils.py:83-95) + REBA.__call__ /
 * RULA.__call__ (lib/utils/reba.py:50-81, lib/utils/rula.py:66-98) per frame, i.e. the
 * loop body of lib/core/base.py:225-229 followed by :151 and :168.
 *   d_pose  [B*72] axis-angle, float32 or float64 (pose_dtype)
 *   d_info  [n_tracks]; d_track_of_frame [B] int32 or NULL (every frame uses d_info[0]);
 *           ids outside [0, n_tracks) set flags bit1 (see prk_score_rec)
 *   which   PRK_SCORE_REBA | PRK_SCORE_RULA
 *   d_out   [B] records
 *   d_euler_out [B][n_debug][3] float64 degrees or NULL: Euler angles of the joints
 *           listed in h_debug_joint_ids (the --debug_joints sequences, base.py:144-146) */
int prk_score_pose(const void* d_pose, int pose_dtype, const prk_addinfo* d_info, int32_t n_tracks,
                   const int32_t* d_track_of_frame, int64_t B, uint32_t which,
                   prk_score_rec* d_out, double* d_euler_out, const int32_t* h_debug_joint_ids,
                   int n_debug, void* stream);

/* REBA.__call__(poses, joint_cams, add_info) / RULA.__call__ on Euler degrees
 * (reba.py:50, rula.py:66): d_euler [B*24*3] float64.  joint_cams is never read by the
 * reference (only inside dead string literals, reba.py:262-290) and has no parameter. */
int prk_score_euler(const double* d_euler, const prk_addinfo* d_info, int32_t n_tracks,
                    const int32_t* d_track_of_frame, int64_t B, uint32_t which,
                    prk_score_rec* d_out, void* stream);

/* axis_angle_to_euler_angle alone (coord_utils.py:83-95): n_rot 3-vectors ->
 * float64 degrees [n_rot*3]; d_bad [n_rot] uint8 (1 = the isRotationMatrix assert would
 * fire) or NULL. */
int prk_euler(const void* d_pose, int pose_dtype, int64_t n_rot, double* d_euler,
              uint8_t* d_bad, void* stream);

/* rot_to_angle (coord_utils.py:24-30, called per frame at base.py:222-231): n_rot row-major 3x3
 * rotation matrices (float32 or float64, e.g. SPIN's pred_rotmat [B][24][3][3]) -> n_rot rotation
 * vectors of the same type, cv2.Rodrigues semantics (the matrix is first replaced by its nearest
 * orthogonal matrix).  d_bad [n_rot] uint8 (1 = singular or non-finite matrix, rvec = 0) or NULL. */
int prk_rot_to_angle(const void* d_rotmat, int dtype, int64_t n_rot, void* d_rvec, uint8_t* d_bad,
                     void* stream);

/* The whole per-frame path of lib/core/base.py:225-239,151,168 in one call on device
 * buffers: prk_smpl_forward + prk_score_pose on the same float32 pose (scoring runs beside the
 * body-model kernels on the handle's own stream and is joined before the call's place in `stream`).
 * d_euler_out / h_debug_joint_ids / n_debug as for prk_score_pose (NULL, NULL, 0: no debug sequences).
 * comm_scores / comm_euler (may be NULL): multi-GPU runs, see "multi-GPU exchange" below -- the B score records
 * (and the B x n_debug x 3 Euler angles) are all-gathered into every rank's buffer at frame `frame_offset`;
 * a rank with B = 0 still takes part.
 * All arguments and the workspace size are checked before anything is launched. */
int prk_pipeline(prk_model* model, const float* d_pose, const float* d_betas,
                 const float* d_trans, int center_idx, const prk_addinfo* d_info, int32_t n_tracks,
                 const int32_t* d_track_of_frame, int64_t B, float* d_verts, int64_t verts_pitch, float* d_joints,
                 prk_score_rec* d_scores, double* d_euler_out, const int32_t* h_debug_joint_ids,
                 int n_debug, struct prk_comm* comm_scores, struct prk_comm* comm_euler,
                 int64_t frame_offset, void* d_workspace, size_t workspace_bytes, void* stream);

/* Same, from HOST buffers (what a reference caller holds, base.py:222): copies pose /
 * betas / trans host->device, runs prk_pipeline, copies scores and joints device->host.
 * Vertices stay in device memory (d_verts, may be NULL).  Staging buffers live in the
 * workspace: size it with prk_host_workspace_bytes.  Host buffers should be pinned for
 * the copies to be asynchronous.
 * Completion is ordered on `stream` (synchronise it before reading h_joints / h_scores),
 * but the copies run on the handle's own copy streams so that back-to-back calls overlap:
 * the inputs of call i+1 are uploaded while the kernels of call i run (two input sets in
 * the workspace), and joints / scores are downloaded while the vertex kernel of the same
 * call is still running.  The host input buffers are read from the moment of the call
 * until `stream` reaches it; they must already hold the data when the call is made.
 * The overlap applies to consecutive prk_pipeline_host calls of one handle on the same
 * workspace; that workspace must not be handed to other work in between. */
size_t prk_host_workspace_bytes(const prk_model* model, int64_t B, uint32_t flags);
/* byte offset, inside the workspace of a prk_pipeline_host call of B frames, of the DEVICE copy
 * of the B score records (valid once `stream` has reached the call): what a multi-GPU caller
 * all-gathers without a second host->device copy. */
size_t prk_host_scores_offset(const prk_model* model, int64_t B);
int prk_pipeline_host(prk_model* model, const float* h_pose, const float* h_betas,
                      const float* h_trans, int center_idx, const prk_addinfo* h_info,
                      int32_t n_tracks, const int32_t* h_track_of_frame, int64_t B,
                      float* d_verts, int64_t verts_pitch, float* h_joints, prk_score_rec* h_scores,
                      struct prk_comm* comm_scores, int64_t frame_offset,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* Predictor.post_processing aggregation (lib/core/base.py:260-271) over one scorer's
 * score sequence: d_hist receives a 64-bin histogram of score values clamped to
 * [-16, 47] (bin = score + 16); mean / top-50% / top-10% / max / mode follow from it on
 * the host.  which = PRK_SCORE_REBA or PRK_SCORE_RULA.  d_hist is zeroed by the call. */
int prk_score_histogram(const prk_score_rec* d_scores, int64_t B, uint32_t which,
                        unsigned long long* d_hist, void* stream);

/* ---- multi-GPU exchange: frames are sharded over the GPUs of one box, one process per
 * GPU, and the only exchange is an all-gather of the per-frame score records and, optionally, of the debug
 * Euler sequences (the reference keeps both as per-frame Python lists: lib/core/base.py:144-151,168).
 * The gather runs over peer memory: every rank owns a gather buffer that its peers map through CUDA IPC
 * (or address directly when they live in the same process); prk_allgather_rows stores this rank's rows into
 * EVERY rank's buffer over NVLink and raises one flag per peer; a one-warp kernel on the communicator's own
 * stream waits for the peers' flags, and the caller's stream is joined with it -- no host round trip, no NCCL
 * kernel competing for SMs.  torch.distributed / MPI / files are only needed once, to
 * hand the opaque handles round.
 *
 *   prk_comm_create        allocates the gather buffer (two slots of slot_bytes each) and the flags
 *   prk_comm_handle_bytes  size of one rank's opaque handle
 *   prk_comm_get_handle    this rank's handle (HOST buffer of prk_comm_handle_bytes())
 *   prk_comm_open_peers    h_all_handles = the world's handles in rank order (world * handle bytes)
 *   prk_allgather_rows     d_local: n_local rows of row_bytes that belong at row
 *                          `row_offset` of the gathered array (row_bytes: a multiple of 8); every rank must call it the same number of
 *                          times.  On return (stream order) *d_gathered_out = the device pointer of the
 *                          slot holding all ranks' rows; readers must be queued before this rank issues its next
 *                          exchange on the communicator (two slots alternate).
 *                          A peer that does not arrive within ~20 s makes a later prk_comm_status()
 *                          return PRK_ERR_PEER.
 *   prk_allgather_scores   the same for prk_score_rec rows (row_bytes = 32)
 *   prk_comm_wait          makes `stream` wait for the most recent exchange (needed after prk_pipeline /
 *                          prk_pipeline_host, which run the exchange on a side stream and do not join it)
 *   prk_comm_gathered      device pointer of the slot written by the most recent exchange
 *   prk_comm_status        PRK_OK, or PRK_ERR_PEER once a wait has timed out (reads a host-mapped flag)
 * prk_pipeline / prk_pipeline_host take the communicators directly (comm_scores, comm_euler, frame_offset): the
 * exchange then runs on the handle's scoring stream right behind the scoring kernel, i.e. underneath the vertex
 * kernel.  It waits for the peers, so the call does NOT join it into the caller's stream (a fast rank may run one call
 * ahead of a slow one): call prk_comm_wait(comm, stream) where the gathered rows are read.  d_scores / d_euler_out must
 * not be modified by other work until then. */
typedef struct prk_comm prk_comm;
int prk_comm_create(prk_comm** out, int rank, int world, int device, size_t slot_bytes);
void prk_comm_destroy(prk_comm* comm);
size_t prk_comm_handle_bytes(void);
int prk_comm_get_handle(prk_comm* comm, void* h_handle_out);
int prk_comm_open_peers(prk_comm* comm, const void* h_all_handles);
int prk_comm_wait(prk_comm* comm, void* stream);
void* prk_comm_gathered(prk_comm* comm);
int prk_allgather_rows(prk_comm* comm, const void* d_local, int64_t n_local, int64_t row_offset,
                       int64_t row_bytes, void** d_gathered_out, void* stream);
int prk_allgather_scores(prk_comm* comm, const prk_score_rec* d_local, int64_t n_local,
                         int64_t frame_offset, prk_score_rec** d_gathered_out, void* stream);
int prk_comm_status(prk_comm* comm);

/* ---- verification hooks (used by tests only; not on the product path) ---- */
/* unit range [*u0, *u1) that CTA (pair) k of n_ranges takes in the vertex kernel when a frame-tile switch inside a range
 * is charged switch_cost16 / 16 units (host arithmetic only, no GPU): n_units = frame tiles (pairs) x 216 vertex tiles */
int prk_debug_unit_range(int64_t n_units, int64_t n_ranges, int switch_cost16, int64_t k, int64_t* u0, int64_t* u1);
/* blend-shape stage alone: v_posed [B][prk_vposed_pitch()] float32 into d_vposed (room for
 * B rows), either through the product kernel run with identity skinning transforms
 * (use_simt = 0) or a plain FFMA loop over the same fp16 / e4m3 operands (use_simt = 1). */
int prk_debug_blend(prk_model* model, const float* d_pose, const float* d_betas, int64_t B,
                    float* d_vposed, int use_simt, void* d_workspace, size_t workspace_bytes,
                    void* stream);
int64_t prk_vposed_pitch(void);
/* Per-stage device timing for bench.py's roofline: between begin and end every stage launch
 * of prk_smpl_forward / prk_pipeline is bracketed by CUDA events on the launch stream.
 * ms_out[4] / launches_out[4]: 0 pose chain, 1 blend GEMM, 2 skinning, 3 scoring.
 * Single-threaded use only. */
int prk_profile_begin(void);
/* the same with event pairs only around the stages in stage_mask (bit k = stage k): timing one kernel without putting
 * events between the others (an event between the pose chain and the vertex kernel undoes their programmatic overlap) */
int prk_profile_begin_stages(uint32_t stage_mask);
int prk_profile_end(double* ms_out, int64_t* launches_out);
/* number of kernels launched by this library since load (all threads) */
uint64_t prk_launch_count(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* POSERISK_B200_H */
