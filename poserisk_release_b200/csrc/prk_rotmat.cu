// rot_to_angle: SPIN's rotation matrices -> axis-angle, the front end of the reference's per-frame
// path (lib/utils/coord_utils.py:24-30, called at lib/core/base.py:222-231), so that pred_rotmat can
// stay on the device instead of going through `.cpu().numpy()` and a Python loop of cv2.Rodrigues.
//
// cv::Rodrigues for a 3x3 input (opencv-python, unpinned in requirements.txt:10; 4.13.0 pinned by
// tests/golden/rotmat.npz): convert to double, replace R by the orthogonal factor U*Vt of its SVD,
// r = (R21-R12, R02-R20, R10-R01), s = |r|/2, c = clamp((trace-1)/2), theta = acos(c);
//   s >= 1e-5            rvec = r * theta / (2 s)
//   s <  1e-5, c > 0     rvec = 0
//   s <  1e-5, c <= 0    axis from the diagonal, signs from R01, R02 (and R12), length theta
// and convert back to the input type.  U*Vt is the orthogonal polar factor; here it comes from
// Newton's iteration X <- (X + X^-T)/2 (quadratic convergence; SPIN's matrices are orthonormal to
// float32 rounding already).  Thread = matrix, double arithmetic with explicit non-contracted
// operations (results do not depend on the compiler's FMA choices), 36 B in / 12 B out per matrix:
// HBM/latency bound.
#include "prk_internal.h"

namespace prk {

namespace {

__device__ __forceinline__ bool inv3_transpose(const double* X, double* Y) {
    const double c00 = __dsub_rn(__dmul_rn(X[4], X[8]), __dmul_rn(X[5], X[7]));
    const double c01 = __dsub_rn(__dmul_rn(X[5], X[6]), __dmul_rn(X[3], X[8]));
    const double c02 = __dsub_rn(__dmul_rn(X[3], X[7]), __dmul_rn(X[4], X[6]));
    const double det = __dadd_rn(__dadd_rn(__dmul_rn(X[0], c00), __dmul_rn(X[1], c01)), __dmul_rn(X[2], c02));
    if (!(fabs(det) > 1e-300) || !isfinite(det)) return false;
    const double id = __ddiv_rn(1.0, det);
    Y[0] = __dmul_rn(c00, id); Y[1] = __dmul_rn(c01, id); Y[2] = __dmul_rn(c02, id);
    Y[3] = __dmul_rn(__dsub_rn(__dmul_rn(X[2], X[7]), __dmul_rn(X[1], X[8])), id);
    Y[4] = __dmul_rn(__dsub_rn(__dmul_rn(X[0], X[8]), __dmul_rn(X[2], X[6])), id);
    Y[5] = __dmul_rn(__dsub_rn(__dmul_rn(X[1], X[6]), __dmul_rn(X[0], X[7])), id);
    Y[6] = __dmul_rn(__dsub_rn(__dmul_rn(X[1], X[5]), __dmul_rn(X[2], X[4])), id);
    Y[7] = __dmul_rn(__dsub_rn(__dmul_rn(X[2], X[3]), __dmul_rn(X[0], X[5])), id);
    Y[8] = __dmul_rn(__dsub_rn(__dmul_rn(X[0], X[4]), __dmul_rn(X[1], X[3])), id);
    return true;
}

template <typename T>
__global__ void __launch_bounds__(128)
rot_to_angle_kernel(const T* __restrict__ rotmat, int64_t n_rot, T* __restrict__ rvec, uint8_t* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rot) return;
    double R[9], Y[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = (double)rotmat[i * 9 + k];
    bool singular = false;
    for (int it = 0; it < 16; ++it) {
        if (!inv3_transpose(R, Y)) { singular = true; break; }
        double d = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double n = __dmul_rn(0.5, __dadd_rn(R[k], Y[k]));
            d = fmax(d, fabs(__dsub_rn(n, R[k])));
            R[k] = n;
        }
        if (d < 1e-16) break;
    }
    double rx = 0.0, ry = 0.0, rz = 0.0;
    if (!singular) {
        rx = __dsub_rn(R[7], R[5]); ry = __dsub_rn(R[2], R[6]); rz = __dsub_rn(R[3], R[1]);
        const double s = sqrt(__dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)), 0.25));
        double c = __dmul_rn(__dsub_rn(__dadd_rn(__dadd_rn(R[0], R[4]), R[8]), 1.0), 0.5);
        c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
        double theta = acos(c);
        if (s < 1e-5) {
            if (c > 0) { rx = ry = rz = 0.0; }
            else {
                double t = __dmul_rn(__dadd_rn(R[0], 1.0), 0.5); rx = sqrt(t > 0.0 ? t : 0.0);
                t = __dmul_rn(__dadd_rn(R[4], 1.0), 0.5); ry = sqrt(t > 0.0 ? t : 0.0) * (R[1] < 0 ? -1.0 : 1.0);
                t = __dmul_rn(__dadd_rn(R[8], 1.0), 0.5); rz = sqrt(t > 0.0 ? t : 0.0) * (R[2] < 0 ? -1.0 : 1.0);
                if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && ((R[5] > 0) != (__dmul_rn(ry, rz) > 0))) rz = -rz;
                theta = __ddiv_rn(theta, sqrt(__dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz))));
                rx = __dmul_rn(rx, theta); ry = __dmul_rn(ry, theta); rz = __dmul_rn(rz, theta);
            }
        } else {
            const double vth = __ddiv_rn(theta, __dmul_rn(2.0, s));
            rx = __dmul_rn(rx, vth); ry = __dmul_rn(ry, vth); rz = __dmul_rn(rz, vth);
        }
    }
    rvec[i * 3 + 0] = (T)rx; rvec[i * 3 + 1] = (T)ry; rvec[i * 3 + 2] = (T)rz;
    if (bad) bad[i] = singular ? 1 : 0;
}

}  // namespace

cudaError_t launch_rot_to_angle(const void* d_rotmat, int dtype, int64_t n_rot, void* d_rvec, uint8_t* d_bad,
                                cudaStream_t s) {
    if (n_rot == 0) return cudaSuccess;
    const unsigned g = (unsigned)((n_rot + 127) / 128);
    if (dtype == PRK_DTYPE_F32)
        rot_to_angle_kernel<float><<<g, 128, 0, s>>>((const float*)d_rotmat, n_rot, (float*)d_rvec, d_bad);
    else
        rot_to_angle_kernel<double><<<g, 128, 0, s>>>((const double*)d_rotmat, n_rot, (double*)d_rvec, d_bad);
    count_launch();
    return cudaGetLastError();
}

}  // namespace prk
