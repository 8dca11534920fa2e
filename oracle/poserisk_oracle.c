/*
 * poserisk_oracle.c -- CPU restatement of the PoseRisk body-model -> risk-score path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * product in poserisk_release_b200/csrc.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * never links, imports or calls anything in oracle/.
 *
 * Pinning: the reference ships no tests or golden vectors.
 * This restatement is pinned against (a) the UNMODIFIED reference Python code
 * executed in the build container (tests/test_oracle_vs_reference.py) and (b)
 * golden vectors generated from that code (tests/golden/make_golden.py ->
 * tests/golden/*.npz, checked on every box by tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Third-party arithmetic that is not in the reference tree:
 *   - cv2.Rodrigues (opencv-python, unpinned in requirements.txt:10; 4.13.0 in
 *     this image): published closed form R = c*I + (1-c)*r r^T + s*[r]x with
 *     theta = |r| in double, identity when theta < DBL_EPSILON, result rounded
 *     to the input's dtype.  Call site: lib/utils/coord_utils.py:86.
 *   - torch CPU fp32 kernels (torch==1.10.2 pinned, 2.11.0 here): plain fp32
 *     arithmetic; accumulation order differs, so SMPL parity is tolerance based.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NV 6890
#define NJ 24
#define NB 10
#define NP 207

/* ---- add_info (example/additional_information.json:1-25) ------------------
 * reba[7]: Legs_bilateral_weight_bearing/walking, Sitting, Load/Force Score,
 *          Arm_supported_leaning_L, Arm_supported_leaning_R, Coupling, Activity_Score
 * rula[9]: Arm_supported_leaning_L, Arm_supported_leaning_R, A_Muscle_use_L,
 *          A_Muscle_use_R, A_Load/Force_L, A_Load/Force_R,
 *          Legs_bilateral_weight_bearing, B_Muscle_use, B_Load/Force           */
typedef struct { int32_t reba[7]; int32_t rula[9]; } orc_addinfo;

/* Per-frame result record (32 bytes).
 * reba_parts: trunk, neck, leg, uaL, uaR, laL, laR, wL, wR   (reba.py:119,137)
 * rula_parts: uaL, uaR, laL, laR, wL, wR, wtL, wtR, neck, trunk, leg (rula.py:139,156)
 * flags bit0: a used joint failed isRotationMatrix (coord_utils.py:70 assert). */
typedef struct {
    int16_t reba_score, rula_score;
    uint8_t reba_parts[9];
    uint8_t rula_parts[11];
    uint8_t flags;
    uint8_t pad[7];
} orc_score_rec;

enum { Torso = 3, L_Knee = 4, R_Knee = 5, Neck = 12, L_Thorax = 13, R_Thorax = 14,
       L_Shoulder = 16, R_Shoulder = 17, L_Elbow = 18, R_Elbow = 19, L_Wrist = 20,
       R_Wrist = 21 };  /* reba.py:9-11, rula.py:9-11 */

int orc_abi_version(void) { return 1; }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU arm of the bench asks for all cores explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static int iclip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ===========================================================================
 * Euler angles: lib/utils/coord_utils.py:62-95
 * ======================================================================== */

/* cv2.Rodrigues(rvec) -> R, computed in double (see header). */
static void cv_rodrigues(const double r_in[3], double R[9]) {
    double x = r_in[0], y = r_in[1], z = r_in[2];
    double theta = sqrt(x * x + y * y + z * z);
    if (theta < DBL_EPSILON) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
    x *= it; y *= it; z *= it;
    /* Matx33d R = c*eye + c1*rrt + s*r_x : products rrt_ij = r_i*r_j are formed first */
    double xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    R[0] = c + c1 * xx;       R[1] = c1 * xy - s * z;   R[2] = c1 * xz + s * y;
    R[3] = c1 * xy + s * z;   R[4] = c + c1 * yy;       R[5] = c1 * yz - s * x;
    R[6] = c1 * xz - s * y;   R[7] = c1 * yz + s * x;   R[8] = c + c1 * zz;
}

/* One joint: axis-angle -> Euler XYZ degrees (coord_utils.py:83-95 body).
 * is_f32: the pose array is float32, so cv2 returns a float32 matrix and the
 * numpy scalar products in `sy` are float32 (coord_utils.py:71).
 * Returns 1 when assert(isRotationMatrix(R)) would fail (coord_utils.py:62-70). */
static int euler_one(const double aa[3], int is_f32, double out_deg[3]) {
    double R[9];
    cv_rodrigues(aa, R);
    double sy;
    if (is_f32) {
        for (int i = 0; i < 9; ++i) R[i] = (double)(float)R[i];
        float a = (float)R[0] * (float)R[0];
        float b = (float)R[3] * (float)R[3];
        float s2 = a + b;
        sy = sqrt((double)s2);
    } else {
        sy = sqrt(R[0] * R[0] + R[3] * R[3]);
    }
    /* isRotationMatrix: || I - R^T R ||_F < 1e-6 */
    double n2 = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double d = 0.0;
            for (int k = 0; k < 3; ++k) d += R[k * 3 + i] * R[k * 3 + j];
            d = (i == j ? 1.0 : 0.0) - d;
            n2 += d * d;
        }
    int bad = !(sqrt(n2) < 1e-6);
    double ex, ey, ez;
    if (!(sy < 1e-6)) {            /* coord_utils.py:72-76 */
        ex = atan2(R[7], R[8]);
        ey = atan2(-R[6], sy);
        ez = atan2(R[3], R[0]);
    } else {                        /* coord_utils.py:77-80 */
        ex = atan2(-R[5], R[4]);
        ey = atan2(-R[6], sy);
        ez = 0.0;
    }
    out_deg[0] = ex * 180.0 / M_PI;  /* coord_utils.py:93 */
    out_deg[1] = ey * 180.0 / M_PI;
    out_deg[2] = ez * 180.0 / M_PI;
    return bad;
}

/* axis_angle_to_euler_angle over n_rot 3-vectors (coord_utils.py:83-95).
 * pose: float32 or float64 [n_rot*3]; euler: float64 [n_rot*3]; bad: uint8[n_rot] or NULL */
void orc_euler(const void* pose, int is_f32, int64_t n_rot, double* euler, uint8_t* bad) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_rot; ++i) {
        double aa[3];
        for (int c = 0; c < 3; ++c)
            aa[c] = is_f32 ? (double)((const float*)pose)[i * 3 + c] : ((const double*)pose)[i * 3 + c];
        int b = euler_one(aa, is_f32, euler + i * 3);
        if (bad) bad[i] = (uint8_t)b;
    }
}

/* ===========================================================================
 * rot_to_angle: lib/utils/coord_utils.py:24-30  (cv2.Rodrigues(3x3) -> rotation vector)
 *
 * The arithmetic lives in opencv-python (unpinned in requirements.txt:10, 4.13.0 in the build
 * container), not in the reference tree.  Published algorithm of cv::Rodrigues for a matrix
 * input (calib3d): convert to double, replace R by the orthogonal factor U*Vt of its SVD,
 * r = (R21-R12, R02-R20, R10-R01), s = |r|/2, c = (trace-1)/2 clamped to [-1,1], theta = acos(c);
 * s >= 1e-5: rvec = r * theta/(2s);  s < 1e-5 and c > 0: rvec = 0;  otherwise (theta ~ pi) the
 * axis comes from the diagonal with signs from R01, R02 (and R12).  The result is converted back
 * to the input's type.  U*Vt is the orthogonal polar factor of R; it is computed here by Newton's
 * iteration X <- (X + X^-T)/2 instead of an SVD.  Pinned against cv2 itself: tests/golden/rotmat.npz
 * (tests/golden/make_golden.py) -- 99.8 % of float32 results are bit-identical, the rest differ by
 * one float32 ulp (2.4e-7 rad).
 * ======================================================================== */
static int inv3_transpose(const double* X, double* Y) {   /* Y = X^-T = cofactor matrix / det; returns 0 if singular */
    const double c00 = X[4] * X[8] - X[5] * X[7], c01 = X[5] * X[6] - X[3] * X[8], c02 = X[3] * X[7] - X[4] * X[6];
    const double det = X[0] * c00 + X[1] * c01 + X[2] * c02;
    if (!(fabs(det) > 1e-300) || !isfinite(det)) return 0;
    const double id = 1.0 / det;
    Y[0] = c00 * id; Y[1] = c01 * id; Y[2] = c02 * id;
    Y[3] = (X[2] * X[7] - X[1] * X[8]) * id; Y[4] = (X[0] * X[8] - X[2] * X[6]) * id; Y[5] = (X[1] * X[6] - X[0] * X[7]) * id;
    Y[6] = (X[1] * X[5] - X[2] * X[4]) * id; Y[7] = (X[2] * X[3] - X[0] * X[5]) * id; Y[8] = (X[0] * X[4] - X[1] * X[3]) * id;
    return 1;
}

/* one matrix (row-major, double) -> rotation vector; returns 1 when the matrix is singular / not finite */
static int rot_to_angle_one(const double* Rin, double* rvec) {
    double R[9], Y[9];
    for (int i = 0; i < 9; ++i) R[i] = Rin[i];
    for (int it = 0; it < 16; ++it) {                      /* orthogonal polar factor = U*Vt */
        if (!inv3_transpose(R, Y)) { rvec[0] = rvec[1] = rvec[2] = 0.0; return 1; }
        double d = 0.0;
        for (int i = 0; i < 9; ++i) { const double n = 0.5 * (R[i] + Y[i]); d = fmax(d, fabs(n - R[i])); R[i] = n; }
        if (d < 1e-16) break;
    }
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { rx = ry = rz = 0.0; }
        else {
            double t = (R[0] + 1.0) * 0.5; rx = sqrt(t > 0.0 ? t : 0.0);
            t = (R[4] + 1.0) * 0.5; ry = sqrt(t > 0.0 ? t : 0.0) * (R[1] < 0 ? -1.0 : 1.0);
            t = (R[8] + 1.0) * 0.5; rz = sqrt(t > 0.0 ? t : 0.0) * (R[2] < 0 ? -1.0 : 1.0);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && ((R[5] > 0) != (ry * rz > 0))) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            rx *= theta; ry *= theta; rz *= theta;
        }
    } else {
        const double vth = theta / (2.0 * s);
        rx *= vth; ry *= vth; rz *= vth;
    }
    rvec[0] = rx; rvec[1] = ry; rvec[2] = rz;
    return 0;
}

/* rot_to_angle over n_rot row-major 3x3 matrices (coord_utils.py:24-30); rvec has the matrices' type */
void orc_rot_to_angle(const void* rotmat, int is_f32, int64_t n_rot, void* rvec, uint8_t* bad) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_rot; ++i) {
        double R[9], r[3];
        for (int k = 0; k < 9; ++k)
            R[k] = is_f32 ? (double)((const float*)rotmat)[i * 9 + k] : ((const double*)rotmat)[i * 9 + k];
        const int b = rot_to_angle_one(R, r);
        for (int k = 0; k < 3; ++k) {
            if (is_f32) ((float*)rvec)[i * 3 + k] = (float)r[k]; else ((double*)rvec)[i * 3 + k] = r[k];
        }
        if (bad) bad[i] = (uint8_t)b;
    }
}

/* ===========================================================================
 * REBA: lib/utils/reba.py
 * ======================================================================== */
static const int8_t REBA_TA[5][3][4] = {  /* reba.py:13-19 */
    {{1,2,3,4},{1,2,3,4},{3,3,5,6}}, {{2,3,4,5},{3,4,5,6},{4,5,6,7}},
    {{2,4,5,6},{4,5,6,7},{5,6,7,8}}, {{3,5,6,7},{5,6,7,8},{6,7,8,9}},
    {{4,6,7,8},{6,7,8,9},{7,8,9,9}}};
static const int8_t REBA_TB[6][2][3] = {  /* reba.py:21-28 */
    {{1,2,2},{1,2,3}}, {{1,2,3},{2,3,4}}, {{3,4,5},{4,5,5}},
    {{4,5,5},{5,6,7}}, {{6,7,8},{7,8,8}}, {{7,8,8},{8,9,9}}};
static const int8_t REBA_TC[12][12] = {   /* reba.py:30-43 */
    {1,1,1,2,3,3,4,5,6,7,7,7}, {1,2,2,3,4,4,5,6,6,7,7,8}, {2,3,3,3,4,5,6,7,7,8,8,8},
    {3,4,4,4,5,6,7,8,8,9,9,9}, {4,4,4,5,6,7,8,8,9,9,9,9}, {6,6,6,7,8,8,9,9,10,10,10,10},
    {7,7,7,8,9,9,9,10,10,11,11,11}, {8,8,8,9,10,10,10,10,10,11,11,11},
    {9,9,9,10,10,10,11,11,11,12,12,12}, {10,10,10,11,11,11,11,12,12,12,12,12},
    {11,11,11,11,12,12,12,12,12,12,12,12}, {12,12,12,12,12,12,12,12,12,12,12,12}};

#define P(j, c) (pose[(j) * 3 + (c)])

static void reba_frame(const double* pose, const int32_t* info, orc_score_rec* out) {
    const int legs = info[0], sitting = info[1], load = info[2], armL = info[3],
              armR = info[4], coupling = info[5], activity = info[6];
    double a, a1, a2, a3, a4, a5, a6;
    int trunk = 0, neck = 0, leg = 0, s;

    /* trunk_bending reba.py:140-148 */
    a = P(Torso, 0);
    if (fabs(a) < 5) s = 1;
    else if ((a > 5 && a < 20) || (a > -20 && a < -5)) s = 2;
    else if ((a > 20 && a < 60) || (a < -20)) s = 3;
    else if (a > 60) s = 4;
    else s = 1;
    trunk += s;
    /* trunk_twist reba.py:158-164 */
    a = P(Torso, 1);
    if (fabs(a) < 10) s = 0; else if (fabs(a) > 10) s = 1; else s = 0;
    trunk += s;
    /* trunk_side_bending reba.py:150-156: every arm returns 0 */
    /* neck_bending reba.py:166-172 */
    a = P(Neck, 0);
    if (a > -5 && a < 20) s = 1; else if (a < 20 || a < -5) s = 2; else s = 1;
    neck += s;
    /* neck_twist reba.py:174-181 */
    a1 = P(Neck, 2); a2 = P(Neck, 1);
    if (fabs(a1) < 10 && fabs(a2) < 10) s = 0;
    else if (fabs(a1) > 10 || fabs(a2) > 10) s = 1;
    else s = 0;
    neck += s;
    /* leg_bending reba.py:183-200 */
    int l1, l2;
    a1 = P(L_Knee, 0);
    if (a1 < 30) l1 = 0; else if (a1 > 30 && a1 < 60) l1 = 1;
    else if (a1 > 60 && sitting > 0) l1 = 2; else l1 = 0;
    a2 = P(R_Knee, 0);
    if (a2 < 30) l2 = 0; else if (a2 > 30 && a2 < 60) l2 = 1;
    else if (a2 > 60 && sitting > 0) l2 = 2; else l2 = 0;
    leg += legs;                       /* reba.py:113 */
    leg += (l1 > l2 ? l1 : l2);
    trunk = iclip(trunk, 1, 5); neck = iclip(neck, 1, 3); leg = iclip(leg, 1, 4);
    int A = REBA_TA[trunk - 1][neck - 1][leg - 1];   /* reba.py:119 */

    int ua[2] = {0, 0}, la[2] = {0, 0}, w[2] = {0, 0};
    int s1, s2;
    /* upper_arm_bending reba.py:202-243 */
    a1 = P(L_Shoulder, 2); a2 = P(L_Shoulder, 1);
    if (a1 > -110 && a1 < -20) {
        if (fabs(a2) < 20) s1 = 1;
        else if (a2 > 20 || (a2 > -45 && a2 < -20)) s1 = 2;
        else if (a2 > -90 && a2 <= -45) s1 = 3;
        else if (a2 < -90) s1 = 4;
        else s1 = 1;
    } else if (a1 > -20) {
        if (fabs(a2) < 20) s1 = 1;
        else if (a2 > 20 || a2 < 70) s1 = 2;
        else if (a2 > 70) s1 = 2;
        else if (a2 > -70 && a2 < -20) s1 = 4;
        else if (a2 < -70) s1 = 4;
        else s1 = 1;
    } else s1 = 1;
    s1 -= armL;
    a3 = P(R_Shoulder, 2); a4 = P(R_Shoulder, 1);
    if (a3 > 20 && a3 < 110) {
        if (fabs(a4) < 20) s2 = 1;
        else if (a4 < -20 || (a4 > 20 && a4 <= 45)) s2 = 2;
        else if (a4 > 45 && a4 <= 90) s2 = 3;
        else if (a4 > 90) s2 = 4;
        else s2 = 1;
    } else if (a1 > -20) {           /* reba.py:232: LEFT angle1/angle2 reused */
        if (fabs(a2) < 20) s2 = 1;
        else if (a2 > 20 || a2 < 70) s2 = 2;
        else if (a2 > 70) s2 = 2;
        else if (a2 > -70 && a2 < -20) s2 = 4;
        else if (a2 < -70) s2 = 4;
        else s2 = 1;
    } else s2 = 1;
    s2 -= armR;
    ua[0] += s1; ua[1] += s2;
    /* shoulder_rise reba.py:245-260 */
    a1 = P(L_Thorax, 2);
    if (fabs(a1) < 10) s1 = 0; else if (fabs(a1) >= 10) s1 = 1; else s1 = 0;
    a2 = P(R_Thorax, 2);
    if (fabs(a2) < 10) s2 = 0; else if (fabs(a2) >= 10) s2 = 1; else s2 = 0;
    ua[0] += s1; ua[1] += s2;
    /* upper_arm_abducted_rotated reba.py:292-335 */
    s1 = 0; s2 = 0;
    a1 = P(L_Shoulder, 2); a2 = P(L_Shoulder, 0); a3 = P(L_Shoulder, 1);
    if (a1 > -110 && a1 < -20) {
        if (a1 < 45 && fabs(a2) < 10) s1 = 0;
        else if (a1 > 45 || fabs(a2) > 10) s1 = 1;
        else s1 = 0;
    } else if (a1 > -20) {
        if (fabs(a3) < 20) s1 = 1;
        else if (a3 > 20 || a3 < 70) s1 = 1;
        else if (a3 > 70) s1 = 0;
        else if (a3 > -70 && a3 < -20) s1 = 1;
        else if (a3 < -70) s1 = 0;
        else s1 = 0;
        if (fabs(a2) > 10) s1 += 1;
    } else s1 = 0;
    a4 = P(R_Shoulder, 2); a5 = P(R_Shoulder, 0); a6 = P(R_Shoulder, 1);
    if (a4 > 20 && a4 < 110) {
        if (a4 > 45 && fabs(a5) < 10) s2 = 0;
        else if (a4 < 45 || fabs(a5) > 10) s2 = 1;
        else s2 = 0;
    } else if (a4 < 20) {
        if (fabs(a6) < 20) s2 = 1;
        else if (a6 > -70 && a6 < -20) s2 = 1;
        else if (a6 < -70) s2 = 0;
        else if (a6 > 20 && a6 < 70) s2 = 1;
        else if (a6 > 70) s2 = 0;
        else s2 = 0;
        if (fabs(a5) > 10) s1 += 1;   /* reba.py:331: increments the LEFT score */
    } else s2 = 0;
    ua[0] += s1; ua[1] += s2;
    /* lower_arm_bending reba.py:337-356 */
    a1 = P(L_Elbow, 1); { double t = P(L_Elbow, 2); if (t > a1) a1 = t; }
    if (a1 > -100 && a1 < -60) s1 = 1;
    else if (a1 < -100 || (a1 > -60 && a1 < 0)) s1 = 2;
    else s1 = 1;
    a2 = P(R_Elbow, 1); { double t = P(R_Elbow, 2); if (t > a2) a2 = t; }
    if (a2 > 60 && a2 < 100) s2 = 1;
    else if (a2 > 100 || (a2 > 0 && a2 < 60)) s2 = 2;
    else s2 = 1;
    la[0] += s1; la[1] += s2;
    /* wrist_bending reba.py:358-373 */
    a1 = P(L_Wrist, 2);
    if (fabs(a1) < 15) s1 = 1; else if (fabs(a1) > 15) s1 = 2; else s1 = 1;
    a2 = P(R_Wrist, 2);
    if (fabs(a2) < 15) s2 = 1; else if (fabs(a2) > 15) s2 = 2; else s2 = 1;
    w[0] += s1; w[1] += s2;
    /* wrist_side_bending_or_twisted reba.py:375-392 */
    a1 = P(L_Wrist, 1); a2 = P(L_Wrist, 0);
    if (fabs(a1) < 10 && fabs(a2) < 10) s1 = 0;
    else if (fabs(a1) > 10 || fabs(a2) > 10) s1 = 1;
    else s1 = 0;
    a3 = P(R_Wrist, 1); a4 = P(R_Wrist, 0);
    if (fabs(a3) < 10 && fabs(a4) < 10) s2 = 0;
    else if (fabs(a3) > 10 || fabs(a4) > 10) s2 = 1;
    else s2 = 0;
    w[0] += s1; w[1] += s2;

    for (int k = 0; k < 2; ++k) {     /* reba.py:131-133 */
        ua[k] = iclip(ua[k], 1, 6); la[k] = iclip(la[k], 1, 2); w[k] = iclip(w[k], 1, 3);
    }
    int BL = REBA_TB[ua[0] - 1][la[0] - 1][w[0] - 1];
    int BR = REBA_TB[ua[1] - 1][la[1] - 1][w[1] - 1];
    int ga = A + load;                                  /* reba.py:59 */
    int gb = (BL > BR ? BL : BR) + coupling;            /* reba.py:63-64 */
    ga = iclip(ga, 1, 12); gb = iclip(gb, 1, 12);       /* reba.py:67-68 */
    out->reba_score = (int16_t)(REBA_TC[ga - 1][gb - 1] + activity);  /* reba.py:69 */
    out->reba_parts[0] = (uint8_t)trunk; out->reba_parts[1] = (uint8_t)neck;
    out->reba_parts[2] = (uint8_t)leg;
    out->reba_parts[3] = (uint8_t)ua[0]; out->reba_parts[4] = (uint8_t)ua[1];
    out->reba_parts[5] = (uint8_t)la[0]; out->reba_parts[6] = (uint8_t)la[1];
    out->reba_parts[7] = (uint8_t)w[0];  out->reba_parts[8] = (uint8_t)w[1];
}

/* ===========================================================================
 * RULA: lib/utils/rula.py
 * ======================================================================== */
static const int8_t RULA_TA[6][3][4][2] = {  /* rula.py:13-39 */
    {{{1,2},{2,2},{2,3},{3,3}}, {{2,2},{2,2},{3,3},{3,3}}, {{2,3},{3,3},{3,3},{4,4}}},
    {{{2,3},{3,3},{3,4},{4,4}}, {{3,3},{3,3},{3,4},{4,4}}, {{3,4},{4,4},{4,4},{5,5}}},
    {{{3,3},{4,4},{4,4},{5,5}}, {{3,4},{4,4},{4,4},{5,5}}, {{4,4},{4,4},{4,5},{5,5}}},
    {{{4,4},{4,4},{4,5},{5,5}}, {{4,4},{4,4},{4,5},{5,5}}, {{4,4},{4,5},{5,5},{6,6}}},
    {{{5,5},{5,5},{5,6},{6,7}}, {{5,6},{6,6},{6,7},{7,7}}, {{6,6},{6,7},{7,7},{7,8}}},
    {{{7,7},{7,7},{7,8},{8,9}}, {{8,8},{8,8},{8,9},{9,9}}, {{9,9},{9,9},{9,9},{9,9}}}};
static const int8_t RULA_TB[6][6][2] = {     /* rula.py:41-48 */
    {{1,3},{2,3},{3,4},{5,5},{6,6},{7,7}}, {{2,3},{2,3},{4,5},{5,5},{6,7},{7,7}},
    {{3,3},{3,4},{4,5},{5,5},{6,7},{7,7}}, {{5,5},{5,6},{6,7},{7,7},{7,7},{8,8}},
    {{7,7},{7,7},{7,8},{8,8},{8,8},{8,8}}, {{8,8},{8,8},{8,8},{8,9},{9,9},{9,9}}};
static const int8_t RULA_TC[7][7] = {        /* rula.py:50-58 */
    {1,2,3,3,4,5,5}, {2,2,3,4,4,5,5}, {3,3,3,4,4,5,6}, {3,3,3,4,5,6,6},
    {4,4,4,5,6,7,7}, {5,5,6,6,7,7,7}, {5,5,6,7,7,7,7}};

static void rula_frame(const double* pose, const int32_t* info, orc_score_rec* out) {
    const int armL = info[0], armR = info[1], musL = info[2], musR = info[3],
              loadL = info[4], loadR = info[5], legs = info[6], bmus = info[7], bload = info[8];
    double a, a1, a2, a3, a4;
    int ua[2] = {0, 0}, la[2] = {0, 0}, w[2] = {0, 0}, wt[2] = {0, 0};
    int s, s1, s2;

    /* upper_arm_bending rula.py:158-199 */
    s1 = 0; s2 = 0;
    a1 = P(L_Shoulder, 2); a2 = P(L_Shoulder, 1);
    if (a1 > -70 && a1 < 110) {
        if (fabs(a2) < 20) s1 = 1;
        else if (a2 > 20 || (a2 > -45 && a2 < -20)) s1 = 2;
        else if (a2 > -90 && a2 <= -45) s1 = 3;
        else if (a2 < -90) s1 = 4;
        else s1 = 1;
    } else if (a1 > -20) {
        if (fabs(a2) < 20) s1 = 1;
        else if (a2 > 20 && a2 < 70) s1 = 2;
        else if (a2 > 70) s1 = 2;
        else if (a2 > -70 && a2 < -20) s1 = 4;
        else if (a2 < -70) s1 = 4;
        else s1 = 1;
    } else s1 = 1;
    s1 -= armL;
    a3 = P(R_Shoulder, 2); a4 = P(R_Shoulder, 1);
    if (a3 > -70 && a3 < 110) {
        if (fabs(a4) < 20) { /* rula.py:183 assigns angle4=1: score2 stays 0 */ }
        else if (a4 < -20 || (a4 > 20 && a4 <= 45)) s2 = 2;
        else if (a4 > 45 && a4 <= 90) s2 = 3;
        else if (a4 > 90) s2 = 4;
        else s2 = 1;
    } else if (a3 < 20) {
        if (fabs(a4) < 20) s2 = 1;
        else if (a4 > -70 && a4 < -20) s2 = 2;
        else if (a4 < -70) s2 = 2;
        else if (a4 > 20 && a4 < 70) s2 = 4;
        else if (a4 > 70) s2 = 4;
        else s2 = 1;
    } else s2 = 1;
    s2 -= armR;
    ua[0] += s1; ua[1] += s2;
    /* shoulder_rise rula.py:201-216 */
    a1 = P(L_Thorax, 2);
    if (fabs(a1) < 10) s1 = 0; else if (fabs(a1) >= 10) s1 = 1; else s1 = 0;
    a2 = P(R_Thorax, 2);
    if (fabs(a2) < 10) s2 = 0; else if (fabs(a2) >= 10) s2 = 1; else s2 = 0;
    ua[0] += s1; ua[1] += s2;
    /* upper_arm_abducted rula.py:249-285 */
    s1 = 0; s2 = 0;
    a1 = P(L_Shoulder, 2); a2 = P(L_Shoulder, 1);
    if (a1 > -110 && a1 < -20) {
        if (a1 < 45) s1 = 0; else if (a1 > 45) s1 = 1; else s1 = 0;
    } else if (a1 > -20) {
        if (fabs(a2) < 20) s1 = 1;
        else if (a2 > 20 && a2 < 70) s1 = 1;
        else if (a2 > 70) s1 = 0;
        else if (a2 > -70 && a2 < -20) s1 = 1;
        else if (a2 < -70) s1 = 0;
        else s1 = 0;
    } else s1 = 0;
    a3 = P(R_Shoulder, 2); a4 = P(R_Shoulder, 1);
    if (a3 > 20 && a3 < 110) {
        if (a3 > 45) s2 = 0; else if (a3 < 45) s2 = 1; else s2 = 0;
    } else if (a3 < 20) {
        if (fabs(a4) < 20) s2 = 1;
        else if (a4 > -70 && a4 < -20) s2 = 1;
        else if (a4 < -70) s2 = 0;
        else if (a4 > 20 && a4 < 70) s2 = 1;
        else if (a4 > 70) s2 = 0;
        else s2 = 0;
    }                                   /* rula.py:276-282: no else */
    ua[0] += s1; ua[1] += s2;
    /* lower_arm_bending rula.py:287-306 */
    a1 = P(L_Elbow, 1); { double t = P(L_Elbow, 2); if (t > a1) a1 = t; }
    if (a1 > -100 && a1 < -60) s1 = 1;
    else if (a1 < -100 || (a1 > -60 && a1 < 0)) s1 = 2;
    else s1 = 1;
    a2 = P(R_Elbow, 1); { double t = P(R_Elbow, 2); if (t > a2) a2 = t; }
    if (a2 > 60 && a2 < 100) s2 = 1;
    else if (a2 > 100 || (a2 > 0 && a2 < 60)) s2 = 2;
    else s2 = 1;
    la[0] += s1; la[1] += s2;
    /* bent_from_midline_or_out_to_side rula.py:308-323 */
    a1 = P(L_Thorax, 0);
    if (a1 < 10 || (a1 > -45 && a1 < -10)) s1 = 0;
    else if (a1 > 10 || a1 < -45) s1 = 1;
    else s1 = 0;
    a2 = P(R_Thorax, 0);
    if (a2 > -10 || (a2 > 10 && a2 < 45)) s2 = 0;
    else if (a2 < -10 || a2 > 45) s2 = 1;
    else s2 = 0;
    la[0] += s1; la[1] += s2;
    /* wrist_bending rula.py:325-342 */
    a1 = P(L_Wrist, 2);
    if (fabs(a1) < 1) s1 = 1; else if (fabs(a1) > 1 && fabs(a1) < 15) s1 = 2;
    else if (fabs(a1) > 15) s1 = 3; else s1 = 1;
    a2 = P(R_Wrist, 2);
    if (fabs(a2) < 1) s2 = 1; else if (fabs(a2) > 1 && fabs(a2) < 15) s2 = 2;
    else if (fabs(a2) > 15) s2 = 3; else s2 = 1;
    w[0] += s1; w[1] += s2;
    /* wrist_side_bending rula.py:344-359 */
    a1 = P(L_Wrist, 1);
    if (fabs(a1) < 10) s1 = 0; else if (fabs(a1) > 10) s1 = 1; else s1 = 0;
    a2 = P(R_Wrist, 1);
    if (fabs(a2) < 10) s2 = 0; else if (fabs(a2) > 10) s2 = 1; else s2 = 0;
    w[0] += s1; w[1] += s2;
    /* wrist_twist rula.py:361-376 */
    a1 = P(L_Wrist, 0);
    if (fabs(a1) < 45) s1 = 1; else if (fabs(a1) > 45) s1 = 2; else s1 = 1;
    a2 = P(R_Wrist, 0);
    if (fabs(a2) < 45) s2 = 1; else if (fabs(a2) > 45) s2 = 2; else s2 = 1;
    wt[0] += s1; wt[1] += s2;

    for (int k = 0; k < 2; ++k) {     /* rula.py:132-135 */
        ua[k] = iclip(ua[k], 1, 6); la[k] = iclip(la[k], 1, 3);
        w[k] = iclip(w[k], 1, 4); wt[k] = iclip(wt[k], 1, 2);
    }
    int AL = RULA_TA[ua[0] - 1][la[0] - 1][w[0] - 1][wt[0] - 1];
    int AR = RULA_TA[ua[1] - 1][la[1] - 1][w[1] - 1][wt[1] - 1];

    int neck = 0, trunk = 0, leg = 0;
    /* neck_bending rula.py:404-412 */
    a = P(Neck, 0);
    if (a > -5 && a < 10) s = 1; else if (a > 10 && a < 20) s = 2;
    else if (a > 20) s = 3; else if (a < -5) s = 4; else s = 1;
    neck += s;
    /* neck_side_bending_twisted rula.py:414-422 */
    a1 = P(Neck, 2); a2 = P(Neck, 1);
    if (fabs(a1) < 10 && fabs(a2) < 10) s = 0;
    else if (fabs(a1) > 10 || fabs(a2) > 10) s = 1;
    else s = 0;
    neck += s;
    /* trunk_bending rula.py:378-386 */
    a = P(Torso, 0);
    if (fabs(a) < 5) s = 1; else if (a > 5 && a < 20) s = 2;
    else if (a > 20 && a < 60) s = 3; else if (a > 60) s = 4; else s = 1;
    trunk += s;
    /* trunk_twisted rula.py:396-402 */
    a = P(Torso, 1);
    if (fabs(a) < 10) s = 0; else if (fabs(a) > 10) s = 1; else s = 0;
    trunk += s;
    /* trunk_side_bending rula.py:388-394 */
    a = P(Torso, 2);
    if (fabs(a) < 10) s = 0; else if (fabs(a) > 10) s = 1; else s = 0;
    trunk += s;
    leg += legs;                        /* rula.py:151 */
    neck = iclip(neck, 1, 6); trunk = iclip(trunk, 1, 6); leg = iclip(leg, 1, 2);
    int B = RULA_TB[neck - 1][trunk - 1][leg - 1];

    AL += musL + loadL; AR += musR + loadR;            /* rula.py:75-76 */
    int ga = AL > AR ? AL : AR;                          /* rula.py:77 */
    int gb = B + bmus + bload;                           /* rula.py:81 */
    ga = iclip(ga, 1, 7); gb = iclip(gb, 1, 7);          /* rula.py:84-85 */
    out->rula_score = (int16_t)RULA_TC[ga - 1][gb - 1];  /* rula.py:86 */
    out->rula_parts[0] = (uint8_t)ua[0]; out->rula_parts[1] = (uint8_t)ua[1];
    out->rula_parts[2] = (uint8_t)la[0]; out->rula_parts[3] = (uint8_t)la[1];
    out->rula_parts[4] = (uint8_t)w[0];  out->rula_parts[5] = (uint8_t)w[1];
    out->rula_parts[6] = (uint8_t)wt[0]; out->rula_parts[7] = (uint8_t)wt[1];
    out->rula_parts[8] = (uint8_t)neck;  out->rula_parts[9] = (uint8_t)trunk;
    out->rula_parts[10] = (uint8_t)leg;
}
#undef P

/* REBA.__call__/RULA.__call__ over n frames of Euler degrees (reba.py:50-81, rula.py:66-98).
 * euler: float64 [n][24][3]; info: [n_tracks]; track_of_frame: int32[n] or NULL (-> track 0). */
void orc_score_euler(const double* euler, const orc_addinfo* info, const int32_t* track_of_frame,
                     int64_t n, orc_score_rec* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const orc_addinfo* ai = info + (track_of_frame ? track_of_frame[i] : 0);
        orc_score_rec r;
        memset(&r, 0, sizeof r);
        reba_frame(euler + i * 72, ai->reba, &r);
        rula_frame(euler + i * 72, ai->rula, &r);
        out[i] = r;
    }
}

/* base.py:225-229 + :151,168 — pose axis-angle -> Euler -> REBA + RULA, per frame.
 * euler_out: float64 [n][24][3] or NULL. */
void orc_score_pose(const void* pose, int is_f32, const orc_addinfo* info,
                    const int32_t* track_of_frame, int64_t n, orc_score_rec* out,
                    double* euler_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double e[72];
        int bad = 0;
        for (int j = 0; j < NJ; ++j) {
            double aa[3];
            for (int c = 0; c < 3; ++c)
                aa[c] = is_f32 ? (double)((const float*)pose)[i * 72 + j * 3 + c]
                               : ((const double*)pose)[i * 72 + j * 3 + c];
            bad |= euler_one(aa, is_f32, e + j * 3);
        }
        if (euler_out) memcpy(euler_out + i * 72, e, sizeof e);
        const orc_addinfo* ai = info + (track_of_frame ? track_of_frame[i] : 0);
        orc_score_rec r;
        memset(&r, 0, sizeof r);
        reba_frame(e, ai->reba, &r);
        rula_frame(e, ai->rula, &r);
        r.flags = (uint8_t)(bad ? 1 : 0);
        out[i] = r;
    }
}

/* ===========================================================================
 * SMPL_Layer.forward: lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py:65-158
 * ======================================================================== */

/* batch_rodrigues + quat2mat, fp32 (rodrigues_layer.py:13-52) */
static void smpl_rodrigues(const float* aa, float* R) {
    float x = aa[0] + 1e-8f, y = aa[1] + 1e-8f, z = aa[2] + 1e-8f;  /* :43 */
    float angle = sqrtf(x * x + y * y + z * z);
    float nx = aa[0] / angle, ny = aa[1] / angle, nz = aa[2] / angle;  /* :45 */
    float half = angle * 0.5f;
    float vc = cosf(half), vs = sinf(half);
    float qw = vc, qx = vs * nx, qy = vs * ny, qz = vs * nz;            /* :49 */
    float qn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);            /* :21 */
    qw /= qn; qx /= qn; qy /= qn; qz /= qn;
    float w2 = qw * qw, x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
    float wx = qw * qx, wy = qw * qy, wz = qw * qz;
    float xy = qx * qy, xz = qx * qz, yz = qy * qz;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;    R[2] = 2 * wy + 2 * xz;   /* :32-36 */
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2;  R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;    R[8] = w2 - x2 - y2 + z2;
}

/* G(3x4) = Gp(3x4, affine) * [R | t] */
static void affine_mul(const float* Gp, const float* R, const float* t, float* G) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            G[r * 4 + c] = Gp[r * 4 + 0] * R[0 * 3 + c] + Gp[r * 4 + 1] * R[1 * 3 + c] +
                           Gp[r * 4 + 2] * R[2 * 3 + c];
        G[r * 4 + 3] = Gp[r * 4 + 0] * t[0] + Gp[r * 4 + 1] * t[1] + Gp[r * 4 + 2] * t[2] + Gp[r * 4 + 3];
    }
}

typedef struct {
    int nnz_max;
    int32_t* widx;  /* [NV][nnz_max] joint ids, -1 padded */
    float* wval;    /* [NV][nnz_max] */
    int32_t* jptr;  /* [NJ+1] CSR of J_regressor */
    int32_t* jidx;
    float* jval;
} sparse_model;

static void build_sparse(const float* weights, const float* J_regressor, sparse_model* sm) {
    int mx = 0;
    for (int v = 0; v < NV; ++v) {
        int c = 0;
        for (int j = 0; j < NJ; ++j) c += weights[v * NJ + j] != 0.0f;
        if (c > mx) mx = c;
    }
    if (mx == 0) mx = 1;
    sm->nnz_max = mx;
    sm->widx = (int32_t*)malloc(sizeof(int32_t) * NV * mx);
    sm->wval = (float*)malloc(sizeof(float) * NV * mx);
    for (int v = 0; v < NV; ++v) {
        int c = 0;
        for (int j = 0; j < NJ; ++j)
            if (weights[v * NJ + j] != 0.0f) { sm->widx[v * mx + c] = j; sm->wval[v * mx + c] = weights[v * NJ + j]; ++c; }
        for (; c < mx; ++c) { sm->widx[v * mx + c] = -1; sm->wval[v * mx + c] = 0.0f; }
    }
    int total = 0;
    for (int i = 0; i < NJ * NV; ++i) total += J_regressor[i] != 0.0f;
    sm->jptr = (int32_t*)malloc(sizeof(int32_t) * (NJ + 1));
    sm->jidx = (int32_t*)malloc(sizeof(int32_t) * (total ? total : 1));
    sm->jval = (float*)malloc(sizeof(float) * (total ? total : 1));
    int p = 0;
    for (int j = 0; j < NJ; ++j) {
        sm->jptr[j] = p;
        for (int v = 0; v < NV; ++v)
            if (J_regressor[j * NV + v] != 0.0f) { sm->jidx[p] = v; sm->jval[p] = J_regressor[j * NV + v]; ++p; }
    }
    sm->jptr[NJ] = p;
}
static void free_sparse(sparse_model* sm) {
    free(sm->widx); free(sm->wval); free(sm->jptr); free(sm->jidx); free(sm->jval);
}

/* One frame given v_shaped-independent inputs.  `vposed` scratch [NV*3]. */
static void smpl_frame(const float* v_template, const float* shapedirs, const float* posedirs,
                       const sparse_model* sm, const int32_t* parents, const float* pose,
                       const float* betas /* [10], already resolved */, int add_trans,
                       const float* trans, int center_idx, float* vposed, float* verts,
                       float* joints) {
    float R[NJ][9];
    for (int j = 0; j < NJ; ++j) smpl_rodrigues(pose + j * 3, R[j]);   /* tensutils.py:11-19 */
    float pm[NP];                                                       /* tensutils.py:41-48 */
    for (int j = 1; j < NJ; ++j)
        for (int k = 0; k < 9; ++k) pm[(j - 1) * 9 + k] = R[j][k] - ((k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f);

    /* v_shaped, v_posed (smpl_layer.py:87-99) */
    for (int i = 0; i < NV * 3; ++i) {
        const float* sd = shapedirs + (size_t)i * NB;
        float acc = 0.0f;
        for (int k = 0; k < NB; ++k) acc += sd[k] * betas[k];
        vposed[i] = v_template[i] + acc;   /* v_shaped for now */
    }
    float J[NJ][3];                         /* th_j = J_regressor @ v_shaped (:91,:95) */
    for (int j = 0; j < NJ; ++j) {
        float a0 = 0, a1 = 0, a2 = 0;
        for (int p = sm->jptr[j]; p < sm->jptr[j + 1]; ++p) {
            const float* v = vposed + (size_t)sm->jidx[p] * 3;
            a0 += sm->jval[p] * v[0]; a1 += sm->jval[p] * v[1]; a2 += sm->jval[p] * v[2];
        }
        J[j][0] = a0; J[j][1] = a1; J[j][2] = a2;
    }
    if (verts) {
        for (int i = 0; i < NV * 3; ++i) {
            const float* pd = posedirs + (size_t)i * NP;
            float acc = 0.0f;
#pragma omp simd reduction(+ : acc)
            for (int k = 0; k < NP; ++k) acc += pd[k] * pm[k];
            vposed[i] += acc;
        }
    }
    /* kinematic chain (:103-119) */
    float G[NJ][12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) G[0][r * 4 + c] = R[0][r * 3 + c];
        G[0][r * 4 + 3] = J[0][r];
    }
    for (int i = 1; i < NJ; ++i) {
        int p = parents[i];
        float t[3] = {J[i][0] - J[p][0], J[i][1] - J[p][1], J[i][2] - J[p][2]};
        affine_mul(G[p], R[i], t, G[i]);
    }
    /* rest-pose removal (:122-132): A = G - pack(G @ [j;0]) */
    float A[NJ][12];
    for (int i = 0; i < NJ; ++i)
        for (int r = 0; r < 3; ++r) {
            float tmp = G[i][r * 4 + 0] * J[i][0] + G[i][r * 4 + 1] * J[i][1] + G[i][r * 4 + 2] * J[i][2];
            A[i][r * 4 + 0] = G[i][r * 4 + 0]; A[i][r * 4 + 1] = G[i][r * 4 + 1];
            A[i][r * 4 + 2] = G[i][r * 4 + 2]; A[i][r * 4 + 3] = G[i][r * 4 + 3] - tmp;
        }
    /* joints (:145) and centring / translation (:148-155) */
    float off[3] = {0, 0, 0};
    if (add_trans) { off[0] = trans[0]; off[1] = trans[1]; off[2] = trans[2]; }
    else if (center_idx >= 0) { off[0] = -G[center_idx][3]; off[1] = -G[center_idx][7]; off[2] = -G[center_idx][11]; }
    for (int i = 0; i < NJ; ++i) {
        joints[i * 3 + 0] = G[i][3] + off[0];
        joints[i * 3 + 1] = G[i][7] + off[1];
        joints[i * 3 + 2] = G[i][11] + off[2];
    }
    if (!verts) return;
    /* LBS (:134-144); zero weights contribute exactly 0 and are skipped */
    const int mx = sm->nnz_max;
    for (int v = 0; v < NV; ++v) {
        float T[12];
        for (int e = 0; e < 12; ++e) T[e] = 0.0f;
        for (int k = 0; k < mx; ++k) {
            int j = sm->widx[v * mx + k];
            if (j < 0) break;
            float w = sm->wval[v * mx + k];
            for (int e = 0; e < 12; ++e) T[e] += A[j][e] * w;
        }
        const float* p = vposed + (size_t)v * 3;
        for (int r = 0; r < 3; ++r)
            verts[v * 3 + r] = T[r * 4 + 0] * p[0] + T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2] + T[r * 4 + 3] + off[r];
    }
}

/* Batch forward.  betas/trans may be NULL.  verts may be NULL (joints only).
 * center_idx < 0 means None.  model_betas: th_betas buffer (smpl_layer.py:40). */
void orc_smpl_forward(const float* v_template, const float* shapedirs, const float* posedirs,
                      const float* J_regressor, const float* weights, const int32_t* parents,
                      const float* model_betas, const float* pose, const float* betas,
                      const float* trans, int center_idx, int64_t B, float* verts, float* joints) {
    /* whole-batch tests (smpl_layer.py:87,148): torch.norm(x) == 0 */
    int betas_nz = 0, trans_nz = 0;
    if (betas) for (int64_t i = 0; i < B * NB; ++i) { float s = betas[i] * betas[i]; if (s != 0.0f) { betas_nz = 1; break; } }
    if (trans) for (int64_t i = 0; i < B * 3; ++i) { float s = trans[i] * trans[i]; if (s != 0.0f) { trans_nz = 1; break; } }
    sparse_model sm;
    build_sparse(weights, J_regressor, &sm);
#pragma omp parallel
    {
        float* vposed = (float*)malloc(sizeof(float) * NV * 3);
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < B; ++b) {
            const float* bt = betas_nz ? betas + b * NB : model_betas;
            smpl_frame(v_template, shapedirs, posedirs, &sm, parents, pose + b * 72, bt, trans_nz,
                       trans_nz ? trans + b * 3 : NULL, center_idx, vposed,
                       verts ? verts + (size_t)b * NV * 3 : NULL, joints + b * 72);
        }
        free(vposed);
    }
    free_sparse(&sm);
}
