// K3: joint-angle extraction + REBA/RULA scoring (integer kernel, tables in constant memory).
//
// Reference behaviour reproduced here (paths relative to the reference root):
//   axis_angle_to_euler_angle        lib/utils/coord_utils.py:83-95  (cv2.Rodrigues + atan2)
//   rotationMatrixToEulerAngles      lib/utils/coord_utils.py:69-81
//   REBA.__call__, group_a/b, rules  lib/utils/reba.py:50-392
//   RULA.__call__, group_a/b, rules  lib/utils/rula.py:66-422
// All the ladders keep the reference's quirks: strict
// inequalities, the right arm reading left-arm angles (reba.py:232), `score1+=1` in the
// right-arm branch (reba.py:331), the `angle4=1` assignment (rula.py:183), the missing
// else (rula.py:276-282).  NaN compares false everywhere, as in Python.
//
// One thread scores one frame.  Angles are computed in float64 with explicitly
// non-contracted multiplies/adds in the cv2.Rodrigues closed form, so that the only
// differences to the CPU reference are the last-ulp behaviour of sin/cos/atan2.  (Large float32 batches screen the
// angles in float32 first and re-evaluate only the frames near a ladder threshold in float64: euler_screen below.)
// HBM traffic: 144 B (12 joints x 3 x f32) in, 32 B out per frame.
#include "prk_internal.h"

#include <math.h>
#include <cstdlib>

namespace prk {

__constant__ int8_t c_reba_ta[5][3][4] = {   // reba.py:13-19
    {{1,2,3,4},{1,2,3,4},{3,3,5,6}}, {{2,3,4,5},{3,4,5,6},{4,5,6,7}},
    {{2,4,5,6},{4,5,6,7},{5,6,7,8}}, {{3,5,6,7},{5,6,7,8},{6,7,8,9}},
    {{4,6,7,8},{6,7,8,9},{7,8,9,9}}};
__constant__ int8_t c_reba_tb[6][2][3] = {   // reba.py:21-28
    {{1,2,2},{1,2,3}}, {{1,2,3},{2,3,4}}, {{3,4,5},{4,5,5}},
    {{4,5,5},{5,6,7}}, {{6,7,8},{7,8,8}}, {{7,8,8},{8,9,9}}};
__constant__ int8_t c_reba_tc[12][12] = {    // reba.py:30-43
    {1,1,1,2,3,3,4,5,6,7,7,7}, {1,2,2,3,4,4,5,6,6,7,7,8}, {2,3,3,3,4,5,6,7,7,8,8,8},
    {3,4,4,4,5,6,7,8,8,9,9,9}, {4,4,4,5,6,7,8,8,9,9,9,9}, {6,6,6,7,8,8,9,9,10,10,10,10},
    {7,7,7,8,9,9,9,10,10,11,11,11}, {8,8,8,9,10,10,10,10,10,11,11,11},
    {9,9,9,10,10,10,11,11,11,12,12,12}, {10,10,10,11,11,11,11,12,12,12,12,12},
    {11,11,11,11,12,12,12,12,12,12,12,12}, {12,12,12,12,12,12,12,12,12,12,12,12}};
__constant__ int8_t c_rula_ta[6][3][4][2] = {  // rula.py:13-39
    {{{1,2},{2,2},{2,3},{3,3}}, {{2,2},{2,2},{3,3},{3,3}}, {{2,3},{3,3},{3,3},{4,4}}},
    {{{2,3},{3,3},{3,4},{4,4}}, {{3,3},{3,3},{3,4},{4,4}}, {{3,4},{4,4},{4,4},{5,5}}},
    {{{3,3},{4,4},{4,4},{5,5}}, {{3,4},{4,4},{4,4},{5,5}}, {{4,4},{4,4},{4,5},{5,5}}},
    {{{4,4},{4,4},{4,5},{5,5}}, {{4,4},{4,4},{4,5},{5,5}}, {{4,4},{4,5},{5,5},{6,6}}},
    {{{5,5},{5,5},{5,6},{6,7}}, {{5,6},{6,6},{6,7},{7,7}}, {{6,6},{6,7},{7,7},{7,8}}},
    {{{7,7},{7,7},{7,8},{8,9}}, {{8,8},{8,8},{8,9},{9,9}}, {{9,9},{9,9},{9,9},{9,9}}}};
__constant__ int8_t c_rula_tb[6][6][2] = {     // rula.py:41-48
    {{1,3},{2,3},{3,4},{5,5},{6,6},{7,7}}, {{2,3},{2,3},{4,5},{5,5},{6,7},{7,7}},
    {{3,3},{3,4},{4,5},{5,5},{6,7},{7,7}}, {{5,5},{5,6},{6,7},{7,7},{7,7},{8,8}},
    {{7,7},{7,7},{7,8},{8,8},{8,8},{8,8}}, {{8,8},{8,8},{8,8},{8,9},{9,9},{9,9}}};
__constant__ int8_t c_rula_tc[7][7] = {        // rula.py:50-58
    {1,2,3,3,4,5,5}, {2,2,3,4,4,5,5}, {3,3,3,4,4,5,6}, {3,3,3,4,5,6,6},
    {4,4,4,5,6,7,7}, {5,5,6,6,7,7,7}, {5,5,6,7,7,7,7}};

// Slots of the 12 scored joints (reba.py:9-11 joint_name order)
enum Slot { S_TORSO = 0, S_LKNEE, S_RKNEE, S_NECK, S_LTHORAX, S_RTHORAX, S_LSHOULDER,
            S_RSHOULDER, S_LELBOW, S_RELBOW, S_LWRIST, S_RWRIST, N_SLOTS };
// slot -> joint id: 3,4,5, 12,13,14, 16..21
__host__ __device__ constexpr int slot_joint(int s) { return s < 3 ? 3 + s : (s < 6 ? 9 + s : 10 + s); }

struct Angles { double a[N_SLOTS][3]; };   // Euler degrees [slot][x,y,z]
// components of a scored joint's Euler triple that some REBA / RULA rule reads (the P(slot, c)
// uses below): knees x only, thorax x and z, elbows y and z, everything else all three
__host__ __device__ constexpr int slot_need(int s) {
    return (s == S_LKNEE || s == S_RKNEE) ? 1 : ((s == S_LTHORAX || s == S_RTHORAX) ? 5 : ((s == S_LELBOW || s == S_RELBOW) ? 6 : 7));
}

__device__ __forceinline__ int iclip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cv2.Rodrigues(rvec)[0] followed by rotationMatrixToEulerAngles and *180/pi
// (coord_utils.py:86,69-81,93).  Returns true when R is not finite, i.e. the reference's
// assert(isRotationMatrix(R)) (coord_utils.py:70) fires.
// kNeed (a constant after inlining / unrolling): bit c set = component c (x, y, z) is wanted; the others come back as 0
// and their atan2 is not evaluated
// (the ladders never read 8 of the 36 angles: knees y/z, thorax y, elbows x -- slot_need below).
template <bool kF32>
__device__ __forceinline__ bool euler_from_axis_angle(double x, double y, double z, double& ex,
                                                      double& ey, double& ez, const int kNeed = 7) {
    const double th2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
    const double theta = sqrt(th2);
    double R00, R10, R20, R21, R22, R11, R12;
    if (theta < 2.220446049250313e-16) {         // DBL_EPSILON: identity
        R00 = 1.0; R10 = 0.0; R20 = 0.0; R21 = 0.0; R22 = 1.0; R11 = 1.0; R12 = 0.0;
    } else {
        double s, c;
        sincos(theta, &s, &c);
        const double c1 = __dsub_rn(1.0, c);
        const double it = __ddiv_rn(1.0, theta);
        x = __dmul_rn(x, it); y = __dmul_rn(y, it); z = __dmul_rn(z, it);
        const double xx = __dmul_rn(x, x), xy = __dmul_rn(x, y), xz = __dmul_rn(x, z);
        const double yy = __dmul_rn(y, y), yz = __dmul_rn(y, z), zz = __dmul_rn(z, z);
        R00 = __dadd_rn(c, __dmul_rn(c1, xx));
        R11 = __dadd_rn(c, __dmul_rn(c1, yy));
        R22 = __dadd_rn(c, __dmul_rn(c1, zz));
        R10 = __dadd_rn(__dmul_rn(c1, xy), __dmul_rn(s, z));
        R20 = __dsub_rn(__dmul_rn(c1, xz), __dmul_rn(s, y));
        R21 = __dadd_rn(__dmul_rn(c1, yz), __dmul_rn(s, x));
        R12 = __dsub_rn(__dmul_rn(c1, yz), __dmul_rn(s, x));
    }
    double sy;
    if (kF32) {   // cv2 returns float32 for float32 input; numpy keeps the products float32
        const float f00 = (float)R00, f10 = (float)R10;
        R00 = f00; R10 = f10; R20 = (float)R20; R21 = (float)R21; R22 = (float)R22;
        R11 = (float)R11; R12 = (float)R12;
        const float s2 = __fadd_rn(__fmul_rn(f00, f00), __fmul_rn(f10, f10));
        sy = sqrt((double)s2);
    } else {
        sy = sqrt(__dadd_rn(__dmul_rn(R00, R00), __dmul_rn(R10, R10)));
    }
    ex = ey = ez = 0.0;
    if (!(sy < 1e-6)) {
        if (kNeed & 1) ex = atan2(R21, R22);
        if (kNeed & 2) ey = atan2(-R20, sy);
        if (kNeed & 4) ez = atan2(R10, R00);
    } else {
        if (kNeed & 1) ex = atan2(-R12, R11);
        if (kNeed & 2) ey = atan2(-R20, sy);
    }
    const double kPi = 3.141592653589793;
    if (kNeed & 1) ex = __ddiv_rn(__dmul_rn(ex, 180.0), kPi);
    if (kNeed & 2) ey = __ddiv_rn(__dmul_rn(ey, 180.0), kPi);
    if (kNeed & 4) ez = __ddiv_rn(__dmul_rn(ez, 180.0), kPi);
    return !isfinite(theta);
}

// ---- float32 screening pass of the large-batch kernel -----------------------------------
// The ladders only COMPARE angles with constants: the multiples of 5 degrees and +-1 (reba.py / rula.py; grep of every
// comparison in reba_frame / rula_frame below).  An angle that is further from all of them than the error of a float32
// evaluation decides every comparison exactly like the float64 value the reference computes, so the large-batch kernel
// first evaluates the Euler angles in float32 (about 150 instructions per joint instead of 700) and marks the frame
// DOUBTFUL when any angle the ladders read lies within kFastBand degrees (scaled by 1 / sy for the two angles whose
// atan2 arguments shrink with sy = cos(pitch)) of a multiple of 5 or of +-1, when sy < 0.02, or when the rotation is tiny,
// above 4 rad or not finite.  Doubtful frames (about 1 % of random poses) are re-evaluated in float64 by the whole warp, one joint per
// lane, before the ladders run: the records stay bit-identical to the all-float64 path (tests/test_gpu_round2.py::
// test_fast_screening_gives_the_exact_records, 2M adversarial frames).  Error budget: R entries <= 7e-7 absolute in
// float32 (sincosf 2 ulp, the c1 * x * y products, one rounding each) plus 4e-7 from theta's own rounding below 4 rad, i.e.
// <= 1.6e-6 / sy rad = 9e-5 / sy degrees on an angle, plus 2 ulp of atan2f and the conversion to degrees (4e-5 degrees); the band
// is 1e-3 degrees.
// A joint whose three components are exactly zero gives exactly zero angles on both paths (identity) and is not doubtful.
constexpr float kFastBand = 1e-3f;
__device__ __forceinline__ bool near_threshold(float a, float scale) {
    const float r = fabsf(a - 5.0f * rintf(a * 0.2f));          // distance to the next multiple of 5
    const float r1 = fabsf(fabsf(a) - 1.0f);                   // ... and to +-1
    return !(fminf(r, r1) * scale >= kFastBand);               // NaN -> doubtful
}
// returns true when the frame needs the float64 evaluation
// (kNeed: a constant after unrolling, as in euler_from_axis_angle)
__device__ __forceinline__ bool euler_screen(float x, float y, float z, double& ex, double& ey, double& ez, const int kNeed) {
    ex = ey = ez = 0.0;
    if (x == 0.0f && y == 0.0f && z == 0.0f) return false;      // identity on both paths (theta < DBL_EPSILON branch)
    const float theta = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
    if (!(theta > 1e-3f && theta < 4.0f)) return true;     // (the absolute error of a float32 theta grows with theta)
    float s, c;
    sincosf(theta, &s, &c);
    const float c1 = 1.0f - c, it = 1.0f / theta;
    x *= it; y *= it; z *= it;
    const float R00 = fmaf(c1 * x, x, c), R10 = fmaf(c1 * x, y, s * z), R20 = fmaf(c1 * x, z, -s * y);
    const float R21 = fmaf(c1 * y, z, s * x), R22 = fmaf(c1 * z, z, c);
    const float sy = sqrtf(fmaf(R00, R00, R10 * R10));
    if (!(sy >= 0.02f)) return true;
    const float kDeg = 57.29577951308232f;
    const float sc = fminf(sy, 1.0f);
    bool doubt = false;
    if (kNeed & 1) { const float a = atan2f(R21, R22) * kDeg; doubt |= near_threshold(a, sc); ex = (double)a; }
    if (kNeed & 2) { const float a = atan2f(-R20, sy) * kDeg; doubt |= near_threshold(a, 1.0f); ey = (double)a; }
    if (kNeed & 4) { const float a = atan2f(R10, R00) * kDeg; doubt |= near_threshold(a, sc); ez = (double)a; }
    return doubt;
}

// One out-of-line copy of the float64 evaluation for the large-batch kernel: inlined twelve times (once per scored joint) the
// kernel was 23,500 instructions = 376 KB of code and spent most of its time waiting for instruction fetches
// (ncu: stalled_no_instruction 8.3 per issue, issue slots 23 % busy); with the slot loop rolled and this function shared by
// the debug joints, the doubtful-frame pass and the all-float64 path it is a quarter of that.
template <bool kF32>
__device__ __noinline__ bool euler_exact(double x, double y, double z, double* e, int need) {
    double ex, ey, ez;
    const bool bad = euler_from_axis_angle<kF32>(x, y, z, ex, ey, ez, need);
    e[0] = ex; e[1] = ey; e[2] = ez;
    return bad;
}

#define P(slot, c) (A.a[slot][c])

__device__ __forceinline__ void reba_frame(const Angles& A, const int32_t* __restrict__ info,
                                           prk_score_rec& out) {
    const int legs = info[0], sitting = info[1], load = info[2], armL = info[3],
              armR = info[4], coupling = info[5], activity = info[6];
    double a, a1, a2, a3, a4, a5, a6;
    int trunk = 0, neck = 0, leg = 0, s;
    // trunk_bending reba.py:140-148
    a = P(S_TORSO, 0);
    if (fabs(a) < 5) s = 1;
    else if ((a > 5 && a < 20) || (a > -20 && a < -5)) s = 2;
    else if ((a > 20 && a < 60) || (a < -20)) s = 3;
    else if (a > 60) s = 4;
    else s = 1;
    trunk += s;
    // trunk_twist reba.py:158-164 (trunk_side_bending :150-156 is always 0)
    a = P(S_TORSO, 1);
    if (fabs(a) < 10) s = 0; else if (fabs(a) > 10) s = 1; else s = 0;
    trunk += s;
    // neck_bending reba.py:166-172
    a = P(S_NECK, 0);
    if (a > -5 && a < 20) s = 1; else if (a < 20 || a < -5) s = 2; else s = 1;
    neck += s;
    // neck_twist reba.py:174-181
    a1 = P(S_NECK, 2); a2 = P(S_NECK, 1);
    if (fabs(a1) < 10 && fabs(a2) < 10) s = 0;
    else if (fabs(a1) > 10 || fabs(a2) > 10) s = 1;
    else s = 0;
    neck += s;
    // leg_bending reba.py:183-200
    int l1, l2;
    a1 = P(S_LKNEE, 0);
    if (a1 < 30) l1 = 0; else if (a1 > 30 && a1 < 60) l1 = 1;
    else if (a1 > 60 && sitting > 0) l1 = 2; else l1 = 0;
    a2 = P(S_RKNEE, 0);
    if (a2 < 30) l2 = 0; else if (a2 > 30 && a2 < 60) l2 = 1;
    else if (a2 > 60 && sitting > 0) l2 = 2; else l2 = 0;
    leg += legs;
    leg += (l1 > l2 ? l1 : l2);
    trunk = iclip(trunk, 1, 5); neck = iclip(neck, 1, 3); leg = iclip(leg, 1, 4);
    const int GA = c_reba_ta[trunk - 1][neck - 1][leg - 1];

    int ua0 = 0, ua1 = 0, la0 = 0, la1 = 0, w0 = 0, w1 = 0, s1, s2;
    // upper_arm_bending reba.py:202-243
    a1 = P(S_LSHOULDER, 2); a2 = P(S_LSHOULDER, 1);
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
    a3 = P(S_RSHOULDER, 2); a4 = P(S_RSHOULDER, 1);
    if (a3 > 20 && a3 < 110) {
        if (fabs(a4) < 20) s2 = 1;
        else if (a4 < -20 || (a4 > 20 && a4 <= 45)) s2 = 2;
        else if (a4 > 45 && a4 <= 90) s2 = 3;
        else if (a4 > 90) s2 = 4;
        else s2 = 1;
    } else if (a1 > -20) {            // reba.py:232 tests the LEFT shoulder's angles
        if (fabs(a2) < 20) s2 = 1;
        else if (a2 > 20 || a2 < 70) s2 = 2;
        else if (a2 > 70) s2 = 2;
        else if (a2 > -70 && a2 < -20) s2 = 4;
        else if (a2 < -70) s2 = 4;
        else s2 = 1;
    } else s2 = 1;
    s2 -= armR;
    ua0 += s1; ua1 += s2;
    // shoulder_rise reba.py:245-260
    a1 = P(S_LTHORAX, 2);
    if (fabs(a1) < 10) s1 = 0; else if (fabs(a1) >= 10) s1 = 1; else s1 = 0;
    a2 = P(S_RTHORAX, 2);
    if (fabs(a2) < 10) s2 = 0; else if (fabs(a2) >= 10) s2 = 1; else s2 = 0;
    ua0 += s1; ua1 += s2;
    // upper_arm_abducted_rotated reba.py:292-335
    s1 = 0; s2 = 0;
    a1 = P(S_LSHOULDER, 2); a2 = P(S_LSHOULDER, 0); a3 = P(S_LSHOULDER, 1);
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
    a4 = P(S_RSHOULDER, 2); a5 = P(S_RSHOULDER, 0); a6 = P(S_RSHOULDER, 1);
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
        if (fabs(a5) > 10) s1 += 1;    // reba.py:331 bumps the LEFT score
    } else s2 = 0;
    ua0 += s1; ua1 += s2;
    // lower_arm_bending reba.py:337-356 (python max(a, b): b only if b > a)
    a1 = P(S_LELBOW, 1); { const double t = P(S_LELBOW, 2); if (t > a1) a1 = t; }
    if (a1 > -100 && a1 < -60) s1 = 1;
    else if (a1 < -100 || (a1 > -60 && a1 < 0)) s1 = 2;
    else s1 = 1;
    a2 = P(S_RELBOW, 1); { const double t = P(S_RELBOW, 2); if (t > a2) a2 = t; }
    if (a2 > 60 && a2 < 100) s2 = 1;
    else if (a2 > 100 || (a2 > 0 && a2 < 60)) s2 = 2;
    else s2 = 1;
    la0 += s1; la1 += s2;
    // wrist_bending reba.py:358-373
    a1 = P(S_LWRIST, 2);
    if (fabs(a1) < 15) s1 = 1; else if (fabs(a1) > 15) s1 = 2; else s1 = 1;
    a2 = P(S_RWRIST, 2);
    if (fabs(a2) < 15) s2 = 1; else if (fabs(a2) > 15) s2 = 2; else s2 = 1;
    w0 += s1; w1 += s2;
    // wrist_side_bending_or_twisted reba.py:375-392
    a1 = P(S_LWRIST, 1); a2 = P(S_LWRIST, 0);
    if (fabs(a1) < 10 && fabs(a2) < 10) s1 = 0;
    else if (fabs(a1) > 10 || fabs(a2) > 10) s1 = 1;
    else s1 = 0;
    a3 = P(S_RWRIST, 1); a4 = P(S_RWRIST, 0);
    if (fabs(a3) < 10 && fabs(a4) < 10) s2 = 0;
    else if (fabs(a3) > 10 || fabs(a4) > 10) s2 = 1;
    else s2 = 0;
    w0 += s1; w1 += s2;

    ua0 = iclip(ua0, 1, 6); ua1 = iclip(ua1, 1, 6);     // reba.py:131-133
    la0 = iclip(la0, 1, 2); la1 = iclip(la1, 1, 2);
    w0 = iclip(w0, 1, 3); w1 = iclip(w1, 1, 3);
    const int BL = c_reba_tb[ua0 - 1][la0 - 1][w0 - 1];
    const int BR = c_reba_tb[ua1 - 1][la1 - 1][w1 - 1];
    int ga = GA + load;                                   // reba.py:59
    int gb = (BL > BR ? BL : BR) + coupling;              // reba.py:63-64
    ga = iclip(ga, 1, 12); gb = iclip(gb, 1, 12);
    out.reba_score = (int16_t)(c_reba_tc[ga - 1][gb - 1] + activity);   // reba.py:69
    out.reba_parts[0] = (uint8_t)trunk; out.reba_parts[1] = (uint8_t)neck;
    out.reba_parts[2] = (uint8_t)leg;
    out.reba_parts[3] = (uint8_t)ua0; out.reba_parts[4] = (uint8_t)ua1;
    out.reba_parts[5] = (uint8_t)la0; out.reba_parts[6] = (uint8_t)la1;
    out.reba_parts[7] = (uint8_t)w0;  out.reba_parts[8] = (uint8_t)w1;
}

__device__ __forceinline__ void rula_frame(const Angles& A, const int32_t* __restrict__ info,
                                           prk_score_rec& out) {
    const int armL = info[0], armR = info[1], musL = info[2], musR = info[3], loadL = info[4],
              loadR = info[5], legs = info[6], bmus = info[7], bload = info[8];
    double a, a1, a2, a3, a4;
    int ua0 = 0, ua1 = 0, la0 = 0, la1 = 0, w0 = 0, w1 = 0, wt0 = 0, wt1 = 0, s, s1, s2;
    // upper_arm_bending rula.py:158-199
    s1 = 0; s2 = 0;
    a1 = P(S_LSHOULDER, 2); a2 = P(S_LSHOULDER, 1);
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
    a3 = P(S_RSHOULDER, 2); a4 = P(S_RSHOULDER, 1);
    if (a3 > -70 && a3 < 110) {
        if (fabs(a4) < 20) { /* rula.py:183 `angle4=1`: score2 keeps its initial 0 */ }
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
    ua0 += s1; ua1 += s2;
    // shoulder_rise rula.py:201-216
    a1 = P(S_LTHORAX, 2);
    if (fabs(a1) < 10) s1 = 0; else if (fabs(a1) >= 10) s1 = 1; else s1 = 0;
    a2 = P(S_RTHORAX, 2);
    if (fabs(a2) < 10) s2 = 0; else if (fabs(a2) >= 10) s2 = 1; else s2 = 0;
    ua0 += s1; ua1 += s2;
    // upper_arm_abducted rula.py:249-285
    s1 = 0; s2 = 0;
    a1 = P(S_LSHOULDER, 2); a2 = P(S_LSHOULDER, 1);
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
    a3 = P(S_RSHOULDER, 2); a4 = P(S_RSHOULDER, 1);
    if (a3 > 20 && a3 < 110) {
        if (a3 > 45) s2 = 0; else if (a3 < 45) s2 = 1; else s2 = 0;
    } else if (a3 < 20) {
        if (fabs(a4) < 20) s2 = 1;
        else if (a4 > -70 && a4 < -20) s2 = 1;
        else if (a4 < -70) s2 = 0;
        else if (a4 > 20 && a4 < 70) s2 = 1;
        else if (a4 > 70) s2 = 0;
        else s2 = 0;
    }                                   // rula.py:276-282 has no else
    ua0 += s1; ua1 += s2;
    // lower_arm_bending rula.py:287-306
    a1 = P(S_LELBOW, 1); { const double t = P(S_LELBOW, 2); if (t > a1) a1 = t; }
    if (a1 > -100 && a1 < -60) s1 = 1;
    else if (a1 < -100 || (a1 > -60 && a1 < 0)) s1 = 2;
    else s1 = 1;
    a2 = P(S_RELBOW, 1); { const double t = P(S_RELBOW, 2); if (t > a2) a2 = t; }
    if (a2 > 60 && a2 < 100) s2 = 1;
    else if (a2 > 100 || (a2 > 0 && a2 < 60)) s2 = 2;
    else s2 = 1;
    la0 += s1; la1 += s2;
    // bent_from_midline_or_out_to_side rula.py:308-323
    a1 = P(S_LTHORAX, 0);
    if (a1 < 10 || (a1 > -45 && a1 < -10)) s1 = 0;
    else if (a1 > 10 || a1 < -45) s1 = 1;
    else s1 = 0;
    a2 = P(S_RTHORAX, 0);
    if (a2 > -10 || (a2 > 10 && a2 < 45)) s2 = 0;
    else if (a2 < -10 || a2 > 45) s2 = 1;
    else s2 = 0;
    la0 += s1; la1 += s2;
    // wrist_bending rula.py:325-342
    a1 = P(S_LWRIST, 2);
    if (fabs(a1) < 1) s1 = 1; else if (fabs(a1) > 1 && fabs(a1) < 15) s1 = 2;
    else if (fabs(a1) > 15) s1 = 3; else s1 = 1;
    a2 = P(S_RWRIST, 2);
    if (fabs(a2) < 1) s2 = 1; else if (fabs(a2) > 1 && fabs(a2) < 15) s2 = 2;
    else if (fabs(a2) > 15) s2 = 3; else s2 = 1;
    w0 += s1; w1 += s2;
    // wrist_side_bending rula.py:344-359
    a1 = P(S_LWRIST, 1);
    if (fabs(a1) < 10) s1 = 0; else if (fabs(a1) > 10) s1 = 1; else s1 = 0;
    a2 = P(S_RWRIST, 1);
    if (fabs(a2) < 10) s2 = 0; else if (fabs(a2) > 10) s2 = 1; else s2 = 0;
    w0 += s1; w1 += s2;
    // wrist_twist rula.py:361-376
    a1 = P(S_LWRIST, 0);
    if (fabs(a1) < 45) s1 = 1; else if (fabs(a1) > 45) s1 = 2; else s1 = 1;
    a2 = P(S_RWRIST, 0);
    if (fabs(a2) < 45) s2 = 1; else if (fabs(a2) > 45) s2 = 2; else s2 = 1;
    wt0 += s1; wt1 += s2;

    ua0 = iclip(ua0, 1, 6); ua1 = iclip(ua1, 1, 6);     // rula.py:132-135
    la0 = iclip(la0, 1, 3); la1 = iclip(la1, 1, 3);
    w0 = iclip(w0, 1, 4); w1 = iclip(w1, 1, 4);
    wt0 = iclip(wt0, 1, 2); wt1 = iclip(wt1, 1, 2);
    int AL = c_rula_ta[ua0 - 1][la0 - 1][w0 - 1][wt0 - 1];
    int AR = c_rula_ta[ua1 - 1][la1 - 1][w1 - 1][wt1 - 1];

    int neck = 0, trunk = 0, leg = 0;
    // neck_bending rula.py:404-412
    a = P(S_NECK, 0);
    if (a > -5 && a < 10) s = 1; else if (a > 10 && a < 20) s = 2;
    else if (a > 20) s = 3; else if (a < -5) s = 4; else s = 1;
    neck += s;
    // neck_side_bending_twisted rula.py:414-422
    a1 = P(S_NECK, 2); a2 = P(S_NECK, 1);
    if (fabs(a1) < 10 && fabs(a2) < 10) s = 0;
    else if (fabs(a1) > 10 || fabs(a2) > 10) s = 1;
    else s = 0;
    neck += s;
    // trunk_bending rula.py:378-386
    a = P(S_TORSO, 0);
    if (fabs(a) < 5) s = 1; else if (a > 5 && a < 20) s = 2;
    else if (a > 20 && a < 60) s = 3; else if (a > 60) s = 4; else s = 1;
    trunk += s;
    // trunk_twisted rula.py:396-402
    a = P(S_TORSO, 1);
    if (fabs(a) < 10) s = 0; else if (fabs(a) > 10) s = 1; else s = 0;
    trunk += s;
    // trunk_side_bending rula.py:388-394
    a = P(S_TORSO, 2);
    if (fabs(a) < 10) s = 0; else if (fabs(a) > 10) s = 1; else s = 0;
    trunk += s;
    leg += legs;                                          // rula.py:151
    neck = iclip(neck, 1, 6); trunk = iclip(trunk, 1, 6); leg = iclip(leg, 1, 2);
    const int GB = c_rula_tb[neck - 1][trunk - 1][leg - 1];

    AL += musL + loadL; AR += musR + loadR;               // rula.py:75-76
    int ga = AL > AR ? AL : AR;
    int gb = GB + bmus + bload;                           // rula.py:81
    ga = iclip(ga, 1, 7); gb = iclip(gb, 1, 7);
    out.rula_score = (int16_t)c_rula_tc[ga - 1][gb - 1];  // rula.py:86
    out.rula_parts[0] = (uint8_t)ua0; out.rula_parts[1] = (uint8_t)ua1;
    out.rula_parts[2] = (uint8_t)la0; out.rula_parts[3] = (uint8_t)la1;
    out.rula_parts[4] = (uint8_t)w0;  out.rula_parts[5] = (uint8_t)w1;
    out.rula_parts[6] = (uint8_t)wt0; out.rula_parts[7] = (uint8_t)wt1;
    out.rula_parts[8] = (uint8_t)neck; out.rula_parts[9] = (uint8_t)trunk;
    out.rula_parts[10] = (uint8_t)leg;
}
#undef P

__device__ __forceinline__ void store_rec(prk_score_rec* dst, const prk_score_rec& r) {
    const uint4* src = reinterpret_cast<const uint4*>(&r);
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = src[0];
    d[1] = src[1];
}

// additional information of frame i's track; an id outside [0, n_tracks) selects track 0 and sets flags bit 1
// (the reference would raise an IndexError there) instead of reading out of bounds
__device__ __forceinline__ const prk_addinfo* track_info(const prk_addinfo* __restrict__ info, int32_t n_tracks,
                                                         const int32_t* __restrict__ track, int64_t i, uint8_t& flags) {
    int32_t t = track ? track[i] : 0;
    if ((uint32_t)t >= (uint32_t)n_tracks) { t = 0; flags |= 2; }
    return info + t;
}

__device__ __forceinline__ void score_frame(const Angles& A, const prk_addinfo* ai, uint32_t which, uint8_t flags,
                                            prk_score_rec& r) {
    uint4* z = reinterpret_cast<uint4*>(&r);
    z[0] = make_uint4(0, 0, 0, 0); z[1] = make_uint4(0, 0, 0, 0);
    if (which & PRK_SCORE_REBA) reba_frame(A, ai->reba, r);
    if (which & PRK_SCORE_RULA) rula_frame(A, ai->rula, r);
    r.flags = flags;
}

__device__ __forceinline__ void score_and_store(const Angles& A, const prk_addinfo* __restrict__ info, int32_t n_tracks,
                                                const int32_t* __restrict__ track, int64_t i,
                                                uint32_t which, uint8_t flags, prk_score_rec* out) {
    const prk_addinfo* ai = track_info(info, n_tracks, track, i, flags);
    alignas(16) prk_score_rec r;
    score_frame(A, ai, which, flags, r);
    store_rec(out + i, r);
}

// pose (axis-angle) -> Euler -> scores, one thread per frame (large batches, and every call with debug joints).
// debug: bit j of `mask` set = also emit joint j's Euler angles to euler_out[i][slot[j]][3].
//
// Memory path: a frame's pose row is 288 B (576 B for float64), so thread-private row reads would touch 32
// different lines per warp-wide load.  Each warp instead copies the 32 consecutive rows of its frames into a
// shared-memory tile with fully coalesced loads (odd pitch: conflict-free per-thread row reads), the debug Euler
// sequences leave through a second tile the same way, and the 32-byte records are exchanged between lane pairs so
// that every store instruction writes 512 contiguous bytes.
constexpr int kScoreWarps = 2;
constexpr int kPosePitch = 73;            // elements per staged pose row (72 + 1)
constexpr int kDebugStageJoints = 6;      // debug joint lists up to this length are staged (longer lists: direct stores)
constexpr int kDebugPitch = kDebugStageJoints * 3 + 1;

#ifndef PRK_SCORE_MINBLOCKS
#define PRK_SCORE_MINBLOCKS 1
#endif
template <typename T, bool kFast, bool kDebug>
__global__ void __launch_bounds__(kScoreWarps * 32, PRK_SCORE_MINBLOCKS)
score_pose_kernel(const T* __restrict__ pose, const prk_addinfo* __restrict__ info, int32_t n_tracks,
                  const int32_t* __restrict__ track, int64_t B, uint32_t which,
                  prk_score_rec* __restrict__ out, double* __restrict__ euler_out,
                  const DebugSlots dbg, int n_debug) {
    // (the debug Euler rows are staged in the pose tile once the poses are dead: a tile of their own left 7 instead of 11 blocks per SM)
    __shared__ __align__(16) T s_pose[kScoreWarps][32 * kPosePitch];
    static_assert(32 * kPosePitch * sizeof(T) >= 32 * kDebugPitch * sizeof(double), "debug rows do not fit in the pose tile");
    __shared__ double s_fix[kFast ? kScoreWarps : 1][N_SLOTS][3];          // float64 angles of the doubtful frame in hand
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t f0 = ((int64_t)blockIdx.x * kScoreWarps + warp) * 32;     // first frame of this warp
    if (f0 >= B) return;
    const int nf = (int)((B - f0) < 32 ? (B - f0) : 32);                   // frames of this warp (warp-uniform)
    T* tile = s_pose[warp];
    {   // coalesced copy of nf consecutive pose rows: flat element q -> row q / 72, column q % 72
        const T* src = pose + f0 * 72;
        const int n = nf * 72;
        int r = 0, c = lane;
        for (int q = lane; q < n; q += 32) {
            tile[r * kPosePitch + c] = src[q];
            c += 32;
            if (c >= 72) { c -= 72; ++r; }
        }
    }
    __syncwarp();
    const bool live = lane < nf;
    const int64_t i = f0 + lane;
    const T* p = tile + lane * kPosePitch;
    const uint32_t debug_mask = kDebug ? dbg.mask : 0u;
    const bool stage_dbg = n_debug <= kDebugStageJoints;
    double* dtile = reinterpret_cast<double*>(tile);
    double dbg_e[kDebug ? kDebugStageJoints * 3 : 1];          // this frame's staged debug rows (local memory: dynamic slot)
    auto emit = [&](int j, double ex, double ey, double ez) {
        const int slot = dbg.slot[j];
        if (stage_dbg) { dbg_e[slot * 3 + 0] = ex; dbg_e[slot * 3 + 1] = ey; dbg_e[slot * 3 + 2] = ez; }
        else if (live) { double* e = euler_out + (i * n_debug + slot) * 3; e[0] = ex; e[1] = ey; e[2] = ez; }
    };
    Angles A;
    bool bad = false, doubt = false;
#pragma unroll 1
    for (int s = 0; s < N_SLOTS; ++s) {     // rolled: one copy of the Euler code (see euler_exact); A lives in local memory
        const int j = slot_joint(s);
        double e[3];
        if (debug_mask & (1u << j)) {                         // warp-uniform: a debug joint needs the whole triple, in float64
            bad |= euler_exact<sizeof(T) == 4>((double)p[j * 3 + 0], (double)p[j * 3 + 1], (double)p[j * 3 + 2], e, 7);
            emit(j, e[0], e[1], e[2]);
        } else if (kFast) {                                   // float32 screening (see euler_screen)
            doubt |= euler_screen((float)p[j * 3 + 0], (float)p[j * 3 + 1], (float)p[j * 3 + 2], e[0], e[1], e[2], slot_need(s));
        } else {
            bad |= euler_exact<sizeof(T) == 4>((double)p[j * 3 + 0], (double)p[j * 3 + 1], (double)p[j * 3 + 2], e, slot_need(s));
        }
        A.a[s][0] = e[0]; A.a[s][1] = e[1]; A.a[s][2] = e[2];
    }
    if (kFast) {   // doubtful frames: the warp evaluates the frame's twelve scored joints in float64, one joint per lane
        unsigned todo = __ballot_sync(0xffffffffu, doubt && live);
        while (todo) {
            const int d = __ffs(todo) - 1;
            todo &= todo - 1;
            bool b = false;
            if (lane < N_SLOTS) {
                const int j = slot_joint(lane);
                const T* pd = tile + d * kPosePitch + j * 3;
                double e[3];
                const int need = (debug_mask & (1u << j)) ? 7 : slot_need(lane);   // unread components stay 0 as on the float64 path
                b = euler_exact<sizeof(T) == 4>((double)pd[0], (double)pd[1], (double)pd[2], e, need);
                s_fix[warp][lane][0] = e[0]; s_fix[warp][lane][1] = e[1]; s_fix[warp][lane][2] = e[2];
            }
            const bool any_bad = __ballot_sync(0xffffffffu, b) != 0;
            __syncwarp();
            if (lane == d) {
                bad |= any_bad;
#pragma unroll 1
                for (int s = 0; s < N_SLOTS; ++s) { A.a[s][0] = s_fix[warp][s][0]; A.a[s][1] = s_fix[warp][s][1]; A.a[s][2] = s_fix[warp][s][2]; }
            }
            __syncwarp();
        }
    }
    // debug joints that are not scored
    uint32_t rest = debug_mask & ~0x3F7038u;   // bits of joints 3,4,5,12,13,14,16..21 cleared
    while (rest) {
        const int j = __ffs(rest) - 1;
        rest &= rest - 1;
        double e[3];
        euler_exact<sizeof(T) == 4>((double)p[j * 3 + 0], (double)p[j * 3 + 1], (double)p[j * 3 + 2], e, 7);
        emit(j, e[0], e[1], e[2]);
    }
    alignas(16) prk_score_rec r;
    {
        uint8_t flags = bad ? 1 : 0;
        const prk_addinfo* ai = track_info(info, n_tracks, track, live ? i : f0, flags);
        score_frame(A, ai, which, flags, r);
    }
    {   // records: lane l holds the two 16-byte halves of frame f0 + l; store k writes the 512 contiguous bytes of
        // frames f0 + 16k .. f0 + 16k + 15 (lane l: frame 16k + l/2, half l & 1)
        const uint4* h = reinterpret_cast<const uint4*>(&r);
        uint4* dst = reinterpret_cast<uint4*>(out + f0);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int src_lane = 16 * k + (lane >> 1);
            uint4 a, b;
            a.x = __shfl_sync(0xffffffffu, h[0].x, src_lane); a.y = __shfl_sync(0xffffffffu, h[0].y, src_lane);
            a.z = __shfl_sync(0xffffffffu, h[0].z, src_lane); a.w = __shfl_sync(0xffffffffu, h[0].w, src_lane);
            b.x = __shfl_sync(0xffffffffu, h[1].x, src_lane); b.y = __shfl_sync(0xffffffffu, h[1].y, src_lane);
            b.z = __shfl_sync(0xffffffffu, h[1].z, src_lane); b.w = __shfl_sync(0xffffffffu, h[1].w, src_lane);
            if (src_lane < nf) dst[32 * k + lane] = (lane & 1) ? b : a;
        }
    }
    if (kDebug && n_debug > 0 && stage_dbg) {   // coalesced copy of the staged Euler sequences: nf rows of n_debug * 3 doubles
        __syncwarp();                           // every lane is done with the pose tile (doubtful-frame pass included)
        for (int k = 0; k < n_debug * 3; ++k) dtile[lane * kDebugPitch + k] = dbg_e[k];
        __syncwarp();
        const int w = n_debug * 3, n = nf * w;
        double* dst = euler_out + f0 * w;
        int rr = 0, c = lane;
        while (c >= w) { c -= w; ++rr; }
        for (int q = lane; q < n; q += 32) {
            dst[q] = dtile[rr * kDebugPitch + c];
            c += 32;
            while (c >= w) { c -= w; ++rr; }
        }
    }
}

__global__ void __launch_bounds__(128)
score_euler_kernel(const double* __restrict__ euler, const prk_addinfo* __restrict__ info, int32_t n_tracks,
                   const int32_t* __restrict__ track, int64_t B, uint32_t which,
                   prk_score_rec* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double* e = euler + i * 72;
    Angles A;
#pragma unroll
    for (int s = 0; s < N_SLOTS; ++s) {
        const int j = slot_joint(s);
        A.a[s][0] = e[j * 3 + 0]; A.a[s][1] = e[j * 3 + 1]; A.a[s][2] = e[j * 3 + 2];
    }
    score_and_store(A, info, n_tracks, track, i, which, 0, out);
}

template <typename T>
__global__ void __launch_bounds__(128)
euler_kernel(const T* __restrict__ pose, int64_t n_rot, double* __restrict__ euler,
             uint8_t* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rot) return;
    double ex, ey, ez;
    const bool b = euler_from_axis_angle<sizeof(T) == 4>((double)pose[i * 3 + 0], (double)pose[i * 3 + 1],
                                                         (double)pose[i * 3 + 2], ex, ey, ez);
    euler[i * 3 + 0] = ex; euler[i * 3 + 1] = ey; euler[i * 3 + 2] = ez;
    if (bad) bad[i] = b ? 1 : 0;
}

// 64-bin histogram of one scorer's scores (bin = clamp(score,-16,47)+16); block-local
// shared histogram then one atomic per non-empty bin (base.py:260-271 aggregation input).
__global__ void __launch_bounds__(256)
score_hist_kernel(const prk_score_rec* __restrict__ recs, int64_t B, uint32_t which,
                  unsigned long long* __restrict__ hist) {
    __shared__ unsigned int sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 0;
    __syncthreads();
    const int16_t* base = reinterpret_cast<const int16_t*>(recs);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B;
         i += (int64_t)gridDim.x * blockDim.x) {
        int v = base[i * 16 + ((which & PRK_SCORE_RULA) ? 1 : 0)];
        v = v < -16 ? -16 : (v > 47 ? 47 : v);
        atomicAdd(&sh[v + 16], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 64 && sh[threadIdx.x])
        atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// ---- latency-optimised variant for small batches ----------------------------------------
// 16 lanes per frame (two frames per warp): lanes 0..11 each turn one scored joint into
// Euler angles (the float64 sincos/atan2 chains are the long pole), the angles meet in
// shared memory, then lane 0 runs the REBA ladders and lane 1 the RULA ladders of the frame.
// Same device functions as the thread-per-frame kernel => bit-identical records.
constexpr int kFramesPerBlockLanes = 8;
constexpr int64_t kLanesVariantMaxFrames = 32768;    // above this the thread-per-frame kernel takes over (scripts/score_threshold.py: 65,536 frames 47 vs 29 us, 131,072 frames 90 vs 45 us)
template <typename T>
__global__ void __launch_bounds__(kFramesPerBlockLanes * 16)
score_pose_lanes_kernel(const T* __restrict__ pose, const prk_addinfo* __restrict__ info, int32_t n_tracks,
                        const int32_t* __restrict__ track, int64_t B, uint32_t which,
                        prk_score_rec* __restrict__ out) {
    __shared__ double s_e[kFramesPerBlockLanes][N_SLOTS][3];
    __shared__ __align__(16) prk_score_rec s_rec[kFramesPerBlockLanes];
    __shared__ int s_bad[kFramesPerBlockLanes];      // flags of the frame: bit 0 non-finite rotation, bit 1 track id out of range
    const int fl = threadIdx.x >> 4, slot = threadIdx.x & 15;
    const int64_t i = (int64_t)blockIdx.x * kFramesPerBlockLanes + fl;
    const bool live = i < B;
    if (slot == 0) {
        s_bad[fl] = 0;
        uint4* z = reinterpret_cast<uint4*>(&s_rec[fl]);
        z[0] = make_uint4(0, 0, 0, 0); z[1] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    if (live && slot < N_SLOTS) {
        const int j = slot_joint(slot);
        const T* p = pose + i * 72 + j * 3;
        double ex, ey, ez;
        const bool bad = euler_from_axis_angle<sizeof(T) == 4>((double)p[0], (double)p[1], (double)p[2], ex, ey, ez);
        s_e[fl][slot][0] = ex; s_e[fl][slot][1] = ey; s_e[fl][slot][2] = ez;
        if (bad) atomicOr(&s_bad[fl], 1);
    }
    __syncwarp();
    if (live && slot < 2) {
        Angles A;
#pragma unroll
        for (int s = 0; s < N_SLOTS; ++s) { A.a[s][0] = s_e[fl][s][0]; A.a[s][1] = s_e[fl][s][1]; A.a[s][2] = s_e[fl][s][2]; }
        uint8_t oob = 0;
        const prk_addinfo* ai = track_info(info, n_tracks, track, i, oob);
        if (oob && slot == 0) atomicOr(&s_bad[fl], 2);
        if (slot == 0 && (which & PRK_SCORE_REBA)) reba_frame(A, ai->reba, s_rec[fl]);
        if (slot == 1 && (which & PRK_SCORE_RULA)) rula_frame(A, ai->rula, s_rec[fl]);
    }
    __syncwarp();
    if (live && slot == 0) {
        s_rec[fl].flags = (uint8_t)s_bad[fl];
        store_rec(out + i, s_rec[fl]);
    }
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

cudaError_t launch_score_pose(const void* d_pose, int pose_dtype, const prk_addinfo* d_info, int32_t n_tracks,
                              const int32_t* d_track, int64_t B, uint32_t which,
                              prk_score_rec* d_out, double* d_euler_out, const DebugSlots& dbg, int n_debug,
                              cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    static const int64_t lanes_max = [] { const char* e = getenv("PRK_SCORE_LANES_MAX"); return e ? atoll(e) : kLanesVariantMaxFrames; }();
    if (n_debug == 0 && B <= lanes_max) {   // latency-bound regime: spread each frame over 16 lanes
        const unsigned g = grid_for(B, kFramesPerBlockLanes);
        if (pose_dtype == PRK_DTYPE_F32)
            score_pose_lanes_kernel<float><<<g, kFramesPerBlockLanes * 16, 0, s>>>((const float*)d_pose, d_info, n_tracks, d_track, B, which, d_out);
        else
            score_pose_lanes_kernel<double><<<g, kFramesPerBlockLanes * 16, 0, s>>>((const double*)d_pose, d_info, n_tracks, d_track, B, which, d_out);
        count_launch();
        return cudaGetLastError();
    }
    const unsigned g = grid_for(B, kScoreWarps * 32);
    // float32 poses: float32 screening + float64 re-evaluation of doubtful frames (PRK_SCORE_EXACT=1: float64 throughout)
    static const bool exact = [] { const char* e = getenv("PRK_SCORE_EXACT"); return e && atoi(e) != 0; }();
#define PRK_SCORE_LAUNCH(T, FAST, DEBUG)                                                            \
    score_pose_kernel<T, FAST, DEBUG><<<g, kScoreWarps * 32, 0, s>>>(                               \
        (const T*)d_pose, d_info, n_tracks, d_track, B, which, d_out, d_euler_out, dbg, n_debug)
    const bool debug = n_debug > 0;
    if (pose_dtype == PRK_DTYPE_F32 && !exact) { if (debug) PRK_SCORE_LAUNCH(float, true, true); else PRK_SCORE_LAUNCH(float, true, false); }
    else if (pose_dtype == PRK_DTYPE_F32) { if (debug) PRK_SCORE_LAUNCH(float, false, true); else PRK_SCORE_LAUNCH(float, false, false); }
    else { if (debug) PRK_SCORE_LAUNCH(double, false, true); else PRK_SCORE_LAUNCH(double, false, false); }
#undef PRK_SCORE_LAUNCH
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_score_euler(const double* d_euler, const prk_addinfo* d_info, int32_t n_tracks,
                               const int32_t* d_track, int64_t B, uint32_t which,
                               prk_score_rec* d_out, cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    score_euler_kernel<<<grid_for(B, 128), 128, 0, s>>>(d_euler, d_info, n_tracks, d_track, B, which, d_out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_euler(const void* d_pose, int pose_dtype, int64_t n_rot, double* d_euler,
                         uint8_t* d_bad, cudaStream_t s) {
    if (n_rot == 0) return cudaSuccess;
    if (pose_dtype == PRK_DTYPE_F32)
        euler_kernel<float><<<grid_for(n_rot, 128), 128, 0, s>>>((const float*)d_pose, n_rot, d_euler, d_bad);
    else
        euler_kernel<double><<<grid_for(n_rot, 128), 128, 0, s>>>((const double*)d_pose, n_rot, d_euler, d_bad);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_score_hist(const prk_score_rec* d_scores, int64_t B, uint32_t which,
                              unsigned long long* d_hist, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_hist, 0, 64 * sizeof(unsigned long long), s);
    if (e != cudaSuccess || B == 0) return e;
    unsigned g = grid_for(B, 256 * 8);
    if (g > 148u * 8u) g = 148u * 8u;
    if (g == 0) g = 1;
    score_hist_kernel<<<g, 256, 0, s>>>(d_scores, B, which, d_hist);
    count_launch();
    return cudaGetLastError();
}

}  // namespace prk
