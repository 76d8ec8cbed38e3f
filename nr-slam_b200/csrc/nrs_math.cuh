// nrs_math.cuh — device/host arithmetic shared by the sm_100a kernels of libnrslam_b200.
//
// Camera models are evaluated in fp32 and widened, SE(3) algebra and robust weights are fp64, exactly like
// the reference runs them (reference paths relative to /root/reference):
//   PinHole::Project / ProjectionJacobian            modules/calibration/pin_hole.cc:27-49
//   KannalaBrandt8::Project / ProjectionJacobian     modules/calibration/kannala_brandt_8.cc:34-51,87-116
//   double overloads that round through float        modules/calibration/camera_model.h:89-95,131-137
//   g2o::SE3Quat exp / map / operator* / normalize   third_party/g2o/g2o/types/slam3d/se3quat.h:96-118,197-255
//   RobustKernelHuber::robustify                     third_party/g2o/g2o/core/robust_kernel_impl.cpp:60-74
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define NRS_HD __host__ __device__ __forceinline__

namespace nrs {

struct Cam {
  int model;   // 0 PinHole (fx fy cx cy), 1 KannalaBrandt8 (fx fy cx cy k0..k3)
  float p[8];
};

// fp32 products / sums that the compiler must not contract into FMAs: the reference evaluates the camera
// model in plain fp32 and the parity oracle is built with -ffp-contract=off.
#ifdef __CUDA_ARCH__
#define NRS_FM(a, b) __fmul_rn((a), (b))
#define NRS_FA(a, b) __fadd_rn((a), (b))
#define NRS_FS(a, b) __fsub_rn((a), (b))
#define NRS_FD(a, b) __fdiv_rn((a), (b))
#else
#define NRS_FM(a, b) ((a) * (b))
#define NRS_FA(a, b) ((a) + (b))
#define NRS_FS(a, b) ((a) - (b))
#define NRS_FD(a, b) ((a) / (b))
#endif

// pi(X) in fp32.
NRS_HD void project_f(const Cam& c, float x, float y, float z, float& u, float& v) {
  if (c.model == 0) {
    u = NRS_FA(NRS_FD(NRS_FM(c.p[0], x), z), c.p[2]);
    v = NRS_FA(NRS_FD(NRS_FM(c.p[1], y), z), c.p[3]);
  } else {
    const float r2 = NRS_FA(NRS_FM(x, x), NRS_FM(y, y));
    const float th = atan2f(sqrtf(r2), z);
    const float psi = atan2f(y, x);
    const float t2 = NRS_FM(th, th), t3 = NRS_FM(th, t2), t5 = NRS_FM(t3, t2), t7 = NRS_FM(t5, t2),
                t9 = NRS_FM(t7, t2);
    const float r = NRS_FA(NRS_FA(NRS_FA(NRS_FA(th, NRS_FM(c.p[4], t3)), NRS_FM(c.p[5], t5)), NRS_FM(c.p[6], t7)),
                           NRS_FM(c.p[7], t9));
    u = NRS_FA(NRS_FM(NRS_FM(c.p[0], r), cosf(psi)), c.p[2]);
    v = NRS_FA(NRS_FM(NRS_FM(c.p[1], r), sinf(psi)), c.p[3]);
  }
}

// d pi / d X in fp32, row-major 2x3.
NRS_HD void projection_jacobian_f(const Cam& c, float x, float y, float z, float J[6]) {
  if (c.model == 0) {
    J[0] = NRS_FD(c.p[0], z);
    J[1] = 0.f;
    J[2] = NRS_FD(NRS_FM(-c.p[0], x), NRS_FM(z, z));
    J[3] = 0.f;
    J[4] = NRS_FD(c.p[1], z);
    J[5] = NRS_FD(NRS_FM(-c.p[1], y), NRS_FM(z, z));
  } else {
    const float fx = c.p[0], fy = c.p[1], k0 = c.p[4], k1 = c.p[5], k2 = c.p[6], k3 = c.p[7];
    const float x2 = NRS_FM(x, x), y2 = NRS_FM(y, y), z2 = NRS_FM(z, z);
    const float r2 = NRS_FA(x2, y2), r = sqrtf(r2), r3 = NRS_FM(r2, r);
    const float th = atan2f(r, z);
    const float t2 = NRS_FM(th, th), t3 = NRS_FM(t2, th), t4 = NRS_FM(t2, t2), t5 = NRS_FM(t4, th);
    const float t6 = NRS_FM(t2, t4), t7 = NRS_FM(t6, th), t8 = NRS_FM(t4, t4), t9 = NRS_FM(t8, th);
    const float f = NRS_FA(NRS_FA(NRS_FA(NRS_FA(th, NRS_FM(t3, k0)), NRS_FM(t5, k1)), NRS_FM(t7, k2)), NRS_FM(t9, k3));
    const float fd = NRS_FA(NRS_FA(NRS_FA(NRS_FA(1.f, NRS_FM(NRS_FM(3.f, k0), t2)), NRS_FM(NRS_FM(5.f, k1), t4)),
                                   NRS_FM(NRS_FM(7.f, k2), t6)),
                            NRS_FM(NRS_FM(9.f, k3), t8));
    const float den = NRS_FM(r2, NRS_FA(r2, z2));
    const float fdz = NRS_FM(fd, z);
    const float a_xx = NRS_FD(NRS_FM(fdz, x2), den), a_yy = NRS_FD(NRS_FM(fdz, y2), den);
    const float a_xy = NRS_FD(NRS_FM(NRS_FM(fdz, y), x), den);
    const float b_xy = NRS_FD(NRS_FM(NRS_FM(f, y), x), r3);
    J[0] = NRS_FM(fx, NRS_FA(a_xx, NRS_FD(NRS_FM(f, y2), r3)));
    J[1] = NRS_FM(fx, NRS_FS(a_xy, b_xy));
    J[2] = NRS_FD(NRS_FM(NRS_FM(-fx, fd), x), NRS_FA(r2, z2));
    J[3] = NRS_FM(fy, NRS_FS(a_xy, b_xy));
    J[4] = NRS_FM(fy, NRS_FA(a_yy, NRS_FD(NRS_FM(f, x2), r3)));
    J[5] = NRS_FD(NRS_FM(NRS_FM(-fy, fd), y), NRS_FA(r2, z2));
  }
}

// Unit quaternion stored x y z w (Eigen coeffs order); rotation matrix row-major.
NRS_HD void quat_to_R(const double q[4], double R[9]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

template <typename T>
NRS_HD void R_to_quat(const T R[9], T q[4]) {
  T t = R[0] + R[4] + R[8];
  if (t > T(0)) {
    t = sqrt(t + T(1));
    q[3] = T(0.5) * t;
    t = T(0.5) / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + T(1));
    q[i] = T(0.5) * t;
    t = T(0.5) / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}

// pose = 7 doubles: q (x y z w), t.
NRS_HD void pose_normalize(double* T) {
  if (T[3] < 0)
    for (int i = 0; i < 4; i++) T[i] = -T[i];
  const double n = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2] + T[3] * T[3]);
  for (int i = 0; i < 4; i++) T[i] /= n;
}

// v rotated by q: v + w*2(qv x v) + qv x 2(qv x v)
NRS_HD void quat_rotate(const double q[4], const double v[3], double o[3]) {
  double a0 = q[1] * v[2] - q[2] * v[1], a1 = q[2] * v[0] - q[0] * v[2], a2 = q[0] * v[1] - q[1] * v[0];
  a0 += a0; a1 += a1; a2 += a2;
  o[0] = v[0] + q[3] * a0 + (q[1] * a2 - q[2] * a1);
  o[1] = v[1] + q[3] * a1 + (q[2] * a0 - q[0] * a2);
  o[2] = v[2] + q[3] * a2 + (q[0] * a1 - q[1] * a0);
}

NRS_HD void pose_map(const double* T, const double X[3], double o[3]) {
  quat_rotate(T, X, o);
  o[0] += T[4]; o[1] += T[5]; o[2] += T[6];
}

// T <- exp([omega, upsilon]) * T   (VertexSE3Expmap::oplusImpl, types/sba/vertex_se3_expmap.cpp)
NRS_HD void pose_oplus(double* T, const double u[6]) {
  const double om0 = u[0], om1 = u[1], om2 = u[2];
  const double theta = sqrt(om0 * om0 + om1 * om1 + om2 * om2);
  const double O[9] = {0, -om2, om1, om2, 0, -om0, -om1, om0, 0};
  double O2[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += O[r * 3 + k] * O[k * 3 + c];
      O2[r * 3 + c] = s;
    }
  double a, b, d;
  if (theta < 0.00001) {
    a = 1.0; b = 0.5; d = 1.0 / 6.0;
  } else {
    a = sin(theta) / theta;
    b = (1 - cos(theta)) / (theta * theta);
    d = (theta - sin(theta)) / (theta * theta * theta);
  }
  double R[9], V[9];
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = I + a * O[i] + b * O2[i];
    V[i] = I + b * O[i] + d * O2[i];
  }
  double E[7];
  R_to_quat<double>(R, E);
  for (int r = 0; r < 3; r++) E[4 + r] = V[r * 3] * u[3] + V[r * 3 + 1] * u[4] + V[r * 3 + 2] * u[5];
  pose_normalize(E);
  // E * T
  double rt[3], out[7];
  quat_rotate(E, T + 4, rt);
  out[4] = E[4] + rt[0]; out[5] = E[5] + rt[1]; out[6] = E[6] + rt[2];
  out[3] = E[3] * T[3] - E[0] * T[0] - E[1] * T[1] - E[2] * T[2];
  out[0] = E[3] * T[0] + E[0] * T[3] + E[1] * T[2] - E[2] * T[1];
  out[1] = E[3] * T[1] + E[1] * T[3] + E[2] * T[0] - E[0] * T[2];
  out[2] = E[3] * T[2] + E[2] * T[3] + E[0] * T[1] - E[1] * T[0];
  pose_normalize(out);
  for (int i = 0; i < 7; i++) T[i] = out[i];
}

// Huber: rho(e), rho'(e) for e = chi2; delta <= 0 means "no robust kernel".
NRS_HD void huber(double e, double delta, double& rho, double& drho) {
  if (delta <= 0 || e <= delta * delta) {
    rho = e;
    drho = 1.0;
  } else {
    const double sq = sqrt(e);
    rho = 2 * sq * delta - delta * delta;
    drho = delta / sq;
  }
}

// Sophus::SE3f (7 floats) -> fp64 pose (g2o_optimization.cc:69-71: cast<double>() + SE3Quat ctor normalises)
NRS_HD void pose_from_f7(const float* p, double* T) {
  for (int i = 0; i < 7; i++) T[i] = p[i];
  pose_normalize(T);
}
// fp64 pose -> Sophus::SE3f through a 4x4 fp32 matrix (g2o_optimization.cc:144-145)
NRS_HD void pose_to_f7(const double* T, float* p) {
  double R[9];
  quat_to_R(T, R);
  float Rf[9], qf[4];
  for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
  R_to_quat<float>(Rf, qf);
  const float n = sqrtf(qf[0] * qf[0] + qf[1] * qf[1] + qf[2] * qf[2] + qf[3] * qf[3]);
  for (int i = 0; i < 4; i++) p[i] = qf[i] / n;
  for (int i = 0; i < 3; i++) p[4 + i] = (float)T[4 + i];
}

}  // namespace nrs
