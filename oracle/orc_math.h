// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libnrslam_b200.so.
// CPU restatement of the arithmetic NR-SLAM's optimisation hot path runs through g2o/Eigen/Sophus.
// Parity status: UNPINNED by the reference (NR-SLAM ships no tests); pinned only by
// g2o's own known-answer vectors (tests/golden/g2o_linear_solver_kat.json), finite differences and an
// independent NumPy restatement (tests/test_oracle_*.py).
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Camera models. fp32 inside, exactly like modules/calibration (camera_model.h:89-95,131-137: the
// double overloads cast the point to float, evaluate in float, cast the result back).
// ---------------------------------------------------------------------------------------------
struct Camera {
  int model;       // 0 = PinHole (fx fy cx cy), 1 = KannalaBrandt8 (fx fy cx cy k0 k1 k2 k3)
  float p[8];
};

// ref: calibration/pin_hole.cc:27-31 ; calibration/kannala_brandt_8.cc:34-51
inline void project_f(const Camera& c, const float X[3], float uv[2]) {
  if (c.model == 0) {
    uv[0] = c.p[0] * X[0] / X[2] + c.p[2];
    uv[1] = c.p[1] * X[1] / X[2] + c.p[3];
  } else {
    const float x2y2 = X[0] * X[0] + X[1] * X[1];
    const float theta = atan2f(sqrtf(x2y2), X[2]);
    const float psi = atan2f(X[1], X[0]);
    const float t2 = theta * theta, t3 = theta * t2, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
    const float r = theta + c.p[4] * t3 + c.p[5] * t5 + c.p[6] * t7 + c.p[7] * t9;
    uv[0] = c.p[0] * r * cosf(psi) + c.p[2];
    uv[1] = c.p[1] * r * sinf(psi) + c.p[3];
  }
}

// ref: calibration/pin_hole.cc:40-49 ; calibration/kannala_brandt_8.cc:87-116. J is row-major 2x3.
inline void projection_jacobian_f(const Camera& c, const float X[3], float J[6]) {
  if (c.model == 0) {
    J[0] = c.p[0] / X[2];
    J[1] = 0.f;
    J[2] = -c.p[0] * X[0] / (X[2] * X[2]);
    J[3] = 0.f;
    J[4] = c.p[1] / X[2];
    J[5] = -c.p[1] * X[1] / (X[2] * X[2]);
  } else {
    const float fx = c.p[0], fy = c.p[1], k0 = c.p[4], k1 = c.p[5], k2 = c.p[6], k3 = c.p[7];
    float x2 = X[0] * X[0], y2 = X[1] * X[1], z2 = X[2] * X[2];
    float r2 = x2 + y2, r = sqrtf(r2), r3 = r2 * r;
    float theta = atan2f(r, X[2]);
    float t2 = theta * theta, t3 = t2 * theta, t4 = t2 * t2, t5 = t4 * theta;
    float t6 = t2 * t4, t7 = t6 * theta, t8 = t4 * t4, t9 = t8 * theta;
    float f = theta + t3 * k0 + t5 * k1 + t7 * k2 + t9 * k3;
    float fd = 1 + 3 * k0 * t2 + 5 * k1 * t4 + 7 * k2 * t6 + 9 * k3 * t8;
    J[0] = fx * (fd * X[2] * x2 / (r2 * (r2 + z2)) + f * y2 / r3);
    J[1] = fx * (fd * X[2] * X[1] * X[0] / (r2 * (r2 + z2)) - f * X[1] * X[0] / r3);
    J[2] = -fx * fd * X[0] / (r2 + z2);
    J[3] = fy * (fd * X[2] * X[1] * X[0] / (r2 * (r2 + z2)) - f * X[1] * X[0] / r3);
    J[4] = fy * (fd * X[2] * y2 / (r2 * (r2 + z2)) + f * x2 / r3);
    J[5] = -fy * fd * X[1] / (r2 + z2);
  }
}

// double-in / double-out overloads through float (camera_model.h:89-95, 131-137)
inline void project_d(const Camera& c, const double X[3], double uv[2]) {
  float Xf[3] = {(float)X[0], (float)X[1], (float)X[2]}, o[2];
  project_f(c, Xf, o);
  uv[0] = o[0];
  uv[1] = o[1];
}
inline void projection_jacobian_d(const Camera& c, const double X[3], double J[6]) {
  float Xf[3] = {(float)X[0], (float)X[1], (float)X[2]}, o[6];
  projection_jacobian_f(c, Xf, o);
  for (int i = 0; i < 6; i++) J[i] = o[i];
}

// ---------------------------------------------------------------------------------------------
// SE3 with unit quaternion, fp64 — restates g2o::SE3Quat (third_party/g2o/g2o/types/slam3d/se3quat.h).
// Quaternion stored (x, y, z, w) like Eigen's coeffs().
// ---------------------------------------------------------------------------------------------
struct SE3 {
  double q[4];  // x y z w
  double t[3];
};

// Eigen::Quaternion::toRotationMatrix, row-major 3x3.
inline void quat_to_R(const double q[4], double R[9]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen::Quaternion(Matrix3) (trace / largest-diagonal branches).
template <typename T>
inline void R_to_quat(const T R[9], T q[4]) {
  T t = R[0] + R[4] + R[8];
  if (t > T(0)) {
    t = std::sqrt(t + T(1));
    q[3] = T(0.5) * t;
    t = T(0.5) / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + T(1));
    q[i] = T(0.5) * t;
    t = T(0.5) / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}

// se3quat.h:250-255
inline void normalize_rotation(SE3& T) {
  if (T.q[3] < 0)
    for (int i = 0; i < 4; i++) T.q[i] = -T.q[i];
  double n = std::sqrt(T.q[0] * T.q[0] + T.q[1] * T.q[1] + T.q[2] * T.q[2] + T.q[3] * T.q[3]);
  for (int i = 0; i < 4; i++) T.q[i] /= n;
}

// Eigen quaternion * vector:  v + w*2(qv x v) + qv x 2(qv x v)
inline void quat_rotate(const double q[4], const double v[3], double o[3]) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  o[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
  o[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
  o[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

inline void quat_mul(const double a[4], const double b[4], double o[4]) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

// se3quat.h:197  map(): _r * xyz + _t
inline void se3_map(const SE3& T, const double X[3], double o[3]) {
  quat_rotate(T.q, X, o);
  o[0] += T.t[0]; o[1] += T.t[1]; o[2] += T.t[2];
}

// se3quat.h:96-102  operator*
inline SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r;
  double rt[3];
  quat_rotate(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  quat_mul(a.q, b.q, r.q);
  normalize_rotation(r);
  return r;
}

// se3quat.h:199-229  exp([omega, upsilon])
inline SE3 se3_exp(const double u[6]) {
  const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
  const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double O2[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += O[r * 3 + k] * O[k * 3 + c];
      O2[r * 3 + c] = s;
    }
  double a, b, c2, d;
  if (theta < 0.00001) {
    a = 1.0; b = 0.5; c2 = 0.5; d = 1.0 / 6.0;
  } else {
    a = std::sin(theta) / theta;
    b = (1 - std::cos(theta)) / (theta * theta);
    c2 = b;
    d = (theta - std::sin(theta)) / std::pow(theta, 3);
  }
  double R[9], V[9];
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = I + a * O[i] + b * O2[i];
    V[i] = I + c2 * O[i] + d * O2[i];
  }
  SE3 T;
  R_to_quat<double>(R, T.q);
  for (int r = 0; r < 3; r++) T.t[r] = V[r * 3] * up[0] + V[r * 3 + 1] * up[1] + V[r * 3 + 2] * up[2];
  normalize_rotation(T);
  return T;
}

// Sophus::SE3f (unit quaternion xyzw + translation, fp32)  ->  g2o::SE3Quat
// ref: optimization/g2o_optimization.cc:69-71 (cast<double>() then SE3Quat ctor normalises)
inline SE3 se3_from_f7(const float p[7]) {
  SE3 T;
  for (int i = 0; i < 4; i++) T.q[i] = p[i];
  for (int i = 0; i < 3; i++) T.t[i] = p[4 + i];
  normalize_rotation(T);
  return T;
}

// g2o::SE3Quat -> Sophus::SE3f through a 4x4 fp32 matrix (g2o_optimization.cc:144-145).
inline void se3_to_f7(const SE3& T, float p[7]) {
  double R[9];
  quat_to_R(T.q, R);
  float Rf[9], qf[4];
  for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
  R_to_quat<float>(Rf, qf);
  float n = std::sqrt(qf[0] * qf[0] + qf[1] * qf[1] + qf[2] * qf[2] + qf[3] * qf[3]);
  for (int i = 0; i < 4; i++) p[i] = qf[i] / n;
  for (int i = 0; i < 3; i++) p[4 + i] = (float)T.t[i];
}

// g2o RobustKernelHuber::robustify  (third_party/g2o/g2o/core/robust_kernel_impl.cpp:60-74)
inline void huber(double e, double delta, double rho[3]) {
  const double dsqr = delta * delta;
  if (e <= dsqr) {
    rho[0] = e; rho[1] = 1.; rho[2] = 0.;
  } else {
    const double sqrte = std::sqrt(e);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e;
  }
}

}  // namespace orc
