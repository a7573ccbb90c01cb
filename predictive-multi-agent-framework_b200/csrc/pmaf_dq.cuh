// pmaf_dq.cuh — dual-quaternion kinematics of the Franka Panda on the device and a batched feasibility score of
// the predicted end-effector paths (the step downstream of the planner).
//
// What the reference evaluates per control cycle for ONE robot on the CPU — pose and pose Jacobian of the arm
// (CoSTPController::calculateControlPreliminaries, src/costp_controller.cpp:111-126, through dqrobotics'
// DQ_SerialManipulator with the modified-DH table of src/franka_robot.cpp:6-22) and the geometric Jacobian
// (geomJ, src/costp_controller.cpp:465-492) — is evaluated here for every predicted path at once: one thread
// tracks one agent's path with a damped-least-squares step per path point (the J^T (J J^T + lambda I)^-1 form of
// src/costp_controller.cpp:134-135, joint limits of :41-44) and reports tracking error, joint-limit margin and
// manipulability, so that the planner can tell which of its best paths the arm can actually follow without
// copying a single path to the host. dqrobotics is not vendored by the reference and absent here: the algebra
// below follows its published definitions, is checked against oracle/dq_oracle.c (itself checked against 4x4
// homogeneous transforms and finite differences), and PARITY WITH A dqrobotics BUILD IS UNPINNED.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pmaf {

struct dq8 {
  double q[8];  // primary w x y z | dual w x y z
};
struct PathScore {  // mirrors pmaf_path_score (include/pmaf.h)
  double max_pos_err, min_joint_margin, min_manipulability;
  int feasible, first_bad_point;
  double q_final[7];
};

__device__ __forceinline__ void quat_mul(const double *a, const double *b, double *o) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
__device__ __forceinline__ dq8 dq_mul(const dq8 &a, const dq8 &b) {
  dq8 o;
  double d1[4], d2[4];
  quat_mul(a.q, b.q, o.q);
  quat_mul(a.q, b.q + 4, d1);
  quat_mul(a.q + 4, b.q, d2);
#pragma unroll
  for (int i = 0; i < 4; ++i) o.q[4 + i] = d1[i] + d2[i];
  return o;
}
__device__ __forceinline__ dq8 dq_conj(const dq8 &a) {
  dq8 o;
  o.q[0] = a.q[0], o.q[1] = -a.q[1], o.q[2] = -a.q[2], o.q[3] = -a.q[3];
  o.q[4] = a.q[4], o.q[5] = -a.q[5], o.q[6] = -a.q[6], o.q[7] = -a.q[7];
  return o;
}

// Panda, modified DH (src/franka_robot.cpp:7-13)
__device__ __constant__ double kPandaD[7] = {0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.2104};
__device__ __constant__ double kPandaA[7] = {0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088};
__device__ __constant__ double kPandaAlphaHalfPis[7] = {0.0, -1.0, 1.0, 1.0, -1.0, 1.0, 1.0};

// constant part of link i, Rot_x(alpha_i) Trans_x(a_i), and the joint axis seen from frame i-1, w_i = A_i k A_i^*
struct PandaModel {
  dq8 A[7], w[7];
};
__device__ __forceinline__ void panda_model(PandaModel &m) {
  for (int i = 0; i < 7; ++i) {
    const double half = 0.5 * (kPandaAlphaHalfPis[i] * 1.57079632679489661923);
    double s, c;
    sincos(half, &s, &c);
    dq8 r{}, t{}, k{};
    r.q[0] = c, r.q[1] = s;
    t.q[0] = 1.0, t.q[5] = 0.5 * kPandaA[i];
    k.q[3] = 1.0;
    m.A[i] = dq_mul(r, t);
    m.w[i] = dq_mul(dq_mul(m.A[i], k), dq_conj(m.A[i]));
  }
}
// x[i] = base * link_0 ... link_{i-1}; x[7] = the end-effector pose (fkm)
__device__ __forceinline__ void panda_chain(const PandaModel &m, const dq8 &base, const double *q, dq8 *x) {
  x[0] = base;
  for (int i = 0; i < 7; ++i) {
    double s, c;
    sincos(0.5 * q[i], &s, &c);
    dq8 B{};
    B.q[0] = c, B.q[3] = s, B.q[4] = -0.5 * kPandaD[i] * s, B.q[7] = 0.5 * kPandaD[i] * c;
    x[i + 1] = dq_mul(x[i], dq_mul(m.A[i], B));
  }
}
__device__ __forceinline__ void dq_translation(const dq8 &x, double *t) {
  const double pc[4] = {x.q[0], -x.q[1], -x.q[2], -x.q[3]};
  double o[4];
  quat_mul(x.q + 4, pc, o);
  t[0] = 2 * o[1], t[1] = 2 * o[2], t[2] = 2 * o[3];
}
// pose Jacobian column i: (1/2) (x_i w_i x_i^*) x_7
__device__ __forceinline__ dq8 pose_jacobian_col(const PandaModel &m, const dq8 *x, int i) {
  dq8 col = dq_mul(dq_mul(dq_mul(x[i], m.w[i]), dq_conj(x[i])), x[7]);
#pragma unroll
  for (int r = 0; r < 8; ++r) col.q[r] *= 0.5;
  return col;
}
// geomJ (src/costp_controller.cpp:465-492) column from a pose Jacobian column: rot = 2 (dP P^*), tra = 2 (dD P^* + D dP^*)
__device__ __forceinline__ void geom_jacobian_col(const dq8 &x, const dq8 &col, double *rot, double *tra) {
  const double pc[4] = {x.q[0], -x.q[1], -x.q[2], -x.q[3]};
  const double dpc[4] = {col.q[0], -col.q[1], -col.q[2], -col.q[3]};
  double a[4], b[4], c[4];
  quat_mul(col.q, pc, a);       // haminus4(P^*) J_P
  quat_mul(col.q + 4, pc, b);   // haminus4(P^*) J_D
  quat_mul(x.q + 4, dpc, c);    // hamiplus4(D) C4 J_P
#pragma unroll
  for (int r = 0; r < 3; ++r) rot[r] = 2 * a[r + 1], tra[r] = 2 * (c[r + 1] + b[r + 1]);
}
__device__ __forceinline__ double det3(const double A[3][3]) {
  return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
         A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}

struct ScoreArgs {
  const double *paths;   // [agents][max_steps][3] predicted paths (device-resident, PlannerDev::paths)
  const int *n_path;     // [agents]
  int max_steps;
  const int *agent_index;  // [n] local agent of every scored path, or null: paths 0..n-1
  int n;
  double base[8], q_start[7], q_lo[7], q_hi[7];
  double damping, tol_pos;
  PathScore *out;        // [n]
};

// one thread per path: the path is a serial recurrence in q, the paths are independent
__global__ void __launch_bounds__(128) dq_score_kernel(const ScoreArgs S) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S.n) return;
  const int a = S.agent_index ? S.agent_index[j] : j;
  PathScore sc;
  sc.max_pos_err = 0.0, sc.min_joint_margin = (double)INFINITY, sc.min_manipulability = (double)INFINITY;
  sc.feasible = 1, sc.first_bad_point = -1;
  double q[7];
  for (int c = 0; c < 7; ++c) q[c] = S.q_start[c];
  if (a >= 0) {
    PandaModel m;
    panda_model(m);
    dq8 base;
    for (int r = 0; r < 8; ++r) base.q[r] = S.base[r];
    const double *row = S.paths + (size_t)a * S.max_steps * 3;
    const int n = S.n_path[a];
    dq8 x[8];
    for (int k = 0; k < n; ++k) {
      const double target[3] = {row[3 * k], row[3 * k + 1], row[3 * k + 2]};
      panda_chain(m, base, q, x);
      double t[3], Jt[3][7];
      dq_translation(x[7], t);
      for (int c = 0; c < 7; ++c) {
        double rot[3], tra[3];
        geom_jacobian_col(x[7], pose_jacobian_col(m, x, c), rot, tra);
        Jt[0][c] = tra[0], Jt[1][c] = tra[1], Jt[2][c] = tra[2];
      }
      double M[3][3], A[3][3];
      for (int i = 0; i < 3; ++i)
        for (int l = 0; l < 3; ++l) {
          double s = 0;
          for (int c = 0; c < 7; ++c) s += Jt[i][c] * Jt[l][c];
          M[i][l] = s, A[i][l] = s + (i == l ? S.damping : 0.0);
        }
      const double manip = sqrt(fmax(det3(M), 0.0));
      const double e[3] = {target[0] - t[0], target[1] - t[1], target[2] - t[2]};
      const double dA = det3(A);
      double y[3];
      for (int c = 0; c < 3; ++c) {  // Cramer's rule, as the oracle
        double B[3][3];
        for (int r = 0; r < 3; ++r)
          for (int l = 0; l < 3; ++l) B[r][l] = l == c ? e[r] : A[r][l];
        y[c] = det3(B) / dA;
      }
      for (int c = 0; c < 7; ++c) q[c] += Jt[0][c] * y[0] + Jt[1][c] * y[1] + Jt[2][c] * y[2];
      panda_chain(m, base, q, x);
      dq_translation(x[7], t);
      double r2 = 0;
      for (int i = 0; i < 3; ++i) r2 += (target[i] - t[i]) * (target[i] - t[i]);
      const double err = sqrt(r2);
      double margin = (double)INFINITY;
      for (int c = 0; c < 7; ++c) margin = fmin(margin, fmin(q[c] - S.q_lo[c], S.q_hi[c] - q[c]));
      if (err > sc.max_pos_err) sc.max_pos_err = err;
      if (margin < sc.min_joint_margin) sc.min_joint_margin = margin;
      if (manip < sc.min_manipulability) sc.min_manipulability = manip;
      if (sc.first_bad_point < 0 && (!(err <= S.tol_pos) || margin < 0.0)) sc.first_bad_point = k;
    }
    sc.feasible = sc.first_bad_point < 0 ? 1 : 0;
  }
  for (int c = 0; c < 7; ++c) sc.q_final[c] = q[c];
  S.out[j] = sc;
}

// pose, pose Jacobian (8 x 7 row-major) and geometric Jacobian (6 x 7 row-major) of one configuration: what
// calculateControlPreliminaries computes per cycle; used by the tests to compare the device algebra with the oracle
__global__ void dq_probe_kernel(const double *base8, const double *q7, double *pose8, double *J56, double *G42) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  PandaModel m;
  panda_model(m);
  dq8 base, x[8];
  for (int r = 0; r < 8; ++r) base.q[r] = base8[r];
  panda_chain(m, base, q7, x);
  for (int r = 0; r < 8; ++r) pose8[r] = x[7].q[r];
  for (int c = 0; c < 7; ++c) {
    const dq8 col = pose_jacobian_col(m, x, c);
    double rot[3], tra[3];
    geom_jacobian_col(x[7], col, rot, tra);
    for (int r = 0; r < 8; ++r) J56[r * 7 + c] = col.q[r];
    for (int r = 0; r < 3; ++r) G42[r * 7 + c] = rot[r], G42[(r + 3) * 7 + c] = tra[r];
  }
}

}  // namespace pmaf
