/* TEST INFRASTRUCTURE — not product code.
 *
 * CPU restatement of the dual-quaternion kinematics the reference's controller evaluates downstream of the
 * planner. Citations are file:line under /root/reference/src/bimanual_planning_ros/.
 *
 *   Panda::kinematics            src/franka_robot.cpp:6-22   modified-DH table of the Franka Panda, base frame
 *   calculateControlPreliminaries src/costp_controller.cpp:111-126  pose (fkm) and pose Jacobian of the arm
 *   geomJ                         src/costp_controller.cpp:465-492  geometric Jacobian from the pose Jacobian
 *   joint limits                  src/costp_controller.cpp:41-44
 *   damped pseudo-inverse         src/costp_controller.cpp:134-135  J^T (J J^T + lambda I)^-1
 *
 * PARITY UNPINNED. The arithmetic lives in the third-party library dqrobotics (`DQ`, `DQ_SerialManipulator`,
 * "modified" DH convention; installed from its stable PPA, version not pinned by the reference, README.md:31),
 * which is absent from /root/reference and from this image, and the reference holds no test, golden vector or
 * known answer for it. What is restated here is dqrobotics' PUBLISHED algorithm — unit dual quaternions
 * x = r + (eps/2) t r with Hamilton products, the link transform of the modified Denavit-Hartenberg convention
 * Rot_x(alpha) Trans_x(a) Rot_z(theta) Trans_z(d), the pose Jacobian column (1/2) (x_{i-1} w_i x_{i-1}^*) x_n with
 * w_i the joint axis seen from frame i-1 — anchored on the reference's own call sites above and checked against
 * first principles (4x4 homogeneous transforms, finite differences) by tests/test_dq_oracle.py. Last-bit
 * agreement with a dqrobotics build is NOT claimed.
 *
 * The path scorer (dqo_score_path) is this repository's use of that kinematics (SURVEY.md f4 "candidate use"): a
 * damped-least-squares tracking of a predicted end-effector path, one step per path point, reporting tracking
 * error, joint-limit margin and manipulability. The CUDA kernel (csrc/pmaf_dq.cuh) is checked against it.
 *
 * Build: gcc -O2 -ffp-contract=off.
 */
#include <math.h>
#include <string.h>

#define DQO_API __attribute__((visibility("default")))

/* Panda, modified DH (src/franka_robot.cpp:7-13): theta offset, d, a, alpha per joint */
static const double kD[7] = {0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.2104};
static const double kA[7] = {0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088};
static const double kAlphaHalfPis[7] = {0.0, -1.0, 1.0, 1.0, -1.0, 1.0, 1.0}; /* alpha = n * pi/2 */

/* ---- quaternion / dual quaternion algebra (q[0..3] primary w,x,y,z; q[4..7] dual) -------------------- */
static void quat_mul(const double *a, const double *b, double *o) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
static void dq_mul(const double *a, const double *b, double *o) {
  double p[4], d1[4], d2[4];
  quat_mul(a, b, p);
  quat_mul(a, b + 4, d1);
  quat_mul(a + 4, b, d2);
  for (int i = 0; i < 4; ++i) o[i] = p[i], o[4 + i] = d1[i] + d2[i];
}
static void dq_conj(const double *a, double *o) {
  o[0] = a[0], o[1] = -a[1], o[2] = -a[2], o[3] = -a[3];
  o[4] = a[4], o[5] = -a[5], o[6] = -a[6], o[7] = -a[7];
}

/* constant part of link i: Rot_x(alpha_i) Trans_x(a_i) */
static void link_const(int i, double *A) {
  const double half = 0.5 * (kAlphaHalfPis[i] * M_PI_2);
  const double c = cos(half), s = sin(half);
  const double r[8] = {c, s, 0, 0, 0, 0, 0, 0};
  const double t[8] = {1, 0, 0, 0, 0, 0.5 * kA[i], 0, 0};
  dq_mul(r, t, A);
}
/* joint-dependent part: Rot_z(theta) Trans_z(d_i) */
static void link_joint(int i, double theta, double *B) {
  const double c = cos(0.5 * theta), s = sin(0.5 * theta);
  B[0] = c, B[1] = 0, B[2] = 0, B[3] = s;
  B[4] = -0.5 * kD[i] * s, B[5] = 0, B[6] = 0, B[7] = 0.5 * kD[i] * c;
}

/* x[i] = base * link_0 ... link_{i-1}, i = 0..7 (x[7] = the end-effector pose: fkm) */
static void chain(const double *base, const double *q, double x[8][8]) {
  memcpy(x[0], base, 8 * sizeof(double));
  for (int i = 0; i < 7; ++i) {
    double A[8], B[8], L[8];
    link_const(i, A);
    link_joint(i, q[i], B);
    dq_mul(A, B, L);
    dq_mul(x[i], L, x[i + 1]);
  }
}

DQO_API void dqo_fkm(const double *base, const double *q, double *out) {
  double x[8][8];
  chain(base, q, x);
  memcpy(out, x[7], 8 * sizeof(double));
}

/* pose Jacobian, 8 x 7 row-major: d vec8(fkm) / d q_i (DQ_SerialManipulator::pose_jacobian) */
DQO_API void dqo_pose_jacobian(const double *base, const double *q, double *J) {
  double x[8][8];
  chain(base, q, x);
  const double k[8] = {0, 0, 0, 1, 0, 0, 0, 0};
  for (int i = 0; i < 7; ++i) {
    double A[8], Ac[8], w[8], xa[8], xac[8], z[8], tmp[8], col[8];
    link_const(i, A); /* the joint axis k of frame i, seen from frame i-1: w = A k A^* */
    dq_conj(A, Ac);
    dq_mul(A, k, tmp);
    dq_mul(tmp, Ac, w);
    dq_conj(x[i], xac);
    dq_mul(x[i], w, xa);
    dq_mul(xa, xac, z);
    dq_mul(z, x[7], col);
    for (int r = 0; r < 8; ++r) J[r * 7 + i] = 0.5 * col[r];
  }
}

/* translation of a unit dual quaternion: t = 2 D P^* (vector part) */
DQO_API void dqo_translation(const double *x, double *t) {
  const double pc[4] = {x[0], -x[1], -x[2], -x[3]};
  double o[4];
  quat_mul(x + 4, pc, o);
  t[0] = 2 * o[1], t[1] = 2 * o[2], t[2] = 2 * o[3];
}

/* Hamilton operators of a quaternion: hamiplus4(a) b = a b, haminus4(a) b = b a */
static void hamiplus4(const double *a, double H[4][4]) {
  const double m[4][4] = {{a[0], -a[1], -a[2], -a[3]}, {a[1], a[0], -a[3], a[2]}, {a[2], a[3], a[0], -a[1]}, {a[3], -a[2], a[1], a[0]}};
  memcpy(H, m, sizeof m);
}
static void haminus4(const double *a, double H[4][4]) {
  const double m[4][4] = {{a[0], -a[1], -a[2], -a[3]}, {a[1], a[0], a[3], -a[2]}, {a[2], -a[3], a[0], a[1]}, {a[3], a[2], -a[1], a[0]}};
  memcpy(H, m, sizeof m);
}

/* geomJ (src/costp_controller.cpp:465-492) for one arm: rows 0-2 = 2 haminus4(P^*) J_P, rows 3-5 =
 * 2 (hamiplus4(D) C4 J_P + haminus4(P^*) J_D), vector parts only; 6 x 7 row-major. Rows 3-5 are the translation
 * Jacobian d t / d q. */
DQO_API void dqo_geom_jacobian(const double *base, const double *q, double *G) {
  double x[8], J[56];
  dqo_fkm(base, q, x);
  dqo_pose_jacobian(base, q, J);
  const double pc[4] = {x[0], -x[1], -x[2], -x[3]};
  double Hm[4][4], Hp[4][4];
  haminus4(pc, Hm);
  hamiplus4(x + 4, Hp);
  const double c4[4] = {1, -1, -1, -1};
  for (int c = 0; c < 7; ++c) {
    for (int r = 1; r < 4; ++r) {
      double rot = 0, tra = 0;
      for (int j = 0; j < 4; ++j) {
        rot += Hm[r][j] * J[j * 7 + c];
        tra += Hp[r][j] * c4[j] * J[j * 7 + c] + Hm[r][j] * J[(4 + j) * 7 + c];
      }
      G[(r - 1) * 7 + c] = 2 * rot;
      G[(r + 2) * 7 + c] = 2 * tra;
    }
  }
}

typedef struct {
  double max_pos_err;         /* largest residual |target - t(q)| after a point's step */
  double min_joint_margin;    /* smallest distance of any joint to its nearer limit (negative: violated) */
  double min_manipulability;  /* smallest sqrt(det(Jt Jt^T)) met along the path */
  int feasible;               /* every point tracked within tol_pos and inside the joint limits */
  int first_bad_point;        /* index of the first point that failed, -1 if none */
  double q_final[7];
} dqo_path_score;

static double det3(const double A[3][3]) {
  return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
         A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}

/* Damped-least-squares tracking of an end-effector path, one step per point:
 *   e = p_k - t(q);  dq = Jt^T (Jt Jt^T + damping I)^-1 e  (src/costp_controller.cpp:134-135's form);  q += dq. */
DQO_API void dqo_score_path(const double *base, const double *q_start, const double *path, int n_points, const double *q_lo,
                            const double *q_hi, double damping, double tol_pos, dqo_path_score *out) {
  double q[7];
  memcpy(q, q_start, sizeof q);
  double max_err = 0.0, min_margin = INFINITY, min_manip = INFINITY;
  int first_bad = -1;
  for (int k = 0; k < n_points; ++k) {
    double x[8], t[3], G[42], e[3];
    dqo_fkm(base, q, x);
    dqo_translation(x, t);
    dqo_geom_jacobian(base, q, G);
    const double *Jt = G + 21; /* rows 3-5 */
    for (int i = 0; i < 3; ++i) e[i] = path[3 * k + i] - t[i];
    double M[3][3], A[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int c = 0; c < 7; ++c) s += Jt[i * 7 + c] * Jt[j * 7 + c];
        M[i][j] = s, A[i][j] = s + (i == j ? damping : 0.0);
      }
    const double manip = sqrt(fmax(det3(M), 0.0));
    /* y = A^-1 e by Cramer's rule */
    const double dA = det3(A);
    double y[3];
    for (int c = 0; c < 3; ++c) {
      double B[3][3];
      memcpy(B, A, sizeof B);
      for (int r = 0; r < 3; ++r) B[r][c] = e[r];
      y[c] = det3(B) / dA;
    }
    for (int c = 0; c < 7; ++c) q[c] += Jt[0 * 7 + c] * y[0] + Jt[1 * 7 + c] * y[1] + Jt[2 * 7 + c] * y[2];
    dqo_fkm(base, q, x);
    dqo_translation(x, t);
    double r2 = 0;
    for (int i = 0; i < 3; ++i) r2 += (path[3 * k + i] - t[i]) * (path[3 * k + i] - t[i]);
    const double err = sqrt(r2);
    double margin = INFINITY;
    for (int c = 0; c < 7; ++c) margin = fmin(margin, fmin(q[c] - q_lo[c], q_hi[c] - q[c]));
    if (err > max_err) max_err = err;
    if (margin < min_margin) min_margin = margin;
    if (manip < min_manip) min_manip = manip;
    if (first_bad < 0 && (!(err <= tol_pos) || margin < 0.0)) first_bad = k;
  }
  out->max_pos_err = max_err, out->min_joint_margin = min_margin, out->min_manipulability = min_manip;
  out->feasible = first_bad < 0, out->first_bad_point = first_bad;
  memcpy(out->q_final, q, sizeof q);
}
