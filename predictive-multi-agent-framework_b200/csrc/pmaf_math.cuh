// pmaf_math.cuh — binary64 building blocks of the circular-field agent step, in the operation
// order of the reference (citations: /root/reference/src/bimanual_planning_ros/src/cf_agent.cpp
// unless noted). Every function is __host__ __device__ so that the same source can be checked on
// the CPU (tests/host_math_check.cpp) and runs inside the sm_100a kernels.
//
// Exactness contract: compile with -fmad=false (device) / -ffp-contract=off (host). The
// reference is built for default x86-64 (no FMA), Eigen 3.3 fixed-size vectors:
//   dot / squaredNorm reduce as (x0*y0 + x1*y1) + x2*y2, norm = sqrt(squaredNorm),
//   normalized(): z = squaredNorm; z > 0 ? v / sqrt(z) : v   (true divisions),
//   cross(): (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0).
// Division and sqrt are IEEE correctly rounded on both sides (nvcc default -prec-div/-prec-sqrt).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "pmaf_exp_table.inc"

#if defined(__CUDACC__)
#define PMAF_HD __host__ __device__ __forceinline__
#else
#define PMAF_HD inline
#endif

namespace pmaf {

// CfAgent::Type, cf_agent.h:59-68
enum AgentType : int {
  REAL_AGENT = 0,
  GOAL_HEURISTIC = 1,
  OBSTACLE_HEURISTIC = 2,
  GOAL_OBSTACLE_HEURISTIC = 3,
  VEL_HEURISTIC = 4,
  RANDOM_AGENT = 5,
  HAD_HEURISTIC = 6,
  UNDEFINED_AGENT = 7
};

// type of the agent at GLOBAL index idx, CfManager::init cf_manager.cpp:70-104
PMAF_HD int agent_type_of_index(int idx) {
  switch (idx) {
    case 0: return HAD_HEURISTIC;
    case 1: return GOAL_HEURISTIC;
    case 2: return OBSTACLE_HEURISTIC;
    case 3: return GOAL_OBSTACLE_HEURISTIC;
    case 4: return VEL_HEURISTIC;
    default: return RANDOM_AGENT;
  }
}

struct v3 {
  double x, y, z;
};

PMAF_HD v3 mk3(double x, double y, double z) {
  v3 r;
  r.x = x, r.y = y, r.z = z;
  return r;
}
PMAF_HD v3 ld3(const double *p) { return mk3(p[0], p[1], p[2]); }
PMAF_HD void st3(double *p, v3 a) { p[0] = a.x, p[1] = a.y, p[2] = a.z; }
PMAF_HD v3 add3(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PMAF_HD v3 sub3(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PMAF_HD v3 mul3(v3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
PMAF_HD v3 div3(v3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
PMAF_HD double dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
PMAF_HD double norm3(v3 a) { return sqrt(dot3(a, a)); }
PMAF_HD v3 cross3(v3 a, v3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Eigen normalized(): given z = squaredNorm and n = sqrt(z) already computed
PMAF_HD v3 normalized_zn(v3 a, double z, double n) { return z > 0.0 ? div3(a, n) : a; }
PMAF_HD v3 normalized3(v3 a) {
  double z = dot3(a, a);
  return z > 0.0 ? div3(a, sqrt(z)) : a;
}
// std::max(d, 1e-5) (cf_agent.cpp:85): NaN stays NaN
PMAF_HD double clamp_dist(double d) { return d < 1e-5 ? 1e-5 : d; }

// ---- exp() of the host libm ------------------------------------------------------------------------
// attractorForceScaling calls std::exp (:220). CUDA's exp() and glibc's differ in the last bit for a
// few percent of the arguments, and the rollout amplifies a one-ulp difference through its hard
// thresholds, so the kernels evaluate exp with glibc's own algorithm (sysdeps/ieee754/dbl-64/e_exp.c,
// ARM optimized-routines: N = 128 table + degree-5 polynomial) in the operation order of the FMA
// build that x86-64 glibc selects at run time (__exp_fma): every a*b+c of the C source is one fused
// multiply-add. tests/test_exp.py checks it bit for bit against the host's exp(). Constants and table
// come from the system libm (tools/gen_exp_table.py).
#if defined(__CUDACC__)
static __device__ const uint64_t d_exp_tab[256] = {PMAF_EXP_TABLE};
#endif
static const uint64_t h_exp_tab[256] = {PMAF_EXP_TABLE};

PMAF_HD uint64_t exp_tab(unsigned i) {
#if defined(__CUDA_ARCH__)
  return d_exp_tab[i];
#else
  return h_exp_tab[i];
#endif
}
PMAF_HD uint64_t bits_of(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, sizeof u);
  return u;
#endif
}
PMAF_HD double double_of(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, sizeof x);
  return x;
#endif
}

PMAF_HD double exp_glibc(double x) {
  const unsigned abstop = (unsigned)(bits_of(x) >> 52) & 0x7ffu;
  // main path of __exp: 2^-54 <= |x| < 512 (no special-casing of the scale needed)
  if (abstop - 0x3c9u >= 0x408u - 0x3c9u) {
    if (abstop < 0x3c9u) return 1.0 + x;  // tiny |x| (WANT_ROUNDING)
    return exp(x);                         // |x| >= 512, inf, nan: outside the planner's domain
  }
  double kd = fma(kExpInvLn2N, x, kExpShift);
  const uint64_t ki = bits_of(kd);
  kd -= kExpShift;
  const double r = fma(kd, kExpNegLn2loN, fma(kd, kExpNegLn2hiN, x));
  const unsigned idx = 2u * (unsigned)(ki % 128u);
  const uint64_t top = ki << (52 - 7);
  const double tail = double_of(exp_tab(idx));
  const uint64_t sbits = exp_tab(idx + 1) + top;
  const double r2 = r * r;
  const double p23 = fma(r, kExpC3, kExpC2), p45 = fma(r, kExpC5, kExpC4);
  const double tmp = fma(r2 * r2, p45, fma(r2, p23, tail + r));
  const double scale = double_of(sbits);
  return fma(scale, tmp, scale);
}

// ---- comparing a norm with a constant without taking the square root ---------------------------------
// sqrt is monotonic and correctly rounded, so "sqrt(z) < c" can be decided on z whenever z is not
// within a few ulps of c*c; only then is the square root evaluated. The decision is always the one
// the reference's `v.norm() < c` makes (NaN falls through to the exact comparison).
struct SqThr {
  double lo, hi, c;
};
PMAF_HD SqThr make_thr(double c) {
  SqThr t;
  t.c = c;
  if (c > 0.0) {
    const double c2 = c * c;
    t.lo = c2 * (1.0 - 1e-15), t.hi = c2 * (1.0 + 1e-15);
  } else {  // c <= 0 or NaN: always take the exact path
    t.lo = -(double)INFINITY, t.hi = (double)INFINITY;
  }
  return t;
}
PMAF_HD bool norm_lt(double z, const SqThr &t) {  // sqrt(z) < c
  if (z < t.lo) return true;
  if (z > t.hi) return false;
  return sqrt(z) < t.c;
}
PMAF_HD bool norm_gt(double z, const SqThr &t) {  // sqrt(z) > c
  if (z > t.hi) return true;
  if (z < t.lo) return false;
  return sqrt(z) > t.c;
}

// Per-agent values the reference recomputes identically every step (same operands, same operation:
// hoisting them changes nothing bitwise).
struct AgentConsts {
  double k_attr, k_circ, k_repel, k_damp;
  double attr_ratio;   // k_attr / k_damp              (attractorForce :189)
  double inv_shell;    // 1.0 / detect_shell_rad_      (repelForce :176)
  double half_vmax;    // 0.5 * vel_max_               (gate :288)
  double vmax90;       // vel_max_ - 0.1 * vel_max_    (attractorForceScaling :216)
  double shell, vel_max, approach_dist, mass;
  bool unit_mass;      // force_ / 1.0 == force_ exactly
};
PMAF_HD AgentConsts make_agent_consts(double k_attr, double k_circ, double k_repel, double k_damp, double shell,
                                      double vel_max, double approach_dist, double mass) {
  AgentConsts c;
  c.k_attr = k_attr, c.k_circ = k_circ, c.k_repel = k_repel, c.k_damp = k_damp;
  c.attr_ratio = k_attr / k_damp;
  c.inv_shell = 1.0 / shell;
  c.half_vmax = 0.5 * vel_max;
  c.vmax90 = vel_max - 0.1 * vel_max;
  c.shell = shell, c.vel_max = vel_max, c.approach_dist = approach_dist, c.mass = mass;
  c.unit_mass = mass == 1.0;
  return c;
}

// `if (current.norm() < 1e-10) current << 0,0,1; current.normalize();` with one square root
PMAF_HD v3 normalized_or_z(v3 a) {
  const double z = dot3(a, a);
  const double n = sqrt(z);
  if (n < 1e-10) return mk3(0.0, 0.0, 1.0);  // (0,0,1).normalize() is (0,0,1)
  return z > 0.0 ? div3(a, n) : a;
}

// ---- rotation vectors (calculateRotationVector) ------------------------------------------------
// to_obs = normalized(o_i - p), the same value circForce already formed for its skip test.

// HAD :599-611 — NaN when d is parallel to the goal vector
PMAF_HD v3 rot_had(v3 p, v3 goal, v3 o_i) {
  v3 goal_vec = sub3(goal, p);
  v3 rob_obs = sub3(o_i, p);
  double gn = norm3(goal_vec);
  double s = dot3(rob_obs, goal_vec) / (gn * gn);
  v3 d = sub3(add3(p, mul3(goal_vec, s)), o_i);
  v3 c = cross3(d, goal_vec);
  return div3(c, norm3(c));
}
// RANDOM :559-566 — not normalised
PMAF_HD v3 rot_random(v3 p, v3 goal, v3 random_i) { return cross3(normalized3(sub3(goal, p)), random_i); }
// OBSTACLE :447-460, o_c = position of the obstacle closest to obstacle i (:434-446)
PMAF_HD v3 rot_obstacle(v3 to_obs, v3 o_i, v3 o_c) {
  v3 obstacle_vec = sub3(o_c, o_i);
  v3 current = sub3(mul3(to_obs, dot3(obstacle_vec, to_obs)), obstacle_vec);
  return normalized3(cross3(current, to_obs));
}
// GOAL_OBSTACLE :493-517
PMAF_HD v3 rot_goal_obstacle(v3 p, v3 goal, v3 to_obs, v3 o_i, v3 o_c) {
  v3 obstacle_vec = sub3(o_c, o_i);
  v3 obst_current = sub3(mul3(to_obs, dot3(obstacle_vec, to_obs)), obstacle_vec);
  v3 goal_vec = sub3(goal, p);
  v3 goal_current = sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec)));
  v3 current = add3(normalized3(goal_current), normalized3(obst_current));
  if (norm3(current) < 1e-10) current = mk3(0.0, 0.0, 1.0);
  current = normalized3(current);
  return normalized3(cross3(current, to_obs));
}

// ---- current vectors (currentVector) -------------------------------------------------------------
// rel = relative velocity (the caller passes rel_vel as agent_vel, :100), nv_eigen = rel.normalized()
PMAF_HD v3 current_vector(int type, v3 p, v3 goal, v3 to_obs, v3 nv_eigen, v3 rot_i) {
  if (type == GOAL_HEURISTIC) {  // :389-406
    v3 goal_vec = sub3(goal, p);
    return normalized_or_z(sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec))));
  }
  if (type == VEL_HEURISTIC) {  // :520-537
    return normalized_or_z(sub3(nv_eigen, mul3(to_obs, dot3(nv_eigen, to_obs))));
  }
  // OBSTACLE :414-426, GOAL_OBSTACLE :463-475, RANDOM :545-557, HAD :585-597
  return normalized3(cross3(to_obs, rot_i));
}

// circular-field force of one in-shell obstacle, :98-104
PMAF_HD v3 circ_force_term(double k_circ, double dist_obs, v3 nv, v3 current) {
  return mul3(cross3(nv, cross3(current, nv)), k_circ / (dist_obs * dist_obs));
}

// ---- scalar parts of one step ----------------------------------------------------------------------
// gate of cfPlanner / cfPrediction, :287-289 / :315-317 / :352-354
// dist_goal = |goal - p|, vn = |v| (both needed again later in the step)
PMAF_HD bool field_gate_open(double dist_goal, double vn, v3 p, v3 init_pos, const AgentConsts &c) {
  if (dist_goal < c.approach_dist) return false;
  if (!(vn < c.half_vmax)) return true;
  const v3 d = sub3(p, init_pos);
  return !norm_lt(dot3(d, d), make_thr(0.2));
}

// repelForce :159-181 on the sentinel (last obstacle); rsum = rad_ + sentinel radius.
PMAF_HD v3 add_repel_force(v3 force, v3 p, v3 o_s, double rsum, const AgentConsts &c) {
  const v3 dv = sub3(p, o_s);
  const double z = dot3(dv, dv);
  v3 repel = mk3(0.0, 0.0, 0.0);
  // out of the shell for sure (absolute margin 1e-9 >> rounding of n and n - rsum): skip the sqrt
  const double far = (c.shell + rsum) + 1e-9;
  if (!(z > far * far * (1.0 + 1e-15))) {
    const double n = sqrt(z);
    const double d = clamp_dist(n - rsum);
    if (d < c.shell) {
      const v3 u = normalized_zn(dv, z, n);
      const double s1 = 1.0 / d - c.inv_shell;
      const double s2 = d * d;
      repel = mk3(c.k_repel * u.x * s1 / s2, c.k_repel * u.y * s1 / s2, c.k_repel * u.z * s1 / s2);
    }
  }
  const v3 total = add3(mk3(0.0, 0.0, 0.0), repel);  // total_repel_force += repel_force (:179)
  return add3(force, total);
}

// attractorForce :183-193
PMAF_HD v3 add_attractor_force(v3 force, v3 goal_vec, v3 v, double k_goal_scale, const AgentConsts &c) {
  if (c.k_attr == 0.0) return force;
  v3 vel_des = mul3(goal_vec, c.attr_ratio);
  const double lim = c.vel_max / norm3(vel_des);
  const double scale_lim = lim < 1.0 ? lim : 1.0;  // std::min(1.0, lim)
  vel_des = mul3(vel_des, scale_lim);
  return add3(force, mul3(sub3(vel_des, v), k_goal_scale * c.k_damp));
}

// tail of attractorForceScaling :212-226 once the closest in-shell obstacle (distance
// closest_d, position o_c) is known; dist_goal = |goal_vec|, vn = |v|
PMAF_HD double attractor_scaling(v3 goal_vec, double dist_goal, v3 p, v3 v, double vn, const AgentConsts &c,
                                 double closest_d, v3 o_c) {
  if (dot3(goal_vec, v) <= 0.0 && vn < c.vmax90 && dist_goal > 0.15) return 0.0;
  const double w1 = 1 - exp_glibc(-sqrt(closest_d) / c.shell);
  const v3 rov = sub3(o_c, p);
  double w2 = 1 - (dot3(goal_vec, rov) / (dist_goal * norm3(rov)));
  w2 = w2 * w2;
  return w1 * w2;
}

// updatePositionAndVelocity :253-268
PMAF_HD void integrate_step(v3 force, double dt, const AgentConsts &c, v3 &p, v3 &v) {
  v3 acc = c.unit_mass ? force : div3(force, c.mass);
  const double zacc = dot3(acc, acc);
  if (norm_gt(zacc, make_thr(13.0))) acc = mul3(acc, 13.0 / sqrt(zacc));
  const v3 np = mk3((p.x + 0.5 * acc.x * dt * dt) + v.x * dt, (p.y + 0.5 * acc.y * dt * dt) + v.y * dt,
                    (p.z + 0.5 * acc.z * dt * dt) + v.z * dt);
  v = add3(v, mul3(acc, dt));
  const double vel_norm = norm3(v);
  if (vel_norm > c.vel_max) v = mul3(v, c.vel_max / vel_norm);
  p = np;
}

// CfAgent::setVelocity :54-61
PMAF_HD v3 clamp_velocity(v3 v, double vel_max) {
  double n = norm3(v);
  return n > vel_max ? mul3(v, vel_max / n) : v;
}

// workspace term of one path point, CfManager::evaluateAgents cf_manager.cpp:302-323
// ws = [x+, x-, y+, y-, z+, z-]
PMAF_HD double add_workspace_cost(double cost, v3 q, const double *ws, double k_workspace) {
  double t;
  if (q.x > ws[0]) {
    t = fabs(q.x - ws[0]) * k_workspace;
    cost += t * t;
  } else if (q.x < ws[1]) {
    t = fabs(q.x - ws[1]) * k_workspace;
    cost += t * t;
  }
  if (q.y > ws[2]) {
    t = fabs(q.y - ws[2]) * k_workspace;
    cost += t * t;
  } else if (q.y < ws[3]) {
    t = fabs(q.y - ws[3]) * k_workspace;
    cost += t * t;
  }
  if (q.z > ws[4]) {
    t = fabs(q.z - ws[4]) * k_workspace;
    cost += t * t;
  } else if (q.z < ws[5]) {
    t = fabs(q.z - ws[5]) * k_workspace;
    cost += t * t;
  }
  return cost;
}

// remaining terms of the per-agent cost, cf_manager.cpp:324-333
PMAF_HD double finish_cost(double ws_cost, double goal_dist, double approach_dist, double k_goal_dist,
                           double path_len, double k_path_len, double k_safe_dist, double min_obs_dist) {
  double cost = ws_cost;
  if (goal_dist > approach_dist) cost += goal_dist * k_goal_dist;
  cost += path_len * k_path_len;
  cost += k_safe_dist / min_obs_dist;
  if (min_obs_dist < 2e-5) cost += 10000.0;
  return cost;
}

}  // namespace pmaf
