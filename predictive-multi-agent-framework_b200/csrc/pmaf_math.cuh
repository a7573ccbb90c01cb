// pmaf_math.cuh — binary64 building blocks of the circular-field agent step, in the operation
// order of the reference (citations: /root/reference/src/bimanual_planning_ros/src/cf_agent.cpp
// unless noted). Every function is __host__ __device__ so that the same source can be checked on
// the CPU (tests/host_step_check.cu, tests/host_exp_check.cu) and runs inside the sm_100a kernels.
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

// ---- pinning loop invariants in registers ------------------------------------------------------------
// Kernel parameters live in the constant bank and ptxas re-reads them (LDCU/LDC) at every use inside
// the step loop — ~36 constant loads per agent-step, each followed by a dependent use. XOR-ing a value
// with a zero word that is only known at run time (loaded once from global memory) makes it an
// ordinary register value for the rest of the kernel without changing a bit of it.
PMAF_HD double keep(double x, unsigned runtime_zero) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(x) ^ (int)runtime_zero, __double2loint(x) ^ (int)runtime_zero);
#else
  (void)runtime_zero;
  return x;
#endif
}
PMAF_HD int keep(int x, unsigned runtime_zero) { return x ^ (int)runtime_zero; }

// ---- arithmetic policies ---------------------------------------------------------------------------
// The reference's x86-64 build rounds every sqrt and division correctly (IEEE). Two ways to get the
// same bits on the GPU:
//   ExactMath  CUDA's built-in IEEE sqrt / division. Always right; each one is a MUFU seed, ~8 FMAs
//              and a range check that branches to a slow-path subroutine — the branches serialise
//              independent chains, which is what bounds a latency-bound rollout.
//   FastMath   the same Newton/Goldschmidt refinements with the Markstein final corrections
//              (residual by FMA, one more FMA), written branch-free; one reciprocal is shared by the
//              three coefficients of a vector. They are proven only while no intermediate
//              under/overflows, so every operand is range-checked with two integer instructions on
//              its exponent field and a failed check raises `flag` instead of branching. The caller
//              evaluates a whole section (a pure function of registers) with FastMath and re-evaluates
//              it with ExactMath if the flag is up (zero / tiny / huge / NaN operands: rare).
//              pmaf_selftest_math compares FastMath with ExactMath on the GPU over random and
//              adversarial operands (tests/test_gpu_math.py). On the host FastMath is ExactMath.
// CUDA's IEEE sqrt / division as out-of-line calls: ExactMath only runs on cold paths (re-evaluation of
// a flagged section, the single real-agent step), and inlining ~40 instructions per operation there
// bloats the step loop's instruction footprint (the rollout is sensitive to instruction-cache misses).
#if defined(__CUDA_ARCH__)
#define PMAF_COLD_FN __device__ __noinline__
#else
#define PMAF_COLD_FN inline
#endif
PMAF_COLD_FN double cold_sqrt(double x) { return sqrt(x); }
PMAF_COLD_FN double cold_div(double a, double b) { return a / b; }
PMAF_COLD_FN double cold_exp(double x) { return exp(x); }

struct ExactMath {
  PMAF_HD double sqrt_(double x) { return cold_sqrt(x); }
  PMAF_HD double div_(double a, double b) { return cold_div(a, b); }
  PMAF_HD v3 div3_(v3 a, double b) { return mk3(cold_div(a.x, b), cold_div(a.y, b), cold_div(a.z, b)); }
  PMAF_HD bool bad() const { return false; }
};

struct FastMath {
  unsigned flag;
  PMAF_HD FastMath() : flag(0u) {}
  PMAF_HD bool bad() const { return flag != 0u; }
#if defined(__CUDA_ARCH__)
  // exponent field of x within [2^lo, 2^hi): also rejects zero, subnormals, negatives, inf, NaN
  __device__ __forceinline__ void need_range(double x, int lo, int hi) {
    const unsigned h = (unsigned)__double2hiint(x);
    flag |= (unsigned)((h - ((unsigned)(1023 + lo) << 20)) >= ((unsigned)(hi - lo) << 20));
  }
  __device__ __forceinline__ void need_range_or_zero(double x, int lo, int hi) {
    const unsigned h = (unsigned)__double2hiint(x) & 0x7fffffffu;
    flag |= (unsigned)(((h - ((unsigned)(1023 + lo) << 20)) >= ((unsigned)(hi - lo) << 20)) && x != 0.0);
  }
  // RN(sqrt(x)) for x in [2^-600, 2^600)
  __device__ __forceinline__ double sqrt_(double x) {
    need_range(x, -600, 600);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g), h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g), h = fma(h, r, h);
    const double d = fma(-g, g, x);  // exact residual
    return fma(d, h, g);
  }
  // reciprocal of b in [2^-300, 2^300), refined to within an ulp
  __device__ __forceinline__ double rcp_(double b) {
    need_range(b, -300, 300);
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    e = fma(-b, y, 1.0);
    return fma(y, e, y);
  }
  // RN(a / b) given y ~ 1/b: product, two residual corrections (the second is Markstein's final step)
  __device__ __forceinline__ double quot_core_(double a, double b, double y) {
    double q = a * y;
    double r = fma(-b, q, a);
    q = fma(r, y, q);
    r = fma(-b, q, a);
    q = fma(r, y, q);
    // b > 0, so the quotient has the numerator's sign; this also restores it on a zero numerator.
    // One LOP3 (bit select): magnitude bits of q, sign bit of a.
    unsigned hi;
    asm("lop3.b32 %0, %1, %2, 0x7fffffff, 0xE4;" : "=r"(hi) : "r"(__double2hiint(q)), "r"(__double2hiint(a)));
    return __hiloint2double((int)hi, __double2loint(q));
  }
  __device__ __forceinline__ double quot_(double a, double b, double y) {
    need_range_or_zero(a, -600, 600);
    return quot_core_(a, b, y);
  }
  // numerator known to be bounded by the (range-checked) divisor — a component of the vector whose norm
  // b is: only a non-zero magnitude below 2^-600 can leave the proven range (a NaN component makes the
  // norm NaN, which the root's own check rejects)
  __device__ __forceinline__ double quot_bounded_(double a, double b, double y) {
    flag |= (unsigned)((fabs(a) < 0x1p-600) & (a != 0.0));
    return quot_core_(a, b, y);
  }
  __device__ __forceinline__ double div_(double a, double b) { return quot_(a, b, rcp_(b)); }
  __device__ __forceinline__ v3 div3_(v3 a, double b) {
    const double y = rcp_(b);
    return mk3(quot_(a.x, b, y), quot_(a.y, b, y), quot_(a.z, b, y));
  }
  // a / |a| given b = |a| and y ~ 1/b (normalisations)
  __device__ __forceinline__ v3 quot3_(v3 a, double b, double y) {
    return mk3(quot_bounded_(a.x, b, y), quot_bounded_(a.y, b, y), quot_bounded_(a.z, b, y));
  }
  // s = RN(sqrt(x)) and y ~ 1/s to within an ulp for x in [2^-600, 2^600): the Goldschmidt iteration of
  // sqrt_ already carries h ~ 1/(2 sqrt(x)), so the reciprocal of the root costs one Newton step on 2h
  // instead of a second MUFU seed and two steps — a normalisation (sqrt, then three divisions by the
  // root) is five dependent operations shorter. s is in [2^-300, 2^300] by construction: no range check.
  __device__ __forceinline__ void sqrt_rcp_(double x, double &s, double &y) {
    need_range(x, -600, 600);
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double g = x * y0, h = 0.5 * y0;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g), h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g), h = fma(h, r, h);
    const double d = fma(-g, g, x);  // exact residual
    s = fma(d, h, g);
    const double y2 = h + h;
    const double e = fma(-s, y2, 1.0);
    y = fma(y2, e, y2);
  }
#else
  double sqrt_(double x) { return sqrt(x); }
  double div_(double a, double b) { return a / b; }
  v3 div3_(v3 a, double b) { return div3(a, b); }
  // host stand-ins (y is unused: the host divides)
  void need_range(double, int, int) {}
  double rcp_(double b) { return 1.0 / b; }
  double quot_(double a, double b, double) { return a / b; }
  v3 quot3_(v3 a, double b, double) { return div3(a, b); }
  void sqrt_rcp_(double x, double &s, double &y) { s = sqrt(x), y = 1.0 / s; }
#endif
};

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
    return cold_exp(x);                    // |x| >= 512, inf, nan: outside the planner's domain
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
  return cold_sqrt(z) < t.c;
}
PMAF_HD bool norm_gt(double z, const SqThr &t) {  // sqrt(z) > c
  if (z > t.hi) return true;
  if (z < t.lo) return false;
  return cold_sqrt(z) > t.c;
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
  double rsum_s;       // rad_ + sentinel radius          (repelForce :168)
  double repel_far2;   // squared distance beyond which the sentinel is certainly out of the shell
  bool unit_mass;      // force_ / 1.0 == force_ exactly
};
PMAF_HD AgentConsts make_agent_consts(double k_attr, double k_circ, double k_repel, double k_damp, double shell,
                                      double vel_max, double approach_dist, double mass, double rsum_s,
                                      unsigned z = 0u) {
  AgentConsts c;
  c.k_attr = k_attr, c.k_circ = k_circ, c.k_repel = k_repel, c.k_damp = k_damp;
  c.attr_ratio = k_attr / k_damp;
  c.inv_shell = 1.0 / shell;
  c.half_vmax = 0.5 * vel_max;
  c.vmax90 = vel_max - 0.1 * vel_max;
  c.shell = shell, c.vel_max = vel_max, c.approach_dist = approach_dist, c.mass = mass;
  c.rsum_s = rsum_s;
  const double far = (shell + rsum_s) + 1e-9;  // absolute margin >> rounding of n and n - rsum
  c.repel_far2 = far * far * (1.0 + 1e-15);
  c.unit_mass = mass == 1.0;
  c.k_attr = keep(c.k_attr, z), c.k_circ = keep(c.k_circ, z), c.k_repel = keep(c.k_repel, z);
  c.k_damp = keep(c.k_damp, z), c.attr_ratio = keep(c.attr_ratio, z), c.inv_shell = keep(c.inv_shell, z);
  c.half_vmax = keep(c.half_vmax, z), c.vmax90 = keep(c.vmax90, z), c.shell = keep(c.shell, z);
  c.vel_max = keep(c.vel_max, z), c.approach_dist = keep(c.approach_dist, z), c.mass = keep(c.mass, z);
  c.rsum_s = keep(c.rsum_s, z), c.repel_far2 = keep(c.repel_far2, z);
  return c;
}

// Eigen normalized() under a policy. The quotient is formed unconditionally (branch-free) and
// selected: z == 0 raises FastMath's flag, and ExactMath's 0/0 is discarded by the select.
template <class M>
PMAF_HD v3 normalized_m(M &m, v3 a) {
  const double z = dot3(a, a);
  const v3 q = m.div3_(a, m.sqrt_(z));
  return z > 0.0 ? q : a;
}
template <class M>
PMAF_HD v3 normalized_zn_m(M &m, v3 a, double z, double n) {
  const v3 q = m.div3_(a, n);
  return z > 0.0 ? q : a;
}
// `if (current.norm() < 1e-10) current << 0,0,1; current.normalize();` with one square root
template <class M>
PMAF_HD v3 normalized_or_z(M &m, v3 a) {
  const double z = dot3(a, a);
  const double n = m.sqrt_(z);
  const v3 q = m.div3_(a, n);
  if (n < 1e-10) return mk3(0.0, 0.0, 1.0);  // (0,0,1).normalize() is (0,0,1)
  return z > 0.0 ? q : a;
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
template <class M>
PMAF_HD v3 current_vector(M &m, int type, v3 p, v3 goal, v3 to_obs, v3 nv_eigen, v3 rot_i) {
  if (type == GOAL_HEURISTIC) {  // :389-406
    v3 goal_vec = sub3(goal, p);
    return normalized_or_z(m, sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec))));
  }
  if (type == VEL_HEURISTIC) {  // :520-537
    return normalized_or_z(m, sub3(nv_eigen, mul3(to_obs, dot3(nv_eigen, to_obs))));
  }
  // OBSTACLE :414-426, GOAL_OBSTACLE :463-475, RANDOM :545-557, HAD :585-597
  return normalized_m(m, cross3(to_obs, rot_i));
}

// circular-field force of one in-shell obstacle, :98-104
template <class M>
PMAF_HD v3 circ_force_term(M &m, double k_circ, double dist_obs, v3 nv, v3 current) {
  return mul3(cross3(nv, cross3(current, nv)), m.div_(k_circ, dist_obs * dist_obs));
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
PMAF_HD v3 add_repel_force(v3 force, v3 p, v3 o_s, const AgentConsts &c) {
  const v3 dv = sub3(p, o_s);
  const double z = dot3(dv, dv);
  v3 repel = mk3(0.0, 0.0, 0.0);
  if (!(z > c.repel_far2)) {  // otherwise out of the shell for sure: skip the sqrt
    const double n = cold_sqrt(z);
    const double d = clamp_dist(n - c.rsum_s);
    if (d < c.shell) {  // rare (self-collision sentinel within reach): out-of-line IEEE operations
      ExactMath em;
      const v3 u = normalized_zn_m(em, dv, z, n);
      const double s1 = cold_div(1.0, d) - c.inv_shell;
      const double s2 = d * d;
      repel = mk3(cold_div(c.k_repel * u.x * s1, s2), cold_div(c.k_repel * u.y * s1, s2),
                  cold_div(c.k_repel * u.z * s1, s2));
    }
  }
  const v3 total = add3(mk3(0.0, 0.0, 0.0), repel);  // total_repel_force += repel_force (:179)
  return add3(force, total);
}

// attractorForce :183-193
// first half: the velocity-limited desired velocity depends on the position only, so it is formed in
// the step prologue, off the critical path of the field evaluation
template <class M>
PMAF_HD v3 desired_velocity(M &m, v3 goal_vec, const AgentConsts &c) {
  v3 vel_des = mul3(goal_vec, c.attr_ratio);
  const double lim = m.div_(c.vel_max, m.sqrt_(dot3(vel_des, vel_des)));
  const double scale_lim = lim < 1.0 ? lim : 1.0;  // std::min(1.0, lim)
  return mul3(vel_des, scale_lim);
}
PMAF_HD v3 add_attractor_force(v3 force, v3 vel_des, v3 v, double k_goal_scale, const AgentConsts &c) {
  if (c.k_attr == 0.0) return force;
  return add3(force, mul3(sub3(vel_des, v), k_goal_scale * c.k_damp));
}

// tail of attractorForceScaling :212-226 once the closest in-shell obstacle (distance
// closest_d, position o_c) is known; dist_goal = |goal_vec|, vn = |v|
template <class M>
PMAF_HD double attractor_scaling(M &m, v3 goal_vec, double dist_goal, v3 p, v3 v, double vn, const AgentConsts &c,
                                 double closest_d, v3 o_c) {
  if (dot3(goal_vec, v) <= 0.0 && vn < c.vmax90 && dist_goal > 0.15) return 0.0;
  const double w1 = 1 - exp_glibc(m.div_(-m.sqrt_(closest_d), c.shell));
  const v3 rov = sub3(o_c, p);
  double w2 = 1 - m.div_(dot3(goal_vec, rov), dist_goal * m.sqrt_(dot3(rov, rov)));
  w2 = w2 * w2;
  return w1 * w2;
}

// updatePositionAndVelocity :253-268
template <class M>
PMAF_HD void integrate_step(M &m, v3 force, double dt, const AgentConsts &c, v3 &p, v3 &v) {
  ExactMath em;
  v3 acc = c.unit_mass ? force : em.div3_(force, c.mass);
  const double zacc = dot3(acc, acc);
  if (norm_gt(zacc, make_thr(13.0))) acc = mul3(acc, cold_div(13.0, cold_sqrt(zacc)));  // rare
  const v3 np = mk3((p.x + 0.5 * acc.x * dt * dt) + v.x * dt, (p.y + 0.5 * acc.y * dt * dt) + v.y * dt,
                    (p.z + 0.5 * acc.z * dt * dt) + v.z * dt);
  v = add3(v, mul3(acc, dt));
  const double vel_norm = m.sqrt_(dot3(v, v));
  const double scale = m.div_(c.vel_max, vel_norm);  // formed unconditionally, used only when clamping
  if (vel_norm > c.vel_max) v = mul3(v, scale);
  p = np;
}

// ---- sections of one step: pure functions of registers, evaluated under an arithmetic policy ------
struct StepNorms {
  double zg, dist_goal, zv, vn;  // |goal - p|^2, |goal - p|, |v|^2, |v|
  double seg_len;                // length of the path segment appended by the previous step (:29)
  v3 vel_des;                    // attractorForce's limited desired velocity (:189-191)
};
// Step prologue: four independent square roots (interleaved by the scheduler) and one division.
// zseg = |p - previous p|^2 (pass has_seg = false on the first step of a rollout).
template <class M>
PMAF_HD StepNorms step_norms(M &m, v3 goal_vec, v3 v, double zseg, bool has_seg, const AgentConsts &c) {
  StepNorms s;
  s.zg = dot3(goal_vec, goal_vec), s.zv = dot3(v, v);
  s.dist_goal = m.sqrt_(s.zg), s.vn = m.sqrt_(s.zv);
  const double seg = m.sqrt_(has_seg ? zseg : 1.0);
  s.seg_len = has_seg ? seg : 0.0;
  s.vel_des = desired_velocity(m, goal_vec, c);
  return s;
}
// goal_vec.normalized() (:79) and, for static scenes, rel_vel / vel_norm (:99) — one reciprocal each
template <bool STATIC_VEL, class M>
PMAF_HD void step_units(M &m, v3 goal_vec, v3 v, const StepNorms &s, v3 &ghat, v3 &nv_static) {
  ghat = normalized_zn_m(m, goal_vec, s.zg, s.dist_goal);
  nv_static = mk3(0.0, 0.0, 0.0);
  if (STATIC_VEL) {
    const v3 q = m.div3_(v, s.vn);
    if (s.vn != 0) nv_static = q;
  }
}
// everything after the circular force and its attractor scaling: repulsion, attraction, integration
template <class M>
PMAF_HD void finish_step(M &m, v3 &force, double k_goal_scale, const StepNorms &s, v3 o_sentinel, double dt,
                         const AgentConsts &c, v3 &p, v3 &v) {
  force = add_repel_force(force, p, o_sentinel, c);
  force = add_attractor_force(force, s.vel_des, v, k_goal_scale, c);
  integrate_step(m, force, dt, c, p, v);
}

// CfAgent::setVelocity :54-61
PMAF_HD v3 clamp_velocity(v3 v, double vel_max) {
  double n = norm3(v);
  return n > vel_max ? mul3(v, vel_max / n) : v;
}

// the same through the out-of-line IEEE operations (once per rollout, in the kernel's prologue)
PMAF_HD v3 clamp_velocity_cold(v3 v, double vel_max) {
  const double n = cold_sqrt(dot3(v, v));
  return n > vel_max ? mul3(v, cold_div(vel_max, n)) : v;
}

// workspace term of one path point, CfManager::evaluateAgents cf_manager.cpp:302-323
// ws = [x+, x-, y+, y-, z+, z-]
struct WsParams {
  double ws[6], k_workspace;
};
PMAF_HD WsParams pin_ws(const double *ws, double k_workspace, unsigned z) {
  WsParams w;
  for (int i = 0; i < 6; ++i) w.ws[i] = keep(ws[i], z);
  w.k_workspace = keep(k_workspace, z);
  return w;
}
PMAF_HD double add_workspace_cost(double cost, v3 q, const double *ws, double k_workspace) {
  // one combined test first: a point inside the workspace (the common case) costs six independent
  // compares and a single branch instead of six dependent compare-and-branch pairs
  const bool outside = (q.x > ws[0]) | (q.x < ws[1]) | (q.y > ws[2]) | (q.y < ws[3]) | (q.z > ws[4]) | (q.z < ws[5]);
  if (!outside) return cost;
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

// The same term without a branch (selects; adding +0.0 for an axis inside its limits is exact because the
// running cost is a sum of squares starting at +0.0 and can never be -0.0): it can share a basic block
// with other work. NaN coordinates add nothing, as in the reference (both comparisons false).
PMAF_HD double add_workspace_cost_bf(double cost, v3 q, const double *ws, double k_workspace) {
  const double qs[3] = {q.x, q.y, q.z};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 3; ++a) {
    const bool hi = qs[a] > ws[2 * a], lo = qs[a] < ws[2 * a + 1];
    const double lim = hi ? ws[2 * a] : ws[2 * a + 1];
    double t = fabs(qs[a] - lim) * k_workspace;
    t = (hi | lo) ? t : 0.0;
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
