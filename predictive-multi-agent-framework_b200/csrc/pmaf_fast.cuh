// pmaf_fast.cuh — straight-line version of one agent step for the latency-bound rollout.
//
// A small population (BASELINE config C2: 256 agents = 256 warps on 592 schedulers) runs ONE warp per
// scheduler: the rollout's time is steps x (latency of one step), and a single warp issues in order.
// ncu on the general step (pmaf_rollout.cuh: agent_step / field_pass) shows ~1100 warp instructions per
// step at ~5 cycles each: ~50 branches per step end the compiler's scheduling regions, so the independent
// chains of a step (current vector, attractor scaling, reductions) run back to back instead of
// interleaved, and every branch costs its resolution plus an instruction-fetch bubble. A single warp
// does overlap independent FP64 work when it is given any (tools/ubench_ilp.cu: 8.8 -> 2.4 cycles per DFMA
// with 1 -> 8 independent chains).
//
// fast_step evaluates the COMMON case of a step (cf_agent.cpp:312-326) as two basic blocks with
// selects instead of branches and no memory side effects. Everything uncommon raises `rare`:
//   first detection of an obstacle (rotation vector latch), more than 32 broad-phase candidates, an
//   operand outside FastMath's proven range, a norm within 1e-15 of a threshold, the sentinel within
//   its shell, the acceleration clamp, non-unit mass, exp() outside its main path.
// A rare step returns false before anything is committed and the caller runs the general step from
// the same state. The arithmetic (operation order, roundings) is the general step's, which is the
// reference's; tests/test_gpu_parity.py compares whole rollouts bit for bit either way.
#pragma once
// included by pmaf_rollout.cuh after StepEnv / Prologue

namespace pmaf {

struct FastConsts {   // per-agent loop invariants of the fast step
  double y_shell;     // refined reciprocal of detect_shell_rad_ (division by the shell radius, :220)
  SqThr thr_force;    // |F| > 1e-5 (:319)
  SqThr thr_acc;      // |a| > 13   (:256)
  SqThr thr_start;    // |p - p0| < 0.2 (:316)
  bool usable;        // unit mass, shell inside FastMath's range
};

#if defined(__CUDACC__)
__device__ __forceinline__ FastConsts make_fast_consts(const AgentConsts &c, unsigned rz) {
  FastConsts f;
  FastMath m;
  f.y_shell = keep(m.rcp_(c.shell), rz);
  f.thr_force = make_thr(1e-5), f.thr_acc = make_thr(13.0), f.thr_start = make_thr(0.2);
  f.thr_force.lo = keep(f.thr_force.lo, rz), f.thr_force.hi = keep(f.thr_force.hi, rz);
  f.thr_acc.lo = keep(f.thr_acc.lo, rz);
  f.thr_start.lo = keep(f.thr_start.lo, rz), f.thr_start.hi = keep(f.thr_start.hi, rz);
  f.usable = !m.bad() && c.unit_mass;
  return f;
}

// glibc exp (exp_glibc, pmaf_math.cuh) on its main path only: 2^-54 <= |x| < 512, flagged otherwise
__device__ __forceinline__ double exp_main(FastMath &m, double x) {
  const unsigned abstop = (unsigned)(bits_of(x) >> 52) & 0x7ffu;
  m.flag |= (unsigned)(abstop - 0x3c9u >= 0x408u - 0x3c9u);
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

// One step of one agent by one full warp. Returns true when the step was taken (p, v, min_obs updated);
// false: nothing changed, run the general step. pr = this step's prologue (norms, unit vectors, broad phase).
template <bool STATIC_VEL>
__device__ __forceinline__ bool fast_step(const Group<32> &g, const StepEnv &P, const SmemObstacles &obs,
                                          const uint16_t *cand, double *fbuf, const KnownBits &known, int type,
                                          const AgentConsts &c, const FastConsts &fc, v3 init_pos,
                                          const double *rot_row, v3 goal_vec, const Prologue &pr, v3 &p, v3 &v,
                                          double &min_obs) {
  const StepNorms &sn = pr.sn;
  const int n_cand = pr.n_cand;
  if (!fc.usable | (n_cand < 0) | (n_cand > 32)) return false;
  bool rare = false;
  // gate (:315-317)
  const v3 d0 = sub3(p, init_pos);
  const double z0 = dot3(d0, d0);
  const bool near_start = z0 < fc.thr_start.lo;
  rare |= !near_start & !(z0 > fc.thr_start.hi);
  const bool gate_open = !(sn.dist_goal < c.approach_dist) & !((sn.vn < c.half_vmax) & near_start);
  // attractorForceScaling's early exit (:215-218) depends on the agent only
  const bool kgs_zero = (dot3(goal_vec, v) <= 0.0) & (sn.vn < c.vmax90) & (sn.dist_goal > 0.15);

  v3 force = mk3(0.0, 0.0, 0.0);
  double min_d = (double)INFINITY, kgs_closest = 1.0;
  bool has_closest = false;
  if (gate_open & (n_cand > 0)) {
    g.sync();  // cand[] was written by the prologue's broad phase
    // ---- narrow phase: one lane per candidate, idle lanes shadow candidate 0 ----
    const bool active = g.gl < n_cand;
    const int i = (int)cand[active ? g.gl : 0];
    const v3 oi = obs.pos(i);
    const double rs = obs.rsum(i);
    const bool is_known = known.test(i);
    const bool uses_rot = (type != GOAL_HEURISTIC) & (type != VEL_HEURISTIC);
    v3 rot_i = mk3(0.0, 0.0, 1.0);
    if (uses_rot & is_known) rot_i = ld3(rot_row + 3 * i);
    const v3 rov = sub3(oi, p);
    const v3 rel = STATIC_VEL ? v : sub3(v, obs.vel(i));
    FastMath fa, fb;
    const double z = dot3(rov, rov);
    double n, yn;
    fa.sqrt_rcp_(z, n, yn);
    const v3 to_obs = fa.quot3_(rov, n, yn);  // z > 0 here (range check)
    const double d = clamp_dist(n - rs);
    const bool skip = (dot3(to_obs, pr.ghat) < -0.01) & (dot3(rov, rel) < -0.01);  // :79-82
    const bool counts = active & !skip;             // :86-88
    const bool close = active & (d < c.shell);      // closest-obstacle search ignores the skip test (:201-211)
    const bool in_shell = close & !skip;            // :91
    const bool first_seen = in_shell & !is_known;   // :92-96 -> general step
    // attractorForceScaling's tail for THIS obstacle (:219-226), used if it turns out to be the closest
    const double sd = fb.sqrt_(d);
    const double w1 = 1 - exp_main(fb, fb.quot_(-sd, c.shell, fc.y_shell));
    double w2 = 1 - fb.div_(dot3(goal_vec, rov), sn.dist_goal * n);
    w2 = w2 * w2;
    const double kgs = kgs_zero ? 0.0 : w1 * w2;
    // currentVector (:389-406, :520-537, others) and the force term (:98-104)
    double zr, vel_norm;
    v3 nv;
    if (STATIC_VEL) {
      zr = sn.zv, vel_norm = sn.vn, nv = pr.nv_static;
    } else {
      double yv;
      zr = dot3(rel, rel);
      fb.sqrt_rcp_(zr, vel_norm, yv);
      nv = fb.quot3_(rel, vel_norm, yv);
    }
    const v3 nv_eigen = zr > 0.0 ? nv : rel;
    v3 cin = cross3(to_obs, rot_i);
    if (type == GOAL_HEURISTIC) cin = sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec)));
    if (type == VEL_HEURISTIC) cin = sub3(nv_eigen, mul3(to_obs, dot3(nv_eigen, to_obs)));
    double nc, yc;
    fb.sqrt_rcp_(dot3(cin, cin), nc, yc);
    v3 current = fb.quot3_(cin, nc, yc);
    if (!uses_rot & (nc < 1e-10)) current = mk3(0.0, 0.0, 1.0);
    const v3 f = mul3(cross3(nv, cross3(current, nv)), fb.div_(c.k_circ, d * d));
    const bool contributes = in_shell & (vel_norm != 0);
    const bool lane_rare = active & (fa.bad() | first_seen | (close & fb.bad()));

    // ---- force_ += curr_force in obstacle order (:106): staged by rank, zero-padded, summed front to back ----
    const unsigned lt_mask = (1u << g.lane) - 1u;
    const unsigned contrib = g.ballot(contributes);
    const int n_contrib = __popc(contrib);
    if (contributes) st3(fbuf + 3 * __popc(contrib & lt_mask), f);
    if (g.gl < kFastSumUnroll) st3(fbuf + 3 * (n_contrib + g.gl), mk3(0.0, 0.0, 0.0));
    // reductions (exact: minima of non-negative doubles)
    min_d = g.min_reduce_nonneg(counts ? d : (double)INFINITY);
    const double mc = g.min_reduce_nonneg(close ? d : (double)INFINITY);
    const unsigned who = g.ballot(close & (d == mc));  // lowest lane = lowest obstacle index
    has_closest = who != 0u;
    kgs_closest = g.bcast(kgs, (__ffs(who) - 1) & 31);
    rare |= g.ballot(lane_rare) != 0u;
    g.sync();
#pragma unroll
    for (int j = 0; j < kFastSumUnroll; ++j) force = add3(force, ld3(fbuf + 3 * j));
    for (int j = kFastSumUnroll; j < n_contrib; j += 4) {
      const v3 a0 = ld3(fbuf + 3 * j), a1 = ld3(fbuf + 3 * j + 3), a2 = ld3(fbuf + 3 * j + 6), a3 = ld3(fbuf + 3 * j + 9);
      force = add3(add3(add3(add3(force, a0), a1), a2), a3);
    }
    g.sync();
  }

  // ---- scalar rest of the step (replicated in every lane) ----
  const double new_min_obs = min_d < min_obs ? min_d : min_obs;
  const double fz = dot3(force, force);
  const bool big = fz > fc.thr_force.hi;  // |F| > 1e-5 (:319)
  rare |= !big & !(fz < fc.thr_force.lo);
  const double k_goal_scale = (has_closest & big) ? kgs_closest : 1.0;
  // repelForce (:159-181): the sentinel must be out of its shell, then the term is +0
  const v3 dvs = sub3(p, obs.pos(P.n_obs - 1));
  rare |= !(dot3(dvs, dvs) > c.repel_far2);
  force = add3(force, mk3(0.0, 0.0, 0.0));
  // attractorForce (:183-193)
  const v3 fa3 = add3(force, mul3(sub3(sn.vel_des, v), k_goal_scale * c.k_damp));
  if (c.k_attr != 0.0) force = fa3;
  // updatePositionAndVelocity (:253-268), unit mass, no acceleration clamp
  rare |= !(dot3(force, force) < fc.thr_acc.lo);
  const double dt = P.pred_dt;
  const v3 np = mk3((p.x + 0.5 * force.x * dt * dt) + v.x * dt, (p.y + 0.5 * force.y * dt * dt) + v.y * dt,
                    (p.z + 0.5 * force.z * dt * dt) + v.z * dt);
  v3 nvel = add3(v, mul3(force, dt));
  FastMath fm;
  double vel_norm, yvn;
  fm.sqrt_rcp_(dot3(nvel, nvel), vel_norm, yvn);
  const double scale = fm.quot_(c.vel_max, vel_norm, yvn);
  if (vel_norm > c.vel_max) nvel = mul3(nvel, scale);
  rare |= fm.bad();
  if (__builtin_expect(rare, 0)) return false;
  p = np, v = nvel, min_obs = new_min_obs;
  return true;
}
#endif  // __CUDACC__

}  // namespace pmaf
