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
// fast_step evaluates the COMMON case of a step (cf_agent.cpp:312-326) as a few large basic blocks with
// selects instead of branches and no memory side effects until its end:
//   * one entry branch: the step is on, the gate is open, 1..32 broad-phase candidates, unit mass
//     (a closed gate takes the general step, which is short then);
//   * first detections latch in place (RANDOM / GOAL / VEL by selects; HAD and the obstacle heuristics in
//     a cold, group-uniform block), the acceleration clamp is a uniform branch;
//   * everything else uncommon raises `rare`: an operand outside FastMath's proven range, a norm within
//     1e-15 of a threshold, the sentinel within its shell, exp() outside its main path.
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
  int tbits;          // kT* bits of the agent's type (opaque to the compiler: no re-derivation from `type`)
};
constexpr int kTUsesRot = 1, kTRandom = 2, kTGoal = 4, kTVel = 8, kTNeedsNN = 16;

#if defined(__CUDACC__)
__device__ __forceinline__ FastConsts make_fast_consts(const AgentConsts &c, int type, unsigned rz) {
  FastConsts f;
  int tb = 0;
  if (type != GOAL_HEURISTIC && type != VEL_HEURISTIC) tb |= kTUsesRot;
  if (type == RANDOM_AGENT) tb |= kTRandom;
  if (type == GOAL_HEURISTIC) tb |= kTGoal;
  if (type == VEL_HEURISTIC) tb |= kTVel;
  if (type == OBSTACLE_HEURISTIC || type == GOAL_OBSTACLE_HEURISTIC) tb |= kTNeedsNN;
  f.tbits = keep(tb, rz);
  FastMath m;
  f.y_shell = keep(m.rcp_(c.shell), rz);
  f.thr_force = make_thr(1e-5), f.thr_acc = make_thr(13.0), f.thr_start = make_thr(0.2);
  f.thr_force.lo = keep(f.thr_force.lo, rz), f.thr_force.hi = keep(f.thr_force.hi, rz);
  f.thr_acc.lo = keep(f.thr_acc.lo, rz), f.thr_acc.hi = keep(f.thr_acc.hi, rz);
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

// path point store by lane 0 as predicated instructions (the compiler makes `if (lane == 0 && pending)` a
// branch, which would end the basic block the store shares with the next step's prologue)
__device__ __forceinline__ void st3_if(double *ptr, v3 a, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %4, 0;\n"
      "@q st.global.f64 [%0], %1;\n"
      "@q st.global.f64 [%0+8], %2;\n"
      "@q st.global.f64 [%0+16], %3;\n"
      "}\n" ::"l"(ptr),
      "d"(a.x), "d"(a.y), "d"(a.z), "r"((unsigned)pred)
      : "memory");
}

// One step of one agent by one full warp. Returns true when the step was taken (p, v, min_obs updated);
// false: nothing changed, run the general step. pr = this step's prologue (norms, unit vectors, broad phase).
template <bool STATIC_VEL>
__device__ __forceinline__ bool fast_step(const Group<32> &g, const StepEnv &P, const SmemObstacles &obs,
                                          const uint16_t *cand, double *fbuf, const KnownBits &known, int type,
                                          const AgentConsts &c, const FastConsts &fc, v3 init_pos,
                                          double *rot_row, const double *random_row, v3 goal_vec, const Prologue &pr,
                                          v3 &p, v3 &v, double &min_obs, unsigned *why = nullptr, bool step_on = true,
                                          const uint16_t *nn_table = nullptr) {
  // why (developer statistics, PMAF_FAST_STATS builds): bit mask of the reasons a step was not taken
#define PMAF_RARE(bit, cond)                   \
  do {                                         \
    const bool c_ = (cond);                    \
    rare |= c_;                                \
    if (why) *why |= c_ ? (1u << (bit)) : 0u;  \
  } while (0)
  const StepNorms &sn = pr.sn;
  const int n_cand = pr.n_cand;
  bool rare = false;
  // gate (:315-317)
  const v3 d0 = sub3(p, init_pos);
  const double z0 = dot3(d0, d0);
  const bool near_start = z0 < fc.thr_start.lo;
  const bool start_ambiguous = !near_start & !(z0 > fc.thr_start.hi);
  const bool gate_open = !(sn.dist_goal < c.approach_dist) & !((sn.vn < c.half_vmax) & near_start);
  // ONE branch decides between this path and the general step: a closed gate (no field pass: the general
  // step is as short), no or too many candidates, an unusable configuration
  if (!(step_on & fc.usable & (n_cand > 0) & (n_cand <= 32) & gate_open & !start_ambiguous)) {
    if (why) *why |= 1u;
    return false;
  }
  // repelForce (:159-181): the sentinel must be out of its shell, then its term is +0
  const v3 dvs = sub3(p, obs.pos(P.n_obs - 1));
  PMAF_RARE(6, !(dot3(dvs, dvs) > c.repel_far2));
  // attractorForceScaling's early exit (:215-218) depends on the agent only
  const bool kgs_zero = (dot3(goal_vec, v) <= 0.0) & (sn.vn < c.vmax90) & (sn.dist_goal > 0.15);

  v3 force = mk3(0.0, 0.0, 0.0);
  double min_d = (double)INFINITY, kgs_closest = 1.0;
  bool has_closest = false, latch = false, latch_mine = false;
  int latch_i = 0;
  v3 latch_rot = force;
  {
    g.sync();  // cand[] was written by the prologue's broad phase
    // ---- narrow phase: one lane per candidate, idle lanes shadow candidate 0 ----
    const bool active = g.gl < n_cand;
    const int i = (int)cand[active ? g.gl : 0];
    const v3 oi = obs.pos(i);
    const double rs = obs.rsum(i);
    const bool is_known = known.test(i);
    const bool uses_rot = (fc.tbits & kTUsesRot) != 0;
    v3 rot_i = mk3(0.0, 0.0, 1.0);  // GOAL :408-412, VEL :539-543; default of a skipped unknown obstacle
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
    const bool first_seen = in_shell & !is_known;   // :92-96
    // first detection latches the rotation vector (calculateRotationVector, :408-611). RANDOM :559-566 is one
    // cross product with the unit goal vector the prologue already holds (bit-identical to rot_random);
    // GOAL / VEL latch (0,0,1); HAD and the two obstacle heuristics (three agents of a population) take a
    // cold, group-uniform branch: cooperative nearest-neighbour scans (:434-446), then the out-of-line
    // IEEE evaluation, exactly as the general step does (eval_candidate).
    if (first_seen & ((fc.tbits & kTRandom) != 0)) rot_i = cross3(pr.ghat, ld3(random_row + 3 * i));
    bool latch_flag = false;
    if ((fc.tbits & (kTUsesRot | kTRandom)) == kTUsesRot) {
      unsigned todo = g.ballot(first_seen);
      if (__builtin_expect(todo != 0u, 0)) {
#if defined(PMAF_FAST_STATS)
        const long long tl0 = clock64();
#endif
        // cold (a few times per rollout of three agents), written inline with the branch-free arithmetic: an
        // out-of-line call here costs more in register saves than the work itself
        FastMath fscan, frot;
        int nn = 0;
        if ((fc.tbits & kTNeedsNN) && nn_table) {  // static scene: per-tick table built by reset_kernel
          nn = nn_table[i];
        } else if (fc.tbits & kTNeedsNN) {  // nearest other field obstacle, serial-scan semantics (:434-446)
          const int n_field = P.n_obs - 1;
          while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int id = g.bcast(i, src);
            const v3 o_id = obs.pos(id);
            double best = 100.0;
            int best_i = 0x7fffffff;
            for (int k = g.gl; k < n_field; k += 32) {
              const v3 dk = sub3(o_id, obs.pos(k));
              const double dist = fscan.sqrt_(k != id ? dot3(dk, dk) : 1.0);
              if ((k != id) & (best > dist)) best = dist, best_i = k;
            }
            g.argmin_reduce_nonneg(best, best_i);
            if (g.lane == src) nn = best_i == 0x7fffffff ? 0 : best_i;
          }
        }
        v3 r;
        if (type == HAD_HEURISTIC) {  // rot_had (:599-611)
          const double sh = frot.div_(dot3(rov, goal_vec), sn.dist_goal * sn.dist_goal);
          const v3 dh = sub3(add3(p, mul3(goal_vec, sh)), oi);
          const v3 ch = cross3(dh, goal_vec);
          double nh, yh;
          frot.sqrt_rcp_(dot3(ch, ch), nh, yh);
          r = frot.quot3_(ch, nh, yh);
        } else {  // rot_obstacle (:447-460), rot_goal_obstacle (:493-517)
          const v3 obstacle_vec = sub3(obs.pos(nn), oi);
          const v3 obst_current = sub3(mul3(to_obs, dot3(obstacle_vec, to_obs)), obstacle_vec);
          v3 cur = obst_current;
          if (type == GOAL_OBSTACLE_HEURISTIC) {
            const v3 goal_current = sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec)));
            double n1, y1, n2, y2, n3, y3;
            frot.sqrt_rcp_(dot3(goal_current, goal_current), n1, y1);
            frot.sqrt_rcp_(dot3(obst_current, obst_current), n2, y2);
            cur = add3(frot.quot3_(goal_current, n1, y1), frot.quot3_(obst_current, n2, y2));
            FastMath fsum;  // a sum below 1e-10 is replaced, whatever its square root did
            fsum.sqrt_rcp_(dot3(cur, cur), n3, y3);
            const v3 q3 = fsum.quot3_(cur, n3, y3);
            const bool tiny = n3 < 1e-10;
            frot.flag |= (fsum.bad() && !(dot3(cur, cur) == 0.0)) ? 1u : 0u;
            cur = (tiny | (dot3(cur, cur) == 0.0)) ? mk3(0.0, 0.0, 1.0) : q3;
          }
          const v3 cr = cross3(cur, to_obs);
          double nr, yr;
          frot.sqrt_rcp_(dot3(cr, cr), nr, yr);
          r = frot.quot3_(cr, nr, yr);
        }
        if (first_seen) rot_i = r;
        latch_flag = fscan.bad() | (first_seen & frot.bad());
#if defined(PMAF_FAST_STATS)
        if (why) why[1] += (unsigned)(clock64() - tl0), why[2] += 1u;
#endif
      }
    }
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
    if (fc.tbits & kTGoal) cin = sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec)));
    if (fc.tbits & kTVel) cin = sub3(nv_eigen, mul3(to_obs, dot3(nv_eigen, to_obs)));
    double nc, yc;
    fb.sqrt_rcp_(dot3(cin, cin), nc, yc);
    v3 current = fb.quot3_(cin, nc, yc);
    if (!uses_rot & (nc < 1e-10)) current = mk3(0.0, 0.0, 1.0);
    const v3 f = mul3(cross3(nv, cross3(current, nv)), fb.div_(c.k_circ, d * d));
    const bool contributes = in_shell & (vel_norm != 0);
    const bool lane_rare = latch_flag | (active & (fa.bad() | (close & fb.bad())));
    if (why) {
      if (g.ballot(active & fa.bad())) *why |= 1u << 2;
      if (g.ballot(active & close & fb.bad())) *why |= 1u << 4;
    }

    // ---- force_ += curr_force in obstacle order (:106): staged by rank, zero-padded, summed front to back ----
    const unsigned lt_mask = (1u << g.lane) - 1u;
    const unsigned contrib = g.ballot(contributes);
    const int n_contrib = __popc(contrib);
    // staging as three component arrays of kFbufSlots doubles (x | y | z): two slots per 16-byte shared load
    constexpr int kFbufSlots = 32 + kFastSumUnroll;
    double *fx = fbuf, *fy = fbuf + kFbufSlots, *fz = fbuf + 2 * kFbufSlots;
    if (contributes) {
      const int rk = __popc(contrib & lt_mask);
      fx[rk] = f.x, fy[rk] = f.y, fz[rk] = f.z;
    }
    if (g.gl < kFastSumUnroll) fx[n_contrib + g.gl] = 0.0, fy[n_contrib + g.gl] = 0.0, fz[n_contrib + g.gl] = 0.0;
    // reductions (exact: minima of non-negative doubles)
    min_d = g.min_reduce_nonneg(counts ? d : (double)INFINITY);
    const double mc = g.min_reduce_nonneg(close ? d : (double)INFINITY);
    const unsigned who = g.ballot(close & (d == mc));  // lowest lane = lowest obstacle index
    has_closest = who != 0u;
    kgs_closest = g.bcast(kgs, (__ffs(who) - 1) & 31);
    rare |= g.ballot(lane_rare) != 0u;
    latch = g.ballot(first_seen) != 0u;
    latch_mine = first_seen, latch_i = i, latch_rot = rot_i;
    g.sync();
    const double2 *fx2 = reinterpret_cast<const double2 *>(fx), *fy2 = reinterpret_cast<const double2 *>(fy),
                  *fz2 = reinterpret_cast<const double2 *>(fz);
#pragma unroll
    for (int j = 0; j < kFastSumUnroll / 2; ++j) {
      const double2 x = fx2[j], y = fy2[j], z = fz2[j];
      force = add3(add3(force, mk3(x.x, y.x, z.x)), mk3(x.y, y.y, z.y));
    }
    if (n_contrib > kFastSumUnroll) {  // second block, zero-padded as well
#pragma unroll
      for (int j = kFastSumUnroll / 2; j < kFastSumUnroll; ++j) {
        const double2 x = fx2[j], y = fy2[j], z = fz2[j];
        force = add3(add3(force, mk3(x.x, y.x, z.x)), mk3(x.y, y.y, z.y));
      }
      for (int j = kFastSumUnroll; 2 * j < n_contrib; j += 2) {
        const double2 x0 = fx2[j], y0 = fy2[j], z0 = fz2[j], x1 = fx2[j + 1], y1 = fy2[j + 1], z1 = fz2[j + 1];
        force = add3(add3(force, mk3(x0.x, y0.x, z0.x)), mk3(x0.y, y0.y, z0.y));
        force = add3(add3(force, mk3(x1.x, y1.x, z1.x)), mk3(x1.y, y1.y, z1.y));
      }
    }
    g.sync();
  }

  // ---- scalar rest of the step (replicated in every lane) ----
  const double new_min_obs = min_d < min_obs ? min_d : min_obs;
  const double fz = dot3(force, force);
  const bool big = fz > fc.thr_force.hi;  // |F| > 1e-5 (:319)
  PMAF_RARE(5, !big & !(fz < fc.thr_force.lo));
  const double k_goal_scale = (has_closest & big) ? kgs_closest : 1.0;
  force = add3(force, mk3(0.0, 0.0, 0.0));  // += total_repel_force (:179), zero here
  // attractorForce (:183-193)
  const v3 fa3 = add3(force, mul3(sub3(sn.vel_des, v), k_goal_scale * c.k_damp));
  if (c.k_attr != 0.0) force = fa3;
  // updatePositionAndVelocity (:253-268), unit mass; the acceleration clamp (|a| > 13) is a uniform branch
  {
    const double zacc = dot3(force, force);
    const bool clamp = zacc > fc.thr_acc.hi;
    PMAF_RARE(7, !clamp & !(zacc < fc.thr_acc.lo));
    if (clamp) {
      FastMath fc2;
      double na, ya;
      fc2.sqrt_rcp_(zacc, na, ya);
      force = mul3(force, fc2.quot_(13.0, na, ya));
      PMAF_RARE(8, fc2.bad());
    }
  }
  const double dt = P.pred_dt;
  const v3 np = mk3((p.x + 0.5 * force.x * dt * dt) + v.x * dt, (p.y + 0.5 * force.y * dt * dt) + v.y * dt,
                    (p.z + 0.5 * force.z * dt * dt) + v.z * dt);
  v3 nvel = add3(v, mul3(force, dt));
  FastMath fm;
  double vel_norm, yvn;
  fm.sqrt_rcp_(dot3(nvel, nvel), vel_norm, yvn);
  const double scale = fm.quot_(c.vel_max, vel_norm, yvn);
  if (vel_norm > c.vel_max) nvel = mul3(nvel, scale);
  PMAF_RARE(8, fm.bad());
  if (__builtin_expect(rare, 0)) return false;
  if (__builtin_expect(latch, 0)) {  // commit the first detections of this step (:93-95)
    if (latch_mine) {
      st3(rot_row + 3 * latch_i, latch_rot);
      known.set(latch_i);
    }
    g.sync();
  }
  p = np, v = nvel, min_obs = new_min_obs;
  return true;
#undef PMAF_RARE
}

// The same step for the THROUGHPUT shapes: obstacle sets whose candidates do not fit one lane each (more than
// 64 field obstacles: the broad phase is a loop) and groups narrower than a warp (G = Group<16> / Group<8>:
// two / four agents share a warp's instruction stream, so the scalar part of the step, the ordered force sum
// and the reductions are issued once for 2 / 4 agents instead of once per agent, and a candidate list of n
// entries occupies ceil(n / LPA) * LPA lane slots instead of ceil(n / 32) * 32). The narrow phase runs in
// chunks of G::kLanes candidates. Per chunk the straight-line narrow phase of fast_step; contributions are
// summed chunk after chunk (candidate order = obstacle order), the per-lane running minima are reduced once
// after the last chunk, and attractorForceScaling's tail is evaluated once for the winner (these populations
// run several warps per scheduler: fewer instructions beat a shorter dependency chain). First detections are
// committed chunk by chunk: a later rare event re-runs the general step, which then finds the obstacle known
// with exactly the rotation vector it would have latched itself. All collectives are the group's (sub-warp
// masks), all branches are group-uniform; groups of one warp may diverge from each other.
template <bool STATIC_VEL, class G>
__device__ __forceinline__ bool fast_step_multi(const G &g, const StepEnv &P, const SmemObstacles &obs,
                                                const uint16_t *cand, double *fbuf, const KnownBits &known, int type,
                                                const AgentConsts &c, const FastConsts &fc, v3 init_pos, double *rot_row,
                                                const double *random_row, v3 goal_vec, const Prologue &pr, v3 &p,
                                                v3 &v, double &min_obs, bool step_on, const uint16_t *nn_table) {
  const StepNorms &sn = pr.sn;
  const int n_cand = pr.n_cand;
  bool rare = false;
  const v3 d0 = sub3(p, init_pos);
  const double z0 = dot3(d0, d0);
  const bool near_start = z0 < fc.thr_start.lo;
  const bool start_ambiguous = !near_start & !(z0 > fc.thr_start.hi);
  const bool gate_open = !(sn.dist_goal < c.approach_dist) & !((sn.vn < c.half_vmax) & near_start);
  if (!(step_on & fc.usable & (n_cand > 0) & gate_open & !start_ambiguous)) return false;
  const v3 dvs = sub3(p, obs.pos(P.n_obs - 1));
  rare |= !(dot3(dvs, dvs) > c.repel_far2);

  v3 force = mk3(0.0, 0.0, 0.0);
  double lmin = (double)INFINITY, lcd = (double)INFINITY;  // per-lane running minima over the chunks
  int lci = 0x7fffffff;                                    // candidate position of the lane's closest obstacle
  constexpr int LPA = G::kLanes;
  static_assert(LPA >= kFastSumUnroll, "the zero padding of the staging buffer is written by the group's first lanes");
  const unsigned lt_mask = g.mask & ((1u << g.lane) - 1u);
  const bool uses_rot = (fc.tbits & kTUsesRot) != 0;
  constexpr int kFbufSlots = LPA + kFastSumUnroll;
  double *fx = fbuf, *fy = fbuf + kFbufSlots, *fz = fbuf + 2 * kFbufSlots;
  const double2 *fx2 = reinterpret_cast<const double2 *>(fx), *fy2 = reinterpret_cast<const double2 *>(fy),
                *fz2 = reinterpret_cast<const double2 *>(fz);
  g.sync();  // cand[] was written by the broad phase
  for (int c0 = 0; c0 < n_cand; c0 += LPA) {
    const int ci = c0 + g.gl;
    const bool active = ci < n_cand;
    const int i = (int)cand[active ? ci : 0];
    const v3 oi = obs.pos(i);
    const double rs = obs.rsum(i);
    const bool is_known = known.test(i);
    v3 rot_i = mk3(0.0, 0.0, 1.0);
    if (uses_rot & is_known) rot_i = ld3(rot_row + 3 * i);
    const v3 rov = sub3(oi, p);
    const v3 rel = STATIC_VEL ? v : sub3(v, obs.vel(i));
    FastMath fa, fb;
    const double z = dot3(rov, rov);
    double n, yn;
    fa.sqrt_rcp_(z, n, yn);
    const v3 to_obs = fa.quot3_(rov, n, yn);
    const double d = clamp_dist(n - rs);
    const bool skip = (dot3(to_obs, pr.ghat) < -0.01) & (dot3(rov, rel) < -0.01);  // :79-82
    const bool counts = active & !skip;
    const bool close = active & (d < c.shell);
    const bool in_shell = close & !skip;
    const bool first_seen = in_shell & !is_known;
    bool latch_flag = false;
    if (first_seen & ((fc.tbits & kTRandom) != 0)) rot_i = cross3(pr.ghat, ld3(random_row + 3 * i));
    if ((fc.tbits & (kTUsesRot | kTRandom)) == kTUsesRot) {
      unsigned todo = g.ballot(first_seen);
      if (__builtin_expect(todo != 0u, 0)) {  // cold: see fast_step
        FastMath fscan, frot;
        int nn = 0;
        if ((fc.tbits & kTNeedsNN) && nn_table) {
          nn = nn_table[i];
        } else if (fc.tbits & kTNeedsNN) {
          const int n_field = P.n_obs - 1;
          while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int id = g.bcast(i, src);
            const v3 o_id = obs.pos(id);
            double best = 100.0;
            int best_i = 0x7fffffff;
            for (int k = g.gl; k < n_field; k += LPA) {
              const v3 dk = sub3(o_id, obs.pos(k));
              const double dist = fscan.sqrt_(k != id ? dot3(dk, dk) : 1.0);
              if ((k != id) & (best > dist)) best = dist, best_i = k;
            }
            g.argmin_reduce_nonneg(best, best_i);
            if (g.lane == src) nn = best_i == 0x7fffffff ? 0 : best_i;
          }
        }
        v3 r;
        if (type == HAD_HEURISTIC) {
          const double sh = frot.div_(dot3(rov, goal_vec), sn.dist_goal * sn.dist_goal);
          const v3 dh = sub3(add3(p, mul3(goal_vec, sh)), oi);
          const v3 ch = cross3(dh, goal_vec);
          double nh, yh;
          frot.sqrt_rcp_(dot3(ch, ch), nh, yh);
          r = frot.quot3_(ch, nh, yh);
        } else {
          const v3 obstacle_vec = sub3(obs.pos(nn), oi);
          const v3 obst_current = sub3(mul3(to_obs, dot3(obstacle_vec, to_obs)), obstacle_vec);
          v3 cur = obst_current;
          if (type == GOAL_OBSTACLE_HEURISTIC) {
            const v3 goal_current = sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec)));
            double n1, y1, n2, y2, n3, y3;
            frot.sqrt_rcp_(dot3(goal_current, goal_current), n1, y1);
            frot.sqrt_rcp_(dot3(obst_current, obst_current), n2, y2);
            cur = add3(frot.quot3_(goal_current, n1, y1), frot.quot3_(obst_current, n2, y2));
            FastMath fsum;
            fsum.sqrt_rcp_(dot3(cur, cur), n3, y3);
            const v3 q3 = fsum.quot3_(cur, n3, y3);
            const bool tiny = n3 < 1e-10;
            frot.flag |= (fsum.bad() && !(dot3(cur, cur) == 0.0)) ? 1u : 0u;
            cur = (tiny | (dot3(cur, cur) == 0.0)) ? mk3(0.0, 0.0, 1.0) : q3;
          }
          const v3 cr = cross3(cur, to_obs);
          double nr, yr;
          frot.sqrt_rcp_(dot3(cr, cr), nr, yr);
          r = frot.quot3_(cr, nr, yr);
        }
        if (first_seen) rot_i = r;
        latch_flag = fscan.bad() | (first_seen & frot.bad());
      }
    }
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
    if (fc.tbits & kTGoal) cin = sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec)));
    if (fc.tbits & kTVel) cin = sub3(nv_eigen, mul3(to_obs, dot3(nv_eigen, to_obs)));
    double nc, yc;
    fb.sqrt_rcp_(dot3(cin, cin), nc, yc);
    v3 current = fb.quot3_(cin, nc, yc);
    if (!uses_rot & (nc < 1e-10)) current = mk3(0.0, 0.0, 1.0);
    const v3 f = mul3(cross3(nv, cross3(current, nv)), fb.div_(c.k_circ, d * d));
    const bool contributes = in_shell & (vel_norm != 0);
    const bool lane_bad = latch_flag | (active & (fa.bad() | (close & fb.bad())));
    // running minima (strict <: the first of equals within a lane has the lower obstacle index)
    if (counts & (d < lmin)) lmin = d;
    if (close & (d < lcd)) lcd = d, lci = ci;
    // ordered sum of this chunk's contributions
    const unsigned contrib = g.ballot(contributes);
    const int n_contrib = __popc(contrib);
    if (contributes) {
      const int rk = __popc(contrib & lt_mask);
      fx[rk] = f.x, fy[rk] = f.y, fz[rk] = f.z;
    }
    if (g.gl < kFastSumUnroll) fx[n_contrib + g.gl] = 0.0, fy[n_contrib + g.gl] = 0.0, fz[n_contrib + g.gl] = 0.0;
    const bool any_bad = g.ballot(lane_bad) != 0u;
    rare |= any_bad;
    // first detections of this chunk (:93-95); not while an operand was out of range
    if (__builtin_expect(g.ballot(first_seen) != 0u, 0)) {
      if (first_seen & !any_bad) {
        st3(rot_row + 3 * i, rot_i);
        known.set(i);
      }
    }
    g.sync();
#pragma unroll
    for (int j = 0; j < kFastSumUnroll / 2; ++j) {
      const double2 x = fx2[j], y = fy2[j], z2 = fz2[j];
      force = add3(add3(force, mk3(x.x, y.x, z2.x)), mk3(x.y, y.y, z2.y));
    }
    if (LPA > kFastSumUnroll && n_contrib > kFastSumUnroll) {  // narrow groups: a chunk never holds more
      for (int j = kFastSumUnroll / 2; 2 * j < n_contrib; j += 2) {
        const double2 x0 = fx2[j], y0 = fy2[j], z0 = fz2[j], x1 = fx2[j + 1], y1 = fy2[j + 1], z1 = fz2[j + 1];
        force = add3(add3(force, mk3(x0.x, y0.x, z0.x)), mk3(x0.y, y0.y, z0.y));
        force = add3(add3(force, mk3(x1.x, y1.x, z1.x)), mk3(x1.y, y1.y, z1.y));
      }
    }
    g.sync();
  }
  // reductions over the lanes' running minima
  const double min_d = g.min_reduce_nonneg(lmin);
  const bool has_closest = g.ballot(lci != 0x7fffffff) != 0u;
  double kgs_closest = 1.0;
  FastMath fm;
  if (has_closest) {  // attractorForceScaling's tail (:212-226) once, for the first obstacle with the smallest distance
    g.argmin_reduce_nonneg(lcd, lci);
    const v3 o_c = obs.pos((int)cand[lci]);
    kgs_closest = attractor_scaling(fm, goal_vec, sn.dist_goal, p, v, sn.vn, c, lcd, o_c);
  }
  // ---- scalar rest of the step (as in fast_step) ----
  const double new_min_obs = min_d < min_obs ? min_d : min_obs;
  const double fzz = dot3(force, force);
  const bool big = fzz > fc.thr_force.hi;
  rare |= !big & !(fzz < fc.thr_force.lo);
  const double k_goal_scale = (has_closest & big) ? kgs_closest : 1.0;
  force = add3(force, mk3(0.0, 0.0, 0.0));
  const v3 fa3 = add3(force, mul3(sub3(sn.vel_des, v), k_goal_scale * c.k_damp));
  if (c.k_attr != 0.0) force = fa3;
  {
    const double zacc = dot3(force, force);
    const bool clamp = zacc > fc.thr_acc.hi;
    rare |= !clamp & !(zacc < fc.thr_acc.lo);
    if (clamp) {
      double na, ya;
      fm.sqrt_rcp_(zacc, na, ya);
      force = mul3(force, fm.quot_(13.0, na, ya));
    }
  }
  const double dt = P.pred_dt;
  const v3 np = mk3((p.x + 0.5 * force.x * dt * dt) + v.x * dt, (p.y + 0.5 * force.y * dt * dt) + v.y * dt,
                    (p.z + 0.5 * force.z * dt * dt) + v.z * dt);
  v3 nvel = add3(v, mul3(force, dt));
  double vel_norm, yvn;
  fm.sqrt_rcp_(dot3(nvel, nvel), vel_norm, yvn);
  const double scale = fm.quot_(c.vel_max, vel_norm, yvn);
  if (vel_norm > c.vel_max) nvel = mul3(nvel, scale);
  rare |= fm.bad();
  if (__builtin_expect(rare, 0)) return false;
  p = np, v = nvel, min_obs = new_min_obs;
  return true;
}

// ---- packed shapes: 2 / 4 agents per warp -------------------------------------------------------------------------
// fast_step_multi's arithmetic with the control flow of a warp that carries several agents: every lane of the
// WARP executes the same instruction stream (chunk loop up to the largest candidate count among the warp's
// groups, warp-uniform branches around the cold blocks) and all collectives are the warp-convergent `_w`
// variants — a group whose agent has no work in a chunk (fewer candidates, closed gate, finished or absent
// agent) runs predicated off on a valid dummy candidate and commits nothing. What the warp issues once now
// serves 2 / 4 agents: the scalar part of the step, the ordered force sum (its dependent adds are the longest
// serial chain of a step), the reductions and the loop overhead; a candidate list of n entries occupies
// ceil(n / LPA) * LPA lane slots instead of ceil(n / 32) * 32. A closed gate (cf_agent.cpp:315-317) is part of
// this path (zero candidates), so that only termination and the rare events leave it.
// Returns true when the group's step was taken (p, v, min_obs updated); false: nothing of the step's state
// changed (first detections may have been latched, exactly as the general step will find them), the caller
// decides between termination and the general step.
template <bool STATIC_VEL, class G>
__device__ __forceinline__ bool fast_step_packed(const G &g, const StepEnv &P, const SmemObstacles &obs,
                                                 const uint16_t *cand, double *fbuf, const KnownBits &known, int type,
                                                 const AgentConsts &c, const FastConsts &fc, v3 init_pos, double *rot_row,
                                                 const double *random_row, v3 goal_vec, const Prologue &pr, v3 &p,
                                                 v3 &v, double &min_obs, bool step_on, const uint16_t *nn_table) {
  constexpr int LPA = G::kLanes;
  static_assert(LPA >= kFastSumUnroll, "the zero padding of the staging buffer is written by the group's first lanes");
  const StepNorms &sn = pr.sn;
  const v3 d0 = sub3(p, init_pos);
  const double z0 = dot3(d0, d0);
  const bool near_start = z0 < fc.thr_start.lo;
  const bool start_ambiguous = !near_start & !(z0 > fc.thr_start.hi);
  const bool gate_open = !(sn.dist_goal < c.approach_dist) & !((sn.vn < c.half_vmax) & near_start);
  const bool eligible = step_on & fc.usable & !start_ambiguous;
  const int n_cand = (eligible & gate_open) ? pr.n_cand : 0;  // closed gate: no field pass at all (:315-318)
  const int n_max = __reduce_max_sync(0xffffffffu, n_cand);
  const v3 dvs = sub3(p, obs.pos(P.n_obs - 1));
  bool rare = !(dot3(dvs, dvs) > c.repel_far2);  // repelForce (:159-181): the sentinel must be out of its shell

  v3 force = mk3(0.0, 0.0, 0.0);
  double lmin = (double)INFINITY, lcd = (double)INFINITY;  // per-lane running minima over the chunks
  int lci = 0x7fffffff;                                    // candidate position of the lane's closest obstacle
  const unsigned lt_mask = (1u << g.gl) - 1u;
  const bool uses_rot = (fc.tbits & kTUsesRot) != 0;
  constexpr int kFbufSlots = LPA + kFastSumUnroll;
  double *fx = fbuf, *fy = fbuf + kFbufSlots, *fz = fbuf + 2 * kFbufSlots;
  const double2 *fx2 = reinterpret_cast<const double2 *>(fx), *fy2 = reinterpret_cast<const double2 *>(fy),
                *fz2 = reinterpret_cast<const double2 *>(fz);
  g.sync_w();  // cand[] was written by the broad phase
  for (int c0 = 0; c0 < n_max; c0 += LPA) {
    const int ci = c0 + g.gl;
    const bool active = ci < n_cand;
    const int ci_safe = active ? ci : 0;
    const int i_raw = (int)cand[ci_safe];
    const int i = active ? i_raw : 0;  // an idle lane shadows obstacle 0 (cand[0] may never have been written)
    const v3 oi = obs.pos(i);
    const double rs = obs.rsum(i);
    const bool is_known = active & known.test(i);
    v3 rot_i = mk3(0.0, 0.0, 1.0);
    if (uses_rot & is_known) rot_i = ld3(rot_row + 3 * i);
    const v3 rov = sub3(oi, p);
    const v3 rel = STATIC_VEL ? v : sub3(v, obs.vel(i));
    FastMath fa, fb;
    const double z = dot3(rov, rov);
    double n, yn;
    fa.sqrt_rcp_(z, n, yn);
    const v3 to_obs = fa.quot3_(rov, n, yn);
    const double d = clamp_dist(n - rs);
    const bool skip = (dot3(to_obs, pr.ghat) < -0.01) & (dot3(rov, rel) < -0.01);  // :79-82
    const bool counts = active & !skip;
    const bool close = active & (d < c.shell);
    const bool in_shell = close & !skip;
    const bool first_seen = in_shell & !is_known;
    bool latch_flag = false;
    if (first_seen & ((fc.tbits & kTRandom) != 0)) rot_i = cross3(pr.ghat, ld3(random_row + 3 * i));
    // HAD and the two obstacle heuristics (agents 0, 2, 3 of the population): cold, warp-uniform test
    const bool cold_mine = first_seen & ((fc.tbits & (kTUsesRot | kTRandom)) == kTUsesRot);
    if (__builtin_expect(g.any_w(cold_mine), 0)) {
      FastMath fscan, frot;
      int nn = 0;
      const bool need_nn = cold_mine & ((fc.tbits & kTNeedsNN) != 0);
      if (nn_table) {  // static scene: per-tick table built by reset_kernel
        if (need_nn) nn = nn_table[i];
      } else {  // nearest other field obstacle, serial-scan semantics (:434-446); the whole warp scans for one lane
        unsigned todo = __ballot_sync(0xffffffffu, need_nn);
        const int n_field = P.n_obs - 1;
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int id = __shfl_sync(0xffffffffu, i, src);
          const v3 o_id = obs.pos(id);
          double best = 100.0;
          int best_i = 0x7fffffff;
          for (int k = g.lane; k < n_field; k += 32) {
            const v3 dk = sub3(o_id, obs.pos(k));
            const double dist = fscan.sqrt_(k != id ? dot3(dk, dk) : 1.0);
            if ((k != id) & (best > dist)) best = dist, best_i = k;
          }
          {  // lexicographic (distance, index) minimum over the warp: redux on the integer order of non-negative doubles
            const unsigned hi = (unsigned)__double2hiint(best), lo = (unsigned)__double2loint(best);
            const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
            const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
            const bool mine = (hi == mhi) & (lo == mlo);
            best_i = (int)__reduce_min_sync(0xffffffffu, mine ? (unsigned)best_i : 0xffffffffu);
          }
          if (g.lane == src) nn = best_i == 0x7fffffff ? 0 : best_i;
        }
      }
      if (cold_mine) {  // lane-divergent, no collectives inside
        v3 r;
        if (type == HAD_HEURISTIC) {  // rot_had (:599-611)
          const double sh = frot.div_(dot3(rov, goal_vec), sn.dist_goal * sn.dist_goal);
          const v3 dh = sub3(add3(p, mul3(goal_vec, sh)), oi);
          const v3 ch = cross3(dh, goal_vec);
          double nh, yh;
          frot.sqrt_rcp_(dot3(ch, ch), nh, yh);
          r = frot.quot3_(ch, nh, yh);
        } else {  // rot_obstacle (:447-460), rot_goal_obstacle (:493-517)
          const v3 obstacle_vec = sub3(obs.pos(nn), oi);
          const v3 obst_current = sub3(mul3(to_obs, dot3(obstacle_vec, to_obs)), obstacle_vec);
          v3 cur = obst_current;
          if (type == GOAL_OBSTACLE_HEURISTIC) {
            const v3 goal_current = sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec)));
            double n1, y1, n2, y2, n3, y3;
            frot.sqrt_rcp_(dot3(goal_current, goal_current), n1, y1);
            frot.sqrt_rcp_(dot3(obst_current, obst_current), n2, y2);
            cur = add3(frot.quot3_(goal_current, n1, y1), frot.quot3_(obst_current, n2, y2));
            FastMath fsum;  // a sum below 1e-10 is replaced, whatever its square root did
            fsum.sqrt_rcp_(dot3(cur, cur), n3, y3);
            const v3 q3 = fsum.quot3_(cur, n3, y3);
            const bool tiny = n3 < 1e-10;
            frot.flag |= (fsum.bad() && !(dot3(cur, cur) == 0.0)) ? 1u : 0u;
            cur = (tiny | (dot3(cur, cur) == 0.0)) ? mk3(0.0, 0.0, 1.0) : q3;
          }
          const v3 cr = cross3(cur, to_obs);
          double nr, yr;
          frot.sqrt_rcp_(dot3(cr, cr), nr, yr);
          r = frot.quot3_(cr, nr, yr);
        }
        rot_i = r;
        latch_flag = frot.bad();
      }
      latch_flag |= g.any_w(fscan.bad());  // an operand of the shared scan out of range: every group re-runs
    }
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
    if (fc.tbits & kTGoal) cin = sub3(goal_vec, mul3(to_obs, dot3(to_obs, goal_vec)));
    if (fc.tbits & kTVel) cin = sub3(nv_eigen, mul3(to_obs, dot3(nv_eigen, to_obs)));
    double nc, yc;
    fb.sqrt_rcp_(dot3(cin, cin), nc, yc);
    v3 current = fb.quot3_(cin, nc, yc);
    if (!uses_rot & (nc < 1e-10)) current = mk3(0.0, 0.0, 1.0);
    const v3 f = mul3(cross3(nv, cross3(current, nv)), fb.div_(c.k_circ, d * d));
    const bool contributes = in_shell & (vel_norm != 0);
    const bool lane_bad = latch_flag | (active & (fa.bad() | (close & fb.bad())));
    // running minima (strict <: the first of equals within a lane has the lower obstacle index)
    if (counts & (d < lmin)) lmin = d;
    if (close & (d < lcd)) lcd = d, lci = ci;
    // ordered sum of this chunk's contributions (group-relative ballots)
    const unsigned contrib = g.ballot_w(contributes);
    const int n_contrib = __popc(contrib);
    if (contributes) {
      const int rk = __popc(contrib & lt_mask);
      fx[rk] = f.x, fy[rk] = f.y, fz[rk] = f.z;
    }
    if (g.gl < kFastSumUnroll) fx[n_contrib + g.gl] = 0.0, fy[n_contrib + g.gl] = 0.0, fz[n_contrib + g.gl] = 0.0;
    const bool any_bad = g.ballot_w(lane_bad) != 0u;
    rare |= any_bad;
    // first detections of this chunk (:93-95); not while an operand of the group was out of range
    if (__builtin_expect(g.any_w(first_seen), 0)) {
      if (first_seen & !any_bad) {
        st3(rot_row + 3 * i, rot_i);
        known.set(i);
      }
    }
    g.sync_w();
#pragma unroll
    for (int j = 0; j < kFastSumUnroll / 2; ++j) {
      const double2 x = fx2[j], y = fy2[j], z2 = fz2[j];
      force = add3(add3(force, mk3(x.x, y.x, z2.x)), mk3(x.y, y.y, z2.y));
    }
    if (LPA > kFastSumUnroll) {  // narrower groups: a chunk never holds more than the unrolled block
      if (g.any_w(n_contrib > kFastSumUnroll)) {  // zero-padded past n_contrib: groups with fewer add +0.0 (exact)
        const int n_more = __reduce_max_sync(0xffffffffu, n_contrib);
        for (int j = kFastSumUnroll / 2; 2 * j < n_more; j += 2) {
          const bool on = 2 * j < n_contrib;  // beyond the group's own padding the buffer holds stale values
          const double2 x0 = fx2[j], y0 = fy2[j], z0 = fz2[j], x1 = fx2[j + 1], y1 = fy2[j + 1], z1 = fz2[j + 1];
          v3 fn = add3(add3(force, mk3(x0.x, y0.x, z0.x)), mk3(x0.y, y0.y, z0.y));
          fn = add3(add3(fn, mk3(x1.x, y1.x, z1.x)), mk3(x1.y, y1.y, z1.y));
          if (on) force = fn;
        }
      }
    }
    g.sync_w();
  }
  // reductions over the lanes' running minima
  const double min_d = g.min_w_nonneg(lmin);
  const bool has_closest = g.ballot_w(lci != 0x7fffffff) != 0u;
  double kgs_closest = 1.0;
  bool kgs_bad = false;
  if (g.any_w(has_closest)) {  // attractorForceScaling's tail (:212-226) once, for the first obstacle with the smallest distance
    g.argmin_w_nonneg(lcd, lci);
    const int ic_raw = (int)cand[has_closest ? lci : 0];
    const v3 o_c = obs.pos(has_closest ? ic_raw : 0);
    FastMath fs;
    kgs_closest = attractor_scaling(fs, goal_vec, sn.dist_goal, p, v, sn.vn, c, has_closest ? lcd : 1.0, o_c);
    kgs_bad = has_closest & fs.bad();
  }
  rare |= kgs_bad;
  // ---- scalar rest of the step (as in fast_step) ----
  const double new_min_obs = min_d < min_obs ? min_d : min_obs;
  const double fzz = dot3(force, force);
  const bool big = fzz > fc.thr_force.hi;
  rare |= !big & !(fzz < fc.thr_force.lo);
  const double k_goal_scale = (has_closest & big) ? kgs_closest : 1.0;
  force = add3(force, mk3(0.0, 0.0, 0.0));
  const v3 fa3 = add3(force, mul3(sub3(sn.vel_des, v), k_goal_scale * c.k_damp));
  if (c.k_attr != 0.0) force = fa3;
  {
    const double zacc = dot3(force, force);
    const bool clamp = zacc > fc.thr_acc.hi;
    rare |= !clamp & !(zacc < fc.thr_acc.lo);
    if (__builtin_expect(g.any_w(clamp), 0)) {
      FastMath fc2;
      double na, ya;
      fc2.sqrt_rcp_(zacc, na, ya);
      const v3 fcl = mul3(force, fc2.quot_(13.0, na, ya));
      if (clamp) force = fcl;
      rare |= clamp & fc2.bad();
    }
  }
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
  if (rare | !eligible) return false;
  p = np, v = nvel, min_obs = new_min_obs;
  return true;
}
#endif  // __CUDACC__

}  // namespace pmaf
