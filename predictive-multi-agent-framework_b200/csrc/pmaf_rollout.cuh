// pmaf_rollout.cuh — the rollout kernel (CfAgent::cfPrediction, cf_agent.cpp:302-341, for every
// agent at once) and the small per-tick kernels around it.
#pragma once
#include "pmaf_kernels.cuh"

namespace pmaf {

// contributions the straight-line step (pmaf_fast.cuh) sums without a loop: the staging buffer is zero-padded by this many slots
constexpr int kFastSumUnroll = 8;

// Shared memory of a rollout CTA:
//   [0, 16)                      mbarrier
//   [16, 16 + img.bytes)         obstacle image (TMA bulk copy of PlannerDev::image)
//   then per group: double fbuf[fbuf_stride] (ordered force sum staging, 3 * (LPA + 8) used), uint16
//   cand[cand_stride], uint32 known[known_words]
// Per-group strides are padded so that the groups of one warp (2 / 4 agents per warp) fall into different banks
// when they read their own list / staging buffer at the same offset: stride = 4 words (16 B) mod 32 words.
__host__ __device__ inline uint32_t rollout_cand_stride(int n_obs, int lanes_per_agent) {
  if (lanes_per_agent >= 32) return (uint32_t)((n_obs + 7) & ~7);  // one group per warp: nothing to separate
  return (uint32_t)(((n_obs + 63) & ~63) + 8);
}
__host__ __device__ inline uint32_t rollout_fbuf_stride(int lanes_per_agent) {
  const uint32_t n = 3u * (uint32_t)(lanes_per_agent + kFastSumUnroll);  // doubles
  if (lanes_per_agent >= 32) return n;
  return n + ((34u - (n & 15u)) & 15u);  // (2 * stride) mod 32 words == 4: stride mod 16 doubles == 2
}
__host__ __device__ inline size_t rollout_smem_bytes(const ObstacleImage &img, int groups, int lanes_per_agent,
                                                     int known_words) {
  size_t b = 16 + img.bytes;
  b += (size_t)groups * rollout_fbuf_stride(lanes_per_agent) * sizeof(double);
  b += (size_t)groups * rollout_cand_stride(img.n_obs, lanes_per_agent) * sizeof(uint16_t);
  b += (size_t)groups * known_words * sizeof(uint32_t);
  return (b + 15) & ~(size_t)15;
}

// loop-invariant planner values, pinned in registers (see keep())
struct StepEnv {
  v3 goal;
  int n_obs;
  double pred_dt;
};

// ---- broad phase: fp32 sphere test, ordered compaction of candidate indices into cand[] --------------
// Straight-line version for up to ROUNDS * kLanes field obstacles (no branch: it is meant to sit in the
// same basic block as the prologue's FP64 chains so that the scheduler interleaves them).
#pragma nv_exec_check_disable
template <int ROUNDS, class G>
PMAF_HDT int broad_phase_unrolled(const G &g, const float4 *bp, int n_field, v3 p, uint16_t *cand) {
  constexpr int LPA = G::kLanes;
  const float fx = (float)p.x, fy = (float)p.y, fz = (float)p.z;
  const unsigned lt_mask = g.mask & ((1u << g.lane) - 1u);
  int n_cand = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < ROUNDS; ++r) {
    const int i = r * LPA + g.gl;
    const bool in_range = i < n_field;
    // unconditional load of a valid record (the image always holds n_field + 1 >= 1 of them): no branch
    // around the shared-memory read, its latency overlaps the prologue's FP64 chains
    const float4 b = bp[i < n_field ? i : n_field];
    const float dx = b.x - fx, dy = b.y - fy, dz = b.z - fz;
    const bool cnd = in_range & (dx * dx + dy * dy + dz * dz < b.w);
    const unsigned m = g.ballot(cnd);
    if (cnd) cand[n_cand + PMAF_POPC(m & lt_mask)] = (uint16_t)i;
    n_cand += PMAF_POPC(m);
  }
  return n_cand;
}
#pragma nv_exec_check_disable
template <class G>
PMAF_HDT int broad_phase_loop(const G &g, const float4 *bp, int n_field, v3 p, uint16_t *cand) {
  constexpr int LPA = G::kLanes;
  const float fx = (float)p.x, fy = (float)p.y, fz = (float)p.z;
  const unsigned lt_mask = g.mask & ((1u << g.lane) - 1u);
  int n_cand = 0;
  // four rounds per iteration: independent loads and compares, a quarter of the loop overhead
  constexpr int R = 4;
  for (int base = 0; base < n_field; base += R * LPA) {
    float4 b[R];
    bool c[R];
    unsigned m[R];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < R; ++r) {  // unconditional loads of valid records (index n_field is the sentinel's)
      const int i = base + r * LPA + g.gl;
      b[r] = bp[i < n_field ? i : n_field];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < R; ++r) {
      const int i = base + r * LPA + g.gl;
      const float dx = b[r].x - fx, dy = b[r].y - fy, dz = b[r].z - fz;
      c[r] = (i < n_field) & (dx * dx + dy * dy + dz * dz < b[r].w);
      m[r] = g.ballot(c[r]);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < R; ++r) {
      if (c[r]) cand[n_cand + PMAF_POPC(m[r] & lt_mask)] = (uint16_t)(base + r * LPA + g.gl);
      n_cand += PMAF_POPC(m[r]);
    }
  }
  return n_cand;
}

// The same loop with warp-convergent collectives (Group::ballot_w): every lane of the warp runs it, each group
// compacts its own list (2 / 4 agents per warp).
template <class G>
__device__ __forceinline__ int broad_phase_loop_w(const G &g, const float4 *bp, int n_field, v3 p, uint16_t *cand) {
  constexpr int LPA = G::kLanes;
  const float fx = (float)p.x, fy = (float)p.y, fz = (float)p.z;
  const unsigned lt_mask = (1u << g.gl) - 1u;
  int n_cand = 0;
  constexpr int R = 4;
  for (int base = 0; base < n_field; base += R * LPA) {
    float4 b[R];
    bool c[R];
    unsigned m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = base + r * LPA + g.gl;
      b[r] = bp[i < n_field ? i : n_field];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = base + r * LPA + g.gl;
      const float dx = b[r].x - fx, dy = b[r].y - fy, dz = b[r].z - fz;
      c[r] = (i < n_field) & (dx * dx + dy * dy + dz * dz < b[r].w);
      m[r] = g.ballot_w(c[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (c[r]) cand[n_cand + __popc(m[r] & lt_mask)] = (uint16_t)(base + r * LPA + g.gl);
      n_cand += __popc(m[r]);
    }
  }
  return n_cand;
}

// Step prologue: everything that depends only on (p, v) — the norms, the unit vectors the field pass
// needs, attractorForce's desired velocity and, for small obstacle sets, the broad phase — in ONE
// basic block: five independent sqrt/division chains plus the fp32/integer broad phase interleave
// instead of running back to back (a single warp issues in order; only the compiler's scheduling inside
// a block overlaps them). The broad phase is speculative (its only side effect is the scratch list).
struct Prologue {
  StepNorms sn;
  v3 ghat, nv_static;
  int n_cand;  // -1: broad phase not done yet (large obstacle sets)
};
constexpr int kBroadUnrolledRounds = 2;
#pragma nv_exec_check_disable
// EAGER (latency build): unit vectors and broad phase are evaluated speculatively here; otherwise
// (throughput builds) only the norms, and the rest on demand in agent_step — fewer instructions.
template <bool STATIC_VEL, bool EAGER, class G>
PMAF_HDT Prologue step_prologue(const G &g, const float4 *bp, int n_field, uint16_t *cand, v3 goal_vec, v3 p, v3 v,
                                double zseg, bool has_seg, const AgentConsts &c) {
  Prologue pr;
  FastMath fm;
  pr.sn = step_norms(fm, goal_vec, v, zseg, has_seg, c);
  pr.n_cand = -1;
  pr.ghat = pr.nv_static = mk3(0.0, 0.0, 0.0);
  if (EAGER) {
    step_units<STATIC_VEL>(fm, goal_vec, v, pr.sn, pr.ghat, pr.nv_static);
    const bool small = n_field <= kBroadUnrolledRounds * G::kLanes;
    pr.n_cand = broad_phase_unrolled<kBroadUnrolledRounds>(g, bp, small ? n_field : 0, p, cand);
    if (!small) pr.n_cand = -1;
  }
  if (__builtin_expect(fm.bad(), 0)) {
    ExactMath em;
    pr.sn = step_norms(em, goal_vec, v, zseg, has_seg, c);
    if (EAGER) step_units<STATIC_VEL>(em, goal_vec, v, pr.sn, pr.ghat, pr.nv_static);
  }
  return pr;
}

// The same without the in-place fallback (rotated loop of the latency build): an operand outside FastMath's
// range only raises `bad`; the caller re-evaluates with redo_prologue_exact on its slow path, so the hot
// block carries no reconvergence region.
template <bool STATIC_VEL, bool BROAD = true, class G>
PMAF_HDT Prologue step_prologue_nofallback(const G &g, const float4 *bp, int n_field, uint16_t *cand, v3 goal_vec, v3 p,
                                           v3 v, double zseg, bool has_seg, const AgentConsts &c, bool &bad) {
  Prologue pr;
  FastMath fm;
  pr.sn = step_norms(fm, goal_vec, v, zseg, has_seg, c);
  step_units<STATIC_VEL>(fm, goal_vec, v, pr.sn, pr.ghat, pr.nv_static);
  pr.n_cand = -1;
  if (BROAD) {  // large obstacle sets (MULTI) run the broad phase as a loop afterwards
    const bool small = n_field <= kBroadUnrolledRounds * G::kLanes;
    pr.n_cand = broad_phase_unrolled<kBroadUnrolledRounds>(g, bp, small ? n_field : 0, p, cand);
    if (!small) pr.n_cand = -1;
  }
  bad = fm.bad();
  return pr;
}
template <bool STATIC_VEL>
PMAF_HDT void redo_prologue_exact(Prologue &pr, v3 goal_vec, v3 v, double zseg, bool has_seg, const AgentConsts &c) {
  ExactMath em;
  pr.sn = step_norms(em, goal_vec, v, zseg, has_seg, c);
  step_units<STATIC_VEL>(em, goal_vec, v, pr.sn, pr.ghat, pr.nv_static);
}

}  // namespace pmaf
#include "pmaf_fast.cuh"
namespace pmaf {

// One integration step of one agent (loop body of cfPrediction, cf_agent.cpp:312-326).
// Returns the new position in p / velocity in v; updates min_obs.
#pragma nv_exec_check_disable
template <bool STATIC_VEL, bool SPEC, class G>
PMAF_HDT void agent_step(const G &g, const StepEnv &P, const SmemObstacles &obs, const float4 *bp, uint16_t *cand,
                         double *fbuf, const KnownBits &known, int type, const AgentConsts &c, v3 init_pos,
                         double *rot_row, const double *random_row, v3 goal_vec, const Prologue &pr, v3 &p, v3 &v,
                         double &min_obs PMAF_T_ARGS) {
  const StepNorms &sn = pr.sn;
  const v3 goal = P.goal;
  const int n_field = P.n_obs - 1;  // the sentinel is excluded from the field loops (:75)
  v3 force = mk3(0.0, 0.0, 0.0);    // resetForce()
  double k_goal_scale = 1.0;
  if (field_gate_open(sn.dist_goal, sn.vn, p, init_pos, c)) {
    PMAF_T(8);
    const int n_cand = pr.n_cand >= 0 ? pr.n_cand : broad_phase_loop(g, bp, n_field, p, cand);
    PMAF_T(1);
    if (n_cand > 0) {
      g.sync();
      // ---- narrow phase ----
      v3 ghat = pr.ghat, nv_static = pr.nv_static;
      if (!SPEC) {  // throughput builds: unit vectors on demand
        FastMath fm;
        step_units<STATIC_VEL>(fm, goal_vec, v, sn, ghat, nv_static);
        if (__builtin_expect(fm.bad(), 0)) {
          ExactMath em;
          step_units<STATIC_VEL>(em, goal_vec, v, sn, ghat, nv_static);
        }
      }
      double min_d, kgs_closest;
      bool has_closest;
      field_pass<STATIC_VEL, SPEC>(g, obs, n_field, cand, n_cand, type, p, v, goal_vec, sn, nv_static, goal, ghat,
                                   c, known, rot_row, random_row, fbuf, min_obs, force, min_d, has_closest,
                                   kgs_closest PMAF_T_PASS);
      if (min_d < min_obs) min_obs = min_d;
      // `if (force_.norm() > 1e-5) k_goal_scale = attractorForceScaling()` (:319-321); no close obstacle: 1 (:212-214)
      if (has_closest && norm_gt(dot3(force, force), make_thr(1e-5))) k_goal_scale = kgs_closest;
    }
  }
  g.sync();  // cand[] is rewritten by the next step's (speculative) broad phase
  const v3 o_s = obs.pos(P.n_obs - 1);
  const v3 p0 = p, v0 = v, f0 = force;
  FastMath fm;
  finish_step(fm, force, k_goal_scale, sn, o_s, P.pred_dt, c, p, v);
  if (__builtin_expect(fm.bad(), 0)) {
    ExactMath em;
    p = p0, v = v0, force = f0;
    finish_step(em, force, k_goal_scale, sn, o_s, P.pred_dt, c, p, v);
  }
  PMAF_T(6);
}

// predictObstacles (:270-276) for the CTA's shared obstacle image: exact positions by repeated addition of
// vel * dt, and the broad-phase centres refreshed from them
__device__ __forceinline__ void advance_obstacles(unsigned char *img, const ObstacleImage &im, int n_obs,
                                                  const SmemObstacles &obs, const float4 *bp) {
  double *px = const_cast<double *>(obs.px), *py = const_cast<double *>(obs.py), *pz = const_cast<double *>(obs.pz);
  const double *dx = reinterpret_cast<const double *>(img + im.off_dx);
  const double *dy = reinterpret_cast<const double *>(img + im.off_dy);
  const double *dz = reinterpret_cast<const double *>(img + im.off_dz);
  float4 *bpw = const_cast<float4 *>(bp);
  for (int i = threadIdx.x; i < n_obs; i += blockDim.x) {
    const double x = px[i] + dx[i], y = py[i] + dy[i], z = pz[i] + dz[i];
    px[i] = x, py[i] = y, pz[i] = z;
    float4 b = bpw[i];
    b.x = (float)x, b.y = (float)y, b.z = (float)z;
    bpw[i] = b;
  }
}

// Rollout of every agent to termination. Each group continues ITS agent from the agent's current
// state (latest path point, velocity, min_obs_dist, known flags, path length, workspace cost) —
// which resetEEAgents made uniform in the normal tick order — so any call order of the reference
// API keeps its meaning; an already terminated agent executes zero steps.
// OCC = resident CTAs per SM the register budget is sized for: 1 (up to 255 registers: the latency-bound
// case, one warp per scheduler), or 3 / 4 CTAs of 128 threads (170 / 128 registers: populations that
// fill the machine trade registers for resident warps).
// MULTI (static latency build only): more than 64 field obstacles — broad phase as a loop, narrow phase in
// chunks of 32 candidates (fast_step_multi).
template <int LPA, bool DYNAMIC, int OCC, bool MULTI = false>
__global__ void __launch_bounds__(OCC == 1 ? 256 : 128, OCC) rollout_kernel(const PlannerDev P) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  unsigned char *img = smem + 16;
  const int groups = blockDim.x / LPA;
  double *fbuf_all = reinterpret_cast<double *>(img + P.img.bytes);
  const uint32_t fbuf_stride = rollout_fbuf_stride(LPA);
  uint16_t *cand_all = reinterpret_cast<uint16_t *>(fbuf_all + (size_t)groups * fbuf_stride);
  const uint32_t cand_stride = rollout_cand_stride(P.n_obs, LPA);
  uint32_t *known_all = reinterpret_cast<uint32_t *>(cand_all + (size_t)groups * cand_stride);

  // stage the obstacle set: one TMA bulk copy per CTA, completion on an mbarrier
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  // programmatic dependent launch (pmaf_tick): everything above overlapped the tail of tick_kernel; nothing it
  // writes (obstacle image, real agent, known flags) is touched before this point. A no-op otherwise.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, P.img.bytes);
    tma_bulk_g2s(img, P.image, P.img.bytes, bar);
  }

  const Group<LPA> g;
  const int group_in_block = threadIdx.x / LPA;
  const int a = blockIdx.x * groups + group_in_block;  // local agent index
  const bool have_agent = a < P.n_agents;
  uint16_t *cand = cand_all + (size_t)group_in_block * cand_stride;
  double *fbuf = fbuf_all + (size_t)group_in_block * fbuf_stride;
  KnownBits known;
  known.w = known_all + (size_t)group_in_block * P.known_words;

  // per-agent state (overlaps the bulk copy)
  v3 p = mk3(0, 0, 0), v = p, init_pos = p;
  double min_obs = 0, path_len = 0, ws_cost = 0;
  int n_path = 0, type = 0;
  AgentConsts k = make_agent_consts(0, 0, 0, 1, P.shell, P.vel_max, P.approach_dist, P.mass, 0);
  double *rot_row = nullptr;
  const double *random_row = nullptr;
  if (have_agent) {
    init_pos = ld3(P.init_pos + 3 * a);
    type = agent_type_of_index(P.first_agent + a);
    rot_row = P.rot + (size_t)a * P.n_obs * 3;
    random_row = P.random_vecs + (size_t)a * P.n_obs * 3;
    if (P.reset_in_prologue) {
      // resetEEAgents (cf_manager.cpp:246-255) for this agent, in registers: setPosition (path := [pos]),
      // setVelocity (clamped, cf_agent.cpp:54-61), setObstacles' known flags (:63-70), resetMinObsDist
      p = ld3(P.reset_real->pos);
      v = clamp_velocity_cold(ld3(P.reset_real->vel), P.vel_max);
      min_obs = P.shell;
      path_len = 0.0;
      ws_cost = P.fused_valid ? add_workspace_cost(0.0, p, P.fused_cost.ws, P.fused_cost.k_workspace) : 0.0;
      n_path = 1;
      if (g.gl == 0) st3(P.paths + (size_t)a * P.max_steps * 3, p);
      for (int w = g.gl; w < P.known_words; w += LPA)
        known.w[w] = (P.known[(size_t)a * P.known_words + w] & P.reset_known_keep[w]) | P.reset_known_bits[w];
    } else {
      p = ld3(P.cur_pos + 3 * a);
      v = ld3(P.vel + 3 * a);
      min_obs = P.min_obs_dist[a];
      path_len = P.path_len[a];
      ws_cost = P.ws_cost[a];
      n_path = P.n_path[a];
      for (int w = g.gl; w < P.known_words; w += LPA) known.w[w] = P.known[(size_t)a * P.known_words + w];
    }
  }
  g.sync();
  mbar_wait(bar, 0);

  SmemObstacles obs;
  obs.px = reinterpret_cast<const double *>(img + P.img.off_px);
  obs.py = reinterpret_cast<const double *>(img + P.img.off_py);
  obs.pz = reinterpret_cast<const double *>(img + P.img.off_pz);
  obs.rs = reinterpret_cast<const double *>(img + P.img.off_rs);
  obs.vx = reinterpret_cast<const double *>(img + P.img.off_vx);
  obs.vy = reinterpret_cast<const double *>(img + P.img.off_vy);
  obs.vz = reinterpret_cast<const double *>(img + P.img.off_vz);
  obs.dynamic = DYNAMIC;
  const float4 *bp = reinterpret_cast<const float4 *>(img + P.img.off_bp);
  // always 0, but for the latency build only known at run time (see keep()); the occupancy builds
  // cannot afford the registers and keep re-reading the constant bank
  const unsigned rz = OCC == 1 ? *P.runtime_zero : 0u;
  if (have_agent) {
    const int ga = P.first_agent + a;  // gains are indexed by GLOBAL agent index
    k = make_agent_consts(P.k_attr[ga], P.k_circ[ga], P.k_repel[ga], P.k_damp[ga], P.shell, P.vel_max,
                          P.approach_dist, P.mass, obs.rsum(P.n_obs - 1), rz);
  }

  StepEnv env;
  env.goal = mk3(keep(P.goal[0], rz), keep(P.goal[1], rz), keep(P.goal[2], rz));
  env.n_obs = keep(P.n_obs, rz), env.pred_dt = keep(P.pred_dt, rz);
  const v3 goal = env.goal;
  const int max_steps = keep(P.max_steps, rz);
  const bool fused = keep(P.fused_valid, rz) != 0;
  const WsParams wsp = pin_ws(P.fused_cost.ws, P.fused_cost.k_workspace, rz);
  // 255-register build: common steps take the straight-line path (pmaf_fast.cuh) — one warp per agent (latency
  // shapes: fast_step, or fast_step_multi for more than 64 field obstacles), or 2 / 4 agents per warp
  // (throughput shapes: always the chunked fast_step_multi)
  constexpr bool FAST = (OCC == 1 && LPA >= 8) || (LPA >= 8 && LPA < 32);  // packed shapes: at every register budget
  constexpr bool CHUNKED = MULTI || LPA < 32;
  static_assert(!MULTI || (OCC == 1 && LPA == 32 && !DYNAMIC), "MULTI exists for the static one-warp-per-agent build only");
  if (FAST && have_agent) {  // the agent's rotation-vector row into L1 now: its first uses sit on the critical path
    const char *row = reinterpret_cast<const char *>(rot_row);
    for (int off = g.gl * 128; off < P.n_obs * 24; off += LPA * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(row + off));
  }
  FastConsts fc;
  if (FAST) fc = make_fast_consts(k, type, rz);
  const unsigned long long t0 = global_timer_ns();
  int steps_run = 0, general_steps = 0;
  bool alive = have_agent;
  double *path_row = have_agent ? P.paths + (size_t)a * P.max_steps * 3 : nullptr;

  double zseg = 1.0;  // |last path segment|^2: its square root joins the next step's prologue
  bool has_seg = false;
  PMAF_T_DECL;
#if defined(PMAF_FAST_STATS)
  long long st_fast_cyc = 0, st_gen_cyc = 0, st_gen = 0, st_cand = 0, st_latch_cyc = 0, st_latch = 0;
#endif
  const uint16_t *nn_table = P.img.nn_valid ? reinterpret_cast<const uint16_t *>(img + P.img.off_nn) : nullptr;
  if constexpr (FAST && LPA < 32) {
    // Packed shapes (2 / 4 agents per warp), static and moving scenes: the rotated loop of the latency build, run
    // by the whole warp in lockstep — the hot part of an iteration (commit of the previous step, prologue, broad
    // phase, fast_step_packed) is executed by all 32 lanes together with warp-convergent collectives; a group
    // whose agent has finished (or does not exist) stays in the loop predicated off until the warp's last agent
    // is done. Only termination and rare events take a group-divergent branch.
    v3 prev = p;
    bool pending = false;  // a step was taken whose commit is outstanding
    for (;;) {
      const bool had_step = pending;  // the previous iteration took a step: commit it now
      const v3 seg = sub3(p, prev);
      const double zs = dot3(seg, seg);
      const v3 goal_vec = sub3(goal, p);
      bool pr_bad;
      Prologue pr = step_prologue_nofallback<!DYNAMIC, false>(g, bp, env.n_obs - 1, cand, goal_vec, p, v, zs, had_step, k, pr_bad);
      const StepNorms &sn = pr.sn;
      const double path_len_before = path_len;
      path_len += sn.seg_len;  // getPathLength term (:29), in path order; +0.0 when no step is outstanding
      if (fused) {
        const double w = add_workspace_cost_bf(ws_cost, p, wsp.ws, wsp.k_workspace);
        ws_cost = had_step ? w : ws_cost;
      }
      st3_if(path_row + (size_t)n_path * 3, p, had_step & (g.gl == 0));
      n_path += had_step ? 1 : 0, steps_run += had_step ? 1 : 0;
      const bool step_on = alive & (sn.dist_goal > 0.1) & (n_path < max_steps) & !pr_bad;  // :310-311
      prev = p;
      pr.n_cand = broad_phase_loop_w(g, bp, env.n_obs - 1, p, cand);
      const bool done = fast_step_packed<!DYNAMIC>(g, env, obs, cand, fbuf, known, type, k, fc, init_pos, rot_row, random_row,
                                                   goal_vec, pr, p, v, min_obs, step_on, DYNAMIC ? nullptr : nn_table);
      pending = done;
      if (__builtin_expect(!done & alive, 0)) {  // group-divergent: termination or a rare event
        if (pr_bad) {  // an operand outside FastMath's range: the prologue again with the IEEE built-ins
          redo_prologue_exact<!DYNAMIC>(pr, goal_vec, v, zs, had_step, k);
          path_len = path_len_before + sn.seg_len;
        }
        if (sn.dist_goal > 0.1 && n_path < max_steps) {
          ++general_steps;
          agent_step<!DYNAMIC, true>(g, env, obs, bp, cand, fbuf, known, type, k, init_pos, rot_row, random_row, goal_vec, pr,
                                     p, v, min_obs PMAF_T_PASS);
          pending = true;
        } else {
          alive = false;
        }
      }
      if (DYNAMIC) {
        // predictObstacles (:270-276): the CTA's shared obstacle image advances in lockstep with its agents
        if (!__syncthreads_or(alive)) break;
        advance_obstacles(img, P.img, env.n_obs, obs, bp);
        __syncthreads();
      } else if (!__any_sync(0xffffffffu, alive)) {
        break;
      }
    }
  } else if constexpr (FAST && !DYNAMIC) {
    // Latency build, static scene: the loop is rotated — the commit of the previous step (path point,
    // workspace cost, counters) shares ONE basic block with this step's prologue, and one branch decides
    // between the straight-line step and everything else (termination, closed gate, rare events).
    if (alive) {
      v3 prev = p;
      bool pending = false;  // a step was taken whose commit is outstanding
      for (;;) {
        const v3 seg = sub3(p, prev);
        const double zs = dot3(seg, seg);
        const v3 goal_vec = sub3(goal, p);
        bool pr_bad;
        Prologue pr = step_prologue_nofallback<true, !CHUNKED>(g, bp, env.n_obs - 1, cand, goal_vec, p, v, zs, pending, k, pr_bad);
        const StepNorms &sn = pr.sn;
        const double path_len_before = path_len;
        path_len += sn.seg_len;  // getPathLength term (:29), in path order; 0 while nothing is pending
        if (fused) {
          const double w = add_workspace_cost_bf(ws_cost, p, wsp.ws, wsp.k_workspace);
          ws_cost = pending ? w : ws_cost;
        }
        st3_if(path_row + (size_t)n_path * 3, p, pending & (g.gl == 0));
        n_path += pending ? 1 : 0, steps_run += pending ? 1 : 0;
        bool step_on = (sn.dist_goal > 0.1) & (n_path < max_steps) & !pr_bad;  // :310-311
        prev = p;
#if defined(PMAF_FAST_STATS)
        unsigned why_arr[3] = {0u, 0u, 0u};
        unsigned *why = why_arr;
        const long long tq0 = clock64();
        st_cand += pr.n_cand > 0 ? pr.n_cand : 0;
#else
        unsigned *why = nullptr;
#endif
        bool done;
        if constexpr (CHUNKED) {
          pr.n_cand = broad_phase_loop(g, bp, env.n_obs - 1, p, cand);
          done = fast_step_multi<true>(g, env, obs, cand, fbuf, known, type, k, fc, init_pos, rot_row, random_row, goal_vec,
                                       pr, p, v, min_obs, step_on, nn_table);
        } else {
          done = fast_step<true>(g, env, obs, cand, fbuf, known, type, k, fc, init_pos, rot_row, random_row, goal_vec, pr,
                                 p, v, min_obs, why, step_on, nn_table);
        }
        if (!done) {
          if (pr_bad) {  // an operand outside FastMath's range: the prologue again with the IEEE built-ins
            redo_prologue_exact<true>(pr, goal_vec, v, zs, pending, k);
            path_len = path_len_before + sn.seg_len;
            step_on = sn.dist_goal > 0.1 && n_path < max_steps;
          }
          if (!step_on) break;
          ++general_steps;
#if defined(PMAF_FAST_STATS)
          if (g.gl == 0)
            for (int b = 0; b < 12; ++b)
              if (why_arr[0] >> b & 1u) atomicAdd(P.step_counter + 4 + b, 1ull);
#endif
          agent_step<true, true>(g, env, obs, bp, cand, fbuf, known, type, k, init_pos, rot_row, random_row, goal_vec, pr,
                                 p, v, min_obs PMAF_T_PASS);
        }
#if defined(PMAF_FAST_STATS)
        if (done) st_fast_cyc += clock64() - tq0; else st_gen_cyc += clock64() - tq0, ++st_gen;
        st_latch_cyc += why_arr[1], st_latch += why_arr[2];
#endif
        pending = true;
      }
    }
  } else if constexpr (FAST && DYNAMIC) {
    // The same rotated loop for moving obstacles: the CTA's warps step in lockstep around the shared
    // obstacle image (a finished agent's warp keeps taking part in the barriers).
    v3 prev = p;
    bool pending = false;
    for (;;) {
      if (alive) {
        const v3 seg = sub3(p, prev);
        const double zs = dot3(seg, seg);
        const v3 goal_vec = sub3(goal, p);
        bool pr_bad;
        const bool had_step = pending;  // the previous iteration took a step: commit it now
        Prologue pr = step_prologue_nofallback<false, !CHUNKED>(g, bp, env.n_obs - 1, cand, goal_vec, p, v, zs, had_step, k, pr_bad);
        const StepNorms &sn = pr.sn;
        const double path_len_before = path_len;
        path_len += sn.seg_len;
        if (fused) {
          const double w = add_workspace_cost_bf(ws_cost, p, wsp.ws, wsp.k_workspace);
          ws_cost = had_step ? w : ws_cost;
        }
        st3_if(path_row + (size_t)n_path * 3, p, had_step & (g.gl == 0));
        n_path += had_step ? 1 : 0, steps_run += had_step ? 1 : 0;
        pending = false;
        bool step_on = (sn.dist_goal > 0.1) & (n_path < max_steps) & !pr_bad;  // :310-311
        prev = p;
        bool done;
        if constexpr (CHUNKED) {
          pr.n_cand = broad_phase_loop(g, bp, env.n_obs - 1, p, cand);
          done = fast_step_multi<false>(g, env, obs, cand, fbuf, known, type, k, fc, init_pos, rot_row, random_row, goal_vec,
                                        pr, p, v, min_obs, step_on, nullptr);
        } else {
          done = fast_step<false>(g, env, obs, cand, fbuf, known, type, k, fc, init_pos, rot_row, random_row, goal_vec, pr,
                                  p, v, min_obs, nullptr, step_on);
        }
        if (!done) {
          if (pr_bad) {
            redo_prologue_exact<false>(pr, goal_vec, v, zs, had_step, k);
            path_len = path_len_before + sn.seg_len;
            step_on = sn.dist_goal > 0.1 && n_path < max_steps;
          }
          if (step_on) {
            ++general_steps;
            agent_step<false, true>(g, env, obs, bp, cand, fbuf, known, type, k, init_pos, rot_row, random_row, goal_vec,
                                    pr, p, v, min_obs PMAF_T_PASS);
            pending = true;
          } else {
            alive = false;
          }
        } else {
          pending = true;
        }
      }
      if (!__syncthreads_or(alive)) break;
      advance_obstacles(img, P.img, env.n_obs, obs, bp);
      __syncthreads();
    }
  } else
  if (DYNAMIC || alive)  // static scenes: an agent's warp leaves the loop directly when its rollout ends
  for (;;) {
    if (!DYNAMIC || alive) {
      const v3 goal_vec = sub3(goal, p);
      const Prologue pr = step_prologue<!DYNAMIC, OCC == 1>(g, bp, env.n_obs - 1, cand, goal_vec, p, v, zseg, has_seg, k);
      const StepNorms &sn = pr.sn;
      PMAF_T(0);
      path_len += sn.seg_len;  // getPathLength term (:29), in path order
      has_seg = false;
      if (sn.dist_goal > 0.1 && n_path < max_steps) {  // :310-311
        const v3 prev = p;
        bool done = false;
#if defined(PMAF_FAST_STATS)
        unsigned why_arr[3] = {0u, 0u, 0u};
        unsigned &why_bits = why_arr[0];
        unsigned *why = why_arr;
#else
        unsigned *why = nullptr;
#endif
#if defined(PMAF_FAST_STATS)
        const long long tq0 = clock64();
        st_cand += pr.n_cand > 0 ? pr.n_cand : 0;
#endif
        if constexpr (FAST)
          done = fast_step<!DYNAMIC>(g, env, obs, cand, fbuf, known, type, k, fc, init_pos, rot_row, random_row, goal_vec,
                                     pr, p, v, min_obs, why);
        if (!done) {
          ++general_steps;
#if defined(PMAF_FAST_STATS)
          if (g.gl == 0)
            for (int b = 0; b < 12; ++b)
              if (why_bits >> b & 1u) atomicAdd(P.step_counter + 4 + b, 1ull);
#endif
          agent_step<!DYNAMIC, OCC == 1>(g, env, obs, bp, cand, fbuf, known, type, k, init_pos, rot_row, random_row,
                                         goal_vec, pr, p, v, min_obs PMAF_T_PASS);
        }
#if defined(PMAF_FAST_STATS)
        if (done) st_fast_cyc += clock64() - tq0; else st_gen_cyc += clock64() - tq0, ++st_gen;
        st_latch_cyc += why_arr[1], st_latch += why_arr[2];
#endif
        const v3 seg = sub3(p, prev);
        zseg = dot3(seg, seg), has_seg = true;
        if (fused) ws_cost = add_workspace_cost(ws_cost, p, wsp.ws, wsp.k_workspace);
        PMAF_T(9);
        if (g.gl == 0) st3(path_row + (size_t)n_path * 3, p);
        ++n_path;
        ++steps_run;
        PMAF_T(7);
      } else {
        if (!DYNAMIC) break;
        alive = false;
      }
    }
    if (DYNAMIC) {
      // predictObstacles (:270-276): every private obstacle copy advances identically, so the CTA
      // keeps ONE copy and steps it in lockstep with its agents
      if (!__syncthreads_or(alive)) break;
      advance_obstacles(img, P.img, env.n_obs, obs, bp);
      __syncthreads();
    }
  }

  if (have_agent) {
    for (int w = g.gl; w < P.known_words; w += LPA) P.known[(size_t)a * P.known_words + w] = known.w[w];
    if (g.gl == 0) {
      st3(P.cur_pos + 3 * a, p);
      st3(P.vel + 3 * a, v);
      P.min_obs_dist[a] = min_obs;
      P.path_len[a] = path_len;
      P.ws_cost[a] = ws_cost;
      P.n_path[a] = n_path;
#if defined(PMAF_FAST_STATS)
      if (a < 64 && P.section_cycles) {
        P.section_cycles[a * 12 + 0] = st_fast_cyc, P.section_cycles[a * 12 + 1] = st_gen_cyc;
        P.section_cycles[a * 12 + 2] = st_gen, P.section_cycles[a * 12 + 3] = st_cand;
        P.section_cycles[a * 12 + 4] = steps_run;
        P.section_cycles[a * 12 + 5] = st_latch_cyc, P.section_cycles[a * 12 + 6] = st_latch;
      }
#endif
#if defined(PMAF_SECTION_TIMERS) && defined(__CUDA_ARCH__)
      if (a < 64 && P.section_cycles)
        for (int q = 0; q < 12; ++q) P.section_cycles[a * 12 + q] = pmaf_sec_[q];
#endif
      if (steps_run > 0) {  // `if (running_)`, :330-337
        P.pred_time_ns[a] = (double)(global_timer_ns() - t0);
        P.reached[a] = norm3(sub3(goal, p)) < 0.100001 ? 1 : 0;
        atomicAdd(P.step_counter, (unsigned long long)steps_run);
        atomicAdd(P.step_counter + 1, (unsigned long long)steps_run);
        atomicAdd(P.step_counter + 2, (unsigned long long)general_steps);
      }
    }
  }
}

// broad-phase record of one obstacle: fp32 centre and squared candidate radius
PMAF_HDT float4 broad_phase_record(v3 pos, double shell, double rsum, float margin) {
  const float thr = (float)(shell + rsum) + margin;
  float4 r;
  r.x = (float)pos.x, r.y = (float)pos.y, r.z = (float)pos.z, r.w = thr * thr;
  return r;
}

// ---- resetEEAgents (cf_manager.cpp:246-255) + obstacle staging -------------------------------------------
struct ResetArgs {
  const double *pos_vel;      // [6] position, velocity handed to resetEEAgents (device copy), or
  const RealState *real;      // ... the real agent's state when from_real != 0 (fused tick)
  int from_real;
  int n_obs_update;           // obstacles[0..n) get new pos/vel (setObstacles, cf_agent.cpp:63-70)
  const double *new_pos, *new_vel;  // [n][3] live obstacle positions / velocities (device)
  double *obs_pos, *obs_vel;  // [O][3] the agents' obstacle copy at rollout start (updated here)
  const double *obs_rad;      // [O] radii from init()
  const unsigned char *real_known;  // [O]
  unsigned char *image;       // staging image to (re)build
  float margin;               // broad-phase safety margin (absolute, metres)
  int do_agents;              // 0: only rebuild the staging image
  int set_known;              // 1: known := real agent's flags for the passed obstacles
  int reset_velocity;         // 1: vel := clamp(v); min_obs_dist := shell
  int agent_blocks;           // blocks [0, agent_blocks) do the work above; further blocks build the nn table
};

// the obstacle positions a rollout will start from, as block 0 of reset_kernel is about to write them
struct ResetObstacles {
  const double *new_pos, *old_pos;
  int n_update;
  PMAF_HDT v3 pos(int i) const { return i < n_update ? ld3(new_pos + 3 * i) : ld3(old_pos + 3 * i); }
};

// the staging image of the obstacle set (ObstacleImage), rebuilt by `nthreads` threads of one CTA
__device__ __forceinline__ void build_obstacle_image(const PlannerDev &P, const ResetArgs &R, int tid, int nthreads) {
  double *px = reinterpret_cast<double *>(R.image + P.img.off_px);
  double *py = reinterpret_cast<double *>(R.image + P.img.off_py);
  double *pz = reinterpret_cast<double *>(R.image + P.img.off_pz);
  double *rs = reinterpret_cast<double *>(R.image + P.img.off_rs);
  float4 *bp = reinterpret_cast<float4 *>(R.image + P.img.off_bp);
  for (int i = tid; i < P.n_obs; i += nthreads) {
    v3 op, ov;
    if (i < R.n_obs_update) {
      op = ld3(R.new_pos + 3 * i), ov = ld3(R.new_vel + 3 * i);
      st3(R.obs_pos + 3 * i, op), st3(R.obs_vel + 3 * i, ov);
    } else {
      op = ld3(R.obs_pos + 3 * i), ov = ld3(R.obs_vel + 3 * i);
    }
    const double rsum = P.rad + R.obs_rad[i];
    px[i] = op.x, py[i] = op.y, pz[i] = op.z, rs[i] = rsum;
    if (P.img.dynamic) {
      reinterpret_cast<double *>(R.image + P.img.off_vx)[i] = ov.x;
      reinterpret_cast<double *>(R.image + P.img.off_vy)[i] = ov.y;
      reinterpret_cast<double *>(R.image + P.img.off_vz)[i] = ov.z;
      reinterpret_cast<double *>(R.image + P.img.off_dx)[i] = ov.x * P.pred_dt;  // getVelocity() * delta_t (:273)
      reinterpret_cast<double *>(R.image + P.img.off_dy)[i] = ov.y * P.pred_dt;
      reinterpret_cast<double *>(R.image + P.img.off_dz)[i] = ov.z * P.pred_dt;
    }
    bp[i] = broad_phase_record(op, P.shell, rsum, R.margin);
  }
}
// nearest-neighbour table of a static scene (ObstacleImage::off_nn): warp `warp` of `warps` takes every warps-th row
__device__ __forceinline__ void build_nn_table(const PlannerDev &P, const ResetArgs &R, int warp, int warps) {
  const Group<32> g;
  ResetObstacles obs;
  obs.new_pos = R.new_pos, obs.old_pos = R.obs_pos, obs.n_update = R.n_obs_update;
  uint16_t *nn = reinterpret_cast<uint16_t *>(R.image + P.img.off_nn);
  const int n_field = P.n_obs - 1;
  for (int row = warp; row < n_field; row += warps) {
    const int found = nearest_other_obstacle(g, obs, n_field, row);
    if (g.gl == 0) nn[row] = (uint16_t)found;
  }
}
// the real agent's known flags as bit words: `bits` for the obstacles of the passed list, `keep` = the bits of the
// agent's own word that survive (setObstacles only touches the obstacles of the passed list, cf_agent.cpp:63-70)
__device__ __forceinline__ void pack_known_word(const unsigned char *real_known, int n_update, int w, uint32_t &bits,
                                                uint32_t &keep) {
  bits = 0, keep = 0;
  for (int b = 0; b < 32; ++b) {
    const int i = w * 32 + b;
    if (i < n_update) bits |= (uint32_t)(real_known[i] != 0) << b;
    else keep |= 1u << b;
  }
}

// grid: ceil(A / blockDim) blocks (1 block when !do_agents); block 0 also rebuilds the staging image.
__global__ void __launch_bounds__(128) reset_kernel(const PlannerDev P, const ResetArgs R) {
  __shared__ uint32_t s_bits[kMaxObstacles / 32], s_keep[kMaxObstacles / 32];
  if ((int)blockIdx.x >= R.agent_blocks) {  // nearest-neighbour table of a static scene (ObstacleImage::off_nn)
    build_nn_table(P, R, (blockIdx.x - R.agent_blocks) * (blockDim.x / 32) + threadIdx.x / 32,
                   (gridDim.x - R.agent_blocks) * (blockDim.x / 32));
    return;
  }
  if (blockIdx.x == 0) build_obstacle_image(P, R, threadIdx.x, blockDim.x);
  if (!R.do_agents) return;
  if (R.set_known) {  // pack the real agent's flags once per block
    for (int w = threadIdx.x; w < P.known_words; w += blockDim.x) pack_known_word(R.real_known, R.n_obs_update, w, s_bits[w], s_keep[w]);
    __syncthreads();
  }
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.n_agents) return;
  v3 pos, vel;
  if (R.from_real) {
    pos = ld3(R.real->pos), vel = ld3(R.real->vel);
  } else {
    pos = ld3(R.pos_vel), vel = ld3(R.pos_vel + 3);
  }
  st3(P.cur_pos + 3 * a, pos);  // setPosition: path := [pos]
  st3(P.paths + (size_t)a * P.max_steps * 3, pos);
  P.n_path[a] = 1;
  P.path_len[a] = 0.0;
  P.ws_cost[a] = P.fused_valid ? add_workspace_cost(0.0, pos, P.fused_cost.ws, P.fused_cost.k_workspace) : 0.0;
  if (R.reset_velocity) {
    st3(P.vel + 3 * a, clamp_velocity(vel, P.vel_max));  // setVelocity
    P.min_obs_dist[a] = P.shell;                         // resetMinObsDist
  }
  if (R.set_known) {
    uint32_t *row = P.known + (size_t)a * P.known_words;
    for (int w = 0; w < P.known_words; ++w) row[w] = (row[w] & s_keep[w]) | s_bits[w];
  }
}

// ---- evaluateAgents (cf_manager.cpp:293-356) ---------------------------------------------------------------
// workspace cost recomputed from the stored paths: used when the cost parameters differ from the
// ones the rollout accumulated under (first tick, parameter change, setInitialPosition).
__global__ void workspace_cost_kernel(const PlannerDev P, const CostParams C) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.n_agents) return;
  const double *row = P.paths + (size_t)a * P.max_steps * 3;
  double cost = 0.0;
  const int n = P.n_path[a];
  for (int k = 0; k < n; ++k) cost = add_workspace_cost(cost, ld3(row + 3 * k), C.ws, C.k_workspace);
  P.ws_cost[a] = cost;
}

struct ArgminRecord {  // one rank's contribution to the global best-agent selection
  double min_cost;     // DBL_MAX if no agent beat it (serial scan start value, :336)
  double incumbent_cost;  // cost of the incumbent if this rank owns it, else NaN
  double cost_agent0;  // cost of global agent 0 if owned (the scan's default result), else NaN
  int min_index;       // GLOBAL index of the local serial argmin, INT_MAX if none
  int owns_incumbent;
  // followed in the exchange buffer by the local argmin agent's random vectors, double[O][3]
};
__host__ __device__ inline size_t argmin_record_bytes(int n_obs) {
  return sizeof(ArgminRecord) + (size_t)n_obs * 3 * sizeof(double);
}

// single block: per-agent costs, serial-order argmin (strict <, lowest index), then — unsharded —
// hysteresis and incumbent update.
__device__ __forceinline__ void evaluate_body(const PlannerDev &P, const CostParams &C, DeviceBest *best,
                                              double *best_random, ArgminRecord *rec, EvalResult *out, int finalize,
                                              HostOut *host, unsigned long long ticket) {
  __shared__ double s_cost[32];
  __shared__ int s_idx[32];
  const v3 goal = ld3(P.goal);
  double bc = 1.7976931348623157e308;  // std::numeric_limits<double>::max()
  int bi = 0x7fffffff;
  for (int a = threadIdx.x; a < P.n_agents; a += blockDim.x) {
    const double goal_dist = norm3(sub3(goal, ld3(P.cur_pos + 3 * a)));
    const double c = finish_cost(P.ws_cost[a], goal_dist, P.approach_dist, C.k_goal_dist, P.path_len[a],
                                 C.k_path_len, C.k_safe_dist, P.min_obs_dist[a]);
    P.cost[a] = c;
    if (c < bc) bc = c, bi = P.first_agent + a;  // ascending a: first index of the minimum
  }
  for (int off = 16; off > 0; off >>= 1) {
    const double oc = __shfl_xor_sync(0xffffffffu, bc, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (oc < bc || (oc == bc && oi < bi)) bc = oc, bi = oi;
  }
  if ((threadIdx.x & 31) == 0) s_cost[threadIdx.x >> 5] = bc, s_idx[threadIdx.x >> 5] = bi;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = (blockDim.x + 31) >> 5;
    bc = threadIdx.x < nw ? s_cost[threadIdx.x] : 1.7976931348623157e308;
    bi = threadIdx.x < nw ? s_idx[threadIdx.x] : 0x7fffffff;
    for (int off = 16; off > 0; off >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, bc, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (oc < bc || (oc == bc && oi < bi)) bc = oc, bi = oi;
    }
    if (threadIdx.x == 0) {
      s_cost[0] = bc, s_idx[0] = bi;
    }
  }
  __syncthreads();
  bc = s_cost[0], bi = s_idx[0];
  const int inc_local = best->present ? best->id - 1 - P.first_agent : -1;
  const bool owns = inc_local >= 0 && inc_local < P.n_agents;
  if (threadIdx.x == 0) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    rec->min_cost = bc;
    rec->min_index = bi;
    rec->owns_incumbent = owns;
    rec->incumbent_cost = owns ? P.cost[inc_local] : qnan;
    rec->cost_agent0 = P.first_agent == 0 ? P.cost[0] : qnan;
  }
  if (!finalize) {  // sharded: ship the local winner's random vectors with the record
    double *row = reinterpret_cast<double *>(rec + 1);
    const double *src = bi == 0x7fffffff ? nullptr : P.random_vecs + (size_t)(bi - P.first_agent) * P.n_obs * 3;
    for (int i = threadIdx.x; i < P.n_obs * 3; i += blockDim.x) row[i] = src ? src[i] : 0.0;
    return;
  }
  // unsharded: min_cost_idx defaults to 0 when no cost is below DBL_MAX (:335-342)
  const int min_idx = bi == 0x7fffffff ? 0 : bi;
  const double min_cost = P.cost[min_idx - P.first_agent];
  bool take = true;
  int result = min_idx;
  if (best->present && owns) {  // hysteresis :343-350 (the reference reads out of range if id-1 >= A)
    if (!(min_cost < 0.9 * P.cost[inc_local])) take = false, result = best->id - 1;
  }
  __syncthreads();
  if (take) {  // best_agent_ = ee_agents_[min]->makeCopy(): type, id and the RANDOM agent's vectors
    const double *src = P.random_vecs + (size_t)(min_idx - P.first_agent) * P.n_obs * 3;
    for (int i = threadIdx.x; i < P.n_obs * 3; i += blockDim.x) best_random[i] = src[i];
  }
  if (threadIdx.x == 0) {
    out->argmin_index = min_idx;
    out->argmin_cost = min_cost;
    out->incumbent_changed = take;
    out->best_index = result;
    out->best_cost = P.cost[result - P.first_agent];
    if (take) {
      best->present = 1;
      best->id = min_idx + 1;
      best->type = agent_type_of_index(min_idx);
    }
    if (host) {  // zero-copy result + ticket (see HostOut)
      host->eval.best_index = result, host->eval.argmin_index = min_idx, host->eval.incumbent_changed = take;
      host->eval.pad = 0, host->eval.best_cost = out->best_cost, host->eval.argmin_cost = min_cost;
      host->best.present = best->present, host->best.id = best->id, host->best.type = best->type, host->best.pad = 0;
      __threadfence_system();
      host->seq[0] = ticket;
    }
  }
}
__global__ void __launch_bounds__(1024) evaluate_kernel(const PlannerDev P, const CostParams C, DeviceBest *best,
                                                        double *best_random, ArgminRecord *rec, EvalResult *out,
                                                        int finalize, HostOut *host, unsigned long long ticket) {
  evaluate_body(P, C, best, best_random, rec, out, finalize, host, ticket);
}

// Sharded planners: after ONE all-gather of the per-rank records every rank repeats the reference's
// serial scan over the ranks in order (ranks own contiguous ascending agent blocks, so "lowest index
// wins" is preserved), applies the hysteresis and updates its replica of the incumbent.
// the replicated selection itself: thread 0 scans the per-rank records in rank order, applies the hysteresis
// and updates this rank's replica of the incumbent; the block then copies the winner's random vectors
__device__ __forceinline__ void select_over_records(const unsigned char *records, size_t stride, int world, int n_obs,
                                                    DeviceBest *best, double *best_random, EvalResult *out) {
  __shared__ int s_take, s_min_rank;
  if (threadIdx.x == 0) {
    double min_cost = 1.7976931348623157e308;
    int min_idx = 0, min_rank = -1;
    double inc_cost = 0.0, cost0 = 0.0;
    bool have_inc = false;
    for (int r = 0; r < world; ++r) {
      // volatile: the records may have been written by peers over NVLink (p2p_select_kernel); never from L1
      const volatile ArgminRecord *rec = reinterpret_cast<const volatile ArgminRecord *>(records + r * stride);
      const double rc = rec->min_cost;
      const int ri = rec->min_index;
      if (ri != 0x7fffffff && rc < min_cost) min_cost = rc, min_idx = ri, min_rank = r;
      if (rec->owns_incumbent) inc_cost = rec->incumbent_cost, have_inc = true;
      if (r == 0) cost0 = rec->cost_agent0;
    }
    if (min_rank < 0) min_cost = cost0;  // no cost below DBL_MAX: index 0 (:335-342)
    bool take = true;
    int result = min_idx;
    if (best->present && have_inc) {
      if (!(min_cost < 0.9 * inc_cost)) take = false, result = best->id - 1;
    }
    out->argmin_index = min_idx, out->argmin_cost = min_cost, out->incumbent_changed = take;
    out->best_index = result, out->best_cost = take ? min_cost : inc_cost;
    if (take) best->present = 1, best->id = min_idx + 1, best->type = agent_type_of_index(min_idx);
    s_take = take, s_min_rank = min_rank < 0 ? 0 : min_rank;
  }
  __syncthreads();
  if (s_take) {
    const double *row = reinterpret_cast<const double *>(records + s_min_rank * stride + sizeof(ArgminRecord));
    for (int i = threadIdx.x; i < n_obs * 3; i += blockDim.x) best_random[i] = __ldcg(row + i);
  }
}

__global__ void __launch_bounds__(256) global_select_kernel(const unsigned char *records, int world, int n_obs,
                                                            DeviceBest *best, double *best_random,
                                                            EvalResult *out) {
  select_over_records(records, argmin_record_bytes(n_obs), world, n_obs, best, best_random, out);
}

// ---- best-agent exchange over NVLink peer memory -----------------------------------------------------------------------
// Instead of an NCCL all-gather between the local scan and the replicated selection, ONE kernel does both:
// every rank stores its record straight into slot [rank] of every peer's exchange block (P2P stores over
// NVLink / NVSwitch; the blocks are cudaIpc-mapped), publishes a sequence number per peer, waits until its
// own block holds this tick's records of all ranks, and runs the selection. Two buffers alternate by tick
// parity: a rank can be at most one tick ahead of another (its next selection needs everyone's next
// record), so a slot is never overwritten before its reader is done.
constexpr int kP2pMaxWorld = 16;
constexpr unsigned long long kP2pWaitNs = 10000000000ull;  // 10 s: longest wait for a peer's record
struct P2pExchange {
  unsigned char *peers[kP2pMaxWorld];  // every rank's exchange block as mapped into this process (own block included)
  int rank, world;
  unsigned long long seq;              // this tick's sequence number (> 0, the same on every rank)
  size_t stride;                       // bytes per record slot
};
__host__ __device__ inline size_t p2p_slot_stride() { return (argmin_record_bytes(kMaxObstacles) + 127) & ~(size_t)127; }
__host__ __device__ inline size_t p2p_flags_offset() { return 2 * (size_t)kP2pMaxWorld * p2p_slot_stride(); }
__host__ __device__ inline size_t p2p_block_bytes() { return p2p_flags_offset() + 2 * kP2pMaxWorld * sizeof(unsigned long long); }

// out_status: 0 ok, 1 a peer's record never arrived (bounded wait)
// returns false (in every thread) when the exchange failed
__device__ __forceinline__ bool p2p_select_body(const unsigned char *local_rec, const P2pExchange &X, int n_obs,
                                                DeviceBest *best, double *best_random, EvalResult *out, HostOut *host,
                                                unsigned long long ticket, int *out_status) {
  __shared__ int s_fail;
  const int parity = (int)(X.seq & 1ull);
  const size_t bytes = argmin_record_bytes(n_obs);
  const size_t slot = ((size_t)parity * kP2pMaxWorld + X.rank) * X.stride;
  // 1. this rank's record into every rank's block (16-byte stores; the record is 8-byte aligned, sizes are multiples of 8)
  for (int r = 0; r < X.world; ++r) {
    double *dst = reinterpret_cast<double *>(X.peers[r] + slot);
    const double *src = reinterpret_cast<const double *>(local_rec);
    for (size_t i = threadIdx.x; i < bytes / sizeof(double); i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish (the publishing thread fences again after the barrier: cumulative over the block's stores)
  if (threadIdx.x < X.world) {
    volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(X.peers[threadIdx.x] + p2p_flags_offset()) +
                                        parity * kP2pMaxWorld + X.rank;
    __threadfence_system();
    *flag = X.seq;
  }
  // 3. wait for everyone's record of this tick in the own block
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  if (threadIdx.x < X.world) {
    volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(X.peers[X.rank] + p2p_flags_offset()) +
                                        parity * kP2pMaxWorld + threadIdx.x;
    // bounded by the global timer (a peer is gone: give up instead of hanging the GPU); the timer is read every
    // 1024 polls only, it costs more than the poll itself
    const unsigned long long t_start = global_timer_ns();
    unsigned spins = 0;
    while (*flag != X.seq) {
      if ((++spins & 1023u) == 0u && global_timer_ns() - t_start > kP2pWaitNs) {
        s_fail = 1;
        break;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (s_fail) {
    // A failed exchange is FATAL for the sharded planner group (pmaf.h, pmaf_p2p_import): the peers that did get
    // every record go on, so the replicas of the incumbent may differ from here on. The host makes the flag sticky.
    if (threadIdx.x == 0) {
      *out_status = 1;
      __threadfence_system();  // the status before the ticket
      if (host) host->seq[0] = ticket;  // wake the host; it reads out_status
    }
    return false;
  }
  // 4. replicated selection over the own block
  select_over_records(X.peers[X.rank] + (size_t)parity * kP2pMaxWorld * X.stride, X.stride, X.world, n_obs, best, best_random, out);
  __syncthreads();
  if (threadIdx.x == 0 && host) {
    host->eval.best_index = out->best_index, host->eval.argmin_index = out->argmin_index;
    host->eval.incumbent_changed = out->incumbent_changed, host->eval.pad = 0;
    host->eval.best_cost = out->best_cost, host->eval.argmin_cost = out->argmin_cost;
    host->best.present = best->present, host->best.id = best->id, host->best.type = best->type, host->best.pad = 0;
    __threadfence_system();
    host->seq[0] = ticket;
  }
  return true;
}
__global__ void __launch_bounds__(256) p2p_select_kernel(const unsigned char *local_rec, const P2pExchange X, int n_obs,
                                                         DeviceBest *best, double *best_random, EvalResult *out,
                                                         HostOut *host, unsigned long long ticket, int *out_status) {
  p2p_select_body(local_rec, X, n_obs, best, best_random, out, host, ticket, out_status);
}

// ---- moveRealEEAgent (cf_manager.cpp:257-263 -> RealCfAgent::cfPlanner, cf_agent.cpp:343-366) ------------------
struct RealArgs {
  RealState *real;
  unsigned char *known;   // [O] real agent's known_obstacles_
  double *rot;            // [O][3] real agent's field_rotation_vecs_
  const DeviceBest *best;
  const double *best_random;  // [O][3] incumbent's random vectors
  const double *obs_pos, *obs_vel, *obs_rad;  // live list
  int n_obs;
  double delta_t;
  int steps;
  int agent_id;           // GLOBAL index whose gains are used; -1: eval->best_index (fused tick)
  const EvalResult *eval;
  double *path_out;       // [steps][3] position after every step (RealCfAgent::setPosition appends)
  double goal[3];
  HostOut *host;          // zero-copy result + ticket (see HostOut), or null
  unsigned long long ticket;
  // fused tick: the evaluate results go out with the real agent's, under ONE system-scope fence
  const EvalResult *pub_eval;
  const DeviceBest *pub_best;
};

// one warp
__device__ __forceinline__ void real_agent_body(const PlannerDev &P, const RealArgs &R) {
  __shared__ double fbuf[3 * 32];
  const Group<32> g;
  LiveObstacles obs;
  obs.p = R.obs_pos, obs.v = R.obs_vel, obs.r = R.obs_rad, obs.agent_rad = P.rad;
  KnownBytes known;
  known.b = R.known;
  int aid = R.agent_id;
  if (aid < 0) aid = R.eval->best_index;
  const AgentConsts k = make_agent_consts(P.k_attr[aid], P.k_circ[aid], P.k_repel[aid], P.k_damp[aid], P.shell,
                                          P.vel_max, P.approach_dist, P.mass, obs.rsum(R.n_obs - 1));
  const int type = R.best->type;
  const v3 goal = ld3(R.goal);
  const v3 init_pos = ld3(R.real->init_pos);
  v3 p = ld3(R.real->pos), v = ld3(R.real->vel);
  v3 force = mk3(0.0, 0.0, 0.0);
  const int n_field = R.n_obs - 1;
  PMAF_T_DECL;
  for (int s = 0; s < R.steps; ++s) {
    force = mk3(0.0, 0.0, 0.0);
    const v3 goal_vec = sub3(goal, p);
    // scalar parts under FastMath (branch-free Newton refinements, bit-identical inside their proven range), the
    // IEEE built-ins when an operand leaves it
    ExactMath em;
    FastMath fm;
    StepNorms sn = step_norms(fm, goal_vec, v, 1.0, false, k);
    v3 ghat, nv_unused;
    step_units<false>(fm, goal_vec, v, sn, ghat, nv_unused);
    if (__builtin_expect(fm.bad(), 0)) {
      sn = step_norms(em, goal_vec, v, 1.0, false, k);
      step_units<false>(em, goal_vec, v, sn, ghat, nv_unused);
    }
    double k_goal_scale = 1.0;
    if (field_gate_open(sn.dist_goal, sn.vn, p, init_pos, k)) {
      double min_d, kgs_closest;
      bool has_closest;
      field_pass<false, false>(g, obs, n_field, nullptr, n_field, type, p, v, goal_vec, sn, nv_unused, goal, ghat, k, known,
                        R.rot, R.best_random, fbuf, 0.0, force, min_d, has_closest, kgs_closest PMAF_T_PASS);
      if (has_closest && norm_gt(dot3(force, force), make_thr(1e-5))) k_goal_scale = kgs_closest;
      g.sync();
    }
    {
      const v3 p0 = p, v0 = v, f0 = force;
      FastMath ff;
      finish_step(ff, force, k_goal_scale, sn, obs.pos(R.n_obs - 1), R.delta_t, k, p, v);
      if (__builtin_expect(ff.bad(), 0)) {
        p = p0, v = v0, force = f0;
        finish_step(em, force, k_goal_scale, sn, obs.pos(R.n_obs - 1), R.delta_t, k, p, v);
      }
    }
    if (g.gl == 0) {
      st3(R.path_out + 3 * s, p);
      if (R.host) st3(R.host->real_path + 3 * s, p);
    }
  }
  if (g.gl == 0 && R.steps > 0) {
    st3(R.real->pos, p), st3(R.real->vel, v), st3(R.real->force, force);
    if (R.host) {
      if (R.pub_eval) {
        const EvalResult e = *R.pub_eval;
        const DeviceBest b = *R.pub_best;
        R.host->eval.best_index = e.best_index, R.host->eval.argmin_index = e.argmin_index;
        R.host->eval.incumbent_changed = e.incumbent_changed, R.host->eval.pad = 0;
        R.host->eval.best_cost = e.best_cost, R.host->eval.argmin_cost = e.argmin_cost;
        R.host->best.present = b.present, R.host->best.id = b.id, R.host->best.type = b.type, R.host->best.pad = 0;
      }
      st3(R.host->real.pos, p), st3(R.host->real.vel, v), st3(R.host->real.force, force), st3(R.host->real.init_pos, init_pos);
      __threadfence_system();
      R.host->seq[1] = R.ticket;
    }
  }
}
__global__ void __launch_bounds__(32) real_agent_kernel(const PlannerDev P, const RealArgs R) { real_agent_body(P, R); }

// ---- the fused control tick ------------------------------------------------------------------------------------------------
// pmaf_tick's device chain in ONE CTA instead of three or four dependent launches (evaluate [+ exchange] ->
// real step -> reset): phase 1 evaluateAgents (cf_manager.cpp:293-356) — costs, serial-order argmin, for sharded
// planners the peer-memory exchange and the replicated selection; phase 2 warp 0 moves the real agent
// (cf_manager.cpp:257-263) while the other warps rebuild the obstacle image / nearest-neighbour table when the
// obstacle list changed and pack the real agent's known flags; the agents' own reset (resetEEAgents,
// cf_manager.cpp:246-255) happens in the prologue of the rollout that follows (PlannerDev::reset_in_prologue),
// from the real agent's state this kernel leaves behind.
// dynamic_obstacle_node's integration step (dynamic_obstacle_node.cpp:357) on the device-resident live list
__device__ __forceinline__ void feed_obstacles(double *live_pos, const double *live_vel, int n_feed, double frequency,
                                               int tid, int nthreads) {
  for (int i = tid; i < 3 * n_feed; i += nthreads) live_pos[i] += live_vel[i] / frequency;
}
__global__ void __launch_bounds__(256) feed_kernel(double *live_pos, const double *live_vel, int n_feed, double frequency) {
  feed_obstacles(live_pos, live_vel, n_feed, frequency, threadIdx.x, blockDim.x);
}

struct TickArgs {
  int feed_n;           // > 0: a pending obstacle feed (pmaf_feed_obstacles) is applied first
  double feed_frequency;
  double *feed_pos;
  const double *feed_vel;
  int eval_mode;        // 0: unsharded (finalize here), 1: local scan + peer-memory exchange + selection,
                        // 2: selection over all-gathered records (NCCL fallback; the local scan ran before)
  DeviceBest *best;
  double *best_random;
  ArgminRecord *rec;
  EvalResult *eval;
  HostOut *host;
  unsigned long long eval_ticket;
  P2pExchange xchg;     // eval_mode 1
  const unsigned char *rec_all;  // eval_mode 2
  int world;
  int *p2p_status;
  int rebuild_image;    // the obstacle list (or the broad-phase margin) changed since the image was built
  int rebuild_nn;
  uint32_t *known_bits, *known_keep;  // [known_words] packed flags of the real agent for the rollout's prologue
};
// at most 256 threads: the real agent's step (warp 0) needs the full register file — at 1024 threads per CTA the
// kernel is held to 64 registers and the step spills (measured: it is the longest phase of the kernel)
__global__ void __launch_bounds__(256, 1) tick_kernel(const PlannerDev P, const CostParams C, const TickArgs T,
                                                    const RealArgs R, const ResetArgs S) {
  asm volatile("griddepcontrol.launch_dependents;");  // the rollout may be scheduled now; it waits for this grid's end
  bool ok = true;
#if defined(PMAF_FAST_STATS)  // developer build: phase time stamps (ns) into step_counter[8..12]
#define PMAF_STAMP(k) do { if (threadIdx.x == 0) P.step_counter[8 + (k)] = global_timer_ns(); } while (0)
#else
#define PMAF_STAMP(k) do { } while (0)
#endif
  PMAF_STAMP(0);
  // a pending obstacle feed: before anything reads the live list (the barrier after the evaluate phase orders it)
  if (T.feed_n > 0) feed_obstacles(T.feed_pos, T.feed_vel, T.feed_n, T.feed_frequency, threadIdx.x, blockDim.x);
  {
    // Everything the dependent phases below will touch, into L2 now: after a cold start (or an L2 flush) every
    // dependent first touch would otherwise be a DRAM round trip on the tick's critical path. Warp 0 takes what
    // the real step needs, the other threads the agents' persistent rows the rollout's prologue reads.
    const int n3 = R.n_obs * 3 * (int)sizeof(double);
    if (threadIdx.x < 32) {
      prefetch_l2(R.real, sizeof(RealState), threadIdx.x, 32), prefetch_l2(R.best, sizeof(DeviceBest), threadIdx.x, 32);
      prefetch_l2(R.obs_pos, n3, threadIdx.x, 32), prefetch_l2(R.obs_vel, n3, threadIdx.x, 32);
      prefetch_l2(R.obs_rad, n3 / 3, threadIdx.x, 32), prefetch_l2(R.known, R.n_obs, threadIdx.x, 32);
      prefetch_l2(R.rot, n3, threadIdx.x, 32), prefetch_l2(R.best_random, n3, threadIdx.x, 32);
    } else {
      const int tid = threadIdx.x - 32, nthreads = blockDim.x - 32;
      const size_t cap = (size_t)1 << 20;  // large populations: the first MiB of each (the rest streams anyway)
      const size_t ng = (size_t)(P.first_agent + P.n_agents) * sizeof(double);
      prefetch_l2(P.k_attr, ng < cap ? ng : cap, tid, nthreads), prefetch_l2(P.k_circ, ng < cap ? ng : cap, tid, nthreads);
      prefetch_l2(P.k_repel, ng < cap ? ng : cap, tid, nthreads), prefetch_l2(P.k_damp, ng < cap ? ng : cap, tid, nthreads);
      const size_t ni = (size_t)P.n_agents * 3 * sizeof(double), nk = (size_t)P.n_agents * P.known_words * sizeof(uint32_t);
      prefetch_l2(P.init_pos, ni < cap ? ni : cap, tid, nthreads), prefetch_l2(P.known, nk < cap ? nk : cap, tid, nthreads);
      prefetch_l2(P.runtime_zero, sizeof(unsigned), tid, nthreads);
      if (!T.rebuild_image) prefetch_l2(S.image, P.img.bytes, tid, nthreads);
    }
  }
  if (T.eval_mode == 0) {
    evaluate_body(P, C, T.best, T.best_random, T.rec, T.eval, 1, nullptr, 0ull);
  } else if (T.eval_mode == 1) {
    evaluate_body(P, C, T.best, T.best_random, T.rec, T.eval, 0, nullptr, 0ull);
    __syncthreads();  // the record (global memory) is complete
    ok = p2p_select_body(reinterpret_cast<const unsigned char *>(T.rec), T.xchg, P.n_obs, T.best, T.best_random, T.eval,
                         nullptr, 0ull, T.p2p_status);
  } else {
    select_over_records(T.rec_all, argmin_record_bytes(P.n_obs), T.world, P.n_obs, T.best, T.best_random, T.eval);
  }
  __syncthreads();  // best / eval / best_random (global) are visible to the whole CTA
  PMAF_STAMP(1);
  if (!ok) {  // the real agent's ticket still has to arrive: the host waits for it, then reads the failure flag
    if (threadIdx.x == 0 && R.host) {
      __threadfence_system();  // the failure flag (p2p_select_body) before the ticket
      R.host->seq[1] = R.ticket;
    }
    return;
  }
  const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  if (warp == 0) {
    real_agent_body(P, R);
  } else {
    const int tid = threadIdx.x - 32, nthreads = blockDim.x - 32;
    if (T.rebuild_image) build_obstacle_image(P, S, tid, nthreads);
    if (T.rebuild_nn) build_nn_table(P, S, warp - 1, warps - 1);
  }
  if (blockDim.x == 32) {  // a one-warp CTA does everything in turn
    if (T.rebuild_image) build_obstacle_image(P, S, threadIdx.x, 32);
    if (T.rebuild_nn) build_nn_table(P, S, 0, 1);
  }
  PMAF_STAMP(2);
  __syncthreads();  // real_known may have gained flags in the real step
  // the real agent's known flags as bit words, one ballot per word (see pack_known_word)
  for (int w = warp; w < P.known_words; w += warps) {
    const int i = w * 32 + (threadIdx.x & 31);
    const bool in_list = i < S.n_obs_update;
    const unsigned bits = __ballot_sync(0xffffffffu, in_list && S.real_known[i] != 0);
    const unsigned keep = __ballot_sync(0xffffffffu, !in_list);
    if ((threadIdx.x & 31) == 0) T.known_bits[w] = bits, T.known_keep[w] = keep;
  }
  if (threadIdx.x == 0) P.step_counter[0] = 0ull;  // executed steps of the rollout that follows
  PMAF_STAMP(3);
#undef PMAF_STAMP
}

// The k cheapest agents of the last evaluate, in (cost, index) order — what the node's predicted-path
// visualisation needs (panda_bimanual_control.cpp:340-347 draws every path; only these are copied out).
// One block; k passes of the serial-order argmin, NaN costs last.
__global__ void __launch_bounds__(1024) topk_kernel(const PlannerDev P, int k, int *out_index) {
  __shared__ double s_cost[32];
  __shared__ int s_idx[32];
  __shared__ double s_last_cost;
  __shared__ int s_last_idx;
  if (threadIdx.x == 0) s_last_cost = -(double)INFINITY, s_last_idx = -1;
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    double bc = (double)INFINITY;
    int bi = 0x7fffffff;
    const double lc = s_last_cost;
    const int li = s_last_idx;
    for (int a = threadIdx.x; a < P.n_agents; a += blockDim.x) {
      double c = P.cost[a];
      if (!(c == c)) c = 1.7976931348623157e308;  // NaN costs sort last (they never win the argmin)
      const bool after = c > lc || (c == lc && a > li);
      if (after && (c < bc || (c == bc && a < bi))) bc = c, bi = a;
    }
    for (int off = 16; off > 0; off >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, bc, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (oc < bc || (oc == bc && oi < bi)) bc = oc, bi = oi;
    }
    if ((threadIdx.x & 31) == 0) s_cost[threadIdx.x >> 5] = bc, s_idx[threadIdx.x >> 5] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
      const int nw = (blockDim.x + 31) >> 5;
      for (int w = 1; w < nw; ++w)
        if (s_cost[w] < bc || (s_cost[w] == bc && s_idx[w] < bi)) bc = s_cost[w], bi = s_idx[w];
      out_index[r] = bi == 0x7fffffff ? -1 : bi;
      s_last_cost = bc, s_last_idx = bi;
    }
    __syncthreads();
  }
}

// fill rot[A][O][3] with the default (0,0,1) (cf_agent.h:92-96) and clear known bits
__global__ void init_state_kernel(const PlannerDev P, const double *init_pos) {
  const size_t n = (size_t)P.n_agents * P.n_obs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    P.rot[3 * i] = 0.0, P.rot[3 * i + 1] = 0.0, P.rot[3 * i + 2] = 1.0;
  }
  const size_t nk = (size_t)P.n_agents * P.known_words;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nk; i += (size_t)gridDim.x * blockDim.x)
    P.known[i] = 0u;
  for (size_t a = blockIdx.x * (size_t)blockDim.x + threadIdx.x; a < (size_t)P.n_agents;
       a += (size_t)gridDim.x * blockDim.x) {
    const v3 ip = ld3(init_pos);
    // CfAgent constructor (cf_agent.h:69-91): path = [agent_pos], vel = (0.01,0,0), init_pos = 0
    st3(P.cur_pos + 3 * a, ip);
    st3(P.paths + a * P.max_steps * 3, ip);
    st3(P.vel + 3 * a, mk3(0.01, 0.0, 0.0));
    st3(P.init_pos + 3 * a, mk3(0.0, 0.0, 0.0));
    P.min_obs_dist[a] = P.shell;
    P.path_len[a] = 0.0;
    P.ws_cost[a] = 0.0;
    P.pred_time_ns[a] = 0.0;
    P.n_path[a] = 1;
    P.reached[a] = 0;
    P.cost[a] = 0.0;
  }
}

// setInitialPosition (cf_manager.cpp:226-236): agents' init_pos := pos, path := [pos]
__global__ void set_initial_position_kernel(const PlannerDev P, const double *pos6) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.n_agents) return;
  const v3 pos = ld3(pos6);
  st3(P.init_pos + 3 * a, pos);
  st3(P.cur_pos + 3 * a, pos);
  st3(P.paths + (size_t)a * P.max_steps * 3, pos);
  P.n_path[a] = 1;
  P.path_len[a] = 0.0;
  P.ws_cost[a] = 0.0;
}

}  // namespace pmaf
