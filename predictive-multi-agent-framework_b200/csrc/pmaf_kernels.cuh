// pmaf_kernels.cuh — sm_100a kernels of the multi-agent predictive rollout.
//
// Data-parallel decomposition: one GROUP of LPA lanes (4..32, a whole warp by
// default) owns one agent for its whole horizon — the horizon is a strict recurrence, the agents
// are independent (cf_manager.cpp:118-123). Inside a step the O(O) obstacle work is spread over
// the group's lanes:
//
//   broad phase  (fp32, all O-1 field obstacles): conservative sphere test "may this obstacle be
//                inside the detection shell?" against obstacle records staged in shared memory by
//                one TMA bulk copy per CTA. An obstacle outside the shell is a strict no-op for the
//                reference step (it cannot lower min_obs_dist_, which starts at the shell radius,
//                adds a zero force, and is ignored by attractorForceScaling), so only candidates
//                go on. Candidates are compacted, in obstacle order, into a per-group list.
//   narrow phase (fp64, candidates only, one lane per candidate): the reference's circForce body
//                (cf_agent.cpp:76-106) bit for bit; forces are then summed in obstacle-index order
//                (contributions staged in shared memory by rank, serial fp64 adds front to back) so
//                that the sum has the reference's rounding; min distance and the closest obstacle are
//                redux.sync reductions on the integer order of non-negative doubles (exact: min /
//                lexicographic min), xor-shuffle butterflies for groups narrower than a warp.
//   scalar part  (gate, repulsion from the sentinel, attraction, integrator, cost accumulators):
//                replicated in every lane of the group.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmaf_math.cuh"

namespace pmaf {

constexpr int kMaxObstacles = 4096;  // candidate indices are uint16; the smem image must fit 227 KB

// The step logic below (field_pass, agent_step) is written against a "group" policy so that the
// SAME source runs on the GPU (Group<LPA>: sub-warp collectives) and, for CPU verification of the
// operation order only, on the host (HostGroup: one lane, trivial collectives; used by
// tests/host_step_check.cu, never by the product).
#if defined(__CUDA_ARCH__)
#define PMAF_FFS(x) __ffs(x)
#define PMAF_POPC(x) __popc(x)
#else
#define PMAF_FFS(x) __builtin_ffs(x)
#define PMAF_POPC(x) __builtin_popcount(x)
#endif
#define PMAF_HDT __host__ __device__ __forceinline__

// Optional per-section cycle counters (build with -DPMAF_SECTION_TIMERS: tools/section_timers.py).
#if defined(PMAF_SECTION_TIMERS) && defined(__CUDA_ARCH__)
#define PMAF_T(k)                                            \
  do {                                                       \
    const long long t_now_ = clock64();                      \
    pmaf_sec_[k] += t_now_ - pmaf_t_;                        \
    pmaf_t_ = t_now_;                                        \
  } while (0)
#define PMAF_T_DECL long long pmaf_sec_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long pmaf_t_ = clock64()
#define PMAF_T_ARGS , long long *pmaf_sec_, long long &pmaf_t_
#define PMAF_T_PASS , pmaf_sec_, pmaf_t_
#else
#define PMAF_T(k) do { } while (0)
#define PMAF_T_DECL do { } while (0)
#define PMAF_T_ARGS
#define PMAF_T_PASS
#endif

// Cost parameters of CfManager::evaluateAgents (cf_manager.cpp:293-297)
struct CostParams {
  double k_goal_dist, k_path_len, k_safe_dist, k_workspace;
  double ws[6];
};

// Shared-memory image of the obstacle set, identical layout in the global staging buffer that the
// CTA copies with one cp.async.bulk (offsets in bytes from the image base, all 16-byte aligned).
struct ObstacleImage {
  int n_obs;         // O, including the trailing sentinel
  int dynamic;       // any obstacle velocity != 0: positions advance every step
  uint32_t off_px, off_py, off_pz;  // double[O]  exact positions at the current rollout step
  uint32_t off_rs;                  // double[O]  rad_ + obstacle radius (cf_agent.cpp:84)
  uint32_t off_vx, off_vy, off_vz;  // double[O]  velocities           (dynamic only)
  uint32_t off_dx, off_dy, off_dz;  // double[O]  velocity * dt        (dynamic only, :273)
  uint32_t off_bp;                  // float4[O]  broad phase: x, y, z, (shell + rs + margin)^2
  uint32_t off_nn;                  // uint16[O]  nearest other field obstacle of each obstacle (:434-446), static
                                    //            scenes with OBSTACLE / GOAL_OBSTACLE agents only (nn_valid)
  int nn_valid;
  uint32_t bytes;                   // image size (multiple of 16)
};

struct DeviceBest {  // what RealCfAgent needs from *best_agent_ (cf_agent.cpp:368-387)
  int present;       // best_agent_ != nullptr
  int id;            // getAgentID() = global index + 1
  int type;
  int pad;
};

struct RealState {  // RealCfAgent scalar state (cf_agent.h:36-41)
  double pos[3], vel[3], force[3], init_pos[3];
};

struct EvalResult {
  int best_index;      // value returned by evaluateAgents (global index)
  int argmin_index;    // serial argmin before hysteresis
  int incumbent_changed;
  int pad;
  double best_cost, argmin_cost;
};

// Everything a tick reads back lives in ONE device block with the layout of its pinned host mirror, so
// that each call's results come back in a single small copy (eval + best; real + its path; all three).
struct HostOut {
  EvalResult eval;
  DeviceBest best;
  RealState real;
  double real_path[256 * 3];
  unsigned long long steps[16];
  // Host mirror only (pinned, mapped): the small kernels of a tick store their results straight into it and
  // then publish a ticket; the host polls the ticket instead of enqueueing a copy and synchronising the
  // stream (a launch + copy + sync round trip costs 30-50 us, two of them per call-by-call tick).
  volatile unsigned long long seq[2];  // [0] evaluate, [1] real agent
  volatile int p2p_fail;               // set by p2p_select_kernel when a peer's record never arrived
};

struct PlannerDev {  // kernel argument, passed by value
  int n_agents;      // local agents
  int first_agent;   // global index of local agent 0
  int n_obs;
  int max_steps;     // H: cap on path points (cf_agent.cpp:311)
  double goal[3];
  double shell, mass, rad, vel_max, approach_dist;
  double pred_dt;    // prediction_freq_multiple * delta_t
  // per-agent arrays [n_agents]
  const double *k_attr, *k_circ, *k_repel, *k_damp;
  double *init_pos;      // [A][3]
  double *cur_pos;       // [A][3]  latest path point
  double *vel;           // [A][3]
  double *min_obs_dist;  // [A]
  double *path_len;      // [A]   running getPathLength()
  double *ws_cost;       // [A]   running workspace cost under fused_cost
  double *pred_time_ns;  // [A]
  int *n_path;           // [A]
  int *reached;          // [A]
  double *cost;          // [A]
  double *paths;         // [A][H][3]
  // per-(agent, obstacle)
  double *rot;           // [A][O][3] field_rotation_vecs_
  const double *random_vecs;  // [A][O][3] (RANDOM agents only)
  uint32_t *known;       // [A][KW] bit i = known_obstacles_[i]
  int known_words;       // KW = ceil(O / 32)
  // obstacles
  const unsigned char *image;  // staging image (global)
  ObstacleImage img;
  // fused cost accumulation
  CostParams fused_cost;
  int fused_valid;
  // outputs
  unsigned long long *step_counter;  // [0] executed integration steps of this rollout, [1] running total
  long long *section_cycles;         // [64][12] per-section cycle counters (PMAF_SECTION_TIMERS builds)
  const unsigned *runtime_zero;      // one word holding 0 (see keep() in pmaf_math.cuh)
  // fused tick: resetEEAgents (cf_manager.cpp:246-255) happens in the rollout's prologue, from the real agent's
  // state and packed known flags that tick_kernel left behind — no separate pass over the agents
  int reset_in_prologue;
  const RealState *reset_real;
  const uint32_t *reset_known_bits, *reset_known_keep;  // [known_words]
};

// ---- small PTX wrappers (TMA bulk copy + mbarrier) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a bulk copy that never lands traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 26)) __trap();
  }
}
// prefetch [ptr, ptr + bytes) into L2, one 128-byte line per participating thread and round
__device__ __forceinline__ void prefetch_l2(const void *ptr, size_t bytes, int tid, int nthreads) {
  const char *c = reinterpret_cast<const char *>(ptr);
  for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)nthreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- group (sub-warp) collectives ------------------------------------------------------------------------
template <int LPA>
struct Group {
  static constexpr int kLanes = LPA;
  unsigned mask;  // lanes of this group inside the warp
  int lane0;      // first lane of the group
  int gl;         // lane index inside the group
  int lane;       // lane index inside the warp
  __device__ __forceinline__ Group() {
    lane = threadIdx.x & 31;
    gl = lane % LPA;
    lane0 = lane - gl;
    mask = LPA == 32 ? 0xffffffffu : (((1u << LPA) - 1u) << lane0);
  }
  __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(mask, p); }
  __device__ __forceinline__ double bcast(double v, int src_lane) const { return __shfl_sync(mask, v, src_lane); }
  __device__ __forceinline__ int bcast(int v, int src_lane) const { return __shfl_sync(mask, v, src_lane); }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  __device__ __forceinline__ double min_reduce(double v) const {  // NaN never wins (a < b false)
#pragma unroll
    for (int off = LPA / 2; off > 0; off >>= 1) {
      double o = __shfl_xor_sync(mask, v, off);
      v = o < v ? o : v;
    }
    return v;
  }
  // minimum of NON-NEGATIVE doubles (or +inf): their bit patterns order like unsigned integers, so
  // two redux.sync.min (high word, then low word among the lanes holding the smallest high word)
  // replace a five-level shuffle tree. Returns the minimum in every lane.
  __device__ __forceinline__ double min_reduce_nonneg(double v) const {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mhi = __reduce_min_sync(mask, hi);
    const unsigned mlo = __reduce_min_sync(mask, hi == mhi ? lo : 0xffffffffu);
    return __hiloint2double((int)mhi, (int)mlo);
  }
  // lexicographic (value, index) minimum for non-negative values, same idea; idx >= 0
  __device__ __forceinline__ void argmin_reduce_nonneg(double &v, int &idx) const {
    const double m = min_reduce_nonneg(v);
    const unsigned mi = __reduce_min_sync(mask, v == m ? (unsigned)idx : 0xffffffffu);
    v = m, idx = (int)mi;
  }
  // ---- warp-convergent variants ("_w") ----------------------------------------------------------------------
  // A collective with a sub-warp member mask that differs between the lanes of a warp (2 / 4 agents per warp)
  // compiles to MATCH.ANY + a divergent loop over the distinct masks: every group is served on its own and
  // the warp's instruction stream is issued once per group. The _w variants are executed by ALL 32 lanes of
  // the warp together (every group at once, constant full mask) and hand each group its own part of the
  // result; ballots are returned GROUP-RELATIVE (bit j = lane j of this group). The straight-line step of the
  // packed shapes (fast_step_packed) uses nothing else; for LPA == 32 they are the plain collectives.
  __device__ __forceinline__ unsigned ballot_w(bool p) const {
    const unsigned m = __ballot_sync(0xffffffffu, p);
    return LPA == 32 ? m : (m >> lane0) & ((1u << (LPA & 31)) - 1u);
  }
  __device__ __forceinline__ bool any_w(bool p) const { return __any_sync(0xffffffffu, p) != 0; }  // the whole WARP
  __device__ __forceinline__ void sync_w() const { __syncwarp(); }
  __device__ __forceinline__ double bcast_w(double v, int group_lane) const { return __shfl_sync(0xffffffffu, v, lane0 + group_lane); }
  __device__ __forceinline__ int bcast_w(int v, int group_lane) const { return __shfl_sync(0xffffffffu, v, lane0 + group_lane); }
  // minimum of non-negative doubles (or +inf) over the group: redux for a whole warp, xor butterflies
  // (which stay inside aligned groups) otherwise
  __device__ __forceinline__ double min_w_nonneg(double v) const {
    if (LPA == 32) return min_reduce_nonneg(v);
#pragma unroll
    for (int off = LPA / 2; off > 0; off >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, v, off);
      v = o < v ? o : v;
    }
    return v;
  }
  __device__ __forceinline__ void argmin_w_nonneg(double &v, int &idx) const {
    if (LPA == 32) {
      argmin_reduce_nonneg(v, idx);
      return;
    }
#pragma unroll
    for (int off = LPA / 2; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, off);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
      if (ov < v || (ov == v && oi < idx)) v = ov, idx = oi;
    }
  }
  // lexicographic (value, index) minimum: smallest value, lowest index among equals
  __device__ __forceinline__ void argmin_reduce(double &v, int &idx) const {
#pragma unroll
    for (int off = LPA / 2; off > 0; off >>= 1) {
      double ov = __shfl_xor_sync(mask, v, off);
      int oi = __shfl_xor_sync(mask, idx, off);
      if (ov < v || (ov == v && oi < idx)) v = ov, idx = oi;
    }
  }
};

// single-lane stand-in with the same interface (host verification of the operation order)
struct HostGroup {
  static constexpr int kLanes = 1;
  unsigned mask = 1u;
  int lane0 = 0, gl = 0, lane = 0;
  PMAF_HDT unsigned ballot(bool p) const { return p ? 1u : 0u; }
  PMAF_HDT double bcast(double v, int) const { return v; }
  PMAF_HDT int bcast(int v, int) const { return v; }
  PMAF_HDT void sync() const {}
  PMAF_HDT double min_reduce(double v) const { return v; }
  PMAF_HDT void argmin_reduce(double &, int &) const {}
  PMAF_HDT double min_reduce_nonneg(double v) const { return v; }
  PMAF_HDT void argmin_reduce_nonneg(double &, int &) const {}
};

// ---- obstacle accessors ---------------------------------------------------------------------------------
struct SmemObstacles {  // rollout: the CTA's shared-memory image
  const double *px, *py, *pz, *rs, *vx, *vy, *vz;
  bool dynamic;
  PMAF_HDT v3 pos(int i) const { return mk3(px[i], py[i], pz[i]); }
  PMAF_HDT v3 vel(int i) const { return dynamic ? mk3(vx[i], vy[i], vz[i]) : mk3(0.0, 0.0, 0.0); }
  PMAF_HDT double rsum(int i) const { return rs[i]; }
};
struct LiveObstacles {  // real agent: the live list passed to moveRealEEAgent, in global memory
  const double *p, *v, *r;
  double agent_rad;
  PMAF_HDT v3 pos(int i) const { return ld3(p + 3 * i); }
  PMAF_HDT v3 vel(int i) const { return ld3(v + 3 * i); }
  PMAF_HDT double rsum(int i) const { return agent_rad + r[i]; }
};

// known_obstacles_ accessors
struct KnownBits {  // bit mask in shared memory (one row per group)
  uint32_t *w;
  PMAF_HDT bool test(int i) const { return (w[i >> 5] >> (i & 31)) & 1u; }
  PMAF_HDT void set(int i) const {
#if defined(__CUDA_ARCH__)
    atomicOr(&w[i >> 5], 1u << (i & 31));  // lanes of a group may share a word
#else
    w[i >> 5] |= 1u << (i & 31);
#endif
  }
};
struct KnownBytes {  // the real agent's flags in global memory
  unsigned char *b;
  PMAF_HDT bool test(int i) const { return b[i] != 0; }
  PMAF_HDT void set(int i) const { b[i] = 1; }
};

// nearest other field obstacle to obstacle `id` (cf_agent.cpp:434-446): serial scan semantics —
// strict '>' from index 0, start value 100.0, default index 0 — evaluated cooperatively.
#pragma nv_exec_check_disable
template <class G, class Obs>
#if defined(__CUDA_ARCH__)
__device__ __noinline__  // only two agents of a population (OBSTACLE / GOAL_OBSTACLE) ever call it
#else
inline
#endif
    int
    nearest_other_obstacle(const G &g, const Obs &obs, int n_field, int id) {
  constexpr int LPA = G::kLanes;
  v3 oi = obs.pos(id);
  double best = 100.0;
  int best_i = 0x7fffffff;
  for (int i = g.gl; i < n_field; i += LPA) {
    if (i != id) {
      double d = norm3(sub3(oi, obs.pos(i)));
      if (best > d) best = d, best_i = i;
    }
  }
  g.argmin_reduce(best, best_i);
  return best_i == 0x7fffffff ? 0 : best_i;
}

// calculateRotationVector of the agent `type` at first detection (cf_agent.cpp:408-611). Cold code: kept out
// of the step loop's instruction footprint.
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
    v3
    first_rotation_vector(int type, v3 p, v3 goal, v3 to_obs, v3 oi, v3 o_nn, const double *random_i) {
  switch (type) {
    case HAD_HEURISTIC: return rot_had(p, goal, oi);
    case RANDOM_AGENT: return rot_random(p, goal, ld3(random_i));
    case OBSTACLE_HEURISTIC: return rot_obstacle(to_obs, oi, o_nn);
    case GOAL_OBSTACLE_HEURISTIC: return rot_goal_obstacle(p, goal, to_obs, oi, o_nn);
    default: return mk3(0.0, 0.0, 1.0);  // GOAL :408-412, VEL :539-543
  }
}

// What one lane learns about its candidate in one chunk of the narrow phase; nothing is written to
// memory until the whole group's evaluation is known to be in the arithmetic policy's proven range.
struct CandEval {
  v3 f, rot_i;
  double d;        // dist_obs (clamped), +inf for an idle lane
  double kgs;      // attractorForceScaling's value IF this obstacle turns out to be the closest one
  bool counts_min, in_shell, first_seen, contributes;
};

// circForce loop body for obstacle i (cf_agent.cpp:76-105), pure: reads p, v, the obstacle, its known
// flag and stored rotation vector; returns the force term and what has to be committed.
//   STATIC_VEL: every obstacle velocity is zero, so rel_vel == v for all of them (v - 0 is v, bit
//   for bit); zv = v.v, vn = sqrt(zv) and nv_static = v / vn come from the caller, once per step.
//   ghat = normalized(goal - p) from the caller (it already holds |goal - p|).
// Every in-shell lane also evaluates the tail of attractorForceScaling (:212-226) for ITS obstacle:
// the chain (sqrt, division, exp; sqrt, division) is independent of the current-vector chain, so the
// two interleave, and after the closest-obstacle reduction the winner's value is just broadcast.
//   SPEC: evaluate the scaling speculatively per lane (latency build); otherwise the caller evaluates it
//   once for the winner after the reduction (throughput builds: fewer instructions).
#pragma nv_exec_check_disable
template <bool STATIC_VEL, bool SPEC, class M, class G, class Obs, class Known>
PMAF_HDT CandEval eval_candidate(M &m, const G &g, const Obs &obs, int n_field, bool active, int i, int type, v3 p,
                                 v3 v, v3 goal_vec, const StepNorms &sn, v3 nv_static, v3 goal, v3 ghat,
                                 const AgentConsts &c, const Known &known, const double *rot_row,
                                 const double *random_row) {
  CandEval r;
  r.f = mk3(0.0, 0.0, 0.0), r.rot_i = r.f, r.d = (double)INFINITY, r.kgs = 1.0;
  r.counts_min = r.in_shell = r.first_seen = r.contributes = false;
  v3 to_obs = r.f, rel = r.f, oi = r.f;
  bool is_known = false;
  if (active) {
    oi = obs.pos(i);
    is_known = known.test(i);
    if (is_known) r.rot_i = ld3(rot_row + 3 * i);  // issued early: the latency hides behind the sqrt / divisions
    const v3 rov = sub3(oi, p);
    rel = STATIC_VEL ? v : sub3(v, obs.vel(i));
    const double z = dot3(rov, rov);
    const double n = m.sqrt_(z);
    to_obs = normalized_zn_m(m, rov, z, n);
    r.d = clamp_dist(n - obs.rsum(i));
    const bool skip = dot3(to_obs, ghat) < -0.01 && dot3(rov, rel) < -0.01;  // :79-82
    if (!skip) {
      r.counts_min = true;           // :86-88
      r.in_shell = r.d < c.shell;    // :91
      r.first_seen = r.in_shell && !is_known;
    }
  }
  int nn = 0;
  if (type == OBSTACLE_HEURISTIC || type == GOAL_OBSTACLE_HEURISTIC) {
    // group-uniform branch: cooperative nearest-neighbour scans, one per newly detected obstacle
    unsigned todo = g.ballot(r.first_seen);
    while (todo) {
      const int src = PMAF_FFS(todo) - 1;
      todo &= todo - 1;
      const int id = g.bcast(i, src);
      const int found = nearest_other_obstacle(g, obs, n_field, id);
      if (g.lane == src) nn = found;
    }
  }
  if (r.d < c.shell) {  // candidates of the closest-obstacle search (:201-211), skipped ones included
    if (__builtin_expect(r.first_seen, 0)) {  // :92-96 (rare, out of line: built-in arithmetic)
      r.rot_i = first_rotation_vector(type, p, goal, to_obs, oi, obs.pos(nn), random_row + 3 * i);
    } else if (!is_known) {
      r.rot_i = mk3(0.0, 0.0, 1.0);  // skipped obstacle: its force is discarded below
    }
    // one straight-line block: scaling chain and force chain are independent
    if (SPEC) r.kgs = attractor_scaling(m, goal_vec, sn.dist_goal, p, v, sn.vn, c, r.d, oi);
    const double zr = STATIC_VEL ? sn.zv : dot3(rel, rel);
    const double vel_norm = STATIC_VEL ? sn.vn : m.sqrt_(zr);
    const v3 nv = STATIC_VEL ? nv_static : m.div3_(rel, vel_norm);
    const v3 nv_eigen = zr > 0.0 ? nv : rel;
    const v3 current = current_vector(m, type, p, goal, to_obs, nv_eigen, r.rot_i);
    const v3 f = circ_force_term(m, c.k_circ, r.d, nv, current);
    if (r.in_shell && vel_norm != 0) {  // :91, :98
      r.f = f;
      r.contributes = true;
    }
  }
  return r;
}

// the same evaluation with the built-in IEEE operations, out of line: taken only when some lane's
// operands fall outside FastMath's range
#pragma nv_exec_check_disable
template <bool STATIC_VEL, bool SPEC, class G, class Obs, class Known>
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
    CandEval
    eval_candidate_exact(const G &g, const Obs &obs, int n_field, bool active, int i, int type, v3 p, v3 v,
                         v3 goal_vec, const StepNorms &sn, v3 nv_static, v3 goal, v3 ghat, const AgentConsts &c,
                         const Known &known, const double *rot_row, const double *random_row) {
  ExactMath em;
  return eval_candidate<STATIC_VEL, SPEC>(em, g, obs, n_field, active, i, type, p, v, goal_vec, sn, nv_static, goal, ghat, c,
                                    known, rot_row, random_row);
}

// CfAgent::circForce / RealCfAgent::circForce (cf_agent.cpp:72-144) over a candidate list, plus the
// closest-obstacle search of attractorForceScaling (:199-211).
//   cand == nullptr: candidates are 0..n_cand-1 themselves.
//   fbuf: per-group staging buffer, 3 * kLanes doubles (shared memory on the GPU).
//   outputs (identical in every lane of the group): force = sum of curr_force in obstacle order,
//   min_d = min over non-skipped candidates of dist_obs (+inf if none), has_closest / kgs_closest =
//   whether an obstacle lies within the shell, and attractorForceScaling's value for the first one
//   with the smallest dist_obs.
#pragma nv_exec_check_disable
template <bool STATIC_VEL, bool SPEC, class G, class Obs, class Known>
PMAF_HDT void field_pass(const G &g, const Obs &obs, int n_field, const uint16_t *cand, int n_cand, int type, v3 p,
                         v3 v, v3 goal_vec, const StepNorms &sn, v3 nv_static, v3 goal, v3 ghat,
                         const AgentConsts &c, const Known &known, double *rot_row, const double *random_row,
                         double *fbuf, double min_obs, v3 &force, double &min_d, bool &has_closest,
                         double &kgs_closest PMAF_T_ARGS) {
  constexpr int LPA = G::kLanes;
  force = mk3(0.0, 0.0, 0.0);
  double lmin = (double)INFINITY;
  double lcd = c.shell, lkgs = 1.0;
  int lci = 0x7fffffff;
  const unsigned lt_mask = g.mask & ((1u << g.lane) - 1u);

  for (int c0 = 0; c0 < n_cand; c0 += LPA) {
    const int ci = c0 + g.gl;
    const bool active = ci < n_cand;
    const int i = active ? (cand ? (int)cand[ci] : ci) : 0;
    FastMath fm;
    CandEval r = eval_candidate<STATIC_VEL, SPEC>(fm, g, obs, n_field, active, i, type, p, v, goal_vec, sn, nv_static, goal,
                                            ghat, c, known, rot_row, random_row);
    if (__builtin_expect(g.ballot(fm.bad()) != 0u, 0)) {
      // the out-of-line call takes the norms by reference: hand it a copy made on this cold path, so that
      // the caller's prologue values stay in registers instead of being written to the stack every step
      const StepNorms sn_copy = sn;
      r = eval_candidate_exact<STATIC_VEL, SPEC>(g, obs, n_field, active, i, type, p, v, goal_vec, sn_copy, nv_static, goal,
                                                 ghat, c, known, rot_row, random_row);
    }
    PMAF_T(3);
    // ---- commit ----
    if (r.d < lcd) lcd = r.d, lci = i, lkgs = r.kgs;  // the search ignores the skip test (:201-211)
    if (r.counts_min && r.d < lmin) lmin = r.d;
    if (r.first_seen) {
      st3(rot_row + 3 * i, r.rot_i);
      known.set(i);
    }
    // force_ += curr_force in obstacle order (:106). Candidates are sorted and lanes are in list order,
    // so the contributions are compacted into the staging buffer by rank and every lane then adds them
    // up front to back: a chain of dependent adds fed by broadcast loads.
    const unsigned contrib = g.ballot(r.contributes);
    if (contrib) {
      if (r.contributes) {
        const int rank = PMAF_POPC(contrib & lt_mask);
        fbuf[3 * rank] = r.f.x, fbuf[3 * rank + 1] = r.f.y, fbuf[3 * rank + 2] = r.f.z;
      }
      g.sync();
      const int n_contrib = PMAF_POPC(contrib);
      // software-pipelined by two: the loads of the next pair are in flight while the current pair is added
      // (the adds are serially dependent, the loads are not)
      v3 a = mk3(fbuf[0], fbuf[1], fbuf[2]);
      v3 b = n_contrib > 1 ? mk3(fbuf[3], fbuf[4], fbuf[5]) : mk3(0.0, 0.0, 0.0);
      int j = 0;
      for (; j + 2 <= n_contrib; j += 2) {
        const int j2 = j + 2 < n_contrib ? j + 2 : 0, j3 = j + 3 < n_contrib ? j + 3 : 0;
        const v3 na = mk3(fbuf[3 * j2], fbuf[3 * j2 + 1], fbuf[3 * j2 + 2]);
        const v3 nb = mk3(fbuf[3 * j3], fbuf[3 * j3 + 1], fbuf[3 * j3 + 2]);
        force = add3(force, a);
        force = add3(force, b);
        a = na, b = nb;
      }
      if (j < n_contrib) force = add3(force, a);
      g.sync();
    }
    PMAF_T(4);
  }
  // reductions only when they can matter. All keys are non-negative (dist_obs >= 1e-5; a lane without a
  // close obstacle holds the shell radius, and has_closest implies shell > 1e-5), so the integer-ordered
  // redux reductions apply.
  // min_obs = the agent's running minimum: the reduction is skipped when no lane can lower it
  if (g.ballot(lmin < min_obs)) min_d = g.min_reduce_nonneg(lmin);
  else min_d = (double)INFINITY;
  has_closest = g.ballot(lci != 0x7fffffff) != 0u;
  kgs_closest = 1.0;
  if (has_closest) {
    const int mine = lci;
    g.argmin_reduce_nonneg(lcd, lci);
    if (SPEC) {  // the lane that evaluated the winning obstacle holds its scaling value
      const unsigned who = g.ballot(mine == lci);
      kgs_closest = g.bcast(lkgs, PMAF_FFS(who) - 1);
    } else {  // attractorForceScaling's tail (:212-226) once, for the winner
      const v3 o_c = obs.pos(lci);
      FastMath fm;
      kgs_closest = attractor_scaling(fm, goal_vec, sn.dist_goal, p, v, sn.vn, c, lcd, o_c);
      if (__builtin_expect(fm.bad(), 0)) {
        ExactMath em;
        kgs_closest = attractor_scaling(em, goal_vec, sn.dist_goal, p, v, sn.vn, c, lcd, o_c);
      }
    }
  }
  PMAF_T(5);
}

}  // namespace pmaf
