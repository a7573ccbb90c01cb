/* pmaf.h — C ABI of libpmaf.so, the B200 (sm_100a) implementation of the multi-agent predictive
 * rollout of riddhiman13/predictive-multi-agent-framework.
 *
 * The reference has no FFI for this path: the planner node holds a
 * ghostplanner::cfplanner::CfManager by value (panda_bimanual_control.h:45) and calls its C++
 * methods. This header is the boundary a drop-in replacement binds: one entry point per
 * CfManager method the node calls, plain pointers and sizes only. The C++ façade
 * include/pmaf/cf_manager.hpp maps the reference's own class onto it 1:1; INTEGRATION.md shows
 * the binding a maintainer adds.
 *
 * Citations are file:line under /root/reference/src/bimanual_planning_ros/
 * (h = include/bimanual_planning_ros/cf_manager.h, cpp = src/cf_manager.cpp,
 *  node = src/panda_bimanual_control.cpp).
 *
 * Conventions
 *   - every function returns 0 on success, a negative pmaf_status on failure; the message of the
 *     last failure on the calling thread is pmaf_last_error(). No exception crosses this ABI
 *     (the reference asserts / throws std::out_of_range: cpp:50-51, cf_agent.cpp:66-68).
 *   - vectors are double[3]; obstacle lists are three arrays pos[n][3], vel[n][3], rad[n], the
 *     layout of Obstacles.msg (msg/Obstacles.msg:1-3). The LAST obstacle is the self-collision
 *     sentinel (cf_agent.cpp:164-180). Agent indices are 0-based (the reference's ids are
 *     index + 1, cpp:70-104).
 *   - the caller owns every input and output buffer; inputs are copied during the call.
 *     The library owns device memory and all persistent per-agent state for the life of the
 *     handle. A handle is not thread-safe (all CfManager calls come from the single ros::spin()
 *     thread, node:475).
 *   - all arithmetic is IEEE binary64 in the reference's operation order; there is no CPU
 *     fallback: every entry point that computes fails with PMAF_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef PMAF_H
#define PMAF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmaf_planner pmaf_planner;

typedef enum {
  PMAF_OK = 0,
  PMAF_ERR_ARG = -1,    /* null pointer, size mismatch, index out of range */
  PMAF_ERR_STATE = -2,  /* call order: e.g. move_real_agent before any evaluate_agents */
  PMAF_ERR_CUDA = -3,   /* CUDA runtime / device error */
  PMAF_ERR_NCCL = -4,   /* best-agent exchange failed: NCCL error, or a peer's record never arrived (sharded planners only) */
  PMAF_ERR_ALLOC = -5
} pmaf_status;

/* CfAgent::Type (cf_agent.h:59-68) as returned by pmaf_get_best_agent_type / summaries */
enum {
  PMAF_REAL_AGENT = 0,
  PMAF_GOAL_HEURISTIC = 1,
  PMAF_OBSTACLE_HEURISTIC = 2,
  PMAF_GOAL_OBSTACLE_HEURISTIC = 3,
  PMAF_VEL_HEURISTIC = 4,
  PMAF_RANDOM_AGENT = 5,
  PMAF_HAD_HEURISTIC = 6,
  PMAF_UNDEFINED = 7
};

const char *pmaf_last_error(void);
/* library / build identification: "pmaf <version> sm_100a <build flags>" */
const char *pmaf_version(void);

/* ---- lifecycle -------------------------------------------------------------------------------- */

/* CfManager() = default (h:50). `device` is the CUDA ordinal the handle's buffers, stream and
 * kernels live on. */
int pmaf_create(pmaf_planner **out, int device);

/* ~CfManager() (h:51): waits for a running rollout, frees device memory. */
int pmaf_destroy(pmaf_planner *p);

/* Optional, before pmaf_init: this handle owns global agents [first_agent, first_agent+n_local)
 * of n_global (contiguous blocks per rank). Agent types follow the GLOBAL index
 * (cpp:70-104). With n_global == n_local (default) the planner is unsharded. */
int pmaf_set_shard(pmaf_planner *p, int n_global, int first_agent, int rank, int world);

/* Sharded planners exchange ONE record per rank per evaluate (per control tick): (min cost, argmin
 * index, incumbent cost) plus the random vectors of the rank's local winner, 40 + 24*n_obs bytes,
 * followed by a replicated serial selection — bit-identical to the reference's serial argmin over
 * the whole population. Transport: NVLink peer memory (pmaf_p2p_export / pmaf_p2p_import below) or
 * an NCCL all-gather. The NCCL communicator is either created here
 * from an id made on one rank and distributed by the host (torch.distributed, MPI, ...), or
 * attached from outside (ncclComm_t as void*). libnccl.so.2 is dlopen()ed on first use. */
int pmaf_nccl_unique_id(unsigned char id_out[128]);
int pmaf_nccl_init(pmaf_planner *p, const unsigned char id[128], int rank, int world);
int pmaf_set_nccl_comm(pmaf_planner *p, void *nccl_comm);
/* Best-agent exchange over NVLink peer memory instead of the NCCL all-gather (GPUs of one node, one process
 * each, at most 16): every rank exports the cudaIpc handle of its exchange block, the host program
 * distributes the handles (any transport), every rank imports all of them ([world][64], its own entry is
 * ignored). From then on evaluate_agents / tick run ONE kernel per rank that stores the rank's record
 * straight into every peer's block, waits for the peers' records and does the replicated selection.
 * Call both before or after pmaf_set_shard / pmaf_init (the block does not depend on the shard); the rank and
 * world given to import must equal the shard's when evaluate runs. Every import needs a fresh export on ALL
 * ranks (the export zeroes the block's sequence numbers and drops earlier mappings): a second import without
 * one returns PMAF_ERR_STATE. If import fails (no peer access), the NCCL path stays in use.
 * Failure: a peer whose record does not arrive within 10 s makes evaluate_agents / tick return PMAF_ERR_NCCL.
 * That is FATAL for the whole sharded group — ranks that did complete the tick have moved on, so the replicas of
 * the incumbent best agent may differ: every later evaluate on this handle fails the same way until all ranks
 * have called pmaf_init again and re-done the export / import handshake. */
int pmaf_p2p_export(pmaf_planner *p, unsigned char handle_out[64]);
int pmaf_p2p_import(pmaf_planner *p, const unsigned char *handles /* [world][64] */, int rank, int world);

/* CfManager::init (h:93-102, cpp:41-124). n_agents = k_attr.size() (at least one agent — HAD —
 * is always created, cpp:70-72); gains are per agent; the incumbent best agent SURVIVES init
 * (best_agent_ is not reset, cpp:41-124). A live handle is fully re-initialised otherwise.
 * Rollouts use prediction_freq_multiple * delta_t (cpp:122). RANDOM agents (index >= 5) draw
 * their random vectors from the handle's generator: std::random_device-seeded like
 * helper_functions.cpp:8-12 unless pmaf_seed_random_vecs was called. */
int pmaf_init(pmaf_planner *p, const double goal[3], double delta_t, int n_obs, const double *obs_pos,
              const double *obs_vel, const double *obs_rad, int n_agents, const double *k_attr,
              const double *k_circ, const double *k_repel, const double *k_damp, const double *k_manip,
              int n_force, const double *k_repel_force, double velocity_max, double approach_dist,
              double detect_shell_rad, uint64_t max_prediction_steps, uint64_t prediction_freq_multiple,
              double agent_mass, double radius);

/* Determinism hooks replacing RandomCfAgent's std::random_device draws (cf_agent.h:338-342):
 * seed the generator used by subsequent pmaf_init calls, or overwrite the vectors of the current
 * agents: vecs[n_agents][n_obs][3], rows of non-RANDOM agents ignored. */
int pmaf_seed_random_vecs(pmaf_planner *p, uint64_t seed);
int pmaf_set_random_vecs(pmaf_planner *p, const double *vecs, int n_agents, int n_obs);
int pmaf_get_random_vecs(pmaf_planner *p, double *vecs, int n_agents, int n_obs);

/* ---- per-tick calls, in the order planCallback makes them (node:329-369) ------------------------ */

/* CfManager::setInitialPosition (cpp:226-236): init_pos of the real agent and of every agent;
 * agents' paths restart at `pos`, the real agent's path is appended to (cf_agent.cpp:34-46). */
int pmaf_set_initial_position(pmaf_planner *p, const double pos[3]);

/* CfManager::setRealEEAgentPosition (cpp:216-218; node:334, open_loop == false only). */
int pmaf_set_real_position(pmaf_planner *p, const double pos[3]);

/* CfManager::startPrediction (h:57-61): launches the rollout of every agent on the handle's
 * stream and returns immediately. Rollouts always run to termination (distance to goal <= 0.1
 * or max_prediction_steps path points, cf_agent.cpp:310-311); the reference's wall-clock
 * dependent early stop is not reproduced. Calling it again without pmaf_reset_agents is a no-op
 * (every agent's stop condition already holds). */
int pmaf_start_prediction(pmaf_planner *p);

/* CfManager::stopPrediction (cpp:126-140): returns when no rollout is running (stream sync). */
int pmaf_stop_prediction(pmaf_planner *p);

/* CfManager::evaluateAgents (cpp:293-356): per-agent cost, serial-order argmin (strict <,
 * lowest index wins), 0.9 hysteresis against the incumbent; *best_index is the returned agent
 * index. The obstacle arguments are accepted and ignored, as in the reference. Implies
 * pmaf_stop_prediction. */
int pmaf_evaluate_agents(pmaf_planner *p, int n_obs, const double *obs_pos, const double *obs_vel,
                         const double *obs_rad, double k_goal_dist, double k_path_len, double k_safe_dist,
                         double k_workspace, const double ws_limits[6], int *best_index);

/* CfManager::moveRealEEAgent (cpp:257-263 -> RealCfAgent::cfPlanner, cf_agent.cpp:343-366): `steps`
 * integration steps of the real agent on the LIVE obstacle list with the gains of agent
 * `agent_id` and the heuristics of the incumbent best agent. */
int pmaf_move_real_agent(pmaf_planner *p, int n_obs, const double *obs_pos, const double *obs_vel,
                         const double *obs_rad, double delta_t, int steps, int agent_id);

/* CfManager::resetEEAgents (cpp:246-255): every agent restarts at `pos` with clamp(vel), obstacle
 * positions/velocities := live ones (radii keep their init values, cf_agent.cpp:63-70), known
 * flags := the real agent's, min_obs_dist := detect_shell_rad. */
int pmaf_reset_agents(pmaf_planner *p, const double pos[3], const double vel[3], int n_obs,
                      const double *obs_pos, const double *obs_vel, const double *obs_rad);

/* The obstacle feed on the device: dynamic_obstacle_node's integration step (src/dynamic_obstacle_node.cpp:
 * 352-369: `cur_pos.at(i) += cur_vel.at(i) / frequency` for the published obstacles 0..n_feed-1, the sentinel is
 * never published) followed by obstacleCallback's overwrite of the planner's list (src/panda_bimanual_control.cpp:
 * 302-309), applied to the DEVICE-resident live list — the one the last call that took an obstacle list
 * uploaded — so that a moving scene costs no host->device copy per tick. The same IEEE operations run on the
 * library's host mirror of the list: a caller that advances its own copy identically and keeps passing it finds
 * it recognised as unchanged (no upload), a caller of pmaf_tick may pass obs_pos = obs_vel = obs_rad = NULL
 * instead. The step is applied by the next pmaf_tick inside its tick kernel (no extra launch), or by a small
 * kernel of its own before any other consumer of the list. */
int pmaf_feed_obstacles(pmaf_planner *p, int n_feed, double frequency);

/* One whole planCallback tick (node:329-369) as a device-resident chain with a single host
 * synchronisation: [set_real_position if measured_pos != NULL] -> stop -> evaluate -> move real
 * agent (1 step, gains of the best agent) -> reset agents (from the real agent's new state) ->
 * start. Outputs: best index, next position (the `goals` message) and next velocity.
 * obs_pos = obs_vel = obs_rad = NULL: use the device-resident live list (n_obs entries) as the last upload /
 * pmaf_feed_obstacles left it. */
int pmaf_tick(pmaf_planner *p, const double *measured_pos, int n_obs, const double *obs_pos,
              const double *obs_vel, const double *obs_rad, double delta_t, double k_goal_dist,
              double k_path_len, double k_safe_dist, double k_workspace, const double ws_limits[6],
              int *best_index, double next_pos[3], double next_vel[3]);

/* ---- getters (h:69-92) ---------------------------------------------------------------------------- */
int pmaf_get_num_agents(pmaf_planner *p, int *n_agents);
int pmaf_get_next_position(pmaf_planner *p, double out[3]);    /* getNextPosition, h:74-76 */
int pmaf_get_next_velocity(pmaf_planner *p, double out[3]);    /* getNextVelocity, h:78 */
int pmaf_get_ee_force(pmaf_planner *p, double out[3]);         /* getEEForce, h:79 */
int pmaf_get_goal_position(pmaf_planner *p, double out[3]);    /* getGoalPosition, h:80 */
int pmaf_get_initial_position(pmaf_planner *p, double out[3]); /* getInitialPosition, h:77 */
int pmaf_get_dist_from_goal(pmaf_planner *p, double *out);     /* getDistFromGoal, h:87-89 */
int pmaf_get_best_agent_type(pmaf_planner *p, int *out);       /* getBestAgentType, h:73; -1 if none */
int pmaf_get_best_agent_id(pmaf_planner *p, int *out);         /* best_agent_->getAgentID(); 0 if none */
int pmaf_get_num_prediction_steps(pmaf_planner *p, int agent, int *out); /* h:81-83 */
int pmaf_get_real_num_prediction_steps(pmaf_planner *p, int *out);       /* h:84-86 */

/* getPredictedPathLengths / getPredictionTimes / getAgentSuccess (cpp:192-214) and friends in one
 * call; any pointer may be NULL. steps = path points per agent. */
int pmaf_get_agent_summaries(pmaf_planner *p, int *steps, double *length, double *min_obs_dist, int *reached,
                             double *pred_time_ns, int *agent_type);

/* getPredictedPaths (cpp:184-190): out[n_agents][stride][3]; only the first steps[a] rows of an
 * agent are written. */
int pmaf_get_predicted_paths(pmaf_planner *p, double *out, int stride);
/* one agent's path (e.g. the best one, for the predicted_paths marker, node:340-347) */
int pmaf_get_predicted_path(pmaf_planner *p, int agent, double *out, int max_points, int *n_points);
int pmaf_get_agent_velocities(pmaf_planner *p, double *out /* [n_agents][3] */);
/* getPlannedTrajectory (h:90-92): the real agent's path */
int pmaf_get_planned_trajectory(pmaf_planner *p, double *out, int max_points, int *n_points);
/* per-(agent, obstacle) latch state: known[(n_agents+1)][n_obs], rot[(n_agents+1)][n_obs][3];
 * the extra last row is the real agent's (cf_agent.h:52-53) */
int pmaf_get_obstacle_state(pmaf_planner *p, int n_obs, int *known, double *rot);
/* The k cheapest agents of the last evaluate_agents (ascending cost, lowest index first among equals,
 * NaN costs last) and their paths decimated to every stride-th point (the last point always kept):
 * agent_index[k] (-1 where fewer than k agents exist), n_points[k], paths[k][max_points][3]. This is the
 * batched export for the node's predicted-path markers (node:340-347, 390-427), which otherwise copies
 * every agent's whole path twice per tick. */
int pmaf_get_best_paths(pmaf_planner *p, int k, int stride, int max_points, int *agent_index, int *n_points,
                        double *paths);
/* costs computed by the last evaluate_agents */
int pmaf_get_costs(pmaf_planner *p, double *costs /* [n_agents] */);

/* ---- dry-run relay ------------------------------------------------------------------------------------ */
/* The in-process equivalent of launch/dry_run.launch (:9,41 relays the planner's `goals` output back as
 * its `position` input): runs `ticks` control ticks — planCallback, panda_bimanual_control.cpp:329-369:
 * stop_prediction, evaluate_agents, move_real_agent, get_next_position/velocity, reset_agents,
 * start_prediction — through the entry points above with the caller's HOST obstacle lists. After every
 * tick the obstacle feed of dynamic_obstacle_node (dynamic_obstacle_node.cpp:352-369) advances
 * obs_pos[0 .. n_feed) in place by obs_vel / feed_frequency (n_feed = 0: static scene; the node feeds
 * every obstacle but the trailing sentinel). Optional outputs per tick: best[ticks], next_pos[ticks][3],
 * next_vel[ticks][3]. *seconds = wall time spent inside the ticks (CLOCK_MONOTONIC around each tick).
 * flags: PMAF_DRY_RUN_WAIT_ROLLOUT waits for each tick's rollout inside its timed region (benchmarks: a
 * tick then costs calls + rollout); PMAF_DRY_RUN_FLUSH_L2 evicts the L2 before each tick, untimed;
 * PMAF_DRY_RUN_PROFILE: seconds must hold 7 doubles, seconds[1..6] = wall time summed per call
 * (stop, evaluate, move_real, get + reset, start, final wait). */
#define PMAF_DRY_RUN_WAIT_ROLLOUT 1
#define PMAF_DRY_RUN_FLUSH_L2 2
#define PMAF_DRY_RUN_PROFILE 4
#define PMAF_DRY_RUN_TICK_TIMES 8 /* seconds must hold 7 + ticks doubles; seconds[7 + t] = wall time of tick t */
#define PMAF_DRY_RUN_DEVICE_FEED 16 /* the feed between ticks also runs on the device (pmaf_feed_obstacles): the
                                     * calls' host lists are recognised as unchanged and not uploaded */
int pmaf_dry_run(pmaf_planner *p, int ticks, int n_obs, double *obs_pos, const double *obs_vel, const double *obs_rad,
                 int n_feed, double feed_frequency, double delta_t, double k_goal_dist, double k_path_len,
                 double k_safe_dist, double k_workspace, const double ws_limits[6], int flags, double *seconds,
                 int *best /* [ticks] or NULL */, double *next_pos /* [ticks][3] or NULL */,
                 double *next_vel /* [ticks][3] or NULL */);

/* ---- instrumentation ------------------------------------------------------------------------------ */
typedef struct {
  uint64_t kernel_launches;  /* kernels of this library launched on the handle since create */
  uint64_t collectives;      /* NCCL collectives enqueued since create */
  uint64_t rollouts;         /* rollout kernels among them */
  uint64_t agent_steps;      /* integration steps executed by the last completed rollout */
  uint64_t agent_steps_total; /* ... by all completed rollouts since create */
  double last_rollout_ms;    /* device time of the last completed rollout kernel (CUDA events) */
  double rollout_ms_total;   /* sum of device times of all completed rollout kernels */
  uint64_t h2d_bytes;        /* host->device bytes copied since create */
  uint64_t d2h_bytes;        /* device->host bytes copied since create */
  int lanes_per_agent;       /* rollout kernel configuration in use */
  int block_threads;
  int grid_blocks;
  int smem_bytes;
  int occupancy_build;
  int reserved_;
  uint64_t general_steps_total; /* steps of all completed rollouts that took the general (branchy) step
                                 * instead of the straight-line one (latency build; csrc/pmaf_fast.cuh) */
} pmaf_counters;
int pmaf_get_counters(pmaf_planner *p, pmaf_counters *out);
/* Rollout kernel timing (CUDA events around every rollout; last_rollout_ms / rollout_ms_total). On by default.
 * Off: no events, and pmaf_tick's rollout starts as a programmatic dependent launch of the tick kernel (its
 * launch latency overlaps the evaluate / real-agent step) — the production setting. */
int pmaf_set_rollout_timing(pmaf_planner *p, int on);
/* Developer instrumentation (libraries built with -DPMAF_FAST_STATS, all zero otherwise): how many steps
 * each reason kept off the straight-line step since create — out[0] unusable / candidate count, [1] start
 * radius threshold, [2] range flag (distance chain), [3] first detection needing the general latch,
 * [4] range flag (force chain), [5] |F| threshold, [6] sentinel in reach, [7] acceleration clamp,
 * [8] range flag (integrator). Call after pmaf_get_counters. */
int pmaf_get_fast_stats(pmaf_planner *p, uint64_t out[12]);
/* rollout kernel shape override for experiments: lanes_per_agent in {0 (auto), 4, 8, 16, 32};
 * block_threads 0 (auto) or a multiple of 32 up to 256; occupancy 0 (auto), 1, 3 or 4 = resident CTAs
 * per SM the kernel's register budget is compiled for (255 / 170 / 128 registers per thread). */
int pmaf_set_tuning(pmaf_planner *p, int lanes_per_agent, int block_threads, int occupancy);
/* By default an obstacle list that is byte-identical to the last one uploaded is not copied to the
 * device again (static scenes); dedup = 0 forces the host->device copy on every call. */
int pmaf_set_upload_dedup(pmaf_planner *p, int dedup);
/* Device-side stopwatch on the handle's stream (CUDA events): start records, stop records, waits
 * for the stream and returns the elapsed device milliseconds between the two. */
int pmaf_timer_start(pmaf_planner *p);
int pmaf_timer_stop(pmaf_planner *p, double *elapsed_ms);
/* Benchmark hygiene: evict the L2 cache by writing a 256 MiB scratch buffer on the handle's stream. */
int pmaf_flush_l2(pmaf_planner *p);
/* Measured FP64 FMA-pipe peak of the device (dependent DFMA chains on every SM), in TFLOP/s:
 * the denominator of the FP-issue roofline of the rollout kernel. */
int pmaf_measure_fp64_peak(pmaf_planner *p, double *tflops);
/* Self-test of the kernels' branch-free sqrt / division (FastMath, csrc/pmaf_math.cuh) against CUDA's
 * IEEE built-ins on ~`samples` random and adversarial operands: out = { sqrt mismatches, division
 * mismatches, shared-reciprocal vector division mismatches, operands rejected by the range check,
 * comparisons made }. Any mismatch is a bug. */
int pmaf_selftest_math(pmaf_planner *p, uint64_t samples, uint64_t seed, uint64_t out[5]);
/* Developer instrumentation: cycles spent per section of the step loop by the first 64 agents in
 * their last rollout, out[64][12]; all zero unless libpmaf was built with -DPMAF_SECTION_TIMERS. */
int pmaf_get_section_cycles(pmaf_planner *p, long long *out);

/* ---- downstream kinematics ---------------------------------------------------------------------
 * Dual-quaternion kinematics of the Franka Panda as the reference's controller evaluates it every cycle for one
 * robot on the CPU — pose (fkm) and pose Jacobian through dqrobotics' DQ_SerialManipulator with the modified-DH
 * table of src/franka_robot.cpp:6-22 (CoSTPController::calculateControlPreliminaries, src/costp_controller.cpp:
 * 111-126) and the geometric Jacobian geomJ (src/costp_controller.cpp:465-492) — on the device. dqrobotics is
 * not vendored by the reference: the algebra follows its published definitions and PARITY WITH A dqrobotics
 * BUILD IS UNPINNED (oracle/dq_oracle.c). Dual quaternions are 8 doubles: primary w x y z, dual w x y z;
 * base_dq is the arm's base frame (r + eps/2 p r, src/franka_robot.cpp:14-20). */
int pmaf_dq_kinematics(pmaf_planner *p, const double base_dq[8], const double q[7], double pose[8],
                       double pose_jacobian[56] /* 8 x 7 row-major */, double geom_jacobian[42] /* 6 x 7 row-major */);
/* Joint limits of the Panda as the controller holds them (src/costp_controller.cpp:41-44). */
void pmaf_panda_joint_limits(double q_lo[7], double q_hi[7]);

typedef struct {
  double max_pos_err;        /* largest residual |path point - end-effector position| after the point's step */
  double min_joint_margin;   /* smallest distance of any joint to its nearer limit along the path (< 0: violated) */
  double min_manipulability; /* smallest sqrt(det(Jt Jt^T)) of the translation Jacobian along the path */
  int feasible;              /* every point tracked within tol_pos, inside the joint limits */
  int first_bad_point;       /* first point that failed, -1 if none */
  double q_final[7];
} pmaf_path_score;
/* Batched feasibility score of the predicted end-effector paths, on the device-resident paths of the last
 * rollout (nothing is copied out but the scores): one damped-least-squares step per path point from q_start,
 * dq = Jt^T (Jt Jt^T + damping I)^-1 e (the form of src/costp_controller.cpp:134-135).
 * k <= 0: every local agent, out[n_agents], agent_index may be NULL. k >= 1 (<= 64): the k cheapest agents of the
 * last evaluate in (cost, index) order as pmaf_get_best_paths, agent_index[k] receives their global indices
 * (-1 = fewer agents than k; that score is of an empty path). */
int pmaf_score_paths(pmaf_planner *p, int k, const double base_dq[8], const double q_start[7], const double q_lo[7],
                     const double q_hi[7], double damping, double tol_pos, int *agent_index, pmaf_path_score *out);

#ifdef __cplusplus
}
#endif
#endif /* PMAF_H */
