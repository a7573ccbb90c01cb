// TEST INFRASTRUCTURE — not product code. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load what this builds.
//
// C-ABI harness around the UNMODIFIED reference planner
// (ghostplanner::cfplanner::CfManager / CfAgent), whose sources are compiled
// where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libcfref.so. Nothing here restates the algorithm: every number
// that comes out of this library was computed by the reference's own
// cf_agent.cpp / cf_manager.cpp code.
//
// What the harness adds from the outside (SURVEY.md §8c "determinism fixes"):
//   1. seeded random vectors: RandomCfAgent::random_vecs_ is overwritten after
//      init() (the reference draws them from std::random_device,
//      helper_functions.cpp:8-12) — compiled with -fno-access-control so the
//      protected member is reachable without editing reference headers;
//   2. run-to-termination rollouts: either the reference's own thread-per-agent
//      driver (cf_manager.cpp:118-123) polled until every agent's stop
//      condition (cf_agent.cpp:310-311) holds, or a pooled driver that calls
//      the reference's public per-step methods in cfPrediction order
//      (cf_agent.cpp:312-327) from an OpenMP team, for A >> host cores;
//   3. a no-op pthread_cancel: joinPredictionThreads() cancels a thread id it
//      has already joined (cf_manager.cpp:152-156), which is undefined and can
//      crash in glibc; interposing the symbol keeps the reference source
//      untouched.
#include <omp.h>
#include <pthread.h>

#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "bimanual_planning_ros/cf_manager.h"

extern "C" int pthread_cancel(pthread_t) { return 0; }

using ghostplanner::cfplanner::CfAgent;
using ghostplanner::cfplanner::CfManager;
using ghostplanner::cfplanner::Obstacle;
using ghostplanner::cfplanner::RandomCfAgent;
using Eigen::Vector3d;

namespace {
struct Ref {
  CfManager mgr;
  double pred_dt = 0.0;  // prediction_freq_multiple * delta_t (cf_manager.cpp:122)
  size_t max_steps = 0;
  bool threads_alive = false;
};

std::vector<Obstacle> make_obstacles(int n, const double *pos, const double *vel,
                                     const double *rad) {
  std::vector<Obstacle> obs;
  obs.reserve(n);
  for (int i = 0; i < n; ++i) {
    obs.emplace_back(Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]),
                     Vector3d(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]), rad[i]);
  }
  return obs;
}

bool agent_terminated(const CfAgent &a, size_t max_steps) {
  return !(a.getDistFromGoal() > 0.1 && a.pos_.size() < max_steps);
}
}  // namespace

extern "C" {

void *cfref_create() { return new Ref(); }

void cfref_destroy(void *h) { delete static_cast<Ref *>(h); }

// CfManager::init (cf_manager.cpp:41-124). keep_threads = 0 joins the
// per-agent prediction threads right after init so that they do not busy-spin
// (cf_agent.cpp:307) while the pooled driver is used.
void cfref_init(void *h, const double *goal, double delta_t, int n_obs,
                const double *obs_pos, const double *obs_vel, const double *obs_rad,
                int n_agents, const double *k_a, const double *k_c, const double *k_r,
                const double *k_d, const double *k_manip, int n_force,
                const double *k_r_force, double vel_max, double approach_dist,
                double detect_shell_rad, unsigned long max_prediction_steps,
                unsigned long prediction_freq_multiple, double agent_mass,
                double radius, int keep_threads) {
  Ref *r = static_cast<Ref *>(h);
  std::vector<double> ka(k_a, k_a + n_agents), kc(k_c, k_c + n_agents),
      kr(k_r, k_r + n_agents), kd(k_d, k_d + n_agents), km(k_manip, k_manip + n_agents),
      kf(k_r_force, k_r_force + n_force);
  r->mgr.init(Vector3d(goal[0], goal[1], goal[2]), delta_t,
              make_obstacles(n_obs, obs_pos, obs_vel, obs_rad), ka, kc, kr, kd, km, kf,
              vel_max, approach_dist, detect_shell_rad, max_prediction_steps,
              prediction_freq_multiple, agent_mass, radius);
  r->pred_dt = prediction_freq_multiple * delta_t;
  r->max_steps = max_prediction_steps;
  r->threads_alive = true;
  if (!keep_threads) {
    r->mgr.joinPredictionThreads();
    r->threads_alive = false;
  }
}

int cfref_num_agents(void *h) { return (int)static_cast<Ref *>(h)->mgr.ee_agents_.size(); }

// Overwrite RandomCfAgent::random_vecs_ (cf_agent.h:326,338-342) for every
// RANDOM agent; vecs is [n_agents][n_obs][3], rows of non-random agents ignored.
void cfref_set_random_vecs(void *h, const double *vecs, int n_obs) {
  Ref *r = static_cast<Ref *>(h);
  for (size_t a = 0; a < r->mgr.ee_agents_.size(); ++a) {
    auto *ra = dynamic_cast<RandomCfAgent *>(r->mgr.ee_agents_[a].get());
    if (!ra) continue;
    for (int i = 0; i < n_obs && i < (int)ra->random_vecs_.size(); ++i) {
      const double *v = vecs + ((size_t)a * n_obs + i) * 3;
      ra->random_vecs_[i] = Vector3d(v[0], v[1], v[2]);
    }
  }
}

void cfref_get_random_vecs(void *h, double *vecs, int n_obs) {
  Ref *r = static_cast<Ref *>(h);
  for (size_t a = 0; a < r->mgr.ee_agents_.size(); ++a) {
    auto *ra = dynamic_cast<RandomCfAgent *>(r->mgr.ee_agents_[a].get());
    for (int i = 0; i < n_obs; ++i) {
      double *v = vecs + ((size_t)a * n_obs + i) * 3;
      if (ra && i < (int)ra->random_vecs_.size()) {
        v[0] = ra->random_vecs_[i].x(), v[1] = ra->random_vecs_[i].y(),
        v[2] = ra->random_vecs_[i].z();
      } else {
        v[0] = v[1] = v[2] = 0.0;
      }
    }
  }
}

void cfref_set_initial_position(void *h, const double *p) {
  static_cast<Ref *>(h)->mgr.setInitialPosition(Vector3d(p[0], p[1], p[2]));
}

void cfref_set_real_position(void *h, const double *p) {
  static_cast<Ref *>(h)->mgr.setRealEEAgentPosition(Vector3d(p[0], p[1], p[2]));
}

void cfref_start_prediction(void *h) { static_cast<Ref *>(h)->mgr.startPrediction(); }
void cfref_stop_prediction(void *h) { static_cast<Ref *>(h)->mgr.stopPrediction(); }

// Reference thread-per-agent driver, run to termination: start, wait until
// every agent's loop condition (cf_agent.cpp:310-311) is false, stop.
// Returns wall seconds from startPrediction() to all-terminated.
double cfref_rollout_threads(void *h) {
  Ref *r = static_cast<Ref *>(h);
  if (!r->threads_alive) return -1.0;
  auto t0 = std::chrono::steady_clock::now();
  auto t1 = t0;
  for (;;) {
    r->mgr.startPrediction();
    // let every thread observe run_prediction_ and finish its inner loop
    int idle_polls = 0;
    while (idle_polls < 3) {
      std::this_thread::sleep_for(std::chrono::microseconds(50));
      bool any = false;
      for (auto &a : r->mgr.ee_agents_) any = any || a->getRunningStatus();
      idle_polls = any ? 0 : idle_polls + 1;
    }
    t1 = std::chrono::steady_clock::now();
    r->mgr.stopPrediction();
    bool all_done = true;
    for (auto &a : r->mgr.ee_agents_) all_done = all_done && agent_terminated(*a, r->max_steps);
    if (all_done) break;
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

// Pooled driver: the body of CfAgent::cfPrediction's inner loop
// (cf_agent.cpp:310-336) executed by an OpenMP team over agents, calling the
// reference's own methods. Returns wall seconds.
double cfref_rollout_pooled(void *h, int n_threads) {
  Ref *r = static_cast<Ref *>(h);
  CfManager &m = r->mgr;
  const int n = (int)m.ee_agents_.size();
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (int i = 0; i < n; ++i) {
    CfAgent &a = *m.ee_agents_[i];
    auto begin = std::chrono::steady_clock::now();
    bool ran = false;
    while (a.getDistFromGoal() > 0.1 && a.pos_.size() < r->max_steps) {
      ran = true;
      a.resetForce();
      double k_goal_scale = 1.0;
      if (!(a.getDistFromGoal() < a.approach_dist_ ||
            (a.vel_.norm() < 0.5 * a.vel_max_ &&
             (a.getLatestPosition() - a.init_pos_).norm() < 0.2))) {
        a.circForce(a.obstacles_, m.k_c_ee_[i]);
        if (a.force_.norm() > 1e-5) {
          k_goal_scale = a.attractorForceScaling(a.obstacles_);
        }
      }
      a.repelForce(a.obstacles_, m.k_r_ee_[i]);
      a.attractorForce(m.k_a_ee_[i], m.k_d_ee_[i], k_goal_scale);
      a.updatePositionAndVelocity(r->pred_dt);
      a.predictObstacles(r->pred_dt);
    }
    auto end = std::chrono::steady_clock::now();
    if (ran) {
      a.prediction_time_ = (end - begin).count();
      a.reached_goal_ = a.getDistFromGoal() < 0.100001;
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int cfref_evaluate_agents(void *h, int n_obs, const double *obs_pos, const double *obs_vel,
                          const double *obs_rad, double k_goal_dist, double k_path_len,
                          double k_safe_dist, double k_workspace, const double *ws) {
  Ref *r = static_cast<Ref *>(h);
  Eigen::Matrix<double, 6, 1> lim;
  for (int i = 0; i < 6; ++i) lim(i) = ws[i];
  return r->mgr.evaluateAgents(make_obstacles(n_obs, obs_pos, obs_vel, obs_rad), k_goal_dist,
                               k_path_len, k_safe_dist, k_workspace, lim);
}

void cfref_move_real_agent(void *h, int n_obs, const double *obs_pos, const double *obs_vel,
                           const double *obs_rad, double delta_t, int steps, int agent_id) {
  static_cast<Ref *>(h)->mgr.moveRealEEAgent(make_obstacles(n_obs, obs_pos, obs_vel, obs_rad),
                                             delta_t, steps, agent_id);
}

void cfref_reset_agents(void *h, const double *pos, const double *vel, int n_obs,
                        const double *obs_pos, const double *obs_vel, const double *obs_rad) {
  static_cast<Ref *>(h)->mgr.resetEEAgents(Vector3d(pos[0], pos[1], pos[2]),
                                           Vector3d(vel[0], vel[1], vel[2]),
                                           make_obstacles(n_obs, obs_pos, obs_vel, obs_rad));
}

void cfref_get_next_position(void *h, double *p) {
  Vector3d v = static_cast<Ref *>(h)->mgr.getNextPosition();
  p[0] = v.x(), p[1] = v.y(), p[2] = v.z();
}
void cfref_get_next_velocity(void *h, double *p) {
  Vector3d v = static_cast<Ref *>(h)->mgr.getNextVelocity();
  p[0] = v.x(), p[1] = v.y(), p[2] = v.z();
}
void cfref_get_ee_force(void *h, double *p) {
  Vector3d v = static_cast<Ref *>(h)->mgr.getEEForce();
  p[0] = v.x(), p[1] = v.y(), p[2] = v.z();
}
double cfref_get_dist_from_goal(void *h) { return static_cast<Ref *>(h)->mgr.getDistFromGoal(); }
int cfref_get_best_agent_type(void *h) {
  Ref *r = static_cast<Ref *>(h);
  return r->mgr.best_agent_ ? r->mgr.getBestAgentType() : -1;
}
int cfref_get_best_agent_id(void *h) {
  Ref *r = static_cast<Ref *>(h);
  return r->mgr.best_agent_ ? r->mgr.best_agent_->getAgentID() : 0;
}
int cfref_get_num_prediction_steps(void *h, int agent) {
  return static_cast<Ref *>(h)->mgr.getNumPredictionSteps(agent);
}
int cfref_get_real_num_steps(void *h) {
  return static_cast<Ref *>(h)->mgr.getRealNumPredictionSteps();
}

// per-agent scalars: steps[A] (path points), length[A], min_obs_dist[A],
// reached[A], pred_time_ns[A], agent_type[A]; any pointer may be null.
void cfref_get_agent_summaries(void *h, int *steps, double *length, double *min_obs_dist,
                               int *reached, double *pred_time_ns, int *agent_type) {
  CfManager &m = static_cast<Ref *>(h)->mgr;
  for (size_t a = 0; a < m.ee_agents_.size(); ++a) {
    CfAgent &ag = *m.ee_agents_[a];
    if (steps) steps[a] = ag.getNumPredictionSteps();
    if (length) length[a] = ag.getPathLength();
    if (min_obs_dist) min_obs_dist[a] = ag.getMinObsDist();
    if (reached) reached[a] = ag.getReachedGoal() ? 1 : 0;
    if (pred_time_ns) pred_time_ns[a] = ag.getPredictionTime();
    if (agent_type) agent_type[a] = (int)ag.getAgentType();
  }
}

// paths as [A][stride][3]; rows beyond an agent's step count are left untouched.
void cfref_get_predicted_paths(void *h, double *out, int stride) {
  CfManager &m = static_cast<Ref *>(h)->mgr;
  auto paths = m.getPredictedPaths();
  for (size_t a = 0; a < paths.size(); ++a) {
    for (size_t k = 0; k < paths[a].size() && (int)k < stride; ++k) {
      double *o = out + ((size_t)a * stride + k) * 3;
      o[0] = paths[a][k].x(), o[1] = paths[a][k].y(), o[2] = paths[a][k].z();
    }
  }
}

void cfref_get_agent_velocities(void *h, double *out) {
  CfManager &m = static_cast<Ref *>(h)->mgr;
  for (size_t a = 0; a < m.ee_agents_.size(); ++a) {
    Vector3d v = m.ee_agents_[a]->getVelocity();
    out[3 * a] = v.x(), out[3 * a + 1] = v.y(), out[3 * a + 2] = v.z();
  }
}

int cfref_get_planned_trajectory(void *h, double *out, int max_points) {
  auto traj = static_cast<Ref *>(h)->mgr.getPlannedTrajectory();
  int n = (int)traj.size();
  for (int k = 0; k < n && k < max_points; ++k) {
    out[3 * k] = traj[k].x(), out[3 * k + 1] = traj[k].y(), out[3 * k + 2] = traj[k].z();
  }
  return n;
}

// per-(agent, obstacle) state: known[A][O] (0/1), rot[A][O][3]; row A (one past
// the last agent) is the real agent's.
void cfref_get_obstacle_state(void *h, int n_obs, int *known, double *rot) {
  CfManager &m = static_cast<Ref *>(h)->mgr;
  size_t n = m.ee_agents_.size();
  for (size_t a = 0; a <= n; ++a) {
    const CfAgent &ag = a < n ? *m.ee_agents_[a] : static_cast<const CfAgent &>(m.real_ee_agent_);
    for (int i = 0; i < n_obs; ++i) {
      size_t k = a * n_obs + i;
      bool have = i < (int)ag.known_obstacles_.size();
      if (known) known[k] = have && ag.known_obstacles_[i] ? 1 : 0;
      if (rot) {
        Vector3d v = have ? ag.field_rotation_vecs_[i] : Vector3d(0, 0, 0);
        rot[3 * k] = v.x(), rot[3 * k + 1] = v.y(), rot[3 * k + 2] = v.z();
      }
    }
  }
}

int cfref_host_threads() { return omp_get_max_threads(); }

}  // extern "C"
