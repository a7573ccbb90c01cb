// cf_manager.hpp — drop-in for bimanual_planning_ros/cf_manager.h over libpmaf.so.
//
// Same namespace, class name and public member signatures as the reference's
// ghostplanner::cfplanner::CfManager (include/bimanual_planning_ros/cf_manager.h:18-137), so that
// src/panda_bimanual_control.cpp compiles against it unchanged: in the catkin package, replace the
// contents of cf_manager.h by `#include <pmaf/cf_manager.hpp>`, drop src/cf_manager.cpp and
// src/cf_agent.cpp from the `utilities` library and link libpmaf.so.
// The rollouts, the cost / best-agent selection and the real-agent step run on the GPU; this class
// only converts Eigen / Obstacle containers to the flat arrays of include/pmaf.h.
//
// Differences from the reference, all deliberate:
//   * rollouts run to termination on startPrediction() (the reference's threads can be cut short by
//     stopPrediction(); results then depend on wall-clock time);
//   * RandomCfAgent vectors come from a std::random_device-seeded generator inside libpmaf unless
//     seedRandomVectors()/setRandomVectors() is used (the reference offers no seeding);
//   * members with no caller in the planner node and no defined behaviour on this path
//     (getLinkForce, moveAgent, moveAgents, moveAgentsPar, evaluatePath, setEEAgentPositions,
//     setEEAgentPosAndVels) throw std::logic_error;
//   * failures throw std::runtime_error carrying pmaf_last_error() (the reference asserts or throws
//     std::out_of_range).
#pragma once

#include <pmaf.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "bimanual_planning_ros/obstacle.h"
#include "eigen3/Eigen/Dense"

namespace ghostplanner {
namespace cfplanner {

class CfManager {
  pmaf_planner *h_ = nullptr;
  int device_ = 0;
  int n_agents_ = 0;
  int n_obstacles_ = 0;
  size_t max_steps_ = 0;
  // getPredictedPaths() is a full [agents][horizon][3] device copy; the node calls it up to 3 x agents times per
  // tick between stopPrediction and resetEEAgents (panda_bimanual_control.cpp:341-343), so the converted paths
  // are cached until a call that can change them
  std::vector<std::vector<Eigen::Vector3d>> paths_cache_;
  bool paths_cached_ = false;

  static void check(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + pmaf_last_error());
  }
  pmaf_planner *handle() {
    if (!h_) check(pmaf_create(&h_, device_), "pmaf_create");
    return h_;
  }
  struct Flat {
    std::vector<double> pos, vel, rad;
    explicit Flat(const std::vector<Obstacle> &obs) {
      pos.reserve(3 * obs.size()), vel.reserve(3 * obs.size()), rad.reserve(obs.size());
      for (const Obstacle &o : obs) {
        const Eigen::Vector3d p = o.getPosition(), v = o.getVelocity();
        pos.insert(pos.end(), {p.x(), p.y(), p.z()});
        vel.insert(vel.end(), {v.x(), v.y(), v.z()});
        rad.push_back(o.getRadius());
      }
    }
    int n() const { return (int)rad.size(); }
  };
  static Eigen::Vector3d vec(const double *p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
  [[noreturn]] static void unsupported(const char *name) {
    throw std::logic_error(std::string("CfManager::") + name + " has no caller on the accelerated path and is not provided");
  }

 public:
  CfManager(const Eigen::Vector3d agent_pos, const Eigen::Vector3d goal_pos, const double delta_t,
            const std::vector<Obstacle> &obstacles, const std::vector<double> &k_a_ee,
            const std::vector<double> &k_c_ee, const std::vector<double> &k_r_ee,
            const std::vector<double> &k_d_ee, const std::vector<double> &k_manip,
            const std::vector<double> &k_r_force, const double velocity_max = 0.5,
            const double approach_dist = 0.25, const double detect_shell_rad = 0.8,
            const size_t max_prediction_steps = 1500, const size_t prediction_freq_multiple = 1,
            const double agent_mass = 1.0, const double radius = 0.01) {
    // the reference's constructor (cf_manager.cpp:16-39) ignores everything after detect_shell_rad
    // when it forwards to init(); agent_pos only seeds the first RealCfAgent, which init() replaces
    (void)agent_pos, (void)max_prediction_steps, (void)prediction_freq_multiple, (void)agent_mass, (void)radius;
    init(goal_pos, delta_t, obstacles, k_a_ee, k_c_ee, k_r_ee, k_d_ee, k_manip, k_r_force, velocity_max,
         approach_dist, detect_shell_rad);
  }
  CfManager() = default;
  explicit CfManager(int cuda_device) : device_(cuda_device) {}
  ~CfManager() { joinPredictionThreads(); }
  CfManager(const CfManager &) = delete;
  CfManager &operator=(const CfManager &) = delete;
  CfManager(CfManager &&o) noexcept { *this = std::move(o); }
  CfManager &operator=(CfManager &&o) noexcept {
    if (this != &o) {
      joinPredictionThreads();
      h_ = o.h_, device_ = o.device_, n_agents_ = o.n_agents_, n_obstacles_ = o.n_obstacles_, max_steps_ = o.max_steps_;
      paths_cache_ = std::move(o.paths_cache_), paths_cached_ = o.paths_cached_;
      o.h_ = nullptr, o.paths_cached_ = false;
    }
    return *this;
  }

  // ---- determinism hooks (not in the reference) ----
  void seedRandomVectors(uint64_t seed) { check(pmaf_seed_random_vecs(handle(), seed), "pmaf_seed_random_vecs"); }
  void setRandomVectors(const std::vector<double> &vecs /* [agents][obstacles][3] */) {
    check(pmaf_set_random_vecs(handle(), vecs.data(), n_agents_, n_obstacles_), "pmaf_set_random_vecs");
  }
  pmaf_planner *nativeHandle() { return handle(); }

  // ---- cf_manager.h:57-68 ----
  void startPrediction() {
    paths_cached_ = false;
    check(pmaf_start_prediction(handle()), "startPrediction");
  }
  void stopPrediction() { check(pmaf_stop_prediction(handle()), "stopPrediction"); }
  void shutdownAllAgents() {}
  void joinPredictionThreads() {
    if (h_) pmaf_destroy(h_);
    h_ = nullptr, paths_cached_ = false;
  }

  // ---- cf_manager.h:69-92 ----
  std::vector<std::vector<Eigen::Vector3d>> getPredictedPaths() {
    if (paths_cached_) return paths_cache_;
    std::vector<int> steps(n_agents_);
    check(pmaf_get_agent_summaries(handle(), steps.data(), nullptr, nullptr, nullptr, nullptr, nullptr), "getPredictedPaths");
    std::vector<double> flat((size_t)n_agents_ * max_steps_ * 3);
    check(pmaf_get_predicted_paths(handle(), flat.data(), (int)max_steps_), "getPredictedPaths");
    std::vector<std::vector<Eigen::Vector3d>> paths(n_agents_);
    for (int a = 0; a < n_agents_; ++a) {
      paths[a].reserve(steps[a]);
      for (int k = 0; k < steps[a]; ++k) paths[a].push_back(vec(&flat[((size_t)a * max_steps_ + k) * 3]));
    }
    paths_cache_ = paths, paths_cached_ = true;
    return paths;
  }
  std::vector<double> getPredictedPathLengths() {
    std::vector<double> v(n_agents_);
    check(pmaf_get_agent_summaries(handle(), nullptr, v.data(), nullptr, nullptr, nullptr, nullptr), "getPredictedPathLengths");
    return v;
  }
  std::vector<double> getPredictionTimes() {
    std::vector<double> v(n_agents_);
    check(pmaf_get_agent_summaries(handle(), nullptr, nullptr, nullptr, nullptr, v.data(), nullptr), "getPredictionTimes");
    return v;
  }
  std::vector<bool> getAgentSuccess() {
    std::vector<int> r(n_agents_);
    check(pmaf_get_agent_summaries(handle(), nullptr, nullptr, nullptr, r.data(), nullptr, nullptr), "getAgentSuccess");
    return std::vector<bool>(r.begin(), r.end());
  }
  int getBestAgentType() {
    int t = 0;
    check(pmaf_get_best_agent_type(handle(), &t), "getBestAgentType");
    return t;
  }
  Eigen::Vector3d getNextPosition() { return get3(pmaf_get_next_position, "getNextPosition"); }
  Eigen::Vector3d getInitialPosition() { return get3(pmaf_get_initial_position, "getInitialPosition"); }
  Eigen::Vector3d getNextVelocity() { return get3(pmaf_get_next_velocity, "getNextVelocity"); }
  Eigen::Vector3d getEEForce() { return get3(pmaf_get_ee_force, "getEEForce"); }
  Eigen::Vector3d getGoalPosition() { return get3(pmaf_get_goal_position, "getGoalPosition"); }
  int getNumPredictionSteps(int agent_id) {
    int n = 0;
    check(pmaf_get_num_prediction_steps(handle(), agent_id, &n), "getNumPredictionSteps");
    return n;
  }
  int getRealNumPredictionSteps() {
    int n = 0;
    check(pmaf_get_real_num_prediction_steps(handle(), &n), "getRealNumPredictionSteps");
    return n;
  }
  double getDistFromGoal() {
    double d = 0;
    check(pmaf_get_dist_from_goal(handle(), &d), "getDistFromGoal");
    return d;
  }
  std::vector<Eigen::Vector3d> getPlannedTrajectory() {
    int n = 0;
    check(pmaf_get_planned_trajectory(handle(), nullptr, 0, &n), "getPlannedTrajectory");
    std::vector<double> flat((size_t)(n > 0 ? n : 1) * 3);
    check(pmaf_get_planned_trajectory(handle(), flat.data(), n, &n), "getPlannedTrajectory");
    std::vector<Eigen::Vector3d> out;
    out.reserve(n);
    for (int k = 0; k < n; ++k) out.push_back(vec(&flat[3 * (size_t)k]));
    return out;
  }

  // ---- cf_manager.h:93-102 ----
  void init(const Eigen::Vector3d goal_pos, const double delta_t, const std::vector<Obstacle> &obstacles,
            const std::vector<double> &k_a_ee, const std::vector<double> &k_c_ee,
            const std::vector<double> &k_r_ee, const std::vector<double> &k_d_ee,
            const std::vector<double> &k_manip, const std::vector<double> &k_r_force,
            const double velocity_max = 0.5, const double approach_dist = 0.25,
            const double detect_shell_rad = 0.8, const size_t max_prediction_steps = 1500,
            const size_t prediction_freq_multiple = 1, const double agent_mass = 1.0,
            const double radius = 0.05) {
    if (!(k_a_ee.size() == k_c_ee.size() && k_c_ee.size() == k_r_ee.size() && k_c_ee.size() == k_manip.size() &&
          k_d_ee.size() == k_a_ee.size()))
      throw std::invalid_argument("CfManager::init: gain vectors differ in length");  // assert at cf_manager.cpp:50-51
    const Flat o(obstacles);
    const double g[3] = {goal_pos.x(), goal_pos.y(), goal_pos.z()};
    check(pmaf_init(handle(), g, delta_t, o.n(), o.pos.data(), o.vel.data(), o.rad.data(), (int)k_a_ee.size(),
                    k_a_ee.data(), k_c_ee.data(), k_r_ee.data(), k_d_ee.data(), k_manip.data(), (int)k_r_force.size(),
                    k_r_force.data(), velocity_max, approach_dist, detect_shell_rad, max_prediction_steps,
                    prediction_freq_multiple, agent_mass, radius),
          "CfManager::init");
    check(pmaf_get_num_agents(h_, &n_agents_), "CfManager::init");
    n_obstacles_ = o.n(), max_steps_ = max_prediction_steps;
  }

  std::vector<Eigen::Vector3d> getLinkForce(const std::vector<Eigen::Vector3d> &, const std::vector<Obstacle> &) {
    unsupported("getLinkForce");
  }

  // ---- cf_manager.h:107-115 ----
  void setRealEEAgentPosition(const Eigen::Vector3d &position) {
    const double p[3] = {position.x(), position.y(), position.z()};
    check(pmaf_set_real_position(handle(), p), "setRealEEAgentPosition");
  }
  void setEEAgentPositions(const Eigen::Vector3d &) { unsupported("setEEAgentPositions"); }
  void setInitialEEPositions(const Eigen::Vector3d &position) { setInitialPosition(position); }
  void setInitialPosition(const Eigen::Vector3d &position) {
    const double p[3] = {position.x(), position.y(), position.z()};
    paths_cached_ = false;
    check(pmaf_set_initial_position(handle(), p), "setInitialPosition");
  }
  void setEEAgentPosAndVels(const Eigen::Vector3d &, const Eigen::Vector3d &) { unsupported("setEEAgentPosAndVels"); }
  void resetEEAgents(const Eigen::Vector3d &position, const Eigen::Vector3d &velocity,
                     const std::vector<Obstacle> &obstacles) {
    const Flat o(obstacles);
    const double p[3] = {position.x(), position.y(), position.z()}, v[3] = {velocity.x(), velocity.y(), velocity.z()};
    paths_cached_ = false;
    check(pmaf_reset_agents(handle(), p, v, o.n(), o.pos.data(), o.vel.data(), o.rad.data()), "resetEEAgents");
  }

  // ---- cf_manager.h:117-136 ----
  void moveRealEEAgent(const std::vector<Obstacle> &obstacles, const double delta_t, const int steps, const int agent_id) {
    const Flat o(obstacles);
    check(pmaf_move_real_agent(handle(), o.n(), o.pos.data(), o.vel.data(), o.rad.data(), delta_t, steps, agent_id),
          "moveRealEEAgent");
  }
  void moveAgent(const std::vector<Obstacle> &, const double, const int, const int) { unsupported("moveAgent"); }
  void moveAgents(const std::vector<Obstacle> &, const double, const int = 1) { unsupported("moveAgents"); }
  void moveAgentsPar(const std::vector<Obstacle> &, const double, const int = 1) { unsupported("moveAgentsPar"); }
  double evaluatePath(const std::vector<Obstacle> &) { unsupported("evaluatePath"); }
  int evaluateAgents(const std::vector<Obstacle> &obstacles, const double k_goal_dist, const double k_path_len,
                     const double k_safe_dist, const double k_workspace,
                     const Eigen::Matrix<double, 6, 1> des_ws_limits) {
    const Flat o(obstacles);
    double ws[6];
    for (int i = 0; i < 6; ++i) ws[i] = des_ws_limits(i);
    int best = 0;
    check(pmaf_evaluate_agents(handle(), o.n(), o.pos.data(), o.vel.data(), o.rad.data(), k_goal_dist, k_path_len,
                               k_safe_dist, k_workspace, ws, &best),
          "evaluateAgents");
    return best;
  }

 private:
  Eigen::Vector3d get3(int (*fn)(pmaf_planner *, double *), const char *what) {
    double v[3];
    check(fn(handle(), v), what);
    return vec(v);
  }
};

}  // namespace cfplanner
}  // namespace ghostplanner
