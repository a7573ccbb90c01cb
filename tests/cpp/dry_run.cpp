// Host program in the reference's language (C++) driving the drop-in CfManager exactly as
// PandaBimanualPlanning does (taskCallback + planCallback, panda_bimanual_control.cpp:329-369,
// :494-521) on the dual_arms_static1 task, with the dry_run.launch relay closed in-process.
// Prints one line per tick: best agent index, next position, next velocity (hex floats).
#include <pmaf/cf_manager.hpp>

#include <cstdio>
#include <cstdlib>
#include <fstream>

using ghostplanner::cfplanner::CfManager;
using ghostplanner::cfplanner::Obstacle;
using Eigen::Vector3d;

int main(int argc, char **argv) {
  const int ticks = argc > 1 ? std::atoi(argv[1]) : 10;
  const int horizon = argc > 2 ? std::atoi(argv[2]) : 1500;
  // config/tasks/dual_arms_static1.yaml
  const double xs[10][3] = {{0.125, 0.0, 1.0}, {0.125, 0.125, 1.0}, {0.125, -0.125, 1.0}, {0.125, 0.0, 0.7},
                            {0.125, 0.125, 0.7}, {0.125, -0.125, 0.7}, {-0.35, 0.0, 0.6}, {-0.35, 0.125, 0.6},
                            {-0.35, -0.125, 0.6}, {100.0, 100.0, 100.0}};
  std::vector<Obstacle> obstacles;
  for (auto &x : xs) obstacles.emplace_back(Vector3d(x[0], x[1], x[2]), Vector3d(0, 0, 0), 0.1);
  const int num_agents = 10;
  const std::vector<double> k_attr(num_agents, 4.0), k_circ(num_agents, 0.025), k_repel(num_agents, 0.08),
      k_damp(num_agents, 3.0), k_manip(num_agents, 0.0), k_repel_body(1, 0.02);
  const double dt = 1.0 / 100.0;
  Eigen::Matrix<double, 6, 1> ws;
  ws << 1.0, -1.0, 0.3, -0.3, 1.1, 0.2;

  CfManager cf;
  cf.seedRandomVectors(1);
  const Vector3d start(-0.6, 0.0, 0.65), goal(0.5, 0.0, 0.7);
  cf.init(goal, dt, obstacles, k_attr, k_circ, k_repel, k_damp, k_manip, k_repel_body, 0.2, 0.25, 0.35, horizon, 1);
  if (argc > 3) {  // RandomCfAgent vectors from a file of raw doubles [agents][obstacles][3]
    std::vector<double> vecs((size_t)num_agents * obstacles.size() * 3);
    std::ifstream f(argv[3], std::ios::binary);
    f.read(reinterpret_cast<char *>(vecs.data()), (std::streamsize)(vecs.size() * sizeof(double)));
    if (!f) return 2;
    cf.setRandomVectors(vecs);
  }
  cf.setInitialPosition(start);
  for (int t = 0; t < ticks; ++t) {
    cf.stopPrediction();
    const int best = cf.evaluateAgents(obstacles, 100.0, 10.0, 0.001, 1.0, ws);
    cf.moveRealEEAgent(obstacles, dt, 1, best);
    const Vector3d p = cf.getNextPosition(), v = cf.getNextVelocity();
    cf.resetEEAgents(p, v, obstacles);
    cf.startPrediction();
    std::printf("%d %a %a %a %a %a %a\n", best, p.x(), p.y(), p.z(), v.x(), v.y(), v.z());
  }
  cf.stopPrediction();
  std::printf("type %d dist %a steps0 %d traj %zu\n", cf.getBestAgentType(), cf.getDistFromGoal(),
              cf.getNumPredictionSteps(0), cf.getPlannedTrajectory().size());
  return 0;
}
