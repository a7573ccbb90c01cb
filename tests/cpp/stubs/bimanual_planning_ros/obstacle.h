// TEST STUB — stands in for the reference package's bimanual_planning_ros/obstacle.h (which needs
// dqrobotics) so that include/pmaf/cf_manager.hpp can be compiled outside a catkin workspace.
// It declares only the members the façade uses.
#pragma once
#include <string>

#include "eigen3/Eigen/Dense"

namespace ghostplanner {
namespace cfplanner {
class Obstacle {
  Eigen::Vector3d pos_, vel_;
  double rad_ = 0.0;

 public:
  Obstacle() = default;
  Obstacle(const Eigen::Vector3d pos, const Eigen::Vector3d vel, const double rad) : pos_(pos), vel_(vel), rad_(rad) {}
  Eigen::Vector3d getPosition() const { return pos_; }
  Eigen::Vector3d getVelocity() const { return vel_; }
  double getRadius() const { return rad_; }
  void setPosition(Eigen::Vector3d p) { pos_ = p; }
  void setVelocity(Eigen::Vector3d v) { vel_ = v; }
};
}  // namespace cfplanner
}  // namespace ghostplanner
