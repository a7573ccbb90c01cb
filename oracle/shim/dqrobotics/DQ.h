// TEST INFRASTRUCTURE — not product code.
//
// Empty stand-in for the dqrobotics headers (stable PPA, version unpinned,
// /root/reference/README.md:31) that obstacle.h:6,9 and helper_functions.h:4-6
// include. The rollout translation units use none of dqrobotics' symbols; they
// only rely on (a) the class names below existing and (b) DQ.h leaking
// `using namespace Eigen;` at global scope (cf_agent.h:343 writes an
// unqualified `Vector3d`).
#pragma once
#include "eigen3/Eigen/Dense"

using namespace Eigen;

namespace DQ_robotics {
class DQ {};
class DQ_Kinematics {};
class DQ_SerialManipulator {};
class DQ_CooperativeDualTaskSpace {};
}  // namespace DQ_robotics

using namespace DQ_robotics;

class DQ_VrepInterface {};
