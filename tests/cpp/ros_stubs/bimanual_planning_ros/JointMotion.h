#pragma once  // TEST STUB (syntax check only)
#include <array>
#include <vector>
namespace bimanual_planning_ros { struct JointMotion { struct Request { std::array<double, 14> goal{}; double v = 0; } request; struct Response { bool success = false; } response; }; }
