#pragma once  // TEST STUB (syntax check only)
#include <array>
namespace bimanual_planning_ros { struct ContactWrench { std::array<double, 6> B_F_ext{}; }; }
