#pragma once  // TEST STUB (syntax check only): msg/Position.msg = float64[3] data
#include <array>
namespace bimanual_planning_ros { struct Position { std::array<double, 3> data{}; }; }
