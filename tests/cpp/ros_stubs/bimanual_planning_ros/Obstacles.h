#pragma once  // TEST STUB (syntax check only): msg/Obstacles.msg = Position[] pos, Position[] vel, float64[] radius
#include <bimanual_planning_ros/Position.h>
#include <vector>
namespace bimanual_planning_ros { struct Obstacles { std::vector<Position> pos, vel; std::vector<double> radius; }; }
