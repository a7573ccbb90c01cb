#pragma once  // TEST STUB (syntax check only)
#include <std_msgs/String.h>
#include <vector>
namespace bimanual_planning_ros { struct ControllerParams { double velocity = 0, low_level_gain = 0; bool switching = false; std::vector<std_msgs::String> controllers; std::vector<double> gains; }; }
