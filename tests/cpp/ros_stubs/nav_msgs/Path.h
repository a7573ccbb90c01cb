#pragma once  // TEST STUB (syntax check only)
#include <geometry_msgs/PoseStamped.h>
#include <vector>
namespace nav_msgs { struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; }; }
