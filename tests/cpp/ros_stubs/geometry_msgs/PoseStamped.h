#pragma once  // TEST STUB (syntax check only)
#include <geometry_msgs/Point.h>
#include <std_msgs/Header.h>
namespace geometry_msgs { struct PoseStamped { std_msgs::Header header; Pose pose; }; }
