#pragma once  // TEST STUB (syntax check only)
#include <std_msgs/Header.h>
#include <vector>
namespace sensor_msgs { struct JointState { std_msgs::Header header; std::vector<std::string> name; std::vector<double> position, velocity, effort; }; }
