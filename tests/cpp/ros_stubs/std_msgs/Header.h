#pragma once  // TEST STUB (syntax check only)
#include <ros/ros.h>
namespace std_msgs { struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; }; }
