#pragma once  // TEST STUB (syntax check only)
#include <std_msgs/Header.h>
namespace franka_msgs { struct ErrorRecoveryActionGoal { std_msgs::Header header; }; }
