#pragma once  // TEST STUB (syntax check only)
namespace franka_gripper {
struct MoveGoal { double width = 0, speed = 0, force = 0; struct { double inner = 0, outer = 0; } epsilon; };
struct MoveAction {};
}
