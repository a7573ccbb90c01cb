#pragma once  // TEST STUB (syntax check only)
#include <geometry_msgs/PoseStamped.h>
#include <vector>
namespace std_msgs { struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, POINTS = 8, MESH_RESOURCE = 10, ADD = 0, MODIFY = 0, DELETE = 2 };
  std_msgs::Header header;
  std::string ns, mesh_resource;
  int id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  struct { double x = 0, y = 0, z = 0; } scale;
  std_msgs::ColorRGBA color;
  ros::Duration lifetime;
  bool frame_locked = false, mesh_use_embedded_materials = false;
  std::vector<geometry_msgs::Point> points;
  std::vector<std_msgs::ColorRGBA> colors;
};
}  // namespace visualization_msgs
