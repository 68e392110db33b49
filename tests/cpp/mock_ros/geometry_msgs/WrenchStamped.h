#pragma once
#include "ros/ros.h"
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct Twist { Vector3 linear, angular; };
struct Wrench { Vector3 force, torque; };
struct Header { ros::Time stamp; };
struct WrenchStamped { Header header; Wrench wrench; };
}
