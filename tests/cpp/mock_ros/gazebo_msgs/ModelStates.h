#pragma once
#include <string>
#include <vector>
#include "geometry_msgs/WrenchStamped.h"
namespace gazebo_msgs { struct ModelStates { std::vector<std::string> name; std::vector<geometry_msgs::Pose> pose; std::vector<geometry_msgs::Twist> twist; }; }
