#pragma once
#include <string>
#include <vector>
namespace sensor_msgs { struct JointState { std::vector<std::string> name; std::vector<double> position, velocity, effort; }; }
