#pragma once
#include <memory>
#include <vector>
#include "geometry_msgs/WrenchStamped.h"
namespace gazebo_msgs {
struct ContactState { geometry_msgs::Wrench total_wrench; };
struct ContactsState { std::vector<ContactState> states; };
typedef std::shared_ptr<const ContactsState> ContactsStateConstPtr;      // boost::shared_ptr in ROS 1
}
