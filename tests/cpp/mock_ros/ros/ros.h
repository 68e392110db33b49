// TEST-ONLY stand-in for roscpp: just enough of NodeHandle / Subscriber / Publisher / Time for include/wbc_ros_adapter.hpp's node
// glue (WBC_WITH_ROS) to compile and run without ROS.  Subscriptions are kept in a process-wide table so that a test can deliver a
// message to a topic; publications keep their last message per topic.  Not ROS, not shipped.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeindex>

namespace mock_ros {
struct Registry {
    std::map<std::string, std::function<void(const void*)>> subs;      // topic -> callback taking the message by address
    std::map<std::string, std::shared_ptr<void>> last;                 // topic -> last published message
    std::map<std::string, int> count;
    static Registry& get() { static Registry r; return r; }
};
template <class M> bool deliver(const std::string& topic, const M& msg)
{
    auto it = Registry::get().subs.find(topic);
    if (it == Registry::get().subs.end()) return false;
    it->second(&msg);
    return true;
}
template <class M> const M* last_published(const std::string& topic)
{
    auto it = Registry::get().last.find(topic);
    return it == Registry::get().last.end() ? nullptr : static_cast<const M*>(it->second.get());
}
inline int publish_count(const std::string& topic) { return Registry::get().count[topic]; }
}  // namespace mock_ros

namespace ros {
struct Time {
    double sec;
    static Time now() { return Time{0.0}; }
};
class Subscriber {};
class Publisher {
public:
    Publisher() {}
    explicit Publisher(const std::string& t) : topic_(t) {}
    template <class M> void publish(const M& m) const
    {
        mock_ros::Registry::get().last[topic_] = std::make_shared<M>(m);
        mock_ros::Registry::get().count[topic_]++;
    }
private:
    std::string topic_;
};
class NodeHandle {
public:
    // member callback taking the message by const reference (sensor_msgs::JointState, gazebo_msgs::ModelStates in the reference)
    template <class M, class T> Subscriber subscribe(const std::string& topic, int, void (T::*fn)(const M&), T* obj)
    {
        mock_ros::Registry::get().subs[topic] = [obj, fn](const void* p) { (obj->*fn)(*static_cast<const M*>(p)); };
        return Subscriber();
    }
    // member callback taking a shared pointer to a const message (gazebo_msgs::ContactsStateConstPtr in the reference)
    template <class M, class T> Subscriber subscribe(const std::string& topic, int, void (T::*fn)(const std::shared_ptr<const M>&), T* obj)
    {
        mock_ros::Registry::get().subs[topic] = [obj, fn](const void* p) { (obj->*fn)(std::make_shared<const M>(*static_cast<const M*>(p))); };
        return Subscriber();
    }
    template <class M> Publisher advertise(const std::string& topic, int) { return Publisher(topic); }
};
}  // namespace ros
