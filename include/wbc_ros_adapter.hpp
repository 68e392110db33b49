// wbc_ros_adapter.hpp -- the ROS/Gazebo side of the controller node, on top of wbc_b200::DogCtrl.
//
// What dogbot_controller's node does around the control cycle (reference paths relative to dogbot_controller/src/client):
//
//   reference                                                     here
//   ------------------------------------------------------------  -----------------------------------------------------
//   topics subscribed / advertised           main.cpp:265-282      Topics (names only), DogbotNode (ROS glue, WBC_WITH_ROS)
//   jointStateCallback: name -> DoF map,     main.cpp:388-414      JointMap::build / gather
//     positions and velocities by DoF id
//   modelStateCallback: quaternion -> R,     main.cpp:418-456      model_state_to_base
//     roll/pitch/yaw, _world_H_base, twist
//   ee??_cb: contact flag and first force    main.cpp:794-834      ContactSample::update
//   publish_cmd: torque reorder              main.cpp:768-779      JointMap::command_order
//
// The part above the message types is plain C++ (this image has no ROS): it is what tests/cpp/ros_adapter_host.cpp exercises.
// The node itself (subscribers, publishers, one step of the control loop) is the block under WBC_WITH_ROS at the end; it needs
// roscpp, sensor_msgs, gazebo_msgs, std_msgs, geometry_msgs.  It has NOT been compiled against ROS here; tests/cpp/ros_node_host.cpp
// compiles and runs it against a minimal stand-in for those headers (tests/cpp/mock_ros): messages in on the reference's topics,
// command and estimate out, checked against DogCtrl driven directly.  INTEGRATION.md section 5 walks through it.
//
// tf is not vendored in the reference: quaternion -> matrix and getRPY below restate tf/LinearMath/Matrix3x3.h (setRotation,
// getEulerYPR solution 1) from its published source; SURVEY.md Appendix D records the convention (fixed-axis XYZ,
// R = Rz(yaw) Ry(pitch) Rx(roll)) and tests/test_ros_adapter.py checks it against scipy's Rotation.
#ifndef WBC_ROS_ADAPTER_HPP
#define WBC_ROS_ADAPTER_HPP

#include <cmath>
#include <string>
#include <vector>

namespace wbc_b200 {
namespace ros_adapter {

// main.cpp:265-282
struct Topics {
    static const char* joint_states() { return "/dogbot/joint_states"; }
    static const char* model_states() { return "/gazebo/model_states"; }
    static const char* contact_back_left() { return "/dogbot/back_left_contactsensor_state"; }
    static const char* contact_front_left() { return "/dogbot/front_left_contactsensor_state"; }
    static const char* contact_back_right() { return "/dogbot/back_right_contactsensor_state"; }
    static const char* contact_front_right() { return "/dogbot/front_right_contactsensor_state"; }
    static const char* command() { return "/dogbot/joint_position_controller/command"; }
    static const char* estimation() { return "estimation_ee"; }
    static const char* model_name() { return "dogbot"; }
};

// Joint names in the C ABI's DoF order (wbc_b200.h: roll BL, BR, FL, FR, then pitch, knee of BL, BR, FL, FR) -- what
// kinDynComp.getDescriptionOfDegreeOfFreedom(i) returns for the order the controller's qmin / qmax tables imply
// (main.cpp:373-384, 612-613; dogbot.urdf:180-933).
inline const char* dof_name(int id)
{
    static const char* const names[12] = {"back_left_roll_joint",   "back_right_roll_joint",  "front_left_roll_joint",  "front_right_roll_joint",
                                          "back_left_pitch_joint",  "back_left_knee_joint",   "back_right_pitch_joint", "back_right_knee_joint",
                                          "front_left_pitch_joint", "front_left_knee_joint",  "front_right_pitch_joint", "front_right_knee_joint"};
    return (id >= 0 && id < 12) ? names[id] : "";
}

// _id2index / _index2id (main.cpp:236-239): DoF id <-> position in the JointState message.
class JointMap {
public:
    JointMap() : ready_(false)
    {
        for (int i = 0; i < 12; i++) id2index_[i] = index2id_[i] = -1;
    }
    // First message only (main.cpp:390-404): look every DoF name up in msg.name.  Returns false (and stays unready) when a DoF is
    // missing -- the reference would throw out of unordered_map::at at line 407.
    bool build(const std::vector<std::string>& msg_names)
    {
        if (ready_) return true;
        int id2[12], idx2[12];
        for (int i = 0; i < 12; i++) id2[i] = idx2[i] = -1;
        for (int i = 0; i < 12; i++) {
            size_t index = 0;
            while (index < msg_names.size() && msg_names[index] != dof_name(i)) index++;
            if (index == msg_names.size()) return false;
            id2[i] = (int)index;
            if (index < 12) idx2[index] = i;
        }
        for (int i = 0; i < 12; i++) { id2index_[i] = id2[i]; index2id_[i] = idx2[i]; }
        ready_ = true;
        return true;
    }
    bool ready() const { return ready_; }
    int index_of(int id) const { return id2index_[id]; }
    int id_of(int index) const { return index2id_[index]; }
    // _jnt_pos(i) = msg.position[_id2index.at(i)] (main.cpp:406-412); the same for velocities
    template <class Msg, class Out> void gather(const Msg& msg_values, Out& out12) const
    {
        for (int i = 0; i < 12; i++) out12[i] = msg_values[id2index_[i]];
    }
    // publish_cmd (main.cpp:768-779): data.push_back(tau(_index2id.at(i))) for i = 11 .. 0 -- the command array is the torque
    // vector in REVERSE message order.  Needs the twelve DoFs to occupy message positions 0..11 (as the reference does).
    template <class Tau, class Out> bool command_order(const Tau& tau12, Out& data12) const
    {
        for (int k = 0; k < 12; k++) {
            const int id = index2id_[11 - k];
            if (id < 0) return false;
            data12[k] = tau12[id];
        }
        return true;
    }

private:
    bool ready_;
    int id2index_[12], index2id_[12];
};

// What modelStateCallback leaves in the members update() and the cycle read (main.cpp:418-456).
struct BaseState {
    double world_H_base[16];   // row-major 4x4, rotation from the normalised quaternion, translation = position
    double base_pos[6];        // x y z roll pitch yaw                (_base_pos, main.cpp:443)
    double base_vel[6];        // linear, angular, as Gazebo reports  (_base_vel, main.cpp:453)
};

// tf::Quaternion::normalize + tf::Matrix3x3(q) + getRPY (main.cpp:432-441)
inline void quaternion_to_rotation_rpy(double qx, double qy, double qz, double qw, double R[9], double rpy[3])
{
    const double len = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    qx /= len; qy /= len; qz /= len; qw /= len;
    const double d = qx * qx + qy * qy + qz * qz + qw * qw, s = 2.0 / d;
    const double xs = qx * s, ys = qy * s, zs = qz * s;
    const double wx = qw * xs, wy = qw * ys, wz = qw * zs, xx = qx * xs, xy = qx * ys, xz = qx * zs, yy = qy * ys, yz = qy * zs, zz = qz * zs;
    R[0] = 1.0 - (yy + zz); R[1] = xy - wz;         R[2] = xz + wy;
    R[3] = xy + wz;         R[4] = 1.0 - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy;         R[7] = yz + wx;         R[8] = 1.0 - (xx + yy);
    double roll, pitch, yaw;
    if (std::fabs(R[6]) >= 1.0) {          // gimbal lock: yaw = 0, roll from the remaining difference of angles
        yaw = 0.0;
        const double delta = std::atan2(R[7], R[8]);
        pitch = R[6] < 0.0 ? M_PI / 2.0 : -M_PI / 2.0;
        roll = delta;
    } else {
        pitch = -std::asin(R[6]);
        const double c = std::cos(pitch);
        roll = std::atan2(R[7] / c, R[8] / c);
        yaw = std::atan2(R[3] / c, R[0] / c);
    }
    rpy[0] = roll; rpy[1] = pitch; rpy[2] = yaw;
}

inline void model_state_to_base(const double position[3], const double orientation_xyzw[4], const double twist_linear[3],
                                const double twist_angular[3], BaseState& out)
{
    double R[9], rpy[3];
    quaternion_to_rotation_rpy(orientation_xyzw[0], orientation_xyzw[1], orientation_xyzw[2], orientation_xyzw[3], R, rpy);
    for (int k = 0; k < 16; k++) out.world_H_base[k] = 0.0;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) out.world_H_base[4 * i + j] = R[3 * i + j];
        out.world_H_base[4 * i + 3] = position[i];
        out.base_pos[i] = position[i];
        out.base_pos[3 + i] = rpy[i];
        out.base_vel[i] = twist_linear[i];
        out.base_vel[3 + i] = twist_angular[i];
    }
    out.world_H_base[15] = 1.0;
}

// ee??_cb (main.cpp:794-834): an empty states array clears the contact flag and KEEPS the last force; otherwise the force of the
// first state is taken.
struct ContactSample {
    bool contact;
    double force[3];
    ContactSample() : contact(false) { force[0] = force[1] = force[2] = 0.0; }
    void update(int n_states, const double* first_force_xyz)
    {
        if (n_states <= 0) { contact = false; return; }
        contact = true;
        for (int k = 0; k < 3; k++) force[k] = first_force_xyz[k];
    }
};

}  // namespace ros_adapter
}  // namespace wbc_b200

#ifdef WBC_WITH_ROS
// ---------------------------------------------------------------------------------------------------------------------------
// The node (this repository's image has no ROS: exercised against tests/cpp/mock_ros only).  One DogCtrl, the reference's topics.
#include <gazebo_msgs/ContactsState.h>
#include <gazebo_msgs/ModelStates.h>
#include <geometry_msgs/WrenchStamped.h>
#include <ros/ros.h>
#include <sensor_msgs/JointState.h>
#include <std_msgs/Float64MultiArray.h>

#include "wbc_dogctrl.hpp"

namespace wbc_b200 {
namespace ros_adapter {

class DogbotNode {
public:
    explicit DogbotNode(int device = 0) : dc_(device), have_joints_(false), have_base_(false)
    {
        typedef Topics T;
        joint_sub_ = nh_.subscribe(T::joint_states(), 1, &DogbotNode::joint_cb, this);
        model_sub_ = nh_.subscribe(T::model_states(), 1, &DogbotNode::model_cb, this);
        bl_sub_ = nh_.subscribe(T::contact_back_left(), 1, &DogbotNode::eebl_cb, this);        // main.cpp:268-271
        fl_sub_ = nh_.subscribe(T::contact_front_left(), 1, &DogbotNode::eefl_cb, this);
        br_sub_ = nh_.subscribe(T::contact_back_right(), 1, &DogbotNode::eebr_cb, this);
        fr_sub_ = nh_.subscribe(T::contact_front_right(), 1, &DogbotNode::eefr_cb, this);
        cmd_pub_ = nh_.advertise<std_msgs::Float64MultiArray>(T::command(), 1);
        est_pub_ = nh_.advertise<geometry_msgs::WrenchStamped>(T::estimation(), 1);
    }
    DogCtrl& controller() { return dc_; }
    bool ready() const { return have_joints_ && have_base_; }
    // One control period: update() from the latest messages, the cycle for `mode`, publish_cmd and the estimate.
    // The caller sets the desired CoM / swing samples on controller() first (or hands the plan over once and passes sampled = true).
    void step(int mode, bool sampled = false)
    {
        struct V { const double* p; double operator()(int i) const { return p[i]; } };
        struct M4 { const double* p; double operator()(int i, int j) const { return p[4 * i + j]; } };
        const double g[3] = {0.0, 0.0, -9.8};                                   // main.cpp:855
        dc_.update(M4{base_.world_H_base}, V{q_}, V{dq_}, V{base_.base_vel}, V{g});
        dc_.set_base_rpy(base_.base_pos[3], base_.base_pos[4], base_.base_pos[5]);
        dc_.set_foot_forces(V{foot_[0].force}, V{foot_[1].force}, V{foot_[2].force}, V{foot_[3].force});   // BR, BL, FL, FR
        if (mode == WBC_MODE_STANCE) dc_.cycle_stance(sampled); else dc_.cycle_swing(mode == WBC_MODE_SWING_BR_FL, sampled);
        std_msgs::Float64MultiArray cmd;
        cmd.data.resize(12);
        if (map_.command_order(dc_.tau(), cmd.data)) cmd_pub_.publish(cmd);      // main.cpp:768-779
        geometry_msgs::WrenchStamped est;                                        // main.cpp:1129-1144
        est.header.stamp = ros::Time::now();
        est.wrench.force.x = dc_.w()[0]; est.wrench.force.y = dc_.w()[1]; est.wrench.force.z = dc_.w()[2];
        est.wrench.torque.x = dc_.w()[3]; est.wrench.torque.y = dc_.w()[4]; est.wrench.torque.z = dc_.w()[5];
        est_pub_.publish(est);
    }

private:
    void joint_cb(const sensor_msgs::JointState& msg)
    {
        if (!map_.build(msg.name)) return;
        map_.gather(msg.position, q_);
        map_.gather(msg.velocity, dq_);
        have_joints_ = true;
    }
    void model_cb(const gazebo_msgs::ModelStates& msg)
    {
        for (size_t k = 0; k < msg.name.size(); k++) {
            if (msg.name[k] != Topics::model_name()) continue;
            const double p[3] = {msg.pose[k].position.x, msg.pose[k].position.y, msg.pose[k].position.z};
            const double o[4] = {msg.pose[k].orientation.x, msg.pose[k].orientation.y, msg.pose[k].orientation.z, msg.pose[k].orientation.w};
            const double v[3] = {msg.twist[k].linear.x, msg.twist[k].linear.y, msg.twist[k].linear.z};
            const double w[3] = {msg.twist[k].angular.x, msg.twist[k].angular.y, msg.twist[k].angular.z};
            model_state_to_base(p, o, v, w, base_);
            have_base_ = true;
            return;
        }
    }
    // stacked foot order BR, BL, FL, FR (main.cpp:1022-1026)
    void eebr_cb(const gazebo_msgs::ContactsStateConstPtr& m) { contact_cb(m, 0); }
    void eebl_cb(const gazebo_msgs::ContactsStateConstPtr& m) { contact_cb(m, 1); }
    void eefl_cb(const gazebo_msgs::ContactsStateConstPtr& m) { contact_cb(m, 2); }
    void eefr_cb(const gazebo_msgs::ContactsStateConstPtr& m) { contact_cb(m, 3); }
    void contact_cb(const gazebo_msgs::ContactsStateConstPtr& m, int stacked_foot)
    {
        double f[3] = {0.0, 0.0, 0.0};
        if (!m->states.empty()) { f[0] = m->states[0].total_wrench.force.x; f[1] = m->states[0].total_wrench.force.y; f[2] = m->states[0].total_wrench.force.z; }
        foot_[stacked_foot].update((int)m->states.size(), f);
    }
    ros::NodeHandle nh_;
    ros::Subscriber joint_sub_, model_sub_, bl_sub_, fl_sub_, br_sub_, fr_sub_;
    ros::Publisher cmd_pub_, est_pub_;
    DogCtrl dc_;
    JointMap map_;
    BaseState base_;
    ContactSample foot_[4];      // stacked order BR, BL, FL, FR (main.cpp:1022-1026)
    double q_[12], dq_[12];
    bool have_joints_, have_base_;
};

}  // namespace ros_adapter
}  // namespace wbc_b200
#endif  // WBC_WITH_ROS

#endif
