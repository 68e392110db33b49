// Runs include/wbc_ros_adapter.hpp's node glue (WBC_WITH_ROS) against the stand-in ROS headers of tests/cpp/mock_ros:
// messages are delivered on the reference's topics (main.cpp:265-282), one control step is taken, and the published command
// and estimate are compared with a DogCtrl driven directly with the same numbers.  Usage:
//   ros_node_host probe   -> "nodev" without an sm_100 device (no CPU fallback), else "ok"
//   ros_node_host run     -> prints "cmd_err <max abs error> est_err <max abs error> published <count>"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define WBC_WITH_ROS 1
#include "wbc_ros_adapter.hpp"

using namespace wbc_b200;
using namespace wbc_b200::ros_adapter;

struct V { const double* p; double operator()(int i) const { return p[i]; } };
struct M4 { const double* p; double operator()(int i, int j) const { return p[4 * i + j]; } };

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    try {
        if (!strcmp(argv[1], "probe")) {
            try { DogbotNode node; puts("ok"); }
            catch (const Error& e) { if (e.code() != WBC_ENODEV) throw; puts("nodev"); }
            return 0;
        }
        DogbotNode node;
        // joint_states as gazebo_ros_control publishes them: alphabetical by joint name
        std::vector<std::string> names;
        for (int i = 0; i < 12; i++) names.push_back(dof_name(i));
        std::sort(names.begin(), names.end());
        const double qnom[12] = {0.000488, 0.000624, -3.2e-05, -0.000513, -0.88425, -1.60390, 0.88620, 1.60326, -0.88481, -1.60396, 0.88635, 1.60361};
        sensor_msgs::JointState js;
        js.name = names;
        js.position.resize(12); js.velocity.resize(12);
        double q[12], dq[12];
        for (int i = 0; i < 12; i++) { q[i] = qnom[i] + 0.02 * std::sin(1.0 + i); dq[i] = 0.1 * std::cos(2.0 + i); }
        for (size_t k = 0; k < names.size(); k++)
            for (int i = 0; i < 12; i++)
                if (names[k] == dof_name(i)) { js.position[k] = q[i]; js.velocity[k] = dq[i]; }
        gazebo_msgs::ModelStates ms;
        ms.name = {"ground_plane", "dogbot"};
        ms.pose.resize(2); ms.twist.resize(2);
        ms.pose[1].position.x = 0.3; ms.pose[1].position.y = -0.2; ms.pose[1].position.z = 0.43;
        ms.pose[1].orientation.x = 0.01; ms.pose[1].orientation.y = -0.02; ms.pose[1].orientation.z = 0.05; ms.pose[1].orientation.w = 2.0;   // unnormalised
        ms.twist[1].linear.x = 0.05; ms.twist[1].linear.y = -0.02; ms.twist[1].angular.z = 0.03;
        gazebo_msgs::ContactsState cs[4];      // BR, BL, FL, FR
        for (int f = 0; f < 4; f++) {
            cs[f].states.resize(1);
            cs[f].states[0].total_wrench.force.x = 1.0 + f; cs[f].states[0].total_wrench.force.y = -0.5 * f; cs[f].states[0].total_wrench.force.z = 50.0 + 2.0 * f;
        }
        bool ok = mock_ros::deliver(Topics::joint_states(), js) && mock_ros::deliver(Topics::model_states(), ms) &&
                  mock_ros::deliver(Topics::contact_back_right(), cs[0]) && mock_ros::deliver(Topics::contact_back_left(), cs[1]) &&
                  mock_ros::deliver(Topics::contact_front_left(), cs[2]) && mock_ros::deliver(Topics::contact_front_right(), cs[3]);
        if (!ok || !node.ready()) { puts("not ready"); return 1; }
        const double cpos[6] = {0.3, -0.2, 0.40, 0.0, 0.0, 0.05}, zero6[6] = {0, 0, 0, 0, 0, 0};
        node.controller().set_com_desired(V{cpos}, V{zero6}, V{zero6});
        node.step(WBC_MODE_STANCE);

        // the same cycle on a DogCtrl of its own
        DogCtrl dc;
        const double p[3] = {0.3, -0.2, 0.43}, o[4] = {0.01, -0.02, 0.05, 2.0}, lin[3] = {0.05, -0.02, 0.0}, ang[3] = {0.0, 0.0, 0.03};
        BaseState b;
        model_state_to_base(p, o, lin, ang, b);
        const double g[3] = {0.0, 0.0, -9.8};
        dc.update(M4{b.world_H_base}, V{q}, V{dq}, V{b.base_vel}, V{g});
        dc.set_base_rpy(b.base_pos[3], b.base_pos[4], b.base_pos[5]);
        double ff[4][3];
        for (int f = 0; f < 4; f++) { ff[f][0] = 1.0 + f; ff[f][1] = -0.5 * f; ff[f][2] = 50.0 + 2.0 * f; }
        dc.set_foot_forces(V{ff[0]}, V{ff[1]}, V{ff[2]}, V{ff[3]});
        dc.set_com_desired(V{cpos}, V{zero6}, V{zero6});
        dc.cycle_stance();
        const std_msgs::Float64MultiArray* cmd = mock_ros::last_published<std_msgs::Float64MultiArray>(Topics::command());
        const geometry_msgs::WrenchStamped* est = mock_ros::last_published<geometry_msgs::WrenchStamped>(Topics::estimation());
        if (!cmd || !est || cmd->data.size() != 12) { puts("nothing published"); return 1; }
        // publish_cmd (main.cpp:768-779): data[k] = tau(id of message position 11 - k)
        double cmd_err = 0.0, tmax = 0.0;
        for (int k = 0; k < 12; k++) {
            int id = -1;
            for (int i = 0; i < 12; i++) if (names[11 - k] == dof_name(i)) id = i;
            cmd_err = std::fmax(cmd_err, std::fabs(cmd->data[k] - dc.tau()[id]));
            tmax = std::fmax(tmax, std::fabs(dc.tau()[id]));
        }
        const double e6[6] = {est->wrench.force.x, est->wrench.force.y, est->wrench.force.z, est->wrench.torque.x, est->wrench.torque.y, est->wrench.torque.z};
        double est_err = 0.0;
        for (int a = 0; a < 6; a++) est_err = std::fmax(est_err, std::fabs(e6[a] - dc.w()[a]));
        printf("cmd_err %.3e est_err %.3e published %d status %d tau_max %.3f\n", cmd_err, est_err, mock_ros::publish_count(Topics::command()), dc.status(), tmax);
        return 0;
    } catch (const std::exception& e) {
        printf("error: %s\n", e.what());
        return 1;
    }
}
