// Exercises include/wbc_ros_adapter.hpp's ROS-free core (no GPU, no ROS).  Usage:
//   ros_adapter_host joints NAMES.txt   -> NAMES.txt: message joint names, one per line.  The message carries position[k] = 100 + k,
//                                           velocity[k] = 200 + k, and the controller's torque of DoF id is 300 + id.
//                                           prints: "q" 12 values, "dq" 12 values, "cmd" 12 values (or "incomplete")
//   ros_adapter_host pose x y z qx qy qz qw vx vy vz wx wy wz -> prints H (16), base_pos (6), base_vel (6)
//   ros_adapter_host contact            -> prints the contact flag / force after: a force message, an empty message, another force
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "wbc_ros_adapter.hpp"

using namespace wbc_b200::ros_adapter;

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    if (!strcmp(argv[1], "joints") && argc >= 3) {
        std::vector<std::string> names;
        std::ifstream f(argv[2]);
        std::string s;
        while (std::getline(f, s)) if (!s.empty()) names.push_back(s);
        std::vector<double> pos(names.size()), vel(names.size());
        for (size_t k = 0; k < names.size(); k++) { pos[k] = 100.0 + k; vel[k] = 200.0 + k; }
        JointMap m;
        if (!m.build(names)) { puts("incomplete"); return 0; }
        double q[12], dq[12], tau[12], cmd[12];
        m.gather(pos, q);
        m.gather(vel, dq);
        for (int i = 0; i < 12; i++) tau[i] = 300.0 + i;
        printf("q");  for (int i = 0; i < 12; i++) printf(" %.17g", q[i]);  printf("\n");
        printf("dq"); for (int i = 0; i < 12; i++) printf(" %.17g", dq[i]); printf("\n");
        if (m.command_order(tau, cmd)) { printf("cmd"); for (int i = 0; i < 12; i++) printf(" %.17g", cmd[i]); printf("\n"); }
        else puts("cmd unavailable");
        return 0;
    }
    if (!strcmp(argv[1], "pose") && argc >= 15) {
        double v[13];
        for (int k = 0; k < 13; k++) v[k] = atof(argv[2 + k]);
        BaseState b;
        model_state_to_base(v, v + 3, v + 7, v + 10, b);
        for (int k = 0; k < 16; k++) printf("%.17g ", b.world_H_base[k]);
        for (int k = 0; k < 6; k++) printf("%.17g ", b.base_pos[k]);
        for (int k = 0; k < 6; k++) printf("%.17g ", b.base_vel[k]);
        printf("\n");
        return 0;
    }
    if (!strcmp(argv[1], "contact")) {
        ContactSample c;
        const double f1[3] = {1.0, 2.0, 30.0}, f2[3] = {-1.0, 0.5, 25.0};
        printf("%d %g %g %g\n", (int)c.contact, c.force[0], c.force[1], c.force[2]);
        c.update(2, f1); printf("%d %g %g %g\n", (int)c.contact, c.force[0], c.force[1], c.force[2]);
        c.update(0, nullptr); printf("%d %g %g %g\n", (int)c.contact, c.force[0], c.force[1], c.force[2]);
        c.update(1, f2); printf("%d %g %g %g\n", (int)c.contact, c.force[0], c.force[1], c.force[2]);
        return 0;
    }
    return 2;
}
