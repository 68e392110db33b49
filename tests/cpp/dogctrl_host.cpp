// Exercises include/wbc_dogctrl.hpp the way ctrl_loop() would (main.cpp:978-1127, 1155-1397), with tiny Eigen-like
// stand-ins (Eigen is not installed here).  Usage:
//   dogctrl_host probe                 -> prints "nodev" if no sm_100 device (the library has no CPU fallback), else "ok"
//   dogctrl_host cycle  IN.bin OUT.bin -> IN: n (int32) then per cycle 112 doubles + mode; OUT: per cycle tau12 w6 x30 obj status
//   dogctrl_host traj   IN.bin OUT.bin -> same IN; per cycle a one-polynomial plan starting at the record's desired pose is handed to
//                                         the GPU and sampled at t = 0.1: OUT per cycle tau12 (host-fed samples) tau12 (device samples)
//   dogctrl_host opt    IN.bin OUT.bin -> IN: nrows (int32), Q 900, c 30, L nrows*31; OUT: x 30
#include <cstdio>
#include <cstring>
#include <vector>

#include "wbc_dogctrl.hpp"

template <int R, int Cc> struct Mat {
    double a[R * Cc];
    double& operator()(int i, int j) { return a[i * Cc + j]; }
    double operator()(int i, int j) const { return a[i * Cc + j]; }
    double& operator()(int i) { return a[i]; }
    double operator()(int i) const { return a[i]; }
};
struct DynMat {
    int r, c;
    std::vector<double> a;
    DynMat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
    double& operator()(int i) { return a[i]; }
    double operator()(int i) const { return a[i]; }
};

static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    try {
        if (!strcmp(argv[1], "probe")) {
            try {
                wbc_b200::DogCtrl dc;
                puts("ok");
            } catch (const wbc_b200::Error& e) {
                if (e.code() != WBC_ENODEV) throw;
                puts("nodev");
            }
            return 0;
        }
        if (argc < 4) return 2;
        FILE* fi = fopen(argv[2], "rb");
        FILE* fo = fopen(argv[3], "wb");
        if (!fi || !fo) return 3;
        if (!strcmp(argv[1], "cycle")) {
            int n = 0;
            if (!rd(fi, &n, 4)) return 3;
            wbc_b200::DogCtrl dc;
            for (int it = 0; it < n; it++) {
                // record: H 16 | q 12 | dq 12 | basevel 6 | gravity 3 | rpy 3 | com p,v,a 18 | forces 12 | swing p,v,a 18 | mode
                Mat<4, 4> H; Mat<12, 1> q, dq; Mat<6, 1> bv, cp, cv, ca, sp, sv, sa; Mat<3, 1> g, rpy, f[4];
                double mode_d = 0;
                bool ok = rd(fi, H.a, 128) && rd(fi, q.a, 96) && rd(fi, dq.a, 96) && rd(fi, bv.a, 48) && rd(fi, g.a, 24) && rd(fi, rpy.a, 24) &&
                          rd(fi, cp.a, 48) && rd(fi, cv.a, 48) && rd(fi, ca.a, 48) && rd(fi, f[0].a, 24) && rd(fi, f[1].a, 24) && rd(fi, f[2].a, 24) &&
                          rd(fi, f[3].a, 24) && rd(fi, sp.a, 48) && rd(fi, sv.a, 48) && rd(fi, sa.a, 48) && rd(fi, &mode_d, 8);
                if (!ok) return 3;
                dc.update(H, q, dq, bv, g);
                dc.set_base_rpy(rpy(0), rpy(1), rpy(2));
                dc.set_com_desired(cp, cv, ca);
                dc.set_foot_forces(f[0], f[1], f[2], f[3]);
                const int mode = (int)mode_d;
                if (mode == WBC_MODE_STANCE) dc.cycle_stance();
                else {
                    dc.set_swing_desired(sp, sv, sa);
                    dc.cycle_swing(mode == WBC_MODE_SWING_BR_FL);
                }
                double st = dc.status(), obj = dc.qp_objective();
                fwrite(dc.tau(), 8, 12, fo); fwrite(dc.w(), 8, 6, fo); fwrite(dc.x(), 8, 30, fo); fwrite(&obj, 8, 1, fo); fwrite(&st, 8, 1, fo);
            }
        } else if (!strcmp(argv[1], "traj")) {
            int n = 0;
            if (!rd(fi, &n, 4)) return 3;
            wbc_b200::DogCtrl dc;
            for (int it = 0; it < n; it++) {
                Mat<4, 4> H; Mat<12, 1> q, dq; Mat<6, 1> bv, cp, cv, ca, sp, sv, sa; Mat<3, 1> g, rpy, f[4];
                double mode_d = 0;
                bool ok = rd(fi, H.a, 128) && rd(fi, q.a, 96) && rd(fi, dq.a, 96) && rd(fi, bv.a, 48) && rd(fi, g.a, 24) && rd(fi, rpy.a, 24) &&
                          rd(fi, cp.a, 48) && rd(fi, cv.a, 48) && rd(fi, ca.a, 48) && rd(fi, f[0].a, 24) && rd(fi, f[1].a, 24) && rd(fi, f[2].a, 24) &&
                          rd(fi, f[3].a, 24) && rd(fi, sp.a, 48) && rd(fi, sv.a, 48) && rd(fi, sa.a, 48) && rd(fi, &mode_d, 8);
                if (!ok) return 3;
                const int mode = (int)mode_d;
                // plan: one polynomial of 0.5 s per spline, from the record's desired pose / swing targets to a point 1 cm on
                double dur[4] = {0.5, 0.5, 0.5, 0.5}, nodes[4 * 2 * 6];
                for (int s_ = 0; s_ < 4; s_++)
                    for (int c = 0; c < 3; c++) {
                        const double p0 = (s_ < 2) ? cp(3 * s_ + c) : sp(3 * (s_ - 2) + c), v0 = (s_ < 2) ? cv(3 * s_ + c) : sv(3 * (s_ - 2) + c);
                        nodes[(s_ * 2 + 0) * 6 + c] = p0; nodes[(s_ * 2 + 0) * 6 + 3 + c] = v0;
                        nodes[(s_ * 2 + 1) * 6 + c] = p0 + 0.01; nodes[(s_ * 2 + 1) * 6 + 3 + c] = 0.0;
                    }
                dc.set_trajectory(1, dur, nodes);
                // (a) samples fetched to the host through the C ABI and fed back as ordinary inputs
                Mat<6, 1> hp, hv, ha, hsp, hsv, hsa;
                wbc_traj_samples out;
                out.com_des_pos = hp.a; out.com_des_vel = hv.a; out.com_des_acc = ha.a;
                out.sw_des_pos = hsp.a; out.sw_des_vel = hsv.a; out.sw_des_acc = hsa.a; out.ld = 1;
                wbc_b200::check(wbc_sample_trajectory(dc.batch().ctx(), 1, nullptr, 0.1, &out, nullptr, WBC_HOST_PTRS), "wbc_sample_trajectory");
                double tau_a[12], tau_b[12];
                for (int pass = 0; pass < 2; pass++) {
                    const double zero[6] = {0, 0, 0, 0, 0, 0};
                    dc.batch().set_observer_state(zero, zero);
                    dc.update(H, q, dq, bv, g);
                    dc.set_base_rpy(rpy(0), rpy(1), rpy(2));
                    dc.set_foot_forces(f[0], f[1], f[2], f[3]);
                    if (pass == 0) { dc.set_com_desired(hp, hv, ha); dc.set_swing_desired(hsp, hsv, hsa); }
                    else dc.sample_trajectory(0.1);                  // (b) samples stay on the device
                    if (mode == WBC_MODE_STANCE) dc.cycle_stance(pass == 1);
                    else dc.cycle_swing(mode == WBC_MODE_SWING_BR_FL, pass == 1);
                    memcpy(pass == 0 ? tau_a : tau_b, dc.tau(), sizeof(tau_a));
                }
                fwrite(tau_a, 8, 12, fo); fwrite(tau_b, 8, 12, fo);
            }
        } else if (!strcmp(argv[1], "opt")) {
            int nrows = 0;
            if (!rd(fi, &nrows, 4) || (nrows != 86 && nrows != 82)) return 3;
            DynMat Q(30, 30), c(30, 1), L(nrows, 31), x(30, 1);
            if (!rd(fi, Q.a.data(), 7200) || !rd(fi, c.a.data(), 240) || !rd(fi, L.a.data(), (size_t)nrows * 31 * 8)) return 3;
            wbc_b200::OPT o(30, 86, 82);       // main.cpp:266
            o.setQ(Q); o.setc(c);
            if (nrows == 86) { o.setL_stance(L); o.opt_stance(x); }
            else { o.setL_swing(L); o.opt_swing(x); }
            fwrite(x.a.data(), 8, 30, fo);
        } else return 2;
        fclose(fi); fclose(fo);
    } catch (const std::exception& e) {
        fprintf(stderr, "dogctrl_host: %s\n", e.what());
        return 1;
    }
    return 0;
}
