// On-device sampling of the planner's solution splines.  The reference samples four towr::Spline objects
// at wall-clock time t every control cycle on the host (main.cpp:1004-1010 base_linear_/base_angular_, 1333-1368
// ee_motion_ of the two swing feet) and the 36 resulting doubles per robot are inputs of the cycle.  Here the node tables
// of the splines live in HBM (uploaded once per plan, i.e. once per half gait cycle) and a thread per instance
// evaluates them, so a cycle needs only t.
//
// A towr::Spline is a chain of CubicHermitePolynomial segments (spline.cc:30-46); segment j runs between node j and
// node j+1 (position p and velocity v per dimension) and lasts durations[j].
//   * segment of a global time: the first j with  sum_{i<=j} T_i >= t - 1e-10  ("at junctions, returns previous
//     spline"), spline.cc:48-66; local time = t - sum_{i<j} T_i, subtracted one duration at a time (68-79).
//     The reference asserts on t < 0 and on t beyond the last junction; here t < 0 is evaluated as 0 and t beyond the
//     end evaluates the last polynomial at its local time (extrapolation) -- the caller re-plans before that.
//   * coefficients (polynomial.cc:98-104):  a = p0, b = v0,
//       c = -(3 (p0 - p1) + T (2 v0 + v1)) / T^2,   d = (2 (p0 - p1) + T (v0 + v1)) / T^3
//   * value (polynomial.cc:50-76):  sum over the coefficients in the order a, b, c, d of  dphi_k(t) * coeff_k  with
//       pos: t^k     vel: k t^(k-1) (0 for k = 0)     acc: k (k-1) t^(k-2) (0 for k < 2)
// Table layout (SoA, instance index fastest, leading dimension ld):
//   durations [4 * nseg][ld]            row s * nseg + j
//   nodes     [4 * (nseg + 1) * 6][ld]  row ((s * (nseg + 1) + k) * 6 + c), c = 0..2 position, 3..5 velocity of node k
// with splines s = 0 base_linear, 1 base_angular, 2 first swing foot, 3 second swing foot (Jsw row order).
// Output: 36 rows [ld]: com_des_pos 6 (linear p, angular p) | com_des_vel 6 | com_des_acc 6 | sw_des_pos 6 | sw_des_vel 6 | sw_des_acc 6.
#pragma once
#include "qp_warp.cuh"      // WBC_HD

namespace wbc {

constexpr int TRAJ_SPLINES = 4;
constexpr int TRAJ_MAX_SEG = 8;
constexpr int TRAJ_OUT_ROWS = 36;
WBC_HD int traj_duration_rows(int nseg) { return TRAJ_SPLINES * nseg; }
WBC_HD int traj_node_rows(int nseg) { return TRAJ_SPLINES * (nseg + 1) * 6; }

// one instance: tables at dur + i, nodes + i (stride ld)
// outp[b]: the six output blocks (com pos, vel, acc, swing pos, vel, acc), each [6][out_ld], already offset to the instance
struct TrajOut { double* p[6]; long ld; };
WBC_HD void sample_trajectory_instance(int nseg, const double* dur, const double* nodes, long ld, double t_global, const TrajOut& o)
{
    if (!(t_global >= 0.0)) t_global = 0.0;
    for (int s = 0; s < TRAJ_SPLINES; s++) {
        // GetSegmentID (spline.cc:48-66) and GetLocalTime (68-79)
        const double eps = 1e-10;
        int id = nseg - 1;
        double acc = 0.0;
        for (int j = 0; j < nseg; j++) {
            acc += dur[(long)(s * nseg + j) * ld];
            if (acc >= t_global - eps) { id = j; break; }
        }
        double tl = t_global;
        for (int j = 0; j < id; j++) tl -= dur[(long)(s * nseg + j) * ld];
        const double T = dur[(long)(s * nseg + id) * ld];
        const double* n0 = nodes + (long)((s * (nseg + 1) + id) * 6) * ld;
        const double* n1 = n0 + 6 * ld;
        const double t2 = tl * tl, t3 = t2 * tl, T2 = T * T, T3 = T2 * T;
        // output rows: s = 0 -> rows 0..2 of each com block, s = 1 -> rows 3..5; s = 2, 3 likewise in the swing blocks
        const int blk = (s < 2) ? 0 : 3;
        const int r0 = (s & 1) * 3;
        for (int c = 0; c < 3; c++) {
            const double p0 = n0[(long)c * ld], v0 = n0[(long)(3 + c) * ld], p1 = n1[(long)c * ld], v1 = n1[(long)(3 + c) * ld];
            const double ca = p0, cb = v0;
            const double cc = -(3.0 * (p0 - p1) + T * (2.0 * v0 + v1)) / T2;
            const double cd = (2.0 * (p0 - p1) + T * (v0 + v1)) / T3;
            double p = 0.0, v = 0.0, a = 0.0;
            p += 1.0 * ca; p += tl * cb; p += t2 * cc; p += t3 * cd;
            v += 0.0 * ca; v += (1.0 * 1.0) * cb; v += (2.0 * tl) * cc; v += (3.0 * t2) * cd;
            a += 0.0 * ca; a += 0.0 * cb; a += (2.0 * 1.0) * cc; a += (6.0 * tl) * cd;
            o.p[blk][(long)(r0 + c) * o.ld] = p;
            o.p[blk + 1][(long)(r0 + c) * o.ld] = v;
            o.p[blk + 2][(long)(r0 + c) * o.ld] = a;
        }
    }
}

}  // namespace wbc
