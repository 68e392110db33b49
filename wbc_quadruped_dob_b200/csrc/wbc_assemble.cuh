// Assembly of the ground-reaction-force QP (Q, c, L) from the compact QP record, and the maps back
// from the QP solution to joint torques and objective value.  Team-cooperative (executor `Ex`, see
// qp_warp.cuh).  Layouts follow the reference exactly:
//   stance  main.cpp:984-1120   x = [ddq_com(6) | ddq_j(12) | f(12)],          L 86 x 31, 18 equalities
//   swing   main.cpp:1163-1389  x = [ddq_com(6) | ddq_j(12) | f_st(6) | g(6)], L 82 x 31, 12 equalities
//   torque  main.cpp:1126, 1396 tau = Mjj ddq_j + h_j - Jst_j' f
#pragma once
#include "qp_warp.cuh"
#include "wbc_types.h"

namespace wbc {

struct QpShape {
    int nrows, neq;
    int nst;          // number of stance feet
    int strow[4];     // first foot-coordinate row of each stance foot in the stacked Jacobian
    int swrow[2];     // first row of each swing foot
};

WBC_HD QpShape qp_shape(int mode)
{
    QpShape s;
    if (mode == MODE_STANCE) {
        s.nrows = 86; s.neq = 18; s.nst = 4;
        s.strow[0] = 0; s.strow[1] = 3; s.strow[2] = 6; s.strow[3] = 9; s.swrow[0] = s.swrow[1] = 0;
    } else if (mode == MODE_SWING_BR_FL) {          // main.cpp:1163-1167
        s.nrows = 82; s.neq = 12; s.nst = 2;
        s.strow[0] = 3; s.strow[1] = 9; s.strow[2] = s.strow[3] = 0; s.swrow[0] = 0; s.swrow[1] = 6;
    } else {                                         // main.cpp:1710-1714
        s.nrows = 82; s.neq = 12; s.nst = 2;
        s.strow[0] = 0; s.strow[1] = 6; s.strow[2] = s.strow[3] = 0; s.swrow[0] = 3; s.swrow[1] = 9;
    }
    return s;
}

// Q: 30x30 row-major with leading dimension LDQ, c: 30, L: nrows x 31 row-major (dense, zero-filled here).
template <int LDQ, class Ex>
WBC_HDN inline void assemble_qp(const Ex& ex, const Params& P, const double* rec, const QpShape& sh, double* Q, double* c,
                                double* L)
{
    using wbcqp::NMAIN;
    const int lane = ex.lane();
    const double* Mc = rec + QR_MC;
    const double* hc = rec + QR_HC;
    const double* hj = rec + QR_HJ;
    const double* Mjj = rec + QR_MJJ;
    const double* Jc = rec + QR_JC;
    const double* Jj = rec + QR_JJ;
    const double* Jd = rec + QR_JDQD;
    const double* Wc = rec + QR_WCOM;
    const int nf = 3 * sh.nst;                 // force variables
#pragma unroll 1
    for (int k = lane; k < 30 * LDQ; k += Ex::NL) Q[k] = 0.0;
#pragma unroll 1
    for (int k = lane; k < 30; k += Ex::NL) c[k] = 0.0;
#pragma unroll 1
    for (int k = lane; k < sh.nrows * 31; k += Ex::NL) L[k] = 0.0;
    ex.sync();
    // Q = T_s' Q1 T_s + R   (main.cpp:994-1001, 1176-1189)
#pragma unroll 1
    for (int k = lane; k < nf * nf; k += Ex::NL) {
        const int a = k / nf, b = k % nf;
        const int ra = sh.strow[a / 3] + a % 3, rb = sh.strow[b / 3] + b % 3;
        double s = 0.0;
#pragma unroll 1
        for (int t = 0; t < 6; t++) s += Jc[ra * 6 + t] * P.q1_weight * Jc[rb * 6 + t];
        Q[(18 + a) * LDQ + 18 + b] = s + (a == b ? 1.0 : 0.0);
    }
#pragma unroll 1
    for (int k = lane; k < 18; k += Ex::NL) Q[k * LDQ + k] = 1.0;
    if (sh.nst == 2)
#pragma unroll 1
        for (int k = 24 + lane; k < 30; k += Ex::NL) Q[k * LDQ + k] = P.slack_weight;      // main.cpp:1187-1189
    // c = -T_s' Q1 Wcom_des   (main.cpp:1033, 1224)
#pragma unroll 1
    for (int a = lane; a < nf; a += Ex::NL) {
        const int ra = sh.strow[a / 3] + a % 3;
        double s = 0.0;
#pragma unroll 1
        for (int t = 0; t < 6; t++) s += Jc[ra * 6 + t] * P.q1_weight * Wc[t];
        c[18 + a] = -s;
    }
    // equality rows (main.cpp:1039-1048, 1231-1241)
    const bool rhs_on = (sh.nst == 4) || P.fix_swing_rhs;
#pragma unroll 1
    for (int k = lane; k < 6 * 31; k += Ex::NL) {
        const int a = k / 31, b = k % 31;
        double v = 0.0;
        if (b < 6) v = Mc[a * 6 + b];
        else if (b >= 18 && b < 18 + nf) { const int t = b - 18; v = -Jc[(sh.strow[t / 3] + t % 3) * 6 + a]; }
        else if (b == 30 && rhs_on) v = -hc[a];
        L[a * 31 + b] = v;
    }
#pragma unroll 1
    for (int k = lane; k < nf * 31; k += Ex::NL) {
        const int a = k / 31, b = k % 31;
        const int ra = sh.strow[a / 3] + a % 3;
        double v = 0.0;
        if (b < 6) v = Jc[ra * 6 + b];
        else if (b < 18) v = Jj[ra * 12 + (b - 6)];
        else if (b == 30 && rhs_on) v = -Jd[ra];
        L[(6 + a) * 31 + b] = v;
    }
    // inequality rows
    const int r0 = sh.neq;                       // first inequality row
    const int nfr = 5 * sh.nst;                  // friction rows (main.cpp:1062-1085, 1266-1298)
#pragma unroll 1
    for (int k = lane; k < nfr * 3; k += Ex::NL) {
        const int r = k / 3, cc = k % 3, f = r / 5, rr = r % 5;
        const int sf = sh.strow[f] / 3;
        L[(r0 + r) * 31 + 18 + 3 * f + cc] = rec[QR_CFR + 15 * sf + 3 * rr + cc];
    }
    const int rt = r0 + nfr;                     // torque limits (main.cpp:1054-1057, 1088-1095)
#pragma unroll 1
    for (int k = lane; k < 12 * 31; k += Ex::NL) {
        const int a = k / 31, b = k % 31;
        double v = 0.0, vn = 0.0;
        if (b >= 6 && b < 18) { v = Mjj[a * 12 + (b - 6)]; vn = -v; }
        else if (b >= 18 && b < 18 + nf) { const int t = b - 18; vn = Jj[(sh.strow[t / 3] + t % 3) * 12 + a]; v = -vn; }
        else if (b == 30) { v = P.tau_max - hj[a]; vn = -(-P.tau_max - hj[a]); }
        L[(rt + a) * 31 + b] = v;
        L[(rt + 12 + a) * 31 + b] = vn;
    }
    int rq = rt + 24;
    if (sh.nst == 2) {                           // swing-foot tracking with slack (main.cpp:1251-1262, 1378-1379)
#pragma unroll 1
        for (int k = lane; k < 6 * 31; k += Ex::NL) {
            const int a = k / 31, b = k % 31;
            const int ra = sh.swrow[a / 3] + a % 3;
            double v = 0.0, vn = 0.0;
            if (b < 6) { v = Jc[ra * 6 + b]; vn = -v; }
            else if (b < 18) { v = Jj[ra * 12 + (b - 6)]; vn = -v; }
            else if (b == 24 + a) { v = -1.0; vn = -1.0; }
            else if (b == 30) { v = rec[QR_SWRHS + a]; vn = -v; }
            L[(rq + a) * 31 + b] = v;
            L[(rq + 6 + a) * 31 + b] = vn;
        }
        rq += 12;
    }
#pragma unroll 1
    for (int a = lane; a < 12; a += Ex::NL) {    // joint-acceleration limits (main.cpp:1058-1059, 1098-1107)
        L[(rq + a) * 31 + 6 + a] = 1.0;
        L[(rq + a) * 31 + 30] = rec[QR_DDQMAX + a];
        L[(rq + 12 + a) * 31 + 6 + a] = -1.0;
        L[(rq + 12 + a) * 31 + 30] = -rec[QR_DDQMIN + a];
    }
    ex.sync();
}

// tau (12) and 0.5 x'Qx + c'x from the record and the solution x (30, contiguous).
template <class Ex>
WBC_HDN inline void torque_and_objective(const Ex& ex, const Params& P, const double* rec, const QpShape& sh, const double* x,
                                         double* tau_out, long tau_ld, double* obj_out)
{
    const int lane = ex.lane();
    const double* hj = rec + QR_HJ;
    const double* Mjj = rec + QR_MJJ;
    const double* Jc = rec + QR_JC;
    const double* Jj = rec + QR_JJ;
    const double* Wc = rec + QR_WCOM;
    const int nf = 3 * sh.nst;
#pragma unroll 1
    for (int a = lane; a < 12; a += Ex::NL) {
        double s = 0.0;
#pragma unroll 1
        for (int b = 0; b < 12; b++) s += Mjj[a * 12 + b] * x[6 + b];
        s += hj[a];
        double jf = 0.0;
#pragma unroll 1
        for (int t = 0; t < nf; t++) jf += Jj[(sh.strow[t / 3] + t % 3) * 12 + a] * x[18 + t];
        tau_out[(long)a * tau_ld] = s - jf;
    }
    if (obj_out) {
        // x'Qx = sum_i R_ii x_i^2 + q1 |Jst_c' f|^2 ;  c'x = -q1 (Jst_c' f) . Wcom_des
        double sq = 0.0;
#pragma unroll 1
        for (int k = lane; k < 30; k += Ex::NL) sq += ((sh.nst == 2 && k >= 24) ? P.slack_weight : 1.0) * x[k] * x[k];
        sq = wbcqp::red_sum1(ex, sq);
        double jj = 0.0, jw = 0.0;
#pragma unroll 1
        for (int t = lane; t < 6; t += Ex::NL) {
            double s = 0.0;
#pragma unroll 1
            for (int a = 0; a < nf; a++) s += Jc[(sh.strow[a / 3] + a % 3) * 6 + t] * x[18 + a];
            jj += s * s;
            jw += s * Wc[t];
        }
        {
            double r2[2] = {jj, jw};
            wbcqp::red_sum<2>(ex, r2);
            jj = r2[0]; jw = r2[1];
        }
        if (lane == 0) *obj_out = 0.5 * (sq + P.q1_weight * jj) - P.q1_weight * jw;
    }
    ex.sync();
}

}  // namespace wbc
