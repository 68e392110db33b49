// Front half of the control cycle, FOUR LANES PER INSTANCE: lane g of an instance's group owns leg g (BL, BR, FL, FR).
//
// Same quantities and the same closed-form change to CoM coordinates as front_cycle (wbc_front.cuh, which stays the
// thread-per-instance form behind wbc_debug_update, the forward-dynamics plant and the host emulation); what changes is who
// computes what:
//   * per leg, in its lane: FK, velocities and bias accelerations of the three links, their Newton-Euler wrenches, the leg's share of
//     the whole-robot sums (mass moment, rotational inertia about the base origin, base rows of h and g), its three CRBA columns
//     (Mbj, the 3x3 diagonal block of Mjj), its bias torques, its foot kinematics and foot Jacobian;
//   * the whole-robot sums run link by link in the thread-per-instance kernel's order, every lane fetching each link's operands
//     from its owner (shuffles inside the group of four): what is left between the two kernels is the compiler's choice of fused
//     multiply-adds (agreement 1e-11 on every field of the QP record; butterfly sums were measured 14 us faster at 4096 instances
//     and moved one robot-cycle of the 47 104 of the rollout parity test from 2e-10 to 1.2e-6 -- the solver amplifies last bits);
//   * the 6x6 base block is factored by every lane (a few dozen flops; splitting it would cost more shuffles than it saves);
//     each lane solves for ITS three columns of P = Mb^-1 Mbj; P dq_j is accumulated in DoF order over the lanes;
//   * the dense joint block Mjj - Mbj' P and the joint columns of the foot Jacobians need every column of P: the lanes pass their
//     6x3 blocks round (shuffles inside the group of four), and each lane forms and stores its three rows, one 3x3 block at a time;
//   * J' Fgrf runs over the twelve foot rows in order (every lane gets every foot's lever arm and force); the observer, Wcom_des and the CoM block are then formed by every lane and stored by lane 0.
// A thread handles a quarter of the instance's arithmetic and its longest dependent chain is a leg, not the robot: the kernel
// is 4 x as many warps of a quarter of the length each, which is what a batch of a few thousand (or one) robot needs; nothing
// lives in local memory.
#pragma once
#if defined(__CUDACC__)
#include "wbc_front.cuh"

namespace wbc {

constexpr unsigned FL_FULL = 0xffffffffu;
// the value lane `leg` of this instance's group holds
__device__ __forceinline__ double leg_get(double v, int leg) { return __shfl_sync(FL_FULL, v, leg, 4); }
__device__ __forceinline__ V3 leg_get(const V3& a, int leg) { return v3(leg_get(a.x, leg), leg_get(a.y, leg), leg_get(a.z, leg)); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// One lane of one instance.  `i` indexes the SoA arrays (clamped by the caller for lanes past the batch, which compute along and
// store nothing: `valid`); `rec` is the instance's QP record.
__device__ __forceinline__ void front_cycle_leg(const Params& P, const DevInputs& in, const FrontState& st, long i, bool valid, int leg,
                                                double* __restrict__ rec, double* __restrict__ w_out, long w_ld)
{
    using namespace dogbot;
    const long ld = in.ld;
#define LD1(ptr, k) (ptr)[(long)(k) * ld + i]
    const V3 p0 = v3(LD1(in.base_pos, 0), LD1(in.base_pos, 1), LD1(in.base_pos, 2));
    M3 R0;
#pragma unroll
    for (int k = 0; k < 9; k++) R0.m[k] = LD1(in.base_rot, k);
    const V3 v0 = v3(LD1(in.base_vel, 0), LD1(in.base_vel, 1), LD1(in.base_vel, 2));
    const V3 w0 = v3(LD1(in.base_vel, 3), LD1(in.base_vel, 4), LD1(in.base_vel, 5));
    const V3 grav = v3(P.gravity[0], P.gravity[1], P.gravity[2]);
    const int mode = in.mode[i];
    const int sf = kFootLeg[leg];                      // stacked foot index of this leg (the table is its own inverse)
    const int d0 = dof_index(leg, 0), d1 = dof_index(leg, 1);          // d2 = d1 + 1
    const double qd3[3] = {LD1(in.dq, d0), LD1(in.dq, d1), LD1(in.dq, d1 + 1)};
    const double q3[3] = {LD1(in.q, d0), LD1(in.q, d1), LD1(in.q, d1 + 1)};

    // ---- this leg: kinematics and Newton-Euler with nu_dot = 0 (front_dynamics, one leg)
    V3 pj[3], z[3], c[3], f[3], n[3];
    S3 Iw[3];
    // what the whole-robot sums take from each link (summed below in the thread-per-instance kernel's order: link by link)
    V3 xdv[3], xcr[3];
    double xI[3][6];
    M3 Rp = R0;
    V3 pp = v3(0, 0, 0), vp = v0, ap = v3(0, 0, 0), wp = w0, alp = v3(0, 0, 0);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const V3 org = v3(kJointOrigin[leg][k][0], kJointOrigin[leg][k][1], kJointOrigin[leg][k][2]);
        const V3 ax = v3(kJointAxis[leg][k][0], kJointAxis[leg][k][1], kJointAxis[leg][k][2]);
        const V3 off = mul(Rp, org);
        pj[k] = pp + off;
        const V3 wxo = cross(wp, off);
        const V3 vj = vp + wxo;
        const V3 aj = ap + cross(alp, off) + cross(wp, wxo);
        z[k] = mul(Rp, ax);
        const double qd = qd3[k];
        const V3 wk = wp + qd * z[k];
        const V3 alk = alp + qd * cross(wp, z[k]);
        const M3 Rk = mul(Rp, axis_rotation(ax, q3[k]));
        const V3 rho = mul(Rk, v3(kLinkCom[leg][k][0], kLinkCom[leg][k][1], kLinkCom[leg][k][2]));
        c[k] = pj[k] + rho;
        const V3 wxr = cross(wk, rho);
        const V3 vc = vj + wxr;
        const V3 ac = aj + cross(alk, rho) + cross(wk, wxr);
        Iw[k] = rotate_inertia(Rk, kLinkInertia[leg][k]);
        const double m = kLinkMass[leg][k];
        xdv[k] = vc - v0;
        const double r2 = dot(c[k], c[k]);
        xI[k][0] = Iw[k].xx + m * (r2 - c[k].x * c[k].x);
        xI[k][1] = Iw[k].yy + m * (r2 - c[k].y * c[k].y);
        xI[k][2] = Iw[k].zz + m * (r2 - c[k].z * c[k].z);
        xI[k][3] = Iw[k].xy - m * c[k].x * c[k].y;
        xI[k][4] = Iw[k].xz - m * c[k].x * c[k].z;
        xI[k][5] = Iw[k].yz - m * c[k].y * c[k].z;
        f[k] = m * (ac - grav);
        n[k] = mul(Iw[k], alk) + cross(wk, mul(Iw[k], wk));
        xcr[k] = cross(c[k], f[k]);
        Rp = Rk; pp = pj[k]; vp = vj; ap = aj; wp = wk; alp = alk;
    }
    // foot frame (fixed to the lower leg) and the foot's Jacobian with respect to the leg's joints
    V3 footp, footv, foota;
    double Jleg[9];
    {
        const V3 off = mul(Rp, v3(kFootOffset[leg][0], kFootOffset[leg][1], kFootOffset[leg][2]));
        const V3 wxo = cross(wp, off);
        footp = pp + off;
        footv = vp + wxo;
        foota = ap + cross(alp, off) + cross(wp, wxo);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const V3 col = cross(z[k], footp - pj[k]);
            Jleg[0 * 3 + k] = col.x; Jleg[1 * 3 + k] = col.y; Jleg[2 * 3 + k] = col.z;
        }
    }
    const M3 footR = Rp;
    // joint-space projections: bias torques, the leg's three CRBA columns, its 3x3 block of Mjj
    double hj3[3], Mbj3[6][3], Mjj3[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        V3 acc = v3(0, 0, 0);
        V3 Lj = v3(0, 0, 0), Bj = v3(0, 0, 0);
#pragma unroll
        for (int b = j; b < 3; b++) {
            const V3 d = c[b] - pj[j];
            acc = acc + n[b] + cross(d, f[b]);
            const V3 l = kLinkMass[leg][b] * cross(z[j], d);
            Lj = Lj + l;
            Bj = Bj + mul(Iw[b], z[j]) + cross(c[b], l);
        }
        hj3[j] = dot(z[j], acc);
        Mbj3[0][j] = Lj.x; Mbj3[1][j] = Lj.y; Mbj3[2][j] = Lj.z;
        Mbj3[3][j] = Bj.x; Mbj3[4][j] = Bj.y; Mbj3[5][j] = Bj.z;
#pragma unroll
        for (int k = 0; k <= j; k++) {
            const double v = dot(z[k], Bj - cross(pj[k], Lj));
            Mjj3[k][j] = v;
            Mjj3[j][k] = v;
        }
    }

    // ---- whole-robot sums: the base body (body + bodytext lumped, CoM at the base origin), then link by link, leg after leg --
    // the same chain of additions as front_dynamics: every lane fetches every link's operands from the lane that owns it
    V3 mr = v3(0, 0, 0), mvrel = v3(0, 0, 0), ftot, ntot;
    double Ibxx, Ibyy, Ibzz, Ibxy, Ibxz, Ibyz;
    {
        const double Ib6[6] = {kBaseInertia[0], kBaseInertia[1], kBaseInertia[2], 0.0, 0.0, 0.0};
        const S3 Ibw = rotate_inertia(R0, Ib6);
        Ibxx = Ibw.xx; Ibyy = Ibw.yy; Ibzz = Ibw.zz; Ibxy = Ibw.xy; Ibxz = Ibw.xz; Ibyz = Ibw.yz;
        ftot = (-kBaseMass) * grav;
        ntot = cross(w0, mul(Ibw, w0));
    }
#pragma unroll
    for (int src = 0; src < 4; src++) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double m = kLinkMass[src][k];
            const V3 cc = leg_get(c[k], src), dd = leg_get(xdv[k], src), ff = leg_get(f[k], src), nn = leg_get(n[k], src),
                     cr = leg_get(xcr[k], src);
            mr = mr + m * cc;
            mvrel = mvrel + m * dd;
            Ibxx += leg_get(xI[k][0], src); Ibyy += leg_get(xI[k][1], src); Ibzz += leg_get(xI[k][2], src);
            Ibxy += leg_get(xI[k][3], src); Ibxz += leg_get(xI[k][4], src); Ibyz += leg_get(xI[k][5], src);
            ftot = ftot + ff;
            ntot = ntot + nn + cr;
        }
    }
    const double mtot = kTotalMass;
    double Mb[36];
#pragma unroll
    for (int k = 0; k < 36; k++) Mb[k] = 0.0;
    Mb[0] = Mb[7] = Mb[14] = mtot;
    Mb[0 * 6 + 4] = mr.z;  Mb[0 * 6 + 5] = -mr.y;
    Mb[1 * 6 + 3] = -mr.z; Mb[1 * 6 + 5] = mr.x;
    Mb[2 * 6 + 3] = mr.y;  Mb[2 * 6 + 4] = -mr.x;
    Mb[3 * 6 + 1] = -mr.z; Mb[3 * 6 + 2] = mr.y;
    Mb[4 * 6 + 0] = mr.z;  Mb[4 * 6 + 2] = -mr.x;
    Mb[5 * 6 + 0] = -mr.y; Mb[5 * 6 + 1] = mr.x;
    Mb[3 * 6 + 3] = Ibxx; Mb[4 * 6 + 4] = Ibyy; Mb[5 * 6 + 5] = Ibzz;
    Mb[3 * 6 + 4] = Mb[4 * 6 + 3] = Ibxy;
    Mb[3 * 6 + 5] = Mb[5 * 6 + 3] = Ibxz;
    Mb[4 * 6 + 5] = Mb[5 * 6 + 4] = Ibyz;
    const double hb[6] = {ftot.x, ftot.y, ftot.z, ntot.x, ntot.y, ntot.z};

    const V3 xbc = (1.0 / mtot) * mr;                 // com - base            main.cpp:518
    const V3 xbcd = (1.0 / mtot) * mvrel;             // com_vel - v_base      main.cpp:538
    const V3 com = p0 + xbc;
    const V3 comv = v0 + xbcd;

    // ---- computeTransformation in closed form (main.cpp:491-568): one 6x6 Cholesky, this lane's three columns of P = Mb^-1 Mbj
    double Lc[36];
#pragma unroll
    for (int k = 0; k < 36; k++) Lc[k] = Mb[k];
    chol6(Lc);
    double P3[6][3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double col[6];
#pragma unroll
        for (int k = 0; k < 6; k++) col[k] = Mbj3[k][j];
        chol6_solve(Lc, col);
#pragma unroll
        for (int k = 0; k < 6; k++) P3[k][j] = col[k];
    }
    // P dq_j, accumulated in DoF order (the four rolls, then pitch and knee leg by leg) like front_cycle's column loop
    double pv[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int src = 0; src < 4; src++) {
        const double qd = leg_get(qd3[0], src);
#pragma unroll
        for (int k = 0; k < 6; k++) pv[k] += leg_get(P3[k][0], src) * qd;
    }
#pragma unroll
    for (int src = 0; src < 4; src++)
#pragma unroll
        for (int j = 1; j < 3; j++) {
            const double qd = leg_get(qd3[j], src);
#pragma unroll
            for (int k = 0; k < 6; k++) pv[k] += leg_get(P3[k][j], src) * qd;
        }
    V3 u3;                                            // first three entries of T_inv_dot dq (main.cpp:565-566, 648, 658)
    {
        const V3 mdr = mtot * xbcd;
        const V3 pvl = v3(pv[0], pv[1], pv[2]), pva = v3(pv[3], pv[4], pv[5]);
        const V3 y0 = (-1.0) * cross(mdr, pva);
        const V3 y1 = cross(mdr, pvl);
        double zz[6] = {y0.x, y0.y, y0.z, y1.x, y1.y, y1.z};
        chol6_solve(Lc, zz);
        const V3 zl = v3(zz[0], zz[1], zz[2]), za = v3(zz[3], zz[4], zz[5]);
        const V3 dJs = (-1.0) * cross(xbcd, pva) - (zl - cross(xbc, za));
        u3 = cross(xbcd, w0) - dJs;
    }
    // BiasCOM = T^-T (h + M T_inv_dot dq): base rows (every lane), this leg's joint rows
    double hb2[6], hc6[6], hcj[3];
#pragma unroll
    for (int k = 0; k < 6; k++) hb2[k] = hb[k] + Mb[k * 6] * u3.x + Mb[k * 6 + 1] * u3.y + Mb[k * 6 + 2] * u3.z;
    {
        const V3 t = cross(xbc, v3(hb2[0], hb2[1], hb2[2]));
        hc6[0] = hb2[0]; hc6[1] = hb2[1]; hc6[2] = hb2[2];
        hc6[3] = hb2[3] - t.x; hc6[4] = hb2[4] - t.y; hc6[5] = hb2[5] - t.z;
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double s = hj3[j] + Mbj3[0][j] * u3.x + Mbj3[1][j] * u3.y + Mbj3[2][j] * u3.z;
#pragma unroll
        for (int k = 0; k < 6; k++) s -= P3[k][j] * hb2[k];
        hcj[j] = s;
    }
    // this foot's rows of JacCOM_lin (base columns) and of JdqdCOM_lin (main.cpp:654-659)
    const V3 rf = footp;                         // foot - base origin
    const V3 rc = rf - xbc;                      // foot - com
    const double nS[9] = {0, rc.z, -rc.y, -rc.z, 0, rc.x, rc.y, -rc.x, 0};
    const double nSf[9] = {0, rf.z, -rf.y, -rf.z, 0, rf.x, rf.y, -rf.x, 0};
    double Jc3[3][6], Jd3[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 3; b++) { Jc3[a][b] = (a == b) ? 1.0 : 0.0; Jc3[a][3 + b] = nS[a * 3 + b]; }
        Jd3[a] = comp(foota, a) + comp(u3, a);
    }

    // ---- the blocks that need every column of P: Mjj_com = Mjj - Mbj' P (this leg's rows), joint columns of this foot's Jacobian rows.
    // One source leg at a time: its 6x3 block of P comes over by shuffles; the two 3x3 blocks are stored as soon as they exist.
    double* recMjj = rec + QR_MJJ;
    double* recJj = rec + QR_JJ + (3 * sf) * 12;
#pragma unroll 1
    for (int src = 0; src < 4; src++) {
        double Ps[6][3];
#pragma unroll
        for (int k = 0; k < 6; k++)
#pragma unroll
            for (int b = 0; b < 3; b++) Ps[k][b] = leg_get(P3[k][b], src);
        const bool own = src == leg;
        const int c0 = dof_index(src, 0), c1 = dof_index(src, 1);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double mv[3], jv[3];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                double s = own ? Mjj3[a][b] : 0.0;
#pragma unroll
                for (int k = 0; k < 6; k++) s -= Mbj3[k][a] * Ps[k][b];
                mv[b] = s;
                const double t = Ps[a][b] + nSf[a * 3 + 0] * Ps[3][b] + nSf[a * 3 + 1] * Ps[4][b] + nSf[a * 3 + 2] * Ps[5][b];
                jv[b] = own ? -t + Jleg[a * 3 + b] : -t;
            }
            if (valid) {
                const int ra = (a == 0) ? d0 : d1 + (a - 1);
                recMjj[ra * 12 + c0] = mv[0];
                st2(recMjj + ra * 12 + c1, mv[1], mv[2]);
                recJj[a * 12 + c0] = jv[0];
                st2(recJj + a * 12 + c1, jv[1], jv[2]);
            }
        }
    }

    // ---- Fgrf of this foot (main.cpp:1022-1026; swing feet zeroed 1218 / 1765) and J' Fgrf
    const bool swing = (mode == MODE_SWING_BR_FL && (sf == 0 || sf == 2)) || (mode == MODE_SWING_BL_FR && (sf == 1 || sf == 3));
    V3 fw = v3(0, 0, 0);
    if (!swing) fw = mul(footR, v3(LD1(in.foot_force, 3 * sf), LD1(in.foot_force, 3 * sf + 1), LD1(in.foot_force, 3 * sf + 2)));
    // every foot's lever arm and force, by stacked foot, so that J' Fgrf (and J'J below) run over the twelve rows in order in every lane
    V3 rcA[4], fwA[4];
#pragma unroll
    for (int sfi = 0; sfi < 4; sfi++) {
        const int lg = (sfi < 2) ? 1 - sfi : sfi;      // kFootLeg
        rcA[sfi] = leg_get(rc, lg);
        fwA[sfi] = leg_get(fw, lg);
    }
    double fc6[6];
#pragma unroll
    for (int a = 0; a < 6; a++) {
        double fc = 0.0;
#pragma unroll
        for (int sfi = 0; sfi < 4; sfi++) {
            const V3 r = rcA[sfi];
            const double nSa[9] = {0, r.z, -r.y, -r.z, 0, r.x, r.y, -r.x, 0};
#pragma unroll
            for (int ax = 0; ax < 3; ax++) {
                const double j = (a < 3) ? ((ax == a) ? 1.0 : 0.0) : nSa[ax * 3 + (a - 3)];
                fc += j * comp(fwA[sfi], ax);
            }
        }
        fc6[a] = fc;
    }

    // Mc = Xi' Mb Xi  (MassMatrixCOM[0:6,0:6], main.cpp:645) -- formed here, after the blocks above have released their registers
    double Mc[36];
    {
        double MX[36];
        const double S[9] = {0, -xbc.z, xbc.y, xbc.z, 0, -xbc.x, -xbc.y, xbc.x, 0};
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
                MX[r * 6 + cc] = Mb[r * 6 + cc];
                MX[r * 6 + 3 + cc] = Mb[r * 6 + 3 + cc] + Mb[r * 6] * S[cc] + Mb[r * 6 + 1] * S[3 + cc] + Mb[r * 6 + 2] * S[6 + cc];
            }
#pragma unroll
        for (int cc = 0; cc < 6; cc++) {
            Mc[0 * 6 + cc] = MX[0 * 6 + cc]; Mc[1 * 6 + cc] = MX[1 * 6 + cc]; Mc[2 * 6 + cc] = MX[2 * 6 + cc];
#pragma unroll
            for (int r = 0; r < 3; r++)
                Mc[(3 + r) * 6 + cc] = MX[(3 + r) * 6 + cc] + S[0 * 3 + r] * MX[0 * 6 + cc] + S[1 * 3 + r] * MX[1 * 6 + cc] + S[2 * 3 + r] * MX[2 * 6 + cc];
        }
    }
    // ---- estimate() (main.cpp:692-725), every lane the same arithmetic; lane 0 stores
    const double comv6[6] = {comv.x, comv.y, comv.z, w0.x, w0.y, w0.z};
    double west[6], rho6[6], dd6[6];
    bool finite_obs = true;
#pragma unroll
    for (int a = 0; a < 6; a++) {
        double rho = 0.0;
#pragma unroll
        for (int b = 0; b < 6; b++) rho += Mc[a * 6 + b] * comv6[b];
        rho6[a] = rho;
        dd6[a] = -mtot * (a == 2 ? P.g_acc : 0.0) + fc6[a];
    }
    if (P.observer_enabled) {
        const double T = P.obs_dt, k0 = in.obs_gain ? in.obs_gain[i] : P.obs_gain;
        const double mgain = (1.0 / (1.0 + k0 * T)) * k0;
        const bool second = P.obs_order == 2, expl = P.obs_form == 1;
        const double k2 = P.obs_gain2;
        double ydn[6], ywn[6], ygn[6];
#pragma unroll
        for (int a = 0; a < 6; a++) {
            const double ywp = st.yw[(long)a * st.ld + i];
            const double yd = st.yd[(long)a * st.ld + i] + dd6[a] * T;
            double wv;
            if (!second) {
                if (!expl) wv = mgain * (rho6[a] - ywp - yd);
                else wv = k0 * (rho6[a] - ywp - yd);
                ygn[a] = 0.0;
            } else {
                const double e = rho6[a] - ywp - yd;
                const double ygp = st.yg[(long)a * st.ld + i];
                if (!expl) {
                    wv = k2 * (ygp + T * k0 * e) / (1.0 + k2 * T + k0 * k2 * T * T);
                    ygn[a] = ygp + T * (k0 * (e - wv * T) - wv);
                } else {
                    ygn[a] = ygp + T * (k0 * e - k2 * ygp);
                    wv = k2 * ygn[a];
                }
            }
            ydn[a] = yd;
            ywn[a] = ywp + wv * T;
            west[a] = wv;
            finite_obs = finite_obs && (yd - yd == 0.0) && (ywn[a] - ywn[a] == 0.0) && (ygn[a] - ygn[a] == 0.0);
        }
        __syncwarp();                                  // every lane of the group has read the carried state
        if (finite_obs && valid && leg == 0) {
#pragma unroll
            for (int a = 0; a < 6; a++) { st.yd[(long)a * st.ld + i] = ydn[a]; st.yw[(long)a * st.ld + i] = ywn[a]; }
            if (second)
#pragma unroll
                for (int a = 0; a < 6; a++) st.yg[(long)a * st.ld + i] = ygn[a];
        }
    } else {
#pragma unroll
        for (int a = 0; a < 6; a++) west[a] = 0.0;
    }
    if (valid && leg == 0)
#pragma unroll
        for (int a = 0; a < 6; a++) w_out[(long)a * w_ld + i] = finite_obs ? west[a] : 0.0;
    if (st.w3) {
        // getw3 (estimator_sem.cpp:64-70): w3 = J (J'J)^-1 w; J'J is the sum of the four feet's 3x6 blocks, each lane maps its rows
        double G[36], y[6];
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = 0; b <= a; b++) {
                double g = 0.0;
#pragma unroll
                for (int sfi = 0; sfi < 4; sfi++) {
                    const V3 r = rcA[sfi];
                    const double nSa[9] = {0, r.z, -r.y, -r.z, 0, r.x, r.y, -r.x, 0};
#pragma unroll
                    for (int ax = 0; ax < 3; ax++) {
                        const double ja = (a < 3) ? ((ax == a) ? 1.0 : 0.0) : nSa[ax * 3 + (a - 3)];
                        const double jb = (b < 3) ? ((ax == b) ? 1.0 : 0.0) : nSa[ax * 3 + (b - 3)];
                        g += ja * jb;
                    }
                }
                G[a * 6 + b] = g; G[b * 6 + a] = g;
            }
        chol6(G);
#pragma unroll
        for (int a = 0; a < 6; a++) y[a] = finite_obs ? west[a] : 0.0;
        chol6_solve(G, y);
        if (valid)
#pragma unroll
            for (int a = 0; a < 3; a++) {
                double v = 0.0;
#pragma unroll
                for (int b = 0; b < 6; b++) v += Jc3[a][b] * y[b];
                st.w3[(long)(3 * sf + a) * st.w3_ld + i] = v;
            }
    }

    // ---- Wcom_des (main.cpp:1012-1032)
    double Wc[6];
    {
        double dx[6], dv[6], ades[6];
        dx[0] = LD1(in.com_des_pos, 0) - com.x; dx[1] = LD1(in.com_des_pos, 1) - com.y; dx[2] = LD1(in.com_des_pos, 2) - com.z;
        const V3 dr = mul(R0, v3(LD1(in.com_des_pos, 3) - LD1(in.base_rpy, 0), LD1(in.com_des_pos, 4) - LD1(in.base_rpy, 1),
                                 LD1(in.com_des_pos, 5) - LD1(in.base_rpy, 2)));
        dx[3] = dr.x; dx[4] = dr.y; dx[5] = dr.z;
#pragma unroll
        for (int a = 0; a < 6; a++) { dv[a] = LD1(in.com_des_vel, a) - comv6[a]; ades[a] = LD1(in.com_des_acc, a); }
#pragma unroll
        for (int a = 0; a < 6; a++) {
            double ma = 0.0;
#pragma unroll
            for (int b = 0; b < 6; b++) ma += Mc[a * 6 + b] * ades[b];
            Wc[a] = P.kcom * dx[a] + P.dcom * dv[a] + mtot * (a == 2 ? P.g_acc : 0.0) + ma - west[a];
        }
    }
    if (!valid) return;

    // ---- the rest of the QP record: this leg's rows, and the shared rows spread over the four lanes
    rec[QR_HJ + d0] = hcj[0]; rec[QR_HJ + d1] = hcj[1]; rec[QR_HJ + d1 + 1] = hcj[2];
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 6; b += 2) st2(rec + QR_JC + (3 * sf + a) * 6 + b, Jc3[a][b], Jc3[a][b + 1]);
        rec[QR_JDQD + 3 * sf + a] = Jd3[a];
    }
    {
        const double dt = P.joint_dt, kk = 2.0 / (dt * dt);                  // main.cpp:1098-1104
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int d = (j == 0) ? d0 : d1 + (j - 1);
            rec[QR_DDQMAX + d] = kk * (kQmax[d] - q3[j] - dt * qd3[j]);
            rec[QR_DDQMIN + d] = kk * (kQmin[d] - q3[j] - dt * qd3[j]);
        }
    }
    {   // swing-foot PD (main.cpp:1327-1379): this foot's slot, if it is one of the two swing feet in Jsw row order
        const int sf0 = (mode == MODE_SWING_BL_FR) ? 1 : 0, sf1 = (mode == MODE_SWING_BL_FR) ? 3 : 2;
        if (sf == sf0 || sf == sf1) {
            const int slot = (sf == sf0) ? 0 : 3;
#pragma unroll
            for (int ax = 0; ax < 3; ax++) {
                const int a = slot + ax;
                const double pos = comp(p0, ax) + comp(footp, ax), vel = comp(footv, ax);
                const double vdot = LD1(in.sw_des_acc, a) + P.kd_sw * (LD1(in.sw_des_vel, a) - vel) + P.kp_sw * (LD1(in.sw_des_pos, a) - pos);
                rec[QR_SWRHS + a] = (mode == MODE_STANCE) ? 0.0 : vdot - Jd3[ax];
            }
        }
    }
    {   // friction pyramid rows of this foot (main.cpp:1062-1078)
        double nn[3] = {0, 0, 1}, t1[3] = {1, 0, 0}, t2[3] = {0, 1, 0}, mu = P.mu;
        if (in.terrain) {
#pragma unroll
            for (int k = 0; k < 3; k++) { nn[k] = LD1(in.terrain, 10 * sf + k); t1[k] = LD1(in.terrain, 10 * sf + 3 + k); t2[k] = LD1(in.terrain, 10 * sf + 6 + k); }
            mu = LD1(in.terrain, 10 * sf + 9);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            rec[QR_CFR + 15 * sf + 0 + k] = -mu * nn[k] + t1[k];
            rec[QR_CFR + 15 * sf + 3 + k] = -mu * nn[k] + t2[k];
            rec[QR_CFR + 15 * sf + 6 + k] = -(mu * nn[k] + t1[k]);
            rec[QR_CFR + 15 * sf + 9 + k] = -(mu * nn[k] + t2[k]);
            rec[QR_CFR + 15 * sf + 12 + k] = -nn[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) rec[QR_FOOTR + 9 * sf + k] = footR.m[k];
    // shared rows: Mc (nine entries per lane), and one small group per lane
#pragma unroll
    for (int g = 0; g < 4; g++)
        if (leg == g) {
#pragma unroll
            for (int k = 0; k < 9; k++) rec[QR_MC + 9 * g + k] = Mc[9 * g + k];
        }
    if (leg == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) rec[QR_HC + k] = hc6[k];
    } else if (leg == 1) {
#pragma unroll
        for (int k = 0; k < 6; k++) rec[QR_WCOM + k] = Wc[k];
    } else if (leg == 2) {
#pragma unroll
        for (int a = 0; a < 6; a++) { rec[QR_RHO + a] = rho6[a]; rec[QR_DD + a] = dd6[a]; }
    } else {
        rec[QR_MODE] = (double)mode;
        rec[QR_XBC] = xbc.x; rec[QR_XBC + 1] = xbc.y; rec[QR_XBC + 2] = xbc.z;
    }
#undef LD1
}

}  // namespace wbc
#endif
