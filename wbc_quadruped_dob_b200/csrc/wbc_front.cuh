// Front half of the batched WBC control cycle for ONE DogBot instance (thread-per-instance):
//   DOGCTRL::update()              main.cpp:572-660  floating-base dynamics + change to CoM coordinates
//   Fgrf                           main.cpp:1022-1026, 1218, 1765
//   DOGCTRL::estimate()            main.cpp:692-725   momentum-observer recurrence
//   Wcom_des, joint-limit rows, swing-foot PD, friction rows   main.cpp:1012-1032, 1062-1107, 1327-1379
// and emission of the compact QP record consumed by the solver kernel (wbc_types.h).
//
// Formulation (deliberately not the reference's, and not the oracle's Jacobian projection):
//   * 13 lumped bodies (base + 4 legs x {hip, upperleg, lowerleg+foot}); all vectors in world axes,
//     positions relative to the base origin, which is exactly iDynTree's MIXED representation
//     (main.cpp:292): nu = [v_base_origin (world), omega_base (world), dq].
//   * mass matrix by composite momenta (CRBA): column j is the momentum of the subtree below joint j
//     under a unit rate of joint j; bias/gravity forces by Newton-Euler with nu_dot = 0 (RNEA).
//   * the change of coordinates T (main.cpp:491-568) is applied in closed form:
//       T^-1 = [Xi, -P; 0, I],  Xi = [I, S(xbc); 0, I],  P = Mb^-1 Mbj  (one 6x6 Cholesky),
//     instead of six dense 18x18 inversions and three SVD solves; T_inv_dot only has its first three
//     rows (main.cpp:565-566), so T_inv_dot*dq collapses to a 3-vector u3.
#pragma once
#include <math.h>

#include "dogbot_model.h"
#include "wbc_types.h"

namespace wbc {

struct V3 {
    double x, y, z;
};
WBC_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
WBC_HD V3 operator+(const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
WBC_HD V3 operator-(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
WBC_HD V3 operator*(double s, const V3& a) { return v3(s * a.x, s * a.y, s * a.z); }
WBC_HD V3 cross(const V3& a, const V3& b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
WBC_HD double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
WBC_HD double comp(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct M3 {
    double m[9];   // row-major
};
WBC_HD V3 mul(const M3& A, const V3& v)
{
    return v3(A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
              A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z);
}
WBC_HD M3 mul(const M3& A, const M3& B)
{
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C.m[i * 3 + j] = A.m[i * 3] * B.m[j] + A.m[i * 3 + 1] * B.m[3 + j] + A.m[i * 3 + 2] * B.m[6 + j];
    return C;
}
// symmetric 3x3: xx, yy, zz, xy, xz, yz
struct S3 {
    double xx, yy, zz, xy, xz, yz;
};
WBC_HD V3 mul(const S3& I, const V3& v)
{
    return v3(I.xx * v.x + I.xy * v.y + I.xz * v.z, I.xy * v.x + I.yy * v.y + I.yz * v.z, I.xz * v.x + I.yz * v.y + I.zz * v.z);
}
// R * Iloc * R' for a symmetric body-frame inertia
WBC_HD S3 rotate_inertia(const M3& R, const double* I6)
{
    M3 L;
    L.m[0] = I6[0]; L.m[1] = I6[3]; L.m[2] = I6[4];
    L.m[3] = I6[3]; L.m[4] = I6[1]; L.m[5] = I6[5];
    L.m[6] = I6[4]; L.m[7] = I6[5]; L.m[8] = I6[2];
    M3 RL = mul(R, L);
    S3 W;
    W.xx = RL.m[0] * R.m[0] + RL.m[1] * R.m[1] + RL.m[2] * R.m[2];
    W.yy = RL.m[3] * R.m[3] + RL.m[4] * R.m[4] + RL.m[5] * R.m[5];
    W.zz = RL.m[6] * R.m[6] + RL.m[7] * R.m[7] + RL.m[8] * R.m[8];
    W.xy = RL.m[0] * R.m[3] + RL.m[1] * R.m[4] + RL.m[2] * R.m[5];
    W.xz = RL.m[0] * R.m[6] + RL.m[1] * R.m[7] + RL.m[2] * R.m[8];
    W.yz = RL.m[3] * R.m[6] + RL.m[4] * R.m[7] + RL.m[5] * R.m[8];
    return W;
}
// rotation by angle th about the unit axis a
WBC_HD M3 axis_rotation(const V3& a, double th)
{
    double s, c;
#if defined(__CUDA_ARCH__)
    sincos(th, &s, &c);
#else
    s = sin(th); c = cos(th);
#endif
    const double v = 1.0 - c;
    M3 R;
    R.m[0] = c + a.x * a.x * v;       R.m[1] = a.x * a.y * v - a.z * s; R.m[2] = a.x * a.z * v + a.y * s;
    R.m[3] = a.y * a.x * v + a.z * s; R.m[4] = c + a.y * a.y * v;       R.m[5] = a.y * a.z * v - a.x * s;
    R.m[6] = a.z * a.x * v - a.y * s; R.m[7] = a.z * a.y * v + a.x * s; R.m[8] = c + a.z * a.z * v;
    return R;
}

// In-place lower Cholesky of a 6x6 SPD matrix (row-major) and solves with it.
WBC_HD void chol6(double* A)
{
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= A[j * 6 + k] * A[j * 6 + k];
        d = sqrt(d);
        A[j * 6 + j] = d;
        const double r = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = A[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= A[i * 6 + k] * A[j * 6 + k];
            A[i * 6 + j] = s * r;
        }
    }
}
WBC_HD void chol6_solve(const double* L, double* x)
{
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[i * 6 + k] * x[k];
        x[i] = s / L[i * 6 + i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = x[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * x[k];
        x[i] = s / L[i * 6 + i];
    }
}

struct FrontState {        // observer state carried across cycles (main.cpp:721-724)
    double* yd;            // [6][ld]
    double* yw;            // [6][ld]
    long ld;
    double* yg;            // [6][ld] ygamma: integrator of the second-order recursion (Params::obs_order == 2), else unused
    double* w3;            // [12][w3_ld] or nullptr: the estimate mapped onto the feet (estimator_sem.cpp:64-70)
    long w3_ld;
};

// Floating-base dynamics of one instance in the MIXED representation (the first half of update(), main.cpp:591-630): mass-matrix
// blocks, bias and gravity forces, foot kinematics.  Shared by the control cycle (front_cycle) and the forward-dynamics plant.
struct FrontDyn {
    V3 p0, v0, w0;
    M3 R0;
    double Mb[36], Mbj[72], Mjj[144];    // mass-matrix blocks (base-base, base-joint, joint-joint)
    double hb[6], hj[12], gb[6], gj[12];
    V3 mr, mvrel;                        // sum m_i r_i (r relative to the base origin); sum m_i (v_ci - v0)
    V3 footp[4], footv[4], foota[4];     // canonical leg order; relative position, velocity, bias acceleration
    M3 footR[4];
    double Jleg[4][9];                   // d(foot velocity)/d(dq of its leg), [axis][k]
};

WBC_DEVFN inline void front_dynamics(const Params& P, const DevInputs& in, long i, FrontDyn& D)
{
    using namespace dogbot;
    const long ld = in.ld;
#define LD1(ptr, k) (ptr)[(long)(k) * ld + i]
    D.p0 = v3(LD1(in.base_pos, 0), LD1(in.base_pos, 1), LD1(in.base_pos, 2));
    M3& R0 = D.R0;
    for (int k = 0; k < 9; k++) R0.m[k] = LD1(in.base_rot, k);
    D.v0 = v3(LD1(in.base_vel, 0), LD1(in.base_vel, 1), LD1(in.base_vel, 2));
    D.w0 = v3(LD1(in.base_vel, 3), LD1(in.base_vel, 4), LD1(in.base_vel, 5));
    const V3 v0 = D.v0, w0 = D.w0;
    const V3 grav = v3(P.gravity[0], P.gravity[1], P.gravity[2]);

    double (&Mb)[36] = D.Mb; double (&Mbj)[72] = D.Mbj; double (&Mjj)[144] = D.Mjj;
    double (&hb)[6] = D.hb; double (&hj)[12] = D.hj; double (&gb)[6] = D.gb; double (&gj)[12] = D.gj;
    for (int k = 0; k < 144; k++) Mjj[k] = 0.0;

    // ---- base body (body + bodytext lumped, CoM at the base origin)
    V3 mr = v3(0, 0, 0);        // sum m_i r_i     (r relative to the base origin)
    V3 mvrel = v3(0, 0, 0);     // sum m_i (v_ci - v0)
    double Ibxx, Ibyy, Ibzz, Ibxy, Ibxz, Ibyz;   // rotational inertia of everything about the base origin
    V3 ftot, ntot, gftot, gntot;
    {
        const double Ib6[6] = {kBaseInertia[0], kBaseInertia[1], kBaseInertia[2], 0.0, 0.0, 0.0};
        const S3 Iw = rotate_inertia(R0, Ib6);
        Ibxx = Iw.xx; Ibyy = Iw.yy; Ibzz = Iw.zz; Ibxy = Iw.xy; Ibxz = Iw.xz; Ibyz = Iw.yz;
        ftot = (-kBaseMass) * grav;
        ntot = cross(w0, mul(Iw, w0));
        gftot = ftot;
        gntot = v3(0, 0, 0);
    }
    V3 (&footp)[4] = D.footp; V3 (&footv)[4] = D.footv; V3 (&foota)[4] = D.foota;
    M3 (&footR)[4] = D.footR;
    double (&Jleg)[4][9] = D.Jleg;

    for (int leg = 0; leg < 4; leg++) {
        V3 pj[3], z[3], c[3], f[3], n[3], fg[3];
        S3 Iw[3];
        M3 Rp = R0;
        V3 pp = v3(0, 0, 0), vp = v0, ap = v3(0, 0, 0), wp = w0, alp = v3(0, 0, 0);
        for (int k = 0; k < 3; k++) {
            const int d = dof_index(leg, k);
            const V3 org = v3(kJointOrigin[leg][k][0], kJointOrigin[leg][k][1], kJointOrigin[leg][k][2]);
            const V3 ax = v3(kJointAxis[leg][k][0], kJointAxis[leg][k][1], kJointAxis[leg][k][2]);
            const V3 off = mul(Rp, org);
            pj[k] = pp + off;
            const V3 wxo = cross(wp, off);
            const V3 vj = vp + wxo;
            const V3 aj = ap + cross(alp, off) + cross(wp, wxo);
            z[k] = mul(Rp, ax);
            const double qd = LD1(in.dq, d);
            const V3 wk = wp + qd * z[k];
            const V3 alk = alp + qd * cross(wp, z[k]);
            const M3 Rk = mul(Rp, axis_rotation(ax, LD1(in.q, d)));
            const V3 rho = mul(Rk, v3(kLinkCom[leg][k][0], kLinkCom[leg][k][1], kLinkCom[leg][k][2]));
            c[k] = pj[k] + rho;
            const V3 wxr = cross(wk, rho);
            const V3 vc = vj + wxr;
            const V3 ac = aj + cross(alk, rho) + cross(wk, wxr);
            Iw[k] = rotate_inertia(Rk, kLinkInertia[leg][k]);
            const double m = kLinkMass[leg][k];
            // accumulate whole-robot quantities about the base origin
            mr = mr + m * c[k];
            mvrel = mvrel + m * (vc - v0);
            const double r2 = dot(c[k], c[k]);
            Ibxx += Iw[k].xx + m * (r2 - c[k].x * c[k].x);
            Ibyy += Iw[k].yy + m * (r2 - c[k].y * c[k].y);
            Ibzz += Iw[k].zz + m * (r2 - c[k].z * c[k].z);
            Ibxy += Iw[k].xy - m * c[k].x * c[k].y;
            Ibxz += Iw[k].xz - m * c[k].x * c[k].z;
            Ibyz += Iw[k].yz - m * c[k].y * c[k].z;
            // Newton-Euler with nu_dot = 0
            f[k] = m * (ac - grav);
            n[k] = mul(Iw[k], alk) + cross(wk, mul(Iw[k], wk));
            fg[k] = (-m) * grav;
            ftot = ftot + f[k];
            ntot = ntot + n[k] + cross(c[k], f[k]);
            gftot = gftot + fg[k];
            gntot = gntot + cross(c[k], fg[k]);
            Rp = Rk; pp = pj[k]; vp = vj; ap = aj; wp = wk; alp = alk;
        }
        {   // foot frame (fixed to the lower leg)
            const V3 off = mul(Rp, v3(kFootOffset[leg][0], kFootOffset[leg][1], kFootOffset[leg][2]));
            const V3 wxo = cross(wp, off);
            footp[leg] = pp + off;
            footv[leg] = vp + wxo;
            foota[leg] = ap + cross(alp, off) + cross(wp, wxo);
            footR[leg] = Rp;
            for (int k = 0; k < 3; k++) {
                const V3 col = cross(z[k], footp[leg] - pj[k]);
                Jleg[leg][0 * 3 + k] = col.x; Jleg[leg][1 * 3 + k] = col.y; Jleg[leg][2 * 3 + k] = col.z;
            }
        }
        // joint-space projections for this leg
        for (int j = 0; j < 3; j++) {
            const int dj = dof_index(leg, j);
            // bias and gravity torque: z_j . sum_{i>=j} (n_i + (c_i - p_j) x f_i)
            V3 acc = v3(0, 0, 0), accg = v3(0, 0, 0);
            // CRBA column j: momentum of bodies i >= j under unit rate of joint j
            V3 Lj = v3(0, 0, 0), Bj = v3(0, 0, 0);   // linear momentum; angular momentum about the base origin
            for (int b = j; b < 3; b++) {
                const V3 d = c[b] - pj[j];
                acc = acc + n[b] + cross(d, f[b]);
                accg = accg + cross(d, fg[b]);
                const V3 l = kLinkMass[leg][b] * cross(z[j], d);
                Lj = Lj + l;
                Bj = Bj + mul(Iw[b], z[j]) + cross(c[b], l);
            }
            hj[dj] = dot(z[j], acc);
            gj[dj] = dot(z[j], accg);
            Mbj[0 * 12 + dj] = Lj.x; Mbj[1 * 12 + dj] = Lj.y; Mbj[2 * 12 + dj] = Lj.z;
            Mbj[3 * 12 + dj] = Bj.x; Mbj[4 * 12 + dj] = Bj.y; Mbj[5 * 12 + dj] = Bj.z;
            for (int k = 0; k <= j; k++) {
                const int dk = dof_index(leg, k);
                const double v = dot(z[k], Bj - cross(pj[k], Lj));
                Mjj[dk * 12 + dj] = v;
                Mjj[dj * 12 + dk] = v;
            }
        }
    }
    // ---- base block of the mass matrix, base rows of h and g
    const double mtot = kTotalMass;
    for (int k = 0; k < 36; k++) Mb[k] = 0.0;
    Mb[0] = Mb[7] = Mb[14] = mtot;
    // Mb[0:3,3:6] = -S(mr), Mb[3:6,0:3] = S(mr)
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
    hb[0] = ftot.x; hb[1] = ftot.y; hb[2] = ftot.z; hb[3] = ntot.x; hb[4] = ntot.y; hb[5] = ntot.z;
    gb[0] = gftot.x; gb[1] = gftot.y; gb[2] = gftot.z; gb[3] = gntot.x; gb[4] = gntot.y; gb[5] = gntot.z;

    D.mr = mr; D.mvrel = mvrel;
#undef LD1
}

// One instance.  `i` indexes the SoA arrays; `rec` points at this instance's QP record.
WBC_DEVFN inline void front_cycle(const Params& P, const DevInputs& in, const FrontState& st, long i, double* rec,
                                double* w_out, long w_ld, const DevDebug* dbg)
{
    using namespace dogbot;
    const long ld = in.ld;
#define LD1(ptr, k) (ptr)[(long)(k) * ld + i]
    FrontDyn D;
    front_dynamics(P, in, i, D);
    const V3 p0 = D.p0, v0 = D.v0, w0 = D.w0;
    const M3& R0 = D.R0;
    const int mode = in.mode[i];
    double (&Mb)[36] = D.Mb; double (&Mbj)[72] = D.Mbj; double (&Mjj)[144] = D.Mjj;
    double (&hb)[6] = D.hb; double (&hj)[12] = D.hj; double (&gb)[6] = D.gb; double (&gj)[12] = D.gj;
    const V3 mr = D.mr, mvrel = D.mvrel;
    V3 (&footp)[4] = D.footp; V3 (&footv)[4] = D.footv; V3 (&foota)[4] = D.foota;
    M3 (&footR)[4] = D.footR;
    double (&Jleg)[4][9] = D.Jleg;
    const double mtot = kTotalMass;
    // CoM (getCenterOfMassPosition / Velocity, main.cpp:597-602)
    const V3 xbc = (1.0 / mtot) * mr;                 // com - base            main.cpp:518
    const V3 xbcd = (1.0 / mtot) * mvrel;             // com_vel - v_base      main.cpp:538
    const V3 com = p0 + xbc;
    const V3 comv = v0 + xbcd;

    // ---- computeTransformation + CoM-coordinate quantities (main.cpp:491-568, 645-659), closed form
    double Lc[36];
    for (int k = 0; k < 36; k++) Lc[k] = Mb[k];
    chol6(Lc);
    double Pm[72];                                    // P = Mb^-1 Mbj   (bdcSvd solve at main.cpp:528)
    double pv[6];                                     // P * dq_j
    for (int k = 0; k < 6; k++) pv[k] = 0.0;
    for (int cidx = 0; cidx < 12; cidx++) {
        double col[6];
        for (int k = 0; k < 6; k++) col[k] = Mbj[k * 12 + cidx];
        chol6_solve(Lc, col);
        const double qd = LD1(in.dq, cidx);
        for (int k = 0; k < 6; k++) { Pm[k * 12 + cidx] = col[k]; pv[k] += col[k] * qd; }
    }
    // u3 = first three entries of T_inv_dot * dq (main.cpp:565-566, 648, 658; dq[3:6] = base omega, quirk E19)
    V3 u3;
    {
        const V3 mdr = mtot * xbcd;                   // main.cpp:539
        const V3 pvl = v3(pv[0], pv[1], pv[2]), pva = v3(pv[3], pv[4], pv[5]);
        const V3 y0 = (-1.0) * cross(mdr, pva);       // dMb * pv, dMb = [0, S(mdr)'; S(mdr), 0]   main.cpp:557-558
        const V3 y1 = cross(mdr, pvl);
        double zz[6] = {y0.x, y0.y, y0.z, y1.x, y1.y, y1.z};
        chol6_solve(Lc, zz);                          // Mb^-1 dMb Mb^-1 Mbj dq_j                   main.cpp:560-561
        const V3 zl = v3(zz[0], zz[1], zz[2]), za = v3(zz[3], zz[4], zz[5]);
        // dJs[0:3,:] dq_j = S(xbcd)' pva - [I, S(xbc)'] zz                                       main.cpp:563
        const V3 dJs = (-1.0) * cross(xbcd, pva) - (zl - cross(xbc, za));
        u3 = cross(xbcd, w0) - dJs;
    }
    // Mc = Xi' Mb Xi  (MassMatrixCOM[0:6,0:6], main.cpp:645)
    double Mc[36];
    {
        double MX[36];
        // MX = Mb * Xi, Xi = [I, S(xbc); 0, I]: columns 3..5 += Mb[:,0:3] * S(xbc)
        const double S[9] = {0, -xbc.z, xbc.y, xbc.z, 0, -xbc.x, -xbc.y, xbc.x, 0};
        for (int r = 0; r < 6; r++)
            for (int cc = 0; cc < 3; cc++) {
                MX[r * 6 + cc] = Mb[r * 6 + cc];
                MX[r * 6 + 3 + cc] = Mb[r * 6 + 3 + cc] + Mb[r * 6] * S[cc] + Mb[r * 6 + 1] * S[3 + cc] + Mb[r * 6 + 2] * S[6 + cc];
            }
        // Mc = Xi' MX: rows 3..5 += S(xbc)' * MX[0:3,:]
        for (int cc = 0; cc < 6; cc++) {
            Mc[0 * 6 + cc] = MX[0 * 6 + cc]; Mc[1 * 6 + cc] = MX[1 * 6 + cc]; Mc[2 * 6 + cc] = MX[2 * 6 + cc];
            for (int r = 0; r < 3; r++)
                Mc[(3 + r) * 6 + cc] = MX[(3 + r) * 6 + cc] + S[0 * 3 + r] * MX[0 * 6 + cc] + S[1 * 3 + r] * MX[1 * 6 + cc] + S[2 * 3 + r] * MX[2 * 6 + cc];
        }
    }
    // Mjj_com = Mjj - Mbj' P  (MassMatrixCOM[6:18,6:18])
    double Mjc[144];
    for (int a = 0; a < 12; a++)
        for (int b = 0; b < 12; b++) {
            double s = Mjj[a * 12 + b];
            for (int k = 0; k < 6; k++) s -= Mbj[k * 12 + a] * Pm[k * 12 + b];
            Mjc[a * 12 + b] = s;
        }
    // BiasCOM = T^-T (h + M T_inv_dot dq),  GravMatrixCOM = T^-T g   (main.cpp:648, 651)
    double hc[18], gc[18];
    {
        double hb2[6];
        for (int k = 0; k < 6; k++) hb2[k] = hb[k] + Mb[k * 6] * u3.x + Mb[k * 6 + 1] * u3.y + Mb[k * 6 + 2] * u3.z;
        const V3 t = cross(xbc, v3(hb2[0], hb2[1], hb2[2]));
        hc[0] = hb2[0]; hc[1] = hb2[1]; hc[2] = hb2[2];
        hc[3] = hb2[3] - t.x; hc[4] = hb2[4] - t.y; hc[5] = hb2[5] - t.z;
        const V3 tg = cross(xbc, v3(gb[0], gb[1], gb[2]));
        gc[0] = gb[0]; gc[1] = gb[1]; gc[2] = gb[2];
        gc[3] = gb[3] - tg.x; gc[4] = gb[4] - tg.y; gc[5] = gb[5] - tg.z;
        for (int a = 0; a < 12; a++) {
            double s = hj[a] + Mbj[0 * 12 + a] * u3.x + Mbj[1 * 12 + a] * u3.y + Mbj[2 * 12 + a] * u3.z;
            double sg = gj[a];
            for (int k = 0; k < 6; k++) { s -= Pm[k * 12 + a] * hb2[k]; sg -= Pm[k * 12 + a] * gb[k]; }
            hc[6 + a] = s;
            gc[6 + a] = sg;
        }
    }
    // JacCOM_lin = Jac_lin T^-1 (main.cpp:654-655), JdqdCOM_lin = Jdqd + Jac T_inv_dot dq (658-659);
    // stacked foot order BR, BL, FL, FR (main.cpp:674-686)
    double Jc[72], Jj[144], Jd[12];
    for (int sf = 0; sf < 4; sf++) {
        const int leg = kFootLeg[sf];
        const V3 rf = footp[leg];                // foot - base origin
        const V3 rc = rf - xbc;                  // foot - com
        // base columns: [I, -S(rc)]
        const double nS[9] = {0, rc.z, -rc.y, -rc.z, 0, rc.x, rc.y, -rc.x, 0};
        // -Jb P + Jj, Jb = [I, -S(rf)]
        const double nSf[9] = {0, rf.z, -rf.y, -rf.z, 0, rf.x, rf.y, -rf.x, 0};
        for (int a = 0; a < 3; a++) {
            const int r = 3 * sf + a;
            for (int b = 0; b < 3; b++) { Jc[r * 6 + b] = (a == b) ? 1.0 : 0.0; Jc[r * 6 + 3 + b] = nS[a * 3 + b]; }
            for (int b = 0; b < 12; b++) {
                double s = Pm[a * 12 + b] + nSf[a * 3 + 0] * Pm[3 * 12 + b] + nSf[a * 3 + 1] * Pm[4 * 12 + b] + nSf[a * 3 + 2] * Pm[5 * 12 + b];
                Jj[r * 12 + b] = -s;
            }
            for (int k = 0; k < 3; k++) Jj[r * 12 + dof_index(leg, k)] += Jleg[leg][a * 3 + k];
            Jd[r] = comp(foota[leg], a) + comp(u3, a);
        }
    }

    // ---- Fgrf (main.cpp:1022-1026; swing feet zeroed 1218 / 1765)
    double Fg[12];
    for (int sf = 0; sf < 4; sf++) {
        const bool swing = (mode == MODE_SWING_BR_FL && (sf == 0 || sf == 2)) || (mode == MODE_SWING_BL_FR && (sf == 1 || sf == 3));
        V3 fw = v3(0, 0, 0);
        if (!swing) fw = mul(footR[kFootLeg[sf]], v3(LD1(in.foot_force, 3 * sf), LD1(in.foot_force, 3 * sf + 1), LD1(in.foot_force, 3 * sf + 2)));
        Fg[3 * sf] = fw.x; Fg[3 * sf + 1] = fw.y; Fg[3 * sf + 2] = fw.z;
    }
    // ---- estimate() (main.cpp:692-725)
    const double comv6[6] = {comv.x, comv.y, comv.z, w0.x, w0.y, w0.z};     // CoM_vel, main.cpp:602
    double west[6], rho6[6], dd6[6];
    bool finite_obs = true;
    for (int a = 0; a < 6; a++) {
        double rho = 0.0, fc = 0.0;
        for (int b = 0; b < 6; b++) rho += Mc[a * 6 + b] * comv6[b];        // 699, 705
        for (int r = 0; r < 12; r++) fc += Jc[r * 6 + a] * Fg[r];           // 696-698
        rho6[a] = rho;
        dd6[a] = -mtot * (a == 2 ? P.g_acc : 0.0) + fc;                     // 700-706
    }
    if (P.observer_enabled) {
        const double T = P.obs_dt, k0 = in.obs_gain ? in.obs_gain[i] : P.obs_gain;
        const double mgain = (1.0 / (1.0 + k0 * T)) * k0;                   // (I + k0 T)^-1 k0, main.cpp:716
        const bool second = P.obs_order == 2, expl = P.obs_form == 1;
        const double k2 = P.obs_gain2;
        double ydn[6], ywn[6], ygn[6];
        for (int a = 0; a < 6; a++) {
            const double yd = st.yd[(long)a * st.ld + i] + dd6[a] * T;      // 717
            double wv;
            if (!second) {
                if (!expl) wv = mgain * (rho6[a] - st.yw[(long)a * st.ld + i] - yd);   // 718
                else wv = k0 * (rho6[a] - st.yw[(long)a * st.ld + i] - yd);            // estimator_sem.cpp:57
                ygn[a] = 0.0;
            } else {
                // second order: gamma1 = k0 (rho - yw - yd), ygamma' = gamma1 - w, w = k2 ygamma, discretised like
                // the first-order forms: backward Euler with the unknown w on both sides (form 0), or forward Euler (form 1)
                const double e = rho6[a] - st.yw[(long)a * st.ld + i] - yd;
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
            ywn[a] = st.yw[(long)a * st.ld + i] + wv * T;                   // 719
            west[a] = wv;
            finite_obs = finite_obs && (yd - yd == 0.0) && (ywn[a] - ywn[a] == 0.0) && (ygn[a] - ygn[a] == 0.0);
        }
        // the carried state (main.cpp:721-724) only advances on finite values: a NaN among the inputs must not poison the
        // instance's observer for the rest of the run.  The estimate itself stays non-finite here, so Wcom_des and with it the
        // QP record are, and the solver kernel flags the instance (WBC_ST_NONFINITE) instead of solving it.
        if (finite_obs) {
            for (int a = 0; a < 6; a++) { st.yd[(long)a * st.ld + i] = ydn[a]; st.yw[(long)a * st.ld + i] = ywn[a]; }
            if (second)
                for (int a = 0; a < 6; a++) st.yg[(long)a * st.ld + i] = ygn[a];
        }
    } else {
        for (int a = 0; a < 6; a++) west[a] = 0.0;
    }
    if (st.w3) {
        // getw3 (estimator_sem.cpp:64-70): w3 = pinv(J)' w, J = JacCOM_lin[:, 0:6] (12 x 6, full column rank whenever the feet span
        // more than a line), so pinv(J)' = J (J'J)^-1: one 6 x 6 Cholesky.  A non-positive pivot (degenerate foot geometry) gives NaN.
        double G[36], y[6];
        for (int a = 0; a < 6; a++)
            for (int b = 0; b <= a; b++) {
                double g = 0.0;
                for (int r = 0; r < 12; r++) g += Jc[r * 6 + a] * Jc[r * 6 + b];
                G[a * 6 + b] = g; G[b * 6 + a] = g;
            }
        chol6(G);
        for (int a = 0; a < 6; a++) y[a] = finite_obs ? west[a] : 0.0;
        chol6_solve(G, y);
        for (int r = 0; r < 12; r++) {
            double v = 0.0;
            for (int a = 0; a < 6; a++) v += Jc[r * 6 + a] * y[a];
            st.w3[(long)r * st.w3_ld + i] = v;
        }
    }
    for (int a = 0; a < 6; a++) w_out[(long)a * w_ld + i] = finite_obs ? west[a] : 0.0;

    // ---- Wcom_des (main.cpp:1012-1032)
    double Wc[6];
    {
        double dx[6], dv[6], ades[6];
        dx[0] = LD1(in.com_des_pos, 0) - com.x; dx[1] = LD1(in.com_des_pos, 1) - com.y; dx[2] = LD1(in.com_des_pos, 2) - com.z;
        const V3 dr = mul(R0, v3(LD1(in.com_des_pos, 3) - LD1(in.base_rpy, 0), LD1(in.com_des_pos, 4) - LD1(in.base_rpy, 1),
                                 LD1(in.com_des_pos, 5) - LD1(in.base_rpy, 2)));                    // main.cpp:1013
        dx[3] = dr.x; dx[4] = dr.y; dx[5] = dr.z;
        for (int a = 0; a < 6; a++) { dv[a] = LD1(in.com_des_vel, a) - comv6[a]; ades[a] = LD1(in.com_des_acc, a); }
        for (int a = 0; a < 6; a++) {
            double ma = 0.0;
            for (int b = 0; b < 6; b++) ma += Mc[a * 6 + b] * ades[b];
            Wc[a] = P.kcom * dx[a] + P.dcom * dv[a] + mtot * (a == 2 ? P.g_acc : 0.0) + ma - west[a];
        }
    }

    // ---- emit the QP record
    for (int k = 0; k < 36; k++) rec[QR_MC + k] = Mc[k];
    for (int k = 0; k < 6; k++) rec[QR_HC + k] = hc[k];
    for (int k = 0; k < 12; k++) rec[QR_HJ + k] = hc[6 + k];
    for (int k = 0; k < 144; k++) rec[QR_MJJ + k] = Mjc[k];
    for (int k = 0; k < 72; k++) rec[QR_JC + k] = Jc[k];
    for (int k = 0; k < 144; k++) rec[QR_JJ + k] = Jj[k];
    for (int k = 0; k < 12; k++) rec[QR_JDQD + k] = Jd[k];
    for (int k = 0; k < 6; k++) rec[QR_WCOM + k] = Wc[k];
    {
        const double dt = P.joint_dt, kk = 2.0 / (dt * dt);                  // main.cpp:1098-1104
        for (int j = 0; j < 12; j++) {
            const double qj = LD1(in.q, j), dqj = LD1(in.dq, j);
            rec[QR_DDQMAX + j] = kk * (kQmax[j] - qj - dt * dqj);
            rec[QR_DDQMIN + j] = kk * (kQmin[j] - qj - dt * dqj);
        }
    }
    {   // swing-foot PD (main.cpp:1327-1379); the two swing feet in Jsw row order
        const int sf0 = (mode == MODE_SWING_BL_FR) ? 1 : 0, sf1 = (mode == MODE_SWING_BL_FR) ? 3 : 2;
        for (int a = 0; a < 6; a++) {
            const int sf = a < 3 ? sf0 : sf1, ax = a % 3, leg = kFootLeg[sf];
            const double pos = comp(p0, ax) + comp(footp[leg], ax), vel = comp(footv[leg], ax);
            const double vdot = LD1(in.sw_des_acc, a) + P.kd_sw * (LD1(in.sw_des_vel, a) - vel) + P.kp_sw * (LD1(in.sw_des_pos, a) - pos);
            rec[QR_SWRHS + a] = (mode == MODE_STANCE) ? 0.0 : vdot - Jd[3 * sf + ax];
        }
    }
    for (int sf = 0; sf < 4; sf++) {   // friction pyramid rows (main.cpp:1062-1078); per-foot terrain frames generalise n=(0,0,1)
        double nn[3] = {0, 0, 1}, t1[3] = {1, 0, 0}, t2[3] = {0, 1, 0}, mu = P.mu;
        if (in.terrain) {
            for (int k = 0; k < 3; k++) { nn[k] = LD1(in.terrain, 10 * sf + k); t1[k] = LD1(in.terrain, 10 * sf + 3 + k); t2[k] = LD1(in.terrain, 10 * sf + 6 + k); }
            mu = LD1(in.terrain, 10 * sf + 9);
        }
        for (int k = 0; k < 3; k++) {
            rec[QR_CFR + 15 * sf + 0 + k] = -mu * nn[k] + t1[k];
            rec[QR_CFR + 15 * sf + 3 + k] = -mu * nn[k] + t2[k];
            rec[QR_CFR + 15 * sf + 6 + k] = -(mu * nn[k] + t1[k]);
            rec[QR_CFR + 15 * sf + 9 + k] = -(mu * nn[k] + t2[k]);
            rec[QR_CFR + 15 * sf + 12 + k] = -nn[k];
        }
    }
    rec[QR_MODE] = (double)mode;
    for (int a = 0; a < 6; a++) { rec[QR_RHO + a] = rho6[a]; rec[QR_DD + a] = dd6[a]; }
    rec[QR_XBC] = xbc.x; rec[QR_XBC + 1] = xbc.y; rec[QR_XBC + 2] = xbc.z;
    for (int sf = 0; sf < 4; sf++)
        for (int k = 0; k < 9; k++) rec[QR_FOOTR + 9 * sf + k] = footR[kFootLeg[sf]].m[k];

    if (dbg) {
        const long dl = dbg->ld;
#define ST1(ptr, k, v) (ptr)[(long)(k) * dl + i] = (v)
        for (int a = 0; a < 18; a++)
            for (int b = 0; b < 18; b++) {
                double v;
                if (a < 6 && b < 6) v = Mb[a * 6 + b];
                else if (a < 6) v = Mbj[a * 12 + (b - 6)];
                else if (b < 6) v = Mbj[b * 12 + (a - 6)];
                else v = Mjj[(a - 6) * 12 + (b - 6)];
                ST1(dbg->M, a * 18 + b, v);
            }
        for (int a = 0; a < 6; a++) { ST1(dbg->h, a, hb[a]); ST1(dbg->g, a, gb[a]); }
        for (int a = 0; a < 12; a++) { ST1(dbg->h, 6 + a, hj[a]); ST1(dbg->g, 6 + a, gj[a]); }
        for (int sf = 0; sf < 4; sf++) {
            const int leg = kFootLeg[sf];
            const V3 rf = footp[leg];
            const double nSf[9] = {0, rf.z, -rf.y, -rf.z, 0, rf.x, rf.y, -rf.x, 0};
            for (int a = 0; a < 3; a++) {
                const int r = 3 * sf + a;
                for (int b = 0; b < 18; b++) ST1(dbg->Jac_lin, r * 18 + b, 0.0);
                ST1(dbg->Jac_lin, r * 18 + a, 1.0);
                for (int b = 0; b < 3; b++) ST1(dbg->Jac_lin, r * 18 + 3 + b, nSf[a * 3 + b]);
                for (int k = 0; k < 3; k++) ST1(dbg->Jac_lin, r * 18 + 6 + dof_index(leg, k), Jleg[leg][a * 3 + k]);
                ST1(dbg->Jdqd_lin, r, comp(foota[leg], a));
                ST1(dbg->foot_pos, r, comp(p0, a) + comp(rf, a));
                ST1(dbg->foot_vel, r, comp(footv[leg], a));
                for (int b = 0; b < 6; b++) ST1(dbg->Jcom_lin, r * 18 + b, Jc[r * 6 + b]);
                for (int b = 0; b < 12; b++) ST1(dbg->Jcom_lin, r * 18 + 6 + b, Jj[r * 12 + b]);
                ST1(dbg->Jdqdcom_lin, r, Jd[r]);
                ST1(dbg->Fgrf, r, Fg[r]);
            }
        }
        for (int a = 0; a < 3; a++) { ST1(dbg->com, a, comp(com, a)); ST1(dbg->com_vel, a, comp(comv, a)); }
        for (int k = 0; k < 36; k++) ST1(dbg->Mcom_b, k, Mc[k]);
        for (int k = 0; k < 144; k++) ST1(dbg->Mcom_j, k, Mjc[k]);
        for (int k = 0; k < 18; k++) { ST1(dbg->hcom, k, hc[k]); ST1(dbg->gcom, k, gc[k]); }
        for (int k = 0; k < 6; k++) ST1(dbg->Wcom_des, k, Wc[k]);
#undef ST1
    }
#undef LD1
}

// ---- Forward dynamics with hard contacts: the closed-loop plant that stands in for Gazebo + ModelPush
// (force_plugin.cpp:124-491).  One control period of instance i under joint torques tau, a world wrench `push` at the CoM and
// rigid bilateral point contacts at the stance feet of its mode; the state arrays of `in` are advanced IN PLACE
// (base_pos, base_rot, base_rpy, base_vel, q, dq) and foot_force receives the contact forces in the sensor frames.
//     M nu_dot + h = S'tau + push_gen + Js' f,     Js nu_dot = -Jdqd_s - gamma Js nu
// solved through the Schur complement of the contact forces: M = L L' (18 x 18 Cholesky), Y = L^-1 Js', A = Y'Y,
// f = A^-1 (c - Js M^-1 b), nu_dot = M^-1 (b + Js' f) -- a different route from the oracle's dense 30 x 30 elimination
// (the test oracle restates the same step with one 30 x 30 Gauss-Jordan).  nsub semi-implicit Euler substeps.  diag (optional, [2][ld]): largest
// contact-constraint residual, smallest normal force.
struct FdynIO {
    double *base_pos, *base_rot, *base_rpy, *base_vel, *q, *dq, *foot_force;   // state, SoA [k][ld], read and written
    const int* mode;
    const double *tau, *push;           // [12][ld], [6][ld]
    double* diag;                       // [2][ld] or nullptr
    long ld;
};

WBC_DEVFN inline void fdyn_step_instance(const Params& P, const FdynIO& io, long i, int nsub, double gamma)
{
    using namespace dogbot;
    const long ld = io.ld;
    DevInputs in;
    in.base_pos = io.base_pos; in.base_rot = io.base_rot; in.base_rpy = io.base_rpy; in.base_vel = io.base_vel; in.q = io.q; in.dq = io.dq;
    in.com_des_pos = in.com_des_vel = in.com_des_acc = in.sw_des_pos = in.sw_des_vel = in.sw_des_acc = nullptr;
    in.foot_force = io.foot_force; in.terrain = nullptr; in.mode = io.mode; in.obs_gain = nullptr; in.ld = ld;
    const int mode = io.mode[i];
    int stance[4], ns = 0;
    for (int sf = 0; sf < 4; sf++) {
        const bool swing = (mode == MODE_SWING_BR_FL && (sf == 0 || sf == 2)) || (mode == MODE_SWING_BL_FR && (sf == 1 || sf == 3));
        stance[sf] = swing ? 0 : 1;
        ns += stance[sf];
    }
    const int nc = 3 * ns;
    const double dt = P.obs_dt / (double)nsub;
    double tau[12], push[6];
    for (int k = 0; k < 12; k++) tau[k] = io.tau[(long)k * ld + i];
    for (int k = 0; k < 6; k++) push[k] = io.push[(long)k * ld + i];
    double worst = 0.0, fzmin = 1.0e300;
    for (int it = 0; it < nsub; it++) {
        FrontDyn D;
        front_dynamics(P, in, i, D);
        // M = [Mb, Mbj; Mbj', Mjj] with Mjj block diagonal (one 3 x 3 block per leg: the legs only couple through the base), so
        // M^-1 is applied by block elimination: four 3 x 3 Cholesky factors, Z_l = Mjj_l^-1 Mbj_l', the 6 x 6 Schur complement
        // Sb = Mb - sum_l Mbj_l Z_l (one more 6 x 6 Cholesky) -- instead of a dense 18 x 18 factorisation.
        double Lj[4][6], Z[4][18];
        double Sb[36];
        for (int k = 0; k < 36; k++) Sb[k] = D.Mb[k];
        for (int leg = 0; leg < 4; leg++) {
            const int i0 = dof_index(leg, 0), i1 = dof_index(leg, 1), i2 = i1 + 1;
            double* l = Lj[leg];                       // lower factor: l00, l10, l11, l20, l21, l22 (diagonal entries stored as reciprocals)
            const double m00 = D.Mjj[i0 * 12 + i0], m10 = D.Mjj[i1 * 12 + i0], m11 = D.Mjj[i1 * 12 + i1];
            const double m20 = D.Mjj[i2 * 12 + i0], m21 = D.Mjj[i2 * 12 + i1], m22 = D.Mjj[i2 * 12 + i2];
            const double d0 = sqrt(m00);
            l[0] = 1.0 / d0; l[1] = m10 * l[0]; l[3] = m20 * l[0];
            const double d1 = sqrt(m11 - l[1] * l[1]);
            l[2] = 1.0 / d1; l[4] = (m21 - l[3] * l[1]) * l[2];
            const double d2 = sqrt(m22 - l[3] * l[3] - l[4] * l[4]);
            l[5] = 1.0 / d2;
            for (int k = 0; k < 6; k++) {
                double y0 = D.Mbj[k * 12 + i0] * l[0];
                double y1 = (D.Mbj[k * 12 + i1] - l[1] * y0) * l[2];
                double y2 = (D.Mbj[k * 12 + i2] - l[3] * y0 - l[4] * y1) * l[5];
                y2 = y2 * l[5]; y1 = (y1 - l[4] * y2) * l[2]; y0 = (y0 - l[1] * y1 - l[3] * y2) * l[0];
                Z[leg][0 * 6 + k] = y0; Z[leg][1 * 6 + k] = y1; Z[leg][2 * 6 + k] = y2;
            }
            for (int r = 0; r < 6; r++)
                for (int c2 = 0; c2 < 6; c2++)
                    Sb[r * 6 + c2] -= D.Mbj[r * 12 + i0] * Z[leg][0 * 6 + c2] + D.Mbj[r * 12 + i1] * Z[leg][1 * 6 + c2] + D.Mbj[r * 12 + i2] * Z[leg][2 * 6 + c2];
        }
        chol6(Sb);
        // x = M^-1 [rb; rj] for a right-hand side whose joint part lives on the legs flagged in `legs` (bit l)
        auto solveM = [&](const double* rb, const double* rj, unsigned legs, double* xb, double* xj) {
            double t[6], y[12];
            for (int k = 0; k < 6; k++) t[k] = rb[k];
            for (int leg = 0; leg < 4; leg++) {
                const int i0 = dof_index(leg, 0), i1 = dof_index(leg, 1), i2 = i1 + 1;
                if (!(legs & (1u << leg))) { y[i0] = y[i1] = y[i2] = 0.0; continue; }
                const double* l = Lj[leg];
                double y0 = rj[i0] * l[0];
                double y1 = (rj[i1] - l[1] * y0) * l[2];
                double y2 = (rj[i2] - l[3] * y0 - l[4] * y1) * l[5];
                y2 = y2 * l[5]; y1 = (y1 - l[4] * y2) * l[2]; y0 = (y0 - l[1] * y1 - l[3] * y2) * l[0];
                y[i0] = y0; y[i1] = y1; y[i2] = y2;
                for (int r = 0; r < 6; r++) t[r] -= D.Mbj[r * 12 + i0] * y0 + D.Mbj[r * 12 + i1] * y1 + D.Mbj[r * 12 + i2] * y2;
            }
            chol6_solve(Sb, t);
            for (int k = 0; k < 6; k++) xb[k] = t[k];
            for (int leg = 0; leg < 4; leg++) {
                const int i0 = dof_index(leg, 0), i1 = dof_index(leg, 1), i2 = i1 + 1;
                const int id[3] = {i0, i1, i2};
                for (int a = 0; a < 3; a++) {
                    double sacc = y[id[a]];
                    for (int k = 0; k < 6; k++) sacc -= Z[leg][a * 6 + k] * t[k];
                    xj[id[a]] = sacc;
                }
            }
        };
        // generalised velocity, right-hand side b = S'tau - h + push_gen
        double nu[18], b[18];
        nu[0] = D.v0.x; nu[1] = D.v0.y; nu[2] = D.v0.z; nu[3] = D.w0.x; nu[4] = D.w0.y; nu[5] = D.w0.z;
        for (int k = 0; k < 12; k++) nu[6 + k] = io.dq[(long)k * ld + i];
        const V3 xbc = (1.0 / kTotalMass) * D.mr;
        const V3 tq = cross(xbc, v3(push[0], push[1], push[2]));
        for (int k = 0; k < 6; k++) b[k] = -D.hb[k] + push[k];
        b[3] += tq.x; b[4] += tq.y; b[5] += tq.z;
        for (int k = 0; k < 12; k++) b[6 + k] = -D.hj[k] + tau[k];
        // stance rows of the linear foot Jacobian: base columns [I, -S(rf)] (Jb, 6 per row) and the leg's three joint columns (Jl)
        double Jb[12 * 6], Jl[12 * 3], c[12];
        int rleg[12];
        int r = 0;
        for (int sf = 0; sf < 4; sf++) {
            if (!stance[sf]) continue;
            const int leg = kFootLeg[sf];
            const V3 rf = D.footp[leg];
            const double nSf[9] = {0, rf.z, -rf.y, -rf.z, 0, rf.x, rf.y, -rf.x, 0};
            for (int a = 0; a < 3; a++, r++) {
                rleg[r] = leg;
                for (int k = 0; k < 3; k++) { Jb[r * 6 + k] = (k == a) ? 1.0 : 0.0; Jb[r * 6 + 3 + k] = nSf[a * 3 + k]; Jl[r * 3 + k] = D.Jleg[leg][a * 3 + k]; }
                double jn = 0.0;
                for (int k = 0; k < 6; k++) jn += Jb[r * 6 + k] * nu[k];
                for (int k = 0; k < 3; k++) jn += Jl[r * 3 + k] * nu[6 + dof_index(leg, k)];
                c[r] = -comp(D.foota[leg], a) - gamma * jn;
            }
        }
        // a0 = M^-1 b
        double a0[18];
        solveM(b, b + 6, 0xfu, a0, a0 + 6);
        // X = M^-1 Js' (column r in Xb[r], Xj[r]), A = Js X, rhs = c - Js a0
        double Xb[12 * 6], Xj[12 * 12], A[12 * 12], f[12];
        for (int rr = 0; rr < nc; rr++) {
            double rj[12];
            for (int k = 0; k < 12; k++) rj[k] = 0.0;
            for (int k = 0; k < 3; k++) rj[dof_index(rleg[rr], k)] = Jl[rr * 3 + k];
            solveM(Jb + rr * 6, rj, 1u << rleg[rr], Xb + rr * 6, Xj + rr * 12);
        }
        for (int p_ = 0; p_ < nc; p_++) {
            const int lp = rleg[p_];
            for (int q_ = 0; q_ <= p_; q_++) {
                double sacc = 0.0;
                for (int k = 0; k < 6; k++) sacc += Jb[p_ * 6 + k] * Xb[q_ * 6 + k];
                for (int k = 0; k < 3; k++) sacc += Jl[p_ * 3 + k] * Xj[q_ * 12 + dof_index(lp, k)];
                A[p_ * 12 + q_] = sacc;
            }
            double sacc = c[p_];
            for (int k = 0; k < 6; k++) sacc -= Jb[p_ * 6 + k] * a0[k];
            for (int k = 0; k < 3; k++) sacc -= Jl[p_ * 3 + k] * a0[6 + dof_index(lp, k)];
            f[p_] = sacc;
        }
        for (int j = 0; j < nc; j++) {                    // Cholesky of A, in place (lower)
            double d = A[j * 12 + j];
            for (int k = 0; k < j; k++) d -= A[j * 12 + k] * A[j * 12 + k];
            d = sqrt(d);
            A[j * 12 + j] = d;
            for (int a = j + 1; a < nc; a++) {
                double sacc = A[a * 12 + j];
                for (int k = 0; k < j; k++) sacc -= A[a * 12 + k] * A[j * 12 + k];
                A[a * 12 + j] = sacc / d;
            }
        }
        for (int a = 0; a < nc; a++) {
            double sacc = f[a];
            for (int k = 0; k < a; k++) sacc -= A[a * 12 + k] * f[k];
            f[a] = sacc / A[a * 12 + a];
        }
        for (int a = nc - 1; a >= 0; a--) {
            double sacc = f[a];
            for (int k = a + 1; k < nc; k++) sacc -= A[k * 12 + a] * f[k];
            f[a] = sacc / A[a * 12 + a];
        }
        // nu_dot = a0 + M^-1 Js' f = a0 + X f
        double nd[18];
        for (int a = 0; a < 18; a++) nd[a] = a0[a];
        for (int rr = 0; rr < nc; rr++) {
            for (int k = 0; k < 6; k++) nd[k] += Xb[rr * 6 + k] * f[rr];
            for (int k = 0; k < 12; k++) nd[6 + k] += Xj[rr * 12 + k] * f[rr];
        }
        for (int rr = 0; rr < nc; rr++) {
            double ja = 0.0;
            for (int k = 0; k < 6; k++) ja += Jb[rr * 6 + k] * nd[k];
            for (int k = 0; k < 3; k++) ja += Jl[rr * 3 + k] * nd[6 + dof_index(rleg[rr], k)];
            worst = fmax(worst, fabs(ja - c[rr]));
        }
        // contact forces in the sensor frames (what the contact sensors report, main.cpp:794-834)
        r = 0;
        for (int sf = 0; sf < 4; sf++) {
            double fs[3] = {0.0, 0.0, 0.0};
            if (stance[sf]) {
                const M3& Rf = D.footR[kFootLeg[sf]];
                for (int k = 0; k < 3; k++) fs[k] = Rf.m[k] * f[r] + Rf.m[3 + k] * f[r + 1] + Rf.m[6 + k] * f[r + 2];
                fzmin = fmin(fzmin, f[r + 2]);
                r += 3;
            }
            for (int k = 0; k < 3; k++) io.foot_force[(long)(3 * sf + k) * ld + i] = fs[k];
        }
        // semi-implicit Euler: velocities first, then the configuration with the new velocities
        for (int a = 0; a < 18; a++) nu[a] += dt * nd[a];
        for (int k = 0; k < 6; k++) io.base_vel[(long)k * ld + i] = nu[k];
        for (int k = 0; k < 12; k++) {
            io.dq[(long)k * ld + i] = nu[6 + k];
            io.q[(long)k * ld + i] += dt * nu[6 + k];
        }
        for (int k = 0; k < 3; k++) io.base_pos[(long)k * ld + i] += dt * nu[k];
        M3 Rn = D.R0;
        const double wn = sqrt(nu[3] * nu[3] + nu[4] * nu[4] + nu[5] * nu[5]);
        if (wn > 0.0) Rn = mul(axis_rotation(v3(nu[3] / wn, nu[4] / wn, nu[5] / wn), wn * dt), D.R0);
        for (int k = 0; k < 9; k++) io.base_rot[(long)k * ld + i] = Rn.m[k];
        io.base_rpy[0 * ld + i] = atan2(Rn.m[7], Rn.m[8]);
        io.base_rpy[1 * ld + i] = atan2(-Rn.m[6], sqrt(Rn.m[7] * Rn.m[7] + Rn.m[8] * Rn.m[8]));
        io.base_rpy[2 * ld + i] = atan2(Rn.m[3], Rn.m[0]);
    }
    if (io.diag) { io.diag[0 * ld + i] = worst; io.diag[1 * ld + i] = (ns > 0) ? fzmin : 0.0; }
}

}  // namespace wbc
