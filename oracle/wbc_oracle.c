/* TEST INFRASTRUCTURE ONLY -- see wbc_oracle.h for scope and parity status.
 *
 * CPU restatement of one control cycle of the reference controller:
 *   update()                main.cpp:572-660  (+ computeTransformation 491-568, computeJac 664-690,
 *                                               computeJacDotQDot 754-766, *linear 727-751)
 *   estimate()              main.cpp:692-725
 *   stance QP assembly      main.cpp:984-1120
 *   swing QP assembly       main.cpp:1163-1389 (second copy 1710-1915)
 *   torque map              main.cpp:1126, 1396
 * Dense, literal arithmetic (18x18 Gauss-Jordan inverse where the reference calls Eigen's
 * .inverse(), 6x6 LU where it calls bdcSvd().solve()); no structure is exploited on purpose, so
 * that this file shares no formulation with the CUDA kernels it checks.
 */
#define _POSIX_C_SOURCE 200809L
#include "wbc_oracle.h"
#include "dogbot_model_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ small dense helpers */
static void mat_mul(const double* A, const double* B, double* C, int m, int k, int n)
{ /* C[m x n] = A[m x k] * B[k x n], row-major */
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int l = 0; l < k; l++) s += A[i * k + l] * B[l * n + j];
            C[i * n + j] = s;
        }
}
static void mat_T(const double* A, double* At, int m, int n)
{
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) At[j * m + i] = A[i * n + j];
}
/* Gauss-Jordan with partial pivoting: solves A X = B in place (A n x n, B n x nrhs). */
static int gj_solve(double* A, double* B, int n, int nrhs)
{
    for (int c = 0; c < n; c++) {
        int piv = c;
        double best = fabs(A[c * n + c]);
        for (int r = c + 1; r < n; r++)
            if (fabs(A[r * n + c]) > best) { best = fabs(A[r * n + c]); piv = r; }
        if (best == 0.0) return -1;
        if (piv != c) {
            for (int j = 0; j < n; j++) { double t = A[c * n + j]; A[c * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            for (int j = 0; j < nrhs; j++) { double t = B[c * nrhs + j]; B[c * nrhs + j] = B[piv * nrhs + j]; B[piv * nrhs + j] = t; }
        }
        double inv = 1.0 / A[c * n + c];
        for (int j = 0; j < n; j++) A[c * n + j] *= inv;
        for (int j = 0; j < nrhs; j++) B[c * nrhs + j] *= inv;
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            double f = A[r * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; j++) A[r * n + j] -= f * A[c * n + j];
            for (int j = 0; j < nrhs; j++) B[r * nrhs + j] -= f * B[c * nrhs + j];
        }
    }
    return 0;
}
static int mat_inv(const double* A, double* Ainv, int n)
{
    double* W = (double*)malloc(sizeof(double) * n * n);
    memcpy(W, A, sizeof(double) * n * n);
    for (int i = 0; i < n * n; i++) Ainv[i] = 0.0;
    for (int i = 0; i < n; i++) Ainv[i * n + i] = 1.0;
    int rc = gj_solve(W, Ainv, n, n);
    free(W);
    return rc;
}
static void cross3(const double* a, const double* b, double* c)
{
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    c[0] = x; c[1] = y; c[2] = z;
}
static void skew3(const double* v, double* S)
{ /* S(v) w = v x w */
    S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}
static void rot_axis_angle(const double* a, double th, double* R)
{ /* Rodrigues, |a| = 1 */
    double c = cos(th), s = sin(th), v = 1.0 - c;
    R[0] = c + a[0] * a[0] * v;        R[1] = a[0] * a[1] * v - a[2] * s; R[2] = a[0] * a[2] * v + a[1] * s;
    R[3] = a[1] * a[0] * v + a[2] * s; R[4] = c + a[1] * a[1] * v;        R[5] = a[1] * a[2] * v - a[0] * s;
    R[6] = a[2] * a[0] * v - a[1] * s; R[7] = a[2] * a[1] * v + a[0] * s; R[8] = c + a[2] * a[2] * v;
}

void wbc_oracle_default_params(wbc_oracle_params* p)
{
    p->kcom = 2500.0; p->dcom = 50.0; p->q1_weight = 50.0; p->slack_weight = 100000000.0;
    p->mu = 0.6; p->tau_max = 60.0; p->joint_dt = 0.025; p->kp_sw = 300.0; p->kd_sw = 20.0;
    p->g_acc = 9.81; p->obs_gain = 10.0; p->obs_dt = 0.0025;
    p->obs_gain2 = 1.0; p->obs_order = 1; p->obs_form = 0;
    p->observer_enabled = 1; p->fix_swing_rhs = 0;
}

/* ------------------------------------------------------------------ rigid-body dynamics
 * iDynTree semantics restated, MIXED velocity representation
 * (main.cpp:292): nu = [v_base (origin velocity, world axes); omega_base (world axes); dq].
 * Everything is computed in world coordinates by projecting each link's Newton-Euler
 * equations through its CoM Jacobian:
 *   M = sum_i  m_i Jv_i' Jv_i + Jw_i' Iw_i Jw_i
 *   h = sum_i  Jv_i' m_i (abias_i - g) + Jw_i' (Iw_i alphabias_i + w_i x Iw_i w_i)
 *   g = sum_i  Jv_i' m_i (-g)
 * (getFreeFloatingMassMatrix main.cpp:620, generalizedBiasForces 622, generalizedGravityForces 628)
 */
typedef struct {
    double R[DB_NLINKS][9], p[DB_NLINKS][3], w[DB_NLINKS][3], v[DB_NLINKS][3];
    double al[DB_NLINKS][3], a[DB_NLINKS][3], z[DB_NLINKS][3];
} kin_t;

static void forward_kinematics(const wbc_oracle_in* in, kin_t* k)
{
    memcpy(k->R[0], in->base_R, sizeof(double) * 9);
    for (int c = 0; c < 3; c++) {
        k->p[0][c] = in->base_pos[c]; k->v[0][c] = in->base_vel[c]; k->w[0][c] = in->base_vel[3 + c];
        k->al[0][c] = 0.0; k->a[0][c] = 0.0; k->z[0][c] = 0.0;
    }
    for (int i = 1; i < DB_NLINKS; i++) {
        int par = DB_PARENT[i], dof = DB_DOF[i];
        double off[3], t[3], t2[3];
        mat_mul(k->R[par], DB_JOINT_XYZ[i], off, 3, 3, 1);
        for (int c = 0; c < 3; c++) k->p[i][c] = k->p[par][c] + off[c];
        cross3(k->w[par], off, t);
        for (int c = 0; c < 3; c++) k->v[i][c] = k->v[par][c] + t[c];
        cross3(k->al[par], off, t);
        double wxo[3];
        cross3(k->w[par], off, wxo);
        cross3(k->w[par], wxo, t2);
        for (int c = 0; c < 3; c++) k->a[i][c] = k->a[par][c] + t[c] + t2[c];
        if (dof >= 0) {
            double Rj[9];
            rot_axis_angle(DB_JOINT_AXIS[i], in->q[dof], Rj);
            mat_mul(k->R[par], Rj, k->R[i], 3, 3, 3);
            mat_mul(k->R[par], DB_JOINT_AXIS[i], k->z[i], 3, 3, 1);
            cross3(k->w[par], k->z[i], t);
            for (int c = 0; c < 3; c++) {
                k->w[i][c] = k->w[par][c] + k->z[i][c] * in->dq[dof];
                k->al[i][c] = k->al[par][c] + t[c] * in->dq[dof];
            }
        } else {
            memcpy(k->R[i], k->R[par], sizeof(double) * 9);
            for (int c = 0; c < 3; c++) { k->w[i][c] = k->w[par][c]; k->al[i][c] = k->al[par][c]; k->z[i][c] = 0.0; }
        }
    }
}

/* Jacobians (3x18 each) of a point P rigidly attached to link `link`. */
static void point_jacobian(const kin_t* k, int link, const double* P, double* Jv, double* Jw)
{
    memset(Jv, 0, sizeof(double) * 54);
    memset(Jw, 0, sizeof(double) * 54);
    double r[3], S[9];
    for (int c = 0; c < 3; c++) r[c] = P[c] - k->p[0][c];
    skew3(r, S);
    for (int a = 0; a < 3; a++) {
        Jv[a * 18 + a] = 1.0;
        for (int b = 0; b < 3; b++) Jv[a * 18 + 3 + b] = -S[a * 3 + b];
        Jw[a * 18 + 3 + a] = 1.0;
    }
    for (int l = link; l > 0; l = DB_PARENT[l]) {
        int dof = DB_DOF[l];
        if (dof < 0) continue;
        double d[3], zc[3];
        for (int c = 0; c < 3; c++) d[c] = P[c] - k->p[l][c];
        cross3(k->z[l], d, zc);
        for (int a = 0; a < 3; a++) { Jv[a * 18 + 6 + dof] = zc[a]; Jw[a * 18 + 6 + dof] = k->z[l][a]; }
    }
}

static void rigid_body_dynamics(const wbc_oracle_in* in, wbc_oracle_dyn* d, kin_t* k)
{
    forward_kinematics(in, k);
    memset(d->M, 0, sizeof(d->M));
    memset(d->h, 0, sizeof(d->h));
    memset(d->g, 0, sizeof(d->g));
    double msum = 0.0;
    for (int c = 0; c < 3; c++) { d->com[c] = 0.0; d->com_vel[c] = 0.0; }
    for (int i = 0; i < DB_NLINKS; i++) {
        double m = DB_MASS[i], rho[3], P[3], Jv[54], Jw[54], t[3], t2[3];
        mat_mul(k->R[i], DB_COM[i], rho, 3, 3, 1);
        for (int c = 0; c < 3; c++) P[c] = k->p[i][c] + rho[c];
        point_jacobian(k, i, P, Jv, Jw);
        /* world inertia about the link CoM */
        double Id[9] = {DB_INERTIA_DIAG[i][0], 0, 0, 0, DB_INERTIA_DIAG[i][1], 0, 0, 0, DB_INERTIA_DIAG[i][2]};
        double RI[9], Rt[9], Iw[9];
        mat_mul(k->R[i], Id, RI, 3, 3, 3);
        mat_T(k->R[i], Rt, 3, 3);
        mat_mul(RI, Rt, Iw, 3, 3, 3);
        /* CoM velocity and bias acceleration */
        double vc[3], ac[3];
        cross3(k->w[i], rho, t);
        for (int c = 0; c < 3; c++) vc[c] = k->v[i][c] + t[c];
        cross3(k->al[i], rho, t2);
        double wxr[3], wwr[3];
        cross3(k->w[i], rho, wxr);
        cross3(k->w[i], wxr, wwr);
        for (int c = 0; c < 3; c++) ac[c] = k->a[i][c] + t2[c] + wwr[c];
        for (int c = 0; c < 3; c++) { d->com[c] += m * P[c]; d->com_vel[c] += m * vc[c]; }
        msum += m;
        /* M += m Jv'Jv + Jw' Iw Jw */
        double IJ[54];
        mat_mul(Iw, Jw, IJ, 3, 3, 18);
        for (int r = 0; r < 18; r++)
            for (int c = 0; c < 18; c++) {
                double s = 0.0;
                for (int a = 0; a < 3; a++) s += m * Jv[a * 18 + r] * Jv[a * 18 + c] + Jw[a * 18 + r] * IJ[a * 18 + c];
                d->M[r * 18 + c] += s;
            }
        /* Newton-Euler residuals with nu_dot = 0 */
        double fb[3], fg[3], nb[3], Iwv[3], Ial[3], wIw[3];
        for (int c = 0; c < 3; c++) { fg[c] = -m * in->gravity[c]; fb[c] = m * ac[c] + fg[c]; }
        mat_mul(Iw, k->w[i], Iwv, 3, 3, 1);
        mat_mul(Iw, k->al[i], Ial, 3, 3, 1);
        cross3(k->w[i], Iwv, wIw);
        for (int c = 0; c < 3; c++) nb[c] = Ial[c] + wIw[c];
        for (int r = 0; r < 18; r++) {
            double sh = 0.0, sg = 0.0;
            for (int a = 0; a < 3; a++) {
                sh += Jv[a * 18 + r] * fb[a] + Jw[a * 18 + r] * nb[a];
                sg += Jv[a * 18 + r] * fg[a];
            }
            d->h[r] += sh;
            d->g[r] += sg;
        }
    }
    for (int c = 0; c < 3; c++) { d->com[c] /= msum; d->com_vel[c] /= msum; }
    /* contact frames: computeJac main.cpp:664-690 (linear rows kept, 727-737), getFrameBiasAcc 754-766 */
    for (int f = 0; f < 4; f++) {
        int link = DB_FOOT_LINK[f];
        double Jv[54], Jw[54];
        point_jacobian(k, link, k->p[link], Jv, Jw);
        memcpy(&d->Jac_lin[f * 54], Jv, sizeof(double) * 54);
        for (int c = 0; c < 3; c++) {
            d->Jdqd_lin[f * 3 + c] = k->a[link][c];
            d->foot_pos[f * 3 + c] = k->p[link][c];
            d->foot_vel[f * 3 + c] = k->v[link][c];
        }
        memcpy(&d->foot_R[f * 9], k->R[link], sizeof(double) * 9);
    }
}

/* ------------------------------------------------------------------ update(), main.cpp:572-660 */
void wbc_oracle_update(const wbc_oracle_in* in, wbc_oracle_dyn* d)
{
    kin_t k;
    rigid_body_dynamics(in, d, &k);
    const double mass = DB_TOTAL_MASS;   /* model.getTotalMass(), main.cpp:296 */

    /* computeTransformation, main.cpp:491-568 */
    double xbc[3], xbc_hat[9], xbc_hat_T[9];
    for (int c = 0; c < 3; c++) xbc[c] = d->com[c] - in->base_pos[c];                /* 518 */
    skew3(xbc, xbc_hat);                                                                /* 521-522 */
    mat_T(xbc_hat, xbc_hat_T, 3, 3);
    double X[36];
    memset(X, 0, sizeof(X));
    for (int a = 0; a < 3; a++) {                                                       /* 524-526 */
        X[a * 6 + a] = 1.0; X[(3 + a) * 6 + 3 + a] = 1.0;
        for (int b = 0; b < 3; b++) X[a * 6 + 3 + b] = xbc_hat_T[a * 3 + b];
    }
    double Mb[36], Mbj[72], Mb_Mj[72];
    for (int a = 0; a < 6; a++) {
        for (int b = 0; b < 6; b++) Mb[a * 6 + b] = d->M[a * 18 + b];
        for (int b = 0; b < 12; b++) Mbj[a * 12 + b] = d->M[a * 18 + 6 + b];
    }
    { /* Mb^-1 Mbj  (bdcSvd().solve at 528) */
        double W[36];
        memcpy(W, Mb, sizeof(W));
        memcpy(Mb_Mj, Mbj, sizeof(Mbj));
        gj_solve(W, Mb_Mj, 6, 12);
    }
    double Js[72];
    mat_mul(X, Mb_Mj, Js, 6, 6, 12);                                                    /* 529 */
    memset(d->T, 0, sizeof(d->T));                                                      /* 532-534 */
    for (int a = 0; a < 18; a++) d->T[a * 18 + a] = 1.0;
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) d->T[a * 18 + 3 + b] = xbc_hat_T[a * 3 + b];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 12; b++) d->T[a * 18 + 6 + b] = Js[a * 12 + b];

    double xbc_dot[3], mdr[3], mdr_hat[9], xbc_hat_dot[9], xbc_hat_dot_T[9];
    for (int c = 0; c < 3; c++) { xbc_dot[c] = d->com_vel[c] - in->base_vel[c]; mdr[c] = mass * xbc_dot[c]; } /* 538-539 */
    skew3(mdr, mdr_hat);                                                                /* 541-543 */
    skew3(xbc_dot, xbc_hat_dot);                                                        /* 546-548 */
    mat_T(xbc_hat_dot, xbc_hat_dot_T, 3, 3);
    double dX[36], dMb[36];
    memset(dX, 0, sizeof(dX));
    memset(dMb, 0, sizeof(dMb));
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            dX[a * 6 + 3 + b] = xbc_hat_dot_T[a * 3 + b];                               /* 551-552 */
            dMb[a * 6 + 3 + b] = mdr_hat[b * 3 + a];                                    /* 557: mdr_hat.transpose() */
            dMb[(3 + a) * 6 + b] = mdr_hat[a * 3 + b];                                  /* 558 */
        }
    double inv_dMb1[36], inv_dMb2[36];
    { /* 560: (Mb' \ dMb')' = dMb Mb^-1 */
        double W[36], B[36], Xs[36];
        mat_T(Mb, W, 6, 6);
        mat_T(dMb, B, 6, 6);
        gj_solve(W, B, 6, 6);
        mat_T(B, Xs, 6, 6);
        memcpy(inv_dMb1, Xs, sizeof(Xs));
    }
    { /* 561: -(Mb \ inv_dMb1) */
        double W[36];
        memcpy(W, Mb, sizeof(W));
        memcpy(inv_dMb2, inv_dMb1, sizeof(inv_dMb1));
        gj_solve(W, inv_dMb2, 6, 6);
        for (int i = 0; i < 36; i++) inv_dMb2[i] = -inv_dMb2[i];
    }
    double dJs[72], t0[72], t1[36], t2[72];
    mat_mul(dX, Mb_Mj, t0, 6, 6, 12);                                                   /* 563 */
    mat_mul(X, inv_dMb2, t1, 6, 6, 6);
    mat_mul(t1, Mbj, t2, 6, 6, 12);
    for (int i = 0; i < 72; i++) dJs[i] = t0[i] + t2[i];
    memset(d->T_inv_dot, 0, sizeof(d->T_inv_dot));                                      /* 565-566 */
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) d->T_inv_dot[a * 18 + 3 + b] = xbc_hat_dot[a * 3 + b];
        for (int b = 0; b < 12; b++) d->T_inv_dot[a * 18 + 6 + b] = -dJs[a * 12 + b];
    }

    /* CoM-coordinate quantities, main.cpp:645-659 */
    double dqv[18];                                                                     /* dq = [CoM_vel; w_base; dq_j], 602, 608 */
    for (int c = 0; c < 3; c++) { dqv[c] = d->com_vel[c]; dqv[3 + c] = in->base_vel[3 + c]; }
    for (int c = 0; c < 12; c++) dqv[6 + c] = in->dq[c];
    double Tinv[324], TinvT[324], Tt[324], A1[324], A2[324], v1[18], v2[18], v3[18];
    mat_inv(d->T, Tinv, 18);
    mat_T(d->T, Tt, 18, 18);
    mat_inv(Tt, TinvT, 18);
    mat_mul(TinvT, d->M, A1, 18, 18, 18);                                               /* 645 */
    mat_mul(A1, Tinv, d->Mcom, 18, 18, 18);
    mat_mul(TinvT, d->h, v1, 18, 18, 1);                                                /* 648 */
    mat_mul(A1, d->T_inv_dot, A2, 18, 18, 18);
    mat_mul(A2, dqv, v2, 18, 18, 1);
    for (int c = 0; c < 18; c++) d->hcom[c] = v1[c] + v2[c];
    mat_mul(TinvT, d->g, d->gcom, 18, 18, 1);                                           /* 651 */
    mat_mul(d->Jac_lin, Tinv, d->Jcom_lin, 12, 18, 18);                                 /* 654-655 (linear rows) */
    double JT[216];
    mat_mul(d->Jac_lin, d->T_inv_dot, JT, 12, 18, 18);                                  /* 658-659 */
    mat_mul(JT, dqv, v3, 12, 18, 1);
    for (int c = 0; c < 12; c++) d->Jdqdcom_lin[c] = d->Jdqd_lin[c] + v3[c];
}

/* ------------------------------------------------------------------ Fgrf, main.cpp:1022-1026, 1218, 1765 */
void wbc_oracle_fgrf(const wbc_oracle_in* in, const wbc_oracle_dyn* d, double* Fgrf)
{
    for (int f = 0; f < 4; f++) {
        /* stacked feet: 0=BR 1=BL 2=FL 3=FR */
        int swing = (in->mode == WBC_MODE_SWING_BR_FL && (f == 0 || f == 2)) ||
                    (in->mode == WBC_MODE_SWING_BL_FR && (f == 1 || f == 3));
        if (swing) { Fgrf[f * 3] = Fgrf[f * 3 + 1] = Fgrf[f * 3 + 2] = 0.0; continue; }
        mat_mul(&d->foot_R[f * 9], &in->foot_force[f * 3], &Fgrf[f * 3], 3, 3, 1);
    }
}

/* ------------------------------------------------------------------ estimate(), main.cpp:692-725 */
static void wbc_oracle_estimate_full(const wbc_oracle_params* p, const double* Mc, const double* qd, const double* fc,
                                     const double* yd_prev, const double* yw_prev, double* w, double* yd, double* yw);
static void wbc_oracle_estimate_sem(const wbc_oracle_params* p, const double* Mc, const double* qd, const double* fc,
                                    const double* yd_prev, const double* yw_prev, const double* yg_prev, double* w, double* yd,
                                    double* yw, double* yg);

void wbc_oracle_estimate(const wbc_oracle_params* p, const wbc_oracle_in* in, const wbc_oracle_dyn* d,
                         const double* Fgrf, double* w, double* yd, double* yw)
{
    double Mc[36], qd[6], fc[6];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) Mc[a * 6 + b] = d->Mcom[a * 18 + b];               /* 694 */
    for (int c = 0; c < 3; c++) { qd[c] = d->com_vel[c]; qd[3 + c] = in->base_vel[3 + c]; }   /* 695, 602 */
    for (int c = 0; c < 6; c++) {                                                       /* 696-698 */
        double s = 0.0;
        for (int r = 0; r < 12; r++) s += d->Jcom_lin[r * 18 + c] * Fgrf[r];
        fc[c] = s;
    }
    wbc_oracle_estimate_full(p, Mc, qd, fc, in->yd_prev, in->yw_prev, w, yd, yw);
}

/* Same inputs, the forms of SURVEY.md 8f-3 (obs_order == 2 or obs_form == 1). */
void wbc_oracle_estimate_ext(const wbc_oracle_params* p, const wbc_oracle_in* in, const wbc_oracle_dyn* d,
                             const double* Fgrf, double* w, double* yd, double* yw, double* yg)
{
    double Mc[36], qd[6], fc[6];
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) Mc[a * 6 + b] = d->Mcom[a * 18 + b];
    for (int c = 0; c < 3; c++) { qd[c] = d->com_vel[c]; qd[3 + c] = in->base_vel[3 + c]; }
    for (int c = 0; c < 6; c++) {
        double s = 0.0;
        for (int r = 0; r < 12; r++) s += d->Jcom_lin[r * 18 + c] * Fgrf[r];
        fc[c] = s;
    }
    wbc_oracle_estimate_sem(p, Mc, qd, fc, in->yd_prev, in->yw_prev, in->yg_prev, w, yd, yw, yg);
}

/* ESTIMATOR_SEM::estimate (estimator_sem.cpp:24-61) and its second-order continuation.
 *   order 1, form 1: the literal lines 53-59: yd = yd_prev + d T; w = k0 (rho - yw_prev - yd); yw = yw_prev + w T
 *                    (its `Ctq` term comes from QUADRUPED::getCtq, declared at main.cpp:155 and defined nowhere: taken as zero,
 *                    which is exact for the centroidal momentum rho = Mcom CoM_vel; its T = 0.001 is p->obs_dt);
 *   order 2:         the r = 2 member of the family the coefficient vector {10, 1} and the ygamma state belong to,
 *                        gamma1 = k1 (rho - int(w + d)),  w = k2 int(gamma1 - w),   w / w_true = k1 k2 / (s^2 + k2 s + k1 k2),
 *                    form 0 discretised as main.cpp:716-719 does the first order (backward Euler, the new w on both sides),
 *                    form 1 as estimator_sem.cpp:55-59 does (forward Euler). */
static void wbc_oracle_estimate_sem(const wbc_oracle_params* p, const double* Mc, const double* qd, const double* fc,
                                    const double* yd_prev, const double* yw_prev, const double* yg_prev, double* w, double* yd,
                                    double* yw, double* yg)
{
    const double mass = DB_TOTAL_MASS, T = p->obs_dt, k1 = p->obs_gain, k2 = p->obs_gain2;
    for (int a = 0; a < 6; a++) {
        double rho = 0.0;
        for (int b = 0; b < 6; b++) rho += Mc[a * 6 + b] * qd[b];
        const double dd = -mass * ((a == 2) ? p->g_acc : 0.0) + fc[a];
        yd[a] = yd_prev[a] + dd * T;
        const double e = rho - yw_prev[a] - yd[a];
        if (p->obs_order != 2) {
            w[a] = (p->obs_form == 1) ? k1 * e : (1.0 / (1.0 + k1 * T)) * k1 * e;
            yg[a] = yg_prev[a];
        } else if (p->obs_form == 1) {
            yg[a] = yg_prev[a] + T * (k1 * e - k2 * yg_prev[a]);
            w[a] = k2 * yg[a];
        } else {
            w[a] = k2 * (yg_prev[a] + T * k1 * e) / (1.0 + k2 * T + k1 * k2 * T * T);
            yg[a] = yg_prev[a] + T * (k1 * (e - w[a] * T) - w[a]);
        }
        yw[a] = yw_prev[a] + w[a] * T;
    }
}

/* The recurrence itself: Mc 6x6 row-major, qd = CoM_vel (6), fc = J' Fgrf (6). */
static void wbc_oracle_estimate_full(const wbc_oracle_params* p, const double* Mc, const double* qd, const double* fc,
                                     const double* yd_prev, const double* yw_prev, double* w, double* yd, double* yw)
{
    const double mass = DB_TOTAL_MASS;                                                  /* 700 */
    const double T = p->obs_dt;                                                         /* 715 */
    const double k0 = p->obs_gain;                                                      /* 708, 712 */
    const double mgain = (1.0 / (1.0 + k0 * T)) * k0;                                   /* 716: (I + k0 T)^-1 k0 */
    for (int a = 0; a < 6; a++) {
        double rho = 0.0;                                                               /* 699, 705 */
        for (int b = 0; b < 6; b++) rho += Mc[a * 6 + b] * qd[b];
        double g_acc = (a == 2) ? p->g_acc : 0.0;                                       /* 701-702 */
        double dd = -mass * g_acc + fc[a];                                              /* 704, 706 */
        yd[a] = yd_prev[a] + dd * T;                                                    /* 717 */
        w[a] = mgain * (rho - yw_prev[a] - yd[a]);                                      /* 718 */
        yw[a] = yw_prev[a] + w[a] * T;                                                  /* 719 */
    }
}

/* ------------------------------------------------------------------ QP assembly */
static void friction_rows(const wbc_oracle_params* p, const wbc_oracle_in* in, int foot, double* cfr /*5x3*/)
{ /* main.cpp:1062-1078 (1266-1288): rows (-mu n + t1)', (-mu n + t2)', -(mu n + t1)', -(mu n + t2)', -n' */
    double n[3] = {0, 0, 1}, t1[3] = {1, 0, 0}, t2[3] = {0, 1, 0}, mu = p->mu;
    if (in->has_terrain) {
        const double* tr = &in->terrain[foot * 10];
        for (int c = 0; c < 3; c++) { n[c] = tr[c]; t1[c] = tr[3 + c]; t2[c] = tr[6 + c]; }
        mu = tr[9];
    }
    for (int c = 0; c < 3; c++) {
        cfr[0 * 3 + c] = -mu * n[c] + t1[c];
        cfr[1 * 3 + c] = -mu * n[c] + t2[c];
        cfr[2 * 3 + c] = -(mu * n[c] + t1[c]);
        cfr[3 * 3 + c] = -(mu * n[c] + t2[c]);
        cfr[4 * 3 + c] = -n[c];
    }
}

void wbc_oracle_assemble(const wbc_oracle_params* p, const wbc_oracle_in* in, const wbc_oracle_dyn* d,
                         const double* w, wbc_oracle_qp* qp)
{
    const double mass = DB_TOTAL_MASS;
    const double* Mcom = d->Mcom;
    const double* J = d->Jcom_lin;      /* 12 x 18 */
    memset(qp->Q, 0, sizeof(qp->Q));
    memset(qp->c, 0, sizeof(qp->c));
    memset(qp->L, 0, sizeof(qp->L));

    /* Wcom_des, main.cpp:1012-1032 (1200-1223) */
    double com6[6], comv6[6], deltax[6], deltav[6], rot[3];
    for (int c = 0; c < 3; c++) { com6[c] = d->com[c]; com6[3 + c] = in->rpy[c]; comv6[c] = d->com_vel[c]; comv6[3 + c] = in->base_vel[3 + c]; }
    for (int c = 0; c < 6; c++) { deltax[c] = in->com_des_pos[c] - com6[c]; deltav[c] = in->com_des_vel[c] - comv6[c]; }
    mat_mul(in->base_R, &deltax[3], rot, 3, 3, 1);                                      /* 1013 */
    for (int c = 0; c < 3; c++) deltax[3 + c] = rot[c];
    for (int a = 0; a < 6; a++) {                                                       /* 1032 */
        double ma = 0.0;
        for (int b = 0; b < 6; b++) ma += Mcom[a * 18 + b] * in->com_des_acc[b];
        double g_acc = (a == 2) ? p->g_acc : 0.0;
        qp->Wcom_des[a] = p->kcom * deltax[a] + p->dcom * deltav[a] + mass * g_acc + ma - w[a];
    }

    const double dt = p->joint_dt;
    /* q (joints) = jointPos, dq (joints) = jointVel; qmin/qmax main.cpp:612-613 */
    static const double qmin[12] = {-1.75, -1.75, -1.75, -1.75, -1.58, -2.62, -3.15, -0.02, -1.58, -2.62, -3.15, -0.02};
    static const double qmax[12] = {1.75, 1.75, 1.75, 1.75, 3.15, 0.02, 1.58, 2.62, 3.15, 0.02, 1.58, 2.62};
    double ddqmin[12], ddqmax[12];
    for (int j = 0; j < 12; j++) {                                                      /* 1103-1104 */
        ddqmin[j] = (2 / pow(dt, 2)) * (qmin[j] - in->q[j] - dt * in->dq[j]);
        ddqmax[j] = (2 / pow(dt, 2)) * (qmax[j] - in->q[j] - dt * in->dq[j]);
    }

    if (in->mode == WBC_MODE_STANCE) {
        qp->nrows = 86; qp->neq = 18;
        double* L = qp->L;
        /* Q = T_s' Q1 T_s + I, T_s = Jstcom' [0 | I12]   (994-1001) */
        for (int i = 0; i < 30; i++) qp->Q[i * 30 + i] = 1.0;
        for (int a = 0; a < 12; a++)
            for (int b = 0; b < 12; b++) {
                double s = 0.0;
                for (int k = 0; k < 6; k++) s += J[a * 18 + k] * p->q1_weight * J[b * 18 + k];
                qp->Q[(18 + a) * 30 + 18 + b] += s;
            }
        /* c = -T_s' Q1' Wcom_des   (1033) */
        for (int a = 0; a < 12; a++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s += J[a * 18 + k] * p->q1_weight * qp->Wcom_des[k];
            qp->c[18 + a] = -s;
        }
        /* equalities (1039-1048) */
        for (int a = 0; a < 6; a++) {
            for (int b = 0; b < 6; b++) L[a * 31 + b] = Mcom[a * 18 + b];
            for (int b = 0; b < 12; b++) L[a * 31 + 18 + b] = -J[b * 18 + a];
            L[a * 31 + 30] = -d->hcom[a];
        }
        for (int a = 0; a < 12; a++) {
            for (int b = 0; b < 18; b++) L[(6 + a) * 31 + b] = J[a * 18 + b];
            L[(6 + a) * 31 + 30] = -d->Jdqdcom_lin[a];
        }
        /* inequalities: D rows start at L row 18 (1051-1107) */
        for (int f = 0; f < 4; f++) {
            double cfr[15];
            friction_rows(p, in, f, cfr);
            for (int r = 0; r < 5; r++)
                for (int c = 0; c < 3; c++) L[(18 + 5 * f + r) * 31 + 18 + 3 * f + c] = cfr[r * 3 + c];
        }
        for (int a = 0; a < 12; a++) {
            for (int b = 0; b < 12; b++) {
                L[(18 + 20 + a) * 31 + 6 + b] = Mcom[(6 + a) * 18 + 6 + b];
                L[(18 + 20 + a) * 31 + 18 + b] = -J[b * 18 + 6 + a];
                L[(18 + 32 + a) * 31 + 6 + b] = -Mcom[(6 + a) * 18 + 6 + b];
                L[(18 + 32 + a) * 31 + 18 + b] = J[b * 18 + 6 + a];
            }
            L[(18 + 44 + a) * 31 + 6 + a] = 1.0;
            L[(18 + 56 + a) * 31 + 6 + a] = -1.0;
            L[(18 + 20 + a) * 31 + 30] = p->tau_max - d->hcom[6 + a];
            L[(18 + 32 + a) * 31 + 30] = -(-p->tau_max - d->hcom[6 + a]);
            L[(18 + 44 + a) * 31 + 30] = ddqmax[a];
            L[(18 + 56 + a) * 31 + 30] = -ddqmin[a];
        }
    } else {
        qp->nrows = 82; qp->neq = 12;
        double* L = qp->L;
        int stl1, stl2, swl1, swl2, stf1, stf2;
        if (in->mode == WBC_MODE_SWING_BR_FL) { swl1 = 0; swl2 = 6; stl1 = 3; stl2 = 9; }   /* 1163-1167 */
        else { swl1 = 3; swl2 = 9; stl1 = 0; stl2 = 6; }                                    /* 1710-1714 */
        stf1 = stl1 / 3; stf2 = stl2 / 3;
        double Jst[6 * 18], Jsw[6 * 18];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 18; c++) {
                Jst[r * 18 + c] = J[(stl1 + r) * 18 + c]; Jst[(3 + r) * 18 + c] = J[(stl2 + r) * 18 + c];
                Jsw[r * 18 + c] = J[(swl1 + r) * 18 + c]; Jsw[(3 + r) * 18 + c] = J[(swl2 + r) * 18 + c];
            }
        /* Q = T_s' Q1 T_s + R, R = I with 1e8 on the slack block (1176-1189) */
        for (int i = 0; i < 24; i++) qp->Q[i * 30 + i] = 1.0;
        for (int i = 24; i < 30; i++) qp->Q[i * 30 + i] = p->slack_weight;
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) {
                double s = 0.0;
                for (int k = 0; k < 6; k++) s += Jst[a * 18 + k] * p->q1_weight * Jst[b * 18 + k];
                qp->Q[(18 + a) * 30 + 18 + b] += s;
            }
        for (int a = 0; a < 6; a++) {                                                   /* 1224 */
            double s = 0.0;
            for (int k = 0; k < 6; k++) s += Jst[a * 18 + k] * p->q1_weight * qp->Wcom_des[k];
            qp->c[18 + a] = -s;
        }
        /* equalities (1231-1241); rhs left at zero in the reference (quirk E2) */
        for (int a = 0; a < 6; a++) {
            for (int b = 0; b < 6; b++) L[a * 31 + b] = Mcom[a * 18 + b];
            for (int b = 0; b < 6; b++) L[a * 31 + 18 + b] = -Jst[b * 18 + a];
            for (int b = 0; b < 18; b++) L[(6 + a) * 31 + b] = Jst[a * 18 + b];
            if (p->fix_swing_rhs) {
                L[a * 31 + 30] = -d->hcom[a];
                L[(6 + a) * 31 + 30] = -d->Jdqdcom_lin[(a < 3 ? stl1 : stl2 - 3) + a];
            }
        }
        /* inequalities: D rows start at L row 12 (1244-1379) */
        int stf[2] = {stf1, stf2};
        for (int f = 0; f < 2; f++) {
            double cfr[15];
            friction_rows(p, in, stf[f], cfr);
            for (int r = 0; r < 5; r++)
                for (int c = 0; c < 3; c++) L[(12 + 5 * f + r) * 31 + 18 + 3 * f + c] = cfr[r * 3 + c];
        }
        for (int a = 0; a < 12; a++) {
            for (int b = 0; b < 12; b++) {
                L[(12 + 10 + a) * 31 + 6 + b] = Mcom[(6 + a) * 18 + 6 + b];
                L[(12 + 22 + a) * 31 + 6 + b] = -Mcom[(6 + a) * 18 + 6 + b];
            }
            for (int b = 0; b < 6; b++) {
                L[(12 + 10 + a) * 31 + 18 + b] = -Jst[b * 18 + 6 + a];
                L[(12 + 22 + a) * 31 + 18 + b] = Jst[b * 18 + 6 + a];
            }
            L[(12 + 46 + a) * 31 + 6 + a] = 1.0;
            L[(12 + 58 + a) * 31 + 6 + a] = -1.0;
            L[(12 + 10 + a) * 31 + 30] = p->tau_max - d->hcom[6 + a];
            L[(12 + 22 + a) * 31 + 30] = -(-p->tau_max - d->hcom[6 + a]);
            L[(12 + 46 + a) * 31 + 30] = ddqmax[a];
            L[(12 + 58 + a) * 31 + 30] = -ddqmin[a];
        }
        /* swing-foot tracking with slack (1251-1262, 1327-1379) */
        int swf[2] = {swl1 / 3, swl2 / 3};
        for (int a = 0; a < 6; a++) {
            int f = swf[a / 3], cidx = a % 3;
            double Jdqdsw = d->Jdqdcom_lin[(a < 3 ? swl1 : swl2) + cidx];
            double posdelta = in->sw_des_pos[a] - d->foot_pos[f * 3 + cidx];
            double veldelta = in->sw_des_vel[a] - d->foot_vel[f * 3 + cidx];
            double vdotswdes = in->sw_des_acc[a] + p->kd_sw * veldelta + p->kp_sw * posdelta;   /* 1375 */
            for (int b = 0; b < 18; b++) {
                L[(12 + 34 + a) * 31 + b] = Jsw[a * 18 + b];
                L[(12 + 40 + a) * 31 + b] = -Jsw[a * 18 + b];
            }
            L[(12 + 34 + a) * 31 + 24 + a] = -1.0;
            L[(12 + 40 + a) * 31 + 24 + a] = -1.0;
            L[(12 + 34 + a) * 31 + 30] = vdotswdes - Jdqdsw;                            /* 1378 */
            L[(12 + 40 + a) * 31 + 30] = -vdotswdes + Jdqdsw;                           /* 1379 */
        }
    }
}

/* ------------------------------------------------------------------ torque map, main.cpp:1126, 1396 */
void wbc_oracle_torque(const wbc_oracle_in* in, const wbc_oracle_dyn* d, const double* x, double* tau)
{
    const double* J = d->Jcom_lin;
    int strow[4], nst;
    if (in->mode == WBC_MODE_STANCE) { nst = 4; strow[0] = 0; strow[1] = 3; strow[2] = 6; strow[3] = 9; }
    else if (in->mode == WBC_MODE_SWING_BR_FL) { nst = 2; strow[0] = 3; strow[1] = 9; }
    else { nst = 2; strow[0] = 0; strow[1] = 6; }
    for (int a = 0; a < 12; a++) {
        double s = 0.0;
        for (int b = 0; b < 12; b++) s += d->Mcom[(6 + a) * 18 + 6 + b] * x[6 + b];
        s += d->hcom[6 + a];
        double jf = 0.0;
        for (int f = 0; f < nst; f++)
            for (int c = 0; c < 3; c++) jf += J[(strow[f] + c) * 18 + 6 + a] * x[18 + 3 * f + c];
        tau[a] = s - jf;
    }
}

/* ------------------------------------------------------------------ one full cycle */
int wbc_oracle_cycle(const wbc_oracle_params* p, const wbc_oracle_in* in, wbc_qp_fn solve,
                     wbc_oracle_out* out, wbc_oracle_dyn* dyn_opt, wbc_oracle_qp* qp_opt)
{
    wbc_oracle_dyn* d = dyn_opt ? dyn_opt : (wbc_oracle_dyn*)malloc(sizeof(wbc_oracle_dyn));
    wbc_oracle_qp* qp = qp_opt ? qp_opt : (wbc_oracle_qp*)malloc(sizeof(wbc_oracle_qp));
    memset(out, 0, sizeof(*out));
    wbc_oracle_update(in, d);
    wbc_oracle_fgrf(in, d, qp->Fgrf);
    if (p->observer_enabled) {
        /* estimate() placed where the reference's commented-out call sits (main.cpp:1029, 1220, 1569, 1767) */
        if (p->obs_order == 2 || p->obs_form == 1) wbc_oracle_estimate_ext(p, in, d, qp->Fgrf, out->w, out->yd, out->yw, out->yg);
        else {
            wbc_oracle_estimate(p, in, d, qp->Fgrf, out->w, out->yd, out->yw);
            for (int c = 0; c < 6; c++) out->yg[c] = in->yg_prev[c];
        }
    } else {
        for (int c = 0; c < 6; c++) { out->w[c] = 0.0; out->yd[c] = in->yd_prev[c]; out->yw[c] = in->yw_prev[c]; out->yg[c] = in->yg_prev[c]; }
    }
    wbc_oracle_assemble(p, in, d, out->w, qp);
    int nchol = 0, term = 0;
    int rc = solve(qp->Q, qp->c, qp->L, qp->nrows, qp->neq, out->x, &nchol, &term);
    out->status = rc;
    out->ncholesky = nchol;
    if (rc == 0) {
        wbc_oracle_torque(in, d, out->x, out->tau);
        double obj = 0.0;
        for (int i = 0; i < 30; i++) {
            double s = 0.0;
            for (int j = 0; j < 30; j++) s += qp->Q[i * 30 + j] * out->x[j];
            obj += 0.5 * out->x[i] * s + qp->c[i] * out->x[i];
        }
        out->qp_obj = obj;
    }
    if (!dyn_opt) free(d);
    if (!qp_opt) free(qp);
    return rc;
}

/* ------------------------------------------------------------------ synthetic plant of BASELINE config 5
 * Not reference code: the reference's plant is Gazebo + the ModelPush plugin (force_plugin/src/force_plugin.cpp:124-491).
 * SURVEY.md 8d row 5 defines the stand-in: a CoM momentum integrator with locked joints,
 *     rho_{k+1} = rho_k + T (-m g_acc e3 + fc + w_true),   rho = Mcom[0:6,0:6] CoM_vel,  fc = J' Fgrf
 * i.e. exactly the balance the observer (main.cpp:692-725) inverts, so that w -> w_true.  The new CoM_vel is
 * mapped back to the base twist of a rigid body: omega = CoM_vel[3:6], v_base = CoM_vel[0:3] - omega x (com - base),
 * and the base position advances by T v_base (orientation and joints stay fixed).  With x != NULL the loop is closed:
 * Fgrf = the commanded forces x[18:30], and foot_force_out = R_foot' f* is what the sensors report next cycle. */
void wbc_oracle_plant_step(const wbc_oracle_params* p, const wbc_oracle_in* in, const double* push, const double* x,
                           double* base_pos_out, double* base_vel_out, double* foot_force_out)
{
    wbc_oracle_dyn* d = (wbc_oracle_dyn*)malloc(sizeof(wbc_oracle_dyn));
    double Fgrf[12], Mc[36], qd[6], rhs[6];
    wbc_oracle_update(in, d);
    if (x) {
        /* closed loop: the ground reacts with the commanded forces (stance layout x[18:30], main.cpp:1126) */
        for (int r = 0; r < 12; r++) Fgrf[r] = x[18 + r];
        for (int f = 0; f < 4; f++) {
            double Rt[9];
            mat_T(&d->foot_R[f * 9], Rt, 3, 3);
            mat_mul(Rt, &Fgrf[f * 3], &foot_force_out[f * 3], 3, 3, 1);
        }
    } else {
        wbc_oracle_fgrf(in, d, Fgrf);
    }
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) Mc[a * 6 + b] = d->Mcom[a * 18 + b];
    for (int c = 0; c < 3; c++) { qd[c] = d->com_vel[c]; qd[3 + c] = in->base_vel[3 + c]; }
    for (int a = 0; a < 6; a++) {
        double rho = 0.0, fc = 0.0;
        for (int b = 0; b < 6; b++) rho += Mc[a * 6 + b] * qd[b];
        for (int r = 0; r < 12; r++) fc += d->Jcom_lin[r * 18 + a] * Fgrf[r];
        const double dd = -DB_TOTAL_MASS * ((a == 2) ? p->g_acc : 0.0) + fc;
        rhs[a] = rho + p->obs_dt * (dd + push[a]);
    }
    gj_solve(Mc, rhs, 6, 1);                                  /* CoM_vel_{k+1} */
    double xbc[3], wx[3];
    for (int c = 0; c < 3; c++) xbc[c] = d->com[c] - in->base_pos[c];
    cross3(&rhs[3], xbc, wx);
    for (int c = 0; c < 3; c++) {
        base_vel_out[c] = rhs[c] - wx[c];
        base_vel_out[3 + c] = rhs[3 + c];
        base_pos_out[c] = in->base_pos[c] + p->obs_dt * base_vel_out[c];
    }
    free(d);
}

/* ------------------------------------------------------------------ forward dynamics with hard contacts
 * Stands in for Gazebo + ModelPush (force_plugin/src/force_plugin.cpp:124-491), which cannot run here; PARITY UNPINNED (the
 * reference for this stage is a physics engine).  One control period (p->obs_dt) of the articulated robot under the commanded
 * joint torques (held over the period, as the 400 Hz loop of main.cpp:861 holds them over Gazebo's 1 ms steps), a world wrench
 * `push` at the CoM (what ModelPush applies to a link, reduced to the CoM) and rigid, bilateral point contacts at the stance
 * feet of in->mode:
 *       [ M  -Js' ] [ nu_dot ]   [ S'tau - h + push_gen          ]
 *       [ Js   0  ] [   f    ] = [ -Jdqd_s - gamma Js nu          ]       (dense, literal, Gauss-Jordan)
 * nu = [v_base; omega; dq] (MIXED), M / h / Js / Jdqd from wbc_oracle_update, gamma = velocity-level constraint stabilisation.
 * `nsub` semi-implicit Euler substeps: nu += dt nu_dot, then q += dt dq, p += dt v, R <- exp([omega dt]) R with the NEW
 * velocities.  Outputs the next state, the contact forces of the last substep in the sensor (foot link) frame -- what the
 * contact sensors report to the next cycle (main.cpp:794-834) -- and the largest |Js nu_dot + Jdqd_s + gamma Js nu| (should be
 * rounding) plus the smallest normal force (unilaterality is NOT enforced: a negative value means the foot would have lifted). */
static void rot_to_rpy(const double* R, double* rpy)
{   /* R = Rz(yaw) Ry(pitch) Rx(roll) */
    rpy[0] = atan2(R[7], R[8]);
    rpy[1] = atan2(-R[6], sqrt(R[7] * R[7] + R[8] * R[8]));
    rpy[2] = atan2(R[3], R[0]);
}
void wbc_oracle_fdyn_step(const wbc_oracle_params* p, const wbc_oracle_in* in0, const double* tau, const double* push, int nsub,
                          double gamma, wbc_oracle_in* out, double* foot_force_out, double* diag /* [2] */)
{
    wbc_oracle_dyn* d = (wbc_oracle_dyn*)malloc(sizeof(wbc_oracle_dyn));
    wbc_oracle_in cur = *in0;
    const double dt = p->obs_dt / (double)nsub;
    int stance[4], ns = 0;
    for (int f = 0; f < 4; f++) {
        const int swing = (in0->mode == 1 && (f == 0 || f == 2)) || (in0->mode == 2 && (f == 1 || f == 3));
        stance[f] = !swing;
        ns += stance[f];
    }
    const int nc = 3 * ns, nk = 18 + nc;
    double* K = (double*)malloc(sizeof(double) * nk * nk);
    double* rhs = (double*)malloc(sizeof(double) * nk);
    double worst = 0.0, fzmin = 1e300;
    for (int c = 0; c < 12; c++) foot_force_out[c] = 0.0;
    for (int it = 0; it < nsub; it++) {
        wbc_oracle_update(&cur, d);
        double nu[18];
        for (int c = 0; c < 6; c++) nu[c] = cur.base_vel[c];
        for (int c = 0; c < 12; c++) nu[6 + c] = cur.dq[c];
        double xbc[3], tq[3];
        for (int c = 0; c < 3; c++) xbc[c] = d->com[c] - cur.base_pos[c];
        cross3(xbc, push, tq);                                  /* wrench at the CoM -> about the base origin */
        memset(K, 0, sizeof(double) * nk * nk);
        for (int a = 0; a < 18; a++) {
            for (int b = 0; b < 18; b++) K[a * nk + b] = d->M[a * 18 + b];
            rhs[a] = -d->h[a];
        }
        for (int c = 0; c < 3; c++) { rhs[c] += push[c]; rhs[3 + c] += push[3 + c] + tq[c]; }
        for (int c = 0; c < 12; c++) rhs[6 + c] += tau[c];
        int r = 0;
        double cres[12];
        for (int f = 0; f < 4; f++) {
            if (!stance[f]) continue;
            for (int a = 0; a < 3; a++, r++) {
                double jn = 0.0;
                for (int b = 0; b < 18; b++) {
                    const double j = d->Jac_lin[(3 * f + a) * 18 + b];
                    K[(18 + r) * nk + b] = j;
                    K[b * nk + 18 + r] = -j;
                    jn += j * nu[b];
                }
                rhs[18 + r] = -d->Jdqd_lin[3 * f + a] - gamma * jn;
                cres[r] = rhs[18 + r];
            }
        }
        gj_solve(K, rhs, nk, 1);                                /* rhs <- [nu_dot; f] */
        /* residual of the contact constraint with the computed acceleration */
        r = 0;
        for (int f = 0; f < 4; f++) {
            if (!stance[f]) continue;
            for (int a = 0; a < 3; a++, r++) {
                double ja = 0.0;
                for (int b = 0; b < 18; b++) ja += d->Jac_lin[(3 * f + a) * 18 + b] * rhs[b];
                worst = fmax(worst, fabs(ja - cres[r]));
            }
        }
        /* contact forces in the sensor frame */
        r = 0;
        for (int f = 0; f < 4; f++) {
            if (!stance[f]) { for (int a = 0; a < 3; a++) foot_force_out[3 * f + a] = 0.0; continue; }
            double Rt[9];
            mat_T(&d->foot_R[f * 9], Rt, 3, 3);
            mat_mul(Rt, &rhs[18 + r], &foot_force_out[3 * f], 3, 3, 1);
            fzmin = fmin(fzmin, rhs[18 + r + 2]);
            r += 3;
        }
        /* semi-implicit Euler */
        for (int c = 0; c < 18; c++) nu[c] += dt * rhs[c];
        for (int c = 0; c < 6; c++) cur.base_vel[c] = nu[c];
        for (int c = 0; c < 12; c++) { cur.dq[c] = nu[6 + c]; cur.q[c] += dt * nu[6 + c]; }
        for (int c = 0; c < 3; c++) cur.base_pos[c] += dt * nu[c];
        {
            const double wn = sqrt(nu[3] * nu[3] + nu[4] * nu[4] + nu[5] * nu[5]);
            if (wn > 0.0) {
                double ax[3] = {nu[3] / wn, nu[4] / wn, nu[5] / wn}, Rw[9], Rn[9];
                rot_axis_angle(ax, wn * dt, Rw);
                mat_mul(Rw, cur.base_R, Rn, 3, 3, 3);
                memcpy(cur.base_R, Rn, sizeof(Rn));
            }
            rot_to_rpy(cur.base_R, cur.rpy);
        }
    }
    *out = cur;
    if (diag) { diag[0] = worst; diag[1] = (ns > 0) ? fzmin : 0.0; }
    free(K); free(rhs); free(d);
}

/* ------------------------------------------------------------------ threaded batch (CPU baseline) */
typedef struct {
    const wbc_oracle_params* p; const wbc_oracle_in* in; wbc_oracle_out* out; wbc_qp_fn solve; int lo, hi;
} batch_arg;
static void* batch_worker(void* a_)
{
    batch_arg* a = (batch_arg*)a_;
    wbc_oracle_dyn d;
    wbc_oracle_qp qp;
    for (int i = a->lo; i < a->hi; i++) wbc_oracle_cycle(a->p, &a->in[i], a->solve, &a->out[i], &d, &qp);
    return NULL;
}
double wbc_oracle_batch(const wbc_oracle_params* p, const wbc_oracle_in* in, int n, int nthreads,
                        wbc_qp_fn solve, wbc_oracle_out* out)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    batch_arg* args = (batch_arg*)malloc(sizeof(batch_arg) * nthreads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        args[t].p = p; args[t].in = in; args[t].out = out; args[t].solve = solve;
        args[t].lo = (int)((long long)n * t / nthreads);
        args[t].hi = (int)((long long)n * (t + 1) / nthreads);
        pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    free(args);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}


/* ------------------------------------------------------------------------------------------------
 * towr spline sampling (the planner side of main.cpp:1004-1010, 1333-1368), literal:
 *   Spline::GetSegmentID   spline.cc:48-66   first i with cumulative duration >= t - 1e-10
 *   Spline::GetLocalTime   spline.cc:68-79   subtract the durations of the earlier polynomials one by one
 *   CubicHermitePolynomial::UpdateCoeff  polynomial.cc:98-104
 *   Polynomial::GetPoint / GetDerivativeWrtCoeff  polynomial.cc:50-76  (std::pow, accumulation from zero in the order A..D) */
int wbc_oracle_spline_point(int nseg, const double* durations, const double* nodes, double t_global, double* p, double* v, double* a)
{
    const double eps = 1e-10;
    if (t_global < 0.0) return -1;                 /* assert(t_global >= 0.0) */
    double t = 0.0;
    int id = -1;
    for (int i = 0; i < nseg; i++) {
        t += durations[i];
        if (t >= t_global - eps) { id = i; break; }
    }
    if (id < 0) return -1;                         /* assert(false): this should never be reached */
    double t_local = t_global;
    for (int i = 0; i < id; i++) t_local -= durations[i];
    const double T = durations[id];
    const double* n0 = nodes + 6 * id;
    const double* n1 = nodes + 6 * (id + 1);
    for (int d = 0; d < 3; d++) {
        double coeff[4];
        coeff[0] = n0[d];
        coeff[1] = n0[3 + d];
        coeff[2] = -(3 * (n0[d] - n1[d]) + T * (2 * n0[3 + d] + n1[3 + d])) / pow(T, 2);
        coeff[3] = (2 * (n0[d] - n1[d]) + T * (n0[3 + d] + n1[3 + d])) / pow(T, 3);
        double op = 0.0, ov = 0.0, oa = 0.0;
        for (int c = 0; c < 4; c++) {
            op += pow(t_local, c) * coeff[c];
            ov += (c >= 1 ? c * pow(t_local, c - 1) : 0.0) * coeff[c];
            oa += (c >= 2 ? c * (c - 1) * pow(t_local, c - 2) : 0.0) * coeff[c];
        }
        p[d] = op; v[d] = ov; a[d] = oa;
    }
    return id;
}
