/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the WBC control-cycle hot path.
 *
 * Plain-C restatement of the reference controller's per-cycle arithmetic
 * (dogbot_controller/src/client/main.cpp:491-766, 984-1127, 1163-1397) for ONE DogBot
 * instance.  Nothing under wbc_quadruped_dob_b200/ may include, link or call this.
 *
 * PARITY STATUS
 *   QP solve      : pinned -- the reference's own vendored ALGLIB is compiled into
 *                   oracle/_ref/libref_alglib_qp.so and called through `wbc_qp_fn`.
 *   rigid-body    : PARITY UNPINNED -- iDynTree is not vendored in the reference and is not
 *   dynamics        available here; the reference has no tests or golden vectors for this
 *                   stage.  This file restates iDynTree's MIXED-representation
 *                   semantics with a Jacobian-projection formulation
 *                   that is independent of the CUDA kernels' recursive formulation; it is
 *                   cross-checked by physics identities in tests/test_oracle_dynamics.py.
 */
#ifndef WBC_ORACLE_H
#define WBC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* contact modes: which feet swing */
#define WBC_MODE_STANCE 0        /* main.cpp:979-1148, 1516-1675 */
#define WBC_MODE_SWING_BR_FL 1   /* main.cpp:1155-1424: stance rows BL,FR; swing rows BR,FL */
#define WBC_MODE_SWING_BL_FR 2   /* main.cpp:1703-1949: stance rows BR,FL; swing rows BL,FR */

typedef struct {
    double kcom, dcom;        /* 2500, 50          main.cpp:1019-1020 */
    double q1_weight;         /* 50                main.cpp:997       */
    double slack_weight;      /* 1e8               main.cpp:1187      */
    double mu;                /* 0.6               main.cpp:1062      */
    double tau_max;           /* 60                main.cpp:1090-1091 */
    double joint_dt;          /* 0.025             main.cpp:1098      */
    double kp_sw, kd_sw;      /* 300, 20           main.cpp:1371-1373 */
    double g_acc;             /* 9.81              main.cpp:702, 1016 */
    double obs_gain;          /* 10                main.cpp:708       */
    double obs_dt;            /* 0.0025            main.cpp:715       */
    int observer_enabled;     /* reference ships 0 (calls commented out, main.cpp:1029); north_star: 1 */
    int fix_swing_rhs;        /* 0 = keep the reference quirk (swing equality rhs = 0, main.cpp:1238-1241) */
    /* SURVEY.md 8f-3 (the dead ESTIMATOR_SEM, estimator_sem.cpp:24-61; second gain main.cpp:707-708) -- PARITY UNPINNED: the
     * reference never runs these forms, there is nothing to compare against except this restatement and the closed-form response */
    double obs_gain2;         /* 1: second entry of the coefficient vector {10, 1}                   */
    int obs_order;            /* 1 = estimate() as main.cpp runs it; 2 = second-order recursion via ygamma */
    int obs_form;             /* 0 = (I + kT)^-1 k (main.cpp:716-718); 1 = k (estimator_sem.cpp:55-57)  */
} wbc_oracle_params;

void wbc_oracle_default_params(wbc_oracle_params* p);

/* One instance, array-of-struct for the oracle's convenience.  Foot stacking order everywhere
 * is BR, BL, FL, FR (main.cpp:674-686).  DoF order is the canonical one (tools/gen_model.py). */
typedef struct {
    double base_pos[3];
    double base_R[9];       /* world_R_base, row-major (rotation block of _world_H_base)            */
    double rpy[3];          /* _base_pos[3:6] -- read from a member at main.cpp:596 (quirk E14)       */
    double base_vel[6];     /* linear (of the base origin) and angular velocity, world axes (MIXED) */
    double q[12], dq[12];
    double gravity[3];      /* (0,0,-9.8) at main.cpp:855                                           */
    double com_des_pos[6], com_des_vel[6], com_des_acc[6];   /* spline samples, main.cpp:1005-1010  */
    int mode;
    double sw_des_pos[6], sw_des_vel[6], sw_des_acc[6];      /* two swing feet in Jsw row order      */
    double foot_force[12];  /* contact-sensor forces, sensor (foot link) frame, main.cpp:794-834    */
    double terrain[40];     /* per stacked foot: n(3) t1(3) t2(3) mu(1); used when has_terrain != 0   */
    int has_terrain;
    double yd_prev[6], yw_prev[6];  /* observer state carried across cycles, main.cpp:721-724       */
    double yg_prev[6];              /* ygamma_prev (main.cpp:243, 724), read by the second-order form only */
} wbc_oracle_in;

typedef struct {
    double tau[12];
    double w[6];
    double yd[6], yw[6];
    double x[30];
    double qp_obj;          /* 0.5 x'Qx + c'x at x                                                */
    int status;             /* 0 ok, <0 QP solver threw                                           */
    int ncholesky;
    double yg[6];           /* ygamma (second-order form), else a copy of yg_prev                 */
} wbc_oracle_out;

/* Intermediates of update(), exposed so tests can compare stage by stage. */
typedef struct {
    double M[18 * 18], h[18], g[18];
    double Jac_lin[12 * 18], Jdqd_lin[12];            /* linear rows of Jac / Jdqd, base coordinates  */
    double com[3], com_vel[3];
    double foot_pos[12], foot_vel[12], foot_R[36];    /* stacked order BR,BL,FL,FR                    */
    double T[18 * 18], T_inv_dot[18 * 18];
    double Mcom[18 * 18], hcom[18], gcom[18];
    double Jcom_lin[12 * 18], Jdqdcom_lin[12];
} wbc_oracle_dyn;

/* QP layouts: Q 30x30 row-major, c 30, L nrows x 31 row-major. */
typedef struct {
    double Q[900], c[30], L[86 * 31];
    int nrows, neq;
    double Fgrf[12], Wcom_des[6];
} wbc_oracle_qp;

typedef int (*wbc_qp_fn)(const double* Q, const double* c, const double* L, int nrows, int neq,
                         double* x, int* ncholesky, int* termtype);

void wbc_oracle_update(const wbc_oracle_in* in, wbc_oracle_dyn* d);
void wbc_oracle_estimate(const wbc_oracle_params* p, const wbc_oracle_in* in, const wbc_oracle_dyn* d,
                         const double* Fgrf, double* w, double* yd, double* yw);
void wbc_oracle_assemble(const wbc_oracle_params* p, const wbc_oracle_in* in, const wbc_oracle_dyn* d,
                         const double* w, wbc_oracle_qp* qp);
void wbc_oracle_fgrf(const wbc_oracle_in* in, const wbc_oracle_dyn* d, double* Fgrf);
void wbc_oracle_torque(const wbc_oracle_in* in, const wbc_oracle_dyn* d, const double* x, double* tau);
int  wbc_oracle_cycle(const wbc_oracle_params* p, const wbc_oracle_in* in, wbc_qp_fn solve,
                      wbc_oracle_out* out, wbc_oracle_dyn* dyn_opt, wbc_oracle_qp* qp_opt);
/* Synthetic plant of BASELINE config 5 (CoM momentum integrator, joints locked; see wbc_oracle.c). push = world wrench
 * at the CoM (6); x = QP solution (30) for the closed loop or NULL; outputs the next base position (3), base twist (6)
 * and, in closed loop, the next sensor-frame foot forces (12). */
void wbc_oracle_plant_step(const wbc_oracle_params* p, const wbc_oracle_in* in, const double* push, const double* x,
                           double* base_pos_out, double* base_vel_out, double* foot_force_out);
/* Forward dynamics with hard point contacts at the stance feet (stands in for Gazebo + ModelPush, parity
 * unpinned): one control period under joint torques tau (12), a world wrench push (6) at the CoM, nsub semi-implicit Euler
 * substeps, velocity-level constraint stabilisation gamma.  out = next state (pose, twist, q, dq; the other fields are copied),
 * foot_force_out (12) = contact forces in the sensor frames, diag[0] = contact-constraint residual, diag[1] = smallest normal force. */
void wbc_oracle_fdyn_step(const wbc_oracle_params* p, const wbc_oracle_in* in, const double* tau, const double* push, int nsub,
                          double gamma, wbc_oracle_in* out, double* foot_force_out, double* diag);
/* towr::Spline::GetPoint(t) for one spline of nseg cubic-Hermite polynomials in 3 dimensions (spline.cc:48-93,
 * polynomial.cc:50-104): durations[nseg]; nodes[(nseg+1)][6] = position(3), velocity(3) per node.
 * Outputs p, v, a (3 each).  Returns the polynomial id, or -1 where the reference would assert (t < 0, t past the end). */
int wbc_oracle_spline_point(int nseg, const double* durations, const double* nodes, double t, double* p, double* v, double* a);
/* n instances on nthreads host threads (disjoint contiguous ranges); returns wall seconds. */
double wbc_oracle_batch(const wbc_oracle_params* p, const wbc_oracle_in* in, int n, int nthreads,
                        wbc_qp_fn solve, wbc_oracle_out* out);

#ifdef __cplusplus
}
#endif
#endif
