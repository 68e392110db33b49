// Shared device/host types of the batched WBC control-cycle path (internal; the public C ABI is
// include/wbc_b200.h).  Layouts:
//   * inputs / outputs: SoA, component-major [k][ld] doubles (instance index fastest) so that the
//     thread-per-instance front kernel loads and stores fully coalesced;
//   * QP record (front kernel -> solver kernel): AoS, one QPREC_DOUBLES-double record per instance,
//     so that the warp-per-instance solver kernel reads it with coalesced 256-byte requests.
#pragma once

#if defined(__CUDACC__)
#define WBC_HD __host__ __device__ __forceinline__
#define WBC_HDN __host__ __device__
#define WBC_DEVFN __device__      // reads __constant__ model tables: device-only under nvcc
#else
#define WBC_HD inline
#define WBC_HDN
#define WBC_DEVFN
#endif

namespace wbc {

constexpr int MODE_STANCE = 0;        // main.cpp:979-1148, 1516-1675
constexpr int MODE_SWING_BR_FL = 1;   // main.cpp:1155-1424: stance rows BL,FR (3,9); swing rows BR,FL (0,6)
constexpr int MODE_SWING_BL_FR = 2;   // main.cpp:1703-1949: stance rows BR,FL (0,6); swing rows BL,FR (3,9)

// Literal gains / limits of the reference (SURVEY.md section 5 "Config / flags").
struct Params {
    double kcom, dcom;        // 2500, 50          main.cpp:1019-1020
    double q1_weight;         // 50                main.cpp:997
    double slack_weight;      // 1e8               main.cpp:1187
    double mu;                // 0.6               main.cpp:1062
    double tau_max;           // 60                main.cpp:1090-1091
    double joint_dt;          // 0.025             main.cpp:1098
    double kp_sw, kd_sw;      // 300, 20           main.cpp:1371-1373
    double g_acc;             // 9.81              main.cpp:702, 1016
    double obs_gain;          // 10                main.cpp:708
    double obs_dt;            // 0.0025            main.cpp:715
    double gravity[3];        // (0,0,-9.8)        main.cpp:855
    double qp_epsx, qp_rho;   // 1e-2, 1e4         lopt.cpp:101
    double obs_gain2;         // 1                 second entry of the coefficient vector {10, 1} (main.cpp:707-708, estimator_sem.cpp:47-48)
    int qp_outerits;          // 5                 lopt.cpp:101
    int observer_enabled;     // north_star: 1 (the reference ships with the call commented out, main.cpp:1029)
    int fix_swing_rhs;        // 0 keeps the reference's zero swing-equality rhs (main.cpp:1238-1241)
    int qp_literal_kkt;       // 0: reduced multiplier update, literal form as fallback; 1: literal form only (opt.cpp:41803-42032)
    int hold_tau_on_failure;  // 1: a failed instance keeps the last good tau of its index (main.cpp:242; lopt.cpp:114-116 swallows the failure)
    int obs_order;            // 1: the observer main.cpp runs; 2: second-order recursion through the ygamma state
    int obs_form;             // 0: main.cpp:716-718, w = (I + kT)^-1 k (...); 1: estimator_sem.cpp:55-57, w = k (...)
};

struct DevInputs {
    const double* base_pos;     // [3][ld]
    const double* base_rot;     // [9][ld] world_R_base row-major
    const double* base_rpy;     // [3][ld]
    const double* base_vel;     // [6][ld]
    const double* q;            // [12][ld]
    const double* dq;           // [12][ld]
    const double* com_des_pos;  // [6][ld]
    const double* com_des_vel;  // [6][ld]
    const double* com_des_acc;  // [6][ld]
    const double* sw_des_pos;   // [6][ld]
    const double* sw_des_vel;   // [6][ld]
    const double* sw_des_acc;   // [6][ld]
    const double* foot_force;   // [12][ld]
    const double* terrain;      // [40][ld] or nullptr (flat ground, main.cpp:1062-1078)
    const int* mode;            // [ld]
    const double* obs_gain;     // [ld] or nullptr: per-instance observer gain k0 (config 5's sweep); nullptr = Params::obs_gain
    long ld;
};

struct DevOutputs {
    double* tau;        // [12][ld]
    double* w;          // [6][ld]
    double* x;          // [30][ld] or nullptr
    double* qp_obj;     // [ld] or nullptr
    int* status;        // [ld] or nullptr
    long ld;
};

// Intermediates of update() for stage-by-stage parity tests (wbc_debug_update).  SoA [k][ld].
struct DevDebug {
    double* M;          // [324] free-floating mass matrix, MIXED representation
    double* h;          // [18]
    double* g;          // [18]
    double* Jac_lin;    // [216]
    double* Jdqd_lin;   // [12]
    double* com;        // [3]
    double* com_vel;    // [3]
    double* Mcom_b;     // [36]  MassMatrixCOM[0:6,0:6]
    double* Mcom_j;     // [144] MassMatrixCOM[6:18,6:18]
    double* hcom;       // [18]
    double* gcom;       // [18]
    double* Jcom_lin;   // [216]
    double* Jdqdcom_lin;// [12]
    double* foot_pos;   // [12]
    double* foot_vel;   // [12]
    double* Fgrf;       // [12]
    double* Wcom_des;   // [6]
    long ld;
};

// QP record offsets (doubles).  Foot-coordinate index r = 3*stacked_foot + axis, stacked order BR,BL,FL,FR.
constexpr int QR_MC = 0;        // [6][6]   Mc  = MassMatrixCOM[0:6,0:6]
constexpr int QR_HC = 36;       // [6]      hc  = BiasCOM[0:6]
constexpr int QR_HJ = 42;       // [12]     hj  = BiasCOM[6:18]
constexpr int QR_MJJ = 54;      // [12][12] Mjj = MassMatrixCOM[6:18,6:18]
constexpr int QR_JC = 198;      // [12][6]  JacCOM_lin[:,0:6]
constexpr int QR_JJ = 270;      // [12][12] JacCOM_lin[:,6:18]
constexpr int QR_JDQD = 414;    // [12]     JdqdCOM_lin
constexpr int QR_WCOM = 426;    // [6]      Wcom_des
constexpr int QR_DDQMAX = 432;  // [12]
constexpr int QR_DDQMIN = 444;  // [12]
constexpr int QR_SWRHS = 456;   // [6]      vdotswdes - Jdqdsw (main.cpp:1378)
constexpr int QR_CFR = 462;     // [4][15]  friction rows per stacked foot (main.cpp:1072-1078)
constexpr int QR_MODE = 522;    // [1]      contact mode as a double
// momentum balance of the cycle, for the synthetic plant of BASELINE config 5 (wbc_plant_kernel); not read by the solver
constexpr int QR_RHO = 523;     // [6]      rho = Mc CoM_vel                       (main.cpp:699, 705)
constexpr int QR_DD = 529;      // [6]      d = -m g_acc e3 + Jc' Fgrf             (main.cpp:700-706)
constexpr int QR_XBC = 535;     // [3]      com - base origin                      (main.cpp:518)
constexpr int QR_FOOTR = 538;   // [4][9]   world_R_foot per stacked foot, row-major (main.cpp:459-487)
constexpr int QPREC_DOUBLES = 576;   // padded to a multiple of 16 doubles (128 B)

}  // namespace wbc
