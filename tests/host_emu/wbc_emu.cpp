// TEST-ONLY host emulation of the device control-cycle path: compiles the product headers
// (wbc_front.cuh, wbc_assemble.cuh, qp_warp.cuh) with g++ and the single-lane `HostEx` executor so
// that the device code's arithmetic and decisions can be unit-tested on a machine without a GPU
// (`pytest -m "not gpu"`).  It is NOT linked into libwbc_b200.so and never used by bench.py or by the
// product path; GPU parity tests call the CUDA kernels through the C ABI instead.
#include "emu_work.h"
#include "../../wbc_quadruped_dob_b200/csrc/wbc_assemble.cuh"
#include "../../wbc_quadruped_dob_b200/csrc/wbc_front.cuh"
#include "../../wbc_quadruped_dob_b200/csrc/wbc_traj.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace wbc;
using namespace wbcqp;

extern "C" {

// Plain-pointer mirror of DevInputs / outputs so ctypes can fill it.
struct EmuIO {
    const double *base_pos, *base_rot, *base_rpy, *base_vel, *q, *dq, *com_des_pos, *com_des_vel, *com_des_acc, *sw_des_pos,
        *sw_des_vel, *sw_des_acc, *foot_force, *terrain;
    const int* mode;
    long ld;
    double *yd, *yw;          // [6][ld] in/out
    double *tau, *w, *x, *qp_obj;   // [12][ld], [6][ld], [30][ld], [ld]
    int* status;              // [ld]
    int* info;                // [8][ld]
    double* rec;              // [ld][QPREC_DOUBLES] (optional)
    double* yg;               // [6][ld] in/out (second-order observer; may be NULL when obs_order == 1)
    double* w3;               // [12][ld] (optional)
};

int emu_cycle(const Params* P, const EmuIO* io, int n)
{
    EmuWork ew;
    DevInputs in;
    in.base_pos = io->base_pos; in.base_rot = io->base_rot; in.base_rpy = io->base_rpy; in.base_vel = io->base_vel;
    in.q = io->q; in.dq = io->dq; in.com_des_pos = io->com_des_pos; in.com_des_vel = io->com_des_vel;
    in.com_des_acc = io->com_des_acc; in.sw_des_pos = io->sw_des_pos; in.sw_des_vel = io->sw_des_vel;
    in.sw_des_acc = io->sw_des_acc; in.foot_force = io->foot_force; in.terrain = io->terrain; in.mode = io->mode; in.obs_gain = nullptr; in.ld = io->ld;
    FrontState st;
    st.yd = io->yd; st.yw = io->yw; st.ld = io->ld; st.yg = io->yg; st.w3 = io->w3; st.w3_ld = io->ld;
    std::vector<double> rec(QPREC_DOUBLES);
    HostEx ex;
    Settings cfg;
    cfg.epsx = P->qp_epsx; cfg.rho = P->qp_rho; cfg.outerits = P->qp_outerits; cfg.kkt_mode = getenv("WBC_EMU_KKT") ? atoi(getenv("WBC_EMU_KKT")) : 1;
    if (getenv("WBC_EMU_PIVTOL")) cfg.kkt_pivtol = atof(getenv("WBC_EMU_PIVTOL"));
    for (long i = 0; i < n; i++) {
        front_cycle(*P, in, st, i, rec.data(), io->w, io->ld, nullptr);
        if (io->rec) memcpy(io->rec + i * QPREC_DOUBLES, rec.data(), sizeof(double) * QPREC_DOUBLES);
        const QpShape sh = qp_shape((int)rec[QR_MODE]);
        assemble_qp<LDH>(ex, *P, rec.data(), sh, W_H(ew.w), W_EXB(ew.w), W_C(ew.w));
        Stats s;
        double xs[30];
        // WBC_EMU_STAGED=1: the three resumable stages the control cycle's solver kernel hands from warp to warp, with the
        // state going through the solve's global block between them (must reproduce solve_denseaul bit for bit)
        if (getenv("WBC_EMU_STAGED") && atoi(getenv("WBC_EMU_STAGED"))) solve_staged(ex, ew.w, cfg, sh.nrows, sh.neq, s);
        else solve_denseaul(ex, ew.w, cfg, sh.nrows, sh.neq, s);
        memcpy(xs, W_XS(ew.w), sizeof(xs));
        if (s.termination != 2) memset(xs, 0, sizeof(xs));
        torque_and_objective(ex, *P, rec.data(), sh, xs, io->tau + i, io->ld, io->qp_obj ? io->qp_obj + i : nullptr);
        if (io->x) for (int k = 0; k < 30; k++) io->x[k * io->ld + i] = xs[k];
        if (io->status) io->status[i] = s.termination == 2 ? 0 : s.termination;
        if (io->info) {
            int v[8] = {s.ncholesky, s.outer_its, s.qqp_calls, s.nicwork, s.kkt_dim_max, s.flags, 0, 0};
            for (int k = 0; k < 8; k++) io->info[k * io->ld + i] = v[k];
        }
    }
    return 0;
}

// Assemble only: dense Q (900), c (30), L (86*31) for instance i of a record array.
int emu_assemble(const Params* P, const double* rec, double* Q, double* c, double* L, int* nrows, int* neq)
{
    HostEx ex;
    const QpShape sh = qp_shape((int)rec[QR_MODE]);
    assemble_qp<30>(ex, *P, rec, sh, Q, c, L);
    *nrows = sh.nrows; *neq = sh.neq;
    return 0;
}

// The device's spline sampling (wbc_traj.cuh) for n instances: out36 is [36][ld] in wbc_traj_kernel's block order.
void emu_sample_trajectory(int n, int nseg, const double* dur, const double* nodes, long ld, const double* t, double* out36)
{
    for (int i = 0; i < n; i++) {
        wbc::TrajOut o;
        for (int b = 0; b < 6; b++) o.p[b] = out36 + (long)b * 6 * ld + i;
        o.ld = ld;
        wbc::sample_trajectory_instance(nseg, dur + i, nodes + i, ld, t[i], o);
    }
}

// The device's forward-dynamics plant step (wbc_front.cuh, fdyn_step_instance) for n instances; all arrays SoA [k][ld], state in place.
void emu_fdyn_step(const Params* P, int n, long ld, double* base_pos, double* base_rot, double* base_rpy, double* base_vel, double* q, double* dq,
                   double* foot_force, const int* mode, const double* tau, const double* push, double* diag, int nsub, double gamma)
{
    wbc::FdynIO io;
    io.base_pos = base_pos; io.base_rot = base_rot; io.base_rpy = base_rpy; io.base_vel = base_vel; io.q = q; io.dq = dq; io.foot_force = foot_force;
    io.mode = mode; io.tau = tau; io.push = push; io.diag = diag; io.ld = ld;
    for (long i = 0; i < n; i++) wbc::fdyn_step_instance(*P, io, i, nsub, gamma);
}

int emu_sizeof_params() { return (int)sizeof(Params); }
int emu_qprec_doubles() { return QPREC_DOUBLES; }
}
