// TEST-ONLY host emulation of the device QP solver: compiles the product header
// wbc_quadruped_dob_b200/csrc/qp_denseaul.cuh with g++ and the single-lane `HostEx` executor so the
// solver's host-visible logic (working-set decisions, phase switching, multiplier update) can be
// unit-tested without a GPU.  It is NOT linked into libwbc_b200.so and is never used by bench.py
// or by the product path; GPU parity tests call the CUDA kernels through the C-ABI instead.
#include "../../wbc_quadruped_dob_b200/csrc/qp_denseaul.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace wbcqp;

extern "C" int emu_qp_solve(const double* Q, const double* c, const double* L, int nrows, int neq,
                            double epsx, double rho, int outerits, int kkt_mode, double* x, int* istats,
                            double* dstats)
{
    static thread_local std::vector<double> buf;
    static thread_local std::vector<int> ibuf;
    const int nqmax = MAXNT + MAXK;
    size_t nd = 900 + 30 + 30 + MAXK * 31 + MAXNIC + 2 * MAXK + 2 * MAXNT + 2 * MAXNT * MAXNT + 12 * MAXNT +
                (size_t)kkt_doubles(nqmax) + 2 * nqmax + 2 + nqmax;
    buf.assign(nd, 0.0);
    ibuf.assign(MAXNIC + 2 * MAXNT, 0);
    double* p = buf.data();
    Work w;
    w.A = p; p += 900; w.b = p; p += 30; w.s = p; p += 30; w.C = p; p += MAXK * 31;
    w.nicerr = p; p += MAXNIC; w.nulc = p; p += MAXK; w.nulcest = p; p += MAXK;
    w.exxc = p; p += MAXNT; w.exb = p; p += MAXNT;
    w.exa = p; p += MAXNT * MAXNT; w.z = p; p += MAXNT * MAXNT;
    w.xc = p; p += MAXNT; w.xp = p; p += MAXNT; w.xf = p; p += MAXNT; w.gc = p; p += MAXNT;
    w.cgc = p; p += MAXNT; w.cgp = p; p += MAXNT; w.dc = p; p += MAXNT; w.dp = p; p += MAXNT;
    w.tmp0 = p; p += MAXNT; w.tmp1 = p; p += MAXNT; w.regdiag = p; p += MAXNT; w.bufr = p; p += MAXNT;
    w.kkt = p; p += kkt_doubles(nqmax); w.qrv = p; p += 2 * nqmax + 2; w.sv0 = p; p += nqmax;
    int* ip = ibuf.data();
    w.nicnact = ip; ip += MAXNIC; w.cstatus = ip; ip += MAXNT; w.isfree = ip; ip += MAXNT;
    Settings cfg;
    cfg.epsx = epsx; cfg.rho = rho; cfg.outerits = outerits; cfg.kkt_mode = kkt_mode;
    Stats st;
    HostEx ex;
    solve_denseaul(ex, w, cfg, Q, 1, c, 1, L, 1, nrows, neq, x, 1, st);
    if (istats) {
        istats[0] = st.termination; istats[1] = st.ncholesky; istats[2] = st.outer_its; istats[3] = st.qqp_calls;
        istats[4] = st.nicwork; istats[5] = st.kkt_dim_max; istats[6] = st.flags;
    }
    if (dstats) dstats[0] = st.flops;
    return st.termination == 2 ? 0 : -1;
}
