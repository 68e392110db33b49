// TEST-ONLY: heap-backed wbcqp::Work for the single-lane host emulation of the device solver.
#pragma once
#include <vector>

#include "../../wbc_quadruped_dob_b200/csrc/qp_team.cuh"

struct EmuWork {
    std::vector<double> buf;
    std::vector<int> ibuf;
    wbcqp::Work w;
    EmuWork()
    {
        using namespace wbcqp;
        const int nqmax = MAXNT + MAXK;
        const size_t nd = 944 + (size_t)MAXNT * LDG + (size_t)NVEC * VLG + 2 * nqmax + 4 + nqmax + (size_t)kkt_doubles(nqmax) +
                          S_DOUBLES + NVEC * NCAP + MAXK * 31 + 32 * 5 + MAXNIC + 2 * MAXK + 2 * 104 + 64;
        buf.assign(nd, 0.0);
        ibuf.assign(72 + 104 + 104 + 8, 0);
        double* p = buf.data();
        w.SA = p; p += 944; w.Sgl = p; p += (size_t)MAXNT * LDG; w.vgl = p; p += NVEC * VLG; w.qrv = p; p += 2 * nqmax + 4;
        w.sv0 = p; p += nqmax; w.kkt = p; p += kkt_doubles(nqmax);
        w.Ssh = p; p += S_DOUBLES; w.vsh = p; p += NVEC * NCAP; w.C = p; p += MAXK * 31;
        w.larinv = p; p += 32; w.ladiag = p; p += 32; w.b = p; p += 32; w.s = p; p += 32; w.xs = p; p += 32;
        w.nicerr = p; p += MAXNIC; w.nulc = p; p += MAXK; w.nulcest = p; p += MAXK; w.exxc = p; p += 104; w.exb = p; p += 104;
        int* ip = ibuf.data();
        w.nicnact = ip; w.cstatus = ip + 72; w.isfree = ip + 72 + 104; w.iscr = ip + 72 + 104 + 104;
    }
};
