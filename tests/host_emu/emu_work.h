// TEST-ONLY: heap-backed wbcqp::Work for the single-lane host emulation of the device solver.
#pragma once
#include <vector>

#include "../../wbc_quadruped_dob_b200/csrc/qp_warp.cuh"

struct EmuWork {
    std::vector<double> gbuf, kbuf, sbuf;
    wbcqp::Work w;
    EmuWork() : gbuf(wbcqp::gl::TOTAL, 0.0), kbuf(wbcqp::gl::KKT_DOUBLES, 0.0), sbuf(wbcqp::sl::TOTAL, 0.0)
    {
        w.g = gbuf.data();
        w.kkt = kbuf.data();
        w.sm = sbuf.data();
    }
};
