// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin C-ABI wrapper around the reference's own vendored ALGLIB 3.16.0, compiled
// from the sources where they lie under /root/reference (see oracle/Makefile).
// It replays the exact call sequence of the reference's OPT::opt_stance /
// OPT::opt_swing (dogbot_controller/src/lopt.cpp:84-117, 121-154):
//   minqpcreate(30) -> minqpsetquadraticterm -> minqpsetlinearterm ->
//   minqpsetlc(L, ct) with ct[i]=0 for i<neq else -1 (lopt.cpp:35-66) ->
//   minqpsetscaleautodiag -> minqpsetalgodenseaul(1e-2, 1e4, 5) -> minqpoptimize
//   -> minqpresults.
// Exceptions are swallowed like the reference does (lopt.cpp:114-116) but are
// reported through the return code so tests can see them.
#include "alglib/optimization.h"
#include <cstring>

extern "C" {

// Q: n x n row-major, c: n, L: nrows x (n+1) row-major, first neq rows are
// equalities, the rest are "<=" rows.  Returns 0 on success, -1 if ALGLIB threw.
int ref_qp_solve_ex(int n, const double* Q, const double* c, const double* L,
                    int nrows, int neq, double epsx, double rho, int outerits,
                    double* x, int* ncholesky, int* termtype)
{
    try {
        alglib::real_2d_array aQ, aL;
        alglib::real_1d_array ac, ax;
        alglib::integer_1d_array ct;
        aQ.setlength(n, n);
        ac.setlength(n);
        aL.setlength(nrows, n + 1);
        ct.setlength(nrows);
        for (int i = 0; i < n; i++) {
            for (int j = 0; j < n; j++) aQ(i, j) = Q[i * n + j];
            ac(i) = c[i];
        }
        for (int i = 0; i < nrows; i++) {
            for (int j = 0; j <= n; j++) aL(i, j) = L[i * (n + 1) + j];
            ct(i) = (i < neq) ? 0 : -1;
        }
        alglib::minqpstate state;
        alglib::minqpreport rep;
        alglib::minqpcreate(n, state);
        alglib::minqpsetquadraticterm(state, aQ);
        alglib::minqpsetlinearterm(state, ac);
        alglib::minqpsetlc(state, aL, ct);
        alglib::minqpsetscaleautodiag(state);
        alglib::minqpsetalgodenseaul(state, epsx, rho, outerits);
        alglib::minqpoptimize(state);
        alglib::minqpresults(state, ax, rep);
        for (int i = 0; i < n; i++) x[i] = ax(i);
        if (ncholesky) *ncholesky = (int)rep.ncholesky;
        if (termtype) *termtype = (int)rep.terminationtype;
        return 0;
    } catch (alglib::ap_error&) {
        return -1;
    }
}

// The reference's settings (lopt.cpp:101, 138).
int ref_qp_solve(const double* Q, const double* c, const double* L, int nrows, int neq,
                 double* x, int* ncholesky, int* termtype)
{
    return ref_qp_solve_ex(30, Q, c, L, nrows, neq, 1.0e-2, 1.0e+4, 5, x, ncholesky, termtype);
}

// A tightly converged solve of the same QP (DENSE-IPM, eps 1e-13) -- used only to
// classify which reference solves were converged.
int ref_qp_solve_exact(int n, const double* Q, const double* c, const double* L,
                       int nrows, int neq, double* x)
{
    try {
        alglib::real_2d_array aQ, aL;
        alglib::real_1d_array ac, ax;
        alglib::integer_1d_array ct;
        aQ.setlength(n, n);
        ac.setlength(n);
        aL.setlength(nrows, n + 1);
        ct.setlength(nrows);
        for (int i = 0; i < n; i++) {
            for (int j = 0; j < n; j++) aQ(i, j) = Q[i * n + j];
            ac(i) = c[i];
        }
        for (int i = 0; i < nrows; i++) {
            for (int j = 0; j <= n; j++) aL(i, j) = L[i * (n + 1) + j];
            ct(i) = (i < neq) ? 0 : -1;
        }
        alglib::minqpstate state;
        alglib::minqpreport rep;
        alglib::minqpcreate(n, state);
        alglib::minqpsetquadraticterm(state, aQ);
        alglib::minqpsetlinearterm(state, ac);
        alglib::minqpsetlc(state, aL, ct);
        alglib::minqpsetscaleautodiag(state);
        alglib::minqpsetalgodenseipm(state, 1.0e-13);
        alglib::minqpoptimize(state);
        alglib::minqpresults(state, ax, rep);
        for (int i = 0; i < n; i++) x[i] = ax(i);
        return (int)rep.terminationtype > 0 ? 0 : -2;
    } catch (alglib::ap_error&) {
        return -1;
    }
}

}  // extern "C"
