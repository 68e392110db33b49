// Batched dense QP solver for the WBC ground-reaction-force QP: a decision-for-decision restatement
// of the algorithm the reference controller runs through ALGLIB 3.16.0
//   minqpoptimize (opt.cpp:48020) -> qpdenseauloptimize (41087) -> qqpoptimize (29675)
// with the reference's settings (lopt.cpp:91-106: autodiag scaling, DENSE-AUL epsx=1e-2, rho=1e4,
// 5 outer iterations, cold start, no box constraints on x).  "opt.cpp" = the reference's
// dogbot_controller/src/alglib/optimization.cpp, "linalg.cpp" likewise.
//
// Execution model: ONE WARP PER QP.  Every routine is written against an executor `Ex` that
// provides lane(), NL (lanes), warp all-reduces and a warp barrier.  All control flow depends only on
// values that are bit-identical in every lane (butterfly reductions), so a warp never diverges on
// an instance's data; different warps follow different instances' iteration counts freely.
// `HostEx` (NL = 1) lets the same source be compiled by g++ for CPU-side unit tests of the host
// logic (tests/host_emu); the shipped library only ever instantiates `WarpEx`.
//
// Specialisation relative to generic ALGLIB (all other cases cannot occur on this path):
//   * dense A, no sparse constraints, x unbounded, start point 0, origin 0;
//   * hence in QQP the only bounds are "slack >= 0" on variables i >= NMAIN.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WBC_HD __host__ __device__ __forceinline__
#define WBC_HDN __host__ __device__
#else
#define WBC_HD inline
#define WBC_HDN
#endif

namespace wbcqp {

constexpr int NMAIN = 30;              // decision variables (main.cpp:266 OPT(30,86,82))
constexpr int MAXK = 88;               // >= 86 constraint rows (stance), 82 (swing)
constexpr int MAXNIC = 72;             // >= 68 / 70 inequality rows
constexpr int MAXNT = NMAIN + MAXNIC;  // extended variable count upper bound
constexpr double MACHEPS = 5.0e-16;    // ae_machineepsilon (ap.cpp: 5E-16, NOT DBL_EPSILON)
constexpr double BIGSTEP = 1.0e50;     // opt.cpp:27458

struct Settings {
    double epsx = 1.0e-2;   // lopt.cpp:101
    double rho = 1.0e4;
    int outerits = 5;
    int kkt_mode = 0;       // 0 = literal stacked KKT least squares (opt.cpp:41803-42032); 1 = reduced
    double kkt_pivtol = 1.0e-5;   // reduced form: relative Schur pivot below which the literal form is used
};

struct Stats {
    int termination;    // 2 = ok (opt.cpp:41583); -9 non-positive diagonal (48178-48181)
    int ncholesky;      // rep.ncholesky (opt.cpp:41325)
    int outer_its;      // outer AUL iterations executed
    int qqp_calls;      // inner QQP solves
    int nicwork;        // final working-set size
    int kkt_dim_max;    // largest (N+K) of the multiplier update
    int flags;          // bit0: A not PD (42500); bit1: rcond retry would trigger (42008); bit2: QQP -4
    double flops;       // instrumented algorithmic flop count (DESIGN.md section "work per solve")
};

// Per-instance scratch.  Pointers so the caller decides what lives in shared vs global memory.
struct Work {
    double* A;        // [30*30]  sclsfta, full symmetric
    double* b;        // [30]     sclsftb
    double* s;        // [30]     variable scales
    double* C;        // [MAXK*31] sclsftcleic (rows physically swapped like opt.cpp:41372)
    double* nicerr;   // [MAXNIC]
    int* nicnact;     // [MAXNIC]
    double* nulc;     // [MAXK]
    double* nulcest;  // [MAXK]
    double* exxc;     // [MAXNT]
    double* exb;      // [MAXNT]
    double* exa;      // [n*n]   extended quadratic term == QQP densea (full symmetric)
    double* z;        // [n*n]   QQP densez (Cholesky factor, upper)
    // QQP vectors, each [MAXNT]
    double *xc, *xp, *xf, *gc, *cgc, *cgp, *dc, *dp, *tmp0, *tmp1, *regdiag, *bufr;
    int* cstatus;     // [MAXNT]
    int* isfree;      // [MAXNT]  constrained-Newton free/fixed flags (yidx in opt.cpp:31091)
    double* kkt;      // multiplier-update buffer, see kkt_doubles()
    double* qrv;      // [2*(MAXNT+MAXK)+2] Householder vector / column workspace
    double* sv0;      // [MAXNT+MAXK]
};

// ------------------------------------------------------------------------------------------------
// executors
struct HostEx {
    static constexpr int NL = 1;
    WBC_HD int lane() const { return 0; }
    WBC_HD void sync() const {}
    WBC_HD double sum(double v) const { return v; }
    WBC_HD double maxv(double v) const { return v; }
    WBC_HD int sumi(int v) const { return v; }
};

#if defined(__CUDACC__)
struct WarpEx {
    static constexpr int NL = 32;
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ double maxv(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ __forceinline__ int sumi(int v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};
#endif

// ------------------------------------------------------------------------------------------------
// warp-cooperative primitives.  Convention: every routine that writes memory ends with ex.sync().
template <class Ex>
WBC_HD double dotp(const Ex& ex, const double* a, const double* b, int n)
{
    double s = 0.0;
    for (int i = ex.lane(); i < n; i += Ex::NL) s += a[i] * b[i];
    return ex.sum(s);
}
template <class Ex>
WBC_HD double maxabs(const Ex& ex, const double* a, int n)
{
    double m = 0.0;
    for (int i = ex.lane(); i < n; i += Ex::NL) m = fmax(m, fabs(a[i]));
    return ex.maxv(m);
}
template <class Ex>
WBC_HD void vcopy(const Ex& ex, double* d, const double* s, int n)
{
    for (int i = ex.lane(); i < n; i += Ex::NL) d[i] = s[i];
    ex.sync();
}
template <class Ex>
WBC_HD void vset(const Ex& ex, double* d, double v, int n)
{
    for (int i = ex.lane(); i < n; i += Ex::NL) d[i] = v;
    ex.sync();
}
// y = A x (+ b), A full symmetric n x n with leading dimension ld.  Lane i owns y[i] and walks
// column i (== row i by symmetry) so that lanes touch consecutive addresses.
template <class Ex>
WBC_HD void symv(const Ex& ex, const double* A, int ld, int n, const double* x, const double* b, double* y)
{
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double s = 0.0;
        for (int j = 0; j < n; j++) s += A[j * ld + i] * x[j];
        y[i] = b ? s + b[i] : s;
    }
    ex.sync();
}
// In-place upper Cholesky A = U'U of the leading n x n block (row-major, ld).  Returns false on a
// non-positive pivot (linalg.cpp:29204-29235).  Only the upper triangle is referenced/written.
template <class Ex>
WBC_HD bool cholesky_upper(const Ex& ex, double* U, int ld, int n)
{
    for (int j = 0; j < n; j++) {
        // row j: U[j][c] = (A[j][c] - sum_{k<j} U[k][j] U[k][c]) / U[j][j],  c >= j
        for (int c = j + ex.lane(); c < n; c += Ex::NL) {
            double s = 0.0;
            for (int k = 0; k < j; k++) s += U[k * ld + j] * U[k * ld + c];
            U[j * ld + c] -= s;
        }
        ex.sync();
        double ajj = U[j * ld + j];
        if (!(ajj > 0.0)) return false;
        ajj = sqrt(ajj);
        double r = 1.0 / ajj;
        ex.sync();   // every lane has read the pivot before it is overwritten
        for (int c = j + ex.lane(); c < n; c += Ex::NL) U[j * ld + c] = (c == j) ? ajj : U[j * ld + c] * r;
        ex.sync();
    }
    return true;
}
// Solve U'U x = rhs in place (linalg.cpp:38395 fblscholeskysolve, upper).
template <class Ex>
WBC_HD void cholesky_solve(const Ex& ex, const double* U, int ld, int n, double* x)
{
    // forward U' y = b, column-oriented: no reductions
    for (int k = 0; k < n; k++) {
        double yk = x[k] / U[k * ld + k];
        ex.sync();
        for (int i = k + ex.lane(); i < n; i += Ex::NL) {
            if (i == k) x[k] = yk; else x[i] -= U[k * ld + i] * yk;
        }
        ex.sync();
    }
    // backward U x = y, row-oriented dot products
    for (int i = n - 1; i >= 0; i--) {
        double s = 0.0;
        for (int j = i + 1 + ex.lane(); j < n; j += Ex::NL) s += U[i * ld + j] * x[j];
        s = ex.sum(s);
        double xi = (x[i] - s) / U[i * ld + i];
        ex.sync();
        if (ex.lane() == 0) x[i] = xi;
        ex.sync();
    }
}

WBC_HD double safeminposrv(double x, double y, double v)
{ // alglibinternal.cpp:1998
    if (y >= 1.0) {
        double r = x / y;
        return (v > r) ? r : v;
    }
    return (x < v * y) ? x / y : v;
}
WBC_HD void generaterotation(double f, double g, double& cs, double& sn, double& r)
{ // alglibinternal.cpp:9101
    if (g == 0.0) { cs = 1.0; sn = 0.0; r = f; return; }
    if (f == 0.0) { cs = 0.0; sn = 1.0; r = g; return; }
    if (fabs(f) > fabs(g)) { double t = g / f; r = fabs(f) * sqrt(1.0 + t * t); }
    else { double t = f / g; r = fabs(g) * sqrt(1.0 + t * t); }
    cs = f / r; sn = g / r;
    if (fabs(f) > fabs(g) && cs < 0.0) { cs = -cs; sn = -sn; r = -r; }
}
WBC_HD void estimateparabolicmodel(double absasum, double absasum2, double mx, double mb, double md,
                                   double d1, double d2, int& d1est, int& d2est)
{ // opt.cpp:23071-23131
    const double eps = 4 * MACHEPS;
    double e1 = eps * md * (mx * absasum + mb);
    double e2 = eps * md * (mx * sqrt(absasum2) + mb);
    double err = sqrt(e1 * e2);
    d1est = (fabs(d1) <= err) ? 0 : (d1 > 0 ? 1 : (d1 < 0 ? -1 : 0));
    e1 = eps * md * md * absasum;
    e2 = eps * md * md * sqrt(absasum2);
    err = sqrt(e1 * e2);
    d2est = (fabs(d2) <= err) ? 0 : (d2 > 0 ? 1 : (d2 < 0 ? -1 : 0));
}

// ------------------------------------------------------------------------------------------------
// QQP (opt.cpp:29675-30566) specialised to: dense A (akind 2, upper), unit scales, zero origin,
// variables [0,NMAIN) free, variables [NMAIN,n) bounded below by 0.
struct QqpState {
    int n;
    int ld;
    double absasum, absasum2;
    int nfree, cnmodelage;
    int ncholesky;
};

template <class Ex>
WBC_HD double qqp_projected_target(const Ex& ex, const Work& w, const QqpState& q, const double* x,
                                   const double* d, double stp, double* t0, double* t1)
{ // opt.cpp:30582-30652
    const int n = q.n;
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v = (stp != 0.0) ? x[i] + stp * d[i] : x[i];
        if (i >= NMAIN && v < 0.0) v = 0.0;
        t0[i] = v;
    }
    ex.sync();
    double lin = dotp(ex, w.exb, t0, n);
    symv(ex, w.exa, q.ld, n, t0, (const double*)nullptr, t1);
    double quad = dotp(ex, t0, t1, n);
    return lin + 0.5 * quad;
}

template <class Ex>
WBC_HD void qqp_quadratic_model(const Ex& ex, const Work& w, const QqpState& q, const double* x, const double* d,
                                const double* g, double& d1, int& d1est, double& d2, int& d2est, double* t0)
{ // opt.cpp:30753-30821
    const int n = q.n;
    double mx = maxabs(ex, x, n), md = maxabs(ex, d, n), mb = maxabs(ex, w.exb, n);
    symv(ex, w.exa, q.ld, n, d, (const double*)nullptr, t0);
    d2 = 0.5 * dotp(ex, d, t0, n);
    d1 = dotp(ex, d, g, n);
    estimateparabolicmodel(q.absasum, q.absasum2, mx, mb, md, d1, d2, d1est, d2est);
}

// sasconstraineddirection (opt.cpp:27856) in the box-only case: zero the active components; everything
// becomes zero once the active count reaches n (28952-28959).  `nact` = #(cstatus>0) at the last
// sasreactivateconstraints, i.e. at the basis rebuild (28356-28367).
template <class Ex>
WBC_HD void sas_constrained_direction(const Ex& ex, const Work& w, int n, int nact, double* d)
{
    for (int i = ex.lane(); i < n; i += Ex::NL)
        if (w.cstatus[i] > 0 || nact >= n) d[i] = 0.0;
    ex.sync();
}
// sasexploredirection (opt.cpp:27433-27528), box-only.  Sequential scan kept literal (first strict
// improvement wins); executed redundantly by all lanes over the <= MAXNIC bounded variables.
template <class Ex>
WBC_HD void sas_explore_direction(const Ex& ex, const Work& w, int n, const double* d, double& stpmax, int& cidx,
                                  double& cval)
{
    (void)ex;
    stpmax = BIGSTEP; cidx = -1; cval = 0.0;
    for (int i = NMAIN; i < n; i++) {
        if (w.cstatus[i] <= 0 && d[i] < 0.0) {
            double prev = stpmax;
            stpmax = safeminposrv(w.xc[i] - 0.0, -d[i], stpmax);
            if (stpmax < prev) { cidx = i; cval = 0.0; }
        }
    }
}
// sasmoveto (opt.cpp:27574-27723), box-only.
template <class Ex>
WBC_HD void sas_moveto(const Ex& ex, const Work& w, int n, const double* xn, bool needact, int cidx, double cval)
{
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double old = w.xc[i];
        double v = xn[i];
        if (needact && i == cidx) { v = cval; w.cstatus[i] = 1; }
        if (i >= NMAIN && v <= 0.0 && v != old) { v = 0.0; w.cstatus[i] = 1; }
        w.xc[i] = v;
    }
    ex.sync();
}
// qqpsolver_findbeststepandmove (opt.cpp:30882-31003).  t0 holds the candidate point.
template <class Ex>
WBC_HD void qqp_find_best_step_and_move(const Ex& ex, const Work& w, const QqpState& q, const double* d, double stp,
                                        bool needact, int cidx, double cval, const double* addsteps, int addcnt)
{
    const int n = q.n;
    double stpbest = stp;
    if (addcnt > 0) {
        double fbest = qqp_projected_target(ex, w, q, w.xc, d, stpbest, w.tmp0, w.tmp1);
        for (int k = 0; k < addcnt; k++) {
            if (addsteps[k] > stp) {
                double fcand = qqp_projected_target(ex, w, q, w.xc, d, addsteps[k], w.tmp0, w.tmp1);
                if (fcand < fbest) { fbest = fcand; stpbest = addsteps[k]; }
            }
        }
    }
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v = w.xc[i] + stpbest * d[i];
        if (i >= NMAIN && v < 0.0) v = 0.0;
        if (needact && i == cidx) v = cval;
        w.tmp0[i] = v;
    }
    ex.sync();
    sas_moveto(ex, w, n, w.tmp0, needact, cidx, cval);
}

// qqpsolver_cnewtonbuild (opt.cpp:31058-31201).  The factor is produced directly in the "scattered"
// n x n layout ALGLIB ends with (identity rows for fixed variables): factoring the masked matrix
// gives bit-identical entries because the extra terms are exact zeros.
template <class Ex>
WBC_HD bool qqp_cnewton_build(const Ex& ex, const Work& w, QqpState& q)
{
    const int n = q.n, ld = q.ld;
    q.cnmodelage = 0;
    int nf = 0;
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        int fr = !(i >= NMAIN && w.xc[i] == 0.0);
        w.isfree[i] = fr;
        nf += fr;
    }
    nf = ex.sumi(nf);
    ex.sync();
    q.nfree = nf;
    if (nf == 0) return false;
    // regdiag[i] = 1e-9 * sum_j |A_ff[i][j]| over free j (31150-31167); lane i walks column i
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v = 0.0;
        if (w.isfree[i]) {
            for (int j = 0; j < n; j++)
                if (w.isfree[j]) v += fabs(w.exa[j * ld + i]);
            if (v == 0.0) v = 1.0;
        }
        w.regdiag[i] = 1.0e-9 * v;
    }
    ex.sync();
    for (int i = 0; i < n; i++) {
        const int fi = w.isfree[i];
        for (int j = i + ex.lane(); j < n; j += Ex::NL) {
            double v;
            if (fi && w.isfree[j]) v = w.exa[i * ld + j] + (i == j ? w.regdiag[i] : 0.0);
            else v = (i == j) ? 1.0 : 0.0;
            w.z[i * ld + j] = v;
        }
    }
    ex.sync();
    q.ncholesky++;
    return cholesky_upper(ex, w.z, ld, n);
}
// qqpsolver_cnewtonupdate (opt.cpp:31314-31426) + spdmatrixcholeskyupdatefixbuf (linalg.cpp:27657-27819, upper).
template <class Ex>
WBC_HD bool qqp_cnewton_update(const Ex& ex, const Work& w, QqpState& q, int cnmaxupdates)
{
    const int n = q.n, ld = q.ld;
    int ntofix = 0;
    for (int i = ex.lane(); i < n; i += Ex::NL)
        if (w.isfree[i] && i >= NMAIN && w.xc[i] == 0.0) ntofix++;
    ntofix = ex.sumi(ntofix);
    if (ntofix == 0 || ntofix == q.nfree) return false;
    if (q.cnmodelage + ntofix > cnmaxupdates) return false;
    for (int k = NMAIN; k < n; k++) {
        if (!(w.isfree[k] && w.xc[k] == 0.0)) continue;
        // fix variable k
        ex.sync();
        if (k == n - 1) {
            for (int i = ex.lane(); i < n; i += Ex::NL) w.z[i * ld + k] = (i == k) ? 1.0 : 0.0;
            if (ex.lane() == 0) w.isfree[k] = 0;
            ex.sync();
            continue;
        }
        for (int j = k + 1 + ex.lane(); j < n; j += Ex::NL) { w.bufr[j] = w.z[k * ld + j]; w.z[k * ld + j] = 0.0; }
        for (int i = ex.lane(); i <= k; i += Ex::NL) w.z[i * ld + k] = (i == k) ? 1.0 : 0.0;
        if (ex.lane() == 0) w.isfree[k] = 0;
        ex.sync();
        for (int i = k + 1; i < n; i++) {
            double bi = w.bufr[i];
            if (bi != 0.0) {
                double cs, sn, r;
                generaterotation(w.z[i * ld + i], bi, cs, sn, r);
                ex.sync();
                for (int j = i + ex.lane(); j < n; j += Ex::NL) {
                    if (j == i) { w.z[i * ld + i] = r; w.bufr[i] = 0.0; }
                    else {
                        double v = w.z[i * ld + j], vv = w.bufr[j];
                        w.z[i * ld + j] = cs * v + sn * vv;
                        w.bufr[j] = -sn * v + cs * vv;
                    }
                }
                ex.sync();
            }
        }
    }
    q.nfree -= ntofix;
    q.cnmodelage += ntofix;
    return true;
}
// qqpsolver_cnewtonstep (opt.cpp:31474-31536), epsg = 0.
template <class Ex>
WBC_HD bool qqp_cnewton_step(const Ex& ex, const Work& w, const QqpState& q, double* g)
{
    const int n = q.n;
    double v = 0.0;
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        if (!w.isfree[i]) g[i] = 0.0;
        v += g[i] * g[i];
    }
    v = ex.sum(v);
    ex.sync();
    if (sqrt(v) <= 0.0) return false;
    for (int i = ex.lane(); i < n; i += Ex::NL) g[i] = -g[i];
    ex.sync();
    cholesky_solve(ex, w.z, q.ld, n, g);
    return true;
}

// One QQP solve from the point w.exxc (in/out), model (w.exa upper -> symmetrised here, w.exb).
// Returns the QQP termination type.
template <class Ex>
WBC_HD int qqp_optimize(const Ex& ex, const Work& w, int n, double epsx, int maxouterits, int& ncholesky,
                        double& flops)
{
    QqpState q;
    q.n = n; q.ld = n; q.ncholesky = 0; q.nfree = 0; q.cnmodelage = 0;
    const int ld = n;
    // settings: qqploaddefaults (opt.cpp:29533-29547) + overrides (41318-41323)
    const int cgminits = 5;
    int cgmaxits = (int)lround(1 + 0.33 * n);
    if (cgmaxits < cgminits) cgmaxits = cgminits;
    const int cnmaxupdates = (int)lround(1 + 0.1 * n);

    // symmetrise + |A| statistics with ALGLIB's k = (i==v ? 1 : 2) quirk (opt.cpp:29893-29915)
    {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < n; i++) {
            for (int j = i + ex.lane(); j < n; j += Ex::NL) {
                double v = w.exa[i * ld + j];
                double vv = fabs(v);
                w.exa[j * ld + i] = v;
                double k = ((double)i == v) ? 1.0 : 2.0;
                s1 += vv * k;
                s2 += vv * vv * k;
            }
        }
        q.absasum = ex.sum(s1);
        q.absasum2 = ex.sum(s2);
        ex.sync();
    }
    // initial point: clip to bounds (29979-29998)
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v = w.exxc[i];
        if (i >= NMAIN && v < 0.0) v = 0.0;
        w.xf[i] = v;     // xs
    }
    ex.sync();
    int term = 0;
    // NOTE: ALGLIB's single-Cholesky fast path for unconstrained problems (opt.cpp:30033-30073) is gated
    // on akind==0 (CQM storage); DENSE-AUL calls QQP with akind==2 (opt.cpp:41324), so the generic
    // CG + constrained-Newton iteration below runs even when there are no slack variables yet.
    {
        // sasstartoptimization (27377-27399)
        for (int i = ex.lane(); i < n; i += Ex::NL) {
            double v = w.xf[i];
            int cs = -1;
            if (i >= NMAIN && v <= 0.0) { v = 0.0; cs = 0; }
            w.xc[i] = v;
            w.cstatus[i] = cs;
        }
        ex.sync();
        int cgmax = cgminits;
        int outerits = 0;
        double stpbuf[3];
        for (;;) {
            if (maxouterits > 0 && outerits >= maxouterits) { term = 5; break; }
            if (outerits > 0) {
                // epsx stopping test (30137-30149); epsf = 0 so the function test is skipped
                double v = 0.0;
                for (int i = ex.lane(); i < n; i += Ex::NL) { double t = w.xp[i] - w.xc[i]; v += t * t; }
                v = ex.sum(v);
                if (sqrt(v) <= epsx) { term = 2; break; }
            }
            outerits++;
            for (int i = ex.lane(); i < n; i += Ex::NL) { w.xp[i] = w.xc[i]; w.cgp[i] = 0.0; w.dp[i] = 0.0; }
            ex.sync();
            for (int cgcnt = 0; cgcnt <= cgmax - 1; cgcnt++) {
                symv(ex, w.exa, ld, n, w.xc, w.exb, w.gc);                     // targetgradient
                flops += 2.0 * n * n;
                // sasreactivateconstraints, box-only (28992-29047)
                int nact = 0;
                for (int i = ex.lane(); i < n; i += Ex::NL) {
                    int cs = -1;
                    if (i >= NMAIN && w.xc[i] == 0.0 && w.gc[i] >= 0.0) { cs = 1; nact++; }
                    w.cstatus[i] = cs;
                    w.cgc[i] = w.gc[i];
                }
                nact = ex.sumi(nact);
                ex.sync();
                sas_constrained_direction(ex, w, n, nact, w.cgc);
                double v = dotp(ex, w.cgc, w.cgc, n);
                if (sqrt(v) <= 0.0) { term = 4; break; }                       // epsg = 0
                // CG direction (30199-30221)
                double vv = 0.0;
                int bflag = 0;
                for (int i = ex.lane(); i < n; i += Ex::NL) {
                    vv += w.cgp[i] * w.cgp[i];
                    if (i >= NMAIN && w.xc[i] == 0.0 && w.dp[i] != 0.0) bflag = 1;
                }
                vv = ex.sum(vv);
                bflag = ex.sumi(bflag);
                bool brst = (bflag != 0) || (vv == 0.0) || (cgcnt % 50 == 0);
                double beta = brst ? 0.0 : v / vv;
                for (int i = ex.lane(); i < n; i += Ex::NL) w.dc[i] = -w.cgc[i] + beta * w.dp[i];
                ex.sync();
                sas_constrained_direction(ex, w, n, nact, w.dc);
                double stpmax, cval; int cidx;
                sas_explore_direction(ex, w, n, w.dc, stpmax, cidx, cval);
                double d1, d2; int d1est, d2est;
                qqp_quadratic_model(ex, w, q, w.xc, w.dc, w.gc, d1, d1est, d2, d2est, w.tmp0);
                flops += 2.0 * n * n;
                if (d1 == 0.0 && d2 == 0.0) { term = 4; break; }
                if (d1est >= 0) { term = 7; break; }
                if (d2est <= 0 && cidx < 0) { term = -4; break; }
                double stp; bool needact; int stpcnt;
                if (d2est > 0) {
                    double fullstp = -d1 / (2 * d2);
                    needact = fullstp >= stpmax;
                    if (needact) { stp = stpmax; stpbuf[0] = stpmax * 4; stpbuf[1] = fullstp; stpbuf[2] = fullstp / 4; stpcnt = 3; }
                    else { stp = fullstp; stpcnt = 0; }
                } else {
                    stp = stpmax; needact = true; stpbuf[0] = 4 * stpmax; stpcnt = 1;
                }
                qqp_find_best_step_and_move(ex, w, q, w.dc, stp, needact, cidx, cval, stpbuf, stpcnt);
                flops += (stpcnt > 0 ? (1 + stpcnt) * 2.0 * n * n : 0.0);
                for (int i = ex.lane(); i < n; i += Ex::NL) { w.dp[i] = w.dc[i]; w.cgp[i] = w.cgc[i]; }
                ex.sync();
            }
            if (term != 0) break;
            cgmax = cgmaxits;
            // constrained Newton phase (30353-30527)
            int newtcnt = 0;
            for (;;) {
                bool b;
                if (newtcnt == 0) {
                    b = qqp_cnewton_build(ex, w, q);
                    flops += (double)n * n * n / 3.0;
                    if (b) cgmax = cgminits;
                } else {
                    b = qqp_cnewton_update(ex, w, q, cnmaxupdates);
                    flops += 3.0 * n * n;
                }
                if (!b) break;
                newtcnt++;
                symv(ex, w.exa, ld, n, w.xc, w.exb, w.gc);
                vcopy(ex, w.dc, w.gc, n);
                if (!qqp_cnewton_step(ex, w, q, w.dc)) break;
                double d1, d2; int d1est, d2est;
                qqp_quadratic_model(ex, w, q, w.xc, w.dc, w.gc, d1, d1est, d2, d2est, w.tmp0);
                flops += 6.0 * n * n;
                if (d1est >= 0) break;
                double stpmax, cval; int cidx;
                if (d2est > 0) {
                    double fullstp = -d1 / (2 * d2);
                    sas_explore_direction(ex, w, n, w.dc, stpmax, cidx, cval);
                    bool needact = fullstp >= stpmax;
                    double stp; int stpcnt;
                    if (needact) { stp = stpmax; stpbuf[0] = stpmax * 4; stpbuf[1] = fullstp; stpbuf[2] = fullstp / 4; stpcnt = 3; }
                    else { stp = fullstp; stpcnt = 0; }
                    qqp_find_best_step_and_move(ex, w, q, w.dc, stp, needact, cidx, cval, stpbuf, stpcnt);
                    flops += (stpcnt > 0 ? (1 + stpcnt) * 2.0 * n * n : 0.0);
                } else {
                    sas_explore_direction(ex, w, n, w.dc, stpmax, cidx, cval);
                    if (cidx < 0) { term = -4; break; }
                    if (stpmax == 0.0) { cgmax = cgmaxits; break; }
                    double f0 = qqp_projected_target(ex, w, q, w.xc, w.dc, 0.0, w.tmp0, w.tmp1);
                    double f1 = qqp_projected_target(ex, w, q, w.xc, w.dc, stpmax, w.tmp0, w.tmp1);
                    if (f1 >= f0) { cgmax = cgmaxits; break; }
                    stpbuf[0] = stpmax * 4; stpbuf[1] = 1.00; stpbuf[2] = 0.25;
                    qqp_find_best_step_and_move(ex, w, q, w.dc, stpmax, true, cidx, cval, stpbuf, 3);
                    flops += 12.0 * n * n;
                }
            }
            if (term != 0) break;
        }
        vcopy(ex, w.xf, w.xc, n);
    }
    // unpack (30546-30565): unit scale, zero origin; slacks clipped / snapped to the bound
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v = w.xf[i];
        if (i >= NMAIN && (v < 0.0 || v == 0.0)) v = 0.0;
        w.exxc[i] = v;
    }
    ex.sync();
    ncholesky += q.ncholesky;
    return term;
}

// ------------------------------------------------------------------------------------------------
// Multiplier update, literal form (opt.cpp:41803-42032): Householder QR of the stacked system
// [K | r ; lambda*mxdiag*I | 0], K = KKT matrix of the equality-constrained model with the columns of
// exactly-active slacks replaced, then back-substitution.  Only the multiplier part of the solution
// is needed, so back-substitution stops at row ntotal.  Structural zeros of the regulariser block
// are skipped (reflector j only touches K rows j.. and regulariser rows 0..j).
WBC_HD int kkt_doubles(int nq) { return 2 * nq * (nq + 1); }

template <class Ex>
WBC_HD void householder_qr_solve_tail(const Ex& ex, double* M, int nq, int ld, int ntail, double* v, double* sol,
                                      double& flops)
{
    // M: 2nq x (nq+1) row-major (ld = nq+1).  On exit the upper triangle holds R and column nq holds Q'r.
    for (int j = 0; j < nq; j++) {
        // rows involved: K rows j..nq-1 and regulariser rows nq..nq+j  -> contiguous range j..nq+j
        const int r0 = j, r1 = nq + j;   // inclusive
        const int len = r1 - r0 + 1;
        // generatereflection (linalg.cpp:19116-19213) on x = M[r0..r1][j]
        double alpha = M[r0 * ld + j];
        double mx = 0.0;
        for (int r = r0 + ex.lane(); r <= r1; r += Ex::NL) { double t = M[r * ld + j]; v[r - r0] = t; mx = fmax(mx, fabs(t)); }
        mx = ex.maxv(mx);
        ex.sync();
        double xnorm = 0.0;
        if (mx != 0.0) {
            double s = 0.0;
            for (int r = 1 + ex.lane(); r < len; r += Ex::NL) { double t = v[r] / mx; s += t * t; }
            s = ex.sum(s);
            xnorm = sqrt(s) * mx;
        }
        double tau = 0.0, beta = alpha;
        if (xnorm != 0.0) {
            double m2 = fmax(fabs(alpha), fabs(xnorm));
            double a = alpha / m2, b = xnorm / m2;
            beta = -m2 * sqrt(a * a + b * b);
            if (alpha < 0.0) beta = -beta;
            tau = (beta - alpha) / beta;
            double sc = 1.0 / (alpha - beta);
            for (int r = 1 + ex.lane(); r < len; r += Ex::NL) v[r] *= sc;
            if (ex.lane() == 0) v[0] = 1.0;
        }
        ex.sync();
        if (tau != 0.0) {
            // apply H = I - tau v v' to columns j+1..nq
            for (int c = j + 1 + ex.lane(); c <= nq; c += Ex::NL) {
                double s = 0.0;
                for (int r = 0; r < len; r++) s += v[r] * M[(r0 + r) * ld + c];
                s *= tau;
                for (int r = 0; r < len; r++) M[(r0 + r) * ld + c] -= s * v[r];
            }
            flops += 4.0 * len * (nq - j);
        }
        if (ex.lane() == 0) M[r0 * ld + j] = beta;
        ex.sync();
    }
    // back-substitution for the last ntail unknowns (42013-42021)
    for (int i = nq - 1; i >= nq - ntail; i--) {
        double s = 0.0;
        for (int jj = i + 1 + ex.lane(); jj < nq; jj += Ex::NL) s += M[i * ld + jj] * sol[jj];
        s = ex.sum(s);
        double xi = (M[i * ld + nq] - s) / M[i * ld + i];
        ex.sync();
        if (ex.lane() == 0) sol[i] = xi;
        ex.sync();
    }
}

template <class Ex>
WBC_HD void update_lagrange_multipliers(const Ex& ex, const Work& w, int nec, int nic, Stats& st)
{
    const int ntotal = NMAIN + nic, ktotal = nec + nic, nq = ntotal + ktotal, ld = nq + 1;
    double* M = w.kkt;
    if (nq > st.kkt_dim_max) st.kkt_dim_max = nq;
    // reference point (X0, L0) (41888-41895)
    for (int i = ex.lane(); i < nq; i += Ex::NL) w.sv0[i] = (i < ntotal) ? w.exxc[i] : w.nulcest[i - ntotal];
    // zero fill
    for (int i = ex.lane(); i < 2 * nq * ld; i += Ex::NL) M[i] = 0.0;
    ex.sync();
    double mxdiag = 0.0;
    for (int i = 0; i < NMAIN; i++) mxdiag = fmax(mxdiag, fabs(w.A[i * NMAIN + i]));
    if (mxdiag == 0.0) mxdiag = 1.0;
    const double lambdareg = 1.0e-8;
    // quadratic term and -b (41919-41927)
    for (int i = 0; i < NMAIN; i++)
        for (int j = ex.lane(); j <= NMAIN; j += Ex::NL)
            M[i * ld + (j < NMAIN ? j : nq)] = (j < NMAIN) ? w.A[i * NMAIN + j] : -w.b[i];
    // constraints (41933-41946)
    for (int i = 0; i < ktotal; i++) {
        for (int j = ex.lane(); j < NMAIN; j += Ex::NL) {
            double c = -w.C[i * 31 + j];
            M[(ntotal + i) * ld + j] = c;
            M[j * ld + ntotal + i] = c;
        }
        if (ex.lane() == 0) {
            if (i >= nec) {
                M[(ntotal + i) * ld + NMAIN + (i - nec)] = -1.0;
                M[(NMAIN + (i - nec)) * ld + ntotal + i] = -1.0;
            }
            M[(ntotal + i) * ld + nq] = -w.C[i * 31 + NMAIN];
        }
    }
    // regulariser rows (41952-41959)
    for (int i = ex.lane(); i < nq; i += Ex::NL) M[(nq + i) * ld + i] = lambdareg * mxdiag;
    ex.sync();
    // subtract reference point: rhs_i -= K[i,:] . sv0  (41964-41968), first nq rows only
    for (int i = ex.lane(); i < nq; i += Ex::NL) {
        double v = 0.0;
        for (int j = 0; j < nq; j++) v += M[i * ld + j] * w.sv0[j];
        M[i * ld + nq] -= v;
    }
    ex.sync();
    // active simple constraints: slack exactly zero (41973-41993)
    for (int i = NMAIN; i < ntotal; i++) {
        if (w.exxc[i] == 0.0) {
            for (int j = ex.lane(); j < 2 * nq; j += Ex::NL) M[j * ld + i] = (j == i) ? -1.0 : 0.0;
        }
    }
    ex.sync();
    st.flops += 2.0 * nq * nq;
    householder_qr_solve_tail(ex, M, nq, ld, ktotal, w.qrv, w.sv0 /*reuse as x1 after reading*/ + 0, st.flops);
    // NOTE: sv0 is overwritten in its tail by the solution; nulcest still holds L0.
    for (int i = ex.lane(); i < ktotal; i += Ex::NL) w.nulcest[i] = w.nulcest[i] + w.sv0[ntotal + i];
    ex.sync();
}


// ------------------------------------------------------------------------------------------------
// Multiplier update, reduced form.  The stacked system above is the KKT system of the equality-
// constrained model  min 1/2 x'Ax + b'x  s.t.  c_r'x = d_r  for r in ACT = {equalities} U {inequality
// rows whose slack is exactly 0}; rows with a free slack get multiplier 0 (their slack-stationarity
// row reads -nu_r = 0).  With A = LA LA' the multipliers solve the Schur-complement system
//     (W W') nu_ACT = d_ACT + W t,   W = C_ACT LA^-T (row m: LA^-1 c_m),  t = LA^-1 b.
// The Tikhonov term of the literal form (lambda = 1e-8 max|A_ii|) changes the result by
// O((lambda/sigma_min(K))^2), far below the parity tolerance unless ACT is (nearly) rank deficient;
// that case is detected by the pivots of the Schur Cholesky and handed to the literal form.
// Returns false when the caller must fall back.
template <class Ex>
WBC_HD bool update_lagrange_multipliers_schur(const Ex& ex, const Work& w, int nec, int nic, Stats& st, double pivtol)
{
    const int ktotal = nec + nic;
    double* LA = w.kkt;                 // [30*30] lower Cholesky factor of A
    double* W = LA + 900;               // [ka][31]: LA^-1 c_m, and entry 30 = rhs d_m
    int nq = NMAIN + nic + ktotal;
    if (nq > st.kkt_dim_max) st.kkt_dim_max = nq;
    // active list (uniform across lanes: reads shared data only)
    int ka = 0;
    int* act = w.isfree;                // reuse: QQP is not running
    ex.sync();
    for (int r = 0; r < ktotal; r++) {
        const bool on = (r < nec) || (w.exxc[NMAIN + (r - nec)] == 0.0);
        if (on) { if (ex.lane() == 0) act[ka] = r; ka++; }
    }
    ex.sync();
    double* Sm = W + (long)ka * 31;     // [ka][ka] Schur complement, lower
    double* tv = Sm + (long)ka * ka;    // [30] t = LA^-1 b, then [ka] solution
    // LA = chol(A), lower, row-major; lane-per-row left-looking columns
    for (int i = ex.lane(); i < NMAIN * NMAIN; i += Ex::NL) LA[i] = w.A[i];
    ex.sync();
    for (int j = 0; j < NMAIN; j++) {
        for (int r = j + ex.lane(); r < NMAIN; r += Ex::NL) {
            double s = LA[r * NMAIN + j];
            for (int k = 0; k < j; k++) s -= LA[r * NMAIN + k] * LA[j * NMAIN + k];
            LA[r * NMAIN + j] = s;
        }
        ex.sync();
        const double d = sqrt(LA[j * NMAIN + j]);
        const double rinv = 1.0 / d;
        ex.sync();
        for (int r = j + ex.lane(); r < NMAIN; r += Ex::NL) LA[r * NMAIN + j] = (r == j) ? d : LA[r * NMAIN + j] * rinv;
        ex.sync();
    }
    // forward substitutions, one lane per right-hand side (ka rows of C and b)
    for (int m = ex.lane(); m <= ka; m += Ex::NL) {
        double* y = (m < ka) ? &W[m * 31] : tv;
        const double* src = (m < ka) ? &w.C[act[m] * 31] : w.b;
        for (int i = 0; i < NMAIN; i++) {
            double s = src[i];
            for (int k = 0; k < i; k++) s -= LA[i * NMAIN + k] * y[k];
            y[i] = s / LA[i * NMAIN + i];
        }
        if (m < ka) y[NMAIN] = src[NMAIN];
    }
    ex.sync();
    // Schur complement (lower) and right-hand side
    for (int e = ex.lane(); e < ka * ka; e += Ex::NL) {
        const int r = e / ka, c = e % ka;
        if (c > r) continue;
        double s = 0.0;
        for (int k = 0; k < NMAIN; k++) s += W[r * 31 + k] * W[c * 31 + k];
        Sm[r * ka + c] = s;
    }
    ex.sync();
    for (int m = ex.lane(); m < ka; m += Ex::NL) {
        double s = W[m * 31 + NMAIN];
        for (int k = 0; k < NMAIN; k++) s += W[m * 31 + k] * tv[k];
        w.sv0[m] = s;
    }
    ex.sync();
    st.flops += 9000.0 + (double)(ka + 1) * NMAIN * NMAIN + (double)ka * ka * NMAIN + 2.0 * ka * NMAIN;
    // delta form: rho = rhs - S nu0 (the literal system is solved for the correction to (X0, L0), and among the
    // solutions of a rank-deficient system its Tikhonov term selects the one of least norm)
    for (int m = ex.lane(); m < ka; m += Ex::NL) {
        double s = 0.0;
        for (int k = 0; k < ka; k++) s += ((k <= m) ? Sm[m * ka + k] : Sm[k * ka + m]) * w.nulcest[act[k]];
        w.sv0[m] = w.sv0[m] - s;
        w.qrv[m] = 0.0;
    }
    ex.sync();
    double* rhs0 = w.qrv + ka;          // copy of rho for the consistency check
    for (int m = ex.lane(); m < ka; m += Ex::NL) rhs0[m] = w.sv0[m];
    // Cholesky of the PSD Schur complement, skipping dependent rows (relative pivot test).  Lt: [ka][ka] lower,
    // columns compacted to the r kept pivots; keep[j] = column of row j or -1.
    double* Lt = tv + 32;
    int* keep = w.cstatus;
    for (int e = ex.lane(); e < ka * ka; e += Ex::NL) Lt[e] = 0.0;
    ex.sync();
    int r = 0;
    bool ambiguous = false;
    for (int j = 0; j < ka; j++) {
        const double d0 = Sm[j * ka + j];
        double piv = d0;
        for (int k = 0; k < r; k++) piv -= Lt[j * ka + k] * Lt[j * ka + k];     // redundant in every lane
        const double ratio = piv / d0;
        if (ratio < pivtol) {
            if (ratio > 1.0e-3 * pivtol) ambiguous = true;
            if (ex.lane() == 0) keep[j] = -1;
            ex.sync();
            continue;
        }
        const double d = sqrt(piv), rinv = 1.0 / d;
        for (int i = j + ex.lane(); i < ka; i += Ex::NL) {
            double v;
            if (i == j) v = d;
            else {
                v = Sm[i * ka + j];
                for (int k = 0; k < r; k++) v -= Lt[i * ka + k] * Lt[j * ka + k];
                v *= rinv;
            }
            Lt[i * ka + r] = v;
        }
        if (ex.lane() == 0) keep[j] = r;
        ex.sync();
        r++;
    }
    if (ambiguous) {
#ifdef WBC_EMU_DEBUG
        printf("schur ambiguous ka=%d r=%d\n", ka, r);
#endif
        return false;
    }
    st.flops += (double)ka * ka * ka / 3.0 + 2.0 * ka * ka;
    if (r == ka) {
        // full rank: L L' delta = rho
        for (int k = 0; k < ka; k++) {
            const double yk = w.sv0[k] / Lt[k * ka + k];
            ex.sync();
            for (int i = k + ex.lane(); i < ka; i += Ex::NL) { if (i == k) w.sv0[k] = yk; else w.sv0[i] -= Lt[i * ka + k] * yk; }
            ex.sync();
        }
        for (int k = ka - 1; k >= 0; k--) {
            const double xk = w.sv0[k] / Lt[k * ka + k];
            ex.sync();
            for (int i = ex.lane(); i <= k; i += Ex::NL) { if (i == k) w.sv0[k] = xk; else w.sv0[i] -= Lt[k * ka + i] * xk; }
            ex.sync();
        }
    } else {
        // rank r < ka: S = Lt Lt' (ka x r), least-norm solution  delta = Lt G^-1 G^-1 Lt' rho,  G = Lt' Lt
        double* G = Lt + (long)ka * ka;      // [r][r] lower
        double* u = G + (long)r * r;         // [r]
        for (int e = ex.lane(); e < r * r; e += Ex::NL) {
            const int a = e / r, c = e % r;
            if (c > a) continue;
            double sacc = 0.0;
            for (int i = 0; i < ka; i++) sacc += Lt[i * ka + a] * Lt[i * ka + c];
            G[a * r + c] = sacc;
        }
        for (int a = ex.lane(); a < r; a += Ex::NL) {
            double sacc = 0.0;
            for (int i = 0; i < ka; i++) sacc += Lt[i * ka + a] * w.sv0[i];
            u[a] = sacc;
        }
        ex.sync();
        for (int j = 0; j < r; j++) {       // Cholesky of G (well conditioned by construction)
            for (int i = j + ex.lane(); i < r; i += Ex::NL) {
                double v = G[i * r + j];
                for (int k = 0; k < j; k++) v -= G[i * r + k] * G[j * r + k];
                G[i * r + j] = v;
            }
            ex.sync();
            const double d = sqrt(G[j * r + j]), rinv = 1.0 / d;
            ex.sync();
            for (int i = j + ex.lane(); i < r; i += Ex::NL) G[i * r + j] = (i == j) ? d : G[i * r + j] * rinv;
            ex.sync();
        }
        for (int pass = 0; pass < 2; pass++) {
            for (int k = 0; k < r; k++) {
                const double yk = u[k] / G[k * r + k];
                ex.sync();
                for (int i = k + ex.lane(); i < r; i += Ex::NL) { if (i == k) u[k] = yk; else u[i] -= G[i * r + k] * yk; }
                ex.sync();
            }
            for (int k = r - 1; k >= 0; k--) {
                const double xk = u[k] / G[k * r + k];
                ex.sync();
                for (int i = ex.lane(); i <= k; i += Ex::NL) { if (i == k) u[k] = xk; else u[i] -= G[k * r + i] * xk; }
                ex.sync();
            }
        }
        for (int i = ex.lane(); i < ka; i += Ex::NL) {
            double sacc = 0.0;
            for (int a = 0; a < r; a++) sacc += Lt[i * ka + a] * u[a];
            w.sv0[i] = sacc;
        }
        ex.sync();
        st.flops += 4.0 * ka * r * r;
        st.flags |= 16;
    }
    // consistency: S delta must reproduce rho (fails only when dependent active rows carry inconsistent right-hand sides)
    double worst = 0.0, scale = 0.0;
    for (int m = ex.lane(); m < ka; m += Ex::NL) {
        double sacc = 0.0;
        for (int k = 0; k < ka; k++) sacc += ((k <= m) ? Sm[m * ka + k] : Sm[k * ka + m]) * w.sv0[k];
        worst = fmax(worst, fabs(sacc - rhs0[m]));
        scale = fmax(scale, fabs(W[m * 31 + NMAIN]));      // |d_m|: rho is a difference of terms of this size
    }
    worst = ex.maxv(worst);
    scale = ex.maxv(scale);
    if (worst > 1.0e-9 * (scale + 1.0)) {
#ifdef WBC_EMU_DEBUG
        printf("schur inconsistent ka=%d r=%d worst=%.3e scale=%.3e\n", ka, r, worst, scale);
#endif
        return false;
    }
    double* nu0 = w.qrv;                    // save nu0 before zeroing
    for (int m = ex.lane(); m < ka; m += Ex::NL) nu0[m] = w.nulcest[act[m]];
    ex.sync();
    for (int i = ex.lane(); i < ktotal; i += Ex::NL) w.nulcest[i] = 0.0;
    ex.sync();
    for (int m = ex.lane(); m < ka; m += Ex::NL) w.nulcest[act[m]] = nu0[m] + w.sv0[m];
    ex.sync();
    return true;
}

// ------------------------------------------------------------------------------------------------
// generateexmodel (opt.cpp:41594-41740): extended box-QP in [x; slacks].  Upper triangle of exa.
template <class Ex>
WBC_HD void generate_ex_model(const Ex& ex, const Work& w, int nec, int nic, double rho, double& flops)
{
    const int n = NMAIN + nic, ld = n, kw = nec + nic;
    // quadratic term, columns < NMAIN: A + rho * C'C
    for (int i = 0; i < NMAIN; i++) {
        for (int j = i + ex.lane(); j < NMAIN; j += Ex::NL) {
            double s = 0.0;
            for (int r = 0; r < kw; r++) s += w.C[r * 31 + i] * w.C[r * 31 + j];
            w.exa[i * ld + j] = w.A[i * NMAIN + j] + rho * s;
        }
        for (int k = ex.lane(); k < nic; k += Ex::NL) w.exa[i * ld + NMAIN + k] = 0.0 + rho * w.C[(nec + k) * 31 + i];
    }
    for (int i = NMAIN; i < n; i++)
        for (int j = i + ex.lane(); j < n; j += Ex::NL) w.exa[i * ld + j] = (i == j) ? 0.0 + rho * 1.0 : 0.0;
    // linear term (41650-41657, 41734-41737): per element, rows in order, two updates per row
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v;
        if (i < NMAIN) {
            v = w.b[i];
            for (int r = 0; r < kw; r++) {
                double c = w.C[r * 31 + i];
                v += c * (-rho * w.C[r * 31 + NMAIN]);
                v += c * (-w.nulc[r]);
            }
        } else {
            int r = nec + (i - NMAIN);
            v = 0.0;
            v += 1.0 * (-rho * w.C[r * 31 + NMAIN]);
            v += 1.0 * (-w.nulc[r]);
        }
        w.exb[i] = v;
    }
    ex.sync();
    flops += (double)NMAIN * NMAIN * kw + 4.0 * NMAIN * kw;
}

// ------------------------------------------------------------------------------------------------
// The solver: Q (30x30 row-major, lower triangle used like minqpsetquadraticterm's default,
// opt.cpp:4962/18959), c (30), L (nrows x 31, first neq rows equalities, rest "<="), result x (30).
// ldq/ldl/... allow strided (SoA) inputs: element (i,j) of Q at Q[(i*30+j)*sq] etc.
template <class Ex>
WBC_HDN void solve_denseaul(const Ex& ex, const Work& w, const Settings& cfg, const double* Q, long sq,
                            const double* c, long sc, const double* L, long sl, int nrows, int neq,
                            double* x, long sx, Stats& st)
{
    const int nec = neq, nictotal = nrows - neq;
    st.termination = 0; st.ncholesky = 0; st.outer_its = 0; st.qqp_calls = 0; st.nicwork = 0;
    st.kkt_dim_max = 0; st.flags = 0; st.flops = 0.0;

    // ---- minqpoptimize: autodiag scale (opt.cpp:48146-48185)
    int bad = 0;
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) {
        double d = Q[(long)(i * NMAIN + i) * sq];
        if (d <= 0.0) bad = 1;
        w.s[i] = 1.0 / sqrt(d);
    }
    bad = ex.sumi(bad);
    ex.sync();
    if (bad) { st.termination = -9; return; }

    // ---- scaleshiftoriginalproblem (opt.cpp:42088-42339)
    for (int i = 0; i < NMAIN; i++)
        for (int j = i + ex.lane(); j < NMAIN; j += Ex::NL) {
            double v = Q[(long)(j * NMAIN + i) * sq] * w.s[i] * w.s[j];
            w.A[i * NMAIN + j] = v;
            w.A[j * NMAIN + i] = v;
        }
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) w.b[i] = c[(long)i * sc] * w.s[i];
    // constraint rows: lane-per-row scaling + normalisation (42219-42314)
    for (int r = ex.lane(); r < nrows; r += Ex::NL) {
        double vv = 0.0;
        for (int j = 0; j < NMAIN; j++) {
            double v = L[(long)(r * 31 + j) * sl] * w.s[j];
            w.C[r * 31 + j] = v;
            vv += v * v;
        }
        double rhs = L[(long)(r * 31 + NMAIN) * sl];
        vv = sqrt(vv);
        if (vv > 0.0) {
            vv = 1.0 / vv;
            for (int j = 0; j < NMAIN; j++) w.C[r * 31 + j] *= vv;
            rhs *= vv;
        }
        w.C[r * 31 + NMAIN] = rhs;
    }
    ex.sync();

    // ---- normalizequadraticterm (opt.cpp:42374-42446)
    double targetscale;
    {
        double an = 0.0;
        for (int i = ex.lane(); i < NMAIN * NMAIN; i += Ex::NL) an += w.A[i] * w.A[i];
        an = sqrt(ex.sum(an));
        double maxcac = 0.0;
        for (int r = ex.lane(); r < nrows; r += Ex::NL) {
            double v = 0.0;
            for (int j = 0; j < NMAIN; j++) {
                double t = 0.0;
                for (int k = 0; k < NMAIN; k++) t += w.C[r * 31 + k] * w.A[k * NMAIN + j];
                v += t * w.C[r * 31 + j];
            }
            maxcac = fmax(maxcac, fabs(v));
        }
        maxcac = ex.maxv(maxcac);
        targetscale = fmax(maxcac, an / NMAIN);
        if (targetscale == 0.0) targetscale = 1.0;
        double v = 1.0 / targetscale;
        ex.sync();
        for (int i = ex.lane(); i < NMAIN * NMAIN; i += Ex::NL) w.A[i] *= v;
        for (int i = ex.lane(); i < NMAIN; i += Ex::NL) w.b[i] *= v;
        ex.sync();
        st.flops += 2.0 * nrows * NMAIN * NMAIN;
    }

    // ---- selectinitialworkingset (opt.cpp:42474-42523)
    int nicwork = 0;
    bool allowevict = true;
    {
        for (int i = 0; i < NMAIN; i++)
            for (int j = i + ex.lane(); j < NMAIN; j += Ex::NL) w.z[i * NMAIN + j] = w.A[i * NMAIN + j];
        ex.sync();
        if (!cholesky_upper(ex, w.z, NMAIN, NMAIN)) { nicwork = nictotal; allowevict = false; st.flags |= 1; }
        st.flops += 9000.0;
    }
    for (int i = ex.lane(); i < nictotal; i += Ex::NL) w.nicnact[i] = (i < nicwork) ? 1 : 0;
    for (int i = ex.lane(); i < nrows; i += Ex::NL) w.nulc[i] = 0.0;
    for (int i = ex.lane(); i < NMAIN + nictotal; i += Ex::NL) w.exxc[i] = 0.0;
    ex.sync();

    double rho = cfg.rho, epsx = cfg.epsx;
    if (epsx <= 0.0) epsx = 1.0e-9;
    const double maxrho = 1.0e12, requestedfeasdecrease = 0.33;
    int goodcounter = 0, stagnationcounter = 0;
    double feaserr = 1.7976931348623157e308;   // ae_maxrealnumber
    for (int outeridx = 0; outeridx < cfg.outerits; outeridx++) {
        st.outer_its++;
        bool extended;
        do {
            const int nwork = NMAIN + nicwork;
            generate_ex_model(ex, w, nec, nicwork, rho, st.flops);
            int term = qqp_optimize(ex, w, nwork, 0.01 * epsx, 50, st.ncholesky, st.flops);
            st.qqp_calls++;
            if (term == -4) st.flags |= 4;
            // violations of all inequality rows w.r.t. the main variables only (41330-41335)
            for (int i = ex.lane(); i < nictotal; i += Ex::NL) {
                const double* row = &w.C[(nec + i) * 31];
                double v = 0.0;
                for (int j = 0; j < NMAIN; j++) v += row[j] * w.exxc[j];
                w.nicerr[i] = v - row[NMAIN];
            }
            ex.sync();
            st.flops += 2.0 * nictotal * NMAIN;
            // working-set expansion (41350-41390): literal sequential selection, executed redundantly
            extended = false;
            int added = 0;
            while ((double)added < 1 + 0.20 * NMAIN && nicwork < nictotal) {
                int k = nicwork;
                for (int j = nicwork; j < nictotal; j++)
                    if (w.nicerr[j] > w.nicerr[k]) k = j;
                if (!(w.nicerr[k] > 0.0)) break;
                ex.sync();
                // swap rows nec+nicwork <-> nec+k of C, and the per-constraint bookkeeping
                if (k != nicwork) {
                    for (int j = ex.lane(); j < 31; j += Ex::NL) {
                        double t = w.C[(nec + nicwork) * 31 + j];
                        w.C[(nec + nicwork) * 31 + j] = w.C[(nec + k) * 31 + j];
                        w.C[(nec + k) * 31 + j] = t;
                    }
                }
                if (ex.lane() == 0) {
                    double t = w.nicerr[nicwork]; w.nicerr[nicwork] = w.nicerr[k]; w.nicerr[k] = t;
                    int ti = w.nicnact[nicwork]; w.nicnact[nicwork] = w.nicnact[k]; w.nicnact[k] = ti;
                    w.exxc[NMAIN + nicwork] = 0.0;
                    w.nulc[nec + nicwork] = 0.0;
                    w.nicnact[nicwork] = w.nicnact[nicwork] + 1;
                }
                ex.sync();
                nicwork++; added++;
                extended = true;
            }
            // working-set eviction (41400-41418)
            if (allowevict) {
                for (int k = nicwork - 1; k >= 0; k--) {
                    if (w.nicerr[k] < -0.01 && w.nicnact[k] <= 1) {
                        ex.sync();
                        const int last = nicwork - 1;
                        if (k != last) {
                            for (int j = ex.lane(); j < 31; j += Ex::NL) {
                                double t = w.C[(nec + last) * 31 + j];
                                w.C[(nec + last) * 31 + j] = w.C[(nec + k) * 31 + j];
                                w.C[(nec + k) * 31 + j] = t;
                            }
                        }
                        if (ex.lane() == 0) {
                            double t = w.nicerr[last]; w.nicerr[last] = w.nicerr[k]; w.nicerr[k] = t;
                            int ti = w.nicnact[last]; w.nicnact[last] = w.nicnact[k]; w.nicnact[k] = ti;
                            t = w.exxc[NMAIN + last]; w.exxc[NMAIN + last] = w.exxc[NMAIN + k]; w.exxc[NMAIN + k] = t;
                            t = w.nulc[nec + last]; w.nulc[nec + last] = w.nulc[nec + k]; w.nulc[nec + k] = t;
                        }
                        ex.sync();
                        nicwork--;
                    }
                }
            }
        } while (extended);

        const int kwork = nec + nicwork;
        // multiplier estimate (41438-41439)
        for (int i = ex.lane(); i < kwork; i += Ex::NL) w.nulcest[i] = w.nulc[i];
        ex.sync();
        if (!(cfg.kkt_mode == 1 && update_lagrange_multipliers_schur(ex, w, nec, nicwork, st, cfg.kkt_pivtol))) {
            if (cfg.kkt_mode == 1) st.flags |= 8;
            update_lagrange_multipliers(ex, w, nec, nicwork, st);
        }
        // feasibility error and multiplier update (41444-41476): lane-per-row, summed in row order
        double feaserrprev = feaserr;
        double fe = 0.0;
        for (int i = ex.lane(); i < kwork; i += Ex::NL) {
            const double* row = &w.C[i * 31];
            double v = 0.0, vv = 0.0;
            for (int j = 0; j < NMAIN; j++) { v += row[j] * w.exxc[j]; vv += row[j] * row[j]; }
            if (i >= nec) { v += w.exxc[NMAIN + (i - nec)]; vv += 1.0; }
            v -= row[NMAIN];
            if (vv == 0.0) vv = 1.0;
            v = v / sqrt(vv);
            fe += v * v;
            w.nulc[i] = w.nulcest[i];
        }
        feaserr = sqrt(ex.sum(fe));
        ex.sync();
        st.flops += 4.0 * kwork * NMAIN;
        if (feaserr < epsx) goodcounter++; else goodcounter = 0;
        if (feaserr > feaserrprev * requestedfeasdecrease) stagnationcounter++; else stagnationcounter = 0;
        if (goodcounter >= 2) break;
        if (stagnationcounter >= 2) rho = fmin(rho * 10.0, maxrho);
        else rho = fmin(rho * 1.41, maxrho);
    }
    st.nicwork = nicwork;
    // unscale (41548-41583): x = s * xc  (+ origin 0); no box constraints on x
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) x[(long)i * sx] = w.s[i] * w.exxc[i] + 0.0;
    ex.sync();
    st.termination = 2;
}

}  // namespace wbcqp
