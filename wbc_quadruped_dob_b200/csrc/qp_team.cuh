// Batched dense QP solver for the WBC ground-reaction-force QP: a decision-for-decision restatement
// of the algorithm the reference controller runs through ALGLIB 3.16.0
//   minqpoptimize (opt.cpp:48020) -> qpdenseauloptimize (41087) -> qqpoptimize (29675)
// with the reference's settings (lopt.cpp:91-106: autodiag scaling, DENSE-AUL epsx=1e-2, rho=1e4,
// 5 outer iterations, cold start, no box constraints on x).  "opt.cpp" = the reference's
// dogbot_controller/src/alglib/optimization.cpp, "linalg.cpp" likewise.
//
// Execution model: ONE TEAM (a whole CTA of NL threads, NL = 64..256) PER QP, everything hot in
// shared memory.  Every routine is written against an executor `Ex` (lane/warp ids, team barrier,
// team all-reduce, warp shuffles).  All control flow depends only on values that are bit-identical in
// every thread (all-reduces, or shared memory read after a barrier), so the team never diverges
// around a barrier.  Matrix work (symmetric mat-vec, Cholesky, model generation) uses every thread;
// short serial recurrences (triangular solves, Givens updates, working-set bookkeeping) run on the
// team's first warp out of registers with shuffles only.  `HostEx` (one lane) lets the same source
// be compiled by g++ for CPU-side unit tests (tests/host_emu); the shipped library only
// instantiates `TeamEx`.
//
// Specialisation relative to generic ALGLIB (all other cases cannot occur on this path):
//   * dense A, no sparse constraints, x unbounded, start point 0, origin 0;
//   * hence in QQP the only bounds are "slack >= 0" on variables i >= NMAIN.
//
// Storage tricks (shared memory is the occupancy limiter):
//   * one n x n array S (leading dimension LDS, odd) holds the QQP quadratic term E in its upper
//     triangle (E is symmetric; the diagonal belongs to E) and the TRANSPOSED Cholesky factor in its
//     strict lower triangle (U[k][c] at S[c*ld+k]), the factor's diagonal lives in a vector.  With an
//     odd leading dimension both row walks and column walks are bank-conflict free;
//   * the same array is the workspace of the multiplier update, when QQP is not running;
//   * instances whose working set outgrows NCAP switch S and the QQP vectors to a global-memory
//     spill copy (same code, different pointers).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WBC_HD __host__ __device__ __forceinline__
#define WBC_HDN __host__ __device__
#define WBC_HDNI __host__ __device__ __noinline__
#else
#define WBC_HD inline
#define WBC_HDN
#define WBC_HDNI
#endif

namespace wbcqp {

constexpr int NMAIN = 30;              // decision variables (main.cpp:266 OPT(30,86,82))
constexpr int MAXK = 88;               // >= 86 constraint rows (stance), 82 (swing)
constexpr int MAXNIC = 72;             // >= 68 / 70 inequality rows
constexpr int MAXNT = NMAIN + MAXNIC;  // extended variable count upper bound
constexpr int NCAP = 48;               // largest extended dimension held in shared memory
constexpr int LDS = 49;                // leading dimension of the shared E/Z array
constexpr int LDG = 103;               // leading dimension of the global spill copy
constexpr int LDA = 31;                // SA: A in the upper triangle, its Cholesky factor (transposed) below
constexpr int KACAP = 33;              // largest active set the reduced multiplier update handles
constexpr int LDM = 35;                // leading dimension of its Schur-complement array
constexpr int NVEC = 16;               // QQP vectors
constexpr int VLG = 104;               // vector length of the global spill copy
constexpr int S_DOUBLES = NCAP * LDS;  // 2352
constexpr double MACHEPS = 5.0e-16;    // ae_machineepsilon (ap.cpp: 5E-16, NOT DBL_EPSILON)
constexpr double BIGSTEP = 1.0e50;     // opt.cpp:27458

struct Settings {
    double epsx = 1.0e-2;   // lopt.cpp:101
    double rho = 1.0e4;
    int outerits = 5;
    int kkt_mode = 1;       // 1 = reduced multiplier update with the literal form as fallback; 0 = literal only
    double kkt_pivtol = 1.0e-5;
};

struct Stats {
    int termination;    // 2 = ok (opt.cpp:41583); -9 non-positive diagonal (48178-48181)
    int ncholesky;      // rep.ncholesky (opt.cpp:41325)
    int outer_its;      // outer AUL iterations executed
    int qqp_calls;      // inner QQP solves
    int nicwork;        // final working-set size
    int kkt_dim_max;    // largest (N+K) of the multiplier update
    int flags;          // bit0: A not PD (42500); bit2: QQP -4; bit3: literal multiplier update used;
                        // bit4: rank-deficient active set (least-norm branch); bit5: spilled to global memory
    double flops;       // instrumented algorithmic flop count (DESIGN.md "work per solve")
};

// Per-team storage.  "sh" = shared memory on the device, "gl" = the team's global scratch.
struct Work {
    double* SA;        // gl [30*31]  upper+diag: scaled A; strict lower: U' of A = U'U (U[k][c] at SA[c*31+k])
    double* larinv;    // sh [32]     1 / U_kk of that factor
    double* ladiag;    // sh [32]     U_kk
    double* C;         // sh [MAXK*31] scaled, normalised constraint rows (physically swapped like opt.cpp:41372)
    double* b;         // sh [32]
    double* s;         // sh [32]     variable scales
    double* nicerr;    // sh [MAXNIC]
    double* nulc;      // sh [MAXK]
    double* nulcest;   // sh [MAXK]
    double* exxc;      // sh [MAXNT]
    double* exb;       // sh [MAXNT]
    double* xs;        // sh [32]     solution
    int* nicnact;      // sh [MAXNIC]
    int* cstatus;      // sh [MAXNT]
    int* isfree;       // sh [MAXNT]
    int* iscr;         // sh [8]
    double* Ssh;       // sh [NCAP*LDS]
    double* vsh;       // sh [NVEC*NCAP]
    double* Sgl;       // gl [MAXNT*LDG]
    double* vgl;       // gl [NVEC*VLG]
    double* kkt;       // gl literal multiplier update, see kkt_doubles()
    double* qrv;       // gl [2*(MAXNT+MAXK)+2]
    double* sv0;       // gl [MAXNT+MAXK]
};

// The QQP view of the storage (shared or spilled).
struct QV {
    double* S;
    int ld;
    double *zd, *zrinv, *xc, *xp, *xf, *gc, *cgc, *cgp, *dc, *dp, *t0, *t1, *t2, *t3, *regdiag, *bufr;
};
WBC_HD QV make_qv(const Work& w, bool spill)
{
    QV q;
    double* v = spill ? w.vgl : w.vsh;
    const int vl = spill ? VLG : NCAP;
    q.S = spill ? w.Sgl : w.Ssh;
    q.ld = spill ? LDG : LDS;
    q.zd = v; q.zrinv = v + vl; q.xc = v + 2 * vl; q.xp = v + 3 * vl; q.xf = v + 4 * vl; q.gc = v + 5 * vl;
    q.cgc = v + 6 * vl; q.cgp = v + 7 * vl; q.dc = v + 8 * vl; q.dp = v + 9 * vl; q.t0 = v + 10 * vl; q.t1 = v + 11 * vl;
    q.t2 = v + 12 * vl; q.t3 = v + 13 * vl; q.regdiag = v + 14 * vl; q.bufr = v + 15 * vl;
    return q;
}

// ------------------------------------------------------------------------------------------------
// executors
struct HostEx {
    static constexpr int NL = 1, WL = 1, NW = 1;
    WBC_HD int lane() const { return 0; }
    WBC_HD int warp() const { return 0; }
    WBC_HD int wlane() const { return 0; }
    WBC_HD void sync() const {}
    WBC_HD void wsync() const {}
    WBC_HD double shfl(double v, int) const { return v; }
    WBC_HD int shfli(int v, int) const { return v; }
    WBC_HD double shfl_xor(double v, int) const { return v; }
    WBC_HD int shfl_xori(int v, int) const { return v; }
    WBC_HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
    template <int NS, int NM>
    WBC_HD void allred(double*, double*) const {}
};

#if defined(__CUDACC__)
// A whole CTA of T threads.  `red` points at 2*(T/32)*8 doubles of shared memory.
template <int T>
struct TeamEx {
    static constexpr int NL = T, WL = 32, NW = T / 32;
    double* red;
    mutable int par;
    __device__ __forceinline__ int lane() const { return threadIdx.x; }
    __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
    __device__ __forceinline__ int wlane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ void sync() const { if (T == 32) __syncwarp(); else __syncthreads(); }
    __device__ __forceinline__ void wsync() const { __syncwarp(); }
    __device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ int shfli(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ double shfl_xor(double v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    __device__ __forceinline__ int shfl_xori(int v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    // all-reduce NS sums and NM maxima at once; every thread ends with bit-identical results (butterflies are
    // commutative at every level; the cross-warp combination runs in warp order in every thread).  One barrier.
    template <int NS, int NM>
    __device__ __forceinline__ void allred(double* s, double* m) const
    {
        static_assert(NS + NM <= 8, "reduction slots");
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < NS; k++) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
#pragma unroll
            for (int k = 0; k < NM; k++) m[k] = fmax(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
        }
        if (T == 32) { __syncwarp(); return; }
        double* buf = red + par * (NW * 8);
        par ^= 1;
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < NS; k++) buf[(threadIdx.x >> 5) * 8 + k] = s[k];
#pragma unroll
            for (int k = 0; k < NM; k++) buf[(threadIdx.x >> 5) * 8 + NS + k] = m[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NS; k++) s[k] = buf[k];
#pragma unroll
        for (int k = 0; k < NM; k++) m[k] = buf[NS + k];
#pragma unroll
        for (int wp = 1; wp < NW; wp++) {
#pragma unroll
            for (int k = 0; k < NS; k++) s[k] += buf[wp * 8 + k];
#pragma unroll
            for (int k = 0; k < NM; k++) m[k] = fmax(m[k], buf[wp * 8 + NS + k]);
        }
    }
};
#endif

template <class Ex> WBC_HD double allsum1(const Ex& ex, double a) { double s[1] = {a}; ex.template allred<1, 0>(s, s); return s[0]; }
template <class Ex> WBC_HD double allmax1(const Ex& ex, double a) { double m[1] = {a}; ex.template allred<0, 1>(m, m); return m[0]; }

// ------------------------------------------------------------------------------------------------
// scalar helpers (verbatim semantics of the ALGLIB routines named)
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
// symmetric matrix in "upper" storage: element (i,j) at S[min*ld + max]
WBC_HD double sym_at(const double* S, int ld, int i, int j) { return (i <= j) ? S[i * ld + j] : S[j * ld + i]; }

// Row-to-thread map of the symmetric products: NL >= 128 -> two threads per row (columns split at `h`),
// otherwise one thread per row.  With LDS odd and h = 24 (mod 16 = 8) the two column segments of a
// half-warp fall on disjoint banks.
template <class Ex>
struct RowMap {
    static constexpr int SEG = (Ex::NL >= 128) ? 2 : 1;
    static constexpr int ROWS = Ex::NL / SEG;
    int s, ioff, j0, j1;
    WBC_HD RowMap(const Ex& ex, int n)
    {
        s = ex.lane() % SEG;
        ioff = ex.lane() / SEG;
        const int h = (SEG == 1) ? n : ((n > 24 && n <= 48) ? 24 : (n + 1) / 2);
        j0 = (s == 0) ? 0 : h;
        j1 = (SEG == 1 || s == 1) ? n : h;
    }
};

// y = E x (+ b): E symmetric, upper storage.  NV right-hand sides share one pass over E.
template <int NV, class Ex>
WBC_HD void symv_multi(const Ex& ex, const double* S, int ld, int n, const double* const* x, const double* b, double* const* y)
{
    const RowMap<Ex> rm(ex, n);
    for (int base = 0; base < n; base += RowMap<Ex>::ROWS) {
        const int i = base + rm.ioff;
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) acc[k] = 0.0;
        if (i < n) {
            for (int j = rm.j0; j < rm.j1; j++) {
                const double e = sym_at(S, ld, i, j);
#pragma unroll
                for (int k = 0; k < NV; k++) acc[k] += e * x[k][j];
            }
        }
        if (RowMap<Ex>::SEG == 2) {
#pragma unroll
            for (int k = 0; k < NV; k++) acc[k] += ex.shfl_xor(acc[k], 1);
        }
        if (i < n && rm.s == 0) {
#pragma unroll
            for (int k = 0; k < NV; k++) y[k][i] = b ? acc[k] + b[i] : acc[k];
        }
    }
    ex.sync();
}
template <class Ex>
WBC_HDNI void symv(const Ex& ex, const double* S, int ld, int n, const double* x, const double* b, double* y)
{
    // single right-hand side with two accumulators for instruction-level parallelism
    const RowMap<Ex> rm(ex, n);
    for (int base = 0; base < n; base += RowMap<Ex>::ROWS) {
        const int i = base + rm.ioff;
        double a0 = 0.0, a1 = 0.0;
        if (i < n) {
            int j = rm.j0;
            for (; j + 1 < rm.j1; j += 2) {
                a0 += sym_at(S, ld, i, j) * x[j];
                a1 += sym_at(S, ld, i, j + 1) * x[j + 1];
            }
            if (j < rm.j1) a0 += sym_at(S, ld, i, j) * x[j];
        }
        double acc = a0 + a1;
        if (RowMap<Ex>::SEG == 2) acc += ex.shfl_xor(acc, 1);
        if (i < n && rm.s == 0) y[i] = b ? acc + b[i] : acc;
    }
    ex.sync();
}

// ------------------------------------------------------------------------------------------------
// Right-looking Cholesky A = U'U on the TRANSPOSED-LOWER storage: on entry Z[c*ld+k] (k<c) = a_kc and
// zd[k] = a_kk; on exit Z[c*ld+k] = U[k][c], zd[k] = U[k][k], zrinv[k] = 1/U[k][k].  One team barrier
// per column; rows are scaled lazily at the end (the trailing update uses a_jr*a_jc/p_j).
// SKIP = false: returns false on a non-positive pivot (linalg.cpp:29204-29235).
// SKIP = true (positive semi-definite input): a pivot below pivtol * (original diagonal) marks a
//   dependent row: dep[j] = 1, its row of U is zero, zd = zrinv = 0; *ambiguous is set when such a pivot
//   is not clearly rounding noise (above 1e-3 * pivtol).
template <bool SKIP, class Ex>
WBC_HDNI bool chol_lowerT(const Ex& ex, double* Z, int ld, int n, double* zd, double* zrinv, double* pinv, int* dep, double pivtol,
                        bool* ambiguous)
{
    constexpr int CW = (Ex::NL >= 64) ? 64 : Ex::NL;
    constexpr int NG = Ex::NL / CW;
    const int cs = ex.lane() % CW, g = ex.lane() / CW;
    // zrinv holds the original diagonal during the factorisation (SKIP), pinv[j] = 1/pivot_j
    for (int k = ex.lane(); k < n; k += Ex::NL) {
        if (SKIP) { zrinv[k] = zd[k]; dep[k] = 0; }
        if (k == 0) pinv[0] = 1.0 / zd[0];
    }
    ex.sync();
    bool amb = false;
    for (int j = 0; j < n; j++) {
        const double p = zd[j];
        if (SKIP) {
            const double d0 = zrinv[j];
            if (!(p > pivtol * d0)) {
                if (p > 1.0e-3 * pivtol * d0) amb = true;
                if (ex.lane() == 0) dep[j] = 1;
                // the next pivot's reciprocal was not produced by a trailing update
                if (j + 1 < n && ex.lane() == 0) pinv[j + 1] = 1.0 / zd[j + 1];
                ex.sync();
                continue;
            }
        } else {
            if (!(p > 0.0)) return false;
        }
        const double inv = pinv[j];
        for (int c = j + 1 + cs; c < n; c += CW) {
            const double ajc = Z[c * ld + j];
            const double t = ajc * inv;
            for (int r = j + 1 + g; r <= c; r += NG) {
                if (r < c) Z[c * ld + r] -= Z[r * ld + j] * t;
                else {
                    const double dn = zd[c] - ajc * t;
                    zd[c] = dn;
                    if (c == j + 1) pinv[c] = 1.0 / dn;
                }
            }
        }
        ex.sync();
    }
    for (int k = ex.lane(); k < n; k += Ex::NL) {
        if (SKIP && dep[k]) { zd[k] = 0.0; pinv[k] = 0.0; }
        else { const double d = sqrt(zd[k]); zd[k] = d; pinv[k] = 1.0 / d; }
    }
    ex.sync();
    for (int c = 1 + ex.warp(); c < n; c += Ex::NW)
        for (int k = ex.wlane(); k < c; k += Ex::WL) Z[c * ld + k] *= pinv[k];
    for (int k = ex.lane(); k < n; k += Ex::NL) zrinv[k] = pinv[k];
    ex.sync();
    if (ambiguous) *ambiguous = amb;
    return true;
}

// Solve U'U x = rhs in place with the factor above; x in shared/global memory, n <= 32*NR.  Runs on the
// team's first warp out of registers (lane l holds x[l], x[l+32], ...), column-oriented both ways, no team
// barrier inside.  Dependent pivots (zrinv = 0) yield a zero component.
template <int NR, class Ex>
WBC_HDNI void tri_solve_regs(const Ex& ex, const double* Z, int ld, int n, const double* zrinv, double* x, bool forward, bool backward)
{
    if (ex.warp() == 0) {
        if (Ex::WL == 1) {
            if (forward)
                for (int k = 0; k < n; k++) {
                    const double yk = x[k] * zrinv[k];
                    x[k] = yk;
                    for (int i = k + 1; i < n; i++) x[i] -= Z[i * ld + k] * yk;
                }
            if (backward)
                for (int k = n - 1; k >= 0; k--) {
                    const double xk = x[k] * zrinv[k];
                    x[k] = xk;
                    for (int i = 0; i < k; i++) x[i] -= Z[k * ld + i] * xk;
                }
        } else {
            const int l = ex.wlane();
            double xr[NR];
#pragma unroll
            for (int s = 0; s < NR; s++) xr[s] = (l + 32 * s < n) ? x[l + 32 * s] : 0.0;
            if (forward)
                for (int k = 0; k < n; k++) {
                    double v = xr[0];
#pragma unroll
                    for (int s = 1; s < NR; s++) if ((k >> 5) == s) v = xr[s];
                    const double yk = ex.shfl(v, k & 31) * zrinv[k];
#pragma unroll
                    for (int s = 0; s < NR; s++) {
                        const int i = l + 32 * s;
                        if (i == k) xr[s] = yk;
                        else if (i > k && i < n) xr[s] -= Z[i * ld + k] * yk;
                    }
                }
            if (backward)
                for (int k = n - 1; k >= 0; k--) {
                    double v = xr[0];
#pragma unroll
                    for (int s = 1; s < NR; s++) if ((k >> 5) == s) v = xr[s];
                    const double xk = ex.shfl(v, k & 31) * zrinv[k];
#pragma unroll
                    for (int s = 0; s < NR; s++) {
                        const int i = l + 32 * s;
                        if (i == k) xr[s] = xk;
                        else if (i < k) xr[s] -= Z[k * ld + i] * xk;
                    }
                }
#pragma unroll
            for (int s = 0; s < NR; s++) if (l + 32 * s < n) x[l + 32 * s] = xr[s];
        }
    }
    ex.sync();
}
template <class Ex>
WBC_HD void tri_solve(const Ex& ex, const double* Z, int ld, int n, const double* zrinv, double* x, bool forward = true, bool backward = true)
{
    if (n <= 64) tri_solve_regs<2>(ex, Z, ld, n, zrinv, x, forward, backward);
    else tri_solve_regs<4>(ex, Z, ld, n, zrinv, x, forward, backward);
}

// ------------------------------------------------------------------------------------------------
// QQP (opt.cpp:29675-30566) specialised to: dense A (akind 2, upper), unit scales, zero origin,
// variables [0,NMAIN) free, variables [NMAIN,n) bounded below by 0.
struct QqpState {
    int n;
    double absasum, absasum2, mb;
    int nfree, cnmodelage;
    int ncholesky;
};

// sasexploredirection (opt.cpp:27433-27528), box-only.  Sequential scan kept literal (first strict
// improvement wins); executed redundantly by every thread over the <= MAXNIC bounded variables.
WBC_HD void sas_explore_direction(const int* cstatus, const double* xc, int n, const double* d, double& stpmax, int& cidx, double& cval)
{
    stpmax = BIGSTEP; cidx = -1; cval = 0.0;
    for (int i = NMAIN; i < n; i++) {
        const double di = d[i];
        if (di < 0.0 && cstatus[i] <= 0) {
            double prev = stpmax;
            stpmax = safeminposrv(xc[i] - 0.0, -di, stpmax);
            if (stpmax < prev) { cidx = i; cval = 0.0; }
        }
    }
}

// qqpsolver_findbeststepandmove (opt.cpp:30882-31003) fused with sasmoveto (27574-27723, box-only) and the
// "previous direction" bookkeeping of the caller.  The candidate steps {stp, addsteps[k] > stp} are evaluated
// together: their projected points share one pass over E (qqpsolver_projectedtargetfunction, 30582-30652).
template <class Ex>
WBC_HDNI void qqp_find_best_step_and_move(const Ex& ex, const Work& w, const QV& q, const QqpState& st, const double* d, double stp,
                                        bool needact, int cidx, double cval, const double* addsteps, int addcnt, double& flops)
{
    const int n = st.n;
    double stpbest = stp;
    if (addcnt > 0) {
        double steps[4] = {stp, stp, stp, stp};
        for (int k = 0; k < addcnt; k++) steps[1 + k] = addsteps[k];
        double* t[4] = {q.t0, q.t1, q.t2, q.t3};
        for (int i = ex.lane(); i < n; i += Ex::NL) {
            const double xi = q.xc[i], di = d[i];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                double v = (steps[k] != 0.0) ? xi + steps[k] * di : xi;
                if (i >= NMAIN && v < 0.0) v = 0.0;
                t[k][i] = v;
            }
        }
        ex.sync();
        // f_k = exb . t_k + 0.5 * t_k . (E t_k): partial sums per row
        double lin[4] = {0, 0, 0, 0}, quad[4] = {0, 0, 0, 0};
        {
            const RowMap<Ex> rm(ex, n);
            for (int base = 0; base < n; base += RowMap<Ex>::ROWS) {
                const int i = base + rm.ioff;
                if (i < n) {
                    double acc[4] = {0, 0, 0, 0};
                    for (int j = rm.j0; j < rm.j1; j++) {
                        const double e = sym_at(q.S, q.ld, i, j);
#pragma unroll
                        for (int k = 0; k < 4; k++) acc[k] += e * t[k][j];
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const double ti = t[k][i];
                        quad[k] += ti * acc[k];
                        if (rm.s == 0) lin[k] += w.exb[i] * ti;
                    }
                }
            }
        }
        double red8[8] = {lin[0], lin[1], lin[2], lin[3], quad[0], quad[1], quad[2], quad[3]};
        ex.template allred<8, 0>(red8, red8);
        double fbest = red8[0] + 0.5 * red8[4];
        for (int k = 0; k < addcnt; k++) {
            if (addsteps[k] > stp) {
                const double fcand = red8[1 + k] + 0.5 * red8[5 + k];
                if (fcand < fbest) { fbest = fcand; stpbest = addsteps[k]; }
            }
        }
        flops += (1 + addcnt) * 2.0 * n * n;
    }
    else ex.sync();      // the caller's scan of cstatus / xc (sasexploredirection) is complete in every thread
    // move (sasmoveto)
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        const double old = q.xc[i];
        double v = old + stpbest * d[i];
        if (i >= NMAIN && v < 0.0) v = 0.0;
        if (needact && i == cidx) { v = cval; w.cstatus[i] = 1; }
        if (i >= NMAIN && v <= 0.0 && v != old) { v = 0.0; w.cstatus[i] = 1; }
        q.xc[i] = v;
    }
    ex.sync();
}

// qqpsolver_cnewtonbuild (opt.cpp:31058-31201).  The factor is produced directly in the "scattered"
// n x n layout ALGLIB ends with (identity rows for fixed variables): factoring the masked matrix
// gives identical entries because the extra terms are exact zeros.
template <class Ex>
WBC_HDNI bool qqp_cnewton_build(const Ex& ex, const Work& w, const QV& q, QqpState& st)
{
    const int n = st.n, ld = q.ld;
    st.cnmodelage = 0;
    double nf = 0.0;
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        const int fr = !(i >= NMAIN && q.xc[i] == 0.0);
        w.isfree[i] = fr;
        nf += fr;
    }
    nf = allsum1(ex, nf);       // barrier inside: isfree visible
    st.nfree = (int)nf;
    if (st.nfree == 0) return false;
    // regdiag[i] = 1e-9 * sum_j |A_ff[i][j]| over free j (31150-31167)
    {
        const RowMap<Ex> rm(ex, n);
        for (int base = 0; base < n; base += RowMap<Ex>::ROWS) {
            const int i = base + rm.ioff;
            double v = 0.0;
            if (i < n && w.isfree[i])
                for (int j = rm.j0; j < rm.j1; j++)
                    if (w.isfree[j]) v += fabs(sym_at(q.S, ld, i, j));
            if (RowMap<Ex>::SEG == 2) v += ex.shfl_xor(v, 1);
            if (i < n && rm.s == 0) {
                if (w.isfree[i]) {
                    if (v == 0.0) v = 1.0;
                    q.zd[i] = q.S[i * ld + i] + 1.0e-9 * v;
                } else q.zd[i] = 1.0;
            }
        }
    }
    // masked copy of the strict upper triangle into the transposed-lower factor storage
    for (int c = 1 + ex.warp(); c < n; c += Ex::NW) {
        const int fc = w.isfree[c];
        for (int k = ex.wlane(); k < c; k += Ex::WL) q.S[c * ld + k] = (fc && w.isfree[k]) ? q.S[k * ld + c] : 0.0;
    }
    ex.sync();
    st.ncholesky++;
    return chol_lowerT<false>(ex, q.S, ld, n, q.zd, q.zrinv, q.regdiag, (int*)nullptr, 0.0, (bool*)nullptr);
}

// qqpsolver_cnewtonupdate (opt.cpp:31314-31426) + spdmatrixcholeskyupdatefixbuf (linalg.cpp:27657-27819, upper):
// fixing variable k = removing row/column k of the factor with a sweep of Givens rotations.  Serial in the row
// index; runs on the first warp with the rotated row held in registers.
template <int NR, class Ex>
WBC_HDNI void givens_fix_regs(const Ex& ex, const QV& q, int n, int k)
{
    const int ld = q.ld;
    double* Z = q.S;
    if (ex.warp() == 0) {
        if (Ex::WL == 1) {
            for (int j = k + 1; j < n; j++) { q.bufr[j] = Z[j * ld + k]; Z[j * ld + k] = 0.0; }
            for (int i = 0; i < k; i++) Z[k * ld + i] = 0.0;
            q.zd[k] = 1.0; q.zrinv[k] = 1.0;
            for (int i = k + 1; i < n; i++) {
                const double bi = q.bufr[i];
                if (bi != 0.0) {
                    double cs, sn, r;
                    generaterotation(q.zd[i], bi, cs, sn, r);
                    q.zd[i] = r; q.zrinv[i] = 1.0 / r; q.bufr[i] = 0.0;
                    for (int j = i + 1; j < n; j++) {
                        const double v = Z[j * ld + i], vv = q.bufr[j];
                        Z[j * ld + i] = cs * v + sn * vv;
                        q.bufr[j] = -sn * v + cs * vv;
                    }
                }
            }
        } else {
            const int l = ex.wlane();
            double br[NR];
#pragma unroll
            for (int s = 0; s < NR; s++) {
                const int j = l + 32 * s;
                br[s] = 0.0;
                if (j > k && j < n) { br[s] = Z[j * ld + k]; Z[j * ld + k] = 0.0; }
                if (j < k) Z[k * ld + j] = 0.0;
            }
            if (l == 0) { q.zd[k] = 1.0; q.zrinv[k] = 1.0; }
            ex.wsync();
            for (int i = k + 1; i < n; i++) {
                double v = br[0];
#pragma unroll
                for (int s = 1; s < NR; s++) if ((i >> 5) == s) v = br[s];
                const double bi = ex.shfl(v, i & 31);
                if (bi != 0.0) {
                    double cs, sn, r;
                    generaterotation(q.zd[i], bi, cs, sn, r);
                    ex.wsync();      // every lane has read zd[i]
#pragma unroll
                    for (int s = 0; s < NR; s++) {
                        const int j = l + 32 * s;
                        if (j == i) { q.zd[i] = r; q.zrinv[i] = 1.0 / r; br[s] = 0.0; }
                        else if (j > i && j < n) {
                            const double zv = Z[j * ld + i], vv = br[s];
                            Z[j * ld + i] = cs * zv + sn * vv;
                            br[s] = -sn * zv + cs * vv;
                        }
                    }
                }
            }
        }
    }
}
template <class Ex>
WBC_HDNI bool qqp_cnewton_update(const Ex& ex, const Work& w, const QV& q, QqpState& st, int cnmaxupdates)
{
    const int n = st.n;
    double ntf = 0.0;
    for (int i = ex.lane(); i < n; i += Ex::NL)
        if (w.isfree[i] && i >= NMAIN && q.xc[i] == 0.0) ntf += 1.0;
    const int ntofix = (int)allsum1(ex, ntf);
    if (ntofix == 0 || ntofix == st.nfree) return false;
    if (st.cnmodelage + ntofix > cnmaxupdates) return false;
    for (int k = NMAIN; k < n; k++) {
        if (!(w.isfree[k] && q.xc[k] == 0.0)) continue;
        if (n <= 64) givens_fix_regs<2>(ex, q, n, k); else givens_fix_regs<4>(ex, q, n, k);
        ex.sync();
        if (ex.lane() == 0) w.isfree[k] = 0;
    }
    ex.sync();
    st.nfree -= ntofix;
    st.cnmodelage += ntofix;
    return true;
}

// One QQP solve from the point w.exxc (in/out) on the model (E in the upper triangle of q.S, w.exb).
// Returns the QQP termination type.
template <class Ex>
WBC_HDNI int qqp_optimize(const Ex& ex, const Work& w, const QV& q, int n, double epsx, int maxouterits, int& ncholesky, double& flops)
{
    QqpState st;
    st.n = n; st.ncholesky = 0; st.nfree = 0; st.cnmodelage = 0;
    const int ld = q.ld;
    // settings: qqploaddefaults (opt.cpp:29533-29547) + overrides (41318-41323)
    const int cgminits = 5;
    int cgmaxits = (int)lround(1 + 0.33 * n);
    if (cgmaxits < cgminits) cgmaxits = cgminits;
    const int cnmaxupdates = (int)lround(1 + 0.1 * n);

    // |A| statistics with ALGLIB's k = (i==v ? 1 : 2) quirk (opt.cpp:29893-29915); max|b|; start point clipped to the
    // bounds (29979-29998) and sasstartoptimization (27377-27399)
    {
        double s1 = 0.0, s2 = 0.0, mb = 0.0;
        const RowMap<Ex> rm(ex, n);
        for (int base = 0; base < n; base += RowMap<Ex>::ROWS) {
            const int i = base + rm.ioff;
            if (i < n) {
                for (int j = (rm.j0 > i ? rm.j0 : i); j < rm.j1; j++) {
                    const double v = q.S[i * ld + j];
                    const double vv = fabs(v);
                    const double k = ((double)i == v) ? 1.0 : 2.0;
                    s1 += vv * k;
                    s2 += vv * vv * k;
                }
            }
        }
        for (int i = ex.lane(); i < n; i += Ex::NL) {
            mb = fmax(mb, fabs(w.exb[i]));
            double v = w.exxc[i];
            int cs = -1;
            if (i >= NMAIN && v <= 0.0) { v = 0.0; cs = 0; }
            q.xc[i] = v;
            w.cstatus[i] = cs;
        }
        double ss[2] = {s1, s2}, mm[1] = {mb};
        ex.template allred<2, 1>(ss, mm);
        st.absasum = ss[0]; st.absasum2 = ss[1]; st.mb = mm[0];
    }
    int term = 0;
    // NOTE: ALGLIB's single-Cholesky fast path for unconstrained problems (opt.cpp:30033-30073) is gated
    // on akind==0 (CQM storage); DENSE-AUL calls QQP with akind==2 (opt.cpp:41324), so the generic
    // CG + constrained-Newton iteration below runs even when there are no slack variables yet.
    int cgmax = cgminits;
    int outerits = 0;
    double stpbuf[3];
    for (;;) {
        if (maxouterits > 0 && outerits >= maxouterits) { term = 5; break; }
        if (outerits > 0) {
            // epsx stopping test (30137-30149); epsf = 0 so the function test is skipped
            double v = 0.0;
            for (int i = ex.lane(); i < n; i += Ex::NL) { const double t = q.xp[i] - q.xc[i]; v += t * t; }
            v = allsum1(ex, v);
            if (sqrt(v) <= epsx) { term = 2; break; }
        }
        outerits++;
        for (int i = ex.lane(); i < n; i += Ex::NL) { q.xp[i] = q.xc[i]; q.cgp[i] = 0.0; q.dp[i] = 0.0; }
        ex.sync();
        for (int cgcnt = 0; cgcnt <= cgmax - 1; cgcnt++) {
            symv(ex, q.S, ld, n, q.xc, w.exb, q.gc);                       // targetgradient
            flops += 2.0 * n * n;
            // sasreactivateconstraints, box-only (28992-29047); constrained gradient; CG coefficients (30199-30221).
            // (sasconstraineddirection's "everything active" clause, 28952-28959, cannot fire: the 30 main
            //  variables are never bounded)
            double r3[3] = {0.0, 0.0, 0.0};
            for (int i = ex.lane(); i < n; i += Ex::NL) {
                const double xi = q.xc[i], g = q.gc[i];
                const bool atb = (i >= NMAIN && xi == 0.0);
                const bool act = atb && g >= 0.0;
                w.cstatus[i] = act ? 1 : -1;
                const double cg = act ? 0.0 : g;
                q.cgc[i] = cg;
                r3[0] += cg * cg;
                const double pv = q.cgp[i];
                r3[1] += pv * pv;
                if (atb && q.dp[i] != 0.0) r3[2] += 1.0;
            }
            ex.template allred<3, 0>(r3, r3);
            const double v = r3[0], vv = r3[1];
            if (sqrt(v) <= 0.0) { term = 4; break; }                       // epsg = 0
            const bool brst = (r3[2] != 0.0) || (vv == 0.0) || (cgcnt % 50 == 0);
            const double beta = brst ? 0.0 : v / vv;
            for (int i = ex.lane(); i < n; i += Ex::NL) {
                double d = -q.cgc[i] + beta * q.dp[i];
                if (w.cstatus[i] > 0) d = 0.0;
                q.dc[i] = d;
            }
            ex.sync();
            double stpmax, cval; int cidx;
            sas_explore_direction(w.cstatus, q.xc, n, q.dc, stpmax, cidx, cval);
            // qqpsolver_quadraticmodel (30753-30821)
            symv(ex, q.S, ld, n, q.dc, (const double*)nullptr, q.t0);
            double d1, d2; int d1est, d2est;
            {
                double ss[2] = {0.0, 0.0}, mm[2] = {0.0, 0.0};
                for (int i = ex.lane(); i < n; i += Ex::NL) {
                    const double di = q.dc[i];
                    ss[0] += di * q.t0[i];
                    ss[1] += di * q.gc[i];
                    mm[0] = fmax(mm[0], fabs(q.xc[i]));
                    mm[1] = fmax(mm[1], fabs(di));
                }
                ex.template allred<2, 2>(ss, mm);
                d2 = 0.5 * ss[0]; d1 = ss[1];
                estimateparabolicmodel(st.absasum, st.absasum2, mm[0], st.mb, mm[1], d1, d2, d1est, d2est);
            }
            flops += 2.0 * n * n;
            if (d1 == 0.0 && d2 == 0.0) { term = 4; break; }
            if (d1est >= 0) { term = 7; break; }
            if (d2est <= 0 && cidx < 0) { term = -4; break; }
            double stp; bool needact; int stpcnt;
            if (d2est > 0) {
                const double fullstp = -d1 / (2 * d2);
                needact = fullstp >= stpmax;
                if (needact) { stp = stpmax; stpbuf[0] = stpmax * 4; stpbuf[1] = fullstp; stpbuf[2] = fullstp / 4; stpcnt = 3; }
                else { stp = fullstp; stpcnt = 0; }
            } else {
                stp = stpmax; needact = true; stpbuf[0] = 4 * stpmax; stpcnt = 1;
            }
            qqp_find_best_step_and_move(ex, w, q, st, q.dc, stp, needact, cidx, cval, stpbuf, stpcnt, flops);
            for (int i = ex.lane(); i < n; i += Ex::NL) { q.dp[i] = q.dc[i]; q.cgp[i] = q.cgc[i]; }
            ex.sync();
        }
        if (term != 0) break;
        cgmax = cgmaxits;
        // constrained Newton phase (30353-30527)
        int newtcnt = 0;
        for (;;) {
            bool b;
            if (newtcnt == 0) {
                b = qqp_cnewton_build(ex, w, q, st);
                flops += (double)n * n * n / 3.0;
                if (b) cgmax = cgminits;
            } else {
                b = qqp_cnewton_update(ex, w, q, st, cnmaxupdates);
                flops += 3.0 * n * n;
            }
            if (!b) break;
            newtcnt++;
            symv(ex, q.S, ld, n, q.xc, w.exb, q.gc);
            // qqpsolver_cnewtonstep (31474-31536), epsg = 0
            double gg = 0.0;
            for (int i = ex.lane(); i < n; i += Ex::NL) {
                const double g = w.isfree[i] ? q.gc[i] : 0.0;
                gg += g * g;
                q.dc[i] = -g;
            }
            gg = allsum1(ex, gg);
            if (sqrt(gg) <= 0.0) break;
            tri_solve(ex, q.S, ld, n, q.zrinv, q.dc);
            symv(ex, q.S, ld, n, q.dc, (const double*)nullptr, q.t0);
            double d1, d2; int d1est, d2est;
            {
                double ss[2] = {0.0, 0.0}, mm[2] = {0.0, 0.0};
                for (int i = ex.lane(); i < n; i += Ex::NL) {
                    const double di = q.dc[i];
                    ss[0] += di * q.t0[i];
                    ss[1] += di * q.gc[i];
                    mm[0] = fmax(mm[0], fabs(q.xc[i]));
                    mm[1] = fmax(mm[1], fabs(di));
                }
                ex.template allred<2, 2>(ss, mm);
                d2 = 0.5 * ss[0]; d1 = ss[1];
                estimateparabolicmodel(st.absasum, st.absasum2, mm[0], st.mb, mm[1], d1, d2, d1est, d2est);
            }
            flops += 6.0 * n * n;
            if (d1est >= 0) break;
            double stpmax, cval; int cidx;
            sas_explore_direction(w.cstatus, q.xc, n, q.dc, stpmax, cidx, cval);
            if (d2est > 0) {
                const double fullstp = -d1 / (2 * d2);
                const bool needact = fullstp >= stpmax;
                double stp; int stpcnt;
                if (needact) { stp = stpmax; stpbuf[0] = stpmax * 4; stpbuf[1] = fullstp; stpbuf[2] = fullstp / 4; stpcnt = 3; }
                else { stp = fullstp; stpcnt = 0; }
                qqp_find_best_step_and_move(ex, w, q, st, q.dc, stp, needact, cidx, cval, stpbuf, stpcnt, flops);
            } else {
                if (cidx < 0) { term = -4; break; }
                if (stpmax == 0.0) { cgmax = cgmaxits; break; }
                // f(x) vs f(x + stpmax d) (30493-30503): evaluated with the candidate machinery
                double f01[2];
                {
                    double* t[2] = {q.t0, q.t1};
                    for (int i = ex.lane(); i < n; i += Ex::NL) {
                        const double xi = q.xc[i];
                        double v0 = xi, v1 = xi + stpmax * q.dc[i];
                        if (i >= NMAIN && v0 < 0.0) v0 = 0.0;
                        if (i >= NMAIN && v1 < 0.0) v1 = 0.0;
                        t[0][i] = v0; t[1][i] = v1;
                    }
                    ex.sync();
                    double lin[2] = {0, 0}, quad[2] = {0, 0};
                    const RowMap<Ex> rm(ex, n);
                    for (int base = 0; base < n; base += RowMap<Ex>::ROWS) {
                        const int i = base + rm.ioff;
                        if (i < n) {
                            double a0 = 0.0, a1 = 0.0;
                            for (int j = rm.j0; j < rm.j1; j++) {
                                const double e = sym_at(q.S, ld, i, j);
                                a0 += e * t[0][j]; a1 += e * t[1][j];
                            }
                            quad[0] += t[0][i] * a0; quad[1] += t[1][i] * a1;
                            if (rm.s == 0) { lin[0] += w.exb[i] * t[0][i]; lin[1] += w.exb[i] * t[1][i]; }
                        }
                    }
                    double r4[4] = {lin[0], lin[1], quad[0], quad[1]};
                    ex.template allred<4, 0>(r4, r4);
                    f01[0] = r4[0] + 0.5 * r4[2]; f01[1] = r4[1] + 0.5 * r4[3];
                }
                if (f01[1] >= f01[0]) { cgmax = cgmaxits; break; }
                stpbuf[0] = stpmax * 4; stpbuf[1] = 1.00; stpbuf[2] = 0.25;
                qqp_find_best_step_and_move(ex, w, q, st, q.dc, stpmax, true, cidx, cval, stpbuf, 3, flops);
                flops += 4.0 * n * n;
            }
        }
        if (term != 0) break;
    }
    // unpack (30546-30565): unit scale, zero origin; slacks clipped / snapped to the bound
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v = q.xc[i];
        if (i >= NMAIN && (v < 0.0 || v == 0.0)) v = 0.0;
        w.exxc[i] = v;
    }
    ex.sync();
    ncholesky += st.ncholesky;
    return term;
}

// ------------------------------------------------------------------------------------------------
// Multiplier update, literal form (opt.cpp:41803-42032): Householder QR of the stacked system
// [K | r ; lambda*mxdiag*I | 0], K = KKT matrix of the equality-constrained model with the columns of
// exactly-active slacks replaced, then back-substitution.  Only the multiplier part of the solution
// is needed, so back-substitution stops at row ntotal.  Structural zeros of the regulariser block
// are skipped (reflector j only touches K rows j.. and regulariser rows 0..j).  Works out of the
// team's global scratch; it is the fallback of the reduced form below.
WBC_HD int kkt_doubles(int nq) { return 2 * nq * (nq + 1); }

template <class Ex>
WBC_HDNI void householder_qr_solve_tail(const Ex& ex, double* M, int nq, int ld, int ntail, double* v, double* sol, double& flops)
{
    // M: 2nq x (nq+1) row-major (ld = nq+1).  On exit the upper triangle holds R and column nq holds Q'r.
    for (int j = 0; j < nq; j++) {
        // rows involved: K rows j..nq-1 and regulariser rows nq..nq+j  -> contiguous range j..nq+j
        const int r0 = j, r1 = nq + j;   // inclusive
        const int len = r1 - r0 + 1;
        // generatereflection (linalg.cpp:19116-19213) on x = M[r0..r1][j]
        const double alpha = M[r0 * ld + j];
        double mx = 0.0;
        for (int r = r0 + ex.lane(); r <= r1; r += Ex::NL) { const double t = M[r * ld + j]; v[r - r0] = t; mx = fmax(mx, fabs(t)); }
        mx = allmax1(ex, mx);
        double xnorm = 0.0;
        if (mx != 0.0) {
            double s = 0.0;
            for (int r = 1 + ex.lane(); r < len; r += Ex::NL) { const double t = v[r] / mx; s += t * t; }
            s = allsum1(ex, s);
            xnorm = sqrt(s) * mx;
        }
        double tau = 0.0, beta = alpha;
        if (xnorm != 0.0) {
            const double m2 = fmax(fabs(alpha), fabs(xnorm));
            const double a = alpha / m2, b = xnorm / m2;
            beta = -m2 * sqrt(a * a + b * b);
            if (alpha < 0.0) beta = -beta;
            tau = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
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
        s = allsum1(ex, s);
        const double xi = (M[i * ld + nq] - s) / M[i * ld + i];
        if (ex.lane() == 0) sol[i] = xi;
        ex.sync();
    }
}

template <class Ex>
WBC_HDNI void update_lagrange_multipliers_literal(const Ex& ex, const Work& w, int nec, int nic, Stats& st)
{
    const int ntotal = NMAIN + nic, ktotal = nec + nic, nq = ntotal + ktotal, ld = nq + 1;
    double* M = w.kkt;
    // reference point (X0, L0) (41888-41895)
    for (int i = ex.lane(); i < nq; i += Ex::NL) w.sv0[i] = (i < ntotal) ? w.exxc[i] : w.nulcest[i - ntotal];
    for (int i = ex.lane(); i < 2 * nq * ld; i += Ex::NL) M[i] = 0.0;
    ex.sync();
    double mxdiag = 0.0;
    for (int i = 0; i < NMAIN; i++) mxdiag = fmax(mxdiag, fabs(w.SA[i * LDA + i]));
    if (mxdiag == 0.0) mxdiag = 1.0;
    const double lambdareg = 1.0e-8;
    // quadratic term and -b (41919-41927)
    for (int i = 0; i < NMAIN; i++)
        for (int j = ex.lane(); j <= NMAIN; j += Ex::NL)
            M[i * ld + (j < NMAIN ? j : nq)] = (j < NMAIN) ? sym_at(w.SA, LDA, i, j) : -w.b[i];
    // constraints (41933-41946)
    for (int i = 0; i < ktotal; i++) {
        for (int j = ex.lane(); j < NMAIN; j += Ex::NL) {
            const double c = -w.C[i * 31 + j];
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
    householder_qr_solve_tail(ex, M, nq, ld, ktotal, w.qrv, w.sv0, st.flops);
    // sv0 is overwritten in its tail by the solution; nulcest still holds L0
    for (int i = ex.lane(); i < ktotal; i += Ex::NL) w.nulcest[i] = w.nulcest[i] + w.sv0[ntotal + i];
    ex.sync();
}

// ------------------------------------------------------------------------------------------------
// Multiplier update, reduced form.  The stacked system above is the KKT system of the equality-
// constrained model  min 1/2 x'Ax + b'x  s.t.  c_r'x = d_r  for r in ACT = {equalities} U {inequality
// rows whose slack is exactly 0}; rows with a free slack get multiplier 0 (their slack-stationarity
// row reads -nu_r = 0).  With A = U'U the multipliers solve the Schur-complement system
//     (W W') nu_ACT = d_ACT + W t,   row m of W = U^-T c_m,  t = U^-T b,
// which is solved for the correction delta = nu_ACT - nu0_ACT: the literal system is posed in corrections to
// (X0, L0), and when ACT is rank deficient (e.g. the whole friction pyramid of an unloaded foot) its
// Tikhonov term (lambda = 1e-8 max|A_ii|) selects the correction of least norm.  That is reproduced by a
// Cholesky factorisation that skips dependent rows, S = Lt Lt', and  delta = Lt G^-1 G^-1 Lt' rho,
// G = Lt'Lt.  Where the two forms could differ by more than rounding -- a pivot that is neither
// clearly independent nor clearly noise, inconsistent dependent rows, an active set larger than KACAP --
// the routine returns false and the caller runs the literal form.  Workspace: the (idle) QQP array.
template <class Ex>
WBC_HDNI bool update_lagrange_multipliers_reduced(const Ex& ex, const Work& w, int nec, int nic, Stats& st, double pivtol)
{
    const int ktotal = nec + nic;
    int* act = w.cstatus;               // QQP is not running: reuse its integer arrays
    int* dep = w.isfree;
    // ---- active list, in row order (first warp, ballot compaction)
    if (ex.warp() == 0) {
        int ka = 0;
        for (int base = 0; base < ktotal; base += Ex::WL) {
            const int r = base + ex.wlane();
            const bool on = (r < ktotal) && ((r < nec) || (w.exxc[NMAIN + (r - nec)] == 0.0));
            const unsigned m = ex.ballot(on);
            int pos;
#if defined(__CUDA_ARCH__)
            pos = ka + __popc(m & ((1u << ex.wlane()) - 1u));
            if (on && pos < KACAP + 1) act[pos] = r;
            ka += __popc(m);
#else
            pos = ka;
            if (on && pos < KACAP + 1) act[pos] = r;
            ka += (int)m;
#endif
        }
        if (ex.wlane() == 0) w.iscr[0] = ka;
    }
    ex.sync();
    const int ka = w.iscr[0];
    if (ka > KACAP) return false;
    double* Sm = w.Ssh;                          // [KACAP][LDM] transposed-lower Schur complement / factor
    double* W = w.Ssh + KACAP * LDM;             // [KACAP+1][31]: U^-T c_m | d_m ; last row: t
    double* G = W;                               // [KACAP][LDM]   (W is dead by then)
    double* sd = w.vsh;                          // vectors in the (idle) QQP vector block
    double* srinv = w.vsh + NCAP;
    double* pinv = w.vsh + 2 * NCAP;
    double* rho = w.vsh + 3 * NCAP;
    double* dl = w.vsh + 4 * NCAP;
    double* u1 = w.vsh + 5 * NCAP;
    double* gd = w.vsh + 6 * NCAP;
    double* grinv = w.vsh + 7 * NCAP;
    double* nu0 = w.vsh + 8 * NCAP;
    // ---- forward substitutions U' y = c_m, one thread per right-hand side
    for (int m = ex.lane(); m <= ka; m += Ex::NL) {
        double* y = &W[m * 31];
        const double* src = (m < ka) ? &w.C[act[m] * 31] : w.b;
        for (int i = 0; i < NMAIN; i++) {
            double sacc = src[i];
            for (int k = 0; k < i; k++) sacc -= w.SA[i * LDA + k] * y[k];
            y[i] = sacc * w.larinv[i];
        }
        y[NMAIN] = (m < ka) ? src[NMAIN] : 0.0;
        if (m < ka) nu0[m] = w.nulcest[act[m]];
    }
    ex.sync();
    // ---- Schur complement (transposed-lower + diagonal vector)
    for (int e = ex.lane(); e < ka * ka; e += Ex::NL) {
        const int r = e / ka, c = e - r * ka;
        if (r > c) continue;
        double sacc = 0.0;
        for (int k = 0; k < NMAIN; k++) sacc += W[r * 31 + k] * W[c * 31 + k];
        if (r == c) sd[r] = sacc; else Sm[c * LDM + r] = sacc;
    }
    ex.sync();
    // rho = d + W t - S nu0
    for (int m = ex.lane(); m < ka; m += Ex::NL) {
        double sacc = W[m * 31 + NMAIN];
        for (int k = 0; k < NMAIN; k++) sacc += W[m * 31 + k] * W[ka * 31 + k];
        double sn = sd[m] * nu0[m];
        for (int k = 0; k < ka; k++)
            if (k != m) sn += ((k < m) ? Sm[m * LDM + k] : Sm[k * LDM + m]) * nu0[k];
        rho[m] = sacc - sn;
        dl[m] = sacc - sn;
    }
    ex.sync();
    st.flops += (double)(ka + 1) * NMAIN * NMAIN + (double)ka * ka * NMAIN + 2.0 * ka * NMAIN + 2.0 * ka * ka;
    bool ambiguous = false;
    chol_lowerT<true>(ex, Sm, LDM, ka, sd, srinv, pinv, dep, pivtol, &ambiguous);
    if (ambiguous) return false;
    st.flops += (double)ka * ka * ka / 3.0 + 2.0 * ka * ka;
    double ndep = 0.0;
    for (int m = ex.lane(); m < ka; m += Ex::NL) ndep += dep[m];
    ndep = allsum1(ex, ndep);
    if (ndep == 0.0) {
        tri_solve(ex, Sm, LDM, ka, srinv, dl);
    } else {
        // G = Lt'Lt over the kept columns (identity on the skipped ones), u = Lt' rho.
        // Lt[i][a] = U[a][i] = Sm[i*LDM+a] (a < i), Lt[a][a] = sd[a]; skipped columns are exact zeros.
        for (int e = ex.lane(); e < ka * ka; e += Ex::NL) {
            const int a = e / ka, c = e - a * ka;
            if (a > c) continue;
            double sacc;
            if (dep[a] || dep[c]) sacc = (a == c) ? 1.0 : 0.0;
            else {
                sacc = ((a == c) ? sd[c] : Sm[c * LDM + a]) * sd[c];
                for (int i = c + 1; i < ka; i++) sacc += Sm[i * LDM + a] * Sm[i * LDM + c];
            }
            if (a == c) gd[a] = sacc; else G[c * LDM + a] = sacc;
        }
        for (int a = ex.lane(); a < ka; a += Ex::NL) {
            double sacc = sd[a] * rho[a];
            for (int i = a + 1; i < ka; i++) sacc += Sm[i * LDM + a] * rho[i];
            u1[a] = dep[a] ? 0.0 : sacc;
        }
        ex.sync();
        if (!chol_lowerT<false>(ex, G, LDM, ka, gd, grinv, pinv, (int*)nullptr, 0.0, (bool*)nullptr)) return false;
        tri_solve(ex, G, LDM, ka, grinv, u1);
        // consistency: Lt u1 is the projection of rho on range(S); it must reproduce rho
        double worst = 0.0, scale = 0.0;
        for (int i = ex.lane(); i < ka; i += Ex::NL) {
            double sacc = sd[i] * u1[i];
            for (int a = 0; a < i; a++) sacc += Sm[i * LDM + a] * u1[a];
            worst = fmax(worst, fabs(sacc - rho[i]));
            scale = fmax(scale, fabs(w.C[act[i] * 31 + NMAIN]));
        }
        {
            double mm[2] = {worst, scale};
            ex.template allred<0, 2>(mm, mm);
            worst = mm[0]; scale = mm[1];
        }
        if (worst > 1.0e-9 * (scale + 1.0)) return false;
        tri_solve(ex, G, LDM, ka, grinv, u1);
        for (int i = ex.lane(); i < ka; i += Ex::NL) {
            double sacc = sd[i] * u1[i];
            for (int a = 0; a < i; a++) sacc += Sm[i * LDM + a] * u1[a];
            dl[i] = sacc;
        }
        ex.sync();
        st.flops += 4.0 * ka * ka * ka / 3.0;
        st.flags |= 16;
    }
    for (int i = ex.lane(); i < ktotal; i += Ex::NL) w.nulcest[i] = 0.0;
    ex.sync();
    for (int m = ex.lane(); m < ka; m += Ex::NL) w.nulcest[act[m]] = nu0[m] + dl[m];
    ex.sync();
    return true;
}

// ------------------------------------------------------------------------------------------------
// generateexmodel (opt.cpp:41594-41740): extended box-QP in [x; slacks].  Upper triangle of q.S.
WBC_HD void tri_index30(int e, int& i, int& j)
{
    // e in [0, 465) -> (i <= j) of the 30 x 30 upper triangle, row-major
    int r = (int)((61.0f - sqrtf(3721.0f - 8.0f * (float)e)) * 0.5f);
    if (r < 0) r = 0;
    if (r > 29) r = 29;
    while (r < 29 && ((r + 1) * 30 - ((r + 1) * r) / 2) <= e) r++;
    while (r > 0 && (r * 30 - (r * (r - 1)) / 2) > e) r--;
    i = r;
    j = r + (e - (r * 30 - (r * (r - 1)) / 2));
}

template <class Ex>
WBC_HDNI void generate_ex_model(const Ex& ex, const Work& w, const QV& q, int nec, int nic, double rho, double& flops)
{
    const int n = NMAIN + nic, ld = q.ld, kw = nec + nic;
    // quadratic term, columns < NMAIN: A + rho * C'C
    for (int e = ex.lane(); e < 465; e += Ex::NL) {
        int i, j;
        tri_index30(e, i, j);
        double s0 = 0.0, s1 = 0.0;
        int r = 0;
        for (; r + 1 < kw; r += 2) {
            s0 += w.C[r * 31 + i] * w.C[r * 31 + j];
            s1 += w.C[(r + 1) * 31 + i] * w.C[(r + 1) * 31 + j];
        }
        if (r < kw) s0 += w.C[r * 31 + i] * w.C[r * 31 + j];
        q.S[i * ld + j] = w.SA[i * LDA + j] + rho * (s0 + s1);
    }
    // slack columns and the slack block
    for (int e = ex.lane(); e < NMAIN * nic; e += Ex::NL) {
        const int k = e / NMAIN, i = e - k * NMAIN;
        q.S[i * ld + NMAIN + k] = 0.0 + rho * w.C[(nec + k) * 31 + i];
    }
    for (int i = NMAIN + ex.warp(); i < n; i += Ex::NW)
        for (int j = i + ex.wlane(); j < n; j += Ex::WL) q.S[i * ld + j] = (i == j) ? 0.0 + rho * 1.0 : 0.0;
    // linear term (41650-41657, 41734-41737): per element, rows in order, two updates per row
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v;
        if (i < NMAIN) {
            v = w.b[i];
            for (int r = 0; r < kw; r++) {
                const double c = w.C[r * 31 + i];
                v += c * (-rho * w.C[r * 31 + NMAIN]);
                v += c * (-w.nulc[r]);
            }
        } else {
            const int r = nec + (i - NMAIN);
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
// Working-set expansion (opt.cpp:41350-41390) and eviction (41400-41418): literal sequential selection on the
// team's first warp (arg-max by butterfly, ties to the lowest index like the reference's strict '>' scan).
// Results: w.iscr[1] = new nicwork, w.iscr[2] = extended flag.
template <class Ex>
WBC_HDNI void update_working_set(const Ex& ex, const Work& w, int nec, int nictotal, int nicwork, bool allowevict)
{
    if (ex.warp() == 0) {
        const int l = ex.wlane();
        int extended = 0, added = 0;
        while ((double)added < 1 + 0.20 * NMAIN && nicwork < nictotal) {
            // k = argmax_{j >= nicwork} nicerr[j], first maximum
            double bv = -1.7976931348623157e308;
            int bk = 0x7fffffff;
            for (int j = nicwork + l; j < nictotal; j += Ex::WL) {
                const double v = w.nicerr[j];
                if (v > bv) { bv = v; bk = j; }
            }
            for (int o = Ex::WL / 2; o > 0; o >>= 1) {
                const double ov = ex.shfl_xor(bv, o);
                const int ok = ex.shfl_xori(bk, o);
                if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
            }
            const int k = bk;
            if (!(bv > 0.0)) break;
            // swap rows nec+nicwork <-> nec+k of C, and the per-constraint bookkeeping
            if (k != nicwork) {
                for (int j = l; j < 31; j += Ex::WL) {
                    const double t = w.C[(nec + nicwork) * 31 + j];
                    w.C[(nec + nicwork) * 31 + j] = w.C[(nec + k) * 31 + j];
                    w.C[(nec + k) * 31 + j] = t;
                }
            }
            if (l == 0) {
                const double t = w.nicerr[nicwork]; w.nicerr[nicwork] = w.nicerr[k]; w.nicerr[k] = t;
                const int ti = w.nicnact[nicwork]; w.nicnact[nicwork] = w.nicnact[k]; w.nicnact[k] = ti;
                w.exxc[NMAIN + nicwork] = 0.0;
                w.nulc[nec + nicwork] = 0.0;
                w.nicnact[nicwork] = w.nicnact[nicwork] + 1;
            }
            ex.wsync();
            nicwork++; added++;
            extended = 1;
        }
        if (allowevict) {
            for (int k = nicwork - 1; k >= 0; k--) {
                if (w.nicerr[k] < -0.01 && w.nicnact[k] <= 1) {
                    const int last = nicwork - 1;
                    ex.wsync();
                    if (k != last) {
                        for (int j = l; j < 31; j += Ex::WL) {
                            const double t = w.C[(nec + last) * 31 + j];
                            w.C[(nec + last) * 31 + j] = w.C[(nec + k) * 31 + j];
                            w.C[(nec + k) * 31 + j] = t;
                        }
                    }
                    if (l == 0) {
                        double t = w.nicerr[last]; w.nicerr[last] = w.nicerr[k]; w.nicerr[k] = t;
                        const int ti = w.nicnact[last]; w.nicnact[last] = w.nicnact[k]; w.nicnact[k] = ti;
                        t = w.exxc[NMAIN + last]; w.exxc[NMAIN + last] = w.exxc[NMAIN + k]; w.exxc[NMAIN + k] = t;
                        t = w.nulc[nec + last]; w.nulc[nec + last] = w.nulc[nec + k]; w.nulc[nec + k] = t;
                    }
                    ex.wsync();
                    nicwork--;
                }
            }
        }
        if (l == 0) { w.iscr[1] = nicwork; w.iscr[2] = extended; }
    }
    ex.sync();
}

// ------------------------------------------------------------------------------------------------
// The solver.  On entry the team has staged the problem:
//   Q (30x30 row-major, lower triangle used like minqpsetquadraticterm's default, opt.cpp:4962/18959) in w.Ssh[0..900),
//   c in w.exb[0..30), L (nrows x 31, first neq rows equalities, rest "<=") in w.C.
// Result: w.xs[0..30).
template <class Ex>
WBC_HDN void solve_denseaul(const Ex& ex, const Work& w, const Settings& cfg, int nrows, int neq, Stats& st)
{
    const int nec = neq, nictotal = nrows - neq;
    st.termination = 0; st.ncholesky = 0; st.outer_its = 0; st.qqp_calls = 0; st.nicwork = 0;
    st.kkt_dim_max = 0; st.flags = 0; st.flops = 0.0;
    const double* Q = w.Ssh;
    double* As = w.Ssh + 912;                   // staging: scaled A, full symmetric 30 x 30 with ld LDA

    // ---- minqpoptimize: autodiag scale (opt.cpp:48146-48185)
    double bad = 0.0;
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) {
        const double d = Q[i * NMAIN + i];
        if (d <= 0.0) bad = 1.0;
        w.s[i] = 1.0 / sqrt(d);
    }
    bad = allsum1(ex, bad);
    if (bad != 0.0) { st.termination = -9; return; }

    // ---- scaleshiftoriginalproblem (opt.cpp:42088-42339)
    double an = 0.0;
    for (int e = ex.lane(); e < NMAIN * NMAIN; e += Ex::NL) {
        const int i = e / NMAIN, j = e - i * NMAIN;
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        const double v = Q[hi * NMAIN + lo] * w.s[lo] * w.s[hi];
        As[i * LDA + j] = v;
        an += v * v;
    }
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) w.b[i] = w.exb[i] * w.s[i];
    // constraint rows: thread-per-row scaling + normalisation (42219-42314)
    for (int r = ex.lane(); r < nrows; r += Ex::NL) {
        double vv = 0.0;
        for (int j = 0; j < NMAIN; j++) {
            const double v = w.C[r * 31 + j] * w.s[j];
            w.C[r * 31 + j] = v;
            vv += v * v;
        }
        double rhs = w.C[r * 31 + NMAIN];
        vv = sqrt(vv);
        if (vv > 0.0) {
            vv = 1.0 / vv;
            for (int j = 0; j < NMAIN; j++) w.C[r * 31 + j] *= vv;
            rhs *= vv;
        }
        w.C[r * 31 + NMAIN] = rhs;
    }
    // ---- normalizequadraticterm (opt.cpp:42374-42446); exact zeros of a row are skipped (adds nothing)
    {
        double mm[1] = {0.0}, ss[1] = {an};
        ex.template allred<1, 0>(ss, ss);        // barrier: As and C complete
        an = sqrt(ss[0]);
        double maxcac = 0.0;
        for (int r = ex.lane(); r < nrows; r += Ex::NL) {
            double v = 0.0;
            for (int j = 0; j < NMAIN; j++) {
                const double cj = w.C[r * 31 + j];
                if (cj == 0.0) continue;
                double t = 0.0;
                for (int k = 0; k < NMAIN; k++) t += w.C[r * 31 + k] * As[k * LDA + j];
                v += t * cj;
            }
            maxcac = fmax(maxcac, fabs(v));
        }
        mm[0] = maxcac;
        ex.template allred<0, 1>(mm, mm);
        double targetscale = fmax(mm[0], an / NMAIN);
        if (targetscale == 0.0) targetscale = 1.0;
        const double v = 1.0 / targetscale;
        // scaled A: upper triangle + diagonal to SA (global), transposed-lower staging for the factorisation
        for (int e = ex.lane(); e < NMAIN * NMAIN; e += Ex::NL) {
            const int i = e / NMAIN, j = e - i * NMAIN;
            const double a = As[i * LDA + j] * v;
            if (i <= j) w.SA[i * LDA + j] = a;
            if (i == j) w.ladiag[i] = a;
            As[i * LDA + j] = a;
        }
        for (int i = ex.lane(); i < NMAIN; i += Ex::NL) w.b[i] *= v;
        ex.sync();
        st.flops += 2.0 * nrows * NMAIN * NMAIN;
    }

    // ---- selectinitialworkingset (opt.cpp:42474-42523).  The factor of A is kept for the multiplier updates.
    int nicwork = 0;
    bool allowevict = true;
    {
        const bool pd = chol_lowerT<false>(ex, As, LDA, NMAIN, w.ladiag, w.larinv, w.nicerr /*scratch*/, (int*)nullptr, 0.0, (bool*)nullptr);
        if (!pd) { nicwork = nictotal; allowevict = false; st.flags |= 1; }
        for (int c = 1 + ex.warp(); c < NMAIN; c += Ex::NW)
            for (int k = ex.wlane(); k < c; k += Ex::WL) w.SA[c * LDA + k] = As[c * LDA + k];
        st.flops += 9000.0;
    }
    const bool have_factor = allowevict;
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
            const bool spill = nwork > NCAP;
            if (spill) st.flags |= 32;
            const QV q = make_qv(w, spill);
            generate_ex_model(ex, w, q, nec, nicwork, rho, st.flops);
            const int term = qqp_optimize(ex, w, q, nwork, 0.01 * epsx, 50, st.ncholesky, st.flops);
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
            update_working_set(ex, w, nec, nictotal, nicwork, allowevict);
            nicwork = w.iscr[1];
            extended = w.iscr[2] != 0;
            ex.sync();
        } while (extended);

        const int kwork = nec + nicwork;
        // multiplier estimate (41438-41439)
        for (int i = ex.lane(); i < kwork; i += Ex::NL) w.nulcest[i] = w.nulc[i];
        ex.sync();
        {
            const int nq = NMAIN + nicwork + kwork;
            if (nq > st.kkt_dim_max) st.kkt_dim_max = nq;
            bool done = false;
            if (cfg.kkt_mode == 1 && have_factor) done = update_lagrange_multipliers_reduced(ex, w, nec, nicwork, st, cfg.kkt_pivtol);
            if (!done) {
                st.flags |= 8;
                for (int i = ex.lane(); i < kwork; i += Ex::NL) w.nulcest[i] = w.nulc[i];
                ex.sync();
                update_lagrange_multipliers_literal(ex, w, nec, nicwork, st);
            }
        }
        // feasibility error and multiplier update (41444-41476): thread-per-row, summed by the team
        const double feaserrprev = feaserr;
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
        feaserr = sqrt(allsum1(ex, fe));
        st.flops += 4.0 * kwork * NMAIN;
        if (feaserr < epsx) goodcounter++; else goodcounter = 0;
        if (feaserr > feaserrprev * requestedfeasdecrease) stagnationcounter++; else stagnationcounter = 0;
        if (goodcounter >= 2) break;
        if (stagnationcounter >= 2) rho = fmin(rho * 10.0, maxrho);
        else rho = fmin(rho * 1.41, maxrho);
    }
    st.nicwork = nicwork;
    // unscale (41548-41583): x = s * xc  (+ origin 0); no box constraints on x
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) w.xs[i] = w.s[i] * w.exxc[i] + 0.0;
    ex.sync();
    st.termination = 2;
}

}  // namespace wbcqp
