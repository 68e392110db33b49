// Device-only fast path of the QQP inner solver (opt.cpp:29675-30566) for the shared-memory case
// (nic <= NICCAP): same decisions and arithmetic as qqp_optimize<false> in qp_warp.cuh, re-expressed for
// a single warp so that it issues few instructions and few branches (the warp runs almost alone on its
// scheduler, so every taken branch and every dependent instruction is exposed latency):
//   * every QQP vector lives in registers in a two-slot form: slot A = main variable `lane` (lanes 0..29),
//     slot B = slack variable 30 + `lane` (lanes 0..nic-1); only vectors that are multiplied by E are
//     mirrored to shared memory (for the broadcast reads of the products);
//   * the products with E = [H, CI'; CI, rho I] are straight-line code over the 30 main columns, with
//     128-bit loads of the vector, four independent accumulation chains, and no predicates on loads (out-of-
//     range lanes read finite in-bounds values and drop the result);
//   * the constrained-Newton factorisation handles two columns per step on a transposed, pair-interleaved factor
//     (conflict-free 128-bit loads); triangular solves keep the running right-hand side in registers.
// Included by qp_warp.cuh; uses its layout constants.
#pragma once
#if defined(__CUDACC__)

namespace wbcqp {
namespace fast {

// Code size is a first-order cost here: ncu shows the SM instruction cache hit rate at 60 % and the GPC-level
// instruction cache at half of its peak request rate when every warp of an SM runs a different phase of a
// 200 KB kernel.  So loops are not unrolled unless they are the product itself, and maxima / arg-minima use the
// warp REDUX unit on the bit patterns (exact for non-negative doubles).
// Which routines are separate functions.  Round 1 made every routine of the solver a non-inlined function to keep the kernel small (the
// one-warp-per-solve kernel was instruction-fetch bound with everything inlined, and its hot code was twice today's).  With SM roles,
// express lanes and the later diets the balance moved, and a call is expensive in a 168-register kernel (live values handed over or
// spilled at every call, no scheduling across it).  Measured in three rounds of interleaved A/B builds, all bit-identical
// (profiles/r02_aw_inlining_ab.txt): everything on the QQP path inlined except the product `symv` (five call sites; inlined it is
// +0.3 % / -1 %), plus the small helpers and the model / working-set / multiplier-update / set-up routines of qp_warp.cuh
// (WBC_SMALL_NI, WBC_HDNI_G, WBC_HDNI_M):  4 096 instances 3.02 -> 2.44 ms, 65 536 instances 37.7 -> 32.3 ms.
// -DWBC_OUTLINE_<NAME> restores a call, -DWBC_INLINE_SYMV inlines the product (A/B experiments).
#ifdef WBC_OUTLINE_CHOL_BUILD30
#define WBC_NI_CHOL_BUILD30 __noinline__
#else
#define WBC_NI_CHOL_BUILD30 __forceinline__
#endif
#ifdef WBC_OUTLINE_TRI_SOLVE30
#define WBC_NI_TRI_SOLVE30 __noinline__
#else
#define WBC_NI_TRI_SOLVE30 __forceinline__
#endif
#ifdef WBC_OUTLINE_RANK1_FIX30
#define WBC_NI_RANK1_FIX30 __noinline__
#else
#define WBC_NI_RANK1_FIX30 __forceinline__
#endif
#ifdef WBC_OUTLINE_NEWTON_DIRECTION
#define WBC_NI_NEWTON_DIRECTION __noinline__
#else
#define WBC_NI_NEWTON_DIRECTION __forceinline__
#endif
#ifdef WBC_OUTLINE_NEWTON_DIAG
#define WBC_NI_NEWTON_DIAG __noinline__
#else
#define WBC_NI_NEWTON_DIAG __forceinline__
#endif
#ifdef WBC_OUTLINE_EVAL4
#define WBC_NI_EVAL4 __noinline__
#else
#define WBC_NI_EVAL4 __forceinline__
#endif
#ifdef WBC_OUTLINE_QUADRATIC_MODEL
#define WBC_NI_QUADRATIC_MODEL __noinline__
#else
#define WBC_NI_QUADRATIC_MODEL __forceinline__
#endif
#ifdef WBC_OUTLINE_EXPLORE
#define WBC_NI_EXPLORE __noinline__
#else
#define WBC_NI_EXPLORE __forceinline__
#endif
#ifdef WBC_OUTLINE_STEP_AND_MOVE
#define WBC_NI_STEP_AND_MOVE __noinline__
#else
#define WBC_NI_STEP_AND_MOVE __forceinline__
#endif
#ifdef WBC_OUTLINE_QQP_OPTIMIZE_FAST
#define WBC_NI_QQP_OPTIMIZE_FAST __noinline__
#else
#define WBC_NI_QQP_OPTIMIZE_FAST __forceinline__
#endif
#ifdef WBC_INLINE_SYMV
#define WBC_NI_SYMV __forceinline__
#else
#define WBC_NI_SYMV __noinline__
#endif
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double bshfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ WBC_SMALL_NI double ddiv(double a, double b) { return a / b; }
__device__ __forceinline__ double dsqrt(double a) { return dsqrt_ni(a); }
// warp sum, every lane gets the same bits
__device__ __forceinline__ double wsum(double v) { return warp_sum_ni(v); }      // the butterfly (unrolled inside)
// warp maximum of non-negative doubles (bit patterns order like the values)
__device__ __forceinline__ double wmax_nn(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(FULL, hi);
    const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

// y = E x for the vector mirrored at `x` (shared, 16-byte aligned, entries >= n finite).  Returns slot A / slot B parts.
// nic2 = nic rounded up to even; rows [nic, nic2) of CI are zero.
__device__ WBC_NI_SYMV double2 symv(const double* __restrict__ x, int nic2, double rho)
{
    const int l = threadIdx.x & 31;
    const double* H = wbc_smem + sl::OFF_H + l;             // column walk of row l (H symmetric)
    const double* Crow = wbc_smem + sl::OFF_CI + (l < NICCAP ? l : NICCAP - 1) * LDH;   // slack row l (lanes past the array: any row, result dropped)
    const double* Ccol = wbc_smem + sl::OFF_CI + l;         // column l of the slack rows
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll 5
    for (int j = 0; j < NMAIN; j += 2) {
        const double2 xv = ld2(x + j);
        a0 += H[j * LDH] * xv.x;
        a1 += H[(j + 1) * LDH] * xv.y;
        b0 += Crow[j] * xv.x;
        b1 += Crow[j + 1] * xv.y;
    }
#pragma unroll 1
    for (int k = 0; k < nic2; k += 2) {
        const double2 xv = ld2(x + NMAIN + k);
        a0 += Ccol[k * LDH] * xv.x;
        a1 += Ccol[(k + 1) * LDH] * xv.y;
    }
    double2 r;
    r.x = a0 + a1;
    r.y = (b0 + b1) + rho * x[NMAIN + l];
    return r;
}

// f_k = exb . t_k + 0.5 t_k . (E t_k) for the four projected points t_k = P(xc + s_k d) (opt.cpp:30582-30652).
// xc, d, exb in two-slot registers.  Results to SP[0..3] (shared).
// The candidates are broadcast from shared memory, interleaved four to an entry: the 30 main entries overlay the mirrors
// of x and d (and the head of EXB), the slack entries overlay EXXC -- all four are dead here (x, exb and the point of the
// extended model live in registers during a QQP call, d is rewritten before its next use); the CALLER restores the mirror of x.
__device__ WBC_NI_EVAL4 void eval4(double xcA, double xcB, double dA, double dB, double exbA, double exbB, int nic, double rho, double s0,
                                  double s1, double s2, double s3)
{
    const int l = threadIdx.x & 31;
    const int nic2 = (nic + 1) & ~1;
    double* t4m = wbc_smem + sl::OFF_XC;                     // [30][4]
    double* t4s = wbc_smem + sl::OFF_EXXC;                   // [NICCAP][4]
    static_assert(NMAIN * 4 <= 3 * 48 && NICCAP * 4 <= 104, "trial points fit the idle arrays");
    const bool vA = l < NMAIN, vB = l < nic;
    double tA[4], tB[4];
    const double s[4] = {s0, s1, s2, s3};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        tA[k] = (s[k] != 0.0) ? xcA + s[k] * dA : xcA;
        double v = (s[k] != 0.0) ? xcB + s[k] * dB : xcB;
        if (v < 0.0) v = 0.0;
        tB[k] = vB ? v : 0.0;
        if (!vA) tA[k] = 0.0;
    }
    __syncwarp();                                            // every lane is done reading the arrays the candidates overlay
    if (vA) {
        *reinterpret_cast<double2*>(t4m + l * 4) = make_double2(tA[0], tA[1]);
        *reinterpret_cast<double2*>(t4m + l * 4 + 2) = make_double2(tA[2], tA[3]);
    }
    if (l < nic2) {      // includes the zero pad entry when nic is odd
        *reinterpret_cast<double2*>(t4s + l * 4) = make_double2(tB[0], tB[1]);
        *reinterpret_cast<double2*>(t4s + l * 4 + 2) = make_double2(tB[2], tB[3]);
    }
    __syncwarp();
    const double* H = wbc_smem + sl::OFF_H + l;
    const double* Crow = wbc_smem + sl::OFF_CI + (l < NICCAP ? l : NICCAP - 1) * LDH;
    const double* Ccol = wbc_smem + sl::OFF_CI + l;
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int j = 0; j < NMAIN; j++) {
        const double2 t01 = ld2(t4m + j * 4), t23 = ld2(t4m + j * 4 + 2);
        const double h = H[j * LDH], c = Crow[j];
        a[0] += h * t01.x; a[1] += h * t01.y; a[2] += h * t23.x; a[3] += h * t23.y;
        b[0] += c * t01.x; b[1] += c * t01.y; b[2] += c * t23.x; b[3] += c * t23.y;
    }
#pragma unroll 1
    for (int k = 0; k < nic2; k++) {
        const double2 t01 = ld2(t4s + k * 4), t23 = ld2(t4s + k * 4 + 2);
        const double c = Ccol[k * LDH];
        a[0] += c * t01.x; a[1] += c * t01.y; a[2] += c * t23.x; a[3] += c * t23.y;
    }
    double r8[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double eb = b[k] + rho * tB[k];
        r8[k] = exbA * tA[k] + exbB * tB[k];                 // tA / tB are zero on invalid lanes
        r8[4 + k] = tA[k] * a[k] + tB[k] * eb;
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 8; k++) r8[k] += __shfl_xor_sync(FULL, r8[k], o);
    }
    __syncwarp();
    if (l == 0) {
        double* f = wbc_smem + sl::OFF_SP;
        f[0] = r8[0] + 0.5 * r8[4]; f[1] = r8[1] + 0.5 * r8[5]; f[2] = r8[2] + 0.5 * r8[6]; f[3] = r8[3] + 0.5 * r8[7];
    }
    __syncwarp();
}

// ---- Constrained-Newton linear algebra, slack block first ------------------------------------------------
// The model matrix of the free variables is  E_FF = [Hreg, B'; B, D]  with B = the rows rho c_k of the free slack
// variables and D = diag(d_k) (d_k = rho + regulariser).  ALGLIB factors it with the slack variables last
// (opt.cpp:31058-31201); the same matrix is factored here with the slack block FIRST: its Cholesky factor is the
// diagonal sqrt(D), the coupling block is T = D^-1/2 B, and only the 30 x 30 Schur complement
//     M = Hreg - T'T = U'U
// is factored densely.  Same matrix, same solution of E_FF d = -g up to rounding; a third of the flops of the
// (30 + nic)^3/3 factorisation, every row owned by exactly one lane, and no dependence on nic in the code shape.
//   * T is not stored: its entries T_ik = CI_ik / sqrt(d_i) are formed where they are used from CI and the vector RS of
//     1/sqrt(d_i) (shared memory is what limits the resident warps, and T would be a fifth of it);
//   * fixing slack k after the build (qqpsolver_cnewtonupdate, opt.cpp:31314-31426) turns row/column k of E_FF into
//     the identity, i.e. M <- M + T_k T_k': a rank-one update of U;
//   * the Newton direction:  dx = M^-1 (-g_x + B' D^-1 g_s),  ds = -D^-1 (g_s + B dx).
//
// Storage of the factor Z = U' (lower triangular, row c = lane c's row): TRANSPOSED AND PAIR-INTERLEAVED.  Pair-row p
// holds, for every c >= 2p, the two entries (Z[c][2p], Z[c][2p+1]) as one 16-byte word at
//     zt_base(p) + 2 (c - 2p),      zt_base(p) = 62 p - 2 p^2      (30 - 2p words per pair-row, 480 doubles in all).
// For a fixed column pair the 30 lanes read consecutive words (conflict-free 128-bit loads: the factorisation's dot
// products, the forward sweep); the words of rows 2p and 2p+1 sit at the head of pair-row p, where the slots on and above
// the diagonal carry the reciprocal diagonal:  head = [1/Z[2p][2p], -, Z[2p+1][2p], 1/Z[2p+1][2p+1]].
// The diagonal itself (used by the rank-one update only) is the vector ZD.
__device__ __forceinline__ int zt_base(int p) { return (62 - 2 * p) * p; }
static_assert(sl::Z_DOUBLES >= 480, "pair-interleaved 30 x 30 factor");

// rsB: 1/sqrt(d_k) of slack lane k (0 when the variable is not free); diagA: regularised diagonal of main variable `lane`.
// Returns false on a non-positive pivot.
__device__ WBC_NI_CHOL_BUILD30 bool chol_build30(double diagA, double rsB, unsigned fmask)
{
    const int l = threadIdx.x & 31;
    double* Z = wbc_smem + sl::OFF_Z;
    double* zd = wbc_smem + sl::OFF_ZD;
    double* rs = wbc_smem + sl::OFF_RS;
    const double* H = wbc_smem + sl::OFF_H + l;             // column l (H symmetric); lanes 30, 31 read finite in-bounds values
    const double* CIb = wbc_smem + sl::OFF_CI;
    const double* CIc = CIb + l;
    if (l < NICCAP) rs[l] = rsB;
    __syncwarp();
    const int c = l;
    const int cr = l < NMAIN ? l : NMAIN - 1;              // lanes 30, 31 shadow row 29 (they never store)
#pragma unroll 1
    for (int k = 0; k < NMAIN; k += 2) {
        // dot products of row c with rows k and k + 1 over the finished columns, one pair-row per step
        double p0 = 0.0, q0 = 0.0, p1 = 0.0, q1 = 0.0;
        double* zq = Z;                                      // zq + 2 r = the word of row r in pair-row q:  Z + zt_base(q) - 4 q
#pragma unroll 1
        for (int q = 0; 2 * q < k; q++) {
            const double2 z0 = ld2(zq + 2 * cr), zk = ld2(zq + 2 * k), zk1 = ld2(zq + 2 * k + 2);
            p0 += z0.x * zk.x; q0 += z0.y * zk.y;
            p1 += z0.x * zk1.x; q1 += z0.y * zk1.y;
            zq += 56 - 4 * q;                                // zt_base(q + 1) - 4 (q + 1) - (zt_base(q) - 4 q)
        }
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 1
        for (unsigned m = fmask; m; m &= m - 1u) {
            const int i = __ffs((int)m) - 1;
            const double r = rs[i];
            const double tc = CIc[i * LDH] * r;              // T_ic, T_ik, T_i,k+1
            const double tk0 = CIb[i * LDH + k] * r, tk1 = CIb[i * LDH + k + 1] * r;
            s0 += tc * tk0; s1 += tc * tk1;
        }
        double v0 = ((c == k) ? diagA : H[k * LDH]) - s0 - (p0 + q0);
        double v1 = ((c == k + 1) ? diagA : H[(k + 1) * LDH]) - s1 - (p1 + q1);
        // the three values the two pivots depend on are broadcast at once and every lane finishes both pivots itself
        // (same operations as lane k + 1 performs on its own entries): one shuffle latency on the chain instead of three
        const double piv0 = bshfl(v0, k);
        const double v0n = bshfl(v0, k + 1), v1n = bshfl(v1, k + 1);
        if (!(piv0 > 0.0)) return false;
        const double ri0 = rsqrt(piv0);
        const double z0k = v0 * ri0;
        const double zk1k = v0n * ri0;
        v1 -= z0k * zk1k;
        const double piv1 = v1n - zk1k * zk1k;
        if (!(piv1 > 0.0)) return false;
        const double ri1 = rsqrt(piv1);
        // lane c writes its word of pair-row k/2: rows k and k + 1 put the reciprocal pivots on their diagonal slots
        if (c >= k && c < NMAIN) {
            double2 wv;
            wv.x = (c == k) ? ri0 : z0k;
            wv.y = (c == k) ? 0.0 : ((c == k + 1) ? ri1 : v1 * ri1);
            *reinterpret_cast<double2*>(zq + 2 * c) = wv;   // zq has arrived at pair-row k/2
        }
        if (l == 0) *reinterpret_cast<double2*>(zd + k) = make_double2(piv0 * ri0, piv1 * ri1);
        __syncwarp();
    }
    return true;
}

// Solve U'U x = rhs (30 x 30 factor); rhs and result in shared memory at x[0..30).
// Two columns per step: both pivot components are broadcast at once and the second one is finished redundantly by every
// lane (y_{k+1} = (x_{k+1} - U_{k,k+1} y_k) / U_{k+1,k+1}), so the dependent chain is one shuffle per two columns instead
// of one per column.  The operations on every component are the ones of the one-column sweep, in the same order.
__device__ WBC_NI_TRI_SOLVE30 void tri_solve30(double* x)
{
    const int l = threadIdx.x & 31;
    const double* Z = wbc_smem + sl::OFF_Z;
    const int ql = (l < NMAIN ? l : NMAIN - 1) >> 1, lo = l & 1;
    const double zri = Z[zt_base(ql) + 3 * lo];             // 1 / diagonal of this lane's row
    double x0 = x[l];                                       // lanes 30, 31 carry finite values that are never broadcast
    {
        const double* hd = Z;                               // head of pair-row k/2
#pragma unroll 1
        for (int k = 0; k < NMAIN; k += 2) {                // forward: U' y = rhs, column oriented
            const double2 h0 = ld2(hd), h1 = ld2(hd + 2);   // (1/U_kk, -), (U_{k,k+1}, 1/U_{k+1,k+1})
            const double2 zz = ld2(hd + 2 * (l - k));       // this lane's entries of columns k, k + 1 (lanes < k: finite, unused)
            const double yk = bshfl(x0, k) * h0.x;
            const double yk1 = (bshfl(x0, k + 1) - h1.x * yk) * h1.y;
            if (l > k) x0 -= zz.x * yk;
            if (l > k + 1) x0 -= zz.y * yk1;
            hd += 60 - 2 * k;                               // zt_base(p + 1) - zt_base(p) = 60 - 4 p
        }
    }
    x0 *= zri;
    {
        // backward: U x = y.  Column k of U is row k of Z: lane l needs Z[k][l] = word of row k in pair-row l/2, entry l & 1
        const double* own = Z + zt_base(ql) - 4 * ql + lo;  // own[2 r] = Z[r][l]
        const double* hd = Z + zt_base(NMAIN / 2 - 1);
#pragma unroll 1
        for (int k = NMAIN - 1; k > 0; k -= 2) {
            const double2 h0 = ld2(hd), h1 = ld2(hd + 2);   // (1/U_{k-1,k-1}, -), (U_{k-1,k}, 1/U_kk)
            const double zk = own[2 * k], zk1 = own[2 * k - 2];
            const double xk = bshfl(x0, k) * h1.y;
            const double xk1 = (bshfl(x0, k - 1) - h1.x * xk) * h0.x;
            if (l < k) x0 -= zk * xk;
            if (l < k - 1) x0 -= zk1 * xk1;
            hd -= 66 - 2 * k;                               // zt_base(p) - zt_base(p - 1) = 64 - 4 p,  p = (k - 1) / 2
        }
    }
    x0 *= zri;
    __syncwarp();
    if (l < NMAIN) x[l] = x0;
    __syncwarp();
}

// U'U <- U'U + T_k T_k'  (slack variable k leaves the free set)
__device__ WBC_NI_RANK1_FIX30 void rank1_fix30(int k)
{
    const int l = threadIdx.x & 31;
    double* Z = wbc_smem + sl::OFF_Z;
    double* zd = wbc_smem + sl::OFF_ZD;
    double x = (l < NMAIN) ? wbc_smem[sl::OFF_CI + k * LDH + l] * wbc_smem[sl::OFF_RS + k] : 0.0;      // T_k
    const unsigned nz = __ballot_sync(FULL, x != 0.0);
    if (nz == 0u) return;
    double* own = Z + 2 * (l < NMAIN ? l : NMAIN - 1);      // own[zt_base(q) - 4 q + (j & 1)] = Z[l][j],  q = j / 2
#pragma unroll 1
    for (int j = __ffs((int)nz) - 1; j < NMAIN; j++) {
        const double xj = bshfl(x, j);
        if (xj == 0.0) continue;
        const int q = j >> 1;
        const int e = (58 - 2 * q) * q + (j & 1);
        double* dslot = Z + zt_base(q) + 3 * (j & 1);       // 1 / Z[j][j]
        const double ljj = zd[j], zri = *dslot;
        const double rr = ljj * ljj + xj * xj;
        const double rinv = rsqrt(rr);
        const double r = rr * rinv;
        const double s = xj * zri, ci = ljj * rinv, cc = r * zri;
        if (l > j && l < NMAIN) {
            const double lcj = (own[e] + s * x) * ci;
            x = cc * x - s * lcj;
            own[e] = lcj;
        }
        __syncwarp();
        if (l == 0) { zd[j] = r; *dslot = rinv; }
    }
    __syncwarp();
}

// Newton direction from the gradient (two-slot registers; gB already zeroed on fixed slack variables, winvB = 1/d_k on the
// free ones and 0 elsewhere).  The direction is written to the mirror `sdc` (main and slack part, pad entry untouched).
__device__ WBC_NI_NEWTON_DIRECTION void newton_direction(double gA, double gB, double winvB, unsigned freemask, int nic)
{
    const int l = threadIdx.x & 31;
    double* sdc = wbc_smem + sl::OFF_DC;
    const double* CIc = wbc_smem + sl::OFF_CI + l;
    const double u = gB * winvB;
    double r = -gA;
#pragma unroll 1
    for (unsigned m = freemask; m; m &= m - 1u) {
        const int k = __ffs((int)m) - 1;
        r += CIc[k * LDH] * bshfl(u, k);
    }
    // lanes past the slack count read a few entries of this vector as the (discarded) tail of the previous product's
    // operand (symv reads x[30 + lane] unconditionally): order those reads before the overwrite
    __syncwarp();
    if (l < NMAIN) sdc[l] = r;
    __syncwarp();
    tri_solve30(sdc);
    const double* row = wbc_smem + sl::OFF_CI + (l < nic ? l : 0) * LDH;
    double t0 = 0.0, t1 = 0.0;
#pragma unroll 1
    for (int j = 0; j < NMAIN; j += 2) {
        const double2 dv = ld2(sdc + j);
        t0 += row[j] * dv.x; t1 += row[j + 1] * dv.y;
    }
    if (l < nic) sdc[NMAIN + l] = -(gB + (t0 + t1)) * winvB;
    __syncwarp();
}

__device__ __forceinline__ void red3(double& a, double& b, double& c)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o);
        b += __shfl_xor_sync(FULL, b, o);
        c += __shfl_xor_sync(FULL, c, o);
    }
}
__device__ __forceinline__ double red1(double a)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(FULL, a, o);
    return a;
}

// qqpsolver_quadraticmodel (opt.cpp:30753-30821) on two-slot registers.  d is mirrored at DC.  Returns the packed sign
// estimates (see estimateparabolicmodel); d1, d2 are left in SP[4..5].
__device__ WBC_NI_QUADRATIC_MODEL int quadratic_model(double dA, double dB, double gA, double gB, double xcA, double xcB, int nic2, double rho,
                                            double absasum, double absasum2, double mb)
{
    const double2 ed = symv(wbc_smem + sl::OFF_DC, nic2, rho);
    double s0 = dA * ed.x + dB * ed.y;       // invalid lanes carry d = 0
    double s1 = dA * gA + dB * gB;
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(FULL, s0, o);
        s1 += __shfl_xor_sync(FULL, s1, o);
    }
    const double m0 = wmax_nn(fmax(fabs(xcA), fabs(xcB))), m1 = wmax_nn(fmax(fabs(dA), fabs(dB)));
    const double d2 = 0.5 * s0, d1 = s1;
    if ((threadIdx.x & 31) == 0) {
        wbc_smem[sl::OFF_SP + 4] = d1;
        wbc_smem[sl::OFF_SP + 5] = d2;
    }
    __syncwarp();
    return estimateparabolicmodel(absasum, absasum2, m0, mb, m1, d1, d2);
}

// sasexploredirection (opt.cpp:27433-27528) on slot B registers (see sas_explore_direction in qp_warp.cuh).
// Returns cidx (-1: no blocking bound); the step is left in SP[6].
__device__ WBC_NI_EXPLORE int explore(double xcB, double dB, int candB)
{
    const int l = threadIdx.x & 31;
    double best = BIGSTEP;
    if (candB && dB < 0.0) {
        // safeminposrv(x, y, BIGSTEP), alglibinternal.cpp:1998
        const double y = -dB;
        if (y >= 1.0 || xcB < BIGSTEP * y) {
            const double r = ddiv(xcB, y);
            if (r < best) best = r;
        }
    }
    // arg-min over non-negative doubles by bit pattern, ties to the lowest variable index
    const unsigned hi = (unsigned)__double2hiint(best), lo = (unsigned)__double2loint(best);
    const unsigned mh = __reduce_min_sync(FULL, hi);
    const unsigned ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
    const unsigned bi = __reduce_min_sync(FULL, (hi == mh && lo == ml) ? (unsigned)(NMAIN + l) : 0x7fffffffu);
    const double bestall = __hiloint2double((int)mh, (int)ml);
    if (l == 0) wbc_smem[sl::OFF_SP + 6] = bestall;
    __syncwarp();
    return (bestall < BIGSTEP) ? (int)bi : -1;
}

// Step selection after the quadratic model (opt.cpp:30259-30302 / 30440-30487), qqpsolver_findbeststepandmove
// (30882-31003) and sasmoveto (27574-27723) on two-slot registers; xc is read from and written back to its shared
// mirror XC.  mode 0: d2est > 0 (full step unless a bound blocks); 1: non-positive curvature in the CG phase
// (step to the bound); 2: the Newton phase's bound step with candidates {4 stpmax, 1, 0.25}.
// Returns the new cstatus of slot B; SP[7] gets the number of extra model evaluations.  (The trial points of eval4 overlay
// the mirror of x, which is read before and rewritten after.)
__device__ WBC_NI_STEP_AND_MOVE int step_and_move(double dA, double dB, double exbA, double exbB, int nic, double rho, int mode, int cidx, int csB)
{
    const int l = threadIdx.x & 31;
    double* xs = wbc_smem + sl::OFF_XC;
    const double* sp = wbc_smem + sl::OFF_SP;
    const double d1 = sp[4], d2 = sp[5], stpmax = sp[6];
    double xcA = (l < NMAIN) ? xs[l] : 0.0, xcB = (l < nic) ? xs[NMAIN + l] : 0.0;      // (lanes past the vector read nothing: lane 0 rewrites xs[0] below)
    double stp, a0 = 0.0, a1 = 0.0, a2 = 0.0;
    bool needact;
    int addcnt;
    if (mode == 0) {
        const double fullstp = ddiv(-d1, 2 * d2);
        needact = fullstp >= stpmax;
        if (needact) { stp = stpmax; a0 = stpmax * 4; a1 = fullstp; a2 = fullstp * 0.25; addcnt = 3; }
        else { stp = fullstp; addcnt = 0; }
    } else if (mode == 1) {
        stp = stpmax; needact = true; a0 = 4 * stpmax; addcnt = 1;
    } else {
        stp = stpmax; needact = true; a0 = stpmax * 4; a1 = 1.00; a2 = 0.25; addcnt = 3;
    }
    double stpbest = stp;
    if (addcnt > 0) {
        eval4(xcA, xcB, dA, dB, exbA, exbB, nic, rho, stp, a0, addcnt > 1 ? a1 : a0, addcnt > 2 ? a2 : a0);
        const double2 f01 = ld2(sp), f23 = ld2(sp + 2);
        double fbest = f01.x;
        if (a0 > stp && f01.y < fbest) { fbest = f01.y; stpbest = a0; }
        if (addcnt > 1 && a1 > stp && f23.x < fbest) { fbest = f23.x; stpbest = a1; }
        if (addcnt > 2 && a2 > stp && f23.y < fbest) { fbest = f23.y; stpbest = a2; }
    }
    xcA = xcA + stpbest * dA;
    {
        const double old = xcB;
        double v = old + stpbest * dB;
        if (v < 0.0) v = 0.0;
        if (needact && (NMAIN + l) == cidx) { v = 0.0; csB = 1; }
        if (v <= 0.0 && v != old) { v = 0.0; csB = 1; }
        xcB = (l < nic) ? v : 0.0;
    }
    if (l < NMAIN) xs[l] = xcA;
    if (l < ((nic + 1) & ~1)) xs[NMAIN + l] = xcB;
    if (l == 0) wbc_smem[sl::OFF_SP + 7] = (double)addcnt;
    __syncwarp();
    return csB;
}

// |A| statistics of E (opt.cpp:29893-29915, with its k = (i==v ? 1 : 2) quirk) over the upper triangle, and max|exb|.
// Results in SP[8..10].
__device__ __noinline__ void qqp_stats(int nic, double rho, double exbA, double exbB)
{
    const int l = threadIdx.x & 31;
    const double* H = wbc_smem + sl::OFF_H;
    const double* CI = wbc_smem + sl::OFF_CI;
    double s1 = 0.0, s2 = 0.0;
    if (l < NMAIN) {
#pragma unroll 1
        for (int j = 0; j < NMAIN; j++) {
            // the upper triangle of row l, read as H[j][l] (generate_ex_model writes both with the same value) with every lane on
            // the same j: the lanes walk a row of H side by side.  (Row l from its diagonal on -- lane l at H[l][l + t] or at
            // H[l + t][l] -- puts all thirty lanes on one bank: that was a 16-way conflict, a quarter of the kernel's replays.)
            const double v = H[j * LDH + l], vv = fabs(v);
            const double k = ((double)l == v) ? 1.0 : 2.0;
            if (j >= l) { s1 += vv * k; s2 += vv * vv * k; }
        }
#pragma unroll 1
        for (int kk = 0; kk < nic; kk++) {
            const double v = CI[kk * LDH + l], vv = fabs(v);
            const double k = ((double)l == v) ? 1.0 : 2.0;
            s1 += vv * k; s2 += vv * vv * k;
        }
    }
    if (l < nic) {
        const double vv = fabs(rho);
        const double k = ((double)(NMAIN + l) == rho) ? 1.0 : 2.0;
        s1 += vv * k; s2 += vv * vv * k;
    }
    s1 = wsum(s1); s2 = wsum(s2);
    const double m1 = wmax_nn(fmax(fabs(exbA), fabs(exbB)));
    if (l == 0) {
        double* sp = wbc_smem + sl::OFF_SP;
        sp[8] = s1; sp[9] = s2; sp[10] = m1;
    }
    __syncwarp();
}

// Diagonal of the constrained-Newton model with its regulariser 1e-9 * sum_j |A_ff[i][j]| over free j (opt.cpp:31150-31167),
// in two-slot form; fixed variables get 1.  fmask: ballot of the free slack variables.
__device__ WBC_NI_NEWTON_DIAG double2 newton_diag(int nic, double rho, unsigned fmask)
{
    const int l = threadIdx.x & 31;
    const double* H = wbc_smem + sl::OFF_H;
    const double* CI = wbc_smem + sl::OFF_CI;
    double2 r = make_double2(0.0, 1.0);
    if (l < NMAIN) {
        double v = 0.0;
#pragma unroll 1
        for (int j = 0; j < NMAIN; j++) v += fabs(H[j * LDH + l]);
#pragma unroll 1
        for (int k = 0; k < nic; k++)
            if ((fmask >> k) & 1u) v += fabs(CI[k * LDH + l]);
        if (v == 0.0) v = 1.0;
        r.x = H[l * LDH + l] + 1.0e-9 * v;
    }
    if (l < nic && ((fmask >> l) & 1u)) {
        const double* row = CI + l * LDH;
        double v = 0.0;
#pragma unroll 1
        for (int i = 0; i < NMAIN; i++) v += fabs(row[i]);
        v += fabs(rho);
        if (v == 0.0) v = 1.0;
        r.y = rho + 1.0e-9 * v;
    }
    return r;
}

// One QQP solve from the point exxc (in/out) on the model (H, CI, rho, exb).  Returns the QQP termination type.
// pre_stats: the model's (absasum, absasum2, max|exb|) when the caller has them already (stage tasks), else NULL.
__device__ WBC_NI_QQP_OPTIMIZE_FAST int qqp_optimize_fast(const Work w, int nic, double rho, double epsx, int maxouterits, int* ncholesky, double* flops_io, int* reused_io,
                                              const double* pre_stats)
{
    const int l = threadIdx.x & 31;
    const int n = NMAIN + nic;
    const int nic2 = (nic + 1) & ~1;
    const bool vA = l < NMAIN, vB = l < nic;
    int nsymv = 0;                      // products with E (2 n^2 flops each), for the instrumented flop count
    int nchol = 0, nfree = 0, cnmodelage = 0;
    int nreused = 0;                    // factorisations skipped because the factor in memory was already the answer
    bool fac_ok = false;
    unsigned fac_mask = 0u;
    double winvB = 0.0;
    int nschur = 0, nfix = 0;           // slack rows folded into Schur complements / solves, rank-one fixes (flop count)
    double* sxc = wbc_smem + sl::OFF_XC;
    double* sdc = wbc_smem + sl::OFF_DC;
    const double* spare = wbc_smem + sl::OFF_SP;
    const double* exb = wbc_smem + sl::OFF_EXB;
    double* exxc = wbc_smem + sl::OFF_EXXC;
    // settings: qqploaddefaults (opt.cpp:29533-29547) + overrides (41318-41323)
    const int cgminits = 5;
    int cgmaxits = (int)(1 + 0.33 * n + 0.5);
    if (cgmaxits < cgminits) cgmaxits = cgminits;
    const int cnmaxupdates = (int)(1 + 0.1 * n + 0.5);
    (void)w;

    const double exbA = vA ? exb[l] : 0.0, exbB = vB ? exb[NMAIN + l] : 0.0;
    // start point clipped to the bounds (29979-29998) and sasstartoptimization (27377-27399)
    double xcA = vA ? exxc[l] : 0.0, xcB = 0.0;
    int csB = -1;
    if (vB) {
        double v = exxc[NMAIN + l];
        if (v <= 0.0) { v = 0.0; csB = 0; }
        xcB = v;
    }
    if (vA) sxc[l] = xcA;
    if (l < nic2) { sxc[NMAIN + l] = xcB; sdc[NMAIN + l] = 0.0; }
    double absasum, absasum2, mb;
    if (pre_stats) { absasum = pre_stats[0]; absasum2 = pre_stats[1]; mb = pre_stats[2]; __syncwarp(); }     // (the mirrors of x written above are read by every lane)
    else { qqp_stats(nic, rho, exbA, exbB); absasum = spare[8]; absasum2 = spare[9]; mb = spare[10]; }
    int term = 0;
    int cgmax = cgminits;
    int outerits = 0;
    double xpA = 0.0, xpB = 0.0;
#pragma unroll 1
    for (;;) {
        if (maxouterits > 0 && outerits >= maxouterits) { term = 5; break; }
        if (outerits > 0) {
            // epsx stopping test (30137-30149)
            const double ta = xpA - xcA, tb = xpB - xcB;
            const double v = wsum(ta * ta + tb * tb);
            if (dsqrt(v) <= epsx) { term = 2; break; }
        }
        outerits++;
        xpA = xcA; xpB = xcB;
        double dpA = 0.0, dpB = 0.0, vprev = 0.0;       // previous direction, |previous constrained gradient|^2
#pragma unroll 1
        for (int cgcnt = 0; cgcnt <= cgmax - 1; cgcnt++) {
            const double2 ex = symv(sxc, nic2, rho);                            // targetgradient
            const double gA = vA ? ex.x + exbA : 0.0, gB = vB ? ex.y + exbB : 0.0;
            // sasreactivateconstraints (28992-29047), constrained gradient, CG coefficients (30199-30221)
            const bool atb = vB && xcB == 0.0;
            const bool act = atb && gB >= 0.0;
            csB = act ? 1 : -1;
            const double cgA = gA, cgB = act ? 0.0 : gB;
            // |cg_prev|^2 is the previous iteration's |cg|^2 (the same sum of the same numbers), 0 at the start of the phase
            const double v = wsum(cgA * cgA + cgB * cgB), vv = vprev;
            const bool bf = __any_sync(FULL, atb && dpB != 0.0);
            if (v <= 0.0) { term = 4; break; }      // sqrt(v) <= 0 with v a sum of squares
            const bool brst = bf || (vv == 0.0) || (cgcnt % 50 == 0);
            const double beta = brst ? 0.0 : ddiv(v, vv);
            const double dA = vA ? -cgA + beta * dpA : 0.0;
            const double dB = (vB && !act) ? -cgB + beta * dpB : 0.0;
            __syncwarp();                               // the product above read (and dropped) a few entries of this vector as the tail of x
            if (vA) sdc[l] = dA;
            if (vB) sdc[NMAIN + l] = dB;
            __syncwarp();
            const int cidx = explore(xcB, dB, vB && csB <= 0);
            const int code = quadratic_model(dA, dB, gA, gB, xcA, xcB, nic2, rho, absasum, absasum2, mb);
            const int d1est = (code >> 2) - 1, d2est = (code & 3) - 1;
            const double d1 = spare[4], d2 = spare[5];
            nsymv += 2;
            if (d1 == 0.0 && d2 == 0.0) { term = 4; break; }
            if (d1est >= 0) { term = 7; break; }
            if (d2est <= 0 && cidx < 0) { term = -4; break; }
            csB = step_and_move(dA, dB, exbA, exbB, nic, rho, d2est > 0 ? 0 : 1, cidx, csB);
            xcA = vA ? sxc[l] : 0.0; xcB = vB ? sxc[NMAIN + l] : 0.0;
            if (spare[7] > 0.0) nsymv += 1 + (int)spare[7];
            dpA = dA; dpB = dB; vprev = v;
        }
        if (term != 0) break;
        cgmax = cgmaxits;
        // constrained Newton phase (30353-30527)
        int newtcnt = 0;
        int freeB = 0;
#pragma unroll 1
        for (;;) {
            bool b;
            if (newtcnt == 0) {
                // qqpsolver_cnewtonbuild (31058-31201): free set, regularised diagonal, factorisation (slack block first)
                freeB = vB ? !(xcB == 0.0) : 0;
                const unsigned fmask = __ballot_sync(FULL, freeB != 0);
                nfree = NMAIN + __popc(fmask);
                cnmodelage = 0;
                nchol++;
                if (fac_ok && fmask == fac_mask) {
                    // same model (H, CI, rho are fixed during this solve) and same free set as the factor in memory, which
                    // no rank-one fix has touched: the factorisation would reproduce it bit for bit
                    nreused++;
                    b = true;
                } else {
                    nschur += __popc(fmask);
                    const double2 dg = newton_diag(nic, rho, fmask);
                    const double rsB = freeB ? rsqrt(dg.y) : 0.0;
                    winvB = rsB * rsB;
                    b = chol_build30(dg.x, rsB, fmask);
                    fac_ok = b;
                    fac_mask = fmask;
                }
                if (b) cgmax = cgminits;
            } else {
                // qqpsolver_cnewtonupdate (31314-31426)
                const bool tofix = vB && freeB && xcB == 0.0;
                const unsigned fixmask = __ballot_sync(FULL, tofix);
                const int ntofix = __popc(fixmask);
                nsymv += 2;
                if (ntofix == 0 || ntofix == nfree) b = false;
                else if (cnmodelage + ntofix > cnmaxupdates) b = false;
                else {
#pragma unroll 1
                    for (unsigned m = fixmask; m; m &= m - 1u) rank1_fix30(__ffs((int)m) - 1);
                    fac_ok = false;                 // the factor no longer equals a fresh factorisation of any free set
                    if (tofix) { freeB = 0; winvB = 0.0; }
                    nfree -= ntofix;
                    cnmodelage += ntofix;
                    nfix += ntofix;
                    b = true;
                }
            }
            if (!b) break;
            newtcnt++;
            const double2 ex = symv(sxc, nic2, rho);
            const double gA = vA ? ex.x + exbA : 0.0, gB = vB ? ex.y + exbB : 0.0;
            // qqpsolver_cnewtonstep (31474-31536), epsg = 0
            const double ngA = gA, ngB = (vB && freeB) ? gB : 0.0;
            const double gg = wsum(ngA * ngA + ngB * ngB);
            if (gg <= 0.0) break;
            newton_direction(ngA, ngB, winvB, __ballot_sync(FULL, freeB != 0), nic);
            nschur += 2;
            const double dA = vA ? sdc[l] : 0.0, dB = vB ? sdc[NMAIN + l] : 0.0;
            const int code = quadratic_model(dA, dB, gA, gB, xcA, xcB, nic2, rho, absasum, absasum2, mb);
            const int d1est = (code >> 2) - 1, d2est = (code & 3) - 1;
            nsymv += 3;
            if (d1est >= 0) break;
            const int cidx = explore(xcB, dB, vB && csB <= 0);
            if (d2est > 0) {
                csB = step_and_move(dA, dB, exbA, exbB, nic, rho, 0, cidx, csB);
                if (spare[7] > 0.0) nsymv += 1 + (int)spare[7];
            } else {
                const double stpmax = spare[6];
                if (cidx < 0) { term = -4; break; }
                if (stpmax == 0.0) { cgmax = cgmaxits; break; }
                // f(x) vs f(x + stpmax d) (30493-30503)
                eval4(xcA, xcB, dA, dB, exbA, exbB, nic, rho, 0.0, stpmax, stpmax, stpmax);
                const double2 f01 = ld2(spare);
                if (vA) sxc[l] = xcA;                   // the trial points overlay the mirror of x: put it back
                if (l < nic2) sxc[NMAIN + l] = xcB;
                __syncwarp();
                if (f01.y >= f01.x) { cgmax = cgmaxits; break; }
                csB = step_and_move(dA, dB, exbA, exbB, nic, rho, 2, cidx, csB);
                nsymv += 6;
            }
            xcA = vA ? sxc[l] : 0.0; xcB = vB ? sxc[NMAIN + l] : 0.0;
        }
        if (term != 0) break;
    }
    // unpack (30546-30565)
    if (vA) exxc[l] = xcA;
    if (vB) exxc[NMAIN + l] = (xcB < 0.0 || xcB == 0.0) ? 0.0 : xcB;
    __syncwarp();
    *ncholesky += nchol;
    *reused_io += nreused;
    // work actually done: products with E, 30^3/3 per factorisation, 30^2 per slack row folded in, 3 * 30^2 per rank-one fix
    *flops_io += 2.0 * n * n * (double)nsymv + (double)(nchol - nreused) * 9000.0 + 900.0 * (double)nschur + 2700.0 * (double)nfix;
    return term;
}

}  // namespace fast
}  // namespace wbcqp
#endif
