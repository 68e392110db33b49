// Batched dense QP solver for the WBC ground-reaction-force QP: a decision-for-decision restatement
// of the algorithm the reference controller runs through ALGLIB 3.16.0
//   minqpoptimize (opt.cpp:48020) -> qpdenseauloptimize (41087) -> qqpoptimize (29675)
// with the reference's settings (lopt.cpp:91-106: autodiag scaling, DENSE-AUL epsx=1e-2, rho=1e4,
// 5 outer iterations, cold start, no box constraints on x).  "opt.cpp" = the reference's
// dogbot_controller/src/alglib/optimization.cpp, "linalg.cpp" likewise.
//
// Execution model: ONE WARP PER QP, one warp per CTA, 18.3 KB of shared memory per warp (12 resident
// per SM).  The design is driven by what ncu showed on the earlier CTA-per-instance version (profiles/):
// the path is bound by instruction issue and instruction fetch, not by the FP64 pipe or by memory --
// so every uniform scalar decision is executed once (one warp), there are no block barriers or
// cross-warp reductions, and nothing is passed between routines by reference (no local-memory traffic):
// storage is addressed as constant offsets from the CTA's shared-memory base and from one global scratch
// pointer.  Which routines are calls and which are inlined was decided by measurement, twice: round 1 made
// all of them calls (code size), round 2's last step inlined all but the product (qp_fast.cuh).
// `HostEx` (one lane) lets the same source be compiled by g++ for CPU-side unit tests
// (tests/host_emu); the shipped library only instantiates `WarpEx`.
//
// Specialisation relative to generic ALGLIB (all other cases cannot occur on this path):
//   * dense A, no sparse constraints, x unbounded, start point 0, origin 0;
//   * hence in QQP the only bounds are "slack >= 0" on variables i >= NMAIN.
//
// Storage (shared memory is the occupancy limiter: 18.3 KB per warp -> 12 resident warps per SM):
//   * the QQP quadratic term E = [H, rho Ci'; rho Ci, rho I] is never formed: H = A + rho C'C (30 x 30,
//     full symmetric, ld 31) and CI = rho * (working inequality rows) (nic x 30, ld 31) are kept and the
//     products with E use the block structure (same number of multiply-adds as the dense n x n form);
//   * the Cholesky factor of the constrained-Newton phase is the 30 x 30 Schur complement only (qp_fast.cuh),
//     stored transposed and pair-interleaved so that every access of the factorisation and of the forward
//     sweep is a conflict-free 128-bit load;
//   * the constraint matrix C (88 x 31) lives in the warp's global scratch (L2-resident) and is staged
//     through shared memory where it is used densely;
//   * the same shared arrays are the workspace of the set-up and of the multiplier update, and the arrays that
//     are idle inside the QQP iteration (exb, exxc, the mirrors of x and d) carry the four trial points of the
//     line search;
//   * instances whose working set outgrows NICCAP switch CI, the factor and the QQP vectors to a
//     global-memory spill copy (same code, SPILL = true instantiation).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WBC_HD __host__ __device__ __forceinline__
#define WBC_HDN __host__ __device__
#define WBC_HDNI __host__ __device__ __noinline__
// groups of routines that are inlined on the device since round 2 (qp_fast.cuh, "Which routines are separate functions");
// -DWBC_OUTLINE_G / _M / _PARAB / _TSW make a group calls again.  G: generate_ex_model, update_working_set, inequality_violations,
// feasibility_error; M: update_lagrange_multipliers_reduced, setup_problem; PARAB: estimateparabolicmodel; TSW: tri_solve_warp32.
#ifdef WBC_OUTLINE_G
#define WBC_HDNI_G WBC_HDNI
#else
#define WBC_HDNI_G __host__ __device__ __forceinline__
#endif
#ifdef WBC_OUTLINE_M
#define WBC_HDNI_M WBC_HDNI
#else
#define WBC_HDNI_M __host__ __device__ __forceinline__
#endif
#ifdef WBC_OUTLINE_PARAB
#define WBC_HDNI_PARAB WBC_HDNI
#else
#define WBC_HDNI_PARAB __host__ __device__ __forceinline__
#endif
#ifdef WBC_OUTLINE_TSW
#define WBC_NI_TSW __noinline__
#else
#define WBC_NI_TSW __forceinline__
#endif
// the small shared helpers (warp sum / max, square root, division): inlined (qp_fast.cuh, "Which routines are separate functions");
// -DWBC_SMALL_OUTLINE makes them calls again
#ifdef WBC_SMALL_OUTLINE
#define WBC_SMALL_NI __noinline__
#else
#define WBC_SMALL_NI __forceinline__
#endif
#else
#define WBC_HD inline
#define WBC_HDN
#define WBC_HDNI
#define WBC_HDNI_G
#define WBC_HDNI_M
#define WBC_HDNI_PARAB
#endif

namespace wbcqp {

constexpr int NMAIN = 30;              // decision variables (main.cpp:266 OPT(30,86,82))
constexpr int MAXK = 88;               // >= 86 constraint rows (stance), 82 (swing)
constexpr int MAXNIC = 72;             // >= 68 / 70 inequality rows
constexpr int MAXNT = NMAIN + MAXNIC;  // extended variable count upper bound
constexpr int NICCAP = 18;             // working inequality rows held in shared memory
constexpr int NCAP = NMAIN + NICCAP;   // 48
constexpr int LDH = 31;                // leading dimension of H, CI, C, A (odd: conflict-free both ways)
constexpr int KACAP = 36;              // largest active set the reduced multiplier update handles
constexpr int NVEC = 16;               // QQP vectors
constexpr int VLS = 48, VLG = 104;     // vector length: shared / spill copy
constexpr double MACHEPS = 5.0e-16;    // ae_machineepsilon (ap.cpp: 5E-16, NOT DBL_EPSILON)
constexpr double BIGSTEP = 1.0e50;     // opt.cpp:27458

// packed lower triangle, row c holds entries (c, k), k < c.  No padding: consecutive rows start at consecutive triangular
// numbers, which are distinct modulo 16 over (almost) any 16 consecutive rows -- a lane-per-row access is conflict free.
WBC_HD int zoff(int c) { return (c * (c - 1)) >> 1; }

// ---- shared-memory layout of one warp (doubles).  Order matters:
//   [H | Z | ZD | XC | DC | EXB | SP | RS]  is one contiguous workspace (sl::BIG doubles) for the multiplier update;
//   [Z .. RS | CI]                          is the staging area for rows of C: generate_ex_model places the working rows so
//                                           that the inequality rows land in CI itself (scaled by rho in place);
//   [XC | DC | EXB] and EXXC                carry the four trial points of the QQP line search (qp_fast.cuh, eval4), which
//                                           never needs them at the same time: exb and exxc live in registers inside a
//                                           QQP call, the mirrors of x and d are rewritten after every line search.
// The host emulation (one lane, generic QQP) needs the full set of QQP vectors and a factor of the whole (30 + NICCAP)-
// dimensional Newton model; they are appended behind the device layout.
namespace sl {
#if defined(__CUDACC__)
constexpr int Z_DOUBLES = 480;                          // fast path: pair-interleaved transposed 30 x 30 factor (qp_fast.cuh)
#else
constexpr int Z_DOUBLES = 1152;                         // generic path: packed factor, zoff(48) = 1128
#endif
constexpr int OFF_H = 0;                                // [30][31]
constexpr int OFF_Z = OFF_H + NMAIN * LDH;              // factor of the Newton model; packed factor of A during set-up
constexpr int OFF_ZD = OFF_Z + Z_DOUBLES;               // [32] diagonal of the Newton factor (fast path)
constexpr int OFF_XC = OFF_ZD + 32;                     // [48] mirror of the QQP point; at the end of a solve: the result x
constexpr int OFF_DC = OFF_XC + 48;                     // [48] mirror of the QQP direction
constexpr int OFF_EXB = OFF_DC + 48;                    // [48] linear term of the extended model (shared-memory case)
constexpr int OFF_SP = OFF_EXB + 48;                    // [16] scalars handed between the QQP routines
constexpr int OFF_RS = OFF_SP + 16;                     // [20] 1/sqrt(d_k) of the free slack variables (fast path)
constexpr int OFF_CI = OFF_RS + 20;                     // [18][31]
constexpr int OFF_EXXC = OFF_CI + NICCAP * LDH;         // [104] point of the extended model
constexpr int OFF_INT = OFF_EXXC + 104;                 // ints: iscr[8]
constexpr int TOTAL_DEV = OFF_INT + 4;
constexpr int BIG = OFF_CI;                             // contiguous workspace in front of CI
constexpr int STAGE0 = OFF_Z;                           // staging area for rows of C: [STAGE0, OFF_EXXC)
constexpr int STAGE_EQ_CAP = OFF_CI - STAGE0;           // room for the equality rows in front of CI
constexpr int STAGE_CAP = OFF_EXXC - STAGE0;
#if defined(__CUDACC__)
constexpr int TOTAL = TOTAL_DEV;
#else
constexpr int OFF_V = TOTAL_DEV;                        // [NVEC][48] QQP vectors of the generic path
constexpr int TOTAL = OFF_V + NVEC * VLS;
#endif
constexpr int BYTES = TOTAL * 8;
static_assert(OFF_Z % 2 == 0 && OFF_ZD % 2 == 0 && OFF_XC % 2 == 0 && OFF_EXXC % 2 == 0 && OFF_SP % 2 == 0, "16-byte alignment of the arrays read with 128-bit loads");
#if defined(__CUDACC__)
static_assert(12 * (BYTES + 1024) <= 228 * 1024, "12 resident solver warps per SM");
#endif
}  // namespace sl
// ---- global scratch layout of one warp (doubles)
namespace gl {
constexpr int NQMAX = MAXNT + MAXK;
constexpr long OFF_C = 0;                                        // [88][31]
constexpr long OFF_A = OFF_C + MAXK * LDH;                       // [30][31] scaled A, full symmetric
constexpr long OFF_LA = OFF_A + 944;                             // packed factor of A (transposed: row c = U[.][c]), zoff(30) = 435, then
constexpr int LA_DOUBLES = 436;                                  //   its reciprocal diagonal [32] at OFF_LA + LA_DOUBLES (one copy fetches both)
constexpr long OFF_B = OFF_LA + 480;                             // [32] scaled linear term
constexpr long OFF_SC = OFF_B + 32;                              // [32] variable scales
constexpr long OFF_NICERR = OFF_SC + 32;                         // [72]
constexpr long OFF_NULC = OFF_NICERR + MAXNIC;                   // [88]
constexpr long OFF_NULCEST = OFF_NULC + MAXK;                    // [88]
constexpr long OFF_NICNACT = OFF_NULCEST + MAXK;                 // [72] ints
constexpr long OFF_CSTAT = OFF_NICNACT + 40;                     // ints: cstatus[104] | isfree[104] of the generic QQP
constexpr long OFF_EXBG = OFF_CSTAT + 104;                       // spill [104] linear term of the extended model
constexpr long OFF_CI = OFF_EXBG + 104;                          // spill [72][31]
constexpr long OFF_Z = OFF_CI + MAXNIC * LDH + 8;                // spill packed, zoff(102) = 5202
constexpr long OFF_V = OFF_Z + 5216;                             // spill [16][104]
constexpr long OFF_QRV = OFF_V + NVEC * VLG;
constexpr long OFF_SV0 = OFF_QRV + 2 * NQMAX + 4;
// Cache of the reduced multiplier update (update_lagrange_multipliers_reduced): everything that depends only on the
// active set -- kept between the outer iterations of one solve, which usually end on the same active set.
constexpr long OFF_MC = ((OFF_SV0 + NQMAX + 15) / 16) * 16;
constexpr int MC_C0 = 0;            // [40]  d_m + W_m . t
constexpr int MC_S0D = 40;          // [40]  diagonal of S = W W'
constexpr int MC_S0 = 80;           // [648] strict lower triangle of S, packed
constexpr int MC_SFD = 728;         // [40]  diagonal of the factor of S
constexpr int MC_SFRINV = 768;      // [40]  its reciprocal
constexpr int MC_SF = 808;          // [648] factor of S, packed
constexpr int MC_GD = 1456;         // [40]  factor of G = Lt'Lt (rank-deficient active sets only)
constexpr int MC_GRINV = 1496;      // [40]
constexpr int MC_G = 1536;          // [648]
constexpr int MC_INT = 2184;        // ints: act[40] | dep[40] | ka, version, ndep
constexpr int MC_TOTAL = 2240;
// State of a solve that is handed from stage to stage (solve_stage_* below): the SolveState header and the point exxc.
constexpr long OFF_HDR = OFF_MC + MC_TOTAL;                      // [32] header | [104] exxc
constexpr int HDR_DOUBLES = 32;
// The extended model of a QQP call handed to the warp that runs it (stage tasks): H [30][31]; its CI rows go to the spill
// CI array and its linear term to the spill EXB array (the spill mode never hands its QQP over).
constexpr long OFF_HQ = OFF_HDR + HDR_DOUBLES + 104;
constexpr long TOTAL = ((OFF_HQ + NMAIN * LDH + 15) / 16) * 16;
// The literal multiplier update's stacked KKT matrix is not part of the block: 2 NQMAX (NQMAX + 1) doubles (580 KB) that
// only the 0.25 % of solves which fall back to it ever touch -- one per resident warp (Work::kkt), not one per solve in flight.
constexpr long KKT_DOUBLES = ((2L * NQMAX * (NQMAX + 1) + 15) / 16) * 16;
}  // namespace gl

// All storage of one instance: the CTA's shared memory (device) and one global scratch block.
struct Work {
    double* g;      // global scratch block of the solve (gl::TOTAL doubles)
    double* kkt;    // global scratch of this warp for the literal multiplier update (gl::KKT_DOUBLES doubles)
    double* sm;     // host emulation only: heap stand-in for the shared memory block
};
#if defined(__CUDACC__)
extern __shared__ __align__(16) double wbc_smem[];
#endif
#if defined(__CUDA_ARCH__)
#define WBC_SM(w) (wbc_smem)
#else
#define WBC_SM(w) ((w).sm)
#endif
#define SM_(w, off) (WBC_SM(w) + (off))
#define W_H(w) SM_(w, sl::OFF_H)
#define W_BIG(w) SM_(w, 0)
#define W_B(w) ((w).g + gl::OFF_B)
#define W_SC(w) ((w).g + gl::OFF_SC)
#define W_LARINV(w) ((w).g + gl::OFF_LA + gl::LA_DOUBLES)
#define W_NICERR(w) ((w).g + gl::OFF_NICERR)
#define W_NULC(w) ((w).g + gl::OFF_NULC)
#define W_NULCEST(w) ((w).g + gl::OFF_NULCEST)
#define W_EXXC(w) SM_(w, sl::OFF_EXXC)
#define W_EXB(w) SM_(w, sl::OFF_EXB)
#define W_XS(w) SM_(w, sl::OFF_XC)
#define W_SPARE(w) SM_(w, sl::OFF_SP)
#define W_NICNACT(w) (reinterpret_cast<int*>((w).g + gl::OFF_NICNACT))
#define W_CSTATUS(w) (reinterpret_cast<int*>((w).g + gl::OFF_CSTAT))
#define W_ISFREE(w) (reinterpret_cast<int*>((w).g + gl::OFF_CSTAT) + 104)
#define W_ISCR(w) (reinterpret_cast<int*>(SM_(w, sl::OFF_INT)))
#define W_C(w) ((w).g + gl::OFF_C)
#define W_A(w) ((w).g + gl::OFF_A)
#define W_LA(w) ((w).g + gl::OFF_LA)

// QQP storage selector: shared memory, or the global spill copy.  The generic QQP's vectors only exist on the host
// (shared-memory case) and in the spill copy; on the device the shared-memory case is the register-resident QQP of
// qp_fast.cuh, which uses the named arrays of `sl` directly.
template <bool SPILL>
struct QS {
    static constexpr int VL = SPILL ? VLG : VLS;
    static WBC_HD double* CI(const Work& w) { return SPILL ? w.g + gl::OFF_CI : SM_(w, sl::OFF_CI); }
    static WBC_HD double* Z(const Work& w) { return SPILL ? w.g + gl::OFF_Z : SM_(w, sl::OFF_Z); }
    static WBC_HD double* EXB(const Work& w) { return SPILL ? w.g + gl::OFF_EXBG : SM_(w, sl::OFF_EXB); }
#if defined(__CUDACC__)
    static WBC_HD double* V(const Work& w, int k)
    {
        static_assert(SPILL, "the generic QQP's shared-memory vectors do not exist on the device");
        return w.g + gl::OFF_V + k * VL;
    }
#else
    static WBC_HD double* V(const Work& w, int k) { return (SPILL ? w.g + gl::OFF_V : SM_(w, sl::OFF_V)) + k * VL; }
#endif
};
// vector slots
enum { V_ZD = 0, V_ZRINV, V_XC, V_DC, V_T0, V_T1, V_T2, V_T3, V_SPARE, V_XP, V_GC, V_CGC, V_CGP, V_DP, V_BUFR, V_REG };

struct Settings {
    double epsx = 1.0e-2;   // lopt.cpp:101
    double rho = 1.0e4;
    int outerits = 5;
    int kkt_mode = 1;       // 1 = reduced multiplier update with the literal form as fallback; 0 = literal only
    // Relative pivot below which an active row counts as dependent in the reduced multiplier update; pivots in
    // (1e-3 * pivtol, pivtol] are "ambiguous" and send the update to the literal form.  Measured against the ALGLIB
    // oracle on 12 000 instances: 0.25 % of swing solves have pivots of 2.8e-8 .. 3.4e-8 there (a pair of opposite
    // swing-tracking rows both active, separated only by the slack column that autodiag scales by 1e-4).  With
    // pivtol = 1e-9 the reduced form keeps those rows and the fallback disappears -- and 1 in 1000 swing torques is off
    // by 1e-5: on exactly these components the literal form's Tikhonov term is not negligible.  So the band stays.
    double kkt_pivtol = 1.0e-5;
    // Hint from the caller that assembled L: rows [dup_start[t], dup_start[t] + dup_count[t]) have, coefficient for
    // coefficient, the negatives of the dup_count[t] rows just before them (upper / lower limits on the same
    // quantity).  c'Ac of a row and of its negative are the same number bit for bit, so set-up's maximum over the
    // rows skips them.  Zero counts (the default, and the dense OPT operator): nothing is assumed.
    int dup_start[2] = {0, 0};
    int dup_count[2] = {0, 0};
};

struct Stats {
    int termination;    // 2 = ok (opt.cpp:41583); -9 non-positive diagonal (48178-48181)
    int ncholesky;      // rep.ncholesky (opt.cpp:41325)
    int outer_its;      // outer AUL iterations executed
    int qqp_calls;      // inner QQP solves
    int nicwork;        // final working-set size
    int kkt_dim_max;    // largest (N+K) of the multiplier update
    int chol_reused;    // constrained-Newton factorisations (counted in ncholesky) served by the factor already in memory
    int flags;          // bit0: A not PD (42500); bit2: QQP -4; bit3: literal multiplier update used;
                        // bit4: rank-deficient active set (least-norm branch); bit5: spilled to global memory;
                        // bit6: a multiplier update reused the cached factorisation of an unchanged active set
    double flops;       // instrumented algorithmic flop count (DESIGN.md "work per solve")
};

// ------------------------------------------------------------------------------------------------
// executors
struct HostEx {
    static constexpr int NL = 1;
    WBC_HD int lane() const { return 0; }
    WBC_HD void sync() const {}
    WBC_HD double shfl(double v, int) const { return v; }
    WBC_HD double shfl_xor(double v, int) const { return v; }
    WBC_HD int shfl_xori(int v, int) const { return v; }
    WBC_HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
    WBC_HD int popc_below(unsigned m) const { (void)m; return 0; }
    WBC_HD int popc(unsigned m) const { return (int)m; }
    // bulk copy into the instance's fast storage (device: global -> shared)
    WBC_HD void copy_in(double* dst, const double* src, int n) const { for (int e = 0; e < n; e++) dst[e] = src[e]; }
    // several copies in flight at once: copy_start ... copy_start, then one copy_wait
    WBC_HD void copy_start(double* dst, const double* src, int n) const { copy_in(dst, src, n); }
    WBC_HD void copy_wait() const {}
};
#if defined(__CUDACC__)
// Bulk copy global -> shared with cp.async (LDGSTS): every element is in flight at once, one L2 round trip for the
// whole block instead of one per unrolled group of register-staged loads (ncu: these copies were 10 % of the
// kernel's stall samples at 0.6 % of its instructions).  8-byte granules: rows have an odd leading dimension.
__device__ __forceinline__ void warp_copy_start(unsigned dst_shared, const double* __restrict__ src, int n)
{
    const int l = threadIdx.x & 31;
    unsigned d = dst_shared + 8u * l;
    const double* g = src + l;
#pragma unroll 4
    for (int e = l; e < n; e += 32, d += 256u, g += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(g) : "memory");
}
// ONE shared, non-inlined copy of that loop for the stage-task kernel: it has 16 call sites (45 instructions each when
// inlined, 12 KB in all), every one of them in once-per-task code, and once-per-task code is what that kernel's instruction
// cache misses on (profiles/r02_ar_icache_misses_by_sm.txt).  Measured (profiles/r02_as_shared_copy_ab.txt): -2.7 % on the
// stage-task kernel at 65 536 instances; the one-warp-per-solve kernel at 4 096 instances LOSES 2.8 % with it, so it keeps
// the inlined form (WarpEx) and the stage-task kernel runs on WarpExS.
__device__ __noinline__ void warp_copy_start_ni(unsigned dst_shared, const double* __restrict__ src, int n) { warp_copy_start(dst_shared, src, n); }
struct WarpEx {
    static constexpr int NL = 32;
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ double shfl_xor(double v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    __device__ __forceinline__ int shfl_xori(int v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ int popc_below(unsigned m) const { return __popc(m & ((1u << (threadIdx.x & 31)) - 1u)); }
    __device__ __forceinline__ int popc(unsigned m) const { return __popc(m); }
    __device__ __forceinline__ void copy_in(double* dst, const double* src, int n) const
    {
        warp_copy_start((unsigned)__cvta_generic_to_shared(dst), src, n);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    }
    // several copies in flight at once: copy_start ... copy_start, then one copy_wait
    __device__ __forceinline__ void copy_start(double* dst, const double* src, int n) const
    {
        warp_copy_start((unsigned)__cvta_generic_to_shared(dst), src, n);
    }
    __device__ __forceinline__ void copy_wait() const
    {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    }
};
// the stage-task kernel's executor: the same warp, bulk copies through the shared routine
struct WarpExS : WarpEx {
    __device__ __forceinline__ void copy_in(double* dst, const double* src, int n) const
    {
        warp_copy_start_ni((unsigned)__cvta_generic_to_shared(dst), src, n);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    }
    __device__ __forceinline__ void copy_start(double* dst, const double* src, int n) const
    {
        warp_copy_start_ni((unsigned)__cvta_generic_to_shared(dst), src, n);
    }
};
#endif
// all-reduce by butterflies: every lane ends with bit-identical results (IEEE add / max are commutative).
// On the device the butterfly is one routine (WBC_SMALL_NI: inlined since round 2, a call before): the solver is instruction-fetch bound
// (DESIGN.md 4.2), and an inlined 5-step double-precision butterfly is 25 instructions at each of ~60 call sites.
#if defined(__CUDACC__)
__device__ WBC_SMALL_NI double warp_sum_ni(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ WBC_SMALL_NI double warp_max_ni(double v)
{
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif
template <int N, class Ex>
WBC_HD void red_sum(const Ex& ex, double* v)
{
#if defined(__CUDA_ARCH__)
    if (Ex::NL == 32) {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = warp_sum_ni(v[k]);
        return;
    }
#endif
    for (int o = Ex::NL / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] += ex.shfl_xor(v[k], o);
    }
}
template <int N, class Ex>
WBC_HD void red_max(const Ex& ex, double* v)
{
#if defined(__CUDA_ARCH__)
    if (Ex::NL == 32) {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = warp_max_ni(v[k]);
        return;
    }
#endif
    for (int o = Ex::NL / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = fmax(v[k], ex.shfl_xor(v[k], o));
    }
}
template <class Ex> WBC_HD double red_sum1(const Ex& ex, double a) { red_sum<1>(ex, &a); return a; }
template <class Ex> WBC_HD double red_max1(const Ex& ex, double a) { red_max<1>(ex, &a); return a; }

WBC_HD int lowest_bit(unsigned m)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

WBC_HD double rsqrt_(double d)
{
#if defined(__CUDA_ARCH__)
    return rsqrt(d);
#else
    return 1.0 / sqrt(d);
#endif
}

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
#if defined(__CUDACC__)
// the double-precision square root (WBC_SMALL_NI: inlined since round 2; as a shared call it saved ~30 instructions per site)
__device__ WBC_SMALL_NI double dsqrt_ni(double a) { return sqrt(a); }
#endif
WBC_HD double sqrt_shared(double a)
{
#if defined(__CUDA_ARCH__)
    return dsqrt_ni(a);
#else
    return sqrt(a);
#endif
}
// returns (d1est + 1) * 4 + (d2est + 1)
WBC_HDNI_PARAB int estimateparabolicmodel(double absasum, double absasum2, double mx, double mb, double md, double d1, double d2)
{ // opt.cpp:23071-23131
    const double eps = 4 * MACHEPS;
    const double sq2 = sqrt_shared(absasum2);
    double e1 = eps * md * (mx * absasum + mb);
    double e2 = eps * md * (mx * sq2 + mb);
    double err = sqrt_shared(e1 * e2);
    const int d1est = (fabs(d1) <= err) ? 0 : (d1 > 0 ? 1 : (d1 < 0 ? -1 : 0));
    e1 = eps * md * md * absasum;
    e2 = eps * md * md * sq2;
    err = sqrt_shared(e1 * e2);
    const int d2est = (fabs(d2) <= err) ? 0 : (d2 > 0 ? 1 : (d2 < 0 ? -1 : 0));
    return (d1est + 1) * 4 + (d2est + 1);
}

// ------------------------------------------------------------------------------------------------
// y = E x (+ b) with E = [H, CI'; CI, rho I] (n = 30 + nic).  Lane i < 30 owns main row i, lane k < nic slack row k.
template <bool SPILL, class Ex>
WBC_HDNI void symv(const Ex ex, const Work w, int nic, double rho, const double* x, const double* b, double* y)
{
    const double* H = W_H(w);
    const double* CI = QS<SPILL>::CI(w);
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 5
        for (int j = 0; j < NMAIN; j += 2) {
            a0 += H[j * LDH + i] * x[j];
            a1 += H[(j + 1) * LDH + i] * x[j + 1];
        }
        int k = 0;
        for (; k + 1 < nic; k += 2) {
            a0 += CI[k * LDH + i] * x[NMAIN + k];
            a1 += CI[(k + 1) * LDH + i] * x[NMAIN + k + 1];
        }
        if (k < nic) a0 += CI[k * LDH + i] * x[NMAIN + k];
        const double acc = a0 + a1;
        y[i] = b ? acc + b[i] : acc;
    }
    for (int k = ex.lane(); k < nic; k += Ex::NL) {
        const double* row = CI + k * LDH;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 5
        for (int i = 0; i < NMAIN; i += 2) {
            a0 += row[i] * x[i];
            a1 += row[i + 1] * x[i + 1];
        }
        const double acc = a0 + a1 + rho * x[NMAIN + k];
        y[NMAIN + k] = b ? acc + b[NMAIN + k] : acc;
    }
    ex.sync();
}

// ------------------------------------------------------------------------------------------------
// Left-looking column Cholesky A = U'U into the packed factor Z (row c of Z = U[0..c-1][c]), diagonal in zd,
// reciprocal diagonal in zrinv.  Column k: every lane c >= k forms a_kc - sum_{m<k} Z[c][m] Z[k][m]; the lane
// with c == k owns the pivot.  `src(k, c)` supplies a_kc (k <= c) -- the factorisation is fused with the
// (masked) read of the matrix, or runs in place.
// SKIP = false: returns false on a non-positive pivot (linalg.cpp:29204-29235).
// SKIP = true (positive semi-definite input): a pivot below pivtol * (original diagonal) marks a dependent
//   row: dep[k] = 1, its row of U is zero, zd = zrinv = 0; *ambiguous is set when such a pivot is not clearly
//   rounding noise (above 1e-3 * pivtol).
template <bool SKIP, class Ex, class Src>
WBC_HD bool chol_cols(const Ex& ex, double* Z, int n, double* zd, double* zrinv, int* dep, double pivtol, bool* ambiguous, const Src& src)
{
    bool amb = false;
#pragma unroll 1
    for (int k = 0; k < n; k++) {
        const double* rk = Z + zoff(k);
        double piv = 0.0, d0 = 0.0;
        // pass over the lanes' rows c = k + lane, k + lane + NL, ...; the pivot lane is the first of the first pass
#pragma unroll 1
        for (int cb = k; cb < n; cb += Ex::NL) {
            const int c = cb + ex.lane();
            double acc = 0.0, a = 0.0;
            if (c < n) {
                a = src(k, c);
                const double* rc = Z + zoff(c);
                double s0 = 0.0, s1 = 0.0;
                int m = 0;
#pragma unroll 1
                for (; m + 1 < k; m += 2) {
                    s0 += rc[m] * rk[m];
                    s1 += rc[m + 1] * rk[m + 1];
                }
                if (m < k) s0 += rc[m] * rk[m];
                acc = a - (s0 + s1);
            }
            if (cb == k) {
                piv = ex.shfl(acc, 0);
                d0 = ex.shfl(a, 0);
            }
            bool skip = false;
            if (SKIP) {
                if (!(piv > pivtol * d0)) {
                    skip = true;
                    if (cb == k && piv > 1.0e-3 * pivtol * d0) {
                        amb = true;
#ifdef WBC_EMU_DEBUG
                        printf("ambiguous pivot k=%d piv/d0=%.3e\n", k, piv / d0);
#endif
                    }
                }
            } else {
                if (!(piv > 0.0)) return false;
            }
            const double rinv = skip ? 0.0 : rsqrt_(piv);
            if (c < n) {
                if (c == k) {
                    zd[k] = skip ? 0.0 : piv * rinv;
                    zrinv[k] = rinv;
                    if (SKIP) dep[k] = skip ? 1 : 0;
                } else {
                    Z[zoff(c) + k] = acc * rinv;
                }
            }
        }
        ex.sync();
    }
    if (ambiguous) *ambiguous = amb;
    return true;
}
struct SrcInPlace {
    const double* Z;
    const double* diag;
    WBC_HD double operator()(int k, int c) const { return (c == k) ? diag[k] : Z[zoff(c) + k]; }
};

// Solve U'U x = rhs in place with the packed factor; x in memory, n <= NL*NR.  Lane l holds x[l], x[l+NL], ... in
// registers, column-oriented both ways.  Dependent pivots (zrinv = 0) yield a zero component.
template <int NR, class Ex>
WBC_HD void tri_solve_regs(const Ex& ex, const double* Z, int n, const double* zrinv, double* x, bool forward, bool backward)
{
    if (Ex::NL == 1) {
        if (forward)
#pragma unroll 1
            for (int k = 0; k < n; k++) {
                const double yk = x[k] * zrinv[k];
                x[k] = yk;
#pragma unroll 1
                for (int i = k + 1; i < n; i++) x[i] -= Z[zoff(i) + k] * yk;
            }
        if (backward)
#pragma unroll 1
            for (int k = n - 1; k >= 0; k--) {
                const double xk = x[k] * zrinv[k];
                x[k] = xk;
                const double* rk = Z + zoff(k);
#pragma unroll 1
                for (int i = 0; i < k; i++) x[i] -= rk[i] * xk;
            }
        return;
    }
    const int l = ex.lane();
    double xr[NR];
    const double* rows[NR];
#pragma unroll
    for (int s = 0; s < NR; s++) {
        const int i = l + Ex::NL * s;
        xr[s] = (i < n) ? x[i] : 0.0;
        rows[s] = Z + zoff(i < n ? i : 0);
    }
    if (forward)
#pragma unroll 1
        for (int k = 0; k < n; k++) {
            double v = xr[0];
#pragma unroll
            for (int s = 1; s < NR; s++) if ((k / Ex::NL) == s) v = xr[s];
            const double yk = ex.shfl(v, k % Ex::NL) * zrinv[k];
#pragma unroll
            for (int s = 0; s < NR; s++) {
                const int i = l + Ex::NL * s;
                if (i == k) xr[s] = yk;
                else if (i > k && i < n) xr[s] -= rows[s][k] * yk;
            }
        }
    if (backward)
#pragma unroll 1
        for (int k = n - 1; k >= 0; k--) {
            double v = xr[0];
#pragma unroll
            for (int s = 1; s < NR; s++) if ((k / Ex::NL) == s) v = xr[s];
            const double xk = ex.shfl(v, k % Ex::NL) * zrinv[k];
            const double* rk = Z + zoff(k);
#pragma unroll
            for (int s = 0; s < NR; s++) {
                const int i = l + Ex::NL * s;
                if (i == k) xr[s] = xk;
                else if (i < k) xr[s] -= rk[i] * xk;
            }
        }
#pragma unroll
    for (int s = 0; s < NR; s++) if (l + Ex::NL * s < n) x[l + Ex::NL * s] = xr[s];
    ex.sync();
}
#if defined(__CUDACC__)
// Device, n <= 32: one component per lane, both sweeps (same operations on the same operands as tri_solve_regs, without
// its slot bookkeeping).  Dependent pivots (zrinv = 0) yield a zero component.
__device__ WBC_NI_TSW void tri_solve_warp32(const double* Z, int n, const double* zrinv, double* x)
{
    const int l = threadIdx.x & 31;
    const double* r0 = Z + zoff(l < n ? l : 0);
    const double zr = (l < n) ? zrinv[l] : 0.0;
    double x0 = (l < n) ? x[l] : 0.0;
#pragma unroll 1
    for (int k = 0; k < n; k++) {
        const double yk = __shfl_sync(0xffffffffu, x0, k) * __shfl_sync(0xffffffffu, zr, k);
        if (l == k) x0 = yk;
        else if (l > k && l < n) x0 -= r0[k] * yk;
    }
#pragma unroll 1
    for (int k = n - 1; k >= 0; k--) {
        const double xk = __shfl_sync(0xffffffffu, x0, k) * __shfl_sync(0xffffffffu, zr, k);
        const double* rk = Z + zoff(k);
        if (l == k) x0 = xk;
        else if (l < k) x0 -= rk[l] * xk;
    }
    if (l < n) x[l] = x0;
    __syncwarp();
}
#endif
template <bool SPILL, class Ex>
WBC_HDNI void tri_solve(const Ex ex, const double* Z, int n, const double* zrinv, double* x, bool forward, bool backward)
{
#if defined(__CUDA_ARCH__)
    if (!SPILL && Ex::NL == 32 && n <= 32 && forward && backward) { tri_solve_warp32(Z, n, zrinv, x); return; }
#endif
    tri_solve_regs<SPILL ? 4 : 2>(ex, Z, n, zrinv, x, forward, backward);
}

// ------------------------------------------------------------------------------------------------
// QQP (opt.cpp:29675-30566) specialised to: dense A (akind 2, upper), unit scales, zero origin,
// variables [0,NMAIN) free, variables [NMAIN,n) bounded below by 0.
struct QqpState {
    int n, nic;
    double rho;
    double absasum, absasum2, mb;
    int nfree, cnmodelage;
    int ncholesky;
};

// sasexploredirection (opt.cpp:27433-27528), box-only: largest feasible step along d and the first bound that
// blocks it.  The reference scans the bounded variables in order and keeps the first strict improvement; here
// every lane evaluates its variables' step (same safeminposrv arithmetic, from the unbounded start) and a
// butterfly picks the minimum with ties to the lowest index.
template <class Ex>
WBC_HD void sas_explore_direction(const Ex& ex, const int* cstatus, const double* xc, int n, const double* d, double& stpmax, int& cidx,
                                  double& cval)
{
    double best = BIGSTEP;
    int bi = 0x7fffffff;
    for (int i = NMAIN + ex.lane(); i < n; i += Ex::NL) {
        const double di = d[i];
        if (di < 0.0 && cstatus[i] <= 0) {
            const double r = safeminposrv(xc[i] - 0.0, -di, BIGSTEP);
            if (r < best) { best = r; bi = i; }
        }
    }
    for (int o = Ex::NL / 2; o > 0; o >>= 1) {
        const double ov = ex.shfl_xor(best, o);
        const int oi = ex.shfl_xori(bi, o);
        if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    stpmax = best;
    cidx = (best < BIGSTEP) ? bi : -1;
    cval = 0.0;
}

// qqpsolver_projectedtargetfunction (opt.cpp:30582-30652) for four steps at once: f_k = exb.t_k + 0.5 t_k.(E t_k),
// t_k = the projection of xc + steps[k] d on the bounds.  One pass over E serves all four.
template <bool SPILL, class Ex>
WBC_HDNI void eval_candidates(const Ex ex, const Work w, int nic, double rho, const double* d, double s0, double s1, double s2, double s3,
                              double* f)
{
    const int n = NMAIN + nic;
    const double* H = W_H(w);
    const double* CI = QS<SPILL>::CI(w);
    const double* xc = QS<SPILL>::V(w, V_XC);
    const double* exb = QS<SPILL>::EXB(w);
    double* t0 = QS<SPILL>::V(w, V_T0);
    double* t1 = QS<SPILL>::V(w, V_T1);
    double* t2 = QS<SPILL>::V(w, V_T2);
    double* t3 = QS<SPILL>::V(w, V_T3);
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        const double xi = xc[i], di = d[i];
        double v0 = (s0 != 0.0) ? xi + s0 * di : xi;
        double v1 = (s1 != 0.0) ? xi + s1 * di : xi;
        double v2 = (s2 != 0.0) ? xi + s2 * di : xi;
        double v3 = (s3 != 0.0) ? xi + s3 * di : xi;
        if (i >= NMAIN) {
            if (v0 < 0.0) v0 = 0.0;
            if (v1 < 0.0) v1 = 0.0;
            if (v2 < 0.0) v2 = 0.0;
            if (v3 < 0.0) v3 = 0.0;
        }
        t0[i] = v0; t1[i] = v1; t2[i] = v2; t3[i] = v3;
    }
    ex.sync();
    double r8[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // lin[4], quad[4]
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 3
        for (int j = 0; j < NMAIN; j++) {
            const double e = H[j * LDH + i];
            a0 += e * t0[j]; a1 += e * t1[j]; a2 += e * t2[j]; a3 += e * t3[j];
        }
        for (int k = 0; k < nic; k++) {
            const double e = CI[k * LDH + i];
            a0 += e * t0[NMAIN + k]; a1 += e * t1[NMAIN + k]; a2 += e * t2[NMAIN + k]; a3 += e * t3[NMAIN + k];
        }
        const double e = exb[i];
        r8[0] += e * t0[i]; r8[1] += e * t1[i]; r8[2] += e * t2[i]; r8[3] += e * t3[i];
        r8[4] += t0[i] * a0; r8[5] += t1[i] * a1; r8[6] += t2[i] * a2; r8[7] += t3[i] * a3;
    }
    for (int k = ex.lane(); k < nic; k += Ex::NL) {
        const double* row = CI + k * LDH;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 3
        for (int i = 0; i < NMAIN; i++) {
            const double e = row[i];
            a0 += e * t0[i]; a1 += e * t1[i]; a2 += e * t2[i]; a3 += e * t3[i];
        }
        const int ii = NMAIN + k;
        a0 += rho * t0[ii]; a1 += rho * t1[ii]; a2 += rho * t2[ii]; a3 += rho * t3[ii];
        const double e = exb[ii];
        r8[0] += e * t0[ii]; r8[1] += e * t1[ii]; r8[2] += e * t2[ii]; r8[3] += e * t3[ii];
        r8[4] += t0[ii] * a0; r8[5] += t1[ii] * a1; r8[6] += t2[ii] * a2; r8[7] += t3[ii] * a3;
    }
    red_sum<8>(ex, r8);
    if (ex.lane() == 0) {
        f[0] = r8[0] + 0.5 * r8[4]; f[1] = r8[1] + 0.5 * r8[5]; f[2] = r8[2] + 0.5 * r8[6]; f[3] = r8[3] + 0.5 * r8[7];
    }
    ex.sync();
}

// qqpsolver_findbeststepandmove (opt.cpp:30882-31003) fused with sasmoveto (27574-27723, box-only).
// The candidate steps {stp, addsteps[k] > stp} are evaluated together (eval_candidates).
template <bool SPILL, class Ex>
WBC_HDNI void qqp_find_best_step_and_move(const Ex ex, const Work w, int nic, double rho, const double* d, double stp, int needact, int cidx,
                                          double cval, double a0, double a1, double a2, int addcnt)
{
    const int n = NMAIN + nic;
    double stpbest = stp;
    if (addcnt > 0) {
        double* f = QS<SPILL>::V(w, V_SPARE);
        eval_candidates<SPILL>(ex, w, nic, rho, d, stp, a0, addcnt > 1 ? a1 : a0, addcnt > 2 ? a2 : a0, f);
        double fbest = f[0];
        if (a0 > stp && f[1] < fbest) { fbest = f[1]; stpbest = a0; }
        if (addcnt > 1 && a1 > stp && f[2] < fbest) { fbest = f[2]; stpbest = a1; }
        if (addcnt > 2 && a2 > stp && f[3] < fbest) { fbest = f[3]; stpbest = a2; }
        ex.sync();
    }
    double* xc = QS<SPILL>::V(w, V_XC);
    int* cstatus = W_CSTATUS(w);
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        const double old = xc[i];
        double v = old + stpbest * d[i];
        if (i >= NMAIN && v < 0.0) v = 0.0;
        if (needact && i == cidx) { v = cval; cstatus[i] = 1; }
        if (i >= NMAIN && v <= 0.0 && v != old) { v = 0.0; cstatus[i] = 1; }
        xc[i] = v;
    }
    ex.sync();
}

// qqpsolver_quadraticmodel (opt.cpp:30753-30821): d1 = d.g, d2 = 0.5 d.(E d), with the noise-aware sign estimates.
// Results in iscr[4..5] (d1est, d2est) and V_SPARE[4..5] (d1, d2).
template <bool SPILL, class Ex>
WBC_HDNI void qqp_quadratic_model(const Ex ex, const Work w, int nic, double rho, double absasum, double absasum2, double mb)
{
    const int n = NMAIN + nic;
    const double* dc = QS<SPILL>::V(w, V_DC);
    const double* gc = QS<SPILL>::V(w, V_GC);
    const double* xc = QS<SPILL>::V(w, V_XC);
    double* t0 = QS<SPILL>::V(w, V_T0);
    symv<SPILL>(ex, w, nic, rho, dc, (const double*)nullptr, t0);
    double ss[2] = {0.0, 0.0}, mm[2] = {0.0, 0.0};
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        const double di = dc[i];
        ss[0] += di * t0[i];
        ss[1] += di * gc[i];
        mm[0] = fmax(mm[0], fabs(xc[i]));
        mm[1] = fmax(mm[1], fabs(di));
    }
    red_sum<2>(ex, ss);
    red_max<2>(ex, mm);
    const double d2 = 0.5 * ss[0], d1 = ss[1];
    const int code = estimateparabolicmodel(absasum, absasum2, mm[0], mb, mm[1], d1, d2);
    const int d1est = (code >> 2) - 1, d2est = (code & 3) - 1;
    ex.sync();
    if (ex.lane() == 0) {
        double* o = QS<SPILL>::V(w, V_SPARE);
        o[4] = d1; o[5] = d2;
        W_ISCR(w)[4] = d1est; W_ISCR(w)[5] = d2est;
    }
    ex.sync();
}

// qqpsolver_cnewtonbuild (opt.cpp:31058-31201).  The factor is produced directly in the "scattered"
// n x n layout ALGLIB ends with (identity rows for fixed variables): factoring the masked matrix
// gives identical entries because the extra terms are exact zeros.
template <bool SPILL>
struct SrcMaskedE {
    const double* H;
    const double* CI;
    const double* reg;
    const int* isfree;
    double rho;
    WBC_HD double operator()(int k, int c) const
    {
        if (c == k) return reg[k];                                   // E_kk + regulariser, or 1 for a fixed variable
        if (c < NMAIN) return H[k * LDH + c];
        if (k < NMAIN) return isfree[c] ? CI[(c - NMAIN) * LDH + k] : 0.0;
        return 0.0;
    }
};
template <bool SPILL, class Ex>
WBC_HDNI bool qqp_cnewton_build(const Ex ex, const Work w, int nic, double rho, int* nfree_out)
{
    const int n = NMAIN + nic;
    const double* H = W_H(w);
    const double* CI = QS<SPILL>::CI(w);
    const double* xc = QS<SPILL>::V(w, V_XC);
    double* reg = QS<SPILL>::V(w, V_REG);
    int* isfree = W_ISFREE(w);
    double nf = 0.0;
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        const int fr = !(i >= NMAIN && xc[i] == 0.0);
        isfree[i] = fr;
        nf += fr;
    }
    nf = red_sum1(ex, nf);
    ex.sync();
    *nfree_out = (int)nf;
    if ((int)nf == 0) return false;
    // diagonal with the regulariser 1e-9 * sum_j |A_ff[i][j]| over free j (31150-31167); main variables are always free
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) {
        double v = 0.0;
        for (int j = 0; j < NMAIN; j++) v += fabs(H[j * LDH + i]);
        for (int k = 0; k < nic; k++)
            if (isfree[NMAIN + k]) v += fabs(CI[k * LDH + i]);
        if (v == 0.0) v = 1.0;
        reg[i] = H[i * LDH + i] + 1.0e-9 * v;
    }
    for (int k = ex.lane(); k < nic; k += Ex::NL) {
        if (isfree[NMAIN + k]) {
            const double* row = CI + k * LDH;
            double v = 0.0;
            for (int i = 0; i < NMAIN; i++) v += fabs(row[i]);
            v += fabs(rho);
            if (v == 0.0) v = 1.0;
            reg[NMAIN + k] = rho + 1.0e-9 * v;
        } else reg[NMAIN + k] = 1.0;
    }
    ex.sync();
    SrcMaskedE<SPILL> src;
    src.H = H; src.CI = CI; src.reg = reg; src.isfree = isfree; src.rho = rho;
    return chol_cols<false>(ex, QS<SPILL>::Z(w), n, QS<SPILL>::V(w, V_ZD), QS<SPILL>::V(w, V_ZRINV), (int*)nullptr, 0.0, (bool*)nullptr, src);
}

// qqpsolver_cnewtonupdate (opt.cpp:31314-31426) + spdmatrixcholeskyupdatefixbuf (linalg.cpp:27657-27819, upper):
// fixing variable k = removing row/column k of the factor with a sweep of Givens rotations.  Serial in the row
// index; the rotated row is held in registers.
template <int NR, class Ex>
WBC_HD void givens_fix_regs(const Ex& ex, double* Z, double* zd, double* zrinv, int n, int k)
{
    if (Ex::NL == 1) {
        double buf[MAXNT];
#pragma unroll 1
        for (int j = k + 1; j < n; j++) { buf[j] = Z[zoff(j) + k]; Z[zoff(j) + k] = 0.0; }
#pragma unroll 1
        for (int i = 0; i < k; i++) Z[zoff(k) + i] = 0.0;
        zd[k] = 1.0; zrinv[k] = 1.0;
#pragma unroll 1
        for (int i = k + 1; i < n; i++) {
            const double bi = buf[i];
            if (bi != 0.0) {
                double cs, sn, r;
                generaterotation(zd[i], bi, cs, sn, r);
                zd[i] = r; zrinv[i] = 1.0 / r; buf[i] = 0.0;
#pragma unroll 1
                for (int j = i + 1; j < n; j++) {
                    const double v = Z[zoff(j) + i], vv = buf[j];
                    Z[zoff(j) + i] = cs * v + sn * vv;
                    buf[j] = -sn * v + cs * vv;
                }
            }
        }
        return;
    }
    const int l = ex.lane();
    double br[NR];
    double* rows[NR];
#pragma unroll
    for (int s = 0; s < NR; s++) {
        const int j = l + Ex::NL * s;
        rows[s] = Z + zoff(j < n ? j : 0);
        br[s] = 0.0;
        if (j > k && j < n) { br[s] = rows[s][k]; rows[s][k] = 0.0; }
        if (j < k) Z[zoff(k) + j] = 0.0;
    }
    if (l == 0) { zd[k] = 1.0; zrinv[k] = 1.0; }
    ex.sync();
#pragma unroll 1
    for (int i = k + 1; i < n; i++) {
        double v = br[0];
#pragma unroll
        for (int s = 1; s < NR; s++) if ((i / Ex::NL) == s) v = br[s];
        const double bi = ex.shfl(v, i % Ex::NL);
        if (bi != 0.0) {
            double cs, sn, r;
            generaterotation(zd[i], bi, cs, sn, r);
            ex.sync();      // every lane has read zd[i]
#pragma unroll
            for (int s = 0; s < NR; s++) {
                const int j = l + Ex::NL * s;
                if (j == i) { zd[i] = r; zrinv[i] = 1.0 / r; br[s] = 0.0; }
                else if (j > i && j < n) {
                    const double zv = rows[s][i], vv = br[s];
                    rows[s][i] = cs * zv + sn * vv;
                    br[s] = -sn * zv + cs * vv;
                }
            }
        }
    }
    ex.sync();
}
template <bool SPILL, class Ex>
WBC_HDNI bool qqp_cnewton_update(const Ex ex, const Work w, int nic, int cnmaxupdates, int* nfree, int* cnmodelage)
{
    const int n = NMAIN + nic;
    const double* xc = QS<SPILL>::V(w, V_XC);
    int* isfree = W_ISFREE(w);
    double ntf = 0.0;
    for (int i = ex.lane(); i < n; i += Ex::NL)
        if (isfree[i] && i >= NMAIN && xc[i] == 0.0) ntf += 1.0;
    const int ntofix = (int)red_sum1(ex, ntf);
    if (ntofix == 0 || ntofix == *nfree) return false;
    if (*cnmodelage + ntofix > cnmaxupdates) return false;
    for (int k = NMAIN; k < n; k++) {
        if (!(isfree[k] && xc[k] == 0.0)) continue;
        givens_fix_regs<SPILL ? 4 : 2>(ex, QS<SPILL>::Z(w), QS<SPILL>::V(w, V_ZD), QS<SPILL>::V(w, V_ZRINV), n, k);
        if (ex.lane() == 0) isfree[k] = 0;
        ex.sync();
    }
    *nfree -= ntofix;
    *cnmodelage += ntofix;
    return true;
}

// One QQP solve from the point exxc (in/out) on the model (H, CI, rho, exb).  Returns the QQP termination type.
template <bool SPILL, class Ex>
WBC_HDNI int qqp_optimize(const Ex ex, const Work w, int nic, double rho, double epsx, int maxouterits, int* ncholesky, double* flops_io)
{
    const int n = NMAIN + nic;
    double flops = 0.0;
    int nchol = 0, nfree = 0, cnmodelage = 0;
    double* xc = QS<SPILL>::V(w, V_XC);
    double* xp = QS<SPILL>::V(w, V_XP);
    double* gc = QS<SPILL>::V(w, V_GC);
    double* cgc = QS<SPILL>::V(w, V_CGC);
    double* cgp = QS<SPILL>::V(w, V_CGP);
    double* dc = QS<SPILL>::V(w, V_DC);
    double* dp = QS<SPILL>::V(w, V_DP);
    double* spare = QS<SPILL>::V(w, V_SPARE);
    double* exb = QS<SPILL>::EXB(w);
    double* exxc = W_EXXC(w);
    int* cstatus = W_CSTATUS(w);
    int* isfree = W_ISFREE(w);
    int* iscr = W_ISCR(w);
    // settings: qqploaddefaults (opt.cpp:29533-29547) + overrides (41318-41323)
    const int cgminits = 5;
    int cgmaxits = (int)(1 + 0.33 * n + 0.5);          // ae_round, positive argument
    if (cgmaxits < cgminits) cgmaxits = cgminits;
    const int cnmaxupdates = (int)(1 + 0.1 * n + 0.5);

    // |A| statistics with ALGLIB's k = (i==v ? 1 : 2) quirk (opt.cpp:29893-29915) over the upper triangle of E;
    // max|b|; start point clipped to the bounds (29979-29998) and sasstartoptimization (27377-27399)
    double absasum, absasum2, mb;
    {
        const double* H = W_H(w);
        const double* CI = QS<SPILL>::CI(w);
        double ss[2] = {0.0, 0.0};
        double m1 = 0.0;
        for (int i = ex.lane(); i < NMAIN; i += Ex::NL) {
            for (int j = i; j < NMAIN; j++) {
                const double v = H[i * LDH + j], vv = fabs(v);
                const double k = ((double)i == v) ? 1.0 : 2.0;
                ss[0] += vv * k; ss[1] += vv * vv * k;
            }
            for (int kk = 0; kk < nic; kk++) {
                const double v = CI[kk * LDH + i], vv = fabs(v);
                const double k = ((double)i == v) ? 1.0 : 2.0;
                ss[0] += vv * k; ss[1] += vv * vv * k;
            }
        }
        for (int kk = ex.lane(); kk < nic; kk += Ex::NL) {
            const double v = rho, vv = fabs(v);
            const double k = ((double)(NMAIN + kk) == v) ? 1.0 : 2.0;
            ss[0] += vv * k; ss[1] += vv * vv * k;
        }
        for (int i = ex.lane(); i < n; i += Ex::NL) {
            m1 = fmax(m1, fabs(exb[i]));
            double v = exxc[i];
            int cs = -1;
            if (i >= NMAIN && v <= 0.0) { v = 0.0; cs = 0; }
            xc[i] = v;
            cstatus[i] = cs;
        }
        red_sum<2>(ex, ss);
        absasum = ss[0]; absasum2 = ss[1];
        mb = red_max1(ex, m1);
        ex.sync();
    }
    int term = 0;
    // NOTE: ALGLIB's single-Cholesky fast path for unconstrained problems (opt.cpp:30033-30073) is gated
    // on akind==0 (CQM storage); DENSE-AUL calls QQP with akind==2 (opt.cpp:41324), so the generic
    // CG + constrained-Newton iteration below runs even when there are no slack variables yet.
    int cgmax = cgminits;
    int outerits = 0;
    for (;;) {
        if (maxouterits > 0 && outerits >= maxouterits) { term = 5; break; }
        if (outerits > 0) {
            // epsx stopping test (30137-30149); epsf = 0 so the function test is skipped
            double v = 0.0;
            for (int i = ex.lane(); i < n; i += Ex::NL) { const double t = xp[i] - xc[i]; v += t * t; }
            v = red_sum1(ex, v);
            if (sqrt(v) <= epsx) { term = 2; break; }
        }
        outerits++;
        for (int i = ex.lane(); i < n; i += Ex::NL) { xp[i] = xc[i]; cgp[i] = 0.0; dp[i] = 0.0; }
        ex.sync();
        for (int cgcnt = 0; cgcnt <= cgmax - 1; cgcnt++) {
            symv<SPILL>(ex, w, nic, rho, xc, exb, gc);                    // targetgradient
            flops += 2.0 * n * n;
            // sasreactivateconstraints, box-only (28992-29047); constrained gradient; CG coefficients (30199-30221).
            // (sasconstraineddirection's "everything active" clause, 28952-28959, cannot fire: the 30 main
            //  variables are never bounded)
            double r3[3] = {0.0, 0.0, 0.0};
            for (int i = ex.lane(); i < n; i += Ex::NL) {
                const double xi = xc[i], g = gc[i];
                const bool atb = (i >= NMAIN && xi == 0.0);
                const bool act = atb && g >= 0.0;
                cstatus[i] = act ? 1 : -1;
                const double cg = act ? 0.0 : g;
                cgc[i] = cg;
                r3[0] += cg * cg;
                const double pv = cgp[i];
                r3[1] += pv * pv;
                if (atb && dp[i] != 0.0) r3[2] += 1.0;
            }
            red_sum<3>(ex, r3);
            const double v = r3[0], vv = r3[1];
            if (sqrt(v) <= 0.0) { term = 4; break; }                      // epsg = 0
            const bool brst = (r3[2] != 0.0) || (vv == 0.0) || (cgcnt % 50 == 0);
            const double beta = brst ? 0.0 : v / vv;
            for (int i = ex.lane(); i < n; i += Ex::NL) {
                double d = -cgc[i] + beta * dp[i];
                if (cstatus[i] > 0) d = 0.0;
                dc[i] = d;
            }
            ex.sync();
            double stpmax, cval; int cidx;
            sas_explore_direction(ex, cstatus, xc, n, dc, stpmax, cidx, cval);
            qqp_quadratic_model<SPILL>(ex, w, nic, rho, absasum, absasum2, mb);
            const double d1 = spare[4], d2 = spare[5];
            const int d1est = iscr[4], d2est = iscr[5];
            flops += 2.0 * n * n;
            if (d1 == 0.0 && d2 == 0.0) { term = 4; break; }
            if (d1est >= 0) { term = 7; break; }
            if (d2est <= 0 && cidx < 0) { term = -4; break; }
            double stp, a0 = 0.0, a1 = 0.0, a2 = 0.0; int needact, stpcnt;
            if (d2est > 0) {
                const double fullstp = -d1 / (2 * d2);
                needact = fullstp >= stpmax;
                if (needact) { stp = stpmax; a0 = stpmax * 4; a1 = fullstp; a2 = fullstp / 4; stpcnt = 3; }
                else { stp = fullstp; stpcnt = 0; }
            } else {
                stp = stpmax; needact = 1; a0 = 4 * stpmax; stpcnt = 1;
            }
            qqp_find_best_step_and_move<SPILL>(ex, w, nic, rho, dc, stp, needact, cidx, cval, a0, a1, a2, stpcnt);
            if (stpcnt > 0) flops += (1 + stpcnt) * 2.0 * n * n;
            for (int i = ex.lane(); i < n; i += Ex::NL) { dp[i] = dc[i]; cgp[i] = cgc[i]; }
            ex.sync();
        }
        if (term != 0) break;
        cgmax = cgmaxits;
        // constrained Newton phase (30353-30527)
        int newtcnt = 0;
        for (;;) {
            bool b;
            if (newtcnt == 0) {
                b = qqp_cnewton_build<SPILL>(ex, w, nic, rho, &nfree);
                cnmodelage = 0;
                nchol++;
                flops += (double)n * n * n / 3.0;
                if (b) cgmax = cgminits;
            } else {
                b = qqp_cnewton_update<SPILL>(ex, w, nic, cnmaxupdates, &nfree, &cnmodelage);
                flops += 3.0 * n * n;
            }
            if (!b) break;
            newtcnt++;
            symv<SPILL>(ex, w, nic, rho, xc, exb, gc);
            // qqpsolver_cnewtonstep (31474-31536), epsg = 0
            double gg = 0.0;
            for (int i = ex.lane(); i < n; i += Ex::NL) {
                const double g = isfree[i] ? gc[i] : 0.0;
                gg += g * g;
                dc[i] = -g;
            }
            gg = red_sum1(ex, gg);
            ex.sync();
            if (sqrt(gg) <= 0.0) break;
            tri_solve<SPILL>(ex, QS<SPILL>::Z(w), n, QS<SPILL>::V(w, V_ZRINV), dc, true, true);
            qqp_quadratic_model<SPILL>(ex, w, nic, rho, absasum, absasum2, mb);
            const double d1 = spare[4], d2 = spare[5];
            const int d1est = iscr[4], d2est = iscr[5];
            flops += 6.0 * n * n;
            if (d1est >= 0) break;
            double stpmax, cval; int cidx;
            sas_explore_direction(ex, cstatus, xc, n, dc, stpmax, cidx, cval);
            if (d2est > 0) {
                const double fullstp = -d1 / (2 * d2);
                const int needact = fullstp >= stpmax;
                double stp, a0 = 0.0, a1 = 0.0, a2 = 0.0; int stpcnt;
                if (needact) { stp = stpmax; a0 = stpmax * 4; a1 = fullstp; a2 = fullstp / 4; stpcnt = 3; }
                else { stp = fullstp; stpcnt = 0; }
                ex.sync();
                qqp_find_best_step_and_move<SPILL>(ex, w, nic, rho, dc, stp, needact, cidx, cval, a0, a1, a2, stpcnt);
                if (stpcnt > 0) flops += (1 + stpcnt) * 2.0 * n * n;
            } else {
                if (cidx < 0) { term = -4; break; }
                if (stpmax == 0.0) { cgmax = cgmaxits; break; }
                // f(x) vs f(x + stpmax d) (30493-30503)
                ex.sync();
                eval_candidates<SPILL>(ex, w, nic, rho, dc, 0.0, stpmax, stpmax, stpmax, spare);
                const double f0 = spare[0], f1 = spare[1];
                ex.sync();
                if (f1 >= f0) { cgmax = cgmaxits; break; }
                qqp_find_best_step_and_move<SPILL>(ex, w, nic, rho, dc, stpmax, 1, cidx, cval, stpmax * 4, 1.00, 0.25, 3);
                flops += 12.0 * n * n;
            }
        }
        if (term != 0) break;
    }
    // unpack (30546-30565): unit scale, zero origin; slacks clipped / snapped to the bound
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v = xc[i];
        if (i >= NMAIN && (v < 0.0 || v == 0.0)) v = 0.0;
        exxc[i] = v;
    }
    ex.sync();
    *ncholesky += nchol;
    *flops_io += flops;
    return term;
}

// ------------------------------------------------------------------------------------------------
// Multiplier update, literal form (opt.cpp:41803-42032): Householder QR of the stacked system
// [K | r ; lambda*mxdiag*I | 0], K = KKT matrix of the equality-constrained model with the columns of
// exactly-active slacks replaced, then back-substitution.  Only the multiplier part of the solution
// is needed, so back-substitution stops at row ntotal.  Structural zeros of the regulariser block
// are skipped (reflector j only touches K rows j.. and regulariser rows 0..j).  Works out of the
// warp's global scratch; it is the fallback of the reduced form below.
WBC_HD int kkt_doubles(int nq) { return 2 * nq * (nq + 1); }

template <class Ex>
WBC_HDNI void update_lagrange_multipliers_literal(const Ex ex, const Work w, int nec, int nic, double* flops_io)
{
    const int ntotal = NMAIN + nic, ktotal = nec + nic, nq = ntotal + ktotal, ld = nq + 1;
    double* M = w.kkt;
    double* sv0 = w.g + gl::OFF_SV0;
    double* v = w.g + gl::OFF_QRV;
    const double* A = W_A(w);
    const double* C = W_C(w);
    const double* b = W_B(w);
    const double* exxc = W_EXXC(w);
    double* nulcest = W_NULCEST(w);
    double flops = 0.0;
    // reference point (X0, L0) (41888-41895)
#pragma unroll 1
    for (int i = ex.lane(); i < nq; i += Ex::NL) sv0[i] = (i < ntotal) ? exxc[i] : nulcest[i - ntotal];
#pragma unroll 1
    for (int i = ex.lane(); i < 2 * nq * ld; i += Ex::NL) M[i] = 0.0;
    ex.sync();
    double mxdiag = 0.0;
#pragma unroll 1
    for (int i = 0; i < NMAIN; i++) mxdiag = fmax(mxdiag, fabs(A[i * LDH + i]));
    if (mxdiag == 0.0) mxdiag = 1.0;
    const double lambdareg = 1.0e-8;
    // quadratic term and -b (41919-41927)
#pragma unroll 1
    for (int i = 0; i < NMAIN; i++)
#pragma unroll 1
        for (int j = ex.lane(); j <= NMAIN; j += Ex::NL)
            M[i * ld + (j < NMAIN ? j : nq)] = (j < NMAIN) ? A[i * LDH + j] : -b[i];
    // constraints (41933-41946)
#pragma unroll 1
    for (int i = 0; i < ktotal; i++) {
#pragma unroll 1
        for (int j = ex.lane(); j < NMAIN; j += Ex::NL) {
            const double c = -C[i * LDH + j];
            M[(ntotal + i) * ld + j] = c;
            M[j * ld + ntotal + i] = c;
        }
        if (ex.lane() == 0) {
            if (i >= nec) {
                M[(ntotal + i) * ld + NMAIN + (i - nec)] = -1.0;
                M[(NMAIN + (i - nec)) * ld + ntotal + i] = -1.0;
            }
            M[(ntotal + i) * ld + nq] = -C[i * LDH + NMAIN];
        }
    }
    // regulariser rows (41952-41959)
#pragma unroll 1
    for (int i = ex.lane(); i < nq; i += Ex::NL) M[(nq + i) * ld + i] = lambdareg * mxdiag;
    ex.sync();
    // subtract reference point: rhs_i -= K[i,:] . sv0  (41964-41968), first nq rows only
#pragma unroll 1
    for (int i = ex.lane(); i < nq; i += Ex::NL) {
        double s = 0.0;
#pragma unroll 1
        for (int j = 0; j < nq; j++) s += M[i * ld + j] * sv0[j];
        M[i * ld + nq] -= s;
    }
    ex.sync();
    // active simple constraints: slack exactly zero (41973-41993)
#pragma unroll 1
    for (int i = NMAIN; i < ntotal; i++) {
        if (exxc[i] == 0.0) {
#pragma unroll 1
            for (int j = ex.lane(); j < 2 * nq; j += Ex::NL) M[j * ld + i] = (j == i) ? -1.0 : 0.0;
        }
    }
    ex.sync();
    flops += 2.0 * nq * nq;
    // Householder QR, M: 2nq x (nq+1) row-major.  On exit the upper triangle holds R and column nq holds Q'r.
#pragma unroll 1
    for (int j = 0; j < nq; j++) {
        // rows involved: K rows j..nq-1 and regulariser rows nq..nq+j  -> contiguous range j..nq+j
        const int r0 = j, r1 = nq + j;   // inclusive
        const int len = r1 - r0 + 1;
        // generatereflection (linalg.cpp:19116-19213) on x = M[r0..r1][j]
        const double alpha = M[r0 * ld + j];
        double mx = 0.0;
#pragma unroll 1
        for (int r = r0 + ex.lane(); r <= r1; r += Ex::NL) { const double t = M[r * ld + j]; v[r - r0] = t; mx = fmax(mx, fabs(t)); }
        mx = red_max1(ex, mx);
        ex.sync();
        double xnorm = 0.0;
        if (mx != 0.0) {
            double s = 0.0;
#pragma unroll 1
            for (int r = 1 + ex.lane(); r < len; r += Ex::NL) { const double t = v[r] / mx; s += t * t; }
            s = red_sum1(ex, s);
            xnorm = sqrt(s) * mx;
        }
        double tau = 0.0, beta = alpha;
        if (xnorm != 0.0) {
            const double m2 = fmax(fabs(alpha), fabs(xnorm));
            const double a = alpha / m2, bb = xnorm / m2;
            beta = -m2 * sqrt(a * a + bb * bb);
            if (alpha < 0.0) beta = -beta;
            tau = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
#pragma unroll 1
            for (int r = 1 + ex.lane(); r < len; r += Ex::NL) v[r] *= sc;
            if (ex.lane() == 0) v[0] = 1.0;
        }
        ex.sync();
        if (tau != 0.0) {
            // apply H = I - tau v v' to columns j+1..nq
#pragma unroll 1
            for (int c = j + 1 + ex.lane(); c <= nq; c += Ex::NL) {
                double s = 0.0;
#pragma unroll 8
                for (int r = 0; r < len; r++) s += v[r] * M[(r0 + r) * ld + c];
                s *= tau;
#pragma unroll 8
                for (int r = 0; r < len; r++) M[(r0 + r) * ld + c] -= s * v[r];
            }
            flops += 4.0 * len * (nq - j);
        }
        if (ex.lane() == 0) M[r0 * ld + j] = beta;
        ex.sync();
    }
    // back-substitution for the last ktotal unknowns (42013-42021)
#pragma unroll 1
    for (int i = nq - 1; i >= nq - ktotal; i--) {
        double s = 0.0;
#pragma unroll 1
        for (int jj = i + 1 + ex.lane(); jj < nq; jj += Ex::NL) s += M[i * ld + jj] * sv0[jj];
        s = red_sum1(ex, s);
        const double xi = (M[i * ld + nq] - s) / M[i * ld + i];
        ex.sync();
        if (ex.lane() == 0) sv0[i] = xi;
        ex.sync();
    }
    // sv0 is overwritten in its tail by the solution; nulcest still holds L0
#pragma unroll 1
    for (int i = ex.lane(); i < ktotal; i += Ex::NL) nulcest[i] = nulcest[i] + sv0[ntotal + i];
    ex.sync();
    *flops_io += flops;
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
// the routine returns false and the caller runs the literal form.  Workspace: the (idle) QQP arrays.
WBC_HD void tri_index_lower(int e, int& c, int& r)
{
    // e in [0, m(m+1)/2) -> (c >= r) of a lower triangle stored row by row
    int cc = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while ((cc + 1) * (cc + 2) / 2 <= e) cc++;
    while (cc * (cc + 1) / 2 > e) cc--;
    c = cc;
    r = e - cc * (cc + 1) / 2;
}

template <class Ex>
WBC_HDNI_M bool update_lagrange_multipliers_reduced(const Ex ex, const Work w, int nec, int nic, double pivtol, int* flags_io, double* flops_io)
{
    const int ktotal = nec + nic;
    // Workspace (the QQP is not running, so everything but EXXC and the ints is free):
    //   big (sl::BIG doubles, in front of CI):  W [KACAP+1][31] | factor of A and its reciprocal diagonal  -- first pass on an active set
    //                                           S, its factor, the factor of G (packed triangles)          -- afterwards
    //   the CI array:                           eight vectors of 40, the active list and the dependency flags
    double* big = W_BIG(w);
    double* civ = SM_(w, sl::OFF_CI);
    constexpr int VS = 40;
    double* sd = civ;
    double* srinv = civ + VS;
    double* rho_ = civ + 2 * VS;
    double* dl = civ + 3 * VS;
    double* u1 = civ + 4 * VS;
    double* gd = civ + 5 * VS;
    double* grinv = civ + 6 * VS;
    double* nu0 = civ + 7 * VS;
    int* act = reinterpret_cast<int*>(civ + 8 * VS);
    int* dep = act + VS;
    static_assert(KACAP < VS && 9 * VS <= NICCAP * LDH, "vectors of the multiplier update fit the CI array");
    const double* exxc = W_EXXC(w);
    double* nulcest = W_NULCEST(w);
    const double* C = W_C(w);
    double flops = 0.0;
    // ---- active list, in row order (ballot compaction)
    int ka = 0;
#pragma unroll 1
    for (int base = 0; base < ktotal; base += Ex::NL) {
        const int r = base + ex.lane();
        const bool on = (r < ktotal) && ((r < nec) || (exxc[NMAIN + (r - nec)] == 0.0));
        const unsigned m = ex.ballot(on);
        const int pos = ka + ex.popc_below(m);
        if (on && pos <= KACAP) act[pos] = r;
        ka += ex.popc(m);
    }
    ex.sync();
    if (ka > KACAP) {
#ifdef WBC_EMU_DEBUG
        printf("fallback ka=%d > KACAP\n", ka);
#endif
        return false;
    }
    constexpr int WROWS = (KACAP + 1) * LDH + 1;   // 1148
    double* Wm = big;                              // [KACAP+1][31]: U^-T c_m | d_m ; last row: t
    double* LAs = big + WROWS;                     // packed factor of A (zoff(30) = 435) | its reciprocal diagonal [32]
    const double* larinv = LAs + gl::LA_DOUBLES;
    static_assert(WROWS + gl::LA_DOUBLES + 32 <= sl::BIG, "reduced multiplier update workspace (first pass)");
    constexpr int TRI = (KACAP * (KACAP - 1)) >> 1;                      // packed triangle of the largest active set, 630
    static_assert(2 * TRI <= sl::BIG && TRI <= 648, "reduced multiplier update workspace (S and G)");
    double* Sm = big;                              // S = W W' (fetched back from the cache once W is dead), factored in place
    double* G = big + TRI;                         // factor of G = Lt'Lt (rank-deficient active sets)
    double* Sfac = Sm;                             // where the factor of S / of G is read from (a cache hit puts them elsewhere)
    double* Gfac = G;
    // ---- the cache: W, S = W W', its factor and (rank-deficient sets) the factor of G depend only on which rows of C
    // are active, not on the multipliers or the point.  The outer iterations of a solve usually end on the same active
    // set (the working set settles in the first one), so the second and third update reuse them bit for bit.
    // Key: the active row list and the working-set version (update_working_set counts every row move).
    double* mc = w.g + gl::OFF_MC;
    int* mci = reinterpret_cast<int*>(mc + gl::MC_INT);
    const int version = W_ISCR(w)[6];
    const int npairs = ka * (ka + 1) / 2;
    const int nz = zoff(ka);
    bool hit = (mci[80] == ka) && (mci[81] == version);
    if (hit) {
        double diff = 0.0;
#pragma unroll 1
        for (int m = ex.lane(); m < ka; m += Ex::NL)
            if (mci[m] != act[m]) diff = 1.0;
        hit = red_sum1(ex, diff) == 0.0;
    }
#ifdef WBC_NO_REUSE
    hit = false;                                   // test builds: always recompute (tests/test_emulation.py compares the two)
#endif
#pragma unroll 1
    for (int m = ex.lane(); m < ka; m += Ex::NL) nu0[m] = nulcest[act[m]];
    int ndep_i = 0;
    bool g_later = false;
    if (!hit) {
        if (ex.lane() == 0) mci[80] = -1;          // invalid until this pass completes
        ex.copy_in(LAs, W_LA(w), gl::LA_DOUBLES + 32);
        ex.sync();
        // ---- forward substitutions U' y = c_m, one lane per right-hand side
#pragma unroll 1
        for (int m = ex.lane(); m <= ka; m += Ex::NL) {
            double* y = &Wm[m * LDH];
            const double* src = (m < ka) ? &C[act[m] * LDH] : W_B(w);
#pragma unroll 1
            for (int i = 0; i < NMAIN; i++) {
                const double* li = LAs + zoff(i);
                double s0 = src[i], s1 = 0.0;
                int k = 0;
#pragma unroll 1
                for (; k + 1 < i; k += 2) { s0 -= li[k] * y[k]; s1 -= li[k + 1] * y[k + 1]; }
                if (k < i) s0 -= li[k] * y[k];
                y[i] = (s0 + s1) * larinv[i];
            }
            y[NMAIN] = (m < ka) ? src[NMAIN] : 0.0;
        }
        ex.sync();
        // ---- Schur complement S = W W' (packed lower + diagonal vector).  W fills the workspace, so S goes straight to the
        // cache (where the later updates on this active set look for it) and comes back once W is dead.
#pragma unroll 1
        for (int e = ex.lane(); e < npairs; e += Ex::NL) {
            int c, r;
            tri_index_lower(e, c, r);
            const double* wr = &Wm[r * LDH];
            const double* wc = &Wm[c * LDH];
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 1
            for (int k = 0; k < NMAIN; k += 2) { s0 += wr[k] * wc[k]; s1 += wr[k + 1] * wc[k + 1]; }
            if (r == c) sd[r] = s0 + s1; else mc[gl::MC_S0 + zoff(c) + r] = s0 + s1;
        }
        ex.sync();
        // c0 = d + W t, for this and the later updates on this active set
#pragma unroll 1
        for (int m = ex.lane(); m < ka; m += Ex::NL) {
            const double* wm = &Wm[m * LDH];
            const double* wt = &Wm[ka * LDH];
            double sacc = wm[NMAIN];
#pragma unroll 1
            for (int k = 0; k < NMAIN; k++) sacc += wm[k] * wt[k];
            mc[gl::MC_C0 + m] = sacc;
            mc[gl::MC_S0D + m] = sd[m];
        }
        ex.sync();                                 // W is dead; every lane's part of S has been written
        ex.copy_in(Sm, mc + gl::MC_S0, nz);
        flops += (double)(ka + 1) * NMAIN * NMAIN + (double)ka * ka * NMAIN + 2.0 * ka * NMAIN;
    } else {
        // the cached blocks are fetched in one round trip: S, the factor of S and (when the three fit) the factor of G
        ndep_i = mci[82];
        Sfac = big + nz;
        Gfac = big + 2 * nz;
        g_later = (ndep_i != 0) && (3 * nz > sl::BIG);
        ex.copy_start(Sm, mc + gl::MC_S0, nz);
        ex.copy_start(Sfac, mc + gl::MC_SF, nz);
        if (ndep_i != 0 && !g_later) ex.copy_start(Gfac, mc + gl::MC_G, nz);
        ex.copy_wait();
#pragma unroll 1
        for (int m = ex.lane(); m < ka; m += Ex::NL) sd[m] = mc[gl::MC_S0D + m];
    }
    ex.sync();
    // rho = d + W t - S nu0
#pragma unroll 1
    for (int m = ex.lane(); m < ka; m += Ex::NL) {
        const double sacc = mc[gl::MC_C0 + m];
        double sn = sd[m] * nu0[m];
#pragma unroll 1
        for (int k = 0; k < ka; k++)
            if (k != m) sn += ((k < m) ? Sm[zoff(m) + k] : Sm[zoff(k) + m]) * nu0[k];
        rho_[m] = sacc - sn;
        dl[m] = sacc - sn;
    }
    ex.sync();
    flops += 2.0 * ka * ka;
    if (!hit) {
        bool ambiguous = false;
        {
            SrcInPlace src;
            src.Z = Sm; src.diag = sd;
            chol_cols<true>(ex, Sm, ka, sd, srinv, dep, pivtol, &ambiguous, src);
        }
        if (ambiguous) {
#ifdef WBC_EMU_DEBUG
            printf("fallback ambiguous ka=%d\n", ka);
#endif
            return false;
        }
        flops += (double)ka * ka * ka / 3.0 + 2.0 * ka * ka;
        double ndep = 0.0;
#pragma unroll 1
        for (int m = ex.lane(); m < ka; m += Ex::NL) ndep += dep[m];
        ndep_i = (int)red_sum1(ex, ndep);
#pragma unroll 1
        for (int m = ex.lane(); m < ka; m += Ex::NL) {
            mc[gl::MC_SFD + m] = sd[m]; mc[gl::MC_SFRINV + m] = srinv[m];
            mci[40 + m] = dep[m];
        }
#pragma unroll 1
        for (int e = ex.lane(); e < nz; e += Ex::NL) mc[gl::MC_SF + e] = Sm[e];
        if (ndep_i != 0) {
            // G = Lt'Lt over the kept columns (identity on the skipped ones).
            // Lt[i][a] = U[a][i] = Sm(i, a) (a < i), Lt[a][a] = sd[a]; skipped columns are exact zeros.
#pragma unroll 1
            for (int e = ex.lane(); e < npairs; e += Ex::NL) {
                int c, a;
                tri_index_lower(e, c, a);            // a <= c
                double sacc;
                if (dep[a] || dep[c]) sacc = (a == c) ? 1.0 : 0.0;
                else {
                    sacc = ((a == c) ? sd[c] : Sm[zoff(c) + a]) * sd[c];
#pragma unroll 1
                    for (int i = c + 1; i < ka; i++) sacc += Sm[zoff(i) + a] * Sm[zoff(i) + c];
                }
                if (a == c) gd[a] = sacc; else G[zoff(c) + a] = sacc;
            }
            ex.sync();
            {
                SrcInPlace src;
                src.Z = G; src.diag = gd;
                if (!chol_cols<false>(ex, G, ka, gd, grinv, (int*)nullptr, 0.0, (bool*)nullptr, src)) return false;
            }
#pragma unroll 1
            for (int m = ex.lane(); m < ka; m += Ex::NL) { mc[gl::MC_GD + m] = gd[m]; mc[gl::MC_GRINV + m] = grinv[m]; }
#pragma unroll 1
            for (int e = ex.lane(); e < nz; e += Ex::NL) mc[gl::MC_G + e] = G[e];
            flops += 2.0 * ka * ka * ka / 3.0;
        }
#pragma unroll 1
        for (int m = ex.lane(); m < ka; m += Ex::NL) mci[m] = act[m];
        ex.sync();
        if (ex.lane() == 0) { mci[81] = version; mci[82] = ndep_i; mci[80] = ka; }
    } else {
#pragma unroll 1
        for (int m = ex.lane(); m < ka; m += Ex::NL) { sd[m] = mc[gl::MC_SFD + m]; srinv[m] = mc[gl::MC_SFRINV + m]; dep[m] = mci[40 + m]; }
        if (ndep_i != 0) {
#pragma unroll 1
            for (int m = ex.lane(); m < ka; m += Ex::NL) { gd[m] = mc[gl::MC_GD + m]; grinv[m] = mc[gl::MC_GRINV + m]; }
            if (g_later) {                         // the three triangles did not fit: the factor of G replaces S, which is dead by now
                Gfac = big;
                ex.copy_in(Gfac, mc + gl::MC_G, nz);
            }
        }
        *flags_io |= 64;
    }
    ex.sync();
    if (ndep_i == 0) {
        tri_solve<false>(ex, Sfac, ka, srinv, dl, true, true);
    } else {
        // u = Lt' rho
#pragma unroll 1
        for (int a = ex.lane(); a < ka; a += Ex::NL) {
            double sacc = sd[a] * rho_[a];
#pragma unroll 1
            for (int i = a + 1; i < ka; i++) sacc += Sfac[zoff(i) + a] * rho_[i];
            u1[a] = dep[a] ? 0.0 : sacc;
        }
        ex.sync();
        tri_solve<false>(ex, Gfac, ka, grinv, u1, true, true);
        // consistency: Lt u1 is the projection of rho on range(S); it must reproduce rho
        double mm[2] = {0.0, 0.0};
#pragma unroll 1
        for (int i = ex.lane(); i < ka; i += Ex::NL) {
            double sacc = sd[i] * u1[i];
#pragma unroll 1
            for (int a = 0; a < i; a++) sacc += Sfac[zoff(i) + a] * u1[a];
            mm[0] = fmax(mm[0], fabs(sacc - rho_[i]));
            mm[1] = fmax(mm[1], fabs(C[act[i] * LDH + NMAIN]));
        }
        red_max<2>(ex, mm);
        if (mm[0] > 1.0e-9 * (mm[1] + 1.0)) {
#ifdef WBC_EMU_DEBUG
            printf("fallback inconsistent ka=%d worst=%.3e scale=%.3e\n", ka, mm[0], mm[1]);
#endif
            return false;
        }
        tri_solve<false>(ex, Gfac, ka, grinv, u1, true, true);
#pragma unroll 1
        for (int i = ex.lane(); i < ka; i += Ex::NL) {
            double sacc = sd[i] * u1[i];
#pragma unroll 1
            for (int a = 0; a < i; a++) sacc += Sfac[zoff(i) + a] * u1[a];
            dl[i] = sacc;
        }
        ex.sync();
        flops += 2.0 * ka * ka * ka / 3.0;
        *flags_io |= 16;
    }
#pragma unroll 1
    for (int i = ex.lane(); i < ktotal; i += Ex::NL) nulcest[i] = 0.0;
    ex.sync();
#pragma unroll 1
    for (int m = ex.lane(); m < ka; m += Ex::NL) nulcest[act[m]] = nu0[m] + dl[m];
    ex.sync();
    *flops_io += flops;
    return true;
}

// ------------------------------------------------------------------------------------------------
// generateexmodel (opt.cpp:41594-41740): extended box-QP in [x; slacks], in block form (H, CI, rho) + exb.
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

// next element of the row-major 30 x 30 upper triangle, `step` positions ahead of (i, j)  (integer only; replaces a
// tri_index30 -- sqrtf and two correction loops -- per element in the loops that walk the triangle with a lane stride)
WBC_HD void tri_advance30(int& i, int& j, int step)
{
    j += step;
    while (j >= NMAIN && i < NMAIN - 1) { j = j - NMAIN + i + 1; i++; }      // past the last element: (i, j) is left invalid, unused
}

template <bool SPILL, class Ex>
WBC_HDNI_G void generate_ex_model(const Ex ex, const Work w, int nec, int nic, double rho)
{
    const int n = NMAIN + nic, kw = nec + nic;
    double* H = W_H(w);
    double* CI = QS<SPILL>::CI(w);
    const double* A = W_A(w);
    const double* nulc = W_NULC(w);
    double* exb = QS<SPILL>::EXB(w);
    // The working rows of C are read in place when spilled.  Otherwise they are staged in front of and INTO the CI array:
    // the equality rows end where CI begins, so the working inequality rows land in CI itself and are scaled by rho in
    // place at the end.  EXB lies inside the staging area: the linear term is collected at the head of the (idle) factor
    // array and moved once the rows are dead.  (Caller guarantees nec * LDH + 48 <= sl::STAGE_EQ_CAP.)
    const double* Cs;
    double* exbt = exb;
    if (SPILL) Cs = W_C(w);
    else {
        double* st = SM_(w, sl::OFF_CI) - nec * LDH;
        ex.copy_in(st, W_C(w), kw * LDH);
        ex.sync();
        Cs = st;
        exbt = SM_(w, sl::OFF_Z);
    }
    // quadratic term, main block: A + rho * C'C
    int i, j;
    tri_index30(ex.lane(), i, j);
#pragma unroll 1
    for (int e = ex.lane(); e < 465; e += Ex::NL, tri_advance30(i, j, Ex::NL)) {
        const double aij = A[i * LDH + j];          // global (L2): issued before the row loop so that its latency is covered
        double s0 = 0.0, s1 = 0.0;
        int r = 0;
#pragma unroll 1
        for (; r + 1 < kw; r += 2) {
            s0 += Cs[r * LDH + i] * Cs[r * LDH + j];
            s1 += Cs[(r + 1) * LDH + i] * Cs[(r + 1) * LDH + j];
        }
        if (r < kw) s0 += Cs[r * LDH + i] * Cs[r * LDH + j];
        const double v = aij + rho * (s0 + s1);
        H[i * LDH + j] = v;
        H[j * LDH + i] = v;
    }
    // linear term (41650-41657, 41734-41737): per element, rows in order, two updates per row
#pragma unroll 1
    for (int i = ex.lane(); i < n; i += Ex::NL) {
        double v;
        if (i < NMAIN) {
            v = W_B(w)[i];
#pragma unroll 1
            for (int r = 0; r < kw; r++) {
                const double c = Cs[r * LDH + i];
                v += c * (-rho * Cs[r * LDH + NMAIN]);
                v += c * (-nulc[r]);
            }
        } else {
            const int r = nec + (i - NMAIN);
            v = 0.0;
            v += 1.0 * (-rho * Cs[r * LDH + NMAIN]);
            v += 1.0 * (-nulc[r]);
        }
        exbt[i] = v;
    }
    ex.sync();
    // slack columns: CI = rho * (working inequality rows); in the shared-memory case that is an in-place scaling, and a zero
    // row pads the count to even for the two-at-a-time products
#pragma unroll 1
    for (int e = ex.lane(); e < nic * LDH; e += Ex::NL) {
        const int k = e / LDH, i = e - k * LDH;
        if (i < NMAIN) CI[k * LDH + i] = 0.0 + rho * Cs[(nec + k) * LDH + i];
    }
    if (!SPILL) {
        if (nic & 1)
#pragma unroll 1
            for (int i = ex.lane(); i < LDH; i += Ex::NL) CI[nic * LDH + i] = 0.0;
#pragma unroll 1
        for (int i = ex.lane(); i < n; i += Ex::NL) exb[i] = exbt[i];
    }
    ex.sync();
}

// ------------------------------------------------------------------------------------------------
// Working-set expansion (opt.cpp:41350-41390) and eviction (41400-41418): literal sequential selection
// (arg-max by butterfly, ties to the lowest index like the reference's strict '>' scan).
// Results: iscr[1] = new nicwork, iscr[2] = extended flag.
template <class Ex>
WBC_HDNI_G void update_working_set(const Ex ex, const Work w, int nec, int nictotal, int nicwork, int allowevict)
{
    const int l = ex.lane();
    double* nicerr = W_NICERR(w);
    int* nicnact = W_NICNACT(w);
    double* C = W_C(w);
    double* exxc = W_EXXC(w);
    double* nulc = W_NULC(w);
    int extended = 0, added = 0, evicted = 0;
#pragma unroll 1
    while ((double)added < 1 + 0.20 * NMAIN && nicwork < nictotal) {
        // k = argmax_{j >= nicwork} nicerr[j], first maximum
        double bv = -1.7976931348623157e308;
        int bk = 0x7fffffff;
#pragma unroll 1
        for (int j = nicwork + l; j < nictotal; j += Ex::NL) {
            const double v = nicerr[j];
            if (v > bv) { bv = v; bk = j; }
        }
#pragma unroll 1
        for (int o = Ex::NL / 2; o > 0; o >>= 1) {
            const double ov = ex.shfl_xor(bv, o);
            const int ok = ex.shfl_xori(bk, o);
            if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
        }
        const int k = bk;
        if (!(bv > 0.0)) break;
        // swap rows nec+nicwork <-> nec+k of C, and the per-constraint bookkeeping
        if (k != nicwork) {
#pragma unroll 1
            for (int j = l; j < LDH; j += Ex::NL) {
                const double t = C[(nec + nicwork) * LDH + j];
                C[(nec + nicwork) * LDH + j] = C[(nec + k) * LDH + j];
                C[(nec + k) * LDH + j] = t;
            }
        }
        if (l == 0) {
            const double t = nicerr[nicwork]; nicerr[nicwork] = nicerr[k]; nicerr[k] = t;
            const int ti = nicnact[nicwork]; nicnact[nicwork] = nicnact[k]; nicnact[k] = ti;
            exxc[NMAIN + nicwork] = 0.0;
            nulc[nec + nicwork] = 0.0;
            nicnact[nicwork] = nicnact[nicwork] + 1;
        }
        ex.sync();
        nicwork++; added++;
        extended = 1;
    }
    if (allowevict) {
#pragma unroll 1
        for (int k = nicwork - 1; k >= 0; k--) {
            if (nicerr[k] < -0.01 && nicnact[k] <= 1) {
                const int last = nicwork - 1;
                ex.sync();
                if (k != last) {
#pragma unroll 1
                    for (int j = l; j < LDH; j += Ex::NL) {
                        const double t = C[(nec + last) * LDH + j];
                        C[(nec + last) * LDH + j] = C[(nec + k) * LDH + j];
                        C[(nec + k) * LDH + j] = t;
                    }
                }
                if (l == 0) {
                    double t = nicerr[last]; nicerr[last] = nicerr[k]; nicerr[k] = t;
                    const int ti = nicnact[last]; nicnact[last] = nicnact[k]; nicnact[k] = ti;
                    t = exxc[NMAIN + last]; exxc[NMAIN + last] = exxc[NMAIN + k]; exxc[NMAIN + k] = t;
                    t = nulc[nec + last]; nulc[nec + last] = nulc[nec + k]; nulc[nec + k] = t;
                }
                ex.sync();
                nicwork--;
                evicted = 1;
            }
        }
    }
    // iscr[6]: working-set version, bumped whenever rows of C moved (keys the multiplier-update cache)
    if (l == 0) { W_ISCR(w)[1] = nicwork; W_ISCR(w)[2] = extended; if (extended || evicted) W_ISCR(w)[6] = W_ISCR(w)[6] + 1; }
    ex.sync();
}

// ------------------------------------------------------------------------------------------------
// Set-up: autodiag scale (opt.cpp:48146-48185), scaleshiftoriginalproblem (42088-42339), normalizequadraticterm
// (42374-42446), selectinitialworkingset (42474-42523).  On entry: Q (lower triangle used, opt.cpp:4962/18959) in the H
// array (ld 31), c in exb[0..30), L rows in the global C array.  Returns 0, or -9 for a non-positive diagonal.
template <class Ex>
WBC_HDNI_M int setup_problem(const Ex ex, const Work w, int nrows, int* pd_out, int dup0s, int dup0c, int dup1s, int dup1c)
{
    double* As = W_H(w);
    double* sc = W_SC(w);
    double* b = W_B(w);
    double* C = W_C(w);
    double* stage = SM_(w, sl::STAGE0);                 // everything behind H but the ints: 43 staged rows on the device
    constexpr int CHUNK = (sl::OFF_INT - sl::STAGE0) / LDH;
    double bad = 0.0;
#pragma unroll 1
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) {
        const double d = As[i * LDH + i];
        if (d <= 0.0) bad = 1.0;
        sc[i] = 1.0 / sqrt(d);
    }
    bad = red_sum1(ex, bad);
    ex.sync();
    if (bad != 0.0) return -9;
    // A <- S A S from the lower triangle, mirrored; Frobenius norm
    double an = 0.0;
    int i, j;
    tri_index30(ex.lane(), i, j);                       // i <= j: element (j, i) of the lower triangle
#pragma unroll 1
    for (int e = ex.lane(); e < 465; e += Ex::NL, tri_advance30(i, j, Ex::NL)) {
        const double v = As[j * LDH + i] * sc[i] * sc[j];
        As[i * LDH + j] = v;
        As[j * LDH + i] = v;
        an += (i == j) ? v * v : 2.0 * (v * v);
    }
#pragma unroll 1
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) b[i] = W_EXB(w)[i] * sc[i];
    an = sqrt(red_sum1(ex, an));
    ex.sync();
    // constraint rows in chunks through shared memory: scale by S, normalise (42219-42314), c_r' A c_r
    double maxcac = 0.0;
#pragma unroll 1
    for (int r0 = 0; r0 < nrows; r0 += CHUNK) {
        const int nr = (nrows - r0 < CHUNK) ? nrows - r0 : CHUNK;
        ex.copy_in(stage, C + r0 * LDH, nr * LDH);
        ex.sync();
#pragma unroll 1
        for (int r = ex.lane(); r < nr; r += Ex::NL) {
            double* row = stage + r * LDH;
            double vv = 0.0;
#pragma unroll 1
            for (int j = 0; j < NMAIN; j++) {
                const double v = row[j] * sc[j];
                row[j] = v;
                vv += v * v;
            }
            double rhs = row[NMAIN];
            vv = sqrt(vv);
            if (vv > 0.0) {
                vv = 1.0 / vv;
#pragma unroll 1
                for (int j = 0; j < NMAIN; j++) row[j] *= vv;
                rhs *= vv;
            }
            row[NMAIN] = rhs;
        }
        ex.sync();
        // c_r' A c_r, one row at a time across the lanes (lane = column).  Exact zeros of c_r add nothing, so only its
        // non-zero entries are visited (in ascending order: same sums); the rows have 1 (joint limits), 3 (friction) or
        // 12..24 (dynamics, torque limits) of them out of 30.
#pragma unroll 1
        for (int r = 0; r < nr; r++) {
            const int gr = r0 + r;
            if ((gr >= dup0s && gr < dup0s + dup0c) || (gr >= dup1s && gr < dup1s + dup1c)) continue;   // negated twin of an earlier row
            const double* row = stage + r * LDH;
            double part = 0.0;
            if (Ex::NL >= NMAIN) {
                const int j = ex.lane();
                const double cj = (j < NMAIN) ? row[j] : 0.0;
                unsigned nz = ex.ballot(cj != 0.0);
                if ((nz & (nz - 1u)) == 0u && nz != 0u) {
                    // one non-zero (the 24 joint-limit rows): the sum over the lanes is the single term (c_k A_kk) c_k
                    const int k = lowest_bit(nz);
                    const double ck = row[k];
                    maxcac = fmax(maxcac, fabs((ck * As[k * LDH + k]) * ck));
                    continue;
                }
                double t = 0.0;
#pragma unroll 1
                while (nz) {
                    const int k = lowest_bit(nz);
                    nz &= nz - 1u;
                    t += row[k] * As[k * LDH + j];          // lanes >= 30 read the pad column / next row: finite, dropped
                }
                part = t * cj;
            } else {
#pragma unroll 1
                for (int j = ex.lane(); j < NMAIN; j += Ex::NL) {
                    const double cj = row[j];
                    double t = 0.0;
#pragma unroll 1
                    for (int k = 0; k < NMAIN; k++) {
                        const double ck = row[k];
                        if (ck != 0.0) t += ck * As[k * LDH + j];
                    }
                    part += t * cj;
                }
            }
            maxcac = fmax(maxcac, fabs(red_sum1(ex, part)));
        }
#pragma unroll 4
        for (int e = ex.lane(); e < nr * LDH; e += Ex::NL) C[r0 * LDH + e] = stage[e];
        ex.sync();
    }
    double targetscale = fmax(maxcac, an / NMAIN);
    if (targetscale == 0.0) targetscale = 1.0;
    const double v = 1.0 / targetscale;
    double* Ag = W_A(w);
#pragma unroll 1
    for (int e = ex.lane(); e < NMAIN * LDH; e += Ex::NL) {
        const double a = As[e] * v;
        As[e] = a;
        Ag[e] = a;
    }
#pragma unroll 1
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) b[i] *= v;
    ex.sync();
    // Cholesky of A: convexity test (42474-42523); the factor is kept for the reduced multiplier update
    double* Zs = SM_(w, sl::OFF_Z);
    double* ladiag = SM_(w, sl::OFF_XC);                // not kept
    struct SrcA {
        const double* A;
        WBC_HD double operator()(int k, int c) const { return A[k * LDH + c]; }
    } src;
    src.A = As;
    const bool pd = chol_cols<false>(ex, Zs, NMAIN, ladiag, W_LARINV(w), (int*)nullptr, 0.0, (bool*)nullptr, src);
    double* LA = W_LA(w);
#pragma unroll 1
    for (int i = ex.lane(); i < zoff(NMAIN); i += Ex::NL) LA[i] = Zs[i];
    ex.sync();
    *pd_out = pd ? 1 : 0;
    return 0;
}

}  // namespace wbcqp
#include "qp_fast.cuh"
namespace wbcqp {

// ------------------------------------------------------------------------------------------------
// One pass of the working-set loop body for a given storage mode: model, QQP, violations.
template <bool SPILL, class Ex>
WBC_HD int model_and_qqp(const Ex& ex, const Work& w, int nec, int nicwork, double rho, double epsx, int* ncholesky, double* flops)
{
    generate_ex_model<SPILL>(ex, w, nec, nicwork, rho);
    *flops += (double)NMAIN * NMAIN * (nec + nicwork) + 4.0 * NMAIN * (nec + nicwork);
    return qqp_optimize<SPILL>(ex, w, nicwork, rho, 0.01 * epsx, 50, ncholesky, flops);
}
#if defined(__CUDACC__)
// device, shared-memory case: the register-resident QQP (the generic one is not instantiated: its vectors do not exist there)
template <class Ex>
__device__ __forceinline__ int model_and_qqp_dev(const Ex& ex, const Work& w, int nec, int nicwork, double rho, double epsx,
                                                 int* ncholesky, double* flops, int* reused)
{
    generate_ex_model<false>(ex, w, nec, nicwork, rho);
    *flops += (double)NMAIN * NMAIN * (nec + nicwork) + 4.0 * NMAIN * (nec + nicwork);
    return fast::qqp_optimize_fast(w, nicwork, rho, 0.01 * epsx, 50, ncholesky, flops, reused, (const double*)nullptr);
}
#endif

// ------------------------------------------------------------------------------------------------
// Row-wise products with the constraint matrix.  C lives in global memory (row-major), so rows are first copied --
// coalesced -- into the idle arrays behind H in shared memory, then each lane takes a row (stride 31: conflict free).
constexpr int STAGE_ROWS = sl::STAGE_CAP / LDH;      // 40 on the device
template <class Ex>
WBC_HD const double* stage_rows(const Ex& ex, const Work& w, int row0, int nr)
{
    double* st = SM_(w, sl::STAGE0);
    ex.copy_in(st, W_C(w) + row0 * LDH, nr * LDH);
    ex.sync();
    return st;
}
// violations of all inequality rows w.r.t. the main variables only (opt.cpp:41330-41335)
template <class Ex>
WBC_HDNI_G void inequality_violations(const Ex ex, const Work w, int nec, int nictotal)
{
    double* nicerr = W_NICERR(w);
    const double* exxc = W_EXXC(w);
#pragma unroll 1
    for (int i0 = 0; i0 < nictotal; i0 += STAGE_ROWS) {
        const int nr = (nictotal - i0 < STAGE_ROWS) ? nictotal - i0 : STAGE_ROWS;
        const double* st = stage_rows(ex, w, nec + i0, nr);
#pragma unroll 1
        for (int i = ex.lane(); i < nr; i += Ex::NL) {
            const double* row = st + i * LDH;
            double v0 = 0.0, v1 = 0.0;
#pragma unroll 1
            for (int j = 0; j < NMAIN; j += 2) { v0 += row[j] * exxc[j]; v1 += row[j + 1] * exxc[j + 1]; }
            nicerr[i0 + i] = (v0 + v1) - row[NMAIN];
        }
        ex.sync();
    }
}
// feasibility error over the working rows and the multiplier hand-over (opt.cpp:41444-41476); returns sum of squares
template <class Ex>
WBC_HDNI_G double feasibility_error(const Ex ex, const Work w, int nec, int kwork)
{
    const double* exxc = W_EXXC(w);
    double* nulc = W_NULC(w);
    const double* nulcest = W_NULCEST(w);
    double fe = 0.0;
#pragma unroll 1
    for (int i0 = 0; i0 < kwork; i0 += STAGE_ROWS) {
        const int nr = (kwork - i0 < STAGE_ROWS) ? kwork - i0 : STAGE_ROWS;
        const double* st = stage_rows(ex, w, i0, nr);
#pragma unroll 1
        for (int ii = ex.lane(); ii < nr; ii += Ex::NL) {
            const int i = i0 + ii;
            const double* row = st + ii * LDH;
            double v = 0.0, vv = 0.0;
#pragma unroll 1
            for (int j = 0; j < NMAIN; j++) { const double c = row[j]; v += c * exxc[j]; vv += c * c; }
            if (i >= nec) { v += exxc[NMAIN + (i - nec)]; vv += 1.0; }
            v -= row[NMAIN];
            if (vv == 0.0) vv = 1.0;
            v = v / sqrt(vv);
            fe += v * v;
            nulc[i] = nulcest[i];
        }
        ex.sync();
    }
    return red_sum1(ex, fe);
}

// The solver.  On entry the warp has staged the problem (see setup_problem).  Result: xs[0..30) in shared memory.
template <class Ex>
WBC_HDN void solve_denseaul(const Ex& ex, const Work& w, const Settings& cfg, int nrows, int neq, Stats& st)
{
    const int nec = neq, nictotal = nrows - neq;
    st.termination = 0; st.ncholesky = 0; st.outer_its = 0; st.qqp_calls = 0; st.nicwork = 0;
    st.kkt_dim_max = 0; st.chol_reused = 0; st.flags = 0; st.flops = 0.0;
    int pd = 0;
    const int rc = setup_problem(ex, w, nrows, &pd, cfg.dup_start[0], cfg.dup_count[0], cfg.dup_start[1], cfg.dup_count[1]);
    if (rc != 0) { st.termination = rc; return; }
    st.flops += 2.0 * nrows * NMAIN * NMAIN + 9000.0;

    if (ex.lane() == 0) {
        W_ISCR(w)[6] = 0;                                                        // working-set version
        reinterpret_cast<int*>(w.g + gl::OFF_MC + gl::MC_INT)[80] = -1;          // multiplier-update cache: empty
    }
    ex.sync();
    int nicwork = 0;
    int allowevict = 1;
    if (!pd) { nicwork = nictotal; allowevict = 0; st.flags |= 1; }
    const bool have_factor = pd != 0;
    double* nulc = W_NULC(w);
    double* nulcest = W_NULCEST(w);
    double* exxc = W_EXXC(w);
#pragma unroll 1
    for (int i = ex.lane(); i < nictotal; i += Ex::NL) W_NICNACT(w)[i] = (i < nicwork) ? 1 : 0;
#pragma unroll 1
    for (int i = ex.lane(); i < nrows; i += Ex::NL) nulc[i] = 0.0;
#pragma unroll 1
    for (int i = ex.lane(); i < NMAIN + nictotal; i += Ex::NL) exxc[i] = 0.0;
    ex.sync();

    double rho = cfg.rho, epsx = cfg.epsx;
    if (epsx <= 0.0) epsx = 1.0e-9;
    const double maxrho = 1.0e12, requestedfeasdecrease = 0.33;
    int goodcounter = 0, stagnationcounter = 0;
    double feaserr = 1.7976931348623157e308;   // ae_maxrealnumber
#pragma unroll 1
    for (int outeridx = 0; outeridx < cfg.outerits; outeridx++) {
        st.outer_its++;
        bool extended;
        do {
            int term;
            // shared-memory capacity: NICCAP working inequality rows in CI, the equality rows (and the linear term) in front of it
            if (nicwork > NICCAP || nec * LDH + 48 > sl::STAGE_EQ_CAP) { st.flags |= 32; term = model_and_qqp<true>(ex, w, nec, nicwork, rho, epsx, &st.ncholesky, &st.flops); }
#if defined(__CUDA_ARCH__)
            else term = model_and_qqp_dev(ex, w, nec, nicwork, rho, epsx, &st.ncholesky, &st.flops, &st.chol_reused);
#else
            else term = model_and_qqp<false>(ex, w, nec, nicwork, rho, epsx, &st.ncholesky, &st.flops);
#endif
            st.qqp_calls++;
            if (term == -4) st.flags |= 4;
            inequality_violations(ex, w, nec, nictotal);
            st.flops += 2.0 * nictotal * NMAIN;
            update_working_set(ex, w, nec, nictotal, nicwork, allowevict);
            nicwork = W_ISCR(w)[1];
            extended = W_ISCR(w)[2] != 0;
            ex.sync();
        } while (extended);

        const int kwork = nec + nicwork;
        // multiplier estimate (41438-41439)
#pragma unroll 1
        for (int i = ex.lane(); i < kwork; i += Ex::NL) nulcest[i] = nulc[i];
        ex.sync();
        {
            const int nq = NMAIN + nicwork + kwork;
            if (nq > st.kkt_dim_max) st.kkt_dim_max = nq;
            bool done = false;
            if (cfg.kkt_mode == 1 && have_factor)
                done = update_lagrange_multipliers_reduced(ex, w, nec, nicwork, cfg.kkt_pivtol, &st.flags, &st.flops);
            if (!done) {
                st.flags |= 8;
#pragma unroll 1
                for (int i = ex.lane(); i < kwork; i += Ex::NL) nulcest[i] = nulc[i];
                ex.sync();
                update_lagrange_multipliers_literal(ex, w, nec, nicwork, &st.flops);
            }
        }
        const double feaserrprev = feaserr;
        feaserr = sqrt(feasibility_error(ex, w, nec, kwork));
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
#pragma unroll 1
    for (int i = ex.lane(); i < NMAIN; i += Ex::NL) W_XS(w)[i] = W_SC(w)[i] * exxc[i] + 0.0;
    ex.sync();
    st.termination = 2;
}

// ------------------------------------------------------------------------------------------------
// The same solver as three resumable stages.  solve_denseaul above runs a solve from start to finish on one warp; the
// control cycle's solver kernel (wbc_b200.cu) instead hands a solve from warp to warp at the two points where the code it
// executes changes completely, so that every SM runs one kind of stage and its instruction cache holds it:
//     SETUP   set-up of the scaled problem                          (solve_stage_setup;  caller assembled Q, c, L)
//     QLOOP   one outer iteration's model / QQP / working-set loop  (solve_stage_qloop)
//     UPDATE  multiplier update, feasibility, penalty update        (solve_stage_update; sets `done` on the last one)
// Everything a later stage needs is already in the solve's global scratch block (C, A, the factor of A, multipliers,
// working-set bookkeeping, the multiplier-update cache) except the point exxc and a few scalars: SolveState, kept at
// gl::OFF_HDR.  Same routines in the same order on the same operands as solve_denseaul: results are bit-identical
// (tests/test_emulation.py runs both on the host; tools/gpu_dump.py on the device).
struct SolveState {
    double rho, epsx, feaserr, flops;
    int nrows, nec, nicwork, allowevict, have_factor, goodcounter, stagnationcounter, outeridx;
    int version;                        // working-set version (keys the multiplier-update cache)
    int termination, ncholesky, outer_its, qqp_calls, kkt_dim_max, chol_reused, flags;
    int done;                           // the solve is complete: result in W_XS, termination set
    int pad_;
    double qstat[3];                    // |A| statistics of the next QQP call's model (qqp_stats), computed where the model is built
};
static_assert(sizeof(SolveState) <= gl::HDR_DOUBLES * sizeof(double), "SolveState fits its header slot");

template <class Ex>
WBC_HDN void stage_store(const Ex& ex, const Work& w, const SolveState& s)
{
    double* hx = w.g + gl::OFF_HDR + gl::HDR_DOUBLES;
    const double* exxc = W_EXXC(w);
    const int n = NMAIN + (s.nrows - s.nec);
#pragma unroll 1
    for (int i = ex.lane(); i < n; i += Ex::NL) hx[i] = exxc[i];
    if (ex.lane() == 0) {
        SolveState t = s;
        t.version = W_ISCR(w)[6];
        *reinterpret_cast<SolveState*>(w.g + gl::OFF_HDR) = t;
    }
    ex.sync();
}
template <class Ex>
WBC_HDN void stage_load(const Ex& ex, const Work& w, SolveState& s)
{
    s = *reinterpret_cast<const SolveState*>(w.g + gl::OFF_HDR);
    const double* hx = w.g + gl::OFF_HDR + gl::HDR_DOUBLES;
    double* exxc = W_EXXC(w);
    const int n = NMAIN + (s.nrows - s.nec);
#pragma unroll 1
    for (int i = ex.lane(); i < n; i += Ex::NL) exxc[i] = hx[i];
    if (ex.lane() == 0) W_ISCR(w)[6] = s.version;
    ex.sync();
}
WBC_HD void stage_stats(const SolveState& s, Stats& st)
{
    st.termination = s.termination; st.ncholesky = s.ncholesky; st.outer_its = s.outer_its; st.qqp_calls = s.qqp_calls;
    st.nicwork = s.nicwork; st.kkt_dim_max = s.kkt_dim_max; st.chol_reused = s.chol_reused; st.flags = s.flags; st.flops = s.flops;
}

// SETUP.  On entry the warp has staged the problem (see setup_problem).  Leaves the state of the first outer iteration.
template <class Ex>
WBC_HDN void solve_stage_setup(const Ex& ex, const Work& w, const Settings& cfg, int nrows, int neq, SolveState& s)
{
    const int nec = neq, nictotal = nrows - neq;
    s.nrows = nrows; s.nec = nec;
    s.termination = 0; s.ncholesky = 0; s.outer_its = 0; s.qqp_calls = 0; s.nicwork = 0;
    s.kkt_dim_max = 0; s.chol_reused = 0; s.flags = 0; s.flops = 0.0; s.done = 0; s.pad_ = 0; s.version = 0;
    s.goodcounter = 0; s.stagnationcounter = 0; s.outeridx = 0; s.allowevict = 1; s.have_factor = 0;
    s.rho = cfg.rho; s.epsx = cfg.epsx; s.feaserr = 1.7976931348623157e308;
    s.qstat[0] = s.qstat[1] = s.qstat[2] = 0.0;
    int pd = 0;
    const int rc = setup_problem(ex, w, nrows, &pd, cfg.dup_start[0], cfg.dup_count[0], cfg.dup_start[1], cfg.dup_count[1]);
    if (rc != 0) { s.termination = rc; s.done = 1; return; }
    s.flops += 2.0 * nrows * NMAIN * NMAIN + 9000.0;
    if (ex.lane() == 0) {
        W_ISCR(w)[6] = 0;
        reinterpret_cast<int*>(w.g + gl::OFF_MC + gl::MC_INT)[80] = -1;
    }
    ex.sync();
    int nicwork = 0;
    if (!pd) { nicwork = nictotal; s.allowevict = 0; s.flags |= 1; }
    s.have_factor = pd != 0;
    s.nicwork = nicwork;
    double* nulc = W_NULC(w);
    double* exxc = W_EXXC(w);
#pragma unroll 1
    for (int i = ex.lane(); i < nictotal; i += Ex::NL) W_NICNACT(w)[i] = (i < nicwork) ? 1 : 0;
#pragma unroll 1
    for (int i = ex.lane(); i < nrows; i += Ex::NL) nulc[i] = 0.0;
#pragma unroll 1
    for (int i = ex.lane(); i < NMAIN + nictotal; i += Ex::NL) exxc[i] = 0.0;
    ex.sync();
    if (s.epsx <= 0.0) s.epsx = 1.0e-9;
}

// QLOOP of outer iteration s.outeridx.
template <class Ex>
WBC_HDN void solve_stage_qloop(const Ex& ex, const Work& w, SolveState& s)
{
    const int nec = s.nec, nictotal = s.nrows - s.nec;
    int nicwork = s.nicwork;
    s.outer_its++;
    bool extended;
    do {
        int term;
        if (nicwork > NICCAP || nec * LDH + 48 > sl::STAGE_EQ_CAP) { s.flags |= 32; term = model_and_qqp<true>(ex, w, nec, nicwork, s.rho, s.epsx, &s.ncholesky, &s.flops); }
#if defined(__CUDA_ARCH__)
        else term = model_and_qqp_dev(ex, w, nec, nicwork, s.rho, s.epsx, &s.ncholesky, &s.flops, &s.chol_reused);
#else
        else term = model_and_qqp<false>(ex, w, nec, nicwork, s.rho, s.epsx, &s.ncholesky, &s.flops);
#endif
        s.qqp_calls++;
        if (term == -4) s.flags |= 4;
        inequality_violations(ex, w, nec, nictotal);
        s.flops += 2.0 * nictotal * NMAIN;
        update_working_set(ex, w, nec, nictotal, nicwork, s.allowevict);
        nicwork = W_ISCR(w)[1];
        extended = W_ISCR(w)[2] != 0;
        ex.sync();
    } while (extended);
    s.nicwork = nicwork;
}

// UPDATE of outer iteration s.outeridx; on the last one the unscaled solution is left in W_XS and s.done is set.
template <class Ex>
WBC_HDN void solve_stage_update(const Ex& ex, const Work& w, const Settings& cfg, SolveState& s)
{
    const int nec = s.nec, nicwork = s.nicwork;
    const int kwork = nec + nicwork;
    double* nulc = W_NULC(w);
    double* nulcest = W_NULCEST(w);
    const double* exxc = W_EXXC(w);
#pragma unroll 1
    for (int i = ex.lane(); i < kwork; i += Ex::NL) nulcest[i] = nulc[i];
    ex.sync();
    {
        const int nq = NMAIN + nicwork + kwork;
        if (nq > s.kkt_dim_max) s.kkt_dim_max = nq;
        bool done = false;
        if (cfg.kkt_mode == 1 && s.have_factor)
            done = update_lagrange_multipliers_reduced(ex, w, nec, nicwork, cfg.kkt_pivtol, &s.flags, &s.flops);
        if (!done) {
            s.flags |= 8;
#pragma unroll 1
            for (int i = ex.lane(); i < kwork; i += Ex::NL) nulcest[i] = nulc[i];
            ex.sync();
            update_lagrange_multipliers_literal(ex, w, nec, nicwork, &s.flops);
        }
    }
    const double maxrho = 1.0e12, requestedfeasdecrease = 0.33;
    const double feaserrprev = s.feaserr;
    s.feaserr = sqrt(feasibility_error(ex, w, nec, kwork));
    ex.sync();
    s.flops += 4.0 * kwork * NMAIN;
    if (s.feaserr < s.epsx) s.goodcounter++; else s.goodcounter = 0;
    if (s.feaserr > feaserrprev * requestedfeasdecrease) s.stagnationcounter++; else s.stagnationcounter = 0;
    bool last = s.goodcounter >= 2;
    if (!last) {
        if (s.stagnationcounter >= 2) s.rho = fmin(s.rho * 10.0, maxrho);
        else s.rho = fmin(s.rho * 1.41, maxrho);
        s.outeridx++;
        last = s.outeridx >= cfg.outerits;
    }
    if (last) {
#pragma unroll 1
        for (int i = ex.lane(); i < NMAIN; i += Ex::NL) W_XS(w)[i] = W_SC(w)[i] * exxc[i] + 0.0;
        ex.sync();
        s.termination = 2;
        s.done = 1;
    }
}

// ---- Finer stages for the solver kernel's SM roles: the QQP iteration alone (its hot code is ~25 KB and fits an SM's
// instruction cache) against everything else.
//   stage_post_and_model  after a QQP call: violations, working set, and -- when the working set has settled -- the multiplier
//                         update; then the extended model of the NEXT QQP call, written to the solve's global block.
//                         Returns 1 when a QQP task must follow, 0 when the solve is complete (s.done).
//                         Spilled working sets (rare) run their generic QQP right here.
//   stage_qqp_load/store  (device) the QQP task: model in, qqp_optimize_fast, point out.
template <class Ex>
WBC_HDN int stage_post_and_model(const Ex& ex, const Work& w, const Settings& cfg, SolveState& s, bool after_qqp)
{
    const int nec = s.nec, nictotal = s.nrows - s.nec;
    for (;;) {
        if (after_qqp) {
            s.qqp_calls++;
            inequality_violations(ex, w, nec, nictotal);
            s.flops += 2.0 * nictotal * NMAIN;
            update_working_set(ex, w, nec, nictotal, s.nicwork, s.allowevict);
            s.nicwork = W_ISCR(w)[1];
            const bool extended = W_ISCR(w)[2] != 0;
            ex.sync();
            if (!extended) {
                solve_stage_update(ex, w, cfg, s);
                if (s.done) return 0;
                s.outer_its++;
            }
        }
        if (s.nicwork > NICCAP || nec * LDH + 48 > sl::STAGE_EQ_CAP) {
            s.flags |= 32;
            const int term = model_and_qqp<true>(ex, w, nec, s.nicwork, s.rho, s.epsx, &s.ncholesky, &s.flops);
            if (term == -4) s.flags |= 4;
            after_qqp = true;
            continue;
        }
        generate_ex_model<false>(ex, w, nec, s.nicwork, s.rho);
        s.flops += (double)NMAIN * NMAIN * (nec + s.nicwork) + 4.0 * NMAIN * (nec + s.nicwork);
        // hand the model over: H, the (even-padded) CI rows, the linear term
        const int nic2 = (s.nicwork + 1) & ~1;
        double* gh = w.g + gl::OFF_HQ;
        double* gc = w.g + gl::OFF_CI;
        double* gb = w.g + gl::OFF_EXBG;
        const double* H = W_H(w);
        const double* CI = SM_(w, sl::OFF_CI);
        const double* exb = W_EXB(w);
#pragma unroll 1
        for (int k = ex.lane(); k < NMAIN * LDH; k += Ex::NL) gh[k] = H[k];
#pragma unroll 1
        for (int k = ex.lane(); k < nic2 * LDH; k += Ex::NL) gc[k] = CI[k];
#pragma unroll 1
        for (int k = ex.lane(); k < NMAIN + s.nicwork; k += Ex::NL) gb[k] = exb[k];
        ex.sync();
#if defined(__CUDA_ARCH__)
        {   // the model's |A| statistics (opt.cpp:29893-29915): once per call, so they are taken here and not on the QQP SMs
            const int l = ex.lane();
            fast::qqp_stats(s.nicwork, s.rho, l < NMAIN ? exb[l] : 0.0, l < s.nicwork ? exb[NMAIN + l] : 0.0);
            const double* sp = SM_(w, sl::OFF_SP);
            s.qstat[0] = sp[8]; s.qstat[1] = sp[9]; s.qstat[2] = sp[10];
        }
#endif
        return 1;
    }
}
#if defined(__CUDACC__)
// The QQP task: stage the model (one cp.async round trip), iterate, leave the point in EXXC for stage_store.
template <class Ex>
__device__ __forceinline__ void stage_qqp(const Ex& ex, const Work& w, SolveState& s)
{
    const int nic2 = (s.nicwork + 1) & ~1;
    ex.copy_start(W_H(w), w.g + gl::OFF_HQ, NMAIN * LDH);
    if (nic2 > 0) ex.copy_start(SM_(w, sl::OFF_CI), w.g + gl::OFF_CI, nic2 * LDH);
    ex.copy_start(W_EXB(w), w.g + gl::OFF_EXBG, NMAIN + s.nicwork);
    ex.copy_wait();
    const int term = fast::qqp_optimize_fast(w, s.nicwork, s.rho, 0.01 * s.epsx, 50, &s.ncholesky, &s.flops, &s.chol_reused, s.qstat);
    if (term == -4) s.flags |= 4;
}
#endif

// The three stages back to back on one executor (the host emulation's check that they reproduce solve_denseaul).
template <class Ex>
WBC_HDN void solve_staged(const Ex& ex, const Work& w, const Settings& cfg, int nrows, int neq, Stats& st)
{
    SolveState s;
    solve_stage_setup(ex, w, cfg, nrows, neq, s);
    while (!s.done) {
        stage_store(ex, w, s);
        stage_load(ex, w, s);
        solve_stage_qloop(ex, w, s);
        stage_store(ex, w, s);
        stage_load(ex, w, s);
        solve_stage_update(ex, w, cfg, s);
    }
    stage_stats(s, st);
}

}  // namespace wbcqp
