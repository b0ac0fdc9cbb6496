// b200mpc solver core: one nonlinear-MPC problem per thread, primal-dual interior point with a
// structure-exploiting Riccati (block-tridiagonal Schur-complement) KKT solve.
//
// What this replaces in the reference (/root/reference):
//   mpc_to_line/solution/MPC.cpp:45-140   FG_eval cost + constraints (here: closed-form residuals, Jacobian
//                                         blocks A_t,B_t and Lagrangian-Hessian blocks, no tape)
//   mpc_to_line/solution/MPC.cpp:149-257  MPC::Solve (start point, bounds, solve, returned 8-vector)
//   Ipopt-3.12.7/Ipopt/src/Algorithm/     the interior-point algorithm with the default options MPC.cpp leaves
//                                         untouched: IpDefaultIterateInitializer.cpp:175-344, IpLeastSquareMults.cpp:40-94,
//                                         IpMonotoneMuUpdate.cpp:132-232, IpPDSearchDirCalc.cpp:60-139,
//                                         IpPDPerturbationHandler.cpp:148-420, IpBacktrackingLineSearch.cpp:261-797,852-941,
//                                         IpFilterLSAcceptor.cpp:227-587,800-813, IpIpoptAlg.cpp:559-727,880-951,
//                                         IpOptErrorConvCheck.cpp:204-329, IpGradientScaling.cpp:69-119,
//                                         IpOrigIpoptNLP.cpp:361-372,466-482,875-883
//                                         soft restoration phase: IpBacktrackingLineSearch.cpp:426-448,498-530,595-603,
//                                         1043-1140, IpIpoptCalculatedQuantities.cpp:2835-2884 (Solver::resto_entry);
//                                         Ipopt's restoration phase proper (IpRestoMinC_1Nrm.cpp) is NOT restated: a
//                                         closed-form restoration step stands in for its nested solve (Solver::do_resto)
//   Ipopt .../LinearSolvers + MUMPS       generic sparse LDL^T of the 348x348 augmented system -> the stage-wise
//                                         Riccati recursion below (inertia test = all 2x2 control blocks positive definite)
//
// The same source is compiled by nvcc for the device and, for the CPU-side algorithm tests only
// (tests/hostsim), by g++.  The product (libb200mpc.so) has no CPU solve path.
//
// Stage structure.  Variables s_t=(x,y,psi,v,cte,epsi), t=0..N-1, u_t=(delta,a), t=0..N-2.  Constraint rows
// at time t+1:  c_{t+1} = s_{t+1} - F(s_t,u_t).  Newton step of the barrier problem (Ipopt's augmented system
// [[W+Sigma+dw I, J^T],[J, 0]] (dx,dlam) = -(grad_x L_mu, c)) is the LQ problem
//     min 1/2 dx^T H dx + r^T dx   s.t.  ds_{t+1} = A_t ds_t + B_t du_t - c_{t+1},  ds_0 = 0
// whose multipliers are dlam.  cte_{t+1} does not influence any later row (column cte of A is zero) so it is
// folded into stage t as a quadratic on a linear output; the (u_{t+1},u_t) smoothness coupling is carried by
// augmenting the recursion state with the previous control: xi_t = (x,y,psi,v,epsi | delta_{t-1},a_{t-1}) in R^7.
//
// Execution structure.  One interior-point iteration = three sweeps over the horizon ("passes"):
//   FACTOR  (backward)  K1 derivative blocks + K2 Riccati factorisation and feed-forward
//   FORWARD (forward)   K2 back-substitution: dx, fraction-to-the-boundary step sizes, barrier slope
//   STEP    (backward)  K1 residuals/objective/barrier at the trial point x + alpha dx, K2 multiplier recovery, the
//                       trial iterate (x, lambda, z with the kappa_sigma reset) and every error norm at it, then
//                       K3: filter / Armijo acceptance, mu update, convergence test.  The trial iterate is written
//                       to the second copy of the iterate arrays and becomes current only if it is accepted
//                       (first trial accepted in >99% of iterations), so a rejection costs nothing extra.
// Every pass touches the per-problem workspace once per stage.  A problem's scalar algorithm state lives in the
// same workspace (record 0), so each pass can run as its own kernel (its own register budget and occupancy) or
// all three can be looped inside one kernel.
#pragma once
#include <float.h>
#include <math.h>

#ifdef __CUDACC__
#define MPC_HD __host__ __device__ __forceinline__
#else
#define MPC_HD inline
#endif

namespace b200mpc {

// Mirrors b200mpc_params (include/b200mpc.h) + constants derived from it once on the host (finalize()).
struct Params {
  int N;
  double dt, Lf, ref_v;
  double w_cte, w_epsi, w_v, w_delta, w_a, w_ddelta, w_da;
  double delta_max, a_max;
  double tol;
  int max_iter;
  int resto;            // after a failed line search: 0 = status -2, 1 = restoration step (Solver::do_resto),
                        // 2 = Ipopt's soft restoration phase first, then the restoration step (Solver::resto_entry)
  // derived
  double dtLf;          // dt / Lf
  double xl[2], xu[2];  // relaxed bounds of (delta, a): +-(b + 1e-8 max(1,|b|)), IpOrigIpoptNLP.cpp:369-372
  double ob[2];         // original bounds
  double u_init[2];     // start value of the controls after the bound push (IpDefaultIterateInitializer.cpp:469-649)
  MPC_HD void finalize() {
    dtLf = dt / Lf;
    ob[0] = delta_max; ob[1] = a_max;
    for (int j = 0; j < 2; ++j) {
      const double b = ob[j], rel = 1e-8 * (fabs(b) > 1.0 ? fabs(b) : 1.0);
      xl[j] = -b - rel; xu[j] = b + rel;
      const double l = xl[j], u = xu[j];
      double v = 0.0 < l ? l : (0.0 > u ? u : 0.0);
      const double ql = 0.01 * (u - l);
      double pl = 0.01 * (fabs(l) > 1.0 ? fabs(l) : 1.0), pu = 0.01 * (fabs(u) > 1.0 ? fabs(u) : 1.0);
      if (ql < pl) pl = ql;
      if (ql < pu) pu = ql;
      v = v < l + pl ? l + pl : (v > u - pu ? u - pu : v);
      u_init[j] = v;
    }
  }
};

// Ipopt return codes used (Ipopt/src/Interfaces/IpReturnCodes_inc.h:16-39)
enum Status {
  kSolveSucceeded = 0,
  kSolvedToAcceptableLevel = 1,
  kSearchDirectionTooSmall = 3,
  kDivergingIterates = 4,
  kMaxIterExceeded = -1,
  kRestorationFailed = -2,      // line search failed at an almost feasible point (or with the restoration switched off)
  kErrorInStepComputation = -3, // inertia correction exhausted
  kInvalidNumberDetected = -13  // NaN / Inf at the starting point
};

// MPC_FUSE_FACTOR 1: the Riccati factorisation of the NEXT iteration can ride on the STEP sweep (step_sweep<true>,
//                     kernel_stepfactor): the trial iterate the sweep has just computed in registers is factorised stage by
//                     stage instead of being read back by a factor sweep, with the feed-forward split as k = ka + mu kb
//                     because the barrier parameter of the next iteration is only known after the sweep.
//                     Built, verified (tests/test_hostsim_golden.py, 18 GPU parity tests) and measured on the B200
//                     (gpurun_out/r2_fuse*): the fused kernel moves 24 % fewer bytes and executes 9 % fewer instructions
//                     than the step and factor kernels together and takes exactly as long (268-280 us against 156 + 114 us
//                     per round of 65 536 problems) -- at 8 warps per SM the sweeps are bound by the length of the dependent
//                     instruction chain of a stage, and the fused chain is the sum of the two -- while the compaction has to
//                     move the factors too.  13 % slower overall, so it is compiled out of the product (0); the host
//                     build of the CPU tests compiles a copy with the switch on and keeps it alive.
#ifndef MPC_FUSE_FACTOR
#define MPC_FUSE_FACTOR 0
#endif
constexpr int kMaxCoef = 4;   // reference polynomial degree <= 3
constexpr int kMaxFilter = 41;   // filter entries (phi, theta); they fill one workspace record of their own (record N+1)
constexpr int kCarry = 24;    // stage-to-stage values of the STEP sweep
constexpr int kRicCarry = 39; // stage-to-stage values of the Riccati recursion when it rides on the STEP sweep (step_sweep<true>)
constexpr int kKF = 13 + 2 * MPC_FUSE_FACTOR;   // rows of the Riccati factors of one stage
constexpr int kStageVals = 24;   // values of one stage the factor sweep stages asynchronously: S U LAM ZL ZU TR (22) + u_{t-1} (2)
constexpr int kStepStageVals = 32;   // step sweep: S U LAM ZL ZU TR of stage t (22), DS of stage t (6), u_{t-1} (2), du_{t-1} (2)
#ifndef MPC_RESTO_BETA
#define MPC_RESTO_BETA 0.05
#endif
constexpr double kRestoBeta = MPC_RESTO_BETA;   // fraction of the constraint defects the restoration step removes

// ---- workspace of one problem, in doubles -----------------------------------------------------------------
// Element i of the problem owned by lane l of a 32-problem group lives at group_base[i*LANES + l] (LANES = 32 on
// the device: every access of a warp is one coalesced 256-byte row).  Record 0 holds the scalar state, record N+1 the
// filter entries, records
// 1..N the per-stage data of time t = 0..N-1 at fixed offsets.
// The iterate block X = {S, U, LAM, ZL, ZU, TR, C} exists twice (current / trial); buffer b starts at 28*b.
enum StageOff {
  xS = 0,      // state s_t (6)
  xU = 6,      // control u_t (2)
  xLAM = 8,    // multipliers of the rows of time t (6)
  xZL = 14, xZU = 16,   // bound multipliers of u_t (2+2)
  xTR = 18,    // sin/cos(psi_t), sin/cos(epsi_t) (4)
  xC = 22,     // constraint residual of the rows of time t (6)
  xPD = 22,    // MPC_STORE_PSIDES: atan(p'(x_t)), the reference heading at s_t (shares the first residual row, unused with MPC_STORE_C 0)
  kX = 28,
  oDS = 56, oDU = 62,   // search direction (6+2)
  oCSOC = 64,  // second-order-correction right-hand side (6)
  oKF = 70,    // Riccati factors of stage t: K (2x4), Lambda^-1 (3), feed-forward k (2); MPC_FUSE_FACTOR: k = ka + mu kb, kb (2)
  kRec = 83 + 2 * MPC_FUSE_FACTOR
};
enum ScalarD {
  dDF, dMU, dTAU, dMUMIN, dDWC, dDWL, dTHMAX, dTHMIN, dF, dTH, dPINF, dDINF, dLAM1, dZ1, dSZMAX, dSZMIN, dSLOG, dXMAX,
  dDLMAX, dRTH, dRBARR, dRGBD, dALPHA, dAMAX, dAMIN, dADU, dATEST, dTRF, dTRTH, dTRPINF, dTRSLOG, dTHSOC, dASOC, dCOBJ,
  dLOBJ, dSOFTA, kNumD
};
enum ScalarI { iPHASE = kNumD, iFLAGS, iCUR, iITER, iSTATUS, iNSTEPS, iSOCCNT, iACCCNT, iNF, iSOFTCNT, iFRCNT, kNumScal };
static_assert((int)kNumScal <= (int)kRec, "scalar record must fit one stage record");

static_assert(2 * kMaxFilter <= kRec, "the filter must fit one record");
MPC_HD int workspace_doubles_per_problem(int N) { return kRec * (N + 2); }

// HBM traffic vs recomputation (the solver kernels are HBM-bound with FP64 issue slots to spare):
//   MPC_STORE_TRIG 0: sin/cos(psi_t), sin/cos(epsi_t) are recomputed by every sweep instead of stored (16 doubles/stage/iter)
//   MPC_STORE_C    0: the constraint residuals are recomputed from the iterate instead of stored (18 doubles/stage/iter)
#ifndef MPC_STORE_TRIG
#define MPC_STORE_TRIG 1
#endif
#ifndef MPC_STORE_C
#define MPC_STORE_C 0
#endif
//   MPC_STORE_PSIDES 1: the reference heading atan(p'(x_t)) of the iterate (MPC.cpp:118) is stored by whoever writes the
//                       iterate (it is evaluated there anyway) instead of being re-evaluated by the factor and forward
//                       sweeps for their residuals: 3 doubles/stage/iter more traffic for two atan() (~130 instructions)
//                       less.  Bit-identical results (the same value, computed once).  Measured: 1.3 % SLOWER
//                       (13.84 -> 13.66 M solves/s, gpurun_out/r2_psd_*.json) -- bytes weigh more than instructions -- so off.
#ifndef MPC_STORE_PSIDES
#define MPC_STORE_PSIDES 0
#endif
#if MPC_STORE_PSIDES && (!MPC_STORE_TRIG || MPC_STORE_C)
#error "MPC_STORE_PSIDES needs MPC_STORE_TRIG 1 and MPC_STORE_C 0 (it uses the first residual row, next to the trig rows)"
#endif
constexpr int kIterRows = MPC_STORE_C ? 28 : (MPC_STORE_PSIDES ? 23 : 22);   // rows of an iterate copy that hold data
// MPC_PREFETCH: 0 = off, 1 = prefetch.global.L1, 2 = prefetch.global.L2 (device only)
#ifndef MPC_PREFETCH
#define MPC_PREFETCH 1
#endif
#ifdef MPC_BOUNDS_CHECK
void mpc_bounds_check(int i);
#endif
template <int LANES>
struct Ws {
  double* b;   // base of this problem: group base + lane
  int lane;    // lane of this problem inside its group (0 on the host)
  MPC_HD double& operator()(int i) const {
#ifdef MPC_BOUNDS_CHECK   // host test build only (tests/hostsim): every workspace index is range-checked
    mpc_bounds_check(i);
#endif
    return b[(size_t)i * LANES];
  }
  // hint: rows [i, i+n) (n <= LANES) of this problem group will be read soon.  Each row of a warp is one 256-byte
  // pair of 128-byte lines; lane l touches both lines of row i+l, so two instructions cover up to 32 rows.
  MPC_HD void prefetch(int i, int n) const {
#if defined(__CUDA_ARCH__) && MPC_PREFETCH
    if (lane < n) {
      const double* row = b - lane + (size_t)(i + lane) * LANES;
#if MPC_PREFETCH == 1
      asm volatile("prefetch.global.L1 [%0];" ::"l"(row));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(row + 16));
#else
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 16));
#endif
    }
#else
    (void)i; (void)n;
#endif
  }
};

// MPC_ASYNC_STAGE (device, per-pass factor sweep): the rows of the NEXT stage are copied into shared memory by
// per-thread asynchronous copies (cp.async / LDGSTS: each thread moves the 8-byte elements of its own problem, so no
// barrier and no cross-thread hand-over is involved -- a thread waits only for its own copy group) while the current
// stage is computed, instead of the L1 prefetch + ordinary loads.  24 KB of shared memory per 64-thread block.  Measured
// with the final kernels of round 2: +0.7 % with overlapped callers, +1.4 % for a lone caller (gpurun_out/r2_misc_*.json;
// +1.2 % / -1.5 % before the chain-length work on the sweeps), so it is on.  MPC_STEP_ASYNC_STAGE: the same in the step
// sweep (32 values per stage, 32 KB per block: +1.1 %, gpurun_out/r2_sa_*.json), on; in the forward sweep it was 1.5 %
// slower with overlapped callers (29 values, 178 KB per SM at 6 blocks -- the L1 it takes away matters more there).
#ifndef MPC_ASYNC_STAGE
#define MPC_ASYNC_STAGE 1
#endif
#if defined(__CUDA_ARCH__) && MPC_ASYNC_STAGE
__device__ __forceinline__ void async_copy8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

// MPC_FAST_DIV (device): reciprocals and quotients of the sweeps without the IEEE slow path.  The compiler's double
// division is MUFU.RCP64H + Newton steps + a range check that CALLS a slow path (denormals, huge exponents): ~20
// instructions and, worse, a branch per division -- 25 of them per stage and iteration cut the stage bodies into
// basic blocks the scheduler cannot interleave.  Here: rcp.approx.ftz.f64 (20+ bits) + two Newton steps (+ one residual
// correction for a quotient), straight-line, within 1 ulp of the rounded result for operands in the normal range; for a
// zero / denormal / infinite divisor the unrefined value is returned (+-inf or 0, what the exact quotient is or rounds
// to for every use here: step-size candidates that are then not the minimum).  Host build: plain division.
// MPC_SPECIALIZE_SWEEPS: the factor / forward sweeps are instantiated per loop-invariant case (multiplier
// initialisation, second-order correction, the usual Newton system) instead of branching inside the stage body
#ifndef MPC_SPECIALIZE_SWEEPS
#define MPC_SPECIALIZE_SWEEPS 1
#endif
#ifndef MPC_STEP_ASYNC_STAGE
#define MPC_STEP_ASYNC_STAGE 1
#endif
#ifndef MPC_FAST_DIV
#define MPC_FAST_DIV 1
#endif
MPC_HD double drcp(double b) {
#if defined(__CUDA_ARCH__) && MPC_FAST_DIV
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(b));
  double e = fma(-b, r0, 1.0);
  double r = fma(r0, e, r0);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return (fabs(r0) <= DBL_MAX && r0 != 0.0) ? r : r0;
#else
  return 1.0 / b;
#endif
}
MPC_HD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__) && MPC_FAST_DIV
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(b));
  double e = fma(-b, r0, 1.0);
  double r = fma(r0, e, r0);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  double q = a * r;
  q = fma(fma(-b, q, a), r, q);
  return (fabs(r0) <= DBL_MAX && r0 != 0.0) ? q : a * r0;
#else
  return a / b;
#endif
}
MPC_HD double dmax(double a, double b) { return a > b ? a : b; }
MPC_HD double dmin(double a, double b) { return a < b ? a : b; }
MPC_HD double dclamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// value and first three derivatives of the reference polynomial (MPC.cpp:117-118 generalised; identical
// arithmetic to the degree-1 shipped form when c2=c3=0).
MPC_HD void poly_eval(const double* c, double x, double& p0, double& p1, double& p2, double& p3) {
  double a = 0.0, b = 0.0, d = 0.0, e = 0.0;
#pragma unroll
  for (int i = kMaxCoef - 1; i >= 0; --i) {
    e = e * x + 3.0 * d;
    d = d * x + 2.0 * b;
    b = b * x + a;
    a = a * x + c[i];
  }
  p0 = a; p1 = b; p2 = d; p3 = e;
}

// IpIpoptCalculatedQuantities.cpp:444-507 (slack safeguard; practically never active)
MPC_HD double safe_slack(double v, double mu, double z, double bnd) {
  if (v < DBL_EPSILON * dmin(mu, 1.0)) {
    const double smin = DBL_EPSILON * dmin(mu, 1.0);
    double t = dmax(mu / z, smin);
    double cap = dmax(fabs(bnd), 1.0) * 1.8189894035458565e-12 /* eps^0.75 */ + dmax(v, 0.0);
    v = dmin(t, cap);
  }
  return v;
}

// The four slacks of one stage (delta and a against their lower / upper bounds) at once: one practically-never-taken
// branch instead of four, so the reciprocals / divisions that follow form independent straight-line chains.
MPC_HD void safe_slack4(double& sl0, double& su0, double& sl1, double& su1, double mu, double zl0, double zu0, double zl1,
                        double zu1, const double* xl, const double* xu) {
  const double thr = DBL_EPSILON * dmin(mu, 1.0);
  if ((sl0 < thr) | (su0 < thr) | (sl1 < thr) | (su1 < thr)) {
    sl0 = safe_slack(sl0, mu, zl0, xl[0]); su0 = safe_slack(su0, mu, zu0, xu[0]);
    sl1 = safe_slack(sl1, mu, zl1, xl[1]); su1 = safe_slack(su1, mu, zu1, xu[1]);
  }
}

// Linearised dynamics of one stage: non-trivial entries of A_t (6x6) and B_t (6x2), MPC.cpp:130-137.
struct Lin {
  double a1, a2, a3, a4, a5, beta, pp, kap, sed, vce;
};
MPC_HD Lin make_lin(const Params& P, double v, double delta, double sp, double cp, double se, double ce, double p1,
                    double p2) {
  Lin L;
  const double vdt = v * P.dt;
  L.a1 = -vdt * sp;           // d x1 / d psi0
  L.a2 = vdt * cp;            // d y1 / d psi0
  L.a3 = cp * P.dt;           // d x1 / d v0
  L.a4 = sp * P.dt;           // d y1 / d v0
  L.a5 = delta * P.dtLf;      // d psi1 / d v0 = d epsi1 / d v0
  L.beta = v * P.dtLf;        // d psi1 / d delta0 = d epsi1 / d delta0
  L.pp = p1;                  // d cte1 / d x0
  L.kap = ddiv(p2, 1.0 + p1 * p1);   // -d epsi1 / d x0 (no branch around the division: 0 for a degree-1 reference)
  L.sed = se * P.dt;          // d cte1 / d v0
  L.vce = vdt * ce;           // d cte1 / d epsi0
  return L;
}
// out = r * Abar over the reduced state (x,y,psi,v,epsi); column epsi of Abar is zero, so 4 outputs.
MPC_HD void applyA(const Lin& L, double rx, double ry, double rp, double rv, double re, double& ox, double& oy,
                   double& op, double& ov) {
  const double rpe = rp + re;
  ox = rx - L.kap * re;
  oy = ry;
  op = L.a1 * rx + L.a2 * ry + rpe;
  ov = L.a3 * rx + L.a4 * ry + L.a5 * rpe + rv;
}
// A_t^T m for the full 6-vector m=(x,y,psi,v,cte,epsi); component cte of the result is 0.
MPC_HD void applyAT6(const Lin& L, const double* m, double* o) {
  o[0] = m[0] + L.pp * m[4] - L.kap * m[5];
  o[1] = m[1] - m[4];
  o[2] = L.a1 * m[0] + L.a2 * m[1] + m[2] + m[5];
  o[3] = L.a3 * m[0] + L.a4 * m[1] + L.a5 * (m[2] + m[5]) + m[3] + L.sed * m[4];
  o[4] = 0.0;
  o[5] = L.vce * m[4];
}

// Lagrangian-Hessian entries contributed by the constraint rows of time t+1 to the variables of time t
// (SURVEY Appendix A), multipliers lam = lambda_{t+1}.
struct Hes {
  double xx, pp, vp, ee, ev, m;
};
MPC_HD Hes make_hes(const Params& P, const double* lam, double v, double sp, double cp, double se, double ce, double p1,
                    double p2, double p3) {
  Hes H;
  {   // 0 for a degree-1 reference (p2 = p3 = 0); evaluated without a branch
    const double q = 1.0 + p1 * p1;
    H.xx = -lam[4] * p2 + ddiv(lam[5] * (p3 * q - 2.0 * p1 * p2 * p2), q * q);
  }
  const double vdt = v * P.dt;
  H.pp = (lam[0] * cp + lam[1] * sp) * vdt;
  H.vp = (lam[0] * sp - lam[1] * cp) * P.dt;
  H.ee = lam[4] * vdt * se;
  H.ev = -lam[4] * ce * P.dt;
  H.m = -(lam[2] + lam[5]) * P.dtLf;
  return H;
}

#define PS(i, j) pss[((i) >= (j)) ? ((i) * ((i) + 1) / 2 + (j)) : ((j) * ((j) + 1) / 2 + (i))]

// Result of one MPC::Solve.
struct Result {
  int status, iters;
  double obj;
  double out8[8];
};

// PH_RESTO: the line search failed; the problem waits for Solver::do_resto (run by the finisher kernels, never by the
// per-pass sweeps, which skip any phase that is not theirs)
enum Phase { PH_FACTOR = 0, PH_FORWARD = 1, PH_STEP = 2, PH_DONE = 3, PH_RESTO = 4 };
// F_SOFT / F_SOFTTRY / F_HARD: soft restoration phase (Params::resto == 2), see Solver::resto_entry
enum Flags { F_INSOC = 1, F_SOCDONE = 2, F_LS = 4, F_LSKEEP = 8, F_TINYLAST = 16, F_TINYFLAG = 32, F_TINYNOW = 64, F_RESTO = 128,
             F_SOFT = 256, F_SOFTTRY = 512, F_HARD = 1024, F_FILTDONE = 2048, F_SOFTFIX = 4096,
             F_LASTREJF = 8192 /* the last rejected trial point was rejected by the filter */ };

// ---- batch compaction ----------------------------------------------------------------------------------------
// Moves one unfinished problem from workspace slot `s` to slot `d` (another workspace region) between two passes.
// A problem that waits for a new factorisation owns nothing but its scalar record and the current copy of the
// iterate block (everything else is rewritten by the passes before it is read); in any other state the whole
// workspace of the problem is moved.
template <int LS_, int LD_>
MPC_HD void repack_rows(const Ws<LS_>& s, const Ws<LD_>& d, int r, int n) {
  int i = 0;
  for (; i + 16 <= n; i += 16) {   // loads first, then stores: 16 rows in flight per lane
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = s(r + i + k);
#pragma unroll
    for (int k = 0; k < 16; ++k) d(r + i + k) = v[k];
  }
  for (; i < n; ++i) d(r + i) = s(r + i);
}
template <int LS_, int LD_>
MPC_HD void repack_problem(const Params& P, const Ws<LS_>& s, const Ws<LD_>& d) {
  repack_rows(s, d, 0, (int)kNumScal);
  repack_rows(s, d, (P.N + 1) * kRec, 2 * (int)s(iNF));   // the filter entries in use
  const int phase = (int)s(iPHASE), flags = (int)s(iFLAGS), cur = (int)s(iCUR);
  if (phase == PH_FACTOR && !(flags & F_INSOC)) {
    constexpr int n = kIterRows;
    const int lo = kX * cur;
    int t = 0;
    for (; t + 2 <= P.N; t += 2) {   // two stages (44 rows) in flight per lane
      const int r0 = (t + 1) * kRec + lo, r1 = r0 + kRec;
      double v[2 * n];
#pragma unroll
      for (int k = 0; k < n; ++k) { v[k] = s(r0 + k); v[n + k] = s(r1 + k); }
#pragma unroll
      for (int k = 0; k < n; ++k) { d(r0 + k) = v[k]; d(r1 + k) = v[n + k]; }
    }
    if (t < P.N) repack_rows(s, d, (t + 1) * kRec + lo, n);
  } else if (phase == PH_FORWARD && !(flags & F_INSOC)) {
    // waits for its forward sweep (the usual state after kernel_stepfactor): the current iterate and the Riccati factors
    const int lo = kX * cur;
    for (int t = 0; t < P.N; ++t) {   // one stage (iterate + factors) in flight per lane
      const int r0 = (t + 1) * kRec + lo, r1 = (t + 1) * kRec + oKF;
      double v[kIterRows + kKF];
#pragma unroll
      for (int k = 0; k < kIterRows; ++k) v[k] = s(r0 + k);
#pragma unroll
      for (int k = 0; k < kKF; ++k) v[kIterRows + k] = s(r1 + k);
#pragma unroll
      for (int k = 0; k < kIterRows; ++k) d(r0 + k) = v[k];
#pragma unroll
      for (int k = 0; k < kKF; ++k) d(r1 + k) = v[kIterRows + k];
    }
  } else {
    for (int t = 0; t < P.N; ++t) repack_rows(s, d, (t + 1) * kRec, (int)kRec);
  }
}

// The per-problem solver.  All "passes" are loops over the horizon that touch the workspace once per stage.
template <int LANES>
struct Solver {
  const Params& P;
  Ws<LANES> w;
  const int N, M;
  double cf[kMaxCoef];
  // ---- scalar algorithm state (persisted in record 0 of the workspace between passes)
  double df, mu, tau, mu_min;
  double dw_curr, dw_last;          // PDPerturbationHandler delta_x
  double theta_max, theta_min;
  double f_cur, theta_cur, priminf, dualinf, lam1, z1, sz_max, sz_min, sumlog, xmaxabs, dlam_max;
  double ref_theta, ref_barr, ref_gbd, alpha, alpha_max, alpha_min, alpha_du, alpha_test;
  double tr_f, tr_theta, tr_priminf, tr_sumlog;   // at the last evaluated trial point
  double theta_soc_old, alpha_soc;
  double curr_obj, last_obj, soft_alpha;
  int phase, flags, cur, iter, status, n_steps, soc_count, acceptable_counter, nf, soft_count, filt_rej_count;
  // transient (within a pass)
  double fw_alpha_pr, fw_alpha_du, fw_gbd;
  bool fw_tiny;
  // carry buffer of the STEP sweep (kCarry doubles, element i at cr[i*cs]): shared memory in the per-pass kernel,
  // a thread-local array elsewhere.  Must be set before step_sweep() runs.
  double* cr;
  int cs;
  // carry buffer of the Riccati recursion that rides on the STEP sweep (kRicCarry doubles, element i at rq[i*rqs]); set
  // before step_sweep<true>() runs
  double* rq;
  int rqs;
  // stage buffers of the asynchronous staging (MPC_ASYNC_STAGE; element i of buffer b at sb[(b * kStageVals + i) * sbs]),
  // null = ordinary loads
  double* sb;
  int sbs;

  MPC_HD Solver(const Params& p, double* base, int lane = 0) : P(p), w{base, lane}, N(p.N), M(p.N - 1), cr(nullptr), cs(1), rq(nullptr), rqs(1), sb(nullptr), sbs(1) {}

  MPC_HD bool fl(int f) const { return (flags & f) != 0; }
  MPC_HD void setfl(int f, bool v) { flags = v ? (flags | f) : (flags & ~f); }
  // record of time t starts at (t+1)*kRec
  MPC_HD int rec(int t) const { return (t + 1) * kRec; }
  MPC_HD double& filt(int i) const { return w(rec(N) + i); }   // record N+1, after the stage records

  MPC_HD void set_coeffs(const double* coef, int ncoef) {
#pragma unroll
    for (int i = 0; i < kMaxCoef; ++i) cf[i] = i < ncoef ? coef[i] : 0.0;
  }

  // ---- scalar state <-> workspace record 0
#define MPC_SCALARS(X)                                                                                                  \
  X(dDF, df) X(dMU, mu) X(dTAU, tau) X(dMUMIN, mu_min) X(dDWC, dw_curr) X(dDWL, dw_last) X(dTHMAX, theta_max)            \
  X(dTHMIN, theta_min) X(dF, f_cur) X(dTH, theta_cur) X(dPINF, priminf) X(dDINF, dualinf) X(dLAM1, lam1) X(dZ1, z1)      \
  X(dSZMAX, sz_max) X(dSZMIN, sz_min) X(dSLOG, sumlog) X(dXMAX, xmaxabs) X(dDLMAX, dlam_max) X(dRTH, ref_theta)          \
  X(dRBARR, ref_barr) X(dRGBD, ref_gbd) X(dALPHA, alpha) X(dAMAX, alpha_max) X(dAMIN, alpha_min) X(dADU, alpha_du)       \
  X(dATEST, alpha_test) X(dTRF, tr_f) X(dTRTH, tr_theta) X(dTRPINF, tr_priminf) X(dTRSLOG, tr_sumlog)                   \
  X(dTHSOC, theta_soc_old) X(dASOC, alpha_soc) X(dCOBJ, curr_obj) X(dLOBJ, last_obj) X(dSOFTA, soft_alpha)
#define MPC_SCALARS_I(X)                                                                                                \
  X(iPHASE, phase) X(iFLAGS, flags) X(iCUR, cur) X(iITER, iter) X(iSTATUS, status) X(iNSTEPS, n_steps)                   \
  X(iSOCCNT, soc_count) X(iACCCNT, acceptable_counter) X(iNF, nf) X(iSOFTCNT, soft_count) X(iFRCNT, filt_rej_count)
  MPC_HD void load_state() {
#define X(idx, name) name = w(idx);
    MPC_SCALARS(X)
#undef X
#define X(idx, name) name = (int)w(idx);
    MPC_SCALARS_I(X)
#undef X
  }
  MPC_HD void store_state() {
#define X(idx, name) w(idx) = name;
    MPC_SCALARS(X)
#undef X
#define X(idx, name) w(idx) = (double)name;
    MPC_SCALARS_I(X)
#undef X
  }
  MPC_HD int load_phase() const { return (int)w(iPHASE); }

  // objective gradient w.r.t. u_t (scaled by df); um/up = u_{t-1}/u_{t+1}
  MPC_HD double grad_u(int j, int t, double u, double um, double up) const {
    const double wq = j == 0 ? P.w_delta : P.w_a, wd = j == 0 ? P.w_ddelta : P.w_da;
    double g = wq * u;
    if (t > 0) g += wd * (u - um);
    if (t < M - 1) g -= wd * (up - u);
    return (2.0 * df) * g;
  }
  MPC_HD double hess_u(int j, int t) const {
    const double wq = j == 0 ? P.w_delta : P.w_a, wd = j == 0 ? P.w_ddelta : P.w_da;
    const double nd = (t > 0 ? 1.0 : 0.0) + (t < M - 1 ? 1.0 : 0.0);
    return 2.0 * df * (wq + nd * wd);
  }
  // constraint residual of the rows of time t+1 (MPC.cpp:130-137): sn - F(s, u)
  MPC_HD void residual(const double* s, const double* u, const double* sn, double sp, double cp, double se, double p0,
                       double psides, double* c) const {
    const double vd = s[3] * u[0] * P.dtLf;
    c[0] = sn[0] - (s[0] + s[3] * cp * P.dt);
    c[1] = sn[1] - (s[1] + s[3] * sp * P.dt);
    c[2] = sn[2] - (s[2] + vd);
    c[3] = sn[3] - (s[3] + u[1] * P.dt);
    c[4] = sn[4] - ((p0 - s[1]) + (s[3] * se * P.dt));
    c[5] = sn[5] - ((s[2] - psides) + vd);
  }
  // sin/cos of psi_t and epsi_t of the iterate copy starting at row r (stored or recomputed)
  MPC_HD void trig_of(int r, const double* s, double& sp, double& cp, double& se, double& ce) const {
#if MPC_STORE_TRIG
    (void)s;
    sp = w(r + xTR); cp = w(r + xTR + 1); se = w(r + xTR + 2); ce = w(r + xTR + 3);
#else
    (void)r;
    sincos(s[2], &sp, &cp);
    sincos(s[5], &se, &ce);
#endif
  }
  // reference heading at s_t of the iterate copy starting at row r (stored or recomputed from p'(x_t))
  MPC_HD double psides_of(int r, double p1) const {
#if MPC_STORE_PSIDES
    (void)p1;
    return w(r + xPD);
#else
    (void)r;
    return atan(p1);
#endif
  }
  MPC_HD void store_psides(int r, double psides) const {
#if MPC_STORE_PSIDES
    w(r + xPD) = psides;
#else
    (void)r; (void)psides;
#endif
  }
  // constraint residual of the rows of time t+1 at the iterate copy `buf` (rare paths only)
  MPC_HD void residual_at(int buf, int t, double* c) const {
#if MPC_STORE_C
    for (int k = 0; k < 6; ++k) c[k] = w(rec(t + 1) + kX * buf + xC + k);
#else
    const int r = rec(t) + kX * buf;
    double s[6], sn[6], u[2], sp, cp, se, ce, p0, p1, p2, p3;
    for (int k = 0; k < 6; ++k) { s[k] = w(r + xS + k); sn[k] = w(r + kRec + xS + k); }
    u[0] = w(r + xU); u[1] = w(r + xU + 1);
    trig_of(r, s, sp, cp, se, ce);
    poly_eval(cf, s[0], p0, p1, p2, p3);
    residual(s, u, sn, sp, cp, se, p0, atan(p1), c);
#endif
  }
  MPC_HD double state_cost(const double* s) const {
    return P.w_cte * (s[4] * s[4]) + P.w_epsi * (s[5] * s[5]) + P.w_v * ((s[3] - P.ref_v) * (s[3] - P.ref_v));
  }

  // ------------------------------------------------------------------------------------------
  // start point, bounds, scaling, z, mu  (MPC.cpp:167-203; IpGradientScaling.cpp:99-116;
  // IpDefaultIterateInitializer.cpp:230-266, 469-649), residuals / objective / barrier at the start point
  MPC_HD void init_scalars(const double* s0, const double* coef, int ncoef) {
    set_coeffs(coef, ncoef);
    double gmax = dmax(fabs(2.0 * P.w_cte * s0[4]), fabs(2.0 * P.w_epsi * s0[5]));
    gmax = dmax(gmax, fabs(2.0 * P.w_v * (s0[3] - P.ref_v)));
    if (N > 1) gmax = dmax(gmax, fabs(2.0 * P.w_v * (0.0 - P.ref_v)));
    df = 1.0;
    if (gmax > 100.0) df = 100.0 / gmax;
    if (df < 1e-8) df = 1e-8;
    mu = 0.1;
    tau = dmax(0.99, 1.0 - mu);
    mu_min = dmin(P.tol, 1e-4 * df) / (10.0 + 1.0);
    theta_max = theta_min = -1.0;
    nf = 0;
    dw_curr = dw_last = 0.0;
    cur = 0; iter = 0; status = -100;
    acceptable_counter = 0;
    curr_obj = last_obj = -1e50;
    flags = F_LS;
    dlam_max = 0.0;
    alpha = alpha_du = alpha_max = alpha_min = alpha_test = 0.0;
    ref_theta = ref_barr = ref_gbd = 0.0;
    theta_soc_old = alpha_soc = 0.0;
    n_steps = soc_count = 0;
    soft_count = 0; soft_alpha = 0.0; filt_rej_count = 0;
    dualinf = lam1 = z1 = sz_max = sz_min = xmaxabs = 0.0;
    tr_f = tr_theta = tr_priminf = tr_sumlog = 0.0;
  }
  MPC_HD void init(const double* s0, const double* coef, int ncoef) {
    init_scalars(s0, coef, ncoef);
    double f = 0.0, th = 0.0, cm = 0.0, sl = 0.0;
    for (int t = 0; t < N; ++t) {
      double ft, tht, cmt, slt;
      init_stage(t, s0, ft, tht, cmt, slt);
      f += ft; th += tht; cm = dmax(cm, cmt); sl += slt;
    }
    init_finish(f, th, cm, sl);
  }
  // stage t of the start point: states are zero except t = 0, controls at their (pushed) start value, z = 1,
  // lambda = 0; partial objective / constraint violation / log-barrier of the stage
  MPC_HD void init_stage(int t, const double* s0, double& f, double& th, double& cm, double& sl) {
    double s[6], sn[6], u[2];
    u[0] = P.u_init[0]; u[1] = P.u_init[1];
    f = 0.0; th = 0.0; cm = 0.0; sl = 0.0;
    {
      const int r = rec(t);
#pragma unroll
      for (int k = 0; k < 6; ++k) { s[k] = t == 0 ? s0[k] : 0.0; sn[k] = 0.0; w(r + xS + k) = s[k]; w(r + xLAM + k) = 0.0; }
      f += state_cost(s);
      if (t < M) {
#pragma unroll
        for (int j = 0; j < 2; ++j) { w(r + xU + j) = u[j]; w(r + xZL + j) = 1.0; w(r + xZU + j) = 1.0; }
        double sp, cp, se, ce, p0, p1, p2, p3, c[6];
        sincos(s[2], &sp, &cp);
        sincos(s[5], &se, &ce);
        poly_eval(cf, s[0], p0, p1, p2, p3);
        const double psides = atan(p1);
        residual(s, u, sn, sp, cp, se, p0, psides, c);
#if MPC_STORE_TRIG
        w(r + xTR) = sp; w(r + xTR + 1) = cp; w(r + xTR + 2) = se; w(r + xTR + 3) = ce;
#endif
        store_psides(r, psides);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#if MPC_STORE_C
          w(r + kRec + xC + k) = c[k];
#endif
          th += fabs(c[k]); cm = dmax(cm, fabs(c[k]));
        }
        f += P.w_delta * (u[0] * u[0]) + P.w_a * (u[1] * u[1]);
        sl = log((u[0] - P.xl[0]) * (P.xu[0] - u[0]) * ((u[1] - P.xl[1]) * (P.xu[1] - u[1])));
      }
    }
  }
  MPC_HD void init_finish(double f, double th, double cm, double sl) {
    f_cur = df * f; theta_cur = th; priminf = cm; sumlog = sl;
    phase = PH_FACTOR;
    // NaN / Inf in the inputs: Ipopt stops at the starting point with Invalid_Number_Detected (IpIpoptAlg.cpp:397-470
    // maps the evaluation error; no iteration is counted)
    if (!(fabs(f_cur) <= DBL_MAX) || !(th <= DBL_MAX)) { status = kInvalidNumberDetected; phase = PH_DONE; }
  }

  // ------------------------------------------------------------------------------------------
  // Warm start of closed-loop step k+1 from the solution of step k (NOT reference behaviour -- the reference
  // cold-starts every call, MPC.cpp:167-177 -- so it is off by default).  The previous solution, which sits in the current
  // copy of the iterate, is shifted by one stage (the new initial state is its t=1 state, main.cpp:66), the last stage is
  // rolled out with the last control, controls are pushed 1e-3 inside their bounds and multipliers kept >= 1e-3 (Ipopt's
  // warm_start_bound_push / warm_start_mult_bound_push), mu starts at mu0.  The first pass over the shifted point is a
  // zero-length STEP (alpha = 0, accepted unconditionally) that evaluates every norm top_of_loop() needs.
  MPC_HD void init_warm(const double* s0, double mu0) {
    const int b = kX * (int)w(iCUR);
    const double df_old = w(dDF);
    // shift in place (ascending t)
    for (int t = 0; t < M; ++t) {
      const int r = rec(t) + b, rn = rec(t + 1) + b;
#pragma unroll
      for (int k = 0; k < 6; ++k) { w(r + xS + k) = t == 0 ? s0[k] : w(rn + xS + k); w(r + xLAM + k) = w(rn + xLAM + k); }
      if (t < M - 1) {
#pragma unroll
        for (int j = 0; j < 2; ++j) { w(r + xU + j) = w(rn + xU + j); w(r + xZL + j) = w(rn + xZL + j); w(r + xZU + j) = w(rn + xZU + j); }
      }
    }
    // objective scaling at the (new) start point, pushes, last stage, residuals / objective / barrier / trig
    double gmax = 0.0, f = 0.0, th = 0.0, cm = 0.0, sl = 0.0;
    double s[6], sn[6], u[2], up[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] = w(rec(0) + b + xS + k);
    for (int t = 0; t < M; ++t) {
      const int r = rec(t) + b, rn = rec(t + 1) + b;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double push = 1e-3 * dmax(1.0, fabs(P.xl[j]));
        u[j] = dclamp(w(r + xU + j), P.xl[j] + push, P.xu[j] - push);
        w(r + xU + j) = u[j];
      }
      double sp, cp, se, ce, p0, p1, p2, p3, c[6];
      sincos(s[2], &sp, &cp);
      sincos(s[5], &se, &ce);
      poly_eval(cf, s[0], p0, p1, p2, p3);
      const double psides = atan(p1);
      if (t == M - 1) {   // last stage: roll the model out so the new rows start feasible
        const double zero[6] = {0, 0, 0, 0, 0, 0};
        residual(s, u, zero, sp, cp, se, p0, psides, c);
#pragma unroll
        for (int k = 0; k < 6; ++k) w(rn + xS + k) = -c[k];
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) sn[k] = w(rn + xS + k);
      residual(s, u, sn, sp, cp, se, p0, psides, c);
#if MPC_STORE_TRIG
      w(r + xTR) = sp; w(r + xTR + 1) = cp; w(r + xTR + 2) = se; w(r + xTR + 3) = ce;
#endif
      store_psides(r, psides);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
#if MPC_STORE_C
        w(rn + xC + k) = c[k];
#endif
        th += fabs(c[k]); cm = dmax(cm, fabs(c[k]));
      }
      f += state_cost(s) + P.w_delta * (u[0] * u[0]) + P.w_a * (u[1] * u[1]);
      if (t > 0) f += P.w_ddelta * ((u[0] - up[0]) * (u[0] - up[0])) + P.w_da * ((u[1] - up[1]) * (u[1] - up[1]));
      sl += log((u[0] - P.xl[0]) * (P.xu[0] - u[0]) * ((u[1] - P.xl[1]) * (P.xu[1] - u[1])));
      gmax = dmax(gmax, dmax(fabs(2.0 * P.w_cte * s[4]), dmax(fabs(2.0 * P.w_epsi * s[5]), fabs(2.0 * P.w_v * (s[3] - P.ref_v)))));
      // control gradient entries are bounded by 2 (w + 2 w_d) |u|max; include them conservatively
      gmax = dmax(gmax, 2.0 * (P.w_delta + 2.0 * P.w_ddelta) * fabs(u[0]));
      gmax = dmax(gmax, 2.0 * (P.w_a + 2.0 * P.w_da) * fabs(u[1]));
      up[0] = u[0]; up[1] = u[1];
#pragma unroll
      for (int k = 0; k < 6; ++k) { s[k] = sn[k]; w(rec(t) + oDS + k) = 0.0; }
      w(rec(t) + oDU) = 0.0; w(rec(t) + oDU + 1) = 0.0;
    }
    f += state_cost(s);
    gmax = dmax(gmax, dmax(fabs(2.0 * P.w_cte * s[4]), dmax(fabs(2.0 * P.w_epsi * s[5]), fabs(2.0 * P.w_v * (s[3] - P.ref_v)))));
#pragma unroll
    for (int k = 0; k < 6; ++k) w(rec(M) + oDS + k) = 0.0;
    const int cur_keep = (int)w(iCUR);
    init_scalars(s0, cf, kMaxCoef);
    cur = cur_keep;
    df = 1.0;
    if (gmax > 100.0) df = 100.0 / gmax;
    if (df < 1e-8) df = 1e-8;
    mu_min = dmin(P.tol, 1e-4 * df) / (10.0 + 1.0);
    mu = dmax(mu0, mu_min);
    tau = dmax(0.99, 1.0 - mu);
    // multipliers of the scaled problem follow the objective scaling; keep the bound multipliers away from zero
    const double ratio = df / df_old;
    for (int t = 0; t < N; ++t) {
      const int r = rec(t) + b;
#pragma unroll
      for (int k = 0; k < 6; ++k) w(r + xLAM + k) = ratio * w(r + xLAM + k);
      if (t < M) {
#pragma unroll
        for (int j = 0; j < 2; ++j) { w(r + xZL + j) = dmax(1e-3, ratio * w(r + xZL + j)); w(r + xZU + j) = dmax(1e-3, ratio * w(r + xZU + j)); }
      }
    }
    f_cur = df * f; theta_cur = th; priminf = cm; sumlog = sl;
    flags = F_TINYNOW;   // zero-length step, accepted without a filter test
    alpha = 0.0; alpha_du = 0.0;
    iter = -1;           // the zero-length step is not an interior-point iteration
    phase = PH_STEP;
  }

#if defined(__CUDA_ARCH__) && MPC_ASYNC_STAGE
  // asynchronous copy of the 22 iterate rows of stage t (copy bX) and of u_{t-1} into stage buffer t & 1
  __device__ __forceinline__ void stage_request(int t, int bX) {
    double* q = sb + (size_t)((t & 1) * kStageVals) * sbs;
    const double* g = &w(rec(t) + bX);
#pragma unroll
    for (int k = 0; k < 22; ++k) async_copy8(q + k * sbs, g + (size_t)k * LANES);
    if (t > 0) {
      const double* gm = &w(rec(t - 1) + bX + xU);
      async_copy8(q + 22 * sbs, gm);
      async_copy8(q + 23 * sbs, gm + LANES);
    }
    async_commit();
  }
  // step sweep: what stage t reads, into stage buffer t & 1 (element i of buffer b at sb[(b * kStepStageVals + i) * sbs])
  __device__ __forceinline__ void step_stage_request(int t, int bO) {
    double* q = sb + (size_t)((t & 1) * kStepStageVals) * sbs;
    const double* g = &w(rec(t) + bO);
#pragma unroll
    for (int k = 0; k < 22; ++k) async_copy8(q + k * sbs, g + (size_t)k * LANES);
    const double* gd = &w(rec(t) + oDS);
#pragma unroll
    for (int k = 0; k < 6; ++k) async_copy8(q + (22 + k) * sbs, gd + (size_t)k * LANES);
    if (t > 0) {
      const double* gu = &w(rec(t - 1) + bO + xU);
      async_copy8(q + 28 * sbs, gu);
      async_copy8(q + 29 * sbs, gu + LANES);
      const double* gdu = &w(rec(t - 1) + oDU);
      async_copy8(q + 30 * sbs, gdu);
      async_copy8(q + 31 * sbs, gdu + LANES);
    }
    async_commit();
  }
#endif
  // ------------------------------------------------------------------------------------------
  // Backward Riccati sweep: factor + feed-forward for the right-hand side (grad L_mu, c).  use_csoc selects the
  // constraint right-hand side (second-order correction).  Returns false on wrong inertia.
  MPC_HD bool factor(double dw, bool use_csoc) { return factor_body(dw, fl(F_LS), use_csoc); }
  // The multiplier-initialisation system (F_LS) and the second-order-correction right-hand side are loop invariants of
  // the sweep: the per-pass kernels run one instantiation per case, which keeps those branches out of the stage body
  // (larger basic blocks to schedule); the looping kernels keep the one generic body (instruction footprint).
  MPC_HD bool factor_spec(double dw, bool use_csoc) {
#if MPC_SPECIALIZE_SWEEPS
    if (fl(F_LS)) return factor_t<true, false>(dw);
    if (use_csoc) return factor_t<false, true>(dw);
    return factor_t<false, false>(dw);
#else
    return factor_body(dw, fl(F_LS), use_csoc);
#endif
  }
  template <bool LS, bool CSOC>
  MPC_HD bool factor_t(double dw) { return factor_body(dw, LS, CSOC); }
  MPC_HD bool factor_body(double dw, const bool ls, const bool use_csoc) {
    const double qv = ls ? 1.0 : 2.0 * P.w_v * df + dw, qe = ls ? 1.0 : 2.0 * P.w_epsi * df + dw,
                 qc = ls ? 1.0 : 2.0 * P.w_cte * df + dw, q0 = ls ? 1.0 : dw;
    const double gv2 = 2.0 * P.w_v * df, ge2 = 2.0 * P.w_epsi * df, gc2 = 2.0 * P.w_cte * df;
    const int bX = kX * cur, bC = use_csoc ? (int)oCSOC : bX + xC;
    double pss[15], psu[4][2], puu00 = 0.0, puu10 = 0.0, puu11 = 0.0, pv[5], pu0 = 0.0, pu1 = 0.0;
#pragma unroll
    for (int i = 0; i < 15; ++i) pss[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) psu[i][0] = psu[i][1] = 0.0;
    double lamn[6];   // lambda_{t+1}
    double rc_next;
    double sT[6];     // s_{t+1}
    {
      const int r = rec(M) + bX;
#pragma unroll
      for (int k = 0; k < 6; ++k) { lamn[k] = ls ? 0.0 : w(r + xLAM + k); sT[k] = w(r + xS + k); }
      PS(0, 0) = q0; PS(1, 1) = q0; PS(2, 2) = q0; PS(3, 3) = qv; PS(4, 4) = qe;
      // r_s at the terminal time: grad f + lambda
      pv[0] = lamn[0]; pv[1] = lamn[1]; pv[2] = lamn[2];
      pv[3] = gv2 * (sT[3] - P.ref_v) + lamn[3];
      pv[4] = ge2 * sT[5] + lamn[5];
      rc_next = gc2 * sT[4] + lamn[4];   // grad L wrt cte_{t+1}
    }
    double un0 = 0.0, un1 = 0.0;         // u_{t+1}
    bool ok = true;
#if defined(__CUDA_ARCH__) && MPC_ASYNC_STAGE && MPC_STORE_TRIG
    const bool staged = sb != nullptr;
    if (staged) stage_request(M - 1, bX);
#else
    const bool staged = false;
#endif
    for (int t = M - 1; t >= 0; --t) {
      const int r = rec(t) + bX, rn = rec(t + 1);
      if (t > 0 && !staged) {   // rows of stage t-1: S,U,LAM,ZL,ZU(,TR) are contiguous, then the residual of rows t
        w.prefetch(r - kRec + xS, MPC_STORE_TRIG ? (MPC_STORE_PSIDES ? 23 : 22) : 18);
        if (!ls && (MPC_STORE_C || use_csoc)) w.prefetch(rec(t) + bC, 6);
      }
      double s[6], lam[6];
      double u0, u1, um0 = 0.0, um1 = 0.0, sp, cp, se, ce, zl0, zl1, zu0, zu1;
#if defined(__CUDA_ARCH__) && MPC_ASYNC_STAGE && MPC_STORE_TRIG
      if (sb) {
        // this stage's rows were requested one stage ago (the prologue for t = M-1); request the next stage's, then use
        async_wait_all();
        const double* q = sb + (size_t)((t & 1) * kStageVals) * sbs;
#pragma unroll
        for (int k = 0; k < 6; ++k) { s[k] = q[(xS + k) * sbs]; lam[k] = ls ? 0.0 : q[(xLAM + k) * sbs]; }
        u0 = q[xU * sbs]; u1 = q[(xU + 1) * sbs];
        zl0 = q[xZL * sbs]; zl1 = q[(xZL + 1) * sbs]; zu0 = q[xZU * sbs]; zu1 = q[(xZU + 1) * sbs];
        sp = q[xTR * sbs]; cp = q[(xTR + 1) * sbs]; se = q[(xTR + 2) * sbs]; ce = q[(xTR + 3) * sbs];
        if (t > 0) { um0 = q[22 * sbs]; um1 = q[23 * sbs]; }
        if (t > 0) stage_request(t - 1, bX);
      } else
#endif
      {
#pragma unroll
        for (int k = 0; k < 6; ++k) { s[k] = w(r + xS + k); lam[k] = ls ? 0.0 : w(r + xLAM + k); }
        u0 = w(r + xU); u1 = w(r + xU + 1);
        if (t > 0) { um0 = w(r - kRec + xU); um1 = w(r - kRec + xU + 1); }
        trig_of(r, s, sp, cp, se, ce);
        zl0 = w(r + xZL); zl1 = w(r + xZL + 1); zu0 = w(r + xZU); zu1 = w(r + xZU + 1);
      }
      double p0, p1, p2, p3;
      poly_eval(cf, s[0], p0, p1, p2, p3);
      // constraint right-hand side of rows t+1
      double rb[5], cc;
      if (ls) { rb[0] = rb[1] = rb[2] = rb[3] = rb[4] = 0.0; cc = 0.0; }
      else if (MPC_STORE_C || use_csoc) {
        rb[0] = -w(rn + bC); rb[1] = -w(rn + bC + 1); rb[2] = -w(rn + bC + 2); rb[3] = -w(rn + bC + 3);
        cc = w(rn + bC + 4); rb[4] = -w(rn + bC + 5);
      } else {
        double c[6];
        const double uu[2] = {u0, u1};
        residual(s, uu, sT, sp, cp, se, p0, staged ? atan(p1) : psides_of(r, p1), c);
        rb[0] = -c[0]; rb[1] = -c[1]; rb[2] = -c[2]; rb[3] = -c[3]; cc = c[4]; rb[4] = -c[5];
      }
      const Lin A = make_lin(P, s[3], u0, sp, cp, se, ce, p1, p2);
      Hes H;
      if (ls) { H.xx = H.pp = H.vp = H.ee = H.ev = H.m = 0.0; }
      else H = make_hes(P, lamn, s[3], sp, cp, se, ce, p1, p2, p3);
      double r0, r1, ru0, ru1;
      double ATl[6];
      applyAT6(A, lamn, ATl);   // A_t^T lambda_{t+1}
      const double gu0 = grad_u(0, t, u0, um0, un0), gu1 = grad_u(1, t, u1, um1, un1);
      if (ls) {
        r0 = 1.0; r1 = 1.0;
        ru0 = gu0 - zl0 + zu0; ru1 = gu1 - zl1 + zu1;
      } else {
        const double bl0 = A.beta * (lamn[2] + lamn[5]), bl1 = P.dt * lamn[3];   // B_t^T lambda_{t+1}
        double sl0 = u0 - P.xl[0], su0 = P.xu[0] - u0, sl1 = u1 - P.xl[1], su1 = P.xu[1] - u1;
        safe_slack4(sl0, su0, sl1, su1, mu, zl0, zu0, zl1, zu1, P.xl, P.xu);
        const double isl0 = drcp(sl0), isu0 = drcp(su0), isl1 = drcp(sl1), isu1 = drcp(su1);
        r0 = hess_u(0, t) + dw + zl0 * isl0 + zu0 * isu0;
        r1 = hess_u(1, t) + dw + zl1 * isl1 + zu1 * isu1;
        ru0 = gu0 - bl0 - mu * isl0 + mu * isu0;
        ru1 = gu1 - bl1 - mu * isl1 + mu * isu1;
      }
      const double gc = rc_next - qc * cc;   // linear coefficient on a_c^T dsigma (cte fold)
      // ---- recursion
      double wv[5], wu0 = pu0, wu1 = pu1;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        double a = pv[i];
#pragma unroll
        for (int j = 0; j < 5; ++j) a += PS(i, j) * rb[j];
        wv[i] = a;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { wu0 += psu[i][0] * rb[i]; wu1 += psu[i][1] * rb[i]; }
      double T0[5], T1[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        T0[j] = A.beta * (PS(2, j) + PS(4, j)) + (j < 4 ? psu[j < 4 ? j : 0][0] : 0.0);
        T1[j] = P.dt * PS(3, j) + (j < 4 ? psu[j < 4 ? j : 0][1] : 0.0);
      }
      const double Tu00 = A.beta * psu[2][0] + puu00, Tu10 = P.dt * psu[3][0] + puu10, Tu11 = P.dt * psu[3][1] + puu11;
      const double L00 = r0 + A.beta * (T0[2] + T0[4]) + Tu00;
      const double L10 = A.beta * (T1[2] + T1[4]) + Tu10;
      const double L11 = r1 + P.dt * T1[3] + Tu11;
      const double det = L00 * L11 - L10 * L10;
      if (!(L00 > 0.0) || !(det > 0.0)) ok = false;
      const double idet = drcp(det);
      const double i00 = L11 * idet, i10 = -L10 * idet, i11 = L00 * idet;
      double G0[4], G1[4];
      applyA(A, T0[0], T0[1], T0[2], T0[3], T0[4], G0[0], G0[1], G0[2], G0[3]);
      applyA(A, T1[0], T1[1], T1[2], T1[3], T1[4], G1[0], G1[1], G1[2], G1[3]);
      G0[3] += H.m;
      const double h0 = ru0 + A.beta * (wv[2] + wv[4]) + wu0, h1 = ru1 + P.dt * wv[3] + wu1;
      double K0[4], K1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { K0[j] = i00 * G0[j] + i10 * G1[j]; K1[j] = i10 * G0[j] + i11 * G1[j]; }
      const double k0 = i00 * h0 + i10 * h1, k1 = i10 * h0 + i11 * h1;
      {
        const int ko = rec(t) + oKF;
#pragma unroll
        for (int j = 0; j < 4; ++j) { w(ko + j) = K0[j]; w(ko + 4 + j) = K1[j]; }
        w(ko + 8) = i00; w(ko + 9) = i10; w(ko + 10) = i11; w(ko + 11) = k0; w(ko + 12) = k1;
#if MPC_FUSE_FACTOR
        w(ko + 13) = 0.0; w(ko + 14) = 0.0;   // the barrier parameter is already in k (see ric_stage for the split form)
#endif
      }
      // Y = Pss * Abar (5x4), S = Abar^T Y (4x4, lower)
      double Y[5][4];
#pragma unroll
      for (int i = 0; i < 5; ++i) applyA(A, PS(i, 0), PS(i, 1), PS(i, 2), PS(i, 3), PS(i, 4), Y[i][0], Y[i][1], Y[i][2], Y[i][3]);
      double Sm[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) applyA(A, Y[0][j], Y[1][j], Y[2][j], Y[3][j], Y[4][j], Sm[0][j], Sm[1][j], Sm[2][j], Sm[3][j]);
      double aw[4];
      applyA(A, wv[0], wv[1], wv[2], wv[3], wv[4], aw[0], aw[1], aw[2], aw[3]);
      const double ac[4] = {A.pp, -1.0, 0.0, A.sed};
      // r_s,t = grad f_s + lambda_t - A^T lambda_{t+1}
      double rs[5];
      rs[0] = lam[0] - ATl[0];
      rs[1] = lam[1] - ATl[1];
      rs[2] = lam[2] - ATl[2];
      rs[3] = gv2 * (s[3] - P.ref_v) + lam[3] - ATl[3];
      rs[4] = ge2 * s[5] + lam[5] - ATl[5];
      // new P, p
      double nss[15];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j)
          nss[i * (i + 1) / 2 + j] = Sm[i][j] + qc * ac[i] * ac[j] - (G0[i] * K0[j] + G1[i] * K1[j]);
      nss[0] += q0 + H.xx;
      nss[2] += q0;
      nss[5] += q0 + H.pp;
      nss[8] += H.vp;          // (v,psi)
      nss[9] += qv;
      const double qvce = qc * A.vce;
#pragma unroll
      for (int j = 0; j < 4; ++j) nss[10 + j] = qvce * ac[j];
      nss[13] += H.ev;         // (epsi,v)
      nss[14] = qe + H.ee + qvce * A.vce;
      double d0 = 0.0, d1 = 0.0;   // coupling with u_{t-1}
      if (t > 0 && !ls) { d0 = 2.0 * df * P.w_ddelta; d1 = 2.0 * df * P.w_da; }
#pragma unroll
      for (int i = 0; i < 4; ++i) { psu[i][0] = K0[i] * d0; psu[i][1] = K1[i] * d1; }
      puu00 = -d0 * d0 * i00; puu10 = -d0 * d1 * i10; puu11 = -d1 * d1 * i11;
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = rs[i] + aw[i] + gc * ac[i] - (G0[i] * k0 + G1[i] * k1);
      pv[4] = rs[4] + gc * A.vce;
      pu0 = d0 * k0; pu1 = d1 * k1;
#pragma unroll
      for (int i = 0; i < 15; ++i) pss[i] = nss[i];
      rc_next = gc2 * s[4] + lam[4];
      un0 = u0; un1 = u1;
#pragma unroll
      for (int k = 0; k < 6; ++k) { lamn[k] = lam[k]; sT[k] = s[k]; }
    }
    return ok;
  }


#if MPC_FUSE_FACTOR
  // ------------------------------------------------------------------------------------------
  // The Riccati recursion riding on the STEP sweep (MPC_FUSE_FACTOR).  ric_stage() is the stage body of factor() for a
  // new system (dw = 0, no second-order correction) at the iterate the STEP sweep has just formed -- its inputs are the
  // trial values the sweep holds in registers -- with the feed-forward split as k = ka + mu kb: of the right-hand side
  // only the barrier term of the controls, -mu / sl + mu / su, depends on mu, and the mu of the next iteration is decided
  // after the sweep (IpMonotoneMuUpdate.cpp:132-232, top_of_loop()).  Carried from stage to stage (RQ(i) = rq[i * rqs]):
  //   0..14 Pss   15..22 Psu   23..25 Puu   26..30 pv   31..32 pu   33..36 the mu-part of pv (its epsi entry is 0)
  //   37..38 the mu-part of pu
  // Returns false on a non-positive-definite control block (wrong inertia: the separate factor sweep takes over with
  // the delta_w ladder).
#define RQ(i) rq[(i) * rqs]
  MPC_HD void ric_terminal(const double* sT, const double* lamT) {
    const double gv2 = 2.0 * P.w_v * df, ge2 = 2.0 * P.w_epsi * df;
#pragma unroll
    for (int i = 0; i < kRicCarry; ++i) RQ(i) = 0.0;
    RQ(9) = gv2; RQ(14) = ge2;   // PS(3,3) = qv, PS(4,4) = qe; q0 = dw = 0
    RQ(26) = lamT[0]; RQ(27) = lamT[1]; RQ(28) = lamT[2];
    RQ(29) = gv2 * (sT[3] - P.ref_v) + lamT[3];
    RQ(30) = ge2 * sT[5] + lamT[5];
  }
  // s, lam: s_t, lambda_t; lamn, cteN: lambda_{t+1}, cte_{t+1}; u, z (zl0, zl1, zu0, zu1), sl (sl0, su0, sl1, su1): controls
  // of time t with their multipliers and slacks; gub0/gub1: grad_u f - B^T lambda_{t+1}; c: residual of the rows of time
  // t+1; A, ATl: linearisation at (s_t, u_t) and A^T lambda_{t+1}
  MPC_HD bool ric_stage(int t, const double* s, const double* lam, const double* lamn, double cteN, double zl0, double zl1,
                        double zu0, double zu1, double sl0, double su0, double sl1, double su1, double gub0, double gub1,
                        double sp, double cp, double se, double ce, double p1, double p2, double p3, const double* c,
                        const Lin& A, const double* ATl) {
    const double gv2 = 2.0 * P.w_v * df, ge2 = 2.0 * P.w_epsi * df, gc2 = 2.0 * P.w_cte * df;
    const double qv = gv2, qe = ge2, qc = gc2;
    double pss[15], psu[4][2], pv[5], pvb[4];
#pragma unroll
    for (int i = 0; i < 15; ++i) pss[i] = RQ(i);
#pragma unroll
    for (int i = 0; i < 4; ++i) { psu[i][0] = RQ(15 + 2 * i); psu[i][1] = RQ(16 + 2 * i); }
    const double puu00 = RQ(23), puu10 = RQ(24), puu11 = RQ(25);
#pragma unroll
    for (int i = 0; i < 5; ++i) pv[i] = RQ(26 + i);
    const double pu0 = RQ(31), pu1 = RQ(32);
#pragma unroll
    for (int i = 0; i < 4; ++i) pvb[i] = RQ(33 + i);
    const double pub0 = RQ(37), pub1 = RQ(38);
    const double rb[5] = {-c[0], -c[1], -c[2], -c[3], -c[5]}, cc = c[4];
    const Hes H = make_hes(P, lamn, s[3], sp, cp, se, ce, p1, p2, p3);
    const double isl0 = drcp(sl0), isu0 = drcp(su0), isl1 = drcp(sl1), isu1 = drcp(su1);
    const double r0 = hess_u(0, t) + zl0 * isl0 + zu0 * isu0;
    const double r1 = hess_u(1, t) + zl1 * isl1 + zu1 * isu1;
    const double rub0 = isu0 - isl0, rub1 = isu1 - isl1;   // d(ru) / d(mu)
    const double gc = (gc2 * cteN + lamn[4]) - qc * cc;   // linear coefficient on a_c^T dsigma (cte fold)
    // ---- recursion
    double wv[5], wu0 = pu0, wu1 = pu1;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double a = pv[i];
#pragma unroll
      for (int j = 0; j < 5; ++j) a += PS(i, j) * rb[j];
      wv[i] = a;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { wu0 += psu[i][0] * rb[i]; wu1 += psu[i][1] * rb[i]; }
    double T0[5], T1[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      T0[j] = A.beta * (PS(2, j) + PS(4, j)) + (j < 4 ? psu[j < 4 ? j : 0][0] : 0.0);
      T1[j] = P.dt * PS(3, j) + (j < 4 ? psu[j < 4 ? j : 0][1] : 0.0);
    }
    const double Tu00 = A.beta * psu[2][0] + puu00, Tu10 = P.dt * psu[3][0] + puu10, Tu11 = P.dt * psu[3][1] + puu11;
    const double L00 = r0 + A.beta * (T0[2] + T0[4]) + Tu00;
    const double L10 = A.beta * (T1[2] + T1[4]) + Tu10;
    const double L11 = r1 + P.dt * T1[3] + Tu11;
    const double det = L00 * L11 - L10 * L10;
    const bool ok = (L00 > 0.0) && (det > 0.0);
    const double idet = drcp(det);
    const double i00 = L11 * idet, i10 = -L10 * idet, i11 = L00 * idet;
    double G0[4], G1[4];
    applyA(A, T0[0], T0[1], T0[2], T0[3], T0[4], G0[0], G0[1], G0[2], G0[3]);
    applyA(A, T1[0], T1[1], T1[2], T1[3], T1[4], G1[0], G1[1], G1[2], G1[3]);
    G0[3] += H.m;
    const double h0 = gub0 + A.beta * (wv[2] + wv[4]) + wu0, h1 = gub1 + P.dt * wv[3] + wu1;
    const double hb0 = rub0 + A.beta * pvb[2] + pub0, hb1 = rub1 + P.dt * pvb[3] + pub1;
    double K0[4], K1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { K0[j] = i00 * G0[j] + i10 * G1[j]; K1[j] = i10 * G0[j] + i11 * G1[j]; }
    const double k0 = i00 * h0 + i10 * h1, k1 = i10 * h0 + i11 * h1;
    const double kb0 = i00 * hb0 + i10 * hb1, kb1 = i10 * hb0 + i11 * hb1;
    {
      const int ko = rec(t) + oKF;
#pragma unroll
      for (int j = 0; j < 4; ++j) { w(ko + j) = K0[j]; w(ko + 4 + j) = K1[j]; }
      w(ko + 8) = i00; w(ko + 9) = i10; w(ko + 10) = i11; w(ko + 11) = k0; w(ko + 12) = k1; w(ko + 13) = kb0; w(ko + 14) = kb1;
    }
    // Y = Pss * Abar (5x4), S = Abar^T Y (4x4, lower)
    double Y[5][4];
#pragma unroll
    for (int i = 0; i < 5; ++i) applyA(A, PS(i, 0), PS(i, 1), PS(i, 2), PS(i, 3), PS(i, 4), Y[i][0], Y[i][1], Y[i][2], Y[i][3]);
    double Sm[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) applyA(A, Y[0][j], Y[1][j], Y[2][j], Y[3][j], Y[4][j], Sm[0][j], Sm[1][j], Sm[2][j], Sm[3][j]);
    double aw[4], awb[4];
    applyA(A, wv[0], wv[1], wv[2], wv[3], wv[4], aw[0], aw[1], aw[2], aw[3]);
    applyA(A, pvb[0], pvb[1], pvb[2], pvb[3], 0.0, awb[0], awb[1], awb[2], awb[3]);
    const double ac[4] = {A.pp, -1.0, 0.0, A.sed};
    double rs[5];
    rs[0] = lam[0] - ATl[0];
    rs[1] = lam[1] - ATl[1];
    rs[2] = lam[2] - ATl[2];
    rs[3] = gv2 * (s[3] - P.ref_v) + lam[3] - ATl[3];
    rs[4] = ge2 * s[5] + lam[5] - ATl[5];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j)
        RQ(i * (i + 1) / 2 + j) = Sm[i][j] + qc * ac[i] * ac[j] - (G0[i] * K0[j] + G1[i] * K1[j]) +
                                  (i == j ? (i == 0 ? H.xx : (i == 2 ? H.pp : (i == 3 ? qv : 0.0))) : (i == 3 && j == 2 ? H.vp : 0.0));
    const double qvce = qc * A.vce;
#pragma unroll
    for (int j = 0; j < 4; ++j) RQ(10 + j) = qvce * ac[j] + (j == 3 ? H.ev : 0.0);
    RQ(14) = qe + H.ee + qvce * A.vce;
    double d0 = 0.0, d1 = 0.0;   // coupling with u_{t-1}
    if (t > 0) { d0 = 2.0 * df * P.w_ddelta; d1 = 2.0 * df * P.w_da; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { RQ(15 + 2 * i) = K0[i] * d0; RQ(16 + 2 * i) = K1[i] * d1; }
    RQ(23) = -d0 * d0 * i00; RQ(24) = -d0 * d1 * i10; RQ(25) = -d1 * d1 * i11;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      RQ(26 + i) = rs[i] + aw[i] + gc * ac[i] - (G0[i] * k0 + G1[i] * k1);
      RQ(33 + i) = awb[i] - (G0[i] * kb0 + G1[i] * kb1);
    }
    RQ(30) = rs[4] + gc * A.vce;
    RQ(31) = d0 * k0; RQ(32) = d1 * k1;
    RQ(37) = d0 * kb0; RQ(38) = d1 * kb1;
    return ok;
  }
#undef RQ
#endif

  // ------------------------------------------------------------------------------------------
  // Forward sweep: dx (-> DS, DU), fraction-to-the-boundary steps, directional derivative of the barrier.
  MPC_HD void forward(bool use_csoc) { forward_body(fl(F_LS), use_csoc); }
  MPC_HD void forward_spec(bool use_csoc) {
#if MPC_SPECIALIZE_SWEEPS
    if (fl(F_LS)) forward_t<true, false>();
    else if (use_csoc) forward_t<false, true>();
    else forward_t<false, false>();
#else
    forward_body(fl(F_LS), use_csoc);
#endif
  }
  template <bool LS, bool CSOC>
  MPC_HD void forward_t() { forward_body(LS, CSOC); }
  MPC_HD void forward_body(const bool ls, const bool use_csoc) {
    const int bX = kX * cur, bC = use_csoc ? (int)oCSOC : bX + xC;
    const double gv2 = 2.0 * P.w_v * df, ge2 = 2.0 * P.w_epsi * df, gc2 = 2.0 * P.w_cte * df;
    const double tinytol = 10.0 * DBL_EPSILON;
    double ds[6] = {0, 0, 0, 0, 0, 0}, dup0 = 0.0, dup1 = 0.0;
    double a_pr = 1.0, a_du = 1.0, gbd = 0.0;
    bool nottiny = false;
    double um0 = 0.0, um1 = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) w(rec(0) + oDS + k) = 0.0;
    double u0 = w(rec(0) + bX + xU), u1 = w(rec(0) + bX + xU + 1);
    double s[6], snx[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] = w(rec(0) + bX + xS + k);
    for (int t = 0; t < M; ++t) {
      const int r = rec(t) + bX, rn = rec(t + 1);
      if (t + 1 < M) {
        w.prefetch(rn + oKF, kKF); w.prefetch(rn + kRec + bX + xS, 8);
        w.prefetch(rn + bX + xZL, MPC_STORE_TRIG ? (MPC_STORE_PSIDES ? 9 : 8) : 4);   // ZL, ZU, the trig values and the reference heading are contiguous
        if (!ls && (MPC_STORE_C || use_csoc)) w.prefetch(rn + kRec + bC, 6);
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) snx[k] = w(rn + bX + xS + k);
      const int ko = rec(t) + oKF;
      double K0[4], K1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { K0[j] = w(ko + j); K1[j] = w(ko + 4 + j); }
      const double i00 = w(ko + 8), i10 = w(ko + 9), i11 = w(ko + 10);
#if MPC_FUSE_FACTOR
      const double k0 = w(ko + 11) + mu * w(ko + 13), k1 = w(ko + 12) + mu * w(ko + 14);   // kb = 0 from factor()
#else
      const double k0 = w(ko + 11), k1 = w(ko + 12);
#endif
      double sp, cp, se, ce;
      trig_of(r, s, sp, cp, se, ce);
      double p0, p1, p2, p3;
      poly_eval(cf, s[0], p0, p1, p2, p3);
      double c[6];
      if (ls) { c[0] = c[1] = c[2] = c[3] = c[4] = c[5] = 0.0; }
      else if (MPC_STORE_C || use_csoc) {
#pragma unroll
        for (int k = 0; k < 6; ++k) c[k] = w(rn + bC + k);
      } else {
        const double uu[2] = {u0, u1};
        residual(s, uu, snx, sp, cp, se, p0, psides_of(r, p1), c);
      }
      double un0 = 0.0, un1 = 0.0;
      if (t < M - 1) { un0 = w(rn + bX + xU); un1 = w(rn + bX + xU + 1); }
      double d0 = 0.0, d1 = 0.0;
      if (t > 0 && !ls) { d0 = 2.0 * df * P.w_ddelta; d1 = 2.0 * df * P.w_da; }
      const double e0 = d0 * dup0, e1 = d1 * dup1;
      const double du0 = -(K0[0] * ds[0] + K0[1] * ds[1] + K0[2] * ds[2] + K0[3] * ds[3] + k0) + (i00 * e0 + i10 * e1);
      const double du1 = -(K1[0] * ds[0] + K1[1] * ds[1] + K1[2] * ds[2] + K1[3] * ds[3] + k1) + (i10 * e0 + i11 * e1);
      w(rec(t) + oDU) = du0; w(rec(t) + oDU + 1) = du1;
      const Lin A = make_lin(P, s[3], u0, sp, cp, se, ce, p1, p2);
      if (!ls) {
        // objective / barrier directional derivative and step bounds for u_t
        const double zl0 = w(r + xZL), zl1 = w(r + xZL + 1), zu0 = w(r + xZU), zu1 = w(r + xZU + 1);
        double sl0 = u0 - P.xl[0], su0 = P.xu[0] - u0, sl1 = u1 - P.xl[1], su1 = P.xu[1] - u1;
        safe_slack4(sl0, su0, sl1, su1, mu, zl0, zu0, zl1, zu1, P.xl, P.xu);
        const double isl0 = drcp(sl0), isu0 = drcp(su0), isl1 = drcp(sl1), isu1 = drcp(su1);
        gbd += (grad_u(0, t, u0, um0, un0) - mu * isl0 + mu * isu0) * du0 + (grad_u(1, t, u1, um1, un1) - mu * isl1 + mu * isu1) * du1;
        gbd += gv2 * (s[3] - P.ref_v) * ds[3] + gc2 * s[4] * ds[4] + ge2 * s[5] * ds[5];
        // fraction to the boundary: alpha <= tau * slack / |du| (IpDenseVector.cpp:928-970)
        // Written without branches so that the six divisions are independent chains the scheduler can interleave:
        // -tau / d == -(tau / d) exactly, one quotient serves both signs of du; a direction component that does not
        // move towards a bound contributes the neutral candidate 2 (alpha <= 1).
        const double dzl0 = (mu - sl0 * zl0 - zl0 * du0) * isl0, dzu0 = (mu - su0 * zu0 + zu0 * du0) * isu0;
        const double dzl1 = (mu - sl1 * zl1 - zl1 * du1) * isl1, dzu1 = (mu - su1 * zu1 + zu1 * du1) * isu1;
        const double qu0 = ddiv(tau, du0), qu1 = ddiv(tau, du1);
        const double qzl0 = ddiv(tau, dzl0), qzu0 = ddiv(tau, dzu0), qzl1 = ddiv(tau, dzl1), qzu1 = ddiv(tau, dzu1);
        const double cu0 = du0 < 0.0 ? -qu0 * sl0 : (du0 > 0.0 ? qu0 * su0 : 2.0);
        const double cu1 = du1 < 0.0 ? -qu1 * sl1 : (du1 > 0.0 ? qu1 * su1 : 2.0);
        a_pr = dmin(a_pr, dmin(cu0, cu1));
        const double czl0 = dzl0 < 0.0 ? -qzl0 * zl0 : 2.0, czu0 = dzu0 < 0.0 ? -qzu0 * zu0 : 2.0;
        const double czl1 = dzl1 < 0.0 ? -qzl1 * zl1 : 2.0, czu1 = dzu1 < 0.0 ? -qzu1 * zu1 : 2.0;
        a_du = dmin(a_du, dmin(dmin(czl0, czu0), dmin(czl1, czu1)));
        // tiny-step test |dx_i| / (|x_i| + 1) <= 10 eps for all i (IpBacktrackingLineSearch.cpp:1145-1200)
        nottiny = nottiny || fabs(du0) > tinytol * (fabs(u0) + 1.0) || fabs(du1) > tinytol * (fabs(u1) + 1.0);
#pragma unroll
        for (int k = 0; k < 6; ++k) nottiny = nottiny || fabs(ds[k]) > tinytol * (fabs(s[k]) + 1.0);
      }
      double dn[6];
      dn[0] = ds[0] + A.a1 * ds[2] + A.a3 * ds[3] - c[0];
      dn[1] = ds[1] + A.a2 * ds[2] + A.a4 * ds[3] - c[1];
      dn[2] = ds[2] + A.a5 * ds[3] + A.beta * du0 - c[2];
      dn[3] = ds[3] + P.dt * du1 - c[3];
      dn[4] = A.pp * ds[0] - ds[1] + A.sed * ds[3] + A.vce * ds[5] - c[4];
      dn[5] = -A.kap * ds[0] + ds[2] + A.a5 * ds[3] + A.beta * du0 - c[5];
#pragma unroll
      for (int k = 0; k < 6; ++k) { ds[k] = dn[k]; w(rn + oDS + k) = dn[k]; s[k] = snx[k]; }
      dup0 = du0; dup1 = du1; um0 = u0; um1 = u1; u0 = un0; u1 = un1;
    }
    if (!ls) {
#pragma unroll
      for (int k = 0; k < 6; ++k) nottiny = nottiny || fabs(ds[k]) > tinytol * (fabs(s[k]) + 1.0);
      gbd += gv2 * (s[3] - P.ref_v) * ds[3] + gc2 * s[4] * ds[4] + ge2 * s[5] * ds[5];
    }
    fw_alpha_pr = a_pr; fw_alpha_du = a_du; fw_gbd = gbd; fw_tiny = !nottiny;
  }

  // ------------------------------------------------------------------------------------------
  // Least-square multiplier initialisation (IpLeastSquareMults.cpp:40-94): lambda from the stationarity rows of the
  // H = I system solved by factor()/forward() in F_LS mode; x, z unchanged.  Also every norm the convergence
  // test / mu update need at the start point.  The first pass (zero = true) only measures ||lambda_LS||_inf and
  // evaluates the norms for lambda = 0 (what Ipopt falls back to when the estimate exceeds constr_mult_init_max =
  // 1000, the common case here) without storing anything; a second pass (zero = false) stores the estimate.
  MPC_HD void accept_ls(bool zero) {
    const double gv2 = 2.0 * P.w_v * df, ge2 = 2.0 * P.w_epsi * df, gc2 = 2.0 * P.w_cte * df;
    const int bX = kX * cur;
    double lp[6], ln[6];
    double dinf = 0.0, l1 = 0.0, zz1 = 0.0, szmx = 0.0, szmn = 1e300, xm = 0.0, dlm = 0.0;
    {
      const int r = rec(M) + bX;
      double s[6], ds[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { s[k] = w(r + xS + k); ds[k] = w(rec(M) + oDS + k); xm = dmax(xm, fabs(s[k])); }
      lp[0] = -ds[0]; lp[1] = -ds[1]; lp[2] = -ds[2];
      lp[3] = -ds[3] - gv2 * (s[3] - P.ref_v);
      lp[4] = -ds[4] - gc2 * s[4];
      lp[5] = -ds[5] - ge2 * s[5];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        dlm = dmax(dlm, fabs(lp[k])); ln[k] = zero ? 0.0 : lp[k];
        if (!zero) w(r + xLAM + k) = ln[k];   // lambda is already zero from init()
        l1 += fabs(ln[k]);
      }
      dinf = dmax(dinf, dmax(fabs(ln[0]), dmax(fabs(ln[1]), fabs(ln[2]))));
      dinf = dmax(dinf, fabs(gv2 * (s[3] - P.ref_v) + ln[3]));
      dinf = dmax(dinf, fabs(gc2 * s[4] + ln[4]));
      dinf = dmax(dinf, fabs(ge2 * s[5] + ln[5]));
    }
    double un0 = 0.0, un1 = 0.0;
    for (int t = M - 1; t >= 0; --t) {
      const int r = rec(t) + bX;
      double s[6], ds[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { s[k] = w(r + xS + k); ds[k] = w(rec(t) + oDS + k); xm = dmax(xm, fabs(s[k])); }
      const double u0 = w(r + xU), u1 = w(r + xU + 1);
      double um0 = 0.0, um1 = 0.0;
      if (t > 0) { um0 = w(r - kRec + xU); um1 = w(r - kRec + xU + 1); }
      double sp, cp, se, ce;
      trig_of(r, s, sp, cp, se, ce);
      const double zl0 = w(r + xZL), zl1 = w(r + xZL + 1), zu0 = w(r + xZU), zu1 = w(r + xZU + 1);
      double p0, p1, p2, p3;
      poly_eval(cf, s[0], p0, p1, p2, p3);
      const Lin A = make_lin(P, s[3], u0, sp, cp, se, ce, p1, p2);
      double at[6], nl[6];
      applyAT6(A, lp, at);
      nl[0] = at[0] - ds[0];
      nl[1] = at[1] - ds[1];
      nl[2] = at[2] - ds[2];
      nl[3] = at[3] - ds[3] - gv2 * (s[3] - P.ref_v);
      nl[4] = -ds[4] - gc2 * s[4];
      nl[5] = at[5] - ds[5] - ge2 * s[5];
      double atn[6];
      applyAT6(A, ln, atn);
      double lnew[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        lp[k] = nl[k]; dlm = dmax(dlm, fabs(nl[k]));
        lnew[k] = zero ? 0.0 : nl[k];
        if (!zero) w(r + xLAM + k) = lnew[k];
        l1 += fabs(lnew[k]);
      }
      xm = dmax(xm, dmax(fabs(u0), fabs(u1)));
      const double sl0 = safe_slack(u0 - P.xl[0], mu, zl0, P.xl[0]), su0 = safe_slack(P.xu[0] - u0, mu, zu0, P.xu[0]);
      const double sl1 = safe_slack(u1 - P.xl[1], mu, zl1, P.xl[1]), su1 = safe_slack(P.xu[1] - u1, mu, zu1, P.xu[1]);
      zz1 += zl0 + zl1 + zu0 + zu1;
      const double c0 = sl0 * zl0, c1 = sl1 * zl1, c2 = su0 * zu0, c3 = su1 * zu1;
      szmx = dmax(szmx, dmax(dmax(c0, c1), dmax(c2, c3)));
      szmn = dmin(szmn, dmin(dmin(c0, c1), dmin(c2, c3)));
      dinf = dmax(dinf, dmax(fabs(lnew[0] - atn[0]), dmax(fabs(lnew[1] - atn[1]), fabs(lnew[2] - atn[2]))));
      dinf = dmax(dinf, fabs(gv2 * (s[3] - P.ref_v) + lnew[3] - atn[3]));
      dinf = dmax(dinf, fabs(gc2 * s[4] + lnew[4]));
      dinf = dmax(dinf, fabs(ge2 * s[5] + lnew[5] - atn[5]));
      const double g0 = grad_u(0, t, u0, um0, un0) - A.beta * (ln[2] + ln[5]) - zl0 + zu0;
      const double g1 = grad_u(1, t, u1, um1, un1) - P.dt * ln[3] - zl1 + zu1;
      dinf = dmax(dinf, dmax(fabs(g0), fabs(g1)));
#pragma unroll
      for (int k = 0; k < 6; ++k) ln[k] = lnew[k];
      un0 = u0; un1 = u1;
    }
    dualinf = dinf; lam1 = l1; z1 = zz1; sz_max = szmx; sz_min = szmn; xmaxabs = xm; dlam_max = dlm;
  }

  // ------------------------------------------------------------------------------------------
  // The fused STEP sweep (backward).  For the trial step sizes (a for x and lambda, a_du for z) it computes, stage by
  // stage: the trial point x + a dx with its trig values, constraint residuals, objective and log-barrier
  // (MPC.cpp:57-138; what Ipopt evaluates in the line search), dlam from the stationarity rows of the CURRENT
  // linearisation, the trial multipliers lambda + a dlam and z + a_du dz with the kappa_sigma reset
  // (IpIpoptAlg.cpp:880-951), and every norm the convergence test / mu update need at the trial iterate
  // (IpIpoptCalculatedQuantities.cpp:2672-2832, 3279-3306).  Everything is written to the OTHER copy of the iterate
  // block; the caller flips `cur` if the trial point is accepted.
  // FUSE (MPC_FUSE_FACTOR): the Riccati factorisation of the next iteration's system at the trial iterate rides on the
  // sweep (ric_stage); the return value says whether it is usable (right inertia, slack safeguard not involved).
  template <bool FUSE = false>
  MPC_HD bool step_sweep(double a, double a_lam, double a_du, double dw) {
    bool ric_ok = true;
    const double qv = 2.0 * P.w_v * df + dw, qe = 2.0 * P.w_epsi * df + dw, qc = 2.0 * P.w_cte * df + dw, q0 = dw;
    const double gv2 = 2.0 * P.w_v * df, ge2 = 2.0 * P.w_epsi * df, gc2 = 2.0 * P.w_cte * df;
    const int bO = kX * cur, bN = kX * (cur ^ 1);
    // Values carried from stage t+1 to stage t live in the carry buffer (shared memory in the per-pass kernel) and are
    // loaded where they are used, which keeps them out of the register file across the heavy parts of the stage body:
    // CR(0..5) lambda^+_{t+1}, CR(6..11) lambda_old_{t+1}, CR(12..17) lambda_new_{t+1}, CR(18..23) s_new_{t+1}
#define CR(i) cr[(i) * cs]
    double dinf = 0.0, l1 = 0.0, zz1 = 0.0, szmx = 0.0, szmn = 1e300, xm = 0.0, dlm = 0.0;
    double f = 0.0, th = 0.0, cm = 0.0, slog = 0.0;
#if defined(__CUDA_ARCH__) && MPC_ASYNC_STAGE && MPC_STEP_ASYNC_STAGE && MPC_STORE_TRIG
    const bool staged = !FUSE && sb != nullptr;
    if (staged) step_stage_request(M - 1, bO);
#else
    const bool staged = false;
#endif
    {
      const int r = rec(M);
      double s[6], ds[6], lam[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { s[k] = w(r + bO + xS + k); ds[k] = w(r + oDS + k); lam[k] = w(r + bO + xLAM + k); }
      double lp[6], ln[6], snn[6];
      lp[0] = -q0 * ds[0]; lp[1] = -q0 * ds[1]; lp[2] = -q0 * ds[2];
      lp[3] = -qv * ds[3] - gv2 * (s[3] - P.ref_v);
      lp[4] = -qc * ds[4] - gc2 * s[4];
      lp[5] = -qe * ds[5] - ge2 * s[5];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        dlm = dmax(dlm, fabs(lp[k] - lam[k]));
        ln[k] = lam[k] + a_lam * (lp[k] - lam[k]);
        snn[k] = s[k] + a * ds[k];
        w(r + bN + xLAM + k) = ln[k];
        w(r + bN + xS + k) = snn[k];
        l1 += fabs(ln[k]);
        xm = dmax(xm, fabs(snn[k]));
        CR(k) = lp[k]; CR(6 + k) = lam[k]; CR(12 + k) = ln[k]; CR(18 + k) = snn[k];
      }
      f += state_cost(snn);
      dinf = dmax(dinf, dmax(fabs(ln[0]), dmax(fabs(ln[1]), fabs(ln[2]))));
      dinf = dmax(dinf, fabs(gv2 * (snn[3] - P.ref_v) + ln[3]));
      dinf = dmax(dinf, fabs(gc2 * snn[4] + ln[4]));
      dinf = dmax(dinf, fabs(ge2 * snn[5] + ln[5]));
#if MPC_FUSE_FACTOR
      if (FUSE) ric_terminal(snn, ln);
#endif
    }
    double unn0 = 0.0, unn1 = 0.0;   // u_new at t+1
    double uc0 = 0.0, uc1 = 0.0, duc0 = 0.0, duc1 = 0.0;
    { const int r = rec(M - 1); uc0 = w(r + bO + xU); uc1 = w(r + bO + xU + 1); duc0 = w(r + oDU); duc1 = w(r + oDU + 1); }
    for (int t = M - 1; t >= 0; --t) {
      const int r = rec(t);
      if (t > 0 && !staged) { w.prefetch(r - kRec + bO + xS, MPC_STORE_TRIG ? 22 : 18); w.prefetch(r - kRec + oDS, 8); }
      double s[6], ds[6], lam[6];
      const double u0 = uc0, u1 = uc1, du0 = duc0, du1 = duc1;
      double um0 = 0.0, um1 = 0.0, dum0 = 0.0, dum1 = 0.0;
      double zl0, zl1, zu0, zu1;
      double trs0 = 0.0, trs1 = 0.0, trs2 = 0.0, trs3 = 0.0;   // staged sin / cos values of the current point
#if defined(__CUDA_ARCH__) && MPC_ASYNC_STAGE && MPC_STEP_ASYNC_STAGE && MPC_STORE_TRIG
      if (staged) {
        // this stage's rows were requested one stage ago (before the terminal block for t = M-1); request the next, then use
        async_wait_all();
        if (t > 0) step_stage_request(t - 1, bO);
        const double* q = sb + (size_t)((t & 1) * kStepStageVals) * sbs;
#pragma unroll
        for (int k = 0; k < 6; ++k) { s[k] = q[(xS + k) * sbs]; lam[k] = q[(xLAM + k) * sbs]; ds[k] = q[(22 + k) * sbs]; }
        zl0 = q[xZL * sbs]; zl1 = q[(xZL + 1) * sbs]; zu0 = q[xZU * sbs]; zu1 = q[(xZU + 1) * sbs];
        trs0 = q[xTR * sbs]; trs1 = q[(xTR + 1) * sbs]; trs2 = q[(xTR + 2) * sbs]; trs3 = q[(xTR + 3) * sbs];
        if (t > 0) { um0 = q[28 * sbs]; um1 = q[29 * sbs]; dum0 = q[30 * sbs]; dum1 = q[31 * sbs]; }
      } else
#endif
      {
#pragma unroll
        for (int k = 0; k < 6; ++k) { s[k] = w(r + bO + xS + k); ds[k] = w(r + oDS + k); lam[k] = w(r + bO + xLAM + k); }
        if (t > 0) { um0 = w(r - kRec + bO + xU); um1 = w(r - kRec + bO + xU + 1); dum0 = w(r - kRec + oDU); dum1 = w(r - kRec + oDU + 1); }
        zl0 = w(r + bO + xZL); zl1 = w(r + bO + xZL + 1); zu0 = w(r + bO + xZU); zu1 = w(r + bO + xZU + 1);
      }
      // ---- current point: lambda^+_t from the stationarity rows
      double lpn[6];   // lambda^+_t
      {
        double lp[6], lo[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { lp[k] = CR(k); lo[k] = CR(6 + k); }
        double spo, cpo, seo, ceo;
        if (staged) { spo = trs0; cpo = trs1; seo = trs2; ceo = trs3; }
        else trig_of(r + bO, s, spo, cpo, seo, ceo);
        double p0, p1, p2, p3;
        poly_eval(cf, s[0], p0, p1, p2, p3);
        const Lin A = make_lin(P, s[3], u0, spo, cpo, seo, ceo, p1, p2);
        const Hes H = make_hes(P, lo, s[3], spo, cpo, seo, ceo, p1, p2, p3);
        double at[6];
        applyAT6(A, lp, at);
        double nl[6];
        nl[0] = at[0] - (q0 + H.xx) * ds[0];
        nl[1] = at[1] - q0 * ds[1];
        nl[2] = at[2] - ((q0 + H.pp) * ds[2] + H.vp * ds[3]);
        nl[3] = at[3] - (H.vp * ds[2] + qv * ds[3] + H.ev * ds[5] + H.m * du0) - gv2 * (s[3] - P.ref_v);
        nl[4] = -qc * ds[4] - gc2 * s[4];
        nl[5] = at[5] - ((qe + H.ee) * ds[5] + H.ev * ds[3]) - ge2 * s[5];
#pragma unroll
        for (int k = 0; k < 6; ++k) { lpn[k] = nl[k]; CR(k) = nl[k]; CR(6 + k) = lam[k]; dlm = dmax(dlm, fabs(nl[k] - lam[k])); }
      }
      // ---- trial point
      double sn[6], lnew[6], un[2];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        sn[k] = s[k] + a * ds[k];
        lnew[k] = lam[k] + a_lam * (lpn[k] - lam[k]);
        w(r + bN + xLAM + k) = lnew[k];
        w(r + bN + xS + k) = sn[k];
        l1 += fabs(lnew[k]);
        xm = dmax(xm, fabs(sn[k]));
      }
      un[0] = u0 + a * du0; un[1] = u1 + a * du1;
      const double umn0 = um0 + a * dum0, umn1 = um1 + a * dum1;
      xm = dmax(xm, dmax(fabs(un[0]), fabs(un[1])));
      w(r + bN + xU) = un[0]; w(r + bN + xU + 1) = un[1];
      {
        double sl0 = u0 - P.xl[0], su0 = P.xu[0] - u0, sl1 = u1 - P.xl[1], su1 = P.xu[1] - u1;
        safe_slack4(sl0, su0, sl1, su1, mu, zl0, zu0, zl1, zu1, P.xl, P.xu);
        // the barrier of the trial point is evaluated with the current z (only matters in the slack safeguard)
        double tl0 = un[0] - P.xl[0], tu0 = P.xu[0] - un[0], tl1 = un[1] - P.xl[1], tu1 = P.xu[1] - un[1];
        safe_slack4(tl0, tu0, tl1, tu1, mu, zl0, zu0, zl1, zu1, P.xl, P.xu);
        const double b0 = tl0 * tu0;
        const double b1 = tl1 * tu1;
        slog += log(b0 * b1);
        zl0 += a_du * ddiv(mu - sl0 * zl0 - zl0 * du0, sl0);
        zu0 += a_du * ddiv(mu - su0 * zu0 + zu0 * du0, su0);
        zl1 += a_du * ddiv(mu - sl1 * zl1 - zl1 * du1, sl1);
        zu1 += a_du * ddiv(mu - su1 * zu1 + zu1 * du1, su1);
      }
      double nsl0 = un[0] - P.xl[0], nsu0 = P.xu[0] - un[0], nsl1 = un[1] - P.xl[1], nsu1 = P.xu[1] - un[1];
      if (FUSE) {   // the slack safeguard depends on mu (practically never active): leave such a system to factor()
        const double thr = DBL_EPSILON * dmin(mu, 1.0);
        if ((nsl0 < thr) | (nsu0 < thr) | (nsl1 < thr) | (nsu1 < thr)) ric_ok = false;
      }
      safe_slack4(nsl0, nsu0, nsl1, nsu1, mu, zl0, zu0, zl1, zu1, P.xl, P.xu);
      {   // kappa_sigma = 1e10 (IpIpoptAlg.cpp:880-951): z stays within [1e-10, 1e10] * mu / slack.  The exact bounds need
          // four divisions; they are evaluated only when the cheap product test says a multiplier is within a factor 4
          // of one of them (any clamp the exact test would apply passes through here: the margin dwarfs the rounding)
        const double lo = 4e-10 * mu, hi = 2.5e9 * mu;
        const double q0 = nsl0 * zl0, q1 = nsu0 * zu0, q2 = nsl1 * zl1, q3 = nsu1 * zu1;
        if (!((q0 > lo) & (q0 < hi) & (q1 > lo) & (q1 < hi) & (q2 > lo) & (q2 < hi) & (q3 > lo) & (q3 < hi))) {
          const double m0 = mu / nsl0, m1 = mu / nsu0, m2 = mu / nsl1, m3 = mu / nsu1;
          zl0 = dclamp(zl0, 1e-10 * m0, 1e10 * m0);
          zu0 = dclamp(zu0, 1e-10 * m1, 1e10 * m1);
          zl1 = dclamp(zl1, 1e-10 * m2, 1e10 * m2);
          zu1 = dclamp(zu1, 1e-10 * m3, 1e10 * m3);
        }
        w(r + bN + xZL) = zl0; w(r + bN + xZL + 1) = zl1; w(r + bN + xZU) = zu0; w(r + bN + xZU + 1) = zu1;
      }
      zz1 += zl0 + zl1 + zu0 + zu1;
      {
        const double c0 = nsl0 * zl0, c1 = nsl1 * zl1, c2 = nsu0 * zu0, c3 = nsu1 * zu1;
        szmx = dmax(szmx, dmax(dmax(c0, c1), dmax(c2, c3)));
        szmn = dmin(szmn, dmin(dmin(c0, c1), dmin(c2, c3)));
      }
      // residuals, objective and grad_x L at the trial point
      {
        double spn, cpn, sen, cen, p0, p1, p2, p3, c[6];
        sincos(sn[2], &spn, &cpn);
        sincos(sn[5], &sen, &cen);
        poly_eval(cf, sn[0], p0, p1, p2, p3);
        double snn[6], ln[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { snn[k] = CR(18 + k); ln[k] = CR(12 + k); }
        const double psides = atan(p1);
        residual(sn, un, snn, spn, cpn, sen, p0, psides, c);
#if MPC_STORE_TRIG
        w(r + bN + xTR) = spn; w(r + bN + xTR + 1) = cpn; w(r + bN + xTR + 2) = sen; w(r + bN + xTR + 3) = cen;
#endif
        store_psides(r + bN, psides);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#if MPC_STORE_C
          w(r + kRec + bN + xC + k) = c[k];
#endif
          th += fabs(c[k]); cm = dmax(cm, fabs(c[k]));
        }
        f += state_cost(sn) + P.w_delta * (un[0] * un[0]) + P.w_a * (un[1] * un[1]);
        if (t > 0) f += P.w_ddelta * ((un[0] - umn0) * (un[0] - umn0)) + P.w_da * ((un[1] - umn1) * (un[1] - umn1));
        const Lin A = make_lin(P, sn[3], un[0], spn, cpn, sen, cen, p1, p2);
        double at[6];
        applyAT6(A, ln, at);
        dinf = dmax(dinf, dmax(fabs(lnew[0] - at[0]), dmax(fabs(lnew[1] - at[1]), fabs(lnew[2] - at[2]))));
        dinf = dmax(dinf, fabs(gv2 * (sn[3] - P.ref_v) + lnew[3] - at[3]));
        dinf = dmax(dinf, fabs(gc2 * sn[4] + lnew[4]));
        dinf = dmax(dinf, fabs(ge2 * sn[5] + lnew[5] - at[5]));
        const double gub0 = grad_u(0, t, un[0], umn0, unn0) - A.beta * (ln[2] + ln[5]);   // grad_u f - B^T lambda_{t+1}
        const double gub1 = grad_u(1, t, un[1], umn1, unn1) - P.dt * ln[3];
        const double g0 = gub0 - zl0 + zu0;
        const double g1 = gub1 - zl1 + zu1;
        dinf = dmax(dinf, dmax(fabs(g0), fabs(g1)));
#if MPC_FUSE_FACTOR
        if (FUSE) {
          const bool ok = ric_stage(t, sn, lnew, ln, snn[4], zl0, zl1, zu0, zu1, nsl0, nsu0, nsl1, nsu1, gub0, gub1, spn, cpn, sen, cen,
                                    p1, p2, p3, c, A, at);
          ric_ok = ric_ok && ok;
        }
#endif
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) { CR(12 + k) = lnew[k]; CR(18 + k) = sn[k]; }
      unn0 = un[0]; unn1 = un[1];
      uc0 = um0; uc1 = um1; duc0 = dum0; duc1 = dum1;
    }
    dualinf = dinf; lam1 = l1; z1 = zz1; sz_max = szmx; sz_min = szmn; xmaxabs = xm; dlam_max = dlm;
    tr_f = df * f; tr_theta = th; tr_priminf = cm; tr_sumlog = slog;
#undef CR
    return ric_ok;
  }

  // ------------------------------------------------------------------------------------------
  // filter (IpFilter.cpp:41-77) and acceptance tests (IpFilterLSAcceptor.cpp:246-382)
  MPC_HD static bool cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(bas); }
  MPC_HD bool filter_ok(double phi, double th) const {
    bool ok = true;
    for (int i = 0; i < nf; ++i)
      if (!(phi <= filt(2 * i) || th <= filt(2 * i + 1))) ok = false;
    return ok;
  }
  MPC_HD void filter_add(double phi, double th) {
    int wr = 0;
    for (int i = 0; i < nf; ++i) {
      const double a = filt(2 * i), b = filt(2 * i + 1);
      if (!(a >= phi && b >= th)) { filt(2 * wr) = a; filt(2 * wr + 1) = b; ++wr; }
    }
    nf = wr;
    if (nf < kMaxFilter) { filt(2 * nf) = phi; filt(2 * nf + 1) = th; ++nf; }
  }
  MPC_HD bool is_ftype(double at) const { return ref_gbd < 0.0 && at * pow(-ref_gbd, 2.3) > 1.0 * pow(ref_theta, 1.1); }
  MPC_HD bool armijo(double at, double trial_barr) const { return cmp_le(trial_barr - ref_barr, 1e-8 * at * ref_gbd, ref_barr); }
  MPC_HD bool check_accept(double at, double trial_theta, double trial_barr) {
    if (theta_max < 0.0) theta_max = 1e4 * dmax(1.0, ref_theta);
    if (theta_min < 0.0) theta_min = 1e-4 * dmax(1.0, ref_theta);
    if (theta_max > 0.0 && trial_theta > theta_max) return false;
    bool acc;
    if (at > 0.0 && is_ftype(at) && ref_theta <= theta_min) acc = armijo(at, trial_barr);
    else {
      acc = true;
      if (trial_barr > ref_barr) {
        double bas = 1.0;
        if (fabs(ref_barr) > 10.0) bas = log10(fabs(ref_barr));
        if (log10(trial_barr - ref_barr) > 5.0 + bas) acc = false;
      }
      if (acc)
        acc = cmp_le(trial_theta, (1.0 - 1e-5) * ref_theta, ref_theta) || cmp_le(trial_barr - ref_barr, -1e-8 * ref_theta, ref_barr);
    }
    if (!acc) { setfl(F_LASTREJF, false); return false; }
    if (!filter_ok(trial_barr, trial_theta)) { setfl(F_LASTREJF, true); return false; }
    // filter reset heuristic (IpFilterLSAcceptor.cpp:357-379; max_filter_resets = 5 is never reached: Ipopt 3.12.7 does
    // not count the resets): 5 successive iterations whose last rejected trial point was rejected by the filter
    if (fl(F_LASTREJF)) {
      if (++filt_rej_count >= 5) { nf = 0; filt_rej_count = 0; }
    } else filt_rej_count = 0;
    setfl(F_LASTREJF, false);
    return true;
  }

  // ------------------------------------------------------------------------------------------
  // Top of Ipopt's main loop at the current iterate: convergence tests and the monotone barrier update.
  // Returns false when the solve is finished (status set).
  MPC_HD bool top_of_loop() {
    const int nb = 2 * M, m = 6 * N;
    const double sc = dmax(100.0, z1 / (2.0 * nb)) / 100.0;
    const double sd = dmax(100.0, (lam1 + z1) / (m + 2.0 * nb)) / 100.0;
    const double compl0 = sz_max;
    const double E0 = dmax(dualinf / sd, dmax(priminf, compl0 / sc));
    const double u_dual = dualinf / df, u_compl = compl0 / df;
    if (E0 <= P.tol && u_dual <= 1.0 && priminf <= 1e-4 && u_compl <= 1e-4) { status = kSolveSucceeded; return false; }
    last_obj = curr_obj; curr_obj = f_cur;
    const bool acc = E0 <= 1e-6 && u_dual <= 1e10 && priminf <= 1e-2 && u_compl <= 1e-2;
    if (acc) { if (++acceptable_counter >= 15) { status = kSolvedToAcceptableLevel; return false; } }
    else acceptable_counter = 0;
    if (xmaxabs > 1e20) { status = kDivergingIterates; return false; }
    if (iter >= P.max_iter) { status = kMaxIterExceeded; return false; }
    // IpMonotoneMuUpdate.cpp:132-232
    double Emu = dmax(dualinf / sd, dmax(priminf, dmax(sz_max - mu, mu - sz_min) / sc));
    bool done = false, tiny = fl(F_TINYFLAG);
    setfl(F_TINYFLAG, false);
    while ((Emu <= 10.0 * mu || tiny) && !done) {
      const double nm = dmax(dmin(0.2 * mu, mu * sqrt(mu)), mu_min);
      const double nt = dmax(0.99, 1.0 - nm);
      const bool changed = nm != mu;
      if (!changed && tiny) { status = kSearchDirectionTooSmall; return false; }
      mu = nm; tau = nt;
      if (!changed) done = true;
      else {
        Emu = dmax(dualinf / sd, dmax(priminf, dmax(sz_max - mu, mu - sz_min) / sc));
        done = Emu > 10.0 * mu;
      }
      if (done && changed) { nf = 0; filt_rej_count = 0; setfl(F_LASTREJF, false); }   // FilterLSAcceptor::Reset
      tiny = false;
    }
    return true;
  }

  // ------------------------------------------------------------------------------------------
  // The three passes with the control logic that follows each.  In the common case a problem runs FACTOR ->
  // FORWARD -> STEP once per interior-point iteration; rare events (inertia correction, backtracking, second order
  // correction) take extra rounds without stalling the other problems.
  MPC_HD void do_factor() {
    const bool ok = factor(dw_curr, fl(F_INSOC));
    if (ok) phase = PH_FORWARD;
    else factor_failed();
  }
  MPC_HD void factor_failed() {   // IpPDPerturbationHandler.cpp:347-391
    if (dw_curr == 0.0) dw_curr = dw_last == 0.0 ? 1e-4 : dmax(1e-20, dw_last / 3.0);
    else dw_curr *= (dw_last == 0.0 || 1e5 * dw_last < dw_curr) ? 100.0 : 8.0;
    if (dw_curr > 1e20) { status = kErrorInStepComputation; phase = PH_DONE; }
  }
  MPC_HD void do_forward() {
    forward(fl(F_INSOC));
    forward_logic();
  }
  MPC_HD void forward_logic() {
    phase = PH_STEP;
    if (fl(F_LS)) return;
    alpha_du = fw_alpha_du;
    if (fl(F_INSOC)) {
      alpha = fw_alpha_pr;   // alpha_primal_soc
    } else if (fl(F_SOCDONE)) {
      // direction restored after a failed SOC: resume backtracking with the alpha set in step_logic
      if (!(alpha > alpha_min)) phase = PH_RESTO;   // line search failed (only reached with Params::resto != 0)
    } else {
      ref_theta = theta_cur;
      ref_barr = f_cur - mu * sumlog;
      ref_gbd = fw_gbd;
      setfl(F_TINYNOW, fw_tiny);
      alpha_max = fw_alpha_pr;
      alpha = alpha_max;
      n_steps = 0;
      if (fl(F_SOFT) && !fw_tiny) {
        // in the soft restoration phase only the damped full step is tried (IpBacktrackingLineSearch.cpp:426-448)
        if (++soft_count > 10) { setfl(F_HARD, true); phase = PH_RESTO; }   // max_soft_resto_iters
        else { soft_alpha = 0.0; alpha_soc = fw_alpha_du; alpha = alpha_du = dmin(alpha_max, fw_alpha_du); setfl(F_SOFTTRY, true); }
        return;
      }
      if (fw_tiny) {
        if (fl(F_TINYLAST)) setfl(F_TINYFLAG, true);
      } else {
        double am = 1e-5;   // CalculateAlphaMin IpFilterLSAcceptor.cpp:393-410
        if (ref_gbd < 0.0) {
          am = dmin(am, 1e-8 * ref_theta / (-ref_gbd));
          if (ref_theta <= theta_min) am = dmin(am, 1.0 * pow(ref_theta, 1.1) / pow(-ref_gbd, 2.3));
        }
        alpha_min = 0.05 * am;
      }
    }
  }
  MPC_HD void do_step() {
    step_pass();
    step_logic();
  }
  // FIN = false: the per-pass sweep kernels, which never see a problem in the soft restoration phase (a failed line
  // search parks the problem for the finisher); the soft branches are compiled out of them.
  template <bool FIN = true>
  MPC_HD void step_pass() {
    if (fl(F_LS)) accept_ls(!fl(F_LSKEEP));
    else step_sweep(alpha, FIN && fl(F_SOFTFIX) ? soft_alpha : alpha, alpha_du, dw_curr);
  }
  template <bool FIN = true>
  MPC_HD void step_logic() {
    if (fl(F_LS)) {
      if (!fl(F_LSKEEP) && dlam_max <= 1000.0) { setfl(F_LSKEEP, true); return; }   // estimate within constr_mult_init_max: store it
      setfl(F_LS, false); setfl(F_LSKEEP, false);
    } else {
      // ---- line search decision on the trial point (IpBacktrackingLineSearch.cpp:637-797)
      const double tbarr = tr_f - mu * tr_sumlog;
      const bool tiny_now = fl(F_TINYNOW), in_soc = fl(F_INSOC);
      bool acc, soft_step = false;
      if (FIN && fl(F_SOFTFIX)) {   // second pass of an 'S' step: multipliers recomputed, accepted as it is
        setfl(F_SOFTFIX, false);
        acc = soft_step = true;
      } else if (FIN && fl(F_SOFTTRY)) {
        // TrySoftRestoStep (IpBacktrackingLineSearch.cpp:1043-1140): the damped full step (same step size for x,
        // lambda and z) is taken if the original acceptance test passes with alpha_test = 0 ('S': back to the regular
        // algorithm) or if it reduces the primal-dual system error by the factor 1 - 1e-4 ('s': stay in the soft
        // phase); the filter is not augmented.  Otherwise the restoration step follows.
        setfl(F_SOFTTRY, false);
        const bool orig = check_accept(0.0, tr_theta, tbarr);
        if (!orig && !(pd_error(cur ^ 1) <= (1.0 - 1e-4) * pd_error(cur))) {
          setfl(F_HARD, true);
          phase = PH_RESTO;
          return;
        }
        setfl(F_SOFT, !orig); setfl(F_FILTDONE, false);
        if (orig) {
          // 'S': back to the regular algorithm.  Ipopt then repeats the dual step with the primal step size its line
          // search variable holds (the last failed one on the first attempt, 0 inside the soft phase) for lambda and
          // the full fraction-to-the-boundary step for z (IpBacktrackingLineSearch.cpp:595-603): one more STEP pass.
          soft_count = 0;
          setfl(F_SOFTFIX, true);
          alpha_du = alpha_soc;
          phase = PH_STEP;
          return;
        }
        acc = soft_step = true;
      } else if (tiny_now) acc = true;
      else {
        if (!in_soc) alpha_test = alpha;
        acc = check_accept(alpha_test, tr_theta, tbarr);
      }
      if (!acc) {
        if (in_soc) {
          ++soc_count;
          if (soc_count < 4 && tr_theta <= 0.99 * theta_soc_old) {   // another correction
            theta_soc_old = tr_theta;
            alpha_soc = alpha;
            build_csoc(alpha_soc);
            phase = PH_FACTOR;
          } else {   // give up: restore the Newton direction, continue backtracking
            setfl(F_INSOC, false); setfl(F_SOCDONE, true);
            alpha = 0.5 * alpha_max; n_steps = 1;
            // DS / DU hold the corrected direction: the Newton direction is recomputed first, also when the halved step
            // is already below alpha_min (the soft restoration step that follows the failure needs it; forward_logic)
            if (alpha > alpha_min || P.resto) phase = PH_FACTOR;
            else line_search_failed();
          }
        } else if (!fl(F_SOCDONE) && alpha == alpha_max && ref_theta <= tr_theta) {   // start SOC (IpFilterLSAcceptor.cpp:473-587)
          setfl(F_INSOC, true); soc_count = 0;
          theta_soc_old = tr_theta;
          alpha_soc = alpha;
          init_csoc();
          build_csoc(alpha_soc);
          phase = PH_FACTOR;
        } else {
          alpha *= 0.5; ++n_steps;
          if (!(alpha > alpha_min)) line_search_failed();
        }
        return;
      }
      if (!tiny_now && !soft_step && (!is_ftype(alpha_test) || !armijo(alpha_test, tbarr)))
        filter_add(ref_barr - 1e-8 * ref_theta, (1.0 - 1e-5) * ref_theta);
      // ---- accept: the trial copy becomes the current iterate
      cur ^= 1;
      f_cur = tr_f; theta_cur = tr_theta; priminf = tr_priminf; sumlog = tr_sumlog;
      setfl(F_TINYLAST, tiny_now && !fl(F_RESTO) ? dlam_max < 1e-2 : false);
      setfl(F_INSOC, false); setfl(F_SOCDONE, false); setfl(F_RESTO, false);
      ++iter;
    }
    phase = top_of_loop() ? PH_FACTOR : PH_DONE;
    if (phase == PH_FACTOR) {   // PDPerturbationHandler::ConsiderNewSystem
      if (dw_curr > 0.0) dw_last = dw_curr;
      dw_curr = 0.0;
    }
  }
  // ------------------------------------------------------------------------------------------
  // Failed line search (alpha <= alpha_min).  Ipopt switches to its restoration phase here
  // (IpBacktrackingLineSearch.cpp:531-585); at an almost feasible point it gives up instead (:548-563).
  MPC_HD void line_search_failed() {
    if (P.resto) phase = PH_RESTO;
    else { status = kRestorationFailed; phase = PH_DONE; }
  }
  // Primal-dual system error of the iterate copy `buf` for the barrier parameter mu
  // (IpIpoptCalculatedQuantities.cpp:2835-2884): 1-norms of grad_x L, c and the relaxed complementarity, each divided
  // by its number of entries.  Rare path: its own sweep over the workspace.
  MPC_HD double pd_error(int buf) const {
    const double gv2 = 2.0 * P.w_v * df, ge2 = 2.0 * P.w_epsi * df, gc2 = 2.0 * P.w_cte * df;
    const int b = kX * buf;
    double dual = 0.0, prim = 0.0, cmpl = 0.0;
    double sn[6], ln[6], un0 = 0.0, un1 = 0.0;
    {
      const int r = rec(M) + b;
#pragma unroll
      for (int k = 0; k < 6; ++k) { sn[k] = w(r + xS + k); ln[k] = w(r + xLAM + k); }
      dual += fabs(ln[0]) + fabs(ln[1]) + fabs(ln[2]) + fabs(gv2 * (sn[3] - P.ref_v) + ln[3]) + fabs(gc2 * sn[4] + ln[4]) +
              fabs(ge2 * sn[5] + ln[5]);
    }
    for (int t = M - 1; t >= 0; --t) {
      const int r = rec(t) + b;
      double s[6], lam[6], u[2], um0 = 0.0, um1 = 0.0, sp, cp, se, ce, p0, p1, p2, p3, c[6], at[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { s[k] = w(r + xS + k); lam[k] = w(r + xLAM + k); }
      u[0] = w(r + xU); u[1] = w(r + xU + 1);
      if (t > 0) { um0 = w(r - kRec + xU); um1 = w(r - kRec + xU + 1); }
      const double zl0 = w(r + xZL), zl1 = w(r + xZL + 1), zu0 = w(r + xZU), zu1 = w(r + xZU + 1);
      trig_of(r, s, sp, cp, se, ce);
      poly_eval(cf, s[0], p0, p1, p2, p3);
      residual(s, u, sn, sp, cp, se, p0, atan(p1), c);
#pragma unroll
      for (int k = 0; k < 6; ++k) prim += fabs(c[k]);
      const Lin A = make_lin(P, s[3], u[0], sp, cp, se, ce, p1, p2);
      applyAT6(A, ln, at);
      dual += fabs(lam[0] - at[0]) + fabs(lam[1] - at[1]) + fabs(lam[2] - at[2]) + fabs(gv2 * (s[3] - P.ref_v) + lam[3] - at[3]) +
              fabs(gc2 * s[4] + lam[4]) + fabs(ge2 * s[5] + lam[5] - at[5]);
      dual += fabs(grad_u(0, t, u[0], um0, un0) - A.beta * (ln[2] + ln[5]) - zl0 + zu0) +
              fabs(grad_u(1, t, u[1], um1, un1) - P.dt * ln[3] - zl1 + zu1);
      cmpl += fabs(safe_slack(u[0] - P.xl[0], mu, zl0, P.xl[0]) * zl0 - mu) + fabs(safe_slack(P.xu[0] - u[0], mu, zu0, P.xu[0]) * zu0 - mu) +
              fabs(safe_slack(u[1] - P.xl[1], mu, zl1, P.xl[1]) * zl1 - mu) + fabs(safe_slack(P.xu[1] - u[1], mu, zu1, P.xu[1]) * zu1 - mu);
#pragma unroll
      for (int k = 0; k < 6; ++k) { sn[k] = s[k]; ln[k] = lam[k]; }
      un0 = u[0]; un1 = u[1];
    }
    return dual / (8.0 * N - 2.0) + prim / (6.0 * N) + cmpl / (4.0 * M);
  }
  // What the finisher does with a problem whose line search failed (phase PH_RESTO).
  //   Params::resto == 2: Ipopt's soft restoration phase first (IpBacktrackingLineSearch.cpp:498-530): the current point
  //   enters the filter and the damped full Newton step is tried (a STEP pass with the flag F_SOFTTRY, decided in
  //   step_logic); while such steps are accepted on the primal-dual error alone the problem stays in the soft phase
  //   (F_SOFT, at most 10 iterations in a row, forward_logic).  A rejected soft step comes back here with F_HARD.
  //   Otherwise, and always for Params::resto == 1: the restoration step do_resto().
  MPC_HD void resto_entry() {
    const bool first = !fl(F_HARD);   // straight from a failed backtracking line search with the Newton direction
    if (!fl(F_FILTDONE) && !fl(F_SOFT))
      filter_add(ref_barr - 1e-8 * ref_theta, (1.0 - 1e-5) * ref_theta);   // FilterLSAcceptor::PrepareRestoPhaseStart
    if (P.resto >= 2 && first && !fl(F_SOFT)) {
      setfl(F_FILTDONE, true);
      setfl(F_SOFTTRY, true);
      setfl(F_SOCDONE, false);
      soft_alpha = alpha; alpha_soc = alpha_du;   // step sizes of the failed line search, for an 'S' step
      alpha = alpha_du = dmin(alpha_max, alpha_du);
      phase = PH_STEP;
      return;
    }
    setfl(F_HARD, false); setfl(F_FILTDONE, false); setfl(F_SOFT, false);
    soft_count = 0;
    if (!(ref_theta > 1e-2 * P.tol)) { status = kRestorationFailed; phase = PH_DONE; return; }   // :548-563
    do_resto();
  }
  MPC_HD void do_resto() {
    const int b = kX * cur;
    const double keep = 1.0 - kRestoBeta;
    // dual step length of z := z + a (mu / slack - z), fraction-to-the-boundary (IpRestoMinC_1Nrm.cpp:283-301)
    double a_du = 1.0;
    for (int t = 0; t < M; ++t) {
      const int r = rec(t) + b;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double u = w(r + xU + j), zl = w(r + xZL + j), zu = w(r + xZU + j);
        const double dl = mu / safe_slack(u - P.xl[j], mu, zl, P.xl[j]) - zl, du = mu / safe_slack(P.xu[j] - u, mu, zu, P.xu[j]) - zu;
        if (dl < 0.0) a_du = dmin(a_du, -tau / dl * zl);
        if (du < 0.0) a_du = dmin(a_du, -tau / du * zu);
      }
    }
    double so[6], sn[6], zmax = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { so[k] = sn[k] = w(rec(0) + b + xS + k); }
    for (int t = 0; t < M; ++t) {
      const int r = rec(t) + b, rn = r + kRec;
      double u[2], son[6], co[6], cn[6], sp, cp, se, ce, p0, p1, p2, p3;
      u[0] = w(r + xU); u[1] = w(r + xU + 1);
#pragma unroll
      for (int k = 0; k < 6; ++k) son[k] = w(rn + xS + k);
      // defect of the rows of time t+1 at x_R
      trig_of(r, so, sp, cp, se, ce);
      poly_eval(cf, so[0], p0, p1, p2, p3);
      residual(so, u, son, sp, cp, se, p0, atan(p1), co);
      // model step from the new s_t
      const double zero[6] = {0, 0, 0, 0, 0, 0};
      sincos(sn[2], &sp, &cp);
      sincos(sn[5], &se, &ce);
      poly_eval(cf, sn[0], p0, p1, p2, p3);
      const double psides = atan(p1);
      residual(sn, u, zero, sp, cp, se, p0, psides, cn);
#if MPC_STORE_TRIG
      w(r + xTR) = sp; w(r + xTR + 1) = cp; w(r + xTR + 2) = se; w(r + xTR + 3) = ce;
#endif
      store_psides(r, psides);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        so[k] = son[k];
        sn[k] = keep * co[k] - cn[k];
        w(rn + xS + k) = sn[k];
        w(r + xLAM + k) = 0.0;
        w(rec(t) + oDS + k) = 0.0;
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double zl = w(r + xZL + j), zu = w(r + xZU + j);
        zl += a_du * (mu / safe_slack(u[j] - P.xl[j], mu, zl, P.xl[j]) - zl);
        zu += a_du * (mu / safe_slack(P.xu[j] - u[j], mu, zu, P.xu[j]) - zu);
        w(r + xZL + j) = zl; w(r + xZU + j) = zu;
        zmax = dmax(zmax, dmax(zl, zu));
        w(rec(t) + oDU + j) = 0.0;
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { w(rec(M) + b + xLAM + k) = 0.0; w(rec(M) + oDS + k) = 0.0; }
    if (zmax > 1000.0)
      for (int t = 0; t < M; ++t) {
        const int r = rec(t) + b;
        w(r + xZL) = 1.0; w(r + xZL + 1) = 1.0; w(r + xZU) = 1.0; w(r + xZU + 1) = 1.0;
      }
    setfl(F_INSOC, false); setfl(F_SOCDONE, false);
    setfl(F_TINYNOW, true); setfl(F_RESTO, true);   // zero-length step, accepted without a filter test
    alpha = 0.0; alpha_du = 0.0;
    phase = PH_STEP;
  }
  MPC_HD void init_csoc() {
    for (int t = 0; t < M; ++t) {
      double c[6];
      residual_at(cur, t, c);
      for (int k = 0; k < 6; ++k) w(rec(t + 1) + oCSOC + k) = c[k];
    }
  }
  MPC_HD void build_csoc(double a) {   // c_soc = c(trial) + alpha_soc * c_soc
    for (int t = 0; t < M; ++t) {
      double c[6];
      residual_at(cur ^ 1, t, c);
      for (int k = 0; k < 6; ++k) w(rec(t + 1) + oCSOC + k) = c[k] + a * w(rec(t + 1) + oCSOC + k);
    }
  }

  // ------------------------------------------------------------------------------------------
  // Per-pass entry points for the per-pass kernels: only the scalars a sweep needs are loaded before it, the
  // ones its control logic needs are loaded after it (so they are not live across the sweep), and only what may
  // have changed is stored.  cf must be set by the caller.  The caller has checked load_phase().
#define LDD(idx, name) name = w(idx)
#define LDI(idx, name) name = (int)w(idx)
#define STD_(idx, name) w(idx) = name
#define STI(idx, name) w(idx) = (double)name
  MPC_HD void kernel_factor() {
    LDI(iFLAGS, flags); LDI(iCUR, cur); LDD(dDF, df); LDD(dMU, mu); LDD(dDWC, dw_curr);
    phase = PH_FACTOR;
    const bool ok = factor_spec(dw_curr, fl(F_INSOC));
    if (ok) { w(iPHASE) = (double)PH_FORWARD; return; }
    LDD(dDWL, dw_last); LDI(iSTATUS, status);
    factor_failed();
    STD_(dDWC, dw_curr); STI(iSTATUS, status); STI(iPHASE, phase);
  }
  MPC_HD void kernel_forward() {
    LDI(iFLAGS, flags); LDI(iCUR, cur); LDD(dDF, df); LDD(dMU, mu); LDD(dTAU, tau);
    phase = PH_FORWARD;
    forward_spec(fl(F_INSOC));
    LDD(dTH, theta_cur); LDD(dF, f_cur); LDD(dSLOG, sumlog); LDD(dTHMIN, theta_min);
    LDD(dALPHA, alpha); LDD(dAMAX, alpha_max); LDD(dAMIN, alpha_min); LDD(dRTH, ref_theta); LDD(dRBARR, ref_barr);
    LDD(dRGBD, ref_gbd); LDD(dADU, alpha_du); LDI(iNSTEPS, n_steps);
    forward_logic();
    STD_(dADU, alpha_du); STD_(dALPHA, alpha); STD_(dAMAX, alpha_max); STD_(dAMIN, alpha_min); STD_(dRTH, ref_theta);
    STD_(dRBARR, ref_barr); STD_(dRGBD, ref_gbd); STI(iNSTEPS, n_steps); STI(iFLAGS, flags); STI(iPHASE, phase);
  }
  MPC_HD void kernel_step() {
    LDI(iFLAGS, flags); LDI(iCUR, cur); LDD(dDF, df); LDD(dMU, mu); LDD(dALPHA, alpha); LDD(dADU, alpha_du); LDD(dDWC, dw_curr);
    phase = PH_STEP;
    step_pass<false>();
    LDD(dRTH, ref_theta); LDD(dRBARR, ref_barr); LDD(dRGBD, ref_gbd); LDD(dTHMAX, theta_max); LDD(dTHMIN, theta_min);
    LDD(dATEST, alpha_test); LDD(dAMAX, alpha_max); LDD(dAMIN, alpha_min); LDD(dTHSOC, theta_soc_old); LDD(dASOC, alpha_soc);
    LDD(dTAU, tau); LDD(dMUMIN, mu_min); LDD(dCOBJ, curr_obj); LDD(dLOBJ, last_obj); LDD(dDWL, dw_last);
    LDD(dF, f_cur); LDD(dTH, theta_cur); LDD(dPINF, priminf); LDD(dSLOG, sumlog);
    LDI(iSOCCNT, soc_count); LDI(iNSTEPS, n_steps); LDI(iNF, nf); LDI(iFRCNT, filt_rej_count); LDI(iSTATUS, status); LDI(iACCCNT, acceptable_counter);
    LDI(iITER, iter);
    step_logic<false>();
    STD_(dTHMAX, theta_max); STD_(dTHMIN, theta_min); STD_(dATEST, alpha_test); STD_(dALPHA, alpha); STD_(dTHSOC, theta_soc_old);
    STD_(dASOC, alpha_soc); STD_(dF, f_cur); STD_(dTH, theta_cur); STD_(dPINF, priminf); STD_(dSLOG, sumlog); STD_(dMU, mu);
    STD_(dTAU, tau); STD_(dCOBJ, curr_obj); STD_(dLOBJ, last_obj); STD_(dDWC, dw_curr); STD_(dDWL, dw_last);
    STI(iSOCCNT, soc_count); STI(iNSTEPS, n_steps); STI(iNF, nf); STI(iFRCNT, filt_rej_count); STI(iACCCNT, acceptable_counter); STI(iITER, iter);
    STI(iCUR, cur); STI(iFLAGS, flags); STI(iSTATUS, status); STI(iPHASE, phase);
  }
#if MPC_FUSE_FACTOR
  // STEP pass with the next iteration's Riccati factorisation riding on it.  If the trial point is accepted and the solve
  // goes on, the factors of the new system are already in the workspace: the problem goes straight to its forward sweep
  // (or, on wrong inertia, to the separate factor sweep with the first delta_w of the ladder -- what kernel_factor would
  // have found).  Anything else (rejected trial point, second-order correction, multiplier initialisation) is
  // kernel_step's business as before.
  MPC_HD void kernel_stepfactor() {
    LDI(iFLAGS, flags); LDI(iCUR, cur); LDD(dDF, df); LDD(dMU, mu); LDD(dALPHA, alpha); LDD(dADU, alpha_du); LDD(dDWC, dw_curr);
    phase = PH_STEP;
    const bool spec = !fl(F_LS);
    bool ric_ok = false;
    if (spec) ric_ok = step_sweep<true>(alpha, alpha, alpha_du, dw_curr);
    else accept_ls(!fl(F_LSKEEP));
    LDD(dRTH, ref_theta); LDD(dRBARR, ref_barr); LDD(dRGBD, ref_gbd); LDD(dTHMAX, theta_max); LDD(dTHMIN, theta_min);
    LDD(dATEST, alpha_test); LDD(dAMAX, alpha_max); LDD(dAMIN, alpha_min); LDD(dTHSOC, theta_soc_old); LDD(dASOC, alpha_soc);
    LDD(dTAU, tau); LDD(dMUMIN, mu_min); LDD(dCOBJ, curr_obj); LDD(dLOBJ, last_obj); LDD(dDWL, dw_last);
    LDD(dF, f_cur); LDD(dTH, theta_cur); LDD(dPINF, priminf); LDD(dSLOG, sumlog);
    LDI(iSOCCNT, soc_count); LDI(iNSTEPS, n_steps); LDI(iNF, nf); LDI(iFRCNT, filt_rej_count); LDI(iSTATUS, status); LDI(iACCCNT, acceptable_counter);
    LDI(iITER, iter);
    const int cur_before = cur;
    step_logic<false>();
    if (spec && phase == PH_FACTOR && cur != cur_before) {   // accepted, not finished: dw_curr = 0, no correction pending
      if (ric_ok) phase = PH_FORWARD;
      else factor_failed();   // may end the solve (delta_w ladder exhausted); the caller looks at phase
    }
    STD_(dTHMAX, theta_max); STD_(dTHMIN, theta_min); STD_(dATEST, alpha_test); STD_(dALPHA, alpha); STD_(dTHSOC, theta_soc_old);
    STD_(dASOC, alpha_soc); STD_(dF, f_cur); STD_(dTH, theta_cur); STD_(dPINF, priminf); STD_(dSLOG, sumlog); STD_(dMU, mu);
    STD_(dTAU, tau); STD_(dCOBJ, curr_obj); STD_(dLOBJ, last_obj); STD_(dDWC, dw_curr); STD_(dDWL, dw_last);
    STI(iSOCCNT, soc_count); STI(iNSTEPS, n_steps); STI(iNF, nf); STI(iFRCNT, filt_rej_count); STI(iACCCNT, acceptable_counter); STI(iITER, iter);
    STI(iCUR, cur); STI(iFLAGS, flags); STI(iSTATUS, status); STI(iPHASE, phase);
  }
#endif
#undef LDD
#undef LDI
#undef STD_
#undef STI

  MPC_HD void trip() {
    if (phase == PH_RESTO) resto_entry();
    if (phase == PH_FACTOR) do_factor();
    if (phase == PH_FORWARD) do_forward();
    if (phase == PH_STEP) do_step();
  }

  // ------------------------------------------------------------------------------------------
  // finalize: honor_original_bounds (IpOrigIpoptNLP.cpp:875-883), unscaled objective, MPC.cpp:253-256.
  // traj (optional): full variable vector in the reference layout (MPC.cpp:36-43), element i at traj[i*tstride].
  MPC_HD void finish(Result& R, double* traj, size_t tstride) {
    const int bX = kX * cur;
    R.status = status; R.iters = iter; R.obj = f_cur / df;
#pragma unroll
    for (int k = 0; k < 6; ++k) R.out8[k] = w(rec(1) + bX + xS + k);
    R.out8[6] = dclamp(w(rec(0) + bX + xU + 0), -P.ob[0], P.ob[0]);
    R.out8[7] = dclamp(w(rec(0) + bX + xU + 1), -P.ob[1], P.ob[1]);
    if (traj) {
      for (int t = 0; t < N; ++t)
#pragma unroll
        for (int k = 0; k < 6; ++k) traj[(size_t)(k * N + t) * tstride] = w(rec(t) + bX + xS + k);
      for (int t = 0; t < M; ++t) {
        traj[(size_t)(6 * N + t) * tstride] = dclamp(w(rec(t) + bX + xU), -P.ob[0], P.ob[0]);
        traj[(size_t)(6 * N + M + t) * tstride] = dclamp(w(rec(t) + bX + xU + 1), -P.ob[1], P.ob[1]);
      }
    }
  }
};

#undef PS

}  // namespace b200mpc
