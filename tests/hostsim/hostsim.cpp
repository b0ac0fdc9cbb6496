// TEST INFRASTRUCTURE ONLY: compiles the solver core (udacitympc_b200/csrc/mpc_core.cuh) for the CPU so the
// algorithm can be checked against the oracle in the GPU-less build container (pytest -m "not gpu").
// It is NOT part of libb200mpc.so and is never used by the product path.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../udacitympc_b200/csrc/mpc_core.cuh"

using namespace b200mpc;

extern "C" {

struct hostsim_params {
  int N;
  double dt, Lf, ref_v, w_cte, w_epsi, w_v, w_delta, w_a, w_ddelta, w_da, delta_max, a_max, tol;
  int max_iter;
};

// trace rows of 8: iter, mu, alpha_pr, alpha_du, dw, f, theta, phase-trips
int hostsim_solve(const hostsim_params* hp, const double* state6, const double* coeffs, int ncoef, double* x_out,
                  double* out8, double* obj, int* iters, double* lam_out, double* trace, int trace_cap, int* trace_rows) {
  Params P;
  P.N = hp->N; P.dt = hp->dt; P.Lf = hp->Lf; P.ref_v = hp->ref_v;
  P.w_cte = hp->w_cte; P.w_epsi = hp->w_epsi; P.w_v = hp->w_v; P.w_delta = hp->w_delta; P.w_a = hp->w_a;
  P.w_ddelta = hp->w_ddelta; P.w_da = hp->w_da; P.delta_max = hp->delta_max; P.a_max = hp->a_max; P.tol = hp->tol;
  P.max_iter = hp->max_iter;
  Layout L(P.N);
  std::vector<double> ws((size_t)L.total, 0.0);
  Solver<1> S(P, ws.data());
  S.init(state6, coeffs, ncoef);
  int rows = 0, trips = 0, last_iter = -1;
  while (S.phase != PH_DONE && trips < 100000) {
    S.trip();
    ++trips;
    if (trace && S.iter != last_iter && rows < trace_cap) {
      double* r = trace + 8 * rows++;
      r[0] = S.iter; r[1] = S.mu; r[2] = S.alpha; r[3] = S.alpha_du; r[4] = S.dw_curr; r[5] = S.f_cur / S.df;
      r[6] = S.theta_cur; r[7] = trips;
      last_iter = S.iter;
    }
  }
  Result R;
  S.finish(R, x_out, 1);
  for (int i = 0; i < 8; ++i) out8[i] = R.out8[i];
  *obj = R.obj; *iters = R.iters;
  if (lam_out) for (int t = 0; t < P.N; ++t) for (int k = 0; k < 6; ++k) lam_out[k * P.N + t] = ws[L.LAM + 6 * t + k] / S.df;
  if (trace_rows) *trace_rows = rows;
  return R.status;
}
}
