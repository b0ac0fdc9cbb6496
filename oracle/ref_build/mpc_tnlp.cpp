// TEST INFRASTRUCTURE ONLY (oracle/_ref) -- never linked into or called by the product path.
//
// Drives the reference's OWN prebuilt Ipopt 3.12.7 + MUMPS 4.10.0 binaries
//   /root/reference/Ipopt-3.12.7/Ipopt/src/Interfaces/.libs/libipopt.so.1.10.7
//   /root/reference/Ipopt-3.12.7/ThirdParty/Mumps/.libs/libcoinmumps.so.1.6.0
// on the NLP that /root/reference/mpc_to_line/solution/MPC.cpp builds:
//   * variables / constraint layout        MPC.cpp:36-43, 161-163
//   * start point (zeros + initial state)  MPC.cpp:167-177
//   * bounds (+-1e19 / +-0.436332 / +-1)   MPC.cpp:185-203
//   * constraint bounds (rows 0 = x0)      MPC.cpp:208-226
//   * objective                            MPC.cpp:57-76
//   * constraints                          MPC.cpp:88-138
//   * Ipopt options: defaults + print_level 0      MPC.cpp:232-235
//   * returned 8-vector                    MPC.cpp:253-256
// The reference obtains derivatives from CppAD (MPC.cpp:241-243), an un-vendored, un-pinned
// dependency that is absent from /root/reference and this image.  AD derivatives are exact up
// to rounding, so this TNLP supplies the closed-form Jacobian / Lagrangian Hessian instead -- the
// same pattern the reference itself uses in
//   mpc_to_line/src/n_steps_kinematics_model/apply_ipopt_n_steps_kinematics_model.cpp:40-413.
// Ipopt's `derivative_test second-order` is run over them by tests/test_ref_oracle.py.
//
// Cost weights: MPC.cpp:57-76 adds the seven squared terms with weight 1 (no factor in front of CppAD::pow).  The
// weights w[7] = (cte, epsi, v, delta, a, ddelta, da) below multiply those terms (ref_mpc_solve_w); ref_mpc_solve passes
// 1.0 for each, and 1.0 * t is t bit for bit, so the as-shipped problem is unchanged.
//
// Generalisation: MPC.cpp:117-118 hard-codes a degree-1 reference (f0 = c0 + c1*x0,
// psides0 = atan(c1)).  For BASELINE configs 4/5 the natural degree-d form
// f0 = p(x0), psides0 = atan(p'(x0)) is used; it is bit-identical to the shipped code when
// ncoef == 2 (the degree-1 branch below evaluates exactly c0 + c1*x0 and atan(c1)).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "IpIpoptApplication.hpp"
#include "IpIpoptCalculatedQuantities.hpp"
#include "IpIpoptData.hpp"
#include "IpTNLP.hpp"

using namespace Ipopt;

namespace {

struct Poly {
  // value and first three derivatives of p(x) = sum c_i x^i
  static void eval(const double* c, int nc, double x, double& p, double& d1, double& d2, double& d3) {
    if (nc == 2) {  // exactly MPC.cpp:117-118
      p = c[0] + c[1] * x; d1 = c[1]; d2 = 0.0; d3 = 0.0; return;
    }
    p = d1 = d2 = d3 = 0.0;
    for (int i = nc - 1; i >= 0; --i) {  // Horner with derivative carries
      d3 = d3 * x + 3.0 * d2;
      d2 = d2 * x + 2.0 * d1;
      d1 = d1 * x + p;
      p = p * x + c[i];
    }
  }
};

struct TraceRow { double v[10]; };  // iter,obj,inf_pr,inf_du,mu,d_norm,regu,alpha_du,alpha_pr,ls_trials

class MpcNlp : public TNLP {
 public:
  int N; double dt, Lf, ref_v, delta_max, a_max;
  double w[7] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0};   // cte, epsi, v, delta, a, ddelta, da
  double s0[6];
  std::vector<double> coef;
  // outputs
  std::vector<double> xsol, lamsol, zL, zU;
  double obj; int status; int iters;
  std::vector<TraceRow> trace;

  int xs() const { return 0; }
  int ys() const { return N; }
  int ps() const { return 2 * N; }
  int vs() const { return 3 * N; }
  int cs() const { return 4 * N; }
  int es() const { return 5 * N; }
  int ds() const { return 6 * N; }
  int as() const { return 7 * N - 1; }
  int nvar() const { return 8 * N - 2; }
  int ncon() const { return 6 * N; }

  bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& style) override {
    n = nvar(); m = ncon();
    nnz_jac_g = 6 + 25 * (N - 1);
    nnz_h_lag = 3 * N + 2 * (N - 1) + 2 * (N - 2) + 6 * (N - 1);
    style = C_STYLE;
    return true;
  }
  bool get_bounds_info(Index n, Number* xl, Number* xu, Index m, Number* gl, Number* gu) override {
    for (int i = 0; i < ds(); ++i) { xl[i] = -1.0e19; xu[i] = 1.0e19; }
    for (int i = ds(); i < as(); ++i) { xl[i] = -delta_max; xu[i] = delta_max; }
    for (int i = as(); i < n; ++i) { xl[i] = -a_max; xu[i] = a_max; }
    for (int i = 0; i < m; ++i) gl[i] = gu[i] = 0.0;
    for (int k = 0; k < 6; ++k) gl[k * N] = gu[k * N] = s0[k];
    return true;
  }
  bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number*, Number*, Index, bool init_lambda,
                          Number*) override {
    if (!init_x || init_z || init_lambda) return false;
    for (int i = 0; i < n; ++i) x[i] = 0.0;
    for (int k = 0; k < 6; ++k) x[k * N] = s0[k];
    return true;
  }
  bool eval_f(Index, const Number* x, bool, Number& f) override {
    f = 0.0;
    for (int t = 0; t < N; ++t) {
      f += w[0] * std::pow(x[cs() + t], 2);
      f += w[1] * std::pow(x[es() + t], 2);
      f += w[2] * std::pow(x[vs() + t] - ref_v, 2);
    }
    for (int t = 0; t < N - 1; ++t) {
      f += w[3] * std::pow(x[ds() + t], 2);
      f += w[4] * std::pow(x[as() + t], 2);
    }
    for (int t = 0; t < N - 2; ++t) {
      f += w[5] * std::pow(x[ds() + t + 1] - x[ds() + t], 2);
      f += w[6] * std::pow(x[as() + t + 1] - x[as() + t], 2);
    }
    return true;
  }
  bool eval_grad_f(Index n, const Number* x, bool, Number* g) override {
    for (int i = 0; i < n; ++i) g[i] = 0.0;
    for (int t = 0; t < N; ++t) {
      g[cs() + t] = w[0] * (2.0 * x[cs() + t]);
      g[es() + t] = w[1] * (2.0 * x[es() + t]);
      g[vs() + t] = w[2] * (2.0 * (x[vs() + t] - ref_v));
    }
    for (int t = 0; t < N - 1; ++t) {
      g[ds() + t] += w[3] * (2.0 * x[ds() + t]);
      g[as() + t] += w[4] * (2.0 * x[as() + t]);
    }
    for (int t = 0; t < N - 2; ++t) {
      double dd = x[ds() + t + 1] - x[ds() + t], da = x[as() + t + 1] - x[as() + t];
      g[ds() + t + 1] += w[5] * (2.0 * dd); g[ds() + t] -= w[5] * (2.0 * dd);
      g[as() + t + 1] += w[6] * (2.0 * da); g[as() + t] -= w[6] * (2.0 * da);
    }
    return true;
  }
  bool eval_g(Index, const Number* x, bool, Index, Number* g) override {
    for (int k = 0; k < 6; ++k) g[k * N] = x[k * N];
    for (int t = 1; t < N; ++t) {
      double x1 = x[xs() + t], y1 = x[ys() + t], psi1 = x[ps() + t], v1 = x[vs() + t], cte1 = x[cs() + t],
             epsi1 = x[es() + t];
      double x0 = x[xs() + t - 1], y0 = x[ys() + t - 1], psi0 = x[ps() + t - 1], v0 = x[vs() + t - 1],
             epsi0 = x[es() + t - 1];
      double delta0 = x[ds() + t - 1], a0 = x[as() + t - 1];
      double f0, d1, d2, d3;
      Poly::eval(coef.data(), (int)coef.size(), x0, f0, d1, d2, d3);
      double psides0 = std::atan(d1);
      g[xs() + t] = x1 - (x0 + v0 * std::cos(psi0) * dt);
      g[ys() + t] = y1 - (y0 + v0 * std::sin(psi0) * dt);
      g[ps() + t] = psi1 - (psi0 + v0 * delta0 / Lf * dt);
      g[vs() + t] = v1 - (v0 + a0 * dt);
      g[cs() + t] = cte1 - ((f0 - y0) + (v0 * std::sin(epsi0) * dt));
      g[es() + t] = epsi1 - ((psi0 - psides0) + v0 * delta0 / Lf * dt);
    }
    return true;
  }
  bool eval_jac_g(Index, const Number* x, bool, Index, Index nele, Index* iRow, Index* jCol, Number* val) override {
    int k = 0;
    auto put = [&](int r, int c, double v) {
      if (val) val[k] = v; else { iRow[k] = r; jCol[k] = c; }
      ++k;
    };
    for (int q = 0; q < 6; ++q) put(q * N, q * N, 1.0);
    for (int t = 1; t < N; ++t) {
      double x0 = 0, psi0 = 0, v0 = 0, epsi0 = 0, delta0 = 0;
      double f0 = 0, d1 = 0, d2 = 0, d3 = 0;
      if (val) {
        x0 = x[xs() + t - 1]; psi0 = x[ps() + t - 1]; v0 = x[vs() + t - 1]; epsi0 = x[es() + t - 1];
        delta0 = x[ds() + t - 1];
        Poly::eval(coef.data(), (int)coef.size(), x0, f0, d1, d2, d3);
      }
      double sp = std::sin(psi0), cp = std::cos(psi0), se = std::sin(epsi0), ce = std::cos(epsi0);
      put(xs() + t, xs() + t, 1.0); put(xs() + t, xs() + t - 1, -1.0);
      put(xs() + t, ps() + t - 1, v0 * sp * dt); put(xs() + t, vs() + t - 1, -cp * dt);
      put(ys() + t, ys() + t, 1.0); put(ys() + t, ys() + t - 1, -1.0);
      put(ys() + t, ps() + t - 1, -v0 * cp * dt); put(ys() + t, vs() + t - 1, -sp * dt);
      put(ps() + t, ps() + t, 1.0); put(ps() + t, ps() + t - 1, -1.0);
      put(ps() + t, vs() + t - 1, -delta0 / Lf * dt); put(ps() + t, ds() + t - 1, -v0 / Lf * dt);
      put(vs() + t, vs() + t, 1.0); put(vs() + t, vs() + t - 1, -1.0); put(vs() + t, as() + t - 1, -dt);
      put(cs() + t, cs() + t, 1.0); put(cs() + t, xs() + t - 1, -d1); put(cs() + t, ys() + t - 1, 1.0);
      put(cs() + t, vs() + t - 1, -se * dt); put(cs() + t, es() + t - 1, -v0 * ce * dt);
      put(es() + t, es() + t, 1.0); put(es() + t, ps() + t - 1, -1.0);
      put(es() + t, xs() + t - 1, d2 / (1.0 + d1 * d1));
      put(es() + t, vs() + t - 1, -delta0 / Lf * dt); put(es() + t, ds() + t - 1, -v0 / Lf * dt);
    }
    return k == nele;
  }
  bool eval_h(Index, const Number* x, bool, Number sig, Index, const Number* lam, bool, Index nele, Index* iRow,
              Index* jCol, Number* val) override {
    int k = 0;
    auto put = [&](int r, int c, double v) {
      if (val) val[k] = v; else { iRow[k] = r; jCol[k] = c; }
      ++k;
    };
    for (int t = 0; t < N; ++t) {
      put(cs() + t, cs() + t, 2.0 * sig * w[0]); put(es() + t, es() + t, 2.0 * sig * w[1]); put(vs() + t, vs() + t, 2.0 * sig * w[2]);
    }
    for (int t = 0; t < N - 1; ++t) {
      const double nd = (t > 0 ? 1.0 : 0.0) + (t < N - 2 ? 1.0 : 0.0);
      put(ds() + t, ds() + t, 2.0 * sig * (w[3] + nd * w[5])); put(as() + t, as() + t, 2.0 * sig * (w[4] + nd * w[6]));
    }
    for (int t = 0; t < N - 2; ++t) { put(ds() + t + 1, ds() + t, -2.0 * sig * w[5]); put(as() + t + 1, as() + t, -2.0 * sig * w[6]); }
    for (int t = 1; t < N; ++t) {
      double x0 = 0, psi0 = 0, v0 = 0, epsi0 = 0, lx = 0, ly = 0, lp = 0, lc = 0, le = 0;
      double f0 = 0, d1 = 0, d2 = 0, d3 = 0;
      if (val) {
        x0 = x[xs() + t - 1]; psi0 = x[ps() + t - 1]; v0 = x[vs() + t - 1]; epsi0 = x[es() + t - 1];
        lx = lam[xs() + t]; ly = lam[ys() + t]; lp = lam[ps() + t]; lc = lam[cs() + t]; le = lam[es() + t];
        Poly::eval(coef.data(), (int)coef.size(), x0, f0, d1, d2, d3);
      }
      double sp = std::sin(psi0), cp = std::cos(psi0), se = std::sin(epsi0), ce = std::cos(epsi0);
      double q = 1.0 + d1 * d1;
      put(ps() + t - 1, ps() + t - 1, lx * v0 * cp * dt + ly * v0 * sp * dt);
      put(vs() + t - 1, ps() + t - 1, lx * sp * dt - ly * cp * dt);
      put(ds() + t - 1, vs() + t - 1, -(lp + le) * dt / Lf);
      put(es() + t - 1, es() + t - 1, lc * v0 * se * dt);
      put(es() + t - 1, vs() + t - 1, -lc * ce * dt);
      put(xs() + t - 1, xs() + t - 1, -lc * d2 + le * (d3 * q - 2.0 * d1 * d2 * d2) / (q * q));
    }
    return k == nele;
  }
  bool intermediate_callback(AlgorithmMode mode, Index iter, Number obj_value, Number inf_pr, Number inf_du,
                             Number mu, Number d_norm, Number regu, Number alpha_du, Number alpha_pr, Index ls_trials,
                             const IpoptData*, IpoptCalculatedQuantities*) override {
    TraceRow r;
    r.v[0] = iter; r.v[1] = obj_value; r.v[2] = inf_pr; r.v[3] = inf_du; r.v[4] = mu; r.v[5] = d_norm;
    r.v[6] = regu; r.v[7] = alpha_du; r.v[8] = alpha_pr; r.v[9] = ls_trials + 100.0 * (mode == RestorationPhaseMode);
    trace.push_back(r);
    return true;
  }
  void finalize_solution(SolverReturn st, Index n, const Number* x, const Number* z_L, const Number* z_U, Index m,
                         const Number*, const Number* lambda, Number obj_value, const IpoptData* ip_data,
                         IpoptCalculatedQuantities*) override {
    xsol.assign(x, x + n); zL.assign(z_L, z_L + n); zU.assign(z_U, z_U + n); lamsol.assign(lambda, lambda + m);
    obj = obj_value; status = (int)st;
    iters = ip_data ? ip_data->iter_count() : -1;
  }
};

}  // namespace

extern "C" {

// Options: `opts` is a newline separated list "name value" applied on top of Ipopt defaults + print_level 0.
// Returns Ipopt's ApplicationReturnStatus (IpReturnCodes_inc.h:16-39).  Any output pointer may be NULL.
int ref_mpc_solve_w(int N, double dt, double Lf, double ref_v, double delta_max, double a_max, const double* weights7,
                    const double* state6, const double* coeffs, int ncoef, const char* opts, double* x_out, double* out8,
                    double* obj_out, int* iters_out, double* lambda_out, double* zl_out, double* zu_out, double* trace_out,
                    int trace_cap, int* trace_rows);

int ref_mpc_solve(int N, double dt, double Lf, double ref_v, double delta_max, double a_max, const double* state6,
                  const double* coeffs, int ncoef, const char* opts, double* x_out /*8N-2*/, double* out8,
                  double* obj_out, int* iters_out, double* lambda_out /*6N*/, double* zl_out /*8N-2*/,
                  double* zu_out /*8N-2*/, double* trace_out /*trace_cap x 10*/, int trace_cap, int* trace_rows) {
  return ref_mpc_solve_w(N, dt, Lf, ref_v, delta_max, a_max, nullptr, state6, coeffs, ncoef, opts, x_out, out8, obj_out, iters_out,
                         lambda_out, zl_out, zu_out, trace_out, trace_cap, trace_rows);
}

// weights7 = (w_cte, w_epsi, w_v, w_delta, w_a, w_ddelta, w_da) multiplying the seven cost terms of MPC.cpp:57-76; NULL = 1.
int ref_mpc_solve_w(int N, double dt, double Lf, double ref_v, double delta_max, double a_max, const double* weights7,
                    const double* state6, const double* coeffs, int ncoef, const char* opts, double* x_out, double* out8,
                    double* obj_out, int* iters_out, double* lambda_out, double* zl_out, double* zu_out, double* trace_out,
                    int trace_cap, int* trace_rows) {
  SmartPtr<MpcNlp> nlp = new MpcNlp();
  if (weights7) std::memcpy(nlp->w, weights7, sizeof(double) * 7);
  nlp->N = N; nlp->dt = dt; nlp->Lf = Lf; nlp->ref_v = ref_v; nlp->delta_max = delta_max; nlp->a_max = a_max;
  std::memcpy(nlp->s0, state6, sizeof(double) * 6);
  nlp->coef.assign(coeffs, coeffs + ncoef);
  nlp->obj = NAN; nlp->status = -999; nlp->iters = -1;

  // A fresh IpoptApplication per solve, as CppAD::ipopt::solve does for every MPC::Solve call.
  SmartPtr<IpoptApplication> app = IpoptApplicationFactory();
  app->Options()->SetIntegerValue("print_level", 0);  // MPC.cpp:233
  app->Options()->SetStringValue("sb", "yes");
  if (opts) {
    std::string all(opts);
    size_t pos = 0;
    while (pos < all.size()) {
      size_t nl = all.find('\n', pos);
      if (nl == std::string::npos) nl = all.size();
      std::string line = all.substr(pos, nl - pos);
      pos = nl + 1;
      size_t sp = line.find(' ');
      if (sp == std::string::npos) continue;
      std::string key = line.substr(0, sp), val = line.substr(sp + 1);
      char* endp = nullptr;
      double num = std::strtod(val.c_str(), &endp);
      bool is_num = endp && *endp == '\0' && !val.empty();
      bool ok = false;
      if (is_num && val.find_first_of(".eE") == std::string::npos) ok = app->Options()->SetIntegerValue(key, (int)num, true, true);
      if (!ok && is_num) ok = app->Options()->SetNumericValue(key, num, true, true);
      if (!ok) ok = app->Options()->SetStringValue(key, val, true, true);
      if (!ok) std::fprintf(stderr, "ref_mpc_solve: could not set option '%s'\n", line.c_str());
    }
  }
  ApplicationReturnStatus st = app->Initialize("");  // "" => do NOT read ./ipopt.opt (hidden input, IpIpoptApplication.cpp:573-581)
  if (st != Solve_Succeeded) return (int)st;
  st = app->OptimizeTNLP(nlp);

  const int n = nlp->nvar(), m = nlp->ncon();
  if ((int)nlp->xsol.size() == n) {
    if (x_out) std::memcpy(x_out, nlp->xsol.data(), sizeof(double) * n);
    if (out8) {  // MPC.cpp:253-256
      const double* x = nlp->xsol.data();
      out8[0] = x[nlp->xs() + 1]; out8[1] = x[nlp->ys() + 1]; out8[2] = x[nlp->ps() + 1]; out8[3] = x[nlp->vs() + 1];
      out8[4] = x[nlp->cs() + 1]; out8[5] = x[nlp->es() + 1]; out8[6] = x[nlp->ds()]; out8[7] = x[nlp->as()];
    }
    if (lambda_out) std::memcpy(lambda_out, nlp->lamsol.data(), sizeof(double) * m);
    if (zl_out) std::memcpy(zl_out, nlp->zL.data(), sizeof(double) * n);
    if (zu_out) std::memcpy(zu_out, nlp->zU.data(), sizeof(double) * n);
  }
  if (obj_out) *obj_out = nlp->obj;
  if (iters_out) *iters_out = nlp->iters;
  if (trace_rows) *trace_rows = (int)nlp->trace.size();
  if (trace_out) {
    int rows = (int)nlp->trace.size() < trace_cap ? (int)nlp->trace.size() : trace_cap;
    for (int i = 0; i < rows; ++i) std::memcpy(trace_out + 10 * i, nlp->trace[i].v, sizeof(double) * 10);
  }
  return (int)st;
}

// Point evaluation of the restated problem functions (used to cross-check the C oracle's and the
// CUDA path's derivative code).  jac/hess are returned as dense row-major (m x n) / (n x n, lower).
int ref_mpc_eval(int N, double dt, double Lf, double ref_v, const double* coeffs, int ncoef, const double* x,
                 const double* lambda, double sigma, double* f, double* grad, double* g, double* jac_dense,
                 double* hess_dense_lower) {
  SmartPtr<MpcNlp> holder = new MpcNlp();
  MpcNlp& nlp = *holder;
  nlp.N = N; nlp.dt = dt; nlp.Lf = Lf; nlp.ref_v = ref_v; nlp.delta_max = 0.436332; nlp.a_max = 1.0;
  nlp.coef.assign(coeffs, coeffs + ncoef);
  Index n, m, nj, nh; TNLP::IndexStyleEnum st;
  nlp.get_nlp_info(n, m, nj, nh, st);
  if (f) nlp.eval_f(n, x, true, *f);
  if (grad) nlp.eval_grad_f(n, x, true, grad);
  if (g) nlp.eval_g(n, x, true, m, g);
  if (jac_dense) {
    std::vector<Index> r(nj), c(nj); std::vector<double> v(nj);
    nlp.eval_jac_g(n, nullptr, true, m, nj, r.data(), c.data(), nullptr);
    nlp.eval_jac_g(n, x, true, m, nj, nullptr, nullptr, v.data());
    std::fill(jac_dense, jac_dense + (size_t)n * m, 0.0);
    for (int k = 0; k < nj; ++k) jac_dense[(size_t)r[k] * n + c[k]] += v[k];
  }
  if (hess_dense_lower) {
    std::vector<Index> r(nh), c(nh); std::vector<double> v(nh);
    nlp.eval_h(n, nullptr, true, sigma, m, nullptr, true, nh, r.data(), c.data(), nullptr);
    nlp.eval_h(n, x, true, sigma, m, lambda, true, nh, nullptr, nullptr, v.data());
    std::fill(hess_dense_lower, hess_dense_lower + (size_t)n * n, 0.0);
    for (int k = 0; k < nh; ++k) hess_dense_lower[(size_t)r[k] * n + c[k]] += v[k];
  }
  return 0;
}

}  // extern "C"
