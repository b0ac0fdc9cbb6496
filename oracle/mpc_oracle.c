/* TEST INFRASTRUCTURE ONLY -- see mpc_oracle.h.  Plain-C restatement of the reference's
 * mpc_to_line path.  Compiled with -ffp-contract=off so the arithmetic is the x86-64 reference's
 * (no FMA contraction).
 *
 * What is restated, and where it lives in /root/reference:
 *   problem definition (FG_eval, bounds, start point)   mpc_to_line/solution/MPC.cpp:45-140, 149-257
 *   Ipopt 3.12.7 algorithm, default options (Ipopt-3.12.7/Ipopt/src/Algorithm/...):
 *     iterate initialisation        IpDefaultIterateInitializer.cpp:175-344, 469-649, 651-718
 *     least-squares multipliers     IpLeastSquareMults.cpp:40-94
 *     NLP scaling / bound relaxing  IpGradientScaling.cpp:69-119, IpOrigIpoptNLP.cpp:361-372,466-482,875-883
 *     monotone mu update            IpMonotoneMuUpdate.cpp:132-232
 *     search direction + refinement IpPDSearchDirCalc.cpp:60-139, IpPDFullSpaceSolver.cpp:132-374,376-651,653-799
 *     inertia correction            IpPDPerturbationHandler.cpp:148-420
 *     backtracking line search      IpBacktrackingLineSearch.cpp:261-635, 637-797, 852-941, 1145-1200
 *     filter acceptor + SOC         IpFilterLSAcceptor.cpp:227-442, 473-587, 800-813, IpFilter.cpp:41-77
 *     accept / kappa_sigma          IpIpoptAlg.cpp:559-727, 880-951
 *     convergence test              IpOptErrorConvCheck.cpp:204-262, 265-329
 *     error measures                IpIpoptCalculatedQuantities.cpp:444-507,697-751,2672-2735,2782-2832,2949-3011,3279-3306
 *   MUMPS's sparse LDL^T (generic; not restated) -> dense Bunch-Kaufman LDL^T with inertia below.
 *   polyfit / polyeval              mpc_to_line/src/helpers.h:13-19, 24-44 (+ Eigen 3.3.3 HouseholderQR)
 *   globalKinematic                 global_kinematic_model/solution/main.cpp:36-62
 * The soft restoration phase (IpBacktrackingLineSearch.cpp:426-448, 498-530, 595-603, 1043-1140) is restated.
 * Not restated (documented gaps): restoration phase proper (IpRestoMinC_1Nrm.cpp; the port returns -2 where Ipopt enters it --
 * the product's own restoration step is checked against the reference binaries' answers instead,
 * tests/golden/resto_N*.npz), watchdog, constraint-row
 * scaling (never triggered for |Jacobian entries| <= 100).  A solve that would need them returns -2.
 */
#include "mpc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXF 64 /* filter entries */

void oracle_default_params(oracle_params* p) {
  p->N = 25; p->dt = 0.05; p->Lf = 2.67; p->ref_v = 40.0;
  p->w_cte = p->w_epsi = p->w_v = p->w_delta = p->w_a = p->w_ddelta = p->w_da = 1.0;
  p->delta_max = 0.436332; p->a_max = 1.0;
  p->tol = 1e-8; p->max_iter = 3000;
}

/* ------------------------------------------------------------------------------------------ */
/* helpers.h                                                                                   */
/* ------------------------------------------------------------------------------------------ */
double oracle_polyeval(const double* c, int nc, double x) {
  double r = 0.0;
  for (int i = 0; i < nc; ++i) r += c[i] * pow(x, i);
  return r;
}

int oracle_polyfit(const double* xs, const double* ys, int m, int order, double* out) {
  int n = order + 1;
  if (!(order >= 1 && order <= m - 1)) return -1; /* helpers.h:26 assert */
  double* A = (double*)malloc(sizeof(double) * (size_t)m * n); /* column major like Eigen */
  double* c = (double*)malloc(sizeof(double) * (size_t)m);
  double* h = (double*)malloc(sizeof(double) * (size_t)n);
  for (int i = 0; i < m; ++i) A[i] = 1.0;
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < order; ++i) A[j + (i + 1) * m] = A[j + i * m] * xs[j];
  int size = m < n ? m : n;
  for (int k = 0; k < size; ++k) { /* HouseholderQR.h:274-286 */
    int rr = m - k;
    double* col = A + k + k * m;
    double tail = 0.0;
    for (int i = 1; i < rr; ++i) tail += col[i] * col[i];
    double c0 = col[0], beta, tau;
    if (tail <= DBL_MIN) { /* Householder.h:79-84 */
      tau = 0.0; beta = c0;
      for (int i = 1; i < rr; ++i) col[i] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0.0) beta = -beta;
      for (int i = 1; i < rr; ++i) col[i] = col[i] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    h[k] = tau; col[0] = beta;
    for (int j = k + 1; j < n; ++j) { /* applyHouseholderOnTheLeft, Householder.h:113-131 */
      double* cj = A + k + j * m;
      if (rr == 1) { cj[0] *= 1.0 - tau; continue; }
      if (tau == 0.0) continue;
      double t = 0.0;
      for (int i = 1; i < rr; ++i) t += col[i] * cj[i];
      t += cj[0];
      cj[0] -= tau * t;
      for (int i = 1; i < rr; ++i) cj[i] -= tau * col[i] * t;
    }
  }
  memcpy(c, ys, sizeof(double) * (size_t)m);
  for (int k = 0; k < size; ++k) { /* Q^T y = H_{size-1} ... H_0 y, HouseholderQR.h:358-362 */
    int rr = m - k;
    double* col = A + k + k * m;
    double tau = h[k];
    if (rr == 1) { c[k] *= 1.0 - tau; continue; }
    if (tau == 0.0) continue;
    double t = 0.0;
    for (int i = 1; i < rr; ++i) t += col[i] * c[k + i];
    t += c[k];
    c[k] -= tau * t;
    for (int i = 1; i < rr; ++i) c[k + i] -= tau * col[i] * t;
  }
  for (int i = size - 1; i >= 0; --i) { /* back substitution on the top triangle */
    double s = c[i];
    for (int j = i + 1; j < size; ++j) s -= A[i + j * m] * out[j];
    out[i] = s / A[i + i * m];
  }
  free(A); free(c); free(h);
  return 0;
}

void oracle_global_kinematic(const double* s, const double* u, double dt, double Lf, double* nx) {
  double x = s[0], y = s[1], psi = s[2], v = s[3], delta = u[0], a = u[1];
  nx[0] = x + v * cos(psi) * dt;
  nx[1] = y + v * sin(psi) * dt;
  nx[2] = psi + v / Lf * delta * dt;
  nx[3] = v + a * dt;
}

/* loops over the two calls above (bench_io.py's cpu_baseline when oracle/_ref is absent) */
void oracle_polyfit_batch(const double* xs, const double* ys, int B, int m, int order, double* coeffs_out) {
  for (int b = 0; b < B; ++b) oracle_polyfit(xs + (size_t)b * m, ys + (size_t)b * m, m, order, coeffs_out + (size_t)b * (order + 1));
}
void oracle_global_kinematic_batch(const double* s, const double* u, int B, double dt, double Lf, double* nx) {
  for (int b = 0; b < B; ++b) oracle_global_kinematic(s + (size_t)b * 4, u + (size_t)b * 2, dt, Lf, nx + (size_t)b * 4);
}

/* ------------------------------------------------------------------------------------------ */
/* The NLP of MPC.cpp                                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const oracle_params* p;
  const double* coef; int ncoef;
  int N, n, m, nb;            /* nb = number of bounded variables = 2(N-1) */
  int xs, ys, ps, vs, cs, es, ds, as;
} nlp_t;

static void nlp_init(nlp_t* q, const oracle_params* p, const double* coef, int ncoef) {
  q->p = p; q->coef = coef; q->ncoef = ncoef;
  int N = p->N;
  q->N = N; q->n = 8 * N - 2; q->m = 6 * N; q->nb = 2 * (N - 1);
  q->xs = 0; q->ys = N; q->ps = 2 * N; q->vs = 3 * N; q->cs = 4 * N; q->es = 5 * N; q->ds = 6 * N; q->as = 7 * N - 1;
}

static void poly_derivs(const double* c, int nc, double x, double* p0, double* p1, double* p2, double* p3) {
  if (nc == 2) { *p0 = c[0] + c[1] * x; *p1 = c[1]; *p2 = 0.0; *p3 = 0.0; return; } /* MPC.cpp:117-118 */
  double a = 0, b = 0, d = 0, e = 0;
  for (int i = nc - 1; i >= 0; --i) { e = e * x + 3.0 * d; d = d * x + 2.0 * b; b = b * x + a; a = a * x + c[i]; }
  *p0 = a; *p1 = b; *p2 = d; *p3 = e;
}

static double nlp_f(const nlp_t* q, const double* x) { /* MPC.cpp:57-76 */
  const oracle_params* p = q->p; int N = q->N; double f = 0.0;
  for (int t = 0; t < N; ++t) {
    f += p->w_cte * pow(x[q->cs + t], 2);
    f += p->w_epsi * pow(x[q->es + t], 2);
    f += p->w_v * pow(x[q->vs + t] - p->ref_v, 2);
  }
  for (int t = 0; t < N - 1; ++t) { f += p->w_delta * pow(x[q->ds + t], 2); f += p->w_a * pow(x[q->as + t], 2); }
  for (int t = 0; t < N - 2; ++t) {
    f += p->w_ddelta * pow(x[q->ds + t + 1] - x[q->ds + t], 2);
    f += p->w_da * pow(x[q->as + t + 1] - x[q->as + t], 2);
  }
  return f;
}

static void nlp_grad(const nlp_t* q, const double* x, double* g) {
  const oracle_params* p = q->p; int N = q->N;
  for (int i = 0; i < q->n; ++i) g[i] = 0.0;
  for (int t = 0; t < N; ++t) {
    g[q->cs + t] = 2.0 * p->w_cte * x[q->cs + t];
    g[q->es + t] = 2.0 * p->w_epsi * x[q->es + t];
    g[q->vs + t] = 2.0 * p->w_v * (x[q->vs + t] - p->ref_v);
  }
  for (int t = 0; t < N - 1; ++t) { g[q->ds + t] += 2.0 * p->w_delta * x[q->ds + t]; g[q->as + t] += 2.0 * p->w_a * x[q->as + t]; }
  for (int t = 0; t < N - 2; ++t) {
    double dd = x[q->ds + t + 1] - x[q->ds + t], da = x[q->as + t + 1] - x[q->as + t];
    g[q->ds + t + 1] += 2.0 * p->w_ddelta * dd; g[q->ds + t] -= 2.0 * p->w_ddelta * dd;
    g[q->as + t + 1] += 2.0 * p->w_da * da; g[q->as + t] -= 2.0 * p->w_da * da;
  }
}

static void nlp_g(const nlp_t* q, const double* x, double* g) { /* MPC.cpp:88-138 */
  const oracle_params* p = q->p; int N = q->N; double dt = p->dt, Lf = p->Lf;
  for (int k = 0; k < 6; ++k) g[k * N] = x[k * N];
  for (int t = 1; t < N; ++t) {
    double x1 = x[q->xs + t], y1 = x[q->ys + t], psi1 = x[q->ps + t], v1 = x[q->vs + t], cte1 = x[q->cs + t], epsi1 = x[q->es + t];
    double x0 = x[q->xs + t - 1], y0 = x[q->ys + t - 1], psi0 = x[q->ps + t - 1], v0 = x[q->vs + t - 1], epsi0 = x[q->es + t - 1];
    double delta0 = x[q->ds + t - 1], a0 = x[q->as + t - 1];
    double f0, d1, d2, d3;
    poly_derivs(q->coef, q->ncoef, x0, &f0, &d1, &d2, &d3);
    double psides0 = atan(d1);
    g[q->xs + t] = x1 - (x0 + v0 * cos(psi0) * dt);
    g[q->ys + t] = y1 - (y0 + v0 * sin(psi0) * dt);
    g[q->ps + t] = psi1 - (psi0 + v0 * delta0 / Lf * dt);
    g[q->vs + t] = v1 - (v0 + a0 * dt);
    g[q->cs + t] = cte1 - ((f0 - y0) + (v0 * sin(epsi0) * dt));
    g[q->es + t] = epsi1 - ((psi0 - psides0) + v0 * delta0 / Lf * dt);
  }
}

/* dense Jacobian, row major m x n */
static void nlp_jac(const nlp_t* q, const double* x, double* J) {
  const oracle_params* p = q->p; int N = q->N, n = q->n; double dt = p->dt, Lf = p->Lf;
  memset(J, 0, sizeof(double) * (size_t)q->m * n);
#define JE(r, c) J[(size_t)(r) * n + (c)]
  for (int k = 0; k < 6; ++k) JE(k * N, k * N) = 1.0;
  for (int t = 1; t < N; ++t) {
    double x0 = x[q->xs + t - 1], psi0 = x[q->ps + t - 1], v0 = x[q->vs + t - 1], epsi0 = x[q->es + t - 1], delta0 = x[q->ds + t - 1];
    double f0, d1, d2, d3;
    poly_derivs(q->coef, q->ncoef, x0, &f0, &d1, &d2, &d3);
    double sp = sin(psi0), cp = cos(psi0), se = sin(epsi0), ce = cos(epsi0);
    JE(q->xs + t, q->xs + t) = 1.0; JE(q->xs + t, q->xs + t - 1) = -1.0;
    JE(q->xs + t, q->ps + t - 1) = v0 * sp * dt; JE(q->xs + t, q->vs + t - 1) = -cp * dt;
    JE(q->ys + t, q->ys + t) = 1.0; JE(q->ys + t, q->ys + t - 1) = -1.0;
    JE(q->ys + t, q->ps + t - 1) = -v0 * cp * dt; JE(q->ys + t, q->vs + t - 1) = -sp * dt;
    JE(q->ps + t, q->ps + t) = 1.0; JE(q->ps + t, q->ps + t - 1) = -1.0;
    JE(q->ps + t, q->vs + t - 1) = -delta0 / Lf * dt; JE(q->ps + t, q->ds + t - 1) = -v0 / Lf * dt;
    JE(q->vs + t, q->vs + t) = 1.0; JE(q->vs + t, q->vs + t - 1) = -1.0; JE(q->vs + t, q->as + t - 1) = -dt;
    JE(q->cs + t, q->cs + t) = 1.0; JE(q->cs + t, q->xs + t - 1) = -d1; JE(q->cs + t, q->ys + t - 1) = 1.0;
    JE(q->cs + t, q->vs + t - 1) = -se * dt; JE(q->cs + t, q->es + t - 1) = -v0 * ce * dt;
    JE(q->es + t, q->es + t) = 1.0; JE(q->es + t, q->ps + t - 1) = -1.0; JE(q->es + t, q->xs + t - 1) = d2 / (1.0 + d1 * d1);
    JE(q->es + t, q->vs + t - 1) = -delta0 / Lf * dt; JE(q->es + t, q->ds + t - 1) = -v0 / Lf * dt;
  }
#undef JE
}

/* dense lower Hessian of sigma*f + lam^T g, row major n x n (only lower triangle written) */
static void nlp_hess(const nlp_t* q, const double* x, const double* lam, double sig, double* H) {
  const oracle_params* p = q->p; int N = q->N, n = q->n; double dt = p->dt, Lf = p->Lf;
  memset(H, 0, sizeof(double) * (size_t)n * n);
#define HE(r, c) H[(size_t)(r) * n + (c)]
  for (int t = 0; t < N; ++t) {
    HE(q->cs + t, q->cs + t) += 2.0 * sig * p->w_cte; HE(q->es + t, q->es + t) += 2.0 * sig * p->w_epsi;
    HE(q->vs + t, q->vs + t) += 2.0 * sig * p->w_v;
  }
  for (int t = 0; t < N - 1; ++t) {
    double nd = (t > 0 ? 1.0 : 0.0) + (t < N - 2 ? 1.0 : 0.0);
    HE(q->ds + t, q->ds + t) += 2.0 * sig * (p->w_delta + nd * p->w_ddelta);
    HE(q->as + t, q->as + t) += 2.0 * sig * (p->w_a + nd * p->w_da);
  }
  for (int t = 0; t < N - 2; ++t) { HE(q->ds + t + 1, q->ds + t) += -2.0 * sig * p->w_ddelta; HE(q->as + t + 1, q->as + t) += -2.0 * sig * p->w_da; }
  for (int t = 1; t < N; ++t) {
    double x0 = x[q->xs + t - 1], psi0 = x[q->ps + t - 1], v0 = x[q->vs + t - 1], epsi0 = x[q->es + t - 1];
    double lx = lam[q->xs + t], ly = lam[q->ys + t], lp = lam[q->ps + t], lc = lam[q->cs + t], le = lam[q->es + t];
    double f0, d1, d2, d3;
    poly_derivs(q->coef, q->ncoef, x0, &f0, &d1, &d2, &d3);
    double sp = sin(psi0), cp = cos(psi0), se = sin(epsi0), ce = cos(epsi0), qq = 1.0 + d1 * d1;
    HE(q->ps + t - 1, q->ps + t - 1) += lx * v0 * cp * dt + ly * v0 * sp * dt;
    HE(q->vs + t - 1, q->ps + t - 1) += lx * sp * dt - ly * cp * dt;
    HE(q->ds + t - 1, q->vs + t - 1) += -(lp + le) * dt / Lf;
    HE(q->es + t - 1, q->es + t - 1) += lc * v0 * se * dt;
    HE(q->es + t - 1, q->vs + t - 1) += -lc * ce * dt;
    HE(q->xs + t - 1, q->xs + t - 1) += -lc * d2 + le * (d3 * qq - 2.0 * d1 * d2 * d2) / (qq * qq);
  }
#undef HE
}

void oracle_mpc_eval(const oracle_params* p, const double* coeffs, int ncoef, const double* x, const double* lambda,
                     double sigma, double* f, double* grad, double* g, double* jac, double* hess) {
  nlp_t q; nlp_init(&q, p, coeffs, ncoef);
  if (f) *f = nlp_f(&q, x);
  if (grad) nlp_grad(&q, x, grad);
  if (g) nlp_g(&q, x, g);
  if (jac) nlp_jac(&q, x, jac);
  if (hess) nlp_hess(&q, x, lambda, sigma, hess);
}

/* ------------------------------------------------------------------------------------------ */
/* Dense symmetric indefinite LDL^T (Bunch-Kaufman partial pivoting, lower, column major)       */
/* stands in for MUMPS (IpMumpsSolverInterface.cpp:402-515): factor, inertia, solve.            */
/* ------------------------------------------------------------------------------------------ */
#define AA(i, j) A[(size_t)(i) + (size_t)(j) * n]
static int ldl_factor(double* A, int n, int* ipiv, int* nneg) {
  const double alpha = (1.0 + sqrt(17.0)) / 8.0;
  int k = 0, info = 0; *nneg = 0;
  while (k < n) {
    int kstep = 1, kp = k, imax = k;
    double absakk = fabs(AA(k, k)), colmax = 0.0;
    for (int i = k + 1; i < n; ++i) if (fabs(AA(i, k)) > colmax) { colmax = fabs(AA(i, k)); imax = i; }
    if ((absakk > colmax ? absakk : colmax) == 0.0) { if (!info) info = k + 1; kp = k; }
    else {
      if (absakk >= alpha * colmax) kp = k;
      else {
        double rowmax = 0.0;
        for (int j = k; j < imax; ++j) if (fabs(AA(imax, j)) > rowmax) rowmax = fabs(AA(imax, j));
        for (int i = imax + 1; i < n; ++i) if (fabs(AA(i, imax)) > rowmax) rowmax = fabs(AA(i, imax));
        if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
        else if (fabs(AA(imax, imax)) >= alpha * rowmax) kp = imax;
        else { kp = imax; kstep = 2; }
      }
      int kk = k + kstep - 1;
      if (kp != kk) {
        for (int i = kp + 1; i < n; ++i) { double t = AA(i, kk); AA(i, kk) = AA(i, kp); AA(i, kp) = t; }
        for (int j = kk + 1; j < kp; ++j) { double t = AA(j, kk); AA(j, kk) = AA(kp, j); AA(kp, j) = t; }
        { double t = AA(kk, kk); AA(kk, kk) = AA(kp, kp); AA(kp, kp) = t; }
        if (kstep == 2) { double t = AA(k + 1, k); AA(k + 1, k) = AA(kp, k); AA(kp, k) = t; }
      }
      if (kstep == 1) {
        if (AA(k, k) < 0.0) ++*nneg;
        if (k < n - 1) {
          double r1 = 1.0 / AA(k, k);
          for (int j = k + 1; j < n; ++j) {
            double w = r1 * AA(j, k);
            if (w != 0.0) for (int i = j; i < n; ++i) AA(i, j) -= AA(i, k) * w;
          }
          for (int i = k + 1; i < n; ++i) AA(i, k) *= r1;
        }
      } else {
        double a11 = AA(k, k), a21 = AA(k + 1, k), a22 = AA(k + 1, k + 1);
        double det = a11 * a22 - a21 * a21, tr = a11 + a22;
        if (det < 0.0) ++*nneg; else if (tr < 0.0) *nneg += 2;
        if (k < n - 2) {
          double d21 = a21, d11 = a22 / d21, d22 = a11 / d21, t = 1.0 / (d11 * d22 - 1.0);
          d21 = t / d21;
          for (int j = k + 2; j < n; ++j) {
            double wk = d21 * (d11 * AA(j, k) - AA(j, k + 1)), wkp1 = d21 * (d22 * AA(j, k + 1) - AA(j, k));
            for (int i = j; i < n; ++i) AA(i, j) -= AA(i, k) * wk + AA(i, k + 1) * wkp1;
            AA(j, k) = wk; AA(j, k + 1) = wkp1;
          }
        }
      }
    }
    if (kstep == 1) ipiv[k] = kp; else { ipiv[k] = -(kp + 1); ipiv[k + 1] = -(kp + 1); }
    k += kstep;
  }
  return info;
}

static void ldl_solve(const double* A, int n, const int* ipiv, double* b) {
  int k = 0;
  while (k < n) { /* solve L D y = P b */
    if (ipiv[k] >= 0) {
      int kp = ipiv[k];
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      for (int i = k + 1; i < n; ++i) b[i] -= AA(i, k) * b[k];
      b[k] /= AA(k, k);
      k += 1;
    } else {
      int kp = -ipiv[k] - 1;
      if (kp != k + 1) { double t = b[k + 1]; b[k + 1] = b[kp]; b[kp] = t; }
      for (int i = k + 2; i < n; ++i) b[i] -= AA(i, k) * b[k] + AA(i, k + 1) * b[k + 1];
      double akm1k = AA(k + 1, k), akm1 = AA(k, k) / akm1k, ak = AA(k + 1, k + 1) / akm1k, denom = akm1 * ak - 1.0;
      double bkm1 = b[k] / akm1k, bk = b[k + 1] / akm1k;
      b[k] = (ak * bkm1 - bk) / denom; b[k + 1] = (akm1 * bk - bkm1) / denom;
      k += 2;
    }
  }
  k = n - 1;
  while (k >= 0) { /* solve L^T x = y, undo P */
    if (ipiv[k] >= 0) {
      double s = 0.0;
      for (int i = k + 1; i < n; ++i) s += AA(i, k) * b[i];
      b[k] -= s;
      int kp = ipiv[k];
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k -= 1;
    } else {
      double s0 = 0.0, s1 = 0.0;
      for (int i = k + 1; i < n; ++i) { s0 += AA(i, k) * b[i]; s1 += AA(i, k - 1) * b[i]; }
      b[k] -= s0; b[k - 1] -= s1;
      int kp = -ipiv[k] - 1;
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k -= 2;
    }
  }
}
#undef AA

/* ------------------------------------------------------------------------------------------ */
/* Interior-point state                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  nlp_t q;
  int n, m, nb, b0;          /* b0 = index of the first bounded variable (delta_start) */
  double df;                 /* objective scaling (IpGradientScaling.cpp:99-116) */
  double *rhs_c;             /* constraint right-hand side (initial state rows) */
  double *xL, *xU;           /* relaxed bounds of the nb bounded variables */
  double *oL, *oU;           /* original bounds */
  /* current iterate */
  double *x, *lam, *zL, *zU;
  /* evaluated at current iterate */
  double f, *grad, *c, *J, *W;
  /* search direction */
  double *dx, *dlam, *dzL, *dzU;
  /* trial */
  double *xt, *ct; double ft;
  /* KKT */
  double *K; int *ipiv; double *sol, *aug;
  double mu, tau;
  /* perturbation handler */
  double dx_curr, dx_last, dc_curr;
  /* filter */
  double fphi[MAXF], fth[MAXF]; int nf;
  int last_rej_filter, count_rej_filter; /* filter reset heuristic (IpFilterLSAcceptor.cpp:357-379) */
  double theta_max, theta_min;
  int regu_tries;
} ip_t;

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static double vmaxabs(const double* v, int n) { double r = 0.0; for (int i = 0; i < n; ++i) if (fabs(v[i]) > r) r = fabs(v[i]); return r; }
static double vasum(const double* v, int n) { double r = 0.0; for (int i = 0; i < n; ++i) r += fabs(v[i]); return r; }

static void eval_point(ip_t* s, const double* x, double* f, double* c) {
  *f = s->df * nlp_f(&s->q, x);
  nlp_g(&s->q, x, c);
  for (int i = 0; i < s->m; ++i) c[i] -= s->rhs_c[i];
}

static void eval_derivs(ip_t* s) { /* at s->x, s->lam */
  nlp_grad(&s->q, s->x, s->grad);
  for (int i = 0; i < s->n; ++i) s->grad[i] *= s->df;
  nlp_jac(&s->q, s->x, s->J);
  nlp_hess(&s->q, s->x, s->lam, s->df, s->W);
}

/* slack of bounded variable k at point x (IpIpoptCalculatedQuantities.cpp:444-507, without the bound move) */
static double slackL(const ip_t* s, const double* x, int k) {
  double v = x[s->b0 + k] - s->xL[k];
  double smin = DBL_EPSILON * (s->mu < 1.0 ? s->mu : 1.0);
  if (v < smin) { double t = s->mu / s->zL[k]; if (t < smin) t = smin; double cap = (fabs(s->xL[k]) > 1.0 ? fabs(s->xL[k]) : 1.0) * pow(DBL_EPSILON, 0.75) + (v > 0 ? v : 0); v = t < cap ? t : cap; }
  return v;
}
static double slackU(const ip_t* s, const double* x, int k) {
  double v = s->xU[k] - x[s->b0 + k];
  double smin = DBL_EPSILON * (s->mu < 1.0 ? s->mu : 1.0);
  if (v < smin) { double t = s->mu / s->zU[k]; if (t < smin) t = smin; double cap = (fabs(s->xU[k]) > 1.0 ? fabs(s->xU[k]) : 1.0) * pow(DBL_EPSILON, 0.75) + (v > 0 ? v : 0); v = t < cap ? t : cap; }
  return v;
}

static double barrier_obj(const ip_t* s, const double* x, double f) { /* :697-751 */
  double sl = 0.0, su = 0.0;
  for (int k = 0; k < s->nb; ++k) sl += log(slackL(s, x, k));
  for (int k = 0; k < s->nb; ++k) su += log(slackU(s, x, k));
  return f + (-s->mu) * (sl + su);
}

/* grad_lag_x = grad f + J^T lam - zL + zU */
static void grad_lag(const ip_t* s, double* r) {
  for (int j = 0; j < s->n; ++j) r[j] = s->grad[j];
  for (int i = 0; i < s->m; ++i) {
    double l = s->lam[i]; if (l == 0.0) continue;
    const double* Ji = s->J + (size_t)i * s->n;
    for (int j = 0; j < s->n; ++j) r[j] += Ji[j] * l;
  }
  for (int k = 0; k < s->nb; ++k) r[s->b0 + k] += -s->zL[k] + s->zU[k];
}

static void err_scaling(const ip_t* s, double* sd, double* sc) { /* :3279-3306, s_max = 100 */
  double zsum = vasum(s->zL, s->nb) + vasum(s->zU, s->nb);
  double c = zsum / (2.0 * s->nb);
  *sc = (c > 100.0 ? c : 100.0) / 100.0;
  double d = (vasum(s->lam, s->m) + zsum) / (s->m + 2.0 * s->nb);
  *sd = (d > 100.0 ? d : 100.0) / 100.0;
}

static double compl_err(const ip_t* s, double mu) {
  double r = 0.0;
  for (int k = 0; k < s->nb; ++k) {
    double a = fabs(slackL(s, s->x, k) * s->zL[k] - mu), b = fabs(slackU(s, s->x, k) * s->zU[k] - mu);
    if (a > r) r = a;
    if (b > r) r = b;
  }
  return r;
}

/* Assemble and factor [[W+Sigma+dx I, J^T],[J, -dc I]]; returns 0 ok, 1 wrong inertia, 2 singular */
static int kkt_factor(ip_t* s, int ls_mode, double delta_x, double delta_c) {
  int n = s->n, m = s->m, d = n + m;
  double* K = s->K;
  memset(K, 0, sizeof(double) * (size_t)d * d);
  for (int i = 0; i < n; ++i) {
    if (!ls_mode) for (int j = 0; j <= i; ++j) K[(size_t)i + (size_t)j * d] = s->W[(size_t)i * n + j];
    K[(size_t)i + (size_t)i * d] += (ls_mode ? 1.0 : 0.0) + delta_x;
  }
  if (!ls_mode)
    for (int k = 0; k < s->nb; ++k) {
      int i = s->b0 + k;
      K[(size_t)i + (size_t)i * d] += s->zL[k] / slackL(s, s->x, k) + s->zU[k] / slackU(s, s->x, k);
    }
  for (int r = 0; r < m; ++r) {
    const double* Jr = s->J + (size_t)r * n;
    for (int j = 0; j < n; ++j) K[(size_t)(n + r) + (size_t)j * d] = Jr[j];
    K[(size_t)(n + r) + (size_t)(n + r) * d] = -delta_c;
  }
  int nneg = 0;
  int info = ldl_factor(K, d, s->ipiv, &nneg);
  if (info) return 2;
  if (nneg != m) return 1;
  return 0;
}

/* PDPerturbationHandler::get_deltas_for_wrong_inertia (:347-391) */
static int next_delta_x(ip_t* s) {
  if (s->dx_curr == 0.0) {
    if (s->dx_last == 0.0) s->dx_curr = 1e-4;
    else { double v = s->dx_last / 3.0; s->dx_curr = v > 1e-20 ? v : 1e-20; }
  } else {
    if (s->dx_last == 0.0 || 1e5 * s->dx_last < s->dx_curr) s->dx_curr *= 100.0; else s->dx_curr *= 8.0;
  }
  if (s->dx_curr > 1e20) { s->dx_last = 0.0; return 0; }
  return 1;
}

/* Factor the primal-dual matrix for the current iterate with inertia correction
 * (PDFullSpaceSolver::SolveOnce :470-626 + PDPerturbationHandler::ConsiderNewSystem :148-236). */
static int pd_factor(ip_t* s) {
  if (s->dx_curr > 0.0) s->dx_last = s->dx_curr;
  s->dx_curr = 0.0; s->dc_curr = 0.0;
  s->regu_tries = 0;
  for (;;) {
    int r = kkt_factor(s, 0, s->dx_curr, s->dc_curr);
    ++s->regu_tries;
    if (r == 0) return 1;
    if (r == 2 && s->dc_curr == 0.0) { s->dc_curr = 1e-8 * pow(s->mu, 0.25); continue; } /* PerturbForSingularity */
    if (!next_delta_x(s)) return 0;
  }
}

/* One augmented-system back-solve for the full 8-block rhs (SolveOnce :415-421, :640-643):
 * rhs = (rx[n], rc[m], rzL[nb], rzU[nb]) -> sol (x, lam, zL, zU).  No sign flip. */
static void pd_backsolve(ip_t* s, const double* rx, const double* rc, const double* rzL, const double* rzU, double* sx,
                         double* sl, double* szL, double* szU) {
  int n = s->n, m = s->m;
  double* b = s->aug;
  for (int j = 0; j < n; ++j) b[j] = rx[j];
  for (int k = 0; k < s->nb; ++k) b[s->b0 + k] += rzL[k] / slackL(s, s->x, k) - rzU[k] / slackU(s, s->x, k);
  for (int i = 0; i < m; ++i) b[n + i] = rc[i];
  ldl_solve(s->K, n + m, s->ipiv, b);
  for (int j = 0; j < n; ++j) sx[j] = b[j];
  for (int i = 0; i < m; ++i) sl[i] = b[n + i];
  for (int k = 0; k < s->nb; ++k) {
    szL[k] = (rzL[k] - s->zL[k] * sx[s->b0 + k]) / slackL(s, s->x, k);
    szU[k] = (rzU[k] + s->zU[k] * sx[s->b0 + k]) / slackU(s, s->x, k);
  }
}

/* residual of the full (unreduced) system, ComputeResiduals :653-775 */
static void pd_resid(ip_t* s, const double* rx, const double* rc, const double* rzL, const double* rzU, const double* sx,
                     const double* sl, const double* szL, const double* szU, double* ox, double* oc, double* ozL, double* ozU) {
  int n = s->n, m = s->m;
  for (int i = 0; i < n; ++i) {
    double a = 0.0;
    for (int j = 0; j <= i; ++j) a += s->W[(size_t)i * n + j] * sx[j];
    for (int j = i + 1; j < n; ++j) a += s->W[(size_t)j * n + i] * sx[j];
    ox[i] = a;
  }
  for (int r = 0; r < m; ++r) {
    const double* Jr = s->J + (size_t)r * n; double l = sl[r], a = 0.0;
    for (int j = 0; j < n; ++j) { ox[j] += Jr[j] * l; a += Jr[j] * sx[j]; }
    oc[r] = a - s->dc_curr * sl[r] - rc[r];
  }
  for (int k = 0; k < s->nb; ++k) ox[s->b0 + k] += -szL[k] + szU[k];
  for (int j = 0; j < n; ++j) ox[j] += s->dx_curr * sx[j] - rx[j];
  for (int k = 0; k < s->nb; ++k) {
    ozL[k] = szL[k] * slackL(s, s->x, k) + s->zL[k] * sx[s->b0 + k] - rzL[k];
    ozU[k] = szU[k] * slackU(s, s->x, k) - s->zU[k] * sx[s->b0 + k] - rzU[k];
  }
}

static double max4(const double* a, int na, const double* b, int nb, const double* c, int nc, const double* d, int nd) {
  double r = vmaxabs(a, na), t;
  t = vmaxabs(b, nb);
  if (t > r) r = t;
  t = vmaxabs(c, nc);
  if (t > r) r = t;
  t = vmaxabs(d, nd);
  if (t > r) r = t;
  return r;
}

/* PDFullSpaceSolver::Solve with iterative refinement (:132-374); matrix must be factored.  Output = -solution. */
static void pd_solve(ip_t* s, const double* rx, const double* rc, const double* rzL, const double* rzU, double* dx,
                     double* dl, double* dzL, double* dzU) {
  int n = s->n, m = s->m, nb = s->nb;
  double* w = (double*)malloc(sizeof(double) * (size_t)(2 * (n + m + 2 * nb)));
  double *ex = w, *ec = ex + n, *ezL = ec + m, *ezU = ezL + nb, *cx = ezU + nb, *cc = cx + n, *czL = cc + m, *czU = czL + nb;
  pd_backsolve(s, rx, rc, rzL, rzU, dx, dl, dzL, dzU);
  pd_resid(s, rx, rc, rzL, rzU, dx, dl, dzL, dzU, ex, ec, ezL, ezU);
  double nrm_rhs = max4(rx, n, rc, m, rzL, nb, rzU, nb);
  double ratio, ratio_old;
  {
    double nres = max4(dx, n, dl, m, dzL, nb, dzU, nb), nresid = max4(ex, n, ec, m, ezL, nb, ezU, nb);
    double mn = nres < 1e6 * nrm_rhs ? nres : 1e6 * nrm_rhs;
    ratio = (nrm_rhs + nres == 0.0) ? nresid : nresid / (mn + nrm_rhs);
  }
  ratio_old = ratio;
  int it = 0;
  while (it < 1 || ratio > 1e-10) {
    pd_backsolve(s, ex, ec, ezL, ezU, cx, cc, czL, czU);
    for (int j = 0; j < n; ++j) dx[j] -= cx[j];
    for (int i = 0; i < m; ++i) dl[i] -= cc[i];
    for (int k = 0; k < nb; ++k) { dzL[k] -= czL[k]; dzU[k] -= czU[k]; }
    pd_resid(s, rx, rc, rzL, rzU, dx, dl, dzL, dzU, ex, ec, ezL, ezU);
    double nres = max4(dx, n, dl, m, dzL, nb, dzU, nb), nresid = max4(ex, n, ec, m, ezL, nb, ezU, nb);
    double mn = nres < 1e6 * nrm_rhs ? nres : 1e6 * nrm_rhs;
    ratio = (nrm_rhs + nres == 0.0) ? nresid : nresid / (mn + nrm_rhs);
    ++it;
    if (ratio > 1e-10 && it > 1 && (it > 10 || ratio > (1.0 - 1e-9) * ratio_old)) break; /* give up, accept */
    ratio_old = ratio;
  }
  for (int j = 0; j < n; ++j) dx[j] = -dx[j];
  for (int i = 0; i < m; ++i) dl[i] = -dl[i];
  for (int k = 0; k < nb; ++k) { dzL[k] = -dzL[k]; dzU[k] = -dzU[k]; }
  free(w);
}

static double frac_to_bound_primal(const ip_t* s, const double* dx, double tau) { /* :2949-3011, IpDenseVector.cpp:928-970 */
  double a = 1.0;
  for (int k = 0; k < s->nb; ++k) {
    double d = dx[s->b0 + k];
    if (d < 0.0) { double t = -tau / d * slackL(s, s->x, k); if (t < a) a = t; }
  }
  for (int k = 0; k < s->nb; ++k) {
    double d = -dx[s->b0 + k];
    if (d < 0.0) { double t = -tau / d * slackU(s, s->x, k); if (t < a) a = t; }
  }
  return a;
}
static double frac_to_bound_dual(const ip_t* s, const double* dzL, const double* dzU, double tau) {
  double a = 1.0;
  for (int k = 0; k < s->nb; ++k) if (dzL[k] < 0.0) { double t = -tau / dzL[k] * s->zL[k]; if (t < a) a = t; }
  for (int k = 0; k < s->nb; ++k) if (dzU[k] < 0.0) { double t = -tau / dzU[k] * s->zU[k]; if (t < a) a = t; }
  return a;
}

static int cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(bas); }

static int filter_ok(const ip_t* s, double phi, double th) {
  for (int i = 0; i < s->nf; ++i) if (!(phi <= s->fphi[i] || th <= s->fth[i])) return 0;
  return 1;
}
static void filter_add(ip_t* s, double phi, double th) {
  int w = 0;
  for (int i = 0; i < s->nf; ++i) if (!(s->fphi[i] >= phi && s->fth[i] >= th)) { s->fphi[w] = s->fphi[i]; s->fth[w] = s->fth[i]; ++w; }
  s->nf = w;
  if (s->nf < MAXF) { s->fphi[s->nf] = phi; s->fth[s->nf] = th; ++s->nf; }
}

typedef struct { double ref_theta, ref_barr, ref_gbd; } ls_ref;

static int is_ftype(const ls_ref* r, double alpha_test) { /* IpFilterLSAcceptor.cpp:246-265 */
  return r->ref_gbd < 0.0 && alpha_test * pow(-r->ref_gbd, 2.3) > 1.0 * pow(r->ref_theta, 1.1);
}
static int armijo(const ls_ref* r, double alpha_test, double trial_barr) {
  return cmp_le(trial_barr - r->ref_barr, 1e-8 * alpha_test * r->ref_gbd, r->ref_barr);
}

/* FilterLSAcceptor::CheckAcceptabilityOfTrialPoint :279-382 */
static int check_accept(ip_t* s, const ls_ref* r, double alpha_test, double trial_theta, double trial_barr) {
  if (s->theta_max < 0.0) s->theta_max = 1e4 * (r->ref_theta > 1.0 ? r->ref_theta : 1.0);
  if (s->theta_min < 0.0) s->theta_min = 1e-4 * (r->ref_theta > 1.0 ? r->ref_theta : 1.0);
  if (s->theta_max > 0 && trial_theta > s->theta_max) return 0;
  int accept;
  if (alpha_test > 0.0 && is_ftype(r, alpha_test) && r->ref_theta <= s->theta_min) accept = armijo(r, alpha_test, trial_barr);
  else {
    accept = 1;
    if (trial_barr > r->ref_barr) { /* obj_max_inc = 5 */
      double bas = 1.0;
      if (fabs(r->ref_barr) > 10.0) bas = log10(fabs(r->ref_barr));
      if (log10(trial_barr - r->ref_barr) > 5.0 + bas) accept = 0;
    }
    if (accept)
      accept = cmp_le(trial_theta, (1.0 - 1e-5) * r->ref_theta, r->ref_theta) ||
               cmp_le(trial_barr - r->ref_barr, -1e-8 * r->ref_theta, r->ref_barr);
  }
  if (!accept) { s->last_rej_filter = 0; return 0; }
  if (!filter_ok(s, trial_barr, trial_theta)) { s->last_rej_filter = 1; return 0; }
  /* filter reset heuristic :357-379 (filter_reset_trigger = 5; Ipopt 3.12.7 never counts the resets, so
   * max_filter_resets = 5 is not reached) */
  if (s->last_rej_filter) {
    if (++s->count_rej_filter >= 5) { s->nf = 0; s->count_rej_filter = 0; }
  } else s->count_rej_filter = 0;
  s->last_rej_filter = 0;
  return 1;
}

/* Primal-dual system error for the barrier parameter mu at (x, lam, zL, zU): 1-norms of grad_x L, c and the relaxed
 * complementarity, each divided by its number of entries (IpIpoptCalculatedQuantities.cpp:2835-2884).  Used by the soft
 * restoration phase only; evaluates the NLP into scratch memory. */
static double pd_system_error(const ip_t* s, const double* x, const double* lam, const double* zL, const double* zU) {
  const int n = s->n, m = s->m, nb = s->nb;
  double* g = (double*)calloc((size_t)n + (size_t)m + (size_t)m * n, sizeof(double));
  double *c = g + n, *J = c + m;
  nlp_grad(&s->q, x, g);
  for (int j = 0; j < n; ++j) g[j] *= s->df;
  nlp_g(&s->q, x, c);
  nlp_jac(&s->q, x, J);
  for (int i = 0; i < m; ++i) {
    c[i] -= s->rhs_c[i];
    const double l = lam[i];
    if (l == 0.0) continue;
    const double* Ji = J + (size_t)i * n;
    for (int j = 0; j < n; ++j) g[j] += Ji[j] * l;
  }
  for (int k = 0; k < nb; ++k) g[s->b0 + k] += -zL[k] + zU[k];
  double cm = 0.0;
  for (int k = 0; k < nb; ++k) cm += fabs(slackL(s, x, k) * zL[k] - s->mu) + fabs(slackU(s, x, k) * zU[k] - s->mu);
  const double r = vasum(g, n) / n + vasum(c, m) / m + cm / (2.0 * nb);
  free(g);
  return r;
}

int oracle_mpc_solve(const oracle_params* p, const double* state6, const double* coeffs, int ncoef, double* x_out,
                     double* out8, double* obj_out, int* iters_out, double* lambda_out, double* trace, int trace_cap,
                     int* trace_rows) {
  ip_t S; ip_t* s = &S;
  memset(s, 0, sizeof(S));
  nlp_init(&s->q, p, coeffs, ncoef);
  int N = p->N, n = s->q.n, m = s->q.m, nb = s->q.nb;
  s->n = n; s->m = m; s->nb = nb; s->b0 = s->q.ds;
  int d = n + m;
  size_t nd = (size_t)(8 * n + 8 * m + 12 * nb + 2 * d) + (size_t)m * n + (size_t)n * n + (size_t)d * d + 16;
  double* mem = (double*)calloc(nd, sizeof(double));
  double* q = mem;
#define TAKE(k) (q += (k), q - (k))
  s->rhs_c = TAKE(m); s->xL = TAKE(nb); s->xU = TAKE(nb); s->oL = TAKE(nb); s->oU = TAKE(nb);
  s->x = TAKE(n); s->lam = TAKE(m); s->zL = TAKE(nb); s->zU = TAKE(nb);
  s->grad = TAKE(n); s->c = TAKE(m); s->J = TAKE((size_t)m * n); s->W = TAKE((size_t)n * n);
  s->dx = TAKE(n); s->dlam = TAKE(m); s->dzL = TAKE(nb); s->dzU = TAKE(nb);
  s->xt = TAKE(n); s->ct = TAKE(m);
  s->K = TAKE((size_t)d * d); s->sol = TAKE(d); s->aug = TAKE(d);
  double *rx = TAKE(n), *rzL = TAKE(nb), *rzU = TAKE(nb), *csoc = TAKE(m);
  double *sdx = TAKE(n), *sdl = TAKE(m), *sdzL = TAKE(nb), *sdzU = TAKE(nb), *glag = TAKE(n);
#undef TAKE
  s->ipiv = (int*)malloc(sizeof(int) * (size_t)d);
  int status = -100, iter = 0, nrows = 0;

  /* bounds: MPC.cpp:185-203; relaxed by 1e-8*max(1,|b|) (IpOrigIpoptNLP.cpp:369-372,466-482) */
  for (int k = 0; k < nb; ++k) {
    double b = k < N - 1 ? p->delta_max : p->a_max;
    s->oL[k] = -b; s->oU[k] = b;
    double rel = 1e-8 * (fabs(b) > 1.0 ? fabs(b) : 1.0);
    s->xL[k] = -b - rel; s->xU[k] = b + rel;
  }
  /* start point: MPC.cpp:167-177; constraint rhs: MPC.cpp:208-226 */
  for (int k = 0; k < 6; ++k) { s->x[k * N] = state6[k]; s->rhs_c[k * N] = state6[k]; }
  /* objective scaling at the user start point (IpGradientScaling.cpp:99-116) */
  s->df = 1.0;
  nlp_grad(&s->q, s->x, s->grad);
  { double g = vmaxabs(s->grad, n); if (g > 100.0) s->df = 100.0 / g; if (s->df < 1e-8) s->df = 1e-8; }
  /* push bounded variables into the interior (IpDefaultIterateInitializer.cpp:469-649; bound_push=bound_frac=0.01) */
  for (int k = 0; k < nb; ++k) {
    double v = s->x[s->b0 + k], l = s->xL[k], u = s->xU[k];
    v = clampd(v, l, u);
    double pl = 0.01 * (fabs(l) > 1.0 ? fabs(l) : 1.0), ql = 0.01 * (u - l);
    if (ql < pl) pl = ql;
    double pu = 0.01 * (fabs(u) > 1.0 ? fabs(u) : 1.0);
    if (ql < pu) pu = ql;
    v = clampd(v, l + pl, u - pu);
    s->x[s->b0 + k] = v;
  }
  for (int k = 0; k < nb; ++k) s->zL[k] = s->zU[k] = 1.0; /* bound_mult_init_val */
  s->mu = 0.1; s->tau = 0.99 > 1.0 - s->mu ? 0.99 : 1.0 - s->mu;
  s->theta_max = s->theta_min = -1.0; s->nf = 0;

  /* least-square multipliers: [[I, J^T],[J, 0]] (sol_x, y) = (zL - zU - grad f, 0)  (IpLeastSquareMults.cpp:51-83) */
  eval_point(s, s->x, &s->f, s->c);
  /* NaN / Inf in the inputs: Ipopt stops at the starting point with Invalid_Number_Detected (-13), no iteration */
  int invalid_start = !(fabs(s->f) <= DBL_MAX);
  for (int i = 0; i < m; ++i) if (!(fabs(s->c[i]) <= DBL_MAX)) invalid_start = 1;
  eval_derivs(s);
  {
    int r = kkt_factor(s, 1, 0.0, 0.0);
    if (r == 0) {
      double* b = s->aug;
      for (int j = 0; j < n; ++j) b[j] = -s->grad[j];
      for (int k = 0; k < nb; ++k) b[s->b0 + k] += s->zL[k] - s->zU[k];
      for (int i = 0; i < m; ++i) b[n + i] = 0.0;
      ldl_solve(s->K, d, s->ipiv, b);
      for (int i = 0; i < m; ++i) s->lam[i] = b[n + i];
      if (vmaxabs(s->lam, m) > 1000.0) for (int i = 0; i < m; ++i) s->lam[i] = 0.0; /* constr_mult_init_max */
    }
  }

  int acceptable_counter = 0; double curr_obj_val = -1e50, last_obj_val = -1e50;
  int tiny_step_last = 0, tiny_step_flag = 0, mu_initialized = 0;
  int in_soft_resto = 0, soft_resto_counter = 0; /* IpBacktrackingLineSearch.cpp: in_soft_resto_phase_, soft_resto_counter_ */
  double info_alpha_pr = 0.0, info_alpha_du = 0.0, info_dnorm = 0.0; int info_ls = 0; double info_regu = 0.0;
  const double mu_min = (p->tol < 1e-4 * s->df ? p->tol : 1e-4 * s->df) / (10.0 + 1.0);

  for (;;) {
    if (invalid_start) { status = -13; break; }
    /* ---- quantities at the current iterate ---- */
    eval_point(s, s->x, &s->f, s->c);
    eval_derivs(s);
    grad_lag(s, glag);
    double sd, sc; err_scaling(s, &sd, &sc);
    double dual_inf = vmaxabs(glag, n), prim_inf = vmaxabs(s->c, m);
    double compl0 = compl_err(s, 0.0);
    double E0 = dual_inf / sd; if (prim_inf > E0) E0 = prim_inf; if (compl0 / sc > E0) E0 = compl0 / sc;
    if (trace && nrows < trace_cap) {
      double* r = trace + 10 * nrows;
      r[0] = iter; r[1] = s->f / s->df; r[2] = prim_inf; r[3] = dual_inf; r[4] = s->mu; r[5] = info_dnorm; r[6] = info_regu;
      r[7] = info_alpha_du; r[8] = info_alpha_pr; r[9] = info_ls;
    }
    ++nrows;
    /* ---- convergence (IpOptErrorConvCheck.cpp:204-262) ---- */
    double u_dual = dual_inf / s->df, u_compl = compl0 / s->df;
    if (E0 <= p->tol && u_dual <= 1.0 && prim_inf <= 1e-4 && u_compl <= 1e-4) { status = 0; break; }
    {
      last_obj_val = curr_obj_val; curr_obj_val = s->f;
      int acc = E0 <= 1e-6 && u_dual <= 1e10 && prim_inf <= 1e-2 && u_compl <= 1e-2 &&
                fabs(curr_obj_val - last_obj_val) / (fabs(curr_obj_val) > 1.0 ? fabs(curr_obj_val) : 1.0) <= 1e20;
      if (acc) { if (++acceptable_counter >= 15) { status = 1; break; } } else acceptable_counter = 0;
    }
    if (vmaxabs(s->x, n) > 1e20) { status = -4; break; } /* Diverging_Iterates */
    if (iter >= p->max_iter) { status = -1; break; }

    /* ---- barrier parameter (IpMonotoneMuUpdate.cpp:132-232) ---- */
    {
      double Emu = dual_inf / sd; if (prim_inf > Emu) Emu = prim_inf;
      double cm = compl_err(s, s->mu) / sc; if (cm > Emu) Emu = cm;
      int done = 0, tiny = tiny_step_flag; tiny_step_flag = 0;
      int stop_tiny = 0;
      while ((Emu <= 10.0 * s->mu || tiny) && !done) {
        double nm = 0.2 * s->mu, pm = pow(s->mu, 1.5);
        if (pm < nm) nm = pm;
        if (nm < mu_min) nm = mu_min;
        double nt = 0.99 > 1.0 - nm ? 0.99 : 1.0 - nm;
        int changed = (nm != s->mu);
        if (!changed && tiny) { stop_tiny = 1; break; }
        s->mu = nm; s->tau = nt;
        if (!changed) done = 1;
        else {
          Emu = dual_inf / sd; if (prim_inf > Emu) Emu = prim_inf;
          cm = compl_err(s, s->mu) / sc; if (cm > Emu) Emu = cm;
          done = Emu > 10.0 * s->mu;
        }
        if (done && changed) { s->nf = 0; s->last_rej_filter = 0; s->count_rej_filter = 0; } /* linesearch_->Reset(): filter cleared */
        tiny = 0;
      }
      (void)mu_initialized; mu_initialized = 1;
      if (stop_tiny) { status = 3; break; }
    }

    /* ---- search direction (IpPDSearchDirCalc.cpp:60-139) ---- */
    for (int j = 0; j < n; ++j) rx[j] = glag[j];
    for (int k = 0; k < nb; ++k) { rzL[k] = slackL(s, s->x, k) * s->zL[k] - s->mu; rzU[k] = slackU(s, s->x, k) * s->zU[k] - s->mu; }
    if (!pd_factor(s)) { status = -3; break; }
    pd_solve(s, rx, s->c, rzL, rzU, s->dx, s->dlam, s->dzL, s->dzU);
    info_regu = s->dx_curr;
    info_dnorm = vmaxabs(s->dx, n);

    /* ---- line search (IpBacktrackingLineSearch.cpp:261-635, 637-797) ---- */
    ls_ref R;
    R.ref_theta = vasum(s->c, m);
    R.ref_barr = barrier_obj(s, s->x, s->f);
    {
      double g = 0.0;
      for (int j = 0; j < n; ++j) g += s->grad[j] * s->dx[j];
      for (int k = 0; k < nb; ++k) g += (-s->mu / slackL(s, s->x, k) + s->mu / slackU(s, s->x, k)) * s->dx[s->b0 + k];
      R.ref_gbd = g;
    }
    double *adl = s->dlam, *adzL = s->dzL, *adzU = s->dzU; /* actual_delta */
    double alpha = 0.0; int n_steps = 0, accept = 0;
    int soft_step = 0; double soft_a_lam = 0.0, soft_a_z = 0.0; /* a soft restoration step was taken: step sizes of lambda, z */
    /* tiny step (:1145-1200, :377-423) */
    int tiny = 1;
    for (int j = 0; j < n && tiny; ++j) if (fabs(s->dx[j]) / (fabs(s->x[j]) + 1.0) > 10.0 * DBL_EPSILON) tiny = 0;
    if (tiny) {
      alpha = frac_to_bound_primal(s, s->dx, s->tau);
      for (int j = 0; j < n; ++j) s->xt[j] = s->x[j] + alpha * s->dx[j];
      if (tiny_step_last) tiny_step_flag = 1;
      tiny_step_last = vmaxabs(s->dlam, m) < 1e-2;
      accept = 1; info_ls = 0;
    } else {
      tiny_step_last = 0;
      double alpha_max = frac_to_bound_primal(s, s->dx, s->tau);
      double alpha_min; /* CalculateAlphaMin :393-410 */
      {
        double am = 1e-5;
        if (R.ref_gbd < 0) {
          double t = 1e-8 * R.ref_theta / (-R.ref_gbd); if (t < am) am = t;
          if (R.ref_theta <= s->theta_min) { t = 1.0 * pow(R.ref_theta, 1.1) / pow(-R.ref_gbd, 2.3); if (t < am) am = t; }
        }
        alpha_min = 0.05 * am;
      }
      /* TrySoftRestoStep (:1043-1140): the damped full step, the same step size for x, lambda and z.  1: acceptable to the
       * original test with alpha_test = 0, 2: reduces the primal-dual system error by 1 - 1e-4, 0: rejected. */
#define TRY_SOFT(result)                                                                                               \
  do {                                                                                                                \
    double a_du_max = frac_to_bound_dual(s, s->dzL, s->dzU, s->tau);                                                  \
    double a_s = alpha_max < a_du_max ? alpha_max : a_du_max;                                                         \
    for (int j = 0; j < n; ++j) s->xt[j] = s->x[j] + a_s * s->dx[j];                                                  \
    eval_point(s, s->xt, &s->ft, s->ct);                                                                              \
    (result) = 0;                                                                                                     \
    if (check_accept(s, &R, 0.0, vasum(s->ct, m), barrier_obj(s, s->xt, s->ft))) (result) = 1;                        \
    else {                                                                                                            \
      for (int i = 0; i < m; ++i) sdl[i] = s->lam[i] + a_s * s->dlam[i];                                              \
      for (int k = 0; k < nb; ++k) { sdzL[k] = s->zL[k] + a_s * s->dzL[k]; sdzU[k] = s->zU[k] + a_s * s->dzU[k]; }    \
      if (pd_system_error(s, s->xt, sdl, sdzL, sdzU) <= (1.0 - 1e-4) * pd_system_error(s, s->x, s->lam, s->zL, s->zU)) \
        (result) = 2;                                                                                                 \
    }                                                                                                                 \
    if (result) {                                                                                                     \
      soft_step = 1;                                                                                                  \
      /* 'S': Ipopt repeats the dual step with the primal step size its line-search variable holds (the last failed   \
         trial, or 0 inside the soft phase) for lambda and the full step for z (:595-603); 's': the soft step size */ \
      soft_a_lam = (result) == 1 ? alpha : a_s; soft_a_z = (result) == 1 ? a_du_max : a_s;                            \
      alpha = a_s; adl = s->dlam; adzL = s->dzL; adzU = s->dzU;                                                       \
    }                                                                                                                 \
  } while (0)
      alpha = alpha_max;
      double alpha_test = alpha;
      if (in_soft_resto) { /* :426-448: only soft steps while in the soft restoration phase, at most 10 in a row */
        alpha = 0.0;
        if (++soft_resto_counter <= 10) {
          int r; TRY_SOFT(r);
          accept = r != 0;
          if (r == 1) { in_soft_resto = 0; soft_resto_counter = 0; }
        }
      } else
      while (alpha > alpha_min || n_steps == 0) {
        for (int j = 0; j < n; ++j) s->xt[j] = s->x[j] + alpha * s->dx[j];
        eval_point(s, s->xt, &s->ft, s->ct);
        double tth = vasum(s->ct, m), tbarr = barrier_obj(s, s->xt, s->ft);
        alpha_test = alpha;
        accept = check_accept(s, &R, alpha_test, tth, tbarr);
        if (accept) break;
        /* second order correction (:758-771, IpFilterLSAcceptor.cpp:473-587) */
        if (alpha == alpha_max && R.ref_theta <= tth) {
          int count_soc = 0; double theta_soc_old = 0.0, theta_trial = tth, alpha_soc = alpha;
          memcpy(csoc, s->c, sizeof(double) * (size_t)m);
          while (count_soc < 4 && !accept && (count_soc == 0 || theta_trial <= 0.99 * theta_soc_old)) {
            theta_soc_old = theta_trial;
            for (int i = 0; i < m; ++i) csoc[i] = s->ct[i] + alpha_soc * csoc[i];
            pd_solve(s, rx, csoc, rzL, rzU, sdx, sdl, sdzL, sdzU);
            alpha_soc = frac_to_bound_primal(s, sdx, s->tau);
            for (int j = 0; j < n; ++j) s->xt[j] = s->x[j] + alpha_soc * sdx[j];
            eval_point(s, s->xt, &s->ft, s->ct);
            theta_trial = vasum(s->ct, m); tbarr = barrier_obj(s, s->xt, s->ft);
            accept = check_accept(s, &R, alpha_test, theta_trial, tbarr);
            if (accept) { alpha = alpha_soc; adl = sdl; adzL = sdzL; adzU = sdzU; }
            else ++count_soc;
          }
          if (accept) break;
        }
        alpha *= 0.5; ++n_steps;
      }
      if (accept && !soft_step) { /* UpdateForNextIteration :800-813 */
        double tbarr = barrier_obj(s, s->xt, s->ft);
        if (!is_ftype(&R, alpha_test) || !armijo(&R, alpha_test, tbarr))
          filter_add(s, R.ref_barr - 1e-8 * R.ref_theta, (1.0 - 1e-5) * R.ref_theta);
      }
      if (!accept && !in_soft_resto) { /* :498-530: start the soft restoration phase; the current point enters the filter */
        filter_add(s, R.ref_barr - 1e-8 * R.ref_theta, (1.0 - 1e-5) * R.ref_theta);
        int r; TRY_SOFT(r);
        accept = r != 0;
        if (r == 2) in_soft_resto = 1;
      }
#undef TRY_SOFT
      info_ls = n_steps + 1;
    }
    if (!accept) { status = -2; break; } /* would enter the restoration phase proper (not restated) */

    /* ---- dual step (PerformDualStep :852-941) ---- */
    double alpha_du = soft_step ? soft_a_z : frac_to_bound_dual(s, adzL, adzU, s->tau);
    const double alpha_lam = soft_step ? soft_a_lam : alpha;
    for (int k = 0; k < nb; ++k) { s->zL[k] += alpha_du * adzL[k]; s->zU[k] += alpha_du * adzU[k]; }
    for (int i = 0; i < m; ++i) s->lam[i] += alpha_lam * adl[i];
    info_alpha_pr = alpha; info_alpha_du = alpha_du;
    /* ---- accept; kappa_sigma safeguard (IpIpoptAlg.cpp:623-681, 880-951) ---- */
    for (int j = 0; j < n; ++j) s->x[j] = s->xt[j];
    for (int k = 0; k < nb; ++k) {
      double sl = slackL(s, s->x, k), su = slackU(s, s->x, k), hi, lo;
      hi = 1e10 * s->mu / sl; lo = s->mu / (1e10 * sl);
      s->zL[k] = clampd(s->zL[k], lo, hi);
      hi = 1e10 * s->mu / su; lo = s->mu / (1e10 * su);
      s->zU[k] = clampd(s->zU[k], lo, hi);
    }
    ++iter;
  }

  /* ---- finalize: honor_original_bounds (IpOrigIpoptNLP.cpp:875-883), unscaled objective ---- */
  double obj = s->f / s->df;
  for (int k = 0; k < nb; ++k) {
    double v = s->x[s->b0 + k];
    s->x[s->b0 + k] = clampd(v, s->oL[k], s->oU[k]);
  }
  if (x_out) memcpy(x_out, s->x, sizeof(double) * (size_t)n);
  if (out8) { /* MPC.cpp:253-256 */
    out8[0] = s->x[s->q.xs + 1]; out8[1] = s->x[s->q.ys + 1]; out8[2] = s->x[s->q.ps + 1]; out8[3] = s->x[s->q.vs + 1];
    out8[4] = s->x[s->q.cs + 1]; out8[5] = s->x[s->q.es + 1]; out8[6] = s->x[s->q.ds]; out8[7] = s->x[s->q.as];
  }
  if (obj_out) *obj_out = obj;
  if (iters_out) *iters_out = iter;
  if (lambda_out) for (int i = 0; i < m; ++i) lambda_out[i] = s->lam[i] / s->df;
  if (trace_rows) *trace_rows = nrows;
  free(mem); free(s->ipiv);
  return status;
}
