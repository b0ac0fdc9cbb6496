/* TEST INFRASTRUCTURE ONLY -- the CPU oracle ("port") for the mpc_to_line hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product path (udacitympc_b200/, include/)
 * never links, imports or executes anything under oracle/.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py, -m "not gpu")
 * against
 *   (a) the reference's own in-source known answers
 *         /root/reference/global_kinematic_model/solution/main.cpp:27-30
 *         /root/reference/polyfit/solution/main.cpp:18-20,33-54
 *   (b) outputs of the reference itself run in the build container (oracle/_ref: the reference's
 *       prebuilt Ipopt 3.12.7 + MUMPS binaries driven on the MPC.cpp NLP, and the reference's
 *       helpers.h / globalKinematic compiled from where they lie), committed as fixtures under
 *       tests/golden/ by tests/golden/make_golden.py.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Problem + algorithm parameters.  Defaults (oracle_default_params) are the values the reference
 * hard-codes: MPC.cpp:14-15 (N, dt), :27 (Lf), :31 (ref_v), :57-76 (all cost weights 1),
 * :194-203 (actuator bounds) and Ipopt 3.12.7 defaults for everything the option string at
 * MPC.cpp:232-235 leaves untouched (tol 1e-8, max_iter 3000). */
typedef struct {
  int N;
  double dt, Lf, ref_v;
  double w_cte, w_epsi, w_v, w_delta, w_a, w_ddelta, w_da;
  double delta_max, a_max;
  double tol;
  int max_iter;
} oracle_params;

void oracle_default_params(oracle_params* p);

/* One MPC::Solve (MPC.cpp:149-257) by a restatement of Ipopt's interior-point algorithm with a
 * dense Bunch-Kaufman LDL^T standing in for MUMPS.
 *   state6   (x, y, psi, v, cte, epsi)              coeffs[ncoef] ascending powers
 *   x_out    full variable vector in the reference layout (MPC.cpp:36-43), 8N-2 doubles, or NULL
 *   out8     (x1,y1,psi1,v1,cte1,epsi1,delta0,a0)   (MPC.cpp:253-256)
 *   trace    optional rows of 10: iter,obj,inf_pr,inf_du,mu,d_norm,regu,alpha_du,alpha_pr,ls_trials
 * returns Ipopt's ApplicationReturnStatus numbering (0 Solve_Succeeded, 1 Solved_To_Acceptable_Level,
 * -1 Maximum_Iterations_Exceeded, -2 Restoration_Failed, -3 Error_In_Step_Computation, 3 Search_Direction_Becomes_Too_Small). */
int oracle_mpc_solve(const oracle_params* p, const double* state6, const double* coeffs, int ncoef, double* x_out,
                     double* out8, double* obj_out, int* iters_out, double* lambda_out, double* trace, int trace_cap,
                     int* trace_rows);

/* Problem functions at a point (reference layout): f, grad f (n), g (m, without subtracting the
 * initial-state right-hand side), dense Jacobian (m x n row major), dense lower Hessian of
 * sigma*f + lambda^T g (n x n row major). Any output may be NULL. */
void oracle_mpc_eval(const oracle_params* p, const double* coeffs, int ncoef, const double* x, const double* lambda,
                     double sigma, double* f, double* grad, double* g, double* jac, double* hess);

/* helpers.h:13-19 and :24-44 (Eigen 3.3.3 HouseholderQR, unblocked path HouseholderQR.h:256-287,
 * solve :350-369, reflector Householder.h:64-94). */
double oracle_polyeval(const double* coeffs, int ncoef, double x);
int oracle_polyfit(const double* xs, const double* ys, int m, int order, double* coeffs_out);

/* global_kinematic_model/solution/main.cpp:36-62 (one Euler step; Lf is a parameter: 2 there, 2.67 in MPC.cpp:27). */
void oracle_global_kinematic(const double* state4, const double* act2, double dt, double Lf, double* next4);
/* B calls of the two functions above: xs, ys [B][m] -> coeffs [B][order+1]; state4 [B][4], act2 [B][2] -> next4 [B][4] */
void oracle_polyfit_batch(const double* xs, const double* ys, int B, int m, int order, double* coeffs_out);
void oracle_global_kinematic_batch(const double* state4, const double* act2, int B, double dt, double Lf, double* next4);

#ifdef __cplusplus
}
#endif
#endif
