// TEST INFRASTRUCTURE ONLY (oracle/_ref) -- never linked into or called by the product path.
//
// C entry points around the reference's own polyfit / polyeval / globalKinematic, compiled FROM THE
// SOURCES WHERE THEY LIE under /root/reference (nothing is copied into this repo):
//   polyeval, polyfit   /root/reference/mpc_to_line/src/helpers.h:13-19, 24-44  (+ vendored Eigen 3.3.3)
//   globalKinematic     /root/reference/global_kinematic_model/solution/main.cpp:36-62 (Lf = 2, :15)
// The quiz file defines main(); it is renamed on inclusion so only the function is used.
#include "Eigen-3.3/Eigen/QR"  // as mpc_to_line/solution/main.cpp:2 does before helpers.h
#include "helpers.h"        // -I/root/reference/mpc_to_line/src

#define main ref_global_kinematic_quiz_main
#define Lf ref_global_kinematic_Lf
#include "solution/main.cpp"  // -I/root/reference/global_kinematic_model
#undef Lf
#undef main

extern "C" {

void ref_polyfit(const double* xs, const double* ys, int m, int order, double* coeffs_out) {
  Eigen::VectorXd xv = Eigen::Map<const Eigen::VectorXd>(xs, m);
  Eigen::VectorXd yv = Eigen::Map<const Eigen::VectorXd>(ys, m);
  Eigen::VectorXd c = polyfit(xv, yv, order);
  for (int i = 0; i <= order; ++i) coeffs_out[i] = c[i];
}

double ref_polyeval(const double* coeffs, int ncoef, double x) {
  Eigen::VectorXd c = Eigen::Map<const Eigen::VectorXd>(coeffs, ncoef);
  return polyeval(c, x);
}

void ref_global_kinematic(const double* state4, const double* act2, double dt, double* next4) {
  Eigen::VectorXd s = Eigen::Map<const Eigen::VectorXd>(state4, 4);
  Eigen::VectorXd a = Eigen::Map<const Eigen::VectorXd>(act2, 2);
  Eigen::VectorXd n = globalKinematic(s, a, dt);
  for (int i = 0; i < 4; ++i) next4[i] = n[i];
}

// loops over the calls above (the CPU baseline of bench_io.py times these, so that the per-call cost is the
// reference's own code and not the Python binding): xs, ys [B][m] -> coeffs [B][order+1]; state4 [B][4], act2 [B][2]
void ref_polyfit_batch(const double* xs, const double* ys, int B, int m, int order, double* coeffs_out) {
  for (int b = 0; b < B; ++b) ref_polyfit(xs + (size_t)b * m, ys + (size_t)b * m, m, order, coeffs_out + (size_t)b * (order + 1));
}
void ref_global_kinematic_batch(const double* state4, const double* act2, int B, double dt, double* next4) {
  for (int b = 0; b < B; ++b) ref_global_kinematic(state4 + (size_t)b * 4, act2 + (size_t)b * 2, dt, next4 + (size_t)b * 4);
}

}  // extern "C"
