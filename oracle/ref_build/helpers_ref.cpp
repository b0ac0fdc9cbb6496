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

}  // extern "C"
