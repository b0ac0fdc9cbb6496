// Drop-in C++ host layer over the C ABI (include/b200mpc.h) with the reference's own interface:
//
//   class MPC { MPC(); virtual ~MPC(); std::vector<double> Solve(const VectorXd& x0, const VectorXd& coeffs); }
//                                                              mpc_to_line/src/MPC.h:7-17 (impl. solution/MPC.cpp:149-257)
//   double   polyeval(const VectorXd& coeffs, double x)        mpc_to_line/src/helpers.h:13-19
//   VectorXd polyfit(const VectorXd& x, const VectorXd& y, int order)          helpers.h:24-44
//   VectorXd globalKinematic(const VectorXd& state, const VectorXd& actuators, double dt)
//                                                              global_kinematic_model/solution/main.cpp:36-62
//
// The functions are templates over the vector type, so mpc_to_line/solution/main.cpp compiles unchanged against this
// header with Eigen::VectorXd (`#include "Eigen-3.3/Eigen/Core"` before it, as the reference does), and programs
// without Eigen can use std::vector<double> or b200mpc::Vec.  Every call runs on the GPU through libb200mpc.so; there
// is no CPU implementation behind these names.
#ifndef B200MPC_MPC_H
#define B200MPC_MPC_H

#include <cassert>
#include <cstddef>
#include <iostream>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../b200mpc.h"

namespace b200mpc {

// Minimal dense vector with the slice of Eigen::VectorXd's interface the reference's drivers use.
class Vec {
 public:
  Vec() {}
  explicit Vec(std::size_t n) : d_(n, 0.0) {}
  Vec(std::initializer_list<double> v) : d_(v) {}
  std::size_t size() const { return d_.size(); }
  double& operator[](std::size_t i) { return d_[i]; }
  double operator[](std::size_t i) const { return d_[i]; }
  double& operator()(std::size_t i) { return d_[i]; }
  double operator()(std::size_t i) const { return d_[i]; }
  const double* data() const { return d_.data(); }
  double* data() { return d_.data(); }

 private:
  std::vector<double> d_;
};

namespace detail {
template <class V>
inline std::vector<double> to_std(const V& v) {
  std::vector<double> o((std::size_t)v.size());
  for (std::size_t i = 0; i < o.size(); ++i) o[i] = v[i];
  return o;
}
inline void check(int rc) {
  if (rc != 0) throw std::runtime_error(std::string("b200mpc: ") + b200mpc_last_error());
}
// process-wide handle for the free helper functions
inline b200mpc_handle* shared_handle() {
  struct Holder {
    b200mpc_handle* h = nullptr;
    Holder() {
      b200mpc_params p;
      b200mpc_default_params(&p);
      check(b200mpc_create(&p, 0, &h));
    }
    ~Holder() { b200mpc_destroy(h); }
  };
  static Holder holder;
  return holder.h;
}
}  // namespace detail
}  // namespace b200mpc

class MPC {
 public:
  // The reference configures N, dt, Lf, ref_v by editing file-scope globals (MPC.cpp:14-31); here they are
  // b200mpc_params with the same defaults.
  MPC() : MPC(default_params()) {}
  explicit MPC(const b200mpc_params& p, int device = 0) : params_(p) {
    b200mpc::detail::check(b200mpc_create(&params_, device, &h_));
  }
  virtual ~MPC() { b200mpc_destroy(h_); }
  MPC(const MPC&) = delete;
  MPC& operator=(const MPC&) = delete;

  static b200mpc_params default_params() {
    b200mpc_params p;
    b200mpc_default_params(&p);
    return p;
  }

  // Solve the model given an initial state.  Return the next state and actuations as a vector
  // {x1, y1, psi1, v1, cte1, epsi1, delta0, a0}  (MPC.cpp:253-256).  Like the reference (MPC.cpp:248-249) the solver
  // status is not checked here; last_status() exposes it.  The reference prints "Cost <obj>" on every call
  // (MPC.cpp:251-252); set print_cost(false) to silence it.
  //
  // Reference polynomial: FG_eval reads coeffs[0] and coeffs[1] only (MPC.cpp:117-118: f0 = coeffs[0] + coeffs[1] * x0,
  // psides0 = atan(coeffs[1])) whatever the length of the vector, and so does this call by default.  The degree-2/3
  // reference of BASELINE configs[3] (f = polyeval(coeffs, x), psides = atan(f'(x))) is an EXTENSION of the reference
  // behaviour: switch it on with use_full_polynomial(true) (at most 4 coefficients), or use SolveBatch / the C ABI,
  // where ncoef is explicit.
  template <class V1, class V2>
  std::vector<double> Solve(const V1& x0, const V2& coeffs) {
    const std::vector<double> s = b200mpc::detail::to_std(x0);
    std::vector<double> c = b200mpc::detail::to_std(coeffs);
    if (s.size() != 6) throw std::invalid_argument("MPC::Solve: state must have 6 entries");
    if (c.size() < 2) throw std::invalid_argument("MPC::Solve: coeffs must have at least 2 entries (MPC.cpp:117-118)");
    if (!full_polynomial_) c.resize(2);
    else if (c.size() > B200MPC_MAX_COEFFS) throw std::invalid_argument("MPC::Solve: the full-polynomial extension takes at most 4 coefficients");
    std::vector<double> out(8);
    b200mpc::detail::check(b200mpc_solve_batch(h_, 1, s.data(), c.data(), (int)c.size(), out.data(), nullptr, &cost_, &status_, &iters_));
    if (print_cost_) std::cout << "Cost " << cost_ << std::endl;
    return out;
  }

  // B problems at once: states B x 6, coeffs B x ncoef (row major) -> out8 B x 8
  std::vector<double> SolveBatch(int B, const std::vector<double>& states, const std::vector<double>& coeffs, int ncoef,
                                 std::vector<int>* status = nullptr, std::vector<double>* cost = nullptr) {
    std::vector<double> out((std::size_t)B * 8);
    if (status) status->resize(B);
    if (cost) cost->resize(B);
    b200mpc::detail::check(b200mpc_solve_batch(h_, B, states.data(), coeffs.data(), ncoef, out.data(), nullptr,
                                               cost ? cost->data() : nullptr, status ? status->data() : nullptr, nullptr));
    return out;
  }

  int last_status() const { return status_; }
  int last_iterations() const { return iters_; }
  double last_cost() const { return cost_; }
  void print_cost(bool on) { print_cost_ = on; }
  void use_full_polynomial(bool on) { full_polynomial_ = on; }
  b200mpc_handle* handle() const { return h_; }

 private:
  b200mpc_params params_;
  b200mpc_handle* h_ = nullptr;
  int status_ = 0, iters_ = 0;
  double cost_ = 0.0;
  bool print_cost_ = true;
  bool full_polynomial_ = false;
};

// The centre line CustomMPC's constructor builds from the roadmap file (mpc_to_line/src/custom_MPC.cpp:201-212 with the
// reader of custom_MPC.h:25-44): cl_x = column 4, cl_y = column 5, cl_phi = atan(column 6).  float_fields = true converts
// the fields through single precision like the reference's std::stof.
inline void readRoadmapCenterline(const std::string& roadmap_file_name, std::vector<double>& cl_x, std::vector<double>& cl_y,
                                  std::vector<double>& cl_phi, bool float_fields = true) {
  int n = 0;
  b200mpc::detail::check(b200mpc_read_roadmap_csv(roadmap_file_name.c_str(), float_fields ? 1 : 0, nullptr, nullptr, 0, &n));
  std::vector<double> cl((std::size_t)2 * n), slope((std::size_t)n);
  b200mpc::detail::check(b200mpc_read_roadmap_csv(roadmap_file_name.c_str(), float_fields ? 1 : 0, cl.data(), slope.data(), n, &n));
  cl_x.resize(n); cl_y.resize(n); cl_phi.resize(n);
  for (int i = 0; i < n; ++i) { cl_x[i] = cl[2 * i]; cl_y[i] = cl[2 * i + 1]; cl_phi[i] = std::atan(slope[i]); }
}

// helpers.h:13-19
template <class V>
inline double polyeval(const V& coeffs, double x) {
  const std::vector<double> c = b200mpc::detail::to_std(coeffs);
  double y = 0.0;
  b200mpc::detail::check(b200mpc_polyeval_batch(b200mpc::detail::shared_handle(), 1, c.data(), (int)c.size(), &x, &y));
  return y;
}

// helpers.h:24-44
template <class V>
inline V polyfit(const V& xvals, const V& yvals, int order) {
  assert(xvals.size() == yvals.size());
  assert(order >= 1 && order <= (int)xvals.size() - 1);
  const std::vector<double> x = b200mpc::detail::to_std(xvals), y = b200mpc::detail::to_std(yvals);
  std::vector<double> c((std::size_t)order + 1);
  b200mpc::detail::check(b200mpc_polyfit_batch(b200mpc::detail::shared_handle(), 1, x.data(), y.data(), (int)x.size(), order, c.data()));
  V out(order + 1);
  for (int i = 0; i <= order; ++i) out[i] = c[i];
  return out;
}

// global_kinematic_model/solution/main.cpp:36-62 (Lf = 2 there, :15)
template <class V>
inline V globalKinematic(const V& state, const V& actuators, double dt, double Lf = 2.0) {
  const std::vector<double> s = b200mpc::detail::to_std(state), a = b200mpc::detail::to_std(actuators);
  std::vector<double> o(4);
  b200mpc::detail::check(b200mpc_rollout_batch(b200mpc::detail::shared_handle(), 1, 1, s.data(), a.data(), dt, Lf, o.data()));
  V out(4);
  for (int i = 0; i < 4; ++i) out[i] = o[i];
  return out;
}

#endif  // B200MPC_MPC_H
