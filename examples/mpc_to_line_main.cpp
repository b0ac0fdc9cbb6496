// The closed loop of the reference's mpc_to_line/solution/main.cpp:14-91 on the B200 library: same waypoints, same
// initial state, 50 MPC::Solve calls each fed the previous call's predicted state, same per-iteration printout.
// The matplotlib tail (:81-91) is replaced by an optional CSV dump of the cte / delta / v traces.
//
//   g++ -std=c++11 -O2 -Iinclude examples/mpc_to_line_main.cpp -Ludacitympc_b200/lib -lb200mpc
//       -Wl,-rpath,$PWD/udacitympc_b200/lib -o mpc_to_line && ./mpc_to_line [trace.csv]
#include <cmath>
#include <fstream>
#include <iostream>
#include <vector>

#include "b200mpc/MPC.h"

using b200mpc::Vec;
using std::cout;
using std::endl;
using std::vector;

int main(int argc, char** argv) {
  MPC mpc;
  int iters = 50;

  Vec ptsx(2), ptsy(2);
  ptsx[0] = -100; ptsx[1] = 100;
  ptsy[0] = -1; ptsy[1] = -1;
  // The polynomial is fitted to a straight line so a polynomial with order 1 is sufficient.
  Vec coeffs = polyfit(ptsx, ptsy, 1);

  double x = -1, y = 10, psi = 0, v = 10;
  double cte = polyeval(coeffs, x) - y;
  double epsi = psi - atan(coeffs[1]);

  Vec state(6);
  state[0] = x; state[1] = y; state[2] = psi; state[3] = v; state[4] = cte; state[5] = epsi;

  vector<double> x_vals = {state[0]}, y_vals = {state[1]}, psi_vals = {state[2]}, v_vals = {state[3]},
                 cte_vals = {state[4]}, epsi_vals = {state[5]}, delta_vals, a_vals;

  for (int i = 0; i < iters; ++i) {
    cout << "Iteration " << i << endl;
    vector<double> vars = mpc.Solve(state, coeffs);
    x_vals.push_back(vars[0]); y_vals.push_back(vars[1]); psi_vals.push_back(vars[2]); v_vals.push_back(vars[3]);
    cte_vals.push_back(vars[4]); epsi_vals.push_back(vars[5]); delta_vals.push_back(vars[6]); a_vals.push_back(vars[7]);
    for (int k = 0; k < 6; ++k) state[k] = vars[k];
    cout << "x = " << vars[0] << endl;
    cout << "y = " << vars[1] << endl;
    cout << "psi = " << vars[2] << endl;
    cout << "v = " << vars[3] << endl;
    cout << "cte = " << vars[4] << endl;
    cout << "epsi = " << vars[5] << endl;
    cout << "delta = " << vars[6] << endl;
    cout << "a = " << vars[7] << endl;
    cout << endl;
  }

  if (argc > 1) {   // what the reference plots: CTE, Delta (Radians), Velocity
    std::ofstream f(argv[1]);
    f << "step,cte,delta,v\n";
    for (size_t i = 0; i < delta_vals.size(); ++i) f << i << "," << cte_vals[i] << "," << delta_vals[i] << "," << v_vals[i] << "\n";
  }
  return 0;
}
