// Closed-loop driver for the C++ drop-in class (include/b200mpc/MPC.h): the experiment of the reference's
// mpc_to_line/solution/main.cpp:14-91 -- a straight reference line y = -1 fitted through two waypoints, the vehicle
// starting 11 m off it at 10 m/s, 50 receding-horizon steps each starting from the state the previous solve predicted
// -- with the output lines that program prints ("Iteration k", "<name> = <value>"), so the two traces can be diffed.
// The matplotlib tail of the reference (:81-91) becomes an optional CSV of the traces it plots (cte, delta, v).
//
//   g++ -std=c++11 -O2 -Iinclude examples/mpc_to_line_main.cpp -Ludacitympc_b200/lib -lb200mpc
//       -Wl,-rpath,$PWD/udacitympc_b200/lib -o mpc_to_line && ./mpc_to_line [trace.csv]
#include <array>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <vector>

#include "b200mpc/MPC.h"

namespace {

constexpr int kSteps = 50;
const char* const kNames[8] = {"x", "y", "psi", "v", "cte", "epsi", "delta", "a"};

// the 6-vector MPC::Solve takes: pose and speed plus the two tracking errors against the fitted line
b200mpc::Vec tracking_state(double x, double y, double psi, double v, const b200mpc::Vec& line) {
  b200mpc::Vec s(6);
  s[0] = x; s[1] = y; s[2] = psi; s[3] = v;
  s[4] = polyeval(line, x) - y;          // cross-track error
  s[5] = psi - std::atan(line[1]);       // heading error
  return s;
}

}  // namespace

int main(int argc, char** argv) {
  b200mpc::Vec wx(2), wy(2);
  wx[0] = -100.0; wx[1] = 100.0;
  wy[0] = wy[1] = -1.0;
  const b200mpc::Vec line = polyfit(wx, wy, 1);   // two waypoints: degree 1

  MPC controller;
  b200mpc::Vec state = tracking_state(-1.0, 10.0, 0.0, 10.0, line);
  std::vector<std::array<double, 8>> trace;   // one row per step: predicted state + first actuators
  trace.reserve(kSteps);

  for (int step = 0; step < kSteps; ++step) {
    std::cout << "Iteration " << step << std::endl;
    const std::vector<double> sol = controller.Solve(state, line);
    std::array<double, 8> row;
    for (int k = 0; k < 8; ++k) {
      row[k] = sol[k];
      std::cout << kNames[k] << " = " << sol[k] << std::endl;
    }
    std::cout << std::endl;
    trace.push_back(row);
    for (int k = 0; k < 6; ++k) state[k] = sol[k];   // feed-forward: the next solve starts where this one predicts
  }

  if (argc > 1) {
    std::ofstream csv(argv[1]);
    csv << "step,cte,delta,v\n";
    for (int step = 0; step < kSteps; ++step) {
      // the reference plots cte / v including the initial state and delta per step: row k pairs the state BEFORE step k
      const double cte = step == 0 ? tracking_state(-1.0, 10.0, 0.0, 10.0, line)[4] : trace[step - 1][4];
      const double v = step == 0 ? 10.0 : trace[step - 1][3];
      csv << step << "," << cte << "," << trace[step][6] << "," << v << "\n";
    }
  }
  return 0;
}
