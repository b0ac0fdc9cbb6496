/* b200mpc -- C ABI of the B200-native batched nonlinear-MPC solver (libb200mpc.so).
 *
 * This is the drop-in boundary for the reference's mpc_to_line hot path.  Every entry point cites the
 * reference interface it replaces (paths relative to the reference checkout, cyscgzx33/UdacityMPC):
 *
 *   b200mpc_solve_batch        class MPC { std::vector<double> Solve(const VectorXd& x0, const VectorXd& coeffs); }
 *                              mpc_to_line/src/MPC.h:7-17, implementation mpc_to_line/solution/MPC.cpp:149-257
 *                              (FG_eval :45-140; CppAD::ipopt::solve :241-243)
 *   b200mpc_polyfit_batch      VectorXd polyfit(const VectorXd& xvals, const VectorXd& yvals, int order)
 *                              mpc_to_line/src/helpers.h:24-44
 *   b200mpc_polyeval_batch     double polyeval(const VectorXd& coeffs, double x)     mpc_to_line/src/helpers.h:13-19
 *   b200mpc_rollout_batch      VectorXd globalKinematic(const VectorXd& state, const VectorXd& actuators, double dt)
 *                              global_kinematic_model/solution/main.cpp:18-19, 36-62
 *   b200mpc_closed_loop_batch  the feed-forward loop of mpc_to_line/solution/main.cpp:51-76
 *   b200mpc_roadmap_reference_batch  nearest centre-line point + local-frame fit: the producer of (coeffs, cte, epsi)
 *                              that mpc_to_line/src/custom_MPC.cpp:177-212 attempts (solution/main.cpp:25-37 for a line)
 *   b200mpc_params             the file-scope globals of MPC.cpp:14-31 (N, dt, Lf, ref_v), the unit cost weights of
 *                              :57-76, the actuator bounds of :194-203, and the Ipopt options tol / max_iter that
 *                              the option string at :232-235 leaves at their defaults
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross this boundary; every function returns 0 on success or a
 *     negative b200mpc_error, and b200mpc_last_error() describes the most recent failure of the calling thread.
 *   - *_batch functions take HOST buffers in the reference's per-problem ("array of structures") order and do the
 *     host<->device copies themselves; *_batch_device functions take DEVICE buffers in the structure-of-arrays
 *     layout the kernels consume (field-major: element k of problem b at ptr[k*B + b]) and are asynchronous on
 *     the given cudaStream_t (passed as void*; NULL = the handle's own stream).
 *   - a handle is bound to one CUDA device and is the unit of thread safety: one handle per host thread AND one solve
 *     in flight per handle (every call on a handle uses the handle's one workspace and its auxiliary streams; a second
 *     *_device call on the same handle, on whatever stream, first waits on the device for the previous one).
 *   - there is no CPU fallback: without a CUDA device b200mpc_create fails with B200MPC_ERR_CUDA.
 *   - per-problem status values are Ipopt's ApplicationReturnStatus numbers
 *     (Ipopt-3.12.7/Ipopt/src/Interfaces/IpReturnCodes_inc.h:16-39): 0 Solve_Succeeded, 1 Solved_To_Acceptable_Level,
 *     3 Search_Direction_Becomes_Too_Small, 4 Diverging_Iterates, -1 Maximum_Iterations_Exceeded,
 *     -2 Restoration_Failed (the line search failed at an almost feasible point, or b200mpc_set_restoration mode 0),
 *     -3 Error_In_Step_Computation, -13 Invalid_Number_Detected (NaN / Inf in the inputs).  Like MPC::Solve (MPC.cpp:248-249) the solution is returned regardless.
 */
#ifndef B200MPC_H
#define B200MPC_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200mpc_params {
  int N;            /* horizon length, MPC.cpp:14 (25) */
  double dt;        /* MPC.cpp:15 (0.05) */
  double Lf;        /* MPC.cpp:27 (2.67) */
  double ref_v;     /* MPC.cpp:31 (40) */
  double w_cte, w_epsi, w_v;     /* MPC.cpp:60-64 (1,1,1) */
  double w_delta, w_a;           /* MPC.cpp:67-70 (1,1) */
  double w_ddelta, w_da;         /* MPC.cpp:73-76 (1,1) */
  double delta_max; /* MPC.cpp:194-195 (0.436332, the literal) */
  double a_max;     /* MPC.cpp:200-201 (1.0) */
  double tol;       /* Ipopt "tol" default 1e-8 */
  int max_iter;     /* Ipopt "max_iter" default 3000 */
} b200mpc_params;

typedef struct b200mpc_handle b200mpc_handle;

enum b200mpc_error {
  B200MPC_OK = 0,
  B200MPC_ERR_ARG = -1,     /* bad argument (null pointer, size out of range, unsupported degree) */
  B200MPC_ERR_CUDA = -2,    /* CUDA runtime error, or no CUDA device */
  B200MPC_ERR_NOMEM = -3    /* device or host allocation failed */
};

#define B200MPC_MAX_COEFFS 4      /* reference polynomial degree <= 3 in the solver */
#define B200MPC_MAX_FIT_POINTS 16 /* polyfit: m <= 16, 1 <= order <= min(m-1, 7) */
#define B200MPC_MAX_FIT_ORDER 7

/* Fills *p with the values the reference hard-codes (see b200mpc_params). */
void b200mpc_default_params(b200mpc_params* p);

/* Creates a solver bound to CUDA device `device`.  Work buffers grow on demand. */
int b200mpc_create(const b200mpc_params* p, int device, b200mpc_handle** out);
void b200mpc_destroy(b200mpc_handle* h);
const char* b200mpc_last_error(void);
/* 8N-2: length of the full variable vector [x|y|psi|v|cte|epsi|delta|a] (MPC.cpp:36-43, 161). */
int b200mpc_num_vars(const b200mpc_handle* h);

/* B independent MPC::Solve calls.
 *   state6  B x 6   (x, y, psi, v, cte, epsi) per problem                 MPC.cpp:152-157
 *   coeffs  B x ncoef, ascending powers, 2 <= ncoef <= 4                  MPC.cpp:117-118 (degree 1 as shipped)
 *   out8    B x 8   (x1, y1, psi1, v1, cte1, epsi1, delta0, a0)           MPC.cpp:253-256
 *   traj    B x (8N-2) full solution in the reference's variable order, or NULL
 *   obj     B       objective value (the "Cost" line, MPC.cpp:251-252), or NULL
 *   status  B       see above, or NULL;   iters  B  interior-point iterations, or NULL */
int b200mpc_solve_batch(b200mpc_handle* h, int B, const double* state6, const double* coeffs, int ncoef, double* out8,
                        double* traj, double* obj, int* status, int* iters);

/* b200mpc_solve_batch without the final wait: the copies in, the solve and the copies out are queued on the handle's own
 * stream and the call returns (truly asynchronous only with pinned host buffers).  The input buffers must stay valid
 * and the output buffers untouched until b200mpc_wait(h) -- which waits for everything queued on the handle -- or a
 * later blocking call on the handle returns.  Lets one host thread keep several handles busy. */
int b200mpc_solve_batch_async(b200mpc_handle* h, int B, const double* state6, const double* coeffs, int ncoef,
                              double* out8, double* traj, double* obj, int* status, int* iters);
int b200mpc_wait(b200mpc_handle* h);

/* Same on device buffers, field-major: d_state6[k*B+b], d_coeffs[i*B+b], d_out8[k*B+b], d_traj[i*B+b]. */
int b200mpc_solve_batch_device(b200mpc_handle* h, int B, const double* d_state6, const double* d_coeffs, int ncoef,
                               double* d_out8, double* d_traj, double* d_obj, int* d_status, int* d_iters,
                               void* stream);

/* The same batch sharded by contiguous index ranges over several handles (normally one per device; several handles on
 * one device are allowed), no inter-device communication: handle g solves
 * problems [g * (B / n), (g + 1) * (B / n)), the last one also the remainder.  Shard 0 is queued and waited for by the
 * calling thread, every other shard by a persistent host thread of the library (created on first use), so the devices
 * start together; one multi call at a time uses those threads (concurrent multi calls are serialised).  Host buffers as
 * b200mpc_solve_batch; pinned ones (cudaHostAlloc / cudaHostRegister) keep the copies asynchronous.
 * Every handle must have been created with the same b200mpc_params and the same restoration mode, and no handle may
 * appear twice (B200MPC_ERR_ARG otherwise: the result rows of the shards would not line up, or two shards would share
 * one workspace). */
int b200mpc_solve_batch_multi(b200mpc_handle* const* hs, int n_handles, int B, const double* state6,
                              const double* coeffs, int ncoef, double* out8, double* traj, double* obj, int* status,
                              int* iters);

/* Closed loop of solution/main.cpp:51-76 for B vehicles: `steps` consecutive solves, each fed the previous solve's
 * predicted state out8[0..5]; the state never leaves the device between steps.
 *   state6  B x 6 initial states, coeffs B x ncoef (constant over the loop)
 *   hist8   steps x B x 8 the returned vector of every step;  cost  steps x B (or NULL);  iters steps x B (or NULL) */
int b200mpc_closed_loop_batch(b200mpc_handle* h, int B, int steps, const double* state6, const double* coeffs,
                              int ncoef, double* hist8, double* cost, int* iters);

/* Warm start for b200mpc_closed_loop_batch (NOT reference behaviour: the reference cold-starts every call,
 * MPC.cpp:167-177, and so does this library unless enabled here).  When enabled, every step after the first starts from
 * the previous step's solution shifted by one stage, with the barrier parameter at mu_init (e.g. 1e-4).  The optimum
 * reached is the same local optimum within the solver tolerance; the iterates, iteration counts and last digits of
 * weakly active bounds differ from a cold start. */
int b200mpc_set_warm_start(b200mpc_handle* h, int enable, double mu_init);

/* Internal concurrency of one large solve call: the batch is cut into `parts` (1..4) contiguous sub-batches whose
 * kernels run concurrently on internal streams (joined before the call's stream continues, so the call keeps its
 * stream-ordered semantics).  Results do not depend on it for a fixed number of rounds (with the automatic setting of
 * b200mpc_set_solver_mode the hand-over to the cooperative kernel adapts to the sub-batch, which can change the last
 * bit of the few problems that finish on the other side of it).  Default 4: best for a caller that issues one call at a
 * time; a caller that already overlaps several calls on several handles / streams should set 1. */
int b200mpc_set_batch_split(b200mpc_handle* h, int parts);

/* Hand-over of the thin tail of a batch to the cooperative (warp-per-problem) kernel: after every round from
 * `from_round` on, once the batch has been compacted to at most `occupied_slots` problems, that kernel finishes them
 * (0 = only after the last round).  The cooperative kernel advances a problem 6x faster than the sweeps do but costs
 * 12x the machine time per iteration, so the threshold trades the latency of a lone call against the throughput of
 * overlapped calls: default 1184 from round 14 (its resident warps on a B200; best for one call at a time); a caller
 * that overlaps several calls on several handles / streams should set about 64-256 (measured with 64, 65 536 / 8 192
 * problems per call: +2 % / +11 % solves/s; a lone 8 192-problem call gets 13 % slower with 0).  Results do not depend on it beyond the last bits a
 * hand-over may change. */
int b200mpc_set_handover(b200mpc_handle* h, int occupied_slots, int from_round);

/* Pipelined solves on one handle (the generalisation of the one-call-at-a-time loop of solution/main.cpp:51-54 to a
 * stream of batches).  Iteration counts have a thin, long tail (N = 25: mean 12, a few problems in 10^5 need 30-50;
 * N = 100: mean 19, 2 % need more than 100, the slowest about 1000), and a batch is done when its slowest problem is.
 * With depth > 0 every b200mpc_solve_batch_device call of more than tail_slots problems is queued in two parts: the
 * bulk runs in the handle's full-size workspace until at most tail_slots problems are still iterating; those move to one
 * of `depth` small tail contexts (used round robin) and finish there, while the next call's bulk already runs.  Calls
 * stay stream-ordered for their caller -- the outputs of a call are complete when the work queued on ITS stream is --
 * so the overlap exists between calls issued on DIFFERENT streams (the bulks are ordered among themselves by the
 * library).  Memory: depth x the workspace of tail_slots problems instead of one full workspace per overlapped call.
 * Results do not depend on it beyond the last bits a hand-over between the thread-per-problem sweeps and the
 * cooperative kernel may change.  depth = 0 (default) switches it off; it synchronises the device. */
int b200mpc_set_pipeline(b200mpc_handle* h, int depth, int tail_slots);

/* Batch compaction of the throughput path: after every round from `from_round` on, when the problems that are still
 * iterating fill at most `max_live_fraction` of the occupied workspace slots, they are moved to consecutive slots, so
 * the following rounds run full warps on whole memory sectors instead of a few lanes per warp.  Results do not
 * depend on it.  Default 0.7 from round 4; 0 switches it off. */
int b200mpc_set_compaction(b200mpc_handle* h, double max_live_fraction, int from_round);

/* What follows a failed line search (Ipopt: IpBacktrackingLineSearch.cpp:498-585 -> soft restoration phase, then
 * IpRestoMinC_1Nrm.cpp).  Never reached on the benchmark workloads at the reference's N = 25; 1 % of problems with
 * initial states far off the road and 2 % of the problems at N = 100 get there.
 *   mode 0: such a problem returns status -2 (Restoration_Failed) at the iteration where Ipopt would switch.
 *   mode 1: the restoration step -- a forward sweep removes 5 % of every constraint defect with the controls
 *           kept, the point left enters the filter, lambda is reset to 0, z moves towards mu / slack -- i.e. what Ipopt
 *           does around its restoration phase, with a closed-form point instead of Ipopt's nested restoration solve.
 *           NOT a restatement of that solve: the iteration count of such a problem differs from Ipopt's; the solution
 *           is the reference's on 346 / 346 (N = 25), 14 / 14 (N = 50), 170 / 179 (N = 100) such problems.
 *   mode 2 (default): as 1, preceded by Ipopt's soft restoration phase (damped full primal-dual steps accepted on the primal-dual
 *           system error, IpBacktrackingLineSearch.cpp:1043-1140), restated exactly: the few problems on which Ipopt
 *           takes such steps then keep its iterates and iteration count. */
int b200mpc_set_restoration(b200mpc_handle* h, int mode);

/* B least-squares polynomial fits (unpivoted Householder QR of the Vandermonde matrix, as Eigen 3.3.3 does for
 * helpers.h:24-44).  xs, ys: B x m;  coeffs_out: B x (order+1).  Requires 1 <= order <= m-1 (helpers.h:26 assert). */
int b200mpc_polyfit_batch(b200mpc_handle* h, int B, const double* xs, const double* ys, int m, int order,
                          double* coeffs_out);
/* Device buffers, field-major: d_xs[j*B+b], d_ys[j*B+b], d_coeffs_out[i*B+b]. */
int b200mpc_polyfit_batch_device(b200mpc_handle* h, int B, const double* d_xs, const double* d_ys, int m, int order,
                                 double* d_coeffs_out, void* stream);

/* y[b] = sum_i coeffs[b][i] * x[b]^i  (helpers.h:13-19).  coeffs: B x ncoef. */
int b200mpc_polyeval_batch(b200mpc_handle* h, int B, const double* coeffs, int ncoef, const double* x, double* y);
int b200mpc_polyeval_batch_device(b200mpc_handle* h, int B, const double* d_coeffs, int ncoef, const double* d_x,
                                  double* d_y, void* stream);

/* B bicycle-model rollouts of H Euler steps (global_kinematic_model/solution/main.cpp:36-62; H=1 is globalKinematic).
 *   state4 B x 4 (x,y,psi,v);  act B x H x 2 (delta,a);  out B x H x 4 (the state after each step). */
int b200mpc_rollout_batch(b200mpc_handle* h, int B, int H, const double* state4, const double* act, double dt,
                          double Lf, double* out);
/* Device buffers, field-major: d_state4[k*B+b], d_act[(s*2+j)*B+b], d_out[(s*4+k)*B+b]. */
int b200mpc_rollout_batch_device(b200mpc_handle* h, int B, int H, const double* d_state4, const double* d_act,
                                 double dt, double Lf, double* d_out, void* stream);

/* Roadmap front-end (the step in front of MPC::Solve): for B vehicle poses in the road's global frame,
 * nearest centre-line point (what mpc_to_line/src/custom_MPC.cpp:177-185 computes), the 6 consecutive centre-line
 * points from there, global -> vehicle frame, degree-3 polyfit, and the MPC inputs in the vehicle frame:
 * state6 = (0, 0, 0, v, cte = p(0), epsi = -atan(p'(0))) as solution/main.cpp:34-37 defines cte / epsi.
 *   pose4  B x 4 (x, y, psi, v);  centerline  n_wp x 2 (x, y), n_wp >= 6 (e.g. columns 4,5 of mpc_to_line/roadmap.csv)
 *   state6_out  B x 6;  coeffs_out  B x 4 -- ready for b200mpc_solve_batch(..., ncoef = 4, ...) */
int b200mpc_roadmap_reference_batch(b200mpc_handle* h, int B, const double* pose4, const double* centerline, int n_wp,
                                    double* state6_out, double* coeffs_out);
/* Device buffers: d_pose4[k*B+b] (field-major), d_centerline[i*2+j] (row-major), d_state6_out[k*B+b], d_coeffs_out[i*B+b]. */
int b200mpc_roadmap_reference_batch_device(b200mpc_handle* h, int B, const double* d_pose4, const double* d_centerline,
                                           int n_wp, double* d_state6_out, double* d_coeffs_out, void* stream);

/* Reads a roadmap file in the reference's format (mpc_to_line/roadmap.csv: 7 comma-separated numbers per line -- left
 * edge x, y, right edge x, y, centre line x, y, slope) the way CustomMPC::readRoadmapFromCSV / parseRoadMapLine intend to
 * (mpc_to_line/src/custom_MPC.h:25-44; the constructor at custom_MPC.cpp:201-212 keeps columns 4, 5 as the centre line
 * and atan(column 6) as its direction): one waypoint per line, fields split at ','.  The reference converts every field
 * with std::stof, i.e. through SINGLE precision; float_fields != 0 reproduces that, 0 parses doubles.  Host-only (no
 * device, no handle).  The reference's reader opens the file NAME as a string stream (custom_MPC.h:28) and so never
 * reads the file; that bug is not reproduced.
 *   centerline_out  max_wp x 2, or NULL;  slope_out  max_wp, or NULL;  *n_wp = waypoints in the file (may exceed
 *   max_wp: only the first max_wp are written) -- feed centerline_out to b200mpc_roadmap_reference_batch. */
int b200mpc_read_roadmap_csv(const char* path, int float_fields, double* centerline_out, double* slope_out, int max_wp, int* n_wp);

/* Execution mode of the solver (tuning; results do not depend on it).
 *   mode 0 (default)  throughput path + latency path.  Batches of at least `fused_below` problems (default 3072) run
 *                     rounds of the per-pass thread-per-problem kernels (factor, forward, step).  By default the
 *                     number of rounds adapts to the batch: after every round from the 14th on (at most 20) the
 *                     cooperative kernel takes the rest over as soon as the compacted batch fits one of its waves
 *                     (1184 problems); a positive `rounds` fixes the hand-over point instead;
 *                     whatever is still iterating then -- the thin tail of the batch and rare 30-50 iteration
 *                     stragglers -- is finished by the cooperative warp-per-problem kernel.  Smaller batches (e.g. the
 *                     reference's one-problem MPC::Solve call: 0.6 ms) use the cooperative kernel alone.
 *   mode 1            the fused thread-per-problem kernel alone (one launch per solve; kept for comparison)
 * rounds <= 0 / fused_below < 0 keep the current value.  The whole launch sequence of a solve is replayed as one CUDA graph. */
int b200mpc_set_solver_mode(b200mpc_handle* h, int mode, int rounds, int fused_below);

/* Measurement helpers (used by bench.py; not part of the reference interface). */
/* Off by default.  When enabled, every solve is bracketed by a pair of CUDA events on its stream (kept until
 * b200mpc_kernel_time_ms reads / resets them; at most 8192 pairs). */
int b200mpc_set_timing(b200mpc_handle* h, int enable);
/* Total device time in ms of the solver's kernels (one interval per batch: first to last kernel of a solve) since the
 * last reset, measured with CUDA events on the launching stream, and the number of intervals (0 with timing off). */
int b200mpc_kernel_time_ms(b200mpc_handle* h, double* total_ms, int* launches, int reset);
/* Sustained FP64 FMA throughput of the device in TFLOP/s (dependent-chain DFMA microbenchmark, 2 FLOP per FMA). */
int b200mpc_measure_fp64_peak(b200mpc_handle* h, double* tflops);
/* Arithmetic self-test (host arrays of n doubles): quot[i] = a[i] / b[i] and rcp[i] = 1 / b[i] as the sweeps compute
 * them -- rcp.approx.ftz.f64 refined by Newton steps, within 1 ulp of the IEEE result for operands in the normal range,
 * +-inf / 0 / NaN like the exact operation for a zero or infinite divisor -- so a caller (tests/test_gpu_parity.py) can
 * hold the library's arithmetic against its own.  No reference counterpart: Ipopt divides with the host FPU. */
int b200mpc_selftest_division(b200mpc_handle* h, int n, const double* a, const double* b, double* quot, double* rcp);
/* Number of kernels this handle has launched since creation. */
long long b200mpc_launch_count(const b200mpc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* B200MPC_H */
