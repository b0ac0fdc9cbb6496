// Interior-point solve kernels for sm_100a (K1 derivative evaluation, K2 Riccati KKT solve, K3 barrier / line
// search / filter / convergence logic).  Replaces CppAD::ipopt::solve at
// /root/reference/mpc_to_line/solution/MPC.cpp:241-243.
//
// Throughput path (mpc_core.cuh): one problem per thread; every per-problem vector lives in a warp-interleaved HBM
// workspace (each workspace access of a warp is one coalesced 256-byte row) and the 7x7 Riccati blocks live in
// registers.  One interior-point iteration = three sweeps over the horizon, each its own kernel
//   init | factor | forward | step        launched round after round on one stream (replayed as one CUDA graph),
// with its own register budget (the Riccati factorisation needs ~250 registers, the step sweep 228 with its
// stage-to-stage values in shared memory, the forward sweep 168) and a small instruction footprint.
// Between rounds mpc_repack_kernel compacts the problems that are still iterating into consecutive workspace slots
// (second workspace region, slot -> problem map, level state in device memory), so late rounds run full warps.
// Latency path (mpc_coop.cuh): one problem per warp, lane <-> stage.  mpc_coop_kernel finishes whatever is still
// iterating once the compacted batch fits one of its waves (or after the last round) and solves small batches on its own.
// mpc_fused_kernel loops the three thread-per-problem sweeps in one launch (comparison / fallback for horizons whose
// per-stage scratch does not fit the cooperative kernel's shared memory).
#include "kernels.h"
#include "mpc_coop.cuh"

namespace b200mpc {

#ifndef MPC_BLOCK
#define MPC_BLOCK 64
#endif
constexpr int kBlock = MPC_BLOCK;
// resident blocks per SM the light sweeps are compiled for (register cap = 65536 / (64 * blocks))
#ifndef MPC_FWD_BLOCKS
#define MPC_FWD_BLOCKS 6
#endif
// The step sweep runs at 4 blocks per SM: 228 registers, nothing spilled.  At 6 blocks (168 registers, 240 B of spill
// stores and 352 B of spill loads per thread and stage) it is 3 % slower with one caller and with overlapped callers
// (gpurun_out/r2_sb*.json: 13.59 -> 14.05 M solves/s; 5 blocks: 13.57 M) -- the sweep is bound by its dependent
// instruction chains, not by the number of resident warps.
#ifndef MPC_STEP_BLOCKS
#define MPC_STEP_BLOCKS 4
#endif
#ifndef MPC_FACTOR_BLOCKS
#define MPC_FACTOR_BLOCKS 4
#endif
// stage-to-stage values of the step sweep: in registers (250 registers, nothing spilled, no shared memory; 0.6-0.8 % faster
// than the shared-memory carry at 4 blocks per SM, gpurun_out/r2_creg_*.json, r2_psd_*.json) or in shared memory
#ifndef MPC_STEP_CARRY_SMEM
#define MPC_STEP_CARRY_SMEM 0
#endif
#ifndef MPC_STEPFACTOR_BLOCKS
#define MPC_STEPFACTOR_BLOCKS 4
#endif
constexpr int kFwdBlocks = MPC_FWD_BLOCKS, kStepBlocks = MPC_STEP_BLOCKS, kFactorBlocks = MPC_FACTOR_BLOCKS;

// One workspace region holds every problem of a batch; the buffer the handle allocates has two of them (the batch
// compaction moves the unfinished problems back and forth), two slot -> problem maps and the counters.
constexpr int kMaxParts = 4;
// per-part compaction state in device memory (ints)
enum Desc { kLevel = 0, kCount, kLive, kAlloc, kTicket, kHanded, kDescInts = 8 };
static size_t region_doubles(int N, int B) {
  size_t groups = ((size_t)B + 31) / 32;
  return groups * (size_t)workspace_doubles_per_problem(N) * 32;
}
static size_t aux_doubles(int B) { return ((size_t)B + 1) / 2 * 2 + kMaxParts * kDescInts; }   // 2 int maps + state
size_t solve_workspace_doubles(int N, int B) { return 2 * region_doubles(N, B) + aux_doubles(B); }

struct SolveArgs {
  int B, steps, step, ncoef;
  int io_aos;             // layout of state6 / coeffs / out8: 0 = field-major [k][B] (device callers), 1 = the reference's
                          // per-problem order [B][k] (host entry points: no separate transpose kernels, K6 is fused here)
  int b0, b1;             // this launch covers problems [b0, b1) of the B-problem batch (B stays the array stride)
  int warm;               // closed loop: steps > 0 start from the shifted previous solution
  double mu_warm;
  const double* state6;   // [6][B]
  const double* coeffs;   // [ncoef][B]
  double* ws;             // workspace region 0: slot == problem (level 0)
  // batch compaction (desc == null: off).  Level k >= 1 lives in region k & 1 with the slot -> problem map (k - 1) & 1;
  // desc = {level, occupied slots, live problems counted by the step kernel, slot allocator, block ticket}
  double* ws1;
  int* map0;
  int* map1;
  int* desc;
  int count_live;         // step kernel: count the problems that are not finished (a compaction attempt follows)
  double* out8;           // [steps][8][B]
  double* traj;           // [8N-2][B] or null (last step)
  double* obj;            // [steps][B] or null
  int* status;            // [B] or null (last step)
  int* iters;             // [steps][B] or null
};

__device__ __forceinline__ double* slot_base(const Params& P, double* ws, int slot) {
  return ws + (size_t)(slot >> 5) * (size_t)workspace_doubles_per_problem(P.N) * 32 + (slot & 31);
}
// thread (or warp) i of a launch -> workspace slot (+ its region) and problem index; false: nothing to do
__device__ __forceinline__ bool locate(const SolveArgs& A, int i, int& slot, int& b, double*& ws, bool sweeps = true) {
  slot = A.b0 + i;
  const int level = A.desc ? A.desc[kLevel] : 0;
  if (sweeps && level > 0 && A.desc[kHanded]) return false;   // the cooperative kernel has taken the rest over
  if (level == 0) {
    if (slot >= A.b1) return false;
    b = slot; ws = A.ws;
  } else {
    if (i >= A.desc[kCount]) return false;
    b = ((level & 1) ? A.map0 : A.map1)[slot];
    ws = (level & 1) ? A.ws1 : A.ws;
  }
  return true;
}
__device__ __forceinline__ void load_coeffs(const SolveArgs& A, int b, double* cf) {
  const size_t sb = A.io_aos ? (size_t)A.ncoef : 1, sk = A.io_aos ? 1 : (size_t)A.B;
#pragma unroll
  for (int i = 0; i < kMaxCoef; ++i) cf[i] = i < A.ncoef ? A.coeffs[(size_t)b * sb + (size_t)i * sk] : 0.0;
}
// element k of the returned 8-vector of problem b at closed-loop step `step`
__device__ __forceinline__ size_t out8_index(const SolveArgs& A, int step, int k, int b) {
  return (size_t)step * 8 * A.B + (A.io_aos ? (size_t)b * 8 + k : (size_t)k * A.B + b);
}
// initial state of closed-loop step `step`: the caller's state for step 0, else the previous step's out8[0..5]
// (solution/main.cpp:66)
__device__ __forceinline__ void load_state6(const SolveArgs& A, int b, double* s0) {
#pragma unroll
  for (int k = 0; k < 6; ++k)
    s0[k] = A.step == 0 ? A.state6[A.io_aos ? (size_t)b * 6 + k : (size_t)k * A.B + b] : A.out8[out8_index(A, A.step - 1, k, b)];
}
__device__ __forceinline__ void write_result(const Params& P, const SolveArgs& A, int b, Solver<32>& S) {
  Result R;
  S.finish(R, (A.traj && A.step == A.steps - 1) ? A.traj + b : nullptr, (size_t)A.B);
#pragma unroll
  for (int k = 0; k < 8; ++k) A.out8[out8_index(A, A.step, k, b)] = R.out8[k];
  if (A.obj) A.obj[(size_t)A.step * A.B + b] = R.obj;
  if (A.iters) A.iters[(size_t)A.step * A.B + b] = R.iters;
  if (A.status && A.step == A.steps - 1) A.status[b] = R.status;
}

// one atomic per group of converged lanes
__device__ __forceinline__ void count_live(int* counter) {
  const unsigned m = __activemask();
  if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(counter, __popc(m));
}

// ---- per-pass kernels ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) mpc_init_kernel(const __grid_constant__ Params P, const __grid_constant__ SolveArgs A) {
  const int b = A.b0 + blockIdx.x * blockDim.x + threadIdx.x;   // always launched on the uncompacted batch
  if (b >= A.b1) return;
  if (b == A.b0 && A.desc)
    for (int k = 0; k < kDescInts; ++k) A.desc[k] = 0;
  Solver<32> S(P, slot_base(P, A.ws, b), b & 31);
  double s0[6], cf[kMaxCoef];
  load_state6(A, b, s0);
  load_coeffs(A, b, cf);
  if (A.warm && A.step > 0) { S.set_coeffs(cf, A.ncoef); S.init_warm(s0, A.mu_warm); }
  else S.init(s0, cf, A.ncoef);
  S.store_state();
  if (S.phase == PH_DONE) write_result(P, A, b, S);   // invalid number at the starting point
}

__global__ void __launch_bounds__(kBlock, kFactorBlocks) mpc_factor_kernel(const __grid_constant__ Params P, const __grid_constant__ SolveArgs A) {
  int slot, b;
  double* ws;
  if (!locate(A, blockIdx.x * blockDim.x + threadIdx.x, slot, b, ws)) return;
  Solver<32> S(P, slot_base(P, ws, slot), slot & 31);
  if (S.load_phase() != PH_FACTOR) return;
#if MPC_ASYNC_STAGE
  __shared__ double stage[2 * kStageVals * kBlock];   // [buffer][value][thread]
  S.sb = stage + threadIdx.x; S.sbs = kBlock;
#endif
  load_coeffs(A, b, S.cf);
  S.kernel_factor();
  if (S.phase == PH_DONE) {   // inertia correction exhausted (IpPDPerturbationHandler.cpp:380-388)
    S.load_state();
    write_result(P, A, b, S);
  }
}

__global__ void __launch_bounds__(kBlock, kFwdBlocks) mpc_forward_kernel(const __grid_constant__ Params P, const __grid_constant__ SolveArgs A) {
  int slot, b;
  double* ws;
  if (!locate(A, blockIdx.x * blockDim.x + threadIdx.x, slot, b, ws)) return;
  Solver<32> S(P, slot_base(P, ws, slot), slot & 31);
  if (S.load_phase() != PH_FORWARD) return;
  load_coeffs(A, b, S.cf);
  S.kernel_forward();
}

__global__ void __launch_bounds__(kBlock, kStepBlocks) mpc_step_kernel(const __grid_constant__ Params P, const __grid_constant__ SolveArgs A) {
  int slot, b;
  double* ws;
  if (!locate(A, blockIdx.x * blockDim.x + threadIdx.x, slot, b, ws)) return;
  Solver<32> S(P, slot_base(P, ws, slot), slot & 31);
  const int ph = S.load_phase();
  if (ph != PH_STEP) {
    if (A.count_live && ph != PH_DONE) count_live(A.desc + kLive);
    return;
  }
#if MPC_STEP_CARRY_SMEM
  __shared__ double carry[kCarry * kBlock];   // stage-to-stage values of the sweep, [value][thread]
  S.cr = carry + threadIdx.x; S.cs = kBlock;
#else
  double carry[kCarry];
  S.cr = carry; S.cs = 1;
#endif
#if MPC_ASYNC_STAGE && MPC_STEP_ASYNC_STAGE
  __shared__ double stage[2 * kStepStageVals * kBlock];   // [buffer][value][thread]
  S.sb = stage + threadIdx.x; S.sbs = kBlock;
#endif
  load_coeffs(A, b, S.cf);
  S.kernel_step();
  if (S.phase == PH_DONE) write_result(P, A, b, S);
  else if (A.count_live) count_live(A.desc + kLive);
}

#if MPC_FUSE_FACTOR
// STEP sweep with the Riccati factorisation of the next iteration riding on it (Solver::kernel_stepfactor): the trial
// iterate is factorised from the registers that have just produced it, so an accepted step is followed by the forward
// sweep directly and the factor kernel only sees the rare cases (first iteration, rejected trial points, second-order
// corrections, inertia corrections).  Both carry sets (24 + 39 doubles per thread) live in shared memory.
__global__ void __launch_bounds__(kBlock, MPC_STEPFACTOR_BLOCKS) mpc_stepfactor_kernel(const __grid_constant__ Params P, const __grid_constant__ SolveArgs A) {
  int slot, b;
  double* ws;
  if (!locate(A, blockIdx.x * blockDim.x + threadIdx.x, slot, b, ws)) return;
  Solver<32> S(P, slot_base(P, ws, slot), slot & 31);
  const int ph = S.load_phase();
  if (ph != PH_STEP) {
    if (A.count_live && ph != PH_DONE) count_live(A.desc + kLive);
    return;
  }
  __shared__ double carry[(kCarry + kRicCarry) * kBlock];   // [value][thread]
  S.cr = carry + threadIdx.x; S.cs = kBlock;
  S.rq = carry + kCarry * kBlock + threadIdx.x; S.rqs = kBlock;
  load_coeffs(A, b, S.cf);
  S.kernel_stepfactor();
  if (S.phase == PH_DONE) write_result(P, A, b, S);
  else if (A.count_live) count_live(A.desc + kLive);
}
#endif
static void launch_step(const Params& P, const SolveArgs& A, const SolveConfig& cfg, int grid, cudaStream_t stream) {
#if MPC_FUSE_FACTOR
  if (cfg.fuse_factor) { mpc_stepfactor_kernel<<<grid, kBlock, 0, stream>>>(P, A); return; }
#endif
  (void)cfg;
  mpc_step_kernel<<<grid, kBlock, 0, stream>>>(P, A);
}

// ---- batch compaction ------------------------------------------------------------------------------------------
// Runs after a step kernel that counted the live (unfinished) problems.  If they fill at most `max_live` of the
// occupied slots they move to consecutive slots of the other region (level + 1); the following launches find the new
// level in desc.  The running lanes of a warp take consecutive destination slots, so the rows they write are
// contiguous; the slot order among warps is first come, first served (results do not depend on the slot a problem
// sits in).  The last block to finish commits the new level and clears the counters.
// Tail hand-off (pipelined solves, launch_solve_bulk / launch_solve_tail): with export_below > 0 the kernel moves the live
// problems OUT of this workspace as soon as at most export_below are left -- into region 1 of a small tail context
// (T.ws1, slot -> problem map T.map0, level 1 in T.desc) -- and leaves this workspace empty (level >= 1, no occupied
// slots: every later launch on it finds nothing to do).  The tail context is finished by its own launches while the
// next batch already runs in this workspace.
__global__ void __launch_bounds__(256) mpc_repack_kernel(const __grid_constant__ Params P, const __grid_constant__ SolveArgs A,
                                                         double max_live, int export_below, double* t_ws1, int* t_map0, int* t_desc) {
  int* desc = A.desc;
  const int level = desc[kLevel], live_n = desc[kLive];
  const int occupied = level == 0 ? A.b1 - A.b0 : desc[kCount];
  const bool exp = export_below > 0 && live_n > 0 && live_n <= export_below;
  const bool go = exp || (live_n > 0 && (double)live_n <= max_live * (double)occupied);
  if (go) {
    int slot = 0, b = 0;
    double* ws = A.ws;
    const bool in = locate(A, blockIdx.x * blockDim.x + threadIdx.x, slot, b, ws);
    const Ws<32> src{slot_base(P, ws, slot), slot & 31};
    const bool live = in && (int)src(iPHASE) != PH_DONE;
    const unsigned m = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (m != 0 && lane == 0) base = atomicAdd(desc + kAlloc, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) {
      const int rank = base + __popc(m & ((1u << lane) - 1u));
      if (exp) {
        t_map0[rank] = b;
        repack_problem(P, src, Ws<32>{slot_base(P, t_ws1, rank), rank & 31});
      } else {
        const int ds = A.b0 + rank;
        (((level + 1) & 1) ? A.map0 : A.map1)[ds] = b;
        repack_problem(P, src, Ws<32>{slot_base(P, ((level + 1) & 1) ? A.ws1 : A.ws, ds), ds & 31});
      }
    }
  }
  __syncthreads();   // every thread of the block has read desc
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(desc + kTicket, 1) == (int)gridDim.x - 1) {
      if (exp) {
        t_desc[kLevel] = 1; t_desc[kCount] = desc[kAlloc];
        desc[kCount] = 0; desc[kLevel] = level > 0 ? level : 1;
      } else if (go) { desc[kCount] = desc[kAlloc]; desc[kLevel] = level + 1; }
      desc[kLive] = 0; desc[kAlloc] = 0; desc[kTicket] = 0;
    }
  }
}
// an empty tail context: level 1 with no occupied slots (level 0 would mean "slot == problem" for every slot)
__global__ void mpc_tail_reset_kernel(int* t_desc) {
  if (threadIdx.x < kDescInts) t_desc[threadIdx.x] = threadIdx.x == kLevel ? 1 : 0;
}

// ---- fused kernel: finishes whatever is still active (fresh == 1: starts from the inputs) ---------------------
__global__ void __launch_bounds__(kBlock) mpc_fused_kernel(const __grid_constant__ Params P, const __grid_constant__ SolveArgs A, int fresh) {
  int slot, b;
  double* ws;
  if (!locate(A, blockIdx.x * blockDim.x + threadIdx.x, slot, b, ws, false)) return;
  Solver<32> S(P, slot_base(P, ws, slot), slot & 31);
  double carry[kCarry];
  S.cr = carry; S.cs = 1;
  load_coeffs(A, b, S.cf);
  if (fresh) {
    double s0[6];
    load_state6(A, b, s0);
    if (A.warm && A.step > 0) S.init_warm(s0, A.mu_warm);
    else S.init(s0, S.cf, kMaxCoef);
  } else {
    if (S.load_phase() == PH_DONE) return;
    S.load_state();
  }
  while (S.phase != PH_DONE) S.trip();
  S.store_state();
  write_result(P, A, b, S);
}

// ---- cooperative kernel: one problem per warp (latency path: small batches, the tail of a batch, stragglers) -----
struct DevExec {
  int lane;
  __device__ __forceinline__ bool lane0() const { return lane == 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  template <class F>
  __device__ __forceinline__ void for_stages(int n, F f) const {
    for (int t = lane; t < n; t += 32) f(t);
  }
};

// take_below > 0: conditional hand-over after a round -- the kernel acts only once the batch has been compacted to at
// most take_below occupied slots (two waves of the cooperative kernel); from then on the sweeps of later rounds find
// nothing to do
__global__ void __launch_bounds__(128) mpc_coop_kernel(const __grid_constant__ Params P, const __grid_constant__ SolveArgs A, int fresh,
                                                       int warps_per_block, int doubles_per_warp, int take_below) {
  extern __shared__ double coop_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (take_below > 0) {
    if (!A.desc || A.desc[kLevel] == 0 || A.desc[kCount] > take_below) return;
    if (threadIdx.x == 0) A.desc[kHanded] = 1;   // only the sweeps read it
  }
  if (wib >= warps_per_block) return;
  double* mine = coop_smem + (size_t)wib * doubles_per_warp;
  CoopStage* st = reinterpret_cast<CoopStage*>(mine);
  CoopPub* pub = reinterpret_cast<CoopPub*>(mine + (size_t)P.N * kCoopStageDoubles);
  // one problem per warp at a time; the grid is capped at a few waves and every warp walks on through the slots (a
  // launch that finds little or nothing to do must not cost thousands of 100 KB-shared-memory blocks)
  for (int idx = blockIdx.x * warps_per_block + wib;; idx += gridDim.x * warps_per_block) {
    int slot, b;
    double* ws;
    if (!locate(A, idx, slot, b, ws, false)) return;
    Solver<32> S(P, slot_base(P, ws, slot), slot & 31);
    load_coeffs(A, b, S.cf);
    CoopSolver<32, DevExec> C(S, st, pub, DevExec{lane});
    if (fresh) {
      double s0[6];
      load_state6(A, b, s0);
      if (A.warm && A.step > 0) {
        if (lane == 0) S.init_warm(s0, A.mu_warm);
        __syncwarp();
      } else C.init(s0);
    } else {
      if (S.load_phase() == PH_DONE) continue;   // whole warp
      if (lane == 0) S.load_state();
      __syncwarp();
    }
    C.run();
    if (lane == 0) {
      S.store_state();
      write_result(P, A, b, S);
    }
    __syncwarp();
  }
}

// per-device, once (b200mpc_create): the cooperative kernel uses up to 200 KB of dynamic shared memory per block
cudaError_t solver_prepare_device() {
  return cudaFuncSetAttribute(mpc_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

static size_t coop_doubles_per_warp(int N) { return (size_t)N * kCoopStageDoubles + (sizeof(CoopPub) + 7) / 8; }
// warps of the cooperative kernel that are resident on the device at once (0: the horizon does not fit): 128-thread blocks
// of up to 4 problems, 2 blocks per SM by registers, per-problem scratch in shared memory
static int coop_resident_warps(int N) {
  const size_t per_warp = coop_doubles_per_warp(N) * sizeof(double), limit = 200 * 1024;
  if (per_warp > limit) return 0;
  int wpb = (int)(limit / per_warp);
  if (wpb > 4) wpb = 4;
  int blocks = (int)((size_t)227 * 1024 / (per_warp * wpb));
  if (blocks > 2) blocks = 2;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * blocks * wpb;
}

// launches the cooperative kernel if the horizon fits in shared memory; returns false otherwise
static bool launch_coop(const Params& P, const SolveArgs& A, int fresh, cudaStream_t stream, cudaError_t* err, int take_below = 0) {
  const size_t per_warp = coop_doubles_per_warp(P.N) * sizeof(double);
  const size_t limit = 200 * 1024;
  if (per_warp > limit) return false;
  int wpb = (int)(limit / per_warp);
  if (wpb > 4) wpb = 4;
  const size_t smem = per_warp * wpb;
  const int n = take_below > 0 && take_below < A.b1 - A.b0 ? take_below : A.b1 - A.b0;
  int grid = (n + wpb - 1) / wpb;
  const int cap = 4 * coop_resident_warps(P.N) / wpb;   // four waves; the warps stride over the rest
  if (cap > 0 && grid > cap) grid = cap;
  mpc_coop_kernel<<<grid, 32 * wpb, smem, stream>>>(P, A, fresh, wpb, (int)coop_doubles_per_warp(P.N), take_below);   // no idle warps: their registers would stay allocated
  *err = cudaGetLastError();
  return true;
}

// launches every kernel of the solve of problems [b0, b1) on one stream
static cudaError_t launch_part(const Params& P, SolveArgs A, const SolveConfig& cfg, int rounds, int take_below, cudaStream_t stream, long long* n) {
  const int nb = A.b1 - A.b0;
  const int grid = (nb + kBlock - 1) / kBlock;
  for (int step = 0; step < A.steps; ++step) {
    A.step = step;
    cudaError_t ce = cudaSuccess;
    if (cfg.mode == kModeFused || A.B < cfg.fused_below) {
      A.desc = nullptr;
      if (!(cfg.coop && cfg.mode != kModeFused && launch_coop(P, A, 1, stream, &ce))) mpc_fused_kernel<<<grid, kBlock, 0, stream>>>(P, A, 1);
      if (ce != cudaSuccess) return ce;
      *n += 1;
    } else {
      mpc_init_kernel<<<grid, kBlock, 0, stream>>>(P, A);
      for (int r = 0; r < rounds; ++r) {
        // a compaction attempt follows every round from compact_from on (not the last: only the finisher is left)
        const bool attempt = A.desc && r + 1 >= cfg.compact_from && r + 1 < rounds;
        SolveArgs Ar = A;
        Ar.count_live = attempt ? 1 : 0;
        mpc_factor_kernel<<<grid, kBlock, 0, stream>>>(P, A);
        mpc_forward_kernel<<<grid, kBlock, 0, stream>>>(P, A);
        launch_step(P, Ar, cfg, grid, stream);
        if (attempt) {
          mpc_repack_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(P, A, cfg.compact_max_live, 0, nullptr, nullptr, nullptr);
          *n += 1;
          if (take_below > 0 && r + 1 >= cfg.handover_from && cfg.coop && launch_coop(P, A, 0, stream, &ce, take_below)) {
            if (ce != cudaSuccess) return ce;
            *n += 1;
          }
        }
      }
      if (!(cfg.coop && launch_coop(P, A, 0, stream, &ce))) mpc_fused_kernel<<<grid, kBlock, 0, stream>>>(P, A, 0);
      if (ce != cudaSuccess) return ce;
      *n += 2 + 3LL * rounds;
    }
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return ce;
  }
  return cudaSuccess;
}

// A large batch is cut into `split` contiguous parts that run concurrently on the caller's stream and on auxiliary
// streams (fork / join with events; capturable into a CUDA graph): the kernels of different parts are in different
// sweeps at any moment, which fills the partial waves of the 252-register factor kernel and hides the thin tail of
// one part behind the bulk of another even when the caller uses a single stream.
cudaError_t launch_solve(const Params& P, int B, int steps, const double* state6, const double* coeffs, int ncoef,
                         double* ws, double* out8, double* traj, double* obj, int* status, int* iters,
                         const SolveConfig& cfg, cudaStream_t stream, const SplitStreams* ss, long long* n_launches, int io_aos) {
  if (B <= 0) return cudaSuccess;
  // buffer layout: region 0 | region 1 | map 0 | map 1 | compaction state of every part
  const size_t reg = region_doubles(P.N, B);
  int* ints = reinterpret_cast<int*>(ws + 2 * reg);
  const size_t map_ints = ((size_t)B + 1) / 2 * 2;
  int* desc = ints + 2 * map_ints;
  // a warm-started step reads the previous solution at the problem's own slot: no compaction then
  if (!(cfg.compact_max_live > 0.0) || cfg.warm_start) desc = nullptr;
  SolveArgs A{B, steps, 0, ncoef, io_aos ? 1 : 0, 0, B, cfg.warm_start ? 1 : 0, cfg.warm_mu, state6, coeffs, ws, ws + reg, ints, ints + map_ints, desc, 0,
              out8, traj, obj, status, iters};
  long long n = 0;
  int parts = 1;
  if (ss && cfg.mode == kModePerPass && cfg.split > 1 && B >= cfg.split * cfg.fused_below) parts = cfg.split < ss->n_aux + 1 ? cfg.split : ss->n_aux + 1;
  // automatic setting: up to handover_max_rounds rounds, and after every late round the cooperative kernel takes the
  // rest over as soon as the compacted batch fits its two waves (the launches after that find nothing to do); a fixed
  // number of rounds (b200mpc_set_solver_mode) hands over exactly there
  const bool adaptive = cfg.rounds <= 0 && desc != nullptr && cfg.coop && cfg.handover_below > 0;
  // Longer horizons have a longer tail of iteration counts (N = 100: 3.6 % of the problems need more than 20, 1.8 % more
  // than 100) and a slower cooperative kernel (0.34 ms per problem and iteration at N = 100, 2 warps per SM), so the sweeps
  // keep the tail longer there: measured with 4 overlapped batches of 65 536 problems at N = 100, 20 rounds 267 ms per
  // batch, 120 rounds 225 ms; N <= 50 does not gain (profiles/r1_handover_sweep_long_horizons.log).
  const int max_rounds = cfg.handover_max_rounds > 0 ? cfg.handover_max_rounds : (P.N <= 50 ? 20 : 2 * P.N - 80);
  const int rounds = cfg.rounds > 0 ? cfg.rounds : (adaptive ? max_rounds : (parts > 1 ? 16 : 18));
  const int take_below = adaptive ? (cfg.handover_below / parts > 64 ? cfg.handover_below / parts : 64) : 0;
  cudaError_t e = cudaSuccess;
  if (parts == 1) {
    e = launch_part(P, A, cfg, rounds, take_below, stream, &n);
  } else {
    const int per = (((B + parts - 1) / parts) + 63) / 64 * 64;
    if ((e = cudaEventRecord(ss->fork, stream)) != cudaSuccess) return e;
    for (int p = 1; p < parts; ++p)
      if ((e = cudaStreamWaitEvent(ss->aux[p - 1], ss->fork, 0)) != cudaSuccess) return e;
    for (int p = 0; p < parts && e == cudaSuccess; ++p) {
      SolveArgs Ap = A;
      Ap.b0 = p * per;
      Ap.b1 = (p + 1) * per < B ? (p + 1) * per : B;
      if (desc) Ap.desc = desc + p * kDescInts;
      if (Ap.b0 >= Ap.b1) continue;
      e = launch_part(P, Ap, cfg, rounds, take_below, p == 0 ? stream : ss->aux[p - 1], &n);
    }
    for (int p = 1; p < parts; ++p) {   // always join, also after an error, so a capture can end cleanly
      cudaError_t e2 = cudaEventRecord(ss->join[p - 1], ss->aux[p - 1]);
      if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(stream, ss->join[p - 1], 0);
      if (e == cudaSuccess) e = e2;
    }
  }
  if (n_launches) *n_launches += n;
  return e;
}

// ---------------------------------------------------------------------------------------------
// Pipelined solves on one handle (b200mpc_set_pipeline).  A solve is cut in two launch sequences:
//   bulk  init + rounds in the handle's main workspace; as soon as at most tail.slots problems are still iterating they
//         are moved to the tail context and the main workspace is free for the next batch;
//   tail  rounds (with compaction and the conditional hand-over to the cooperative kernel) + finisher on the tail context.
// The caller orders them with events: bulk(i+1) after bulk(i), tail(i) after bulk(i), bulk(i+depth) after tail(i), so
// the thin tail of one batch -- latency bound: a straggler needs up to hundreds of iterations at long horizons -- runs
// under the bulk of the following batches instead of holding a full-size workspace.  Results do not depend on where a
// problem is finished.
static SolveArgs tail_args(const Params& P, int B, const double* state6, const double* coeffs, int ncoef, double* out8, double* traj,
                           double* obj, int* status, int* iters, int io_aos, const TailCtx& tail) {
  const size_t reg = region_doubles(P.N, tail.slots);
  int* ints = reinterpret_cast<int*>(tail.ws + 2 * reg);
  const size_t map_ints = ((size_t)tail.slots + 1) / 2 * 2;
  return SolveArgs{B, 1, 0, ncoef, io_aos ? 1 : 0, 0, tail.slots, 0, 0.0, state6, coeffs, tail.ws, tail.ws + reg, ints, ints + map_ints,
                   ints + 2 * map_ints, 0, out8, traj, obj, status, iters};
}
static int auto_rounds(const Params& P, const SolveConfig& cfg) {
  return cfg.rounds > 0 ? cfg.rounds : (cfg.handover_max_rounds > 0 ? cfg.handover_max_rounds : (P.N <= 50 ? 20 : 2 * P.N - 80));
}

cudaError_t launch_solve_bulk(const Params& P, int B, const double* state6, const double* coeffs, int ncoef, double* ws, double* out8,
                              double* traj, double* obj, int* status, int* iters, const SolveConfig& cfg, cudaStream_t stream,
                              long long* n_launches, int io_aos, const TailCtx& tail) {
  if (B <= 0) return cudaSuccess;
  const size_t reg = region_doubles(P.N, B);
  int* ints = reinterpret_cast<int*>(ws + 2 * reg);
  const size_t map_ints = ((size_t)B + 1) / 2 * 2;
  SolveArgs A{B, 1, 0, ncoef, io_aos ? 1 : 0, 0, B, 0, 0.0, state6, coeffs, ws, ws + reg, ints, ints + map_ints, ints + 2 * map_ints, 0,
              out8, traj, obj, status, iters};
  const SolveArgs T = tail_args(P, B, state6, coeffs, ncoef, out8, traj, obj, status, iters, io_aos, tail);
  const int grid = (B + kBlock - 1) / kBlock;
  const int rounds = auto_rounds(P, cfg);
  long long n = 0;
  mpc_tail_reset_kernel<<<1, 32, 0, stream>>>(T.desc);
  mpc_init_kernel<<<grid, kBlock, 0, stream>>>(P, A);
  n += 2;
  for (int r = 0; r < rounds; ++r) {
    const bool attempt = r + 1 >= cfg.compact_from;
    SolveArgs Ar = A;
    Ar.count_live = attempt ? 1 : 0;
    mpc_factor_kernel<<<grid, kBlock, 0, stream>>>(P, A);
    mpc_forward_kernel<<<grid, kBlock, 0, stream>>>(P, A);
    launch_step(P, Ar, cfg, grid, stream);
    n += 3;
    if (attempt) {
      mpc_repack_kernel<<<(B + 255) / 256, 256, 0, stream>>>(P, A, cfg.compact_max_live, tail.slots, T.ws1, T.map0, T.desc);
      n += 1;
    }
  }
  // Whatever never fitted the tail context is finished in place, by the thread loop: in the normal case the workspace
  // is empty by now and the launch must be cheap -- a cooperative-kernel launch, even one that finds nothing to do,
  // asks for 100-200 KB of shared memory per block and has to wait for SMs to drain while the tails of the previous
  // batches keep them busy (measured at N = 100 with 32 tails in flight: ~150 ms per bulk).
  mpc_fused_kernel<<<grid, kBlock, 0, stream>>>(P, A, 0);
  n += 1;
  cudaError_t ce = cudaGetLastError();
  if (n_launches) *n_launches += n;
  return ce;
}

cudaError_t launch_solve_tail(const Params& P, int B, const double* state6, const double* coeffs, int ncoef, double* out8, double* traj,
                              double* obj, int* status, int* iters, const SolveConfig& cfg, cudaStream_t stream, long long* n_launches,
                              int io_aos, const TailCtx& tail) {
  if (B <= 0) return cudaSuccess;
  const SolveArgs A = tail_args(P, B, state6, coeffs, ncoef, out8, traj, obj, status, iters, io_aos, tail);
  const int nb = tail.slots, grid = (nb + kBlock - 1) / kBlock;
  // Up to N = 50 the tail is short (the slowest problems need a few dozen iterations): a few rounds, after each of which
  // the cooperative kernel takes the rest over once it fits its resident warps, as in launch_solve.  Above that the tail
  // is long (N = 100: 2 % of the problems need more than 100 iterations, the slowest about 1000) and the tails of many
  // batches are in flight at once: they stay with the compacted thread-per-problem sweeps -- full warps whose blocks come
  // and go, so the tails of all batches share the machine like one large batch -- because the cooperative kernel holds
  // 100 KB of shared memory per problem there (2 resident warps per SM) and a thread loop that runs to completion holds its
  // registers until the slowest problem of the block is done (measured: with either as the finisher of 32 tails in flight
  // the bulk of the following batches runs at a third of its speed).  Only the last few problems go to the cooperative kernel.
  const bool long_tail = P.N > 50;
  const int rounds = cfg.tail_rounds > 0 ? cfg.tail_rounds : (long_tail ? 6 * P.N : 10);
  int take_below = 0;
  if (cfg.coop && cfg.handover_below > 0) {
    const int resident = coop_resident_warps(P.N);
    take_below = long_tail ? (cfg.tail_take_below > 0 ? cfg.tail_take_below : 32) : (cfg.handover_below < resident ? cfg.handover_below : resident);
    if (take_below > resident) take_below = resident;
    if (take_below > nb) take_below = nb;
  }
  const int coop_every = long_tail ? 8 : 1;
  long long n = 0;
  cudaError_t ce = cudaSuccess;
  for (int r = 0; r < rounds; ++r) {
    SolveArgs Ar = A;
    Ar.count_live = 1;
    mpc_factor_kernel<<<grid, kBlock, 0, stream>>>(P, A);
    mpc_forward_kernel<<<grid, kBlock, 0, stream>>>(P, A);
    launch_step(P, Ar, cfg, grid, stream);
    mpc_repack_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(P, A, cfg.compact_max_live, 0, nullptr, nullptr, nullptr);
    n += 4;
    if (take_below > 0 && (r + 1) % coop_every == 0 && launch_coop(P, A, 0, stream, &ce, take_below)) {
      if (ce != cudaSuccess) return ce;
      n += 1;
    }
  }
  if (take_below > 0 && launch_coop(P, A, 0, stream, &ce, take_below)) {
    if (ce != cudaSuccess) return ce;
    n += 1;
  }
  // what is left after the last round: the thread loop (the cooperative kernel took over above if the rest fits it)
  mpc_fused_kernel<<<grid, kBlock, 0, stream>>>(P, A, 0);
  n += 1;
  if (ce == cudaSuccess) ce = cudaGetLastError();
  if (n_launches) *n_launches += n;
  return ce;
}

// ---------------------------------------------------------------------------------------------
// FP64 peak: 8 independent dependent-FMA chains per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) sink[0] = s;   // keeps the chains live, never true in practice
}

// ---------------------------------------------------------------------------------------------
// Arithmetic self-test: the sweeps' reciprocal / quotient (drcp / ddiv, mpc_core.cuh) on caller-supplied operands, so a
// test can hold them against the IEEE results.
__global__ void division_selftest_kernel(int n, const double* a, const double* b, double* quot, double* rcp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  quot[i] = ddiv(a[i], b[i]);
  rcp[i] = drcp(b[i]);
}
cudaError_t launch_division_selftest(int n, const double* a, const double* b, double* quot, double* rcp, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  division_selftest_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, a, b, quot, rcp);
  return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double* sink, int blocks, int threads, int iters, cudaStream_t stream, double* flop) {
  fp64_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters);
  *flop = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
  return cudaGetLastError();
}

}  // namespace b200mpc
