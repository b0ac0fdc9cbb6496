#!/usr/bin/env python
"""bench.py -- MPC solves/sec (BASELINE.json metric) for the B200-native batched MPC solver.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--workload roadmap|line]

A "step" is one pass of the hot path over one batch of synthetic input: B independent MPC::Solve problems
(N=25 horizon) solved by one launch of the fused interior-point kernel.  Default workload = BASELINE.json
configs[3]: 65 536 problems with a degree-3 reference fitted (on the GPU, K4) to roadmap.csv segments.
Under torchrun (N>1) every rank drives its own GPU with its own 65 536 problems (weak scaling, no data-path
collective); torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the device time.

value    whole-job solves/s with inputs already resident in HBM (device-pointer C-ABI call, CUDA events)
e2e      the same through the host-buffer C-ABI call the reference's MPC::Solve would bind: pinned host buffers,
         H2D + D2H inside the timed region
"""
import argparse
import json
import multiprocessing as mp_
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpc_solves_per_sec"
UNIT = "solves/s"
HORIZON = 25


def f_iter(N):
    """Algorithmic FLOP per interior-point iteration per problem: (N-1)(F_ric + F_eval) + F_vec N."""
    return (N - 1) * (2421 + 150) + 20 * N


def profiled_traffic_bytes():
    """DRAM bytes (read + write) of all solver kernels of ONE step, from the committed ncu launch list of this same
    command (profiles/r1_compact_launches_time_dram.csv: one init ... next init).  None if the file is missing."""
    import csv
    path = os.path.join(ROOT, "profiles", "r1_compact_launches_time_dram.csv")
    try:
        rows = [r for r in csv.reader(open(path)) if len(r) > 5]
        hdr = rows[0]
        ik, iv, im, iu, iid = (hdr.index(c) for c in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit", "ID"))
        per = {}
        for r in rows[1:]:
            if "dram__bytes" not in r[im]:
                continue
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iu], 1.0)
            d = per.setdefault(int(r[iid]), [r[ik], 0.0])
            d[1] += float(r[iv].replace(",", "")) * scale
        seq = [per[k] for k in sorted(per)]
        starts = [i for i, (k, _) in enumerate(seq) if "mpc_init_kernel" in k]
        if len(starts) < 2:
            return None
        return float(sum(b for k, b in seq[starts[0]:starts[1]] if "mpc_" in k))
    except Exception:
        return None


def make_workload(kind, B, seed_shift=0, mpc=None):
    """Returns (states (B,6), coeffs (B,ncoef)).  Degree-3 coefficients come from the GPU polyfit when a handle is
    given (that is the product path); the CPU-only reference arm uses numpy's QR for its own inputs."""
    from udacitympc_b200 import synth
    rng = synth.MT19937_64(synth.SEED + 1000 * seed_shift)
    if kind == "line":
        return synth.line_problems(B, rng)
    xs, ys = synth.roadmap_windows(B, rng)
    if mpc is not None:
        import udacitympc_b200 as m
        fit = m.polyfit_batch(xs, ys, 3, mpc=mpc)
    else:
        V = np.stack([xs ** i for i in range(4)], axis=2)
        fit = np.stack([np.linalg.lstsq(V[b], ys[b], rcond=None)[0] for b in range(B)])
    st = synth.roadmap_problems(B, fit, synth.MT19937_64(synth.SEED + 1 + 1000 * seed_shift))
    return st, np.ascontiguousarray(fit)


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the reference's own Ipopt 3.12.7 + MUMPS binaries (oracle/_ref), one process per host core.
def _cpu_worker(args):
    kind, states, coeffs = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bindings as ob
    t0 = time.perf_counter()
    its = 0
    for b in range(len(states)):
        r = ob.ref_solve(states[b], coeffs[b]) if kind == "reference" else ob.port_solve(states[b], coeffs[b])
        its += r["iters"]
    return time.perf_counter() - t0, its


def cpu_reference_rate(states, coeffs, per_core, cores=None):
    """Times the reference CPU implementation on a bounded sample: `per_core` solves on each of `cores` processes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bindings as ob
    kind = "reference" if ob.ref_available() else "port"
    cores = cores or os.cpu_count() or 1
    n = min(len(states), cores * per_core)
    per = max(1, n // cores)
    chunks = [(kind, states[i * per:(i + 1) * per], coeffs[i * per:(i + 1) * per]) for i in range(cores)]
    chunks = [c for c in chunks if len(c[1])]
    ctx = mp_.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(chunks)) as pool:
        res = pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    solved = sum(len(c[1]) for c in chunks)
    busy = max(r[0] for r in res)
    iters = sum(r[1] for r in res)
    return dict(value=solved / busy, unit=UNIT, cores=len(chunks), kind=kind,
                sample=f"first {solved} problems of the workload, {per} per process, one process per host core "
                       f"(Ipopt 3.12.7 + MUMPS 4.10.0 reference binaries, hand-derived derivatives standing in for CppAD)"
                if kind == "reference" else f"first {solved} problems, C port of the reference algorithm",
                wall_s=wall, mean_iters=iters / max(1, solved))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    states, coeffs = make_workload(args.workload, min(args.batch, 4096))
    cores = os.cpu_count() or 1
    per_core = 100   # solves per host core per step: ~0.6 s of CPU work per step, pool start-up amortised
    times = []
    rate = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(states, coeffs, per_core, cores)
        if i >= args.warmup:
            times.append(per_core * r["cores"] / r["value"])
            rate = r
    solved_per_step = per_core * rate["cores"]
    ms = 1e3 * float(np.mean(times))
    value = solved_per_step / (ms * 1e-3)
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=config_dict(args, args.batch),
                cpu_baseline=dict(value=value, unit=UNIT, cores=rate["cores"], kind=rate["kind"], sample=rate["sample"]),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


def config_dict(args, batch):
    return dict(workload=("BASELINE configs[3]: batched MPC, degree-3 reference fitted to roadmap.csv segments"
                          if args.workload == "roadmap" else "BASELINE configs[2]: batched mpc_to_line, degree-1 reference y=-1"),
                horizon_N=HORIZON, dt=0.05, batch_per_gpu=batch, problems_per_step=batch * args.gpus,
                sharding="contiguous index ranges per GPU, no collective",
                l2="per-step working set (solver workspace region, 17.9 KB/problem = 1.18 GB at 65 536) >> 126 MB L2; "
                   "inputs cycle over 4 distinct batches, the same on every rank")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(max(mx)) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import udacitympc_b200 as mpcmod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, K, W = args.batch, args.steps, args.warmup

    S = max(1, args.streams)
    # concurrency comes either from the caller (S overlapped handles, no internal split) or from the library's
    # internal batch split (one handle)
    split = args.split if args.split > 0 else (1 if S > 1 else 4)
    mpcs = [mpcmod.MPC(device=local) for _ in range(S)]
    for m in mpcs:
        m.set_solver_mode({"perpass": 0, "fused": 1}[args.mode], args.rounds, -1)
        m.set_batch_split(split)
    mpc = mpcs[0]
    # Four distinct 65 536-problem input sets, cycled step by step.  Every rank solves the SAME four sets (weak
    # scaling with identical per-GPU work by construction); about half of such sets contain a 30-50 iteration
    # straggler, so cycling several makes the single-GPU figure representative of the workload.
    nsets = args.input_sets
    sets = [make_workload(args.workload, B, seed_shift=s, mpc=mpc) for s in range(nsets)]
    ncoef = sets[0][1].shape[1]
    # device-resident, field-major inputs and outputs (one output set per stream)
    d_in = [(torch.from_numpy(np.ascontiguousarray(st.T)).to(dev), torch.from_numpy(np.ascontiguousarray(cf.T)).to(dev))
            for st, cf in sets]
    d_out = [dict(out8=torch.empty((8, B), dtype=torch.float64, device=dev), obj=torch.empty(B, dtype=torch.float64, device=dev),
                  status=torch.empty(B, dtype=torch.int32, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
             for _ in range(S)]
    # real (non-default) streams: kernels and timing events share them.  With S > 1 consecutive steps alternate
    # between S solver handles / streams so that the thin tail of one batch (few problems still iterating)
    # overlaps the bulk of the next.
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    torch.cuda.synchronize()

    def dev_step(i):
        st, cf = d_in[i % nsets]
        o = d_out[i % S]
        mpcs[i % S].solve_batch_device(B, st.data_ptr(), cf.data_ptr(), ncoef, o["out8"].data_ptr(), 0, o["obj"].data_ptr(),
                                       o["status"].data_ptr(), o["iters"].data_ptr(), streams[i % S].cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = mpc.fp64_peak_tflops()
    # warm-up: at least W steps, and every (solver handle, input set) pair once -- each pair is its own CUDA graph,
    # captured and instantiated at its first use
    import math
    for i in range(max(W, S * nsets // math.gcd(S, nsets))):
        dev_step(i)
    barrier()
    for m in mpcs:
        m.kernel_time_ms(reset=True)
    launches0 = sum(m.launch_count() for m in mpcs)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s_ in streams[1:]:
        s_.wait_event(e0)
    for i in range(K):
        dev_step(i)
    for s_ in streams[1:]:
        ev = torch.cuda.Event()
        ev.record(s_)
        streams[0].wait_event(ev)
    e1.record(streams[0])
    barrier()
    ms_total = e0.elapsed_time(e1)
    kern = [m.kernel_time_ms(reset=True) for m in mpcs]
    kern_ms, kern_n = sum(k[0] for k in kern), sum(k[1] for k in kern)
    launches = sum(m.launch_count() for m in mpcs) - launches0
    # mean iterations of the two input sets (for the algorithmic FLOP count)
    iters_mean = []
    ok_frac = []
    for s in range(nsets):
        st, cf = d_in[s]
        o = d_out[0]
        mpc.solve_batch_device(B, st.data_ptr(), cf.data_ptr(), ncoef, o["out8"].data_ptr(), 0, o["obj"].data_ptr(),
                               o["status"].data_ptr(), o["iters"].data_ptr(), streams[0].cuda_stream)
        torch.cuda.synchronize()
        iters_mean.append(float(o["iters"].double().mean().item()))
        ok_frac.append(float((o["status"] == 0).double().mean().item()))
    mpc.kernel_time_ms(reset=True)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    per_rank_ms = [ms_total / K]
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x.item()) / K for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * K / (ms_total * 1e-3)

    # ---- e2e: host buffers (pinned), H2D + D2H inside the timed region, through the host-buffer C-ABI call;
    # S host threads, one per solver handle (the call blocks until its results are in the caller's buffers)
    import ctypes
    from concurrent.futures import ThreadPoolExecutor
    pin = [(torch.from_numpy(st).pin_memory(), torch.from_numpy(cf).pin_memory()) for st, cf in sets]
    hout = [dict(out8=torch.empty((B, 8), dtype=torch.float64).pin_memory(), obj=torch.empty(B, dtype=torch.float64).pin_memory(),
                 status=torch.empty(B, dtype=torch.int32).pin_memory(), iters=torch.empty(B, dtype=torch.int32).pin_memory())
            for _ in range(S)]
    lib = mpcmod.load_library()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)

    def host_step(i, k=None):
        k = i % S if k is None else k
        st, cf = pin[i % nsets]
        o = hout[k]
        rc = lib.b200mpc_solve_batch(mpcs[k].handle, B, ctypes.cast(st.data_ptr(), dp), ctypes.cast(cf.data_ptr(), dp), ncoef,
                                     ctypes.cast(o["out8"].data_ptr(), dp), None, ctypes.cast(o["obj"].data_ptr(), dp),
                                     ctypes.cast(o["status"].data_ptr(), ip), ctypes.cast(o["iters"].data_ptr(), ip))
        if rc:
            raise RuntimeError(lib.b200mpc_last_error().decode())

    # host threads of the end-to-end leg (one solver handle each): up to 6, but not more than the rank's share of the
    # host cores (never fewer than 3)
    T = args.e2e_threads if args.e2e_threads > 0 else min(6, max(3, (os.cpu_count() or 6) // world))
    T = max(1, min(S, T))

    def worker(k, n):
        for i in range(k, n, T):
            host_step(i, k)

    for i in range(max(3, W, T)):   # every handle of the leg allocates its buffers and captures its graph here
        host_step(i, i % T)
    barrier()
    with ThreadPoolExecutor(T) as ex:
        t0 = time.perf_counter()
        list(ex.map(lambda k: worker(k, K), range(T)))
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    # unloaded latency of one host-buffer call at a time, on a handle with the library's default settings (a lone
    # call is cut into concurrent sub-batches by the library itself)
    lat = []
    with mpcmod.MPC(device=local) as lone:
        lone.set_solver_mode({"perpass": 0, "fused": 1}[args.mode], args.rounds, -1)
        mpcs.append(lone)
        hout.append(hout[0])
        for i in range(3):
            host_step(i, len(mpcs) - 1)
        for i in range(max(K, args.latency_reps)):
            a = time.perf_counter()
            host_step(i, len(mpcs) - 1)
            lat.append(time.perf_counter() - a)
        mpcs.pop()
        hout.pop()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    p99 = torch.tensor([float(np.percentile(lat, 99))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(p99, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / float(t.item())
    for m in mpcs:
        m.kernel_time_ms(reset=True)
    clocks = sampler.stop()
    per_rank_mhz = [clocks.get("sm_mhz") or 0.0]
    if world > 1:
        c = torch.tensor([clocks.get("sm_mhz") or 0.0], dtype=torch.float64, device=dev)
        allc = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(allc, c)
        per_rank_mhz = [float(x.item()) for x in allc]

    if rank == 0:
        mean_it = float(np.mean(iters_mean))
        flop_per_launch = f_iter(HORIZON) * mean_it * B
        # device time of the solver kernels per step: with one stream, the first-to-last-kernel interval of each
        # step (CUDA events inside the C ABI); with overlapped streams the intervals overlap, so the per-step share of
        # the timed region is used instead
        avg_kernel_ms = kern_ms / max(1, kern_n) if S == 1 else ms_total / K
        achieved = flop_per_launch / (avg_kernel_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / K,
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
            config=config_dict(args, B),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=B * (6 + ncoef) * 8, d2h_bytes_per_step=B * (8 + 1) * 8 + 2 * B * 4,
                     p99_batch_latency_ms=1e3 * float(p99.item()), latency_reps=len(lat)),
            gpu_launches=int(launches),
            clocks=clocks, per_rank_ms_per_step=per_rank_ms, per_rank_sm_mhz=per_rank_mhz,
            roofline=dict(bound="fp64", achieved=achieved, peak=fp64_peak, unit="TFLOP/s", frac=achieved / fp64_peak if fp64_peak else None,
                          traffic=profiled_traffic_bytes(), traffic_unit="bytes of DRAM read+write per step (ncu, profiles/r1_compact_launches_time_dram.csv)",
                          hbm_achieved_gbs=(profiled_traffic_bytes() or 0.0) / (avg_kernel_ms * 1e-3) / 1e9 if profiled_traffic_bytes() else None,
                          hbm_frac=((profiled_traffic_bytes() or 0.0) / (avg_kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]
                                    if profiled_traffic_bytes() and peaks.get("hbm_gbs") else None),
                          kernel=("mpc_{init,factor,forward,step,coop}_kernel: all solver kernels of one step (one CUDA graph), first to last"
                                  if args.mode == "perpass" else "mpc_fused_kernel"),
                          avg_kernel_ms=avg_kernel_ms, launches_timed=kern_n, solver_mode=args.mode, streams=S, batch_split=split, e2e_host_threads=T,
                          flop_per_launch=flop_per_launch, mean_ip_iters=mean_it,
                          peak_source="DFMA microbenchmark in this run (b200mpc_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                          hbm_peak_gbs=peaks.get("hbm_gbs"), algorithmic_io_bytes_per_solve=(6 + ncoef + 8 + 2) * 8,
                          note="SURVEY 8d classifies the solver as FP64 bound (intensity >> ridge, no tensor cores): achieved = F_iter(25)=62204 FLOP x mean "
                               "iterations x B / step device time over the measured DFMA peak.  The kernels exploit the sparsity of the dynamics "
                               "(~4x fewer executed FLOP) and in practice sit between the issue and the HBM roof: hbm_frac = measured DRAM "
                               "traffic of one step (ncu) / step time / measured copy bandwidth"),
            solved_fraction=float(np.mean(ok_frac)),
        )
        if world == 1 and not args.no_cpu_baseline:
            st, cf = sets[0]
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_rate(st[:4096], cf[:4096], args.cpu_per_core).items()}
        emit(line)
    for m in mpcs:
        m.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """Rank 0 must print exactly ONE line on stdout, but libraries write there too (NCCL's version banner comes from C
    code).  Everything written to fd 1 from here on goes to stderr; emit() writes the JSON line to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU per step")
    ap.add_argument("--workload", default="roadmap", choices=["roadmap", "line"])
    ap.add_argument("--latency-reps", type=int, default=100)
    ap.add_argument("--cpu-per-core", type=int, default=150, help="cpu_baseline: solves per host core in the sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--input-sets", type=int, default=4, help="distinct synthetic input batches cycled over the steps")
    ap.add_argument("--e2e-threads", type=int, default=0, help="host threads (one solver handle each) of the end-to-end leg (0 = auto)")
    ap.add_argument("--split", type=int, default=0, help="internal batch split of one solve call (0 = 1 with several streams, 4 with one)")
    ap.add_argument("--streams", type=int, default=6, help="solver handles / CUDA streams consecutive steps alternate between")
    ap.add_argument("--mode", default="perpass", choices=["perpass", "fused"], help="solver execution mode (include/b200mpc.h)")
    ap.add_argument("--rounds", type=int, default=0, help="per-pass mode: rounds before the fused finisher (0 = library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
