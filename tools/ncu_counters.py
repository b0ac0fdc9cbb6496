#!/usr/bin/env python
"""Condenses an ncu launch list (the --metrics ... --csv pass of tools/r2_run2.sh over `bench.py --streams 1 --split 1`)
into the per-step counters bench.py reports next to its live timings (profiles/r2_counters.json):

    python tools/ncu_counters.py gpurun_out/r2_launches_counters.csv profiles/r2_counters.json

One step = the launches from one mpc_init_kernel up to the next (one 65 536-problem solve).  Counted per step: DRAM
bytes (read + write), executed FP64 FLOP (2 x DFMA + DADD + DMUL thread instructions, predicated-on), warp instructions,
summed kernel time (cold-cache, serialised: its per-kernel SHARE is what is comparable with a live run, not its sum)."""
import collections
import csv
import json
import sys


def main(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iu, iv, iid = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(int(r[iid]), dict(kernel=r[ik].split("(")[0]))
        d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    seq = [per[k] for k in sorted(per)]
    starts = [i for i, d in enumerate(seq) if "mpc_init_kernel" in d["kernel"]]
    if len(starts) < 2:
        raise SystemExit("need at least two mpc_init_kernel launches (one complete step)")
    step = seq[starts[0]:starts[1]]
    kernels = collections.OrderedDict()
    tot = collections.Counter()
    for d in step:
        k = kernels.setdefault(d["kernel"], collections.Counter())
        k["launches"] += 1
        flop = 2.0 * d.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0) + \
            d.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0) + d.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
        vals = dict(ms=d.get("gpu__time_duration.sum", 0.0), dram_bytes=d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0),
                    fp64_flop=flop, warp_instructions=d.get("smsp__inst_executed.sum", 0.0))
        for n, v in vals.items():
            k[n] += v
            tot[n] += v
    out = dict(source=src, launches_per_step=len(step),
               dram_bytes_per_step=tot["dram_bytes"], executed_flop_per_step=tot["fp64_flop"], warp_instructions_per_step=tot["warp_instructions"],
               summed_kernel_ms_per_step=tot["ms"],
               traffic_unit="bytes of DRAM read+write of all solver kernels of one 65 536-problem step (ncu launch list, " + src.split("/")[-1] + ")",
               executed_flop_source="2 x DFMA + DADD + DMUL thread instructions (smsp__sass_thread_inst_executed_op_d*_pred_on.sum) of all solver kernels "
                                    "of one 65 536-problem step, ncu launch list " + src.split("/")[-1],
               per_kernel={k: dict(launches=int(v["launches"]), ms=v["ms"], share_of_time=v["ms"] / tot["ms"], dram_bytes=v["dram_bytes"],
                                   fp64_flop=v["fp64_flop"], warp_instructions=v["warp_instructions"]) for k, v in kernels.items()})
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "per_kernel"}, indent=1))
    for k, v in out["per_kernel"].items():
        print(f"{k:24s} x{v['launches']:3d}  {v['ms']:7.3f} ms ({100 * v['share_of_time']:4.1f} %)  {v['dram_bytes'] / 1e9:6.2f} GB  {v['fp64_flop'] / 1e9:6.2f} GFLOP")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
