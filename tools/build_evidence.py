#!/usr/bin/env python
"""Build evidence for profiles/: per-kernel ptxas resource table (registers, spills, stack, shared memory) from the
build log of udacitympc_b200/lib/libb200mpc.so, and a SASS opcode histogram per kernel from cuobjdump -sass.
    python tools/build_evidence.py > profiles/r2_build_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from udacitympc_b200 import build as b
b.build(force=True)
log = open(os.path.join(b.LIBDIR, "build.log")).read()
print("# ptxas -v (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo), udacitympc_b200/lib/build.log")
print(f"{'kernel':58s} {'regs':>5s} {'stack':>6s} {'spill st':>9s} {'spill ld':>9s} {'smem':>7s}")
cur = None; props = {}
for line in log.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m: cur = m.group(1); props[cur] = dict(regs=0, stack=0, sst=0, sld=0, smem=0); continue
    if cur is None: continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and "Function properties" not in line and props[cur]["stack"] == 0 and props[cur]["regs"] == 0:
        props[cur].update(stack=int(m.group(1)), sst=int(m.group(2)), sld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m and props[cur]["regs"] == 0:
        props[cur]["regs"] = int(m.group(1))
        m2 = re.search(r"(\d+) bytes smem", line)
        if m2: props[cur]["smem"] = int(m2.group(1))
def demangle(n):
    try: return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception: return n
for k, p in props.items():
    print(f"{demangle(k)[:58]:58s} {p['regs']:5d} {p['stack']:6d} {p['sst']:9d} {p['sld']:9d} {p['smem']:7d}")
print()
print("# SASS opcode histogram per kernel (cuobjdump -sass libb200mpc.so; static instruction counts)")
sass = subprocess.run(["cuobjdump", "-sass", b.LIB], capture_output=True, text=True).stdout
kern = None; hist = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: kern = demangle(m.group(1)); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern: hist[kern][m.group(1)] += 1
groups = [("FP64 (DFMA DMUL DADD DSETP MUFU.RCP64H..)", lambda o: o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") or o.startswith("MUFU")),
          ("global memory (LDG STG)", lambda o: o in ("LDG", "STG", "LD", "ST", "CCTL", "LDGSTS")),
          ("local memory = spills (LDL STL)", lambda o: o in ("LDL", "STL")),
          ("shared memory (LDS STS)", lambda o: o in ("LDS", "STS", "LDSM")),
          ("warp shuffles / votes (SHFL VOTE MATCH)", lambda o: o in ("SHFL", "VOTE", "MATCH", "REDUX")),
          ("branches / calls (BRA CALL RET BSSY BSYNC..)", lambda o: o in ("BRA", "CALL", "RET", "BSSY", "BSYNC", "EXIT", "BRX", "JMP", "WARPSYNC")),
          ("TMA / tensor (UTMALDG UBLKCP UTC*MMA LDTM)", lambda o: o.startswith("UTMA") or o.startswith("UBLKCP") or o.startswith("UTC") or o in ("LDTM", "STTM", "HMMA"))]
for k, h in hist.items():
    if not k.startswith("b200mpc::mpc_") and "polyfit_kernel<6, 4>" not in k and "rollout" not in k and "soa" not in k and "roadmap" not in k: continue
    tot = sum(h.values())
    print(f"{k}: {tot} instructions")
    for name, f in groups:
        n = sum(v for o, v in h.items() if f(o))
        print(f"    {name:48s} {n:6d}  {100.0 * n / max(1, tot):5.1f} %")
    print("    top opcodes: " + ", ".join(f"{o} {v}" for o, v in h.most_common(12)))
