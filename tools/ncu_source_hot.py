#!/usr/bin/env python
"""Per-CUDA-line summary of an `ncu --page source --print-source cuda,sass --csv` export: stall samples, executed
instructions and the dominant stall reasons per source line, per kernel.  usage: ncu_source_hot.py file.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
kern = None; hdr = None; fname = ''
data = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] in ("Function Name", "Kernel Name"):
        kern = r[1].split("(")[0]; data.setdefault(kern, []); hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or kern is None: continue
    if r[0] != "":   # a CUDA source line summary row
        d = dict(zip(hdr[:2] + hdr[4:], r[:2] + r[4:])); d["file"] = fname
        data[kern].append(d)
for k, lines in data.items():
    tot_s = sum(int(l["# Samples"] or 0) for l in lines); tot_i = sum(int(l["Instructions Executed"] or 0) for l in lines)
    print(f"=== {k}: samples {tot_s}, warp instructions {tot_i}")
    stall_cols = [c for c in lines[0] if c.startswith("stall_") and "Not Issued" not in c]
    agg = collections.Counter()
    for l in lines:
        for c in stall_cols: agg[c] += int(l[c] or 0)
    print("   stalls:", ", ".join(f"{c[6:]} {100*v/max(1,tot_s):.0f}%" for c, v in agg.most_common(8)))
    for l in sorted(lines, key=lambda l: -int(l["# Samples"] or 0))[:top]:
        s = int(l["# Samples"] or 0); i = int(l["Instructions Executed"] or 0)
        st = sorted(((int(l[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"   {100*s/max(1,tot_s):5.1f}% smp {100*i/max(1,tot_i):5.1f}% ins  {l['file'][:12]:12s}:{l['Line No']:>5} {l['Source'].strip()[:100]:100s} | " + " ".join(f"{n}:{100*v/max(1,s):.0f}%" for v, n in st))
