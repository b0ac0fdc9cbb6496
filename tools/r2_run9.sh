#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pipe_one.py 100 65536 3
B200MPC_NO_GRAPHS=1 timeout 300 python tools/pipe_one.py 100 65536 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mpc_ --csv --log-file gpurun_out/r2_pipe_one_launches.csv python tools/pipe_one.py 100 65536 1 > gpurun_out/r2_pipe_one_ncu.log 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_pipe_one_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); ig=hdr.index('Grid Size')
seq=[(r[ik].split('(')[0], float(r[iv].replace(',','')), r[ig]) for r in rows[1:]]
print(len(seq),'launches, total ms', sum(t for _,t,_ in seq)/1e6)
agg=collections.defaultdict(lambda:[0,0.0])
for k,t,g in seq:
    agg[(k,g)][0]+=1; agg[(k,g)][1]+=t
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(k, v[0], round(v[1]/1e6,2),'ms', round(v[1]/v[0]/1e3,1),'us avg')
# timeline of tail step kernels
tail=[t for k,t,g in seq if k=='mpc_step_kernel' and g.startswith('(64,')]
print('tail step kernel us at rounds 0,10,50,100,200,300,400,500:', [round(tail[i]/1e3,1) for i in (0,10,50,100,200,300,400,500) if i < len(tail)])
P
