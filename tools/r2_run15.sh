#!/bin/bash
# round 2, GPU call 15 (2 GPUs): async host-buffer API, multi entry with persistent worker threads, bench at N=2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "multi or shard or asynchronous or pipelined" > gpurun_out/r2_tests15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests15.log
tail -4 gpurun_out/r2_tests15.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 > gpurun_out/r2_scale_2gpu_b.json 2> gpurun_out/r2_scale_2gpu_b.err; echo "bench2 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_scale_2gpu_b.json'))
print('N=2', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), d['e2e'].get('entry'), d['e2e']['host_threads_per_gpu'], 'multi',d.get('one_process_multi_gpu'))
P
# the 8-GPU shard size on one GPU with few host threads: does the asynchronous e2e leg keep up with the device leg?
python bench.py --batch 8192 --e2e-threads 4 --no-cpu-baseline --latency-reps 20 > gpurun_out/r2_b8192_t4.json 2>> gpurun_out/r2_scale_2gpu_b.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b8192_t4.json'))
print('B=8192 T=4', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), d['e2e'].get('entry'), 'streams', d['timed']['streams'])
P
tail -3 gpurun_out/r2_scale_2gpu_b.err
