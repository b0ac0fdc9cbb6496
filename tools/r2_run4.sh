#!/bin/bash
# round 2, GPU call 4 (2 GPUs): the multi-device tests, bench.py under torchrun at N=2 (strong scaling + one-process multi leg + weak leg), bench_io.py
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q -k "multi or shard" > gpurun_out/r2_tests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests4.log
tail -4 gpurun_out/r2_tests4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2_scale_2gpu.json 2> gpurun_out/r2_scale_2gpu.err; echo "bench2 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_scale_2gpu.json'))
print('N=2', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3),'multi',d.get('one_process_multi_gpu'),'weak',d.get('weak_scaling'))
P
python bench_io.py > gpurun_out/r2_io.json 2> gpurun_out/r2_io.err; echo "bench_io rc=$?"; tail -3 gpurun_out/r2_io.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_io.json'))
print('io', round(d['value']/1e6,1),'M units/s e2e',round(d['e2e']['value']/1e6,2), d['roofline']['kernels'], d['parity'], d['cpu_baseline']['value'])
P
