#!/bin/bash
# 8-GPU record: bench.py at 8 x 4096 and 8 x 8192 (= 65 536) envs, S = 4 and S = 20, and the config-4 stand-in
mkdir -p gpurun_out/final
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29601 bench.py --gpus 8 --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/final/bench_c2_n8.json
$T --master-port 29602 bench.py --gpus 8 --steps 100 --warmup 10 --envs-per-gpu 8192 2>/dev/null | tail -1 > gpurun_out/final/bench_65536envs_n8.json
$T --master-port 29603 bench.py --gpus 8 --steps 50 --warmup 10 --envs-per-gpu 8192 --substeps 20 --workload c3 2>/dev/null | tail -1 > gpurun_out/final/bench_65536envs_c3_s20_n8.json
$T --master-port 29604 bench_c4.py --gpus 8 --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final/c4_n8.json
for f in bench_c2_n8 bench_65536envs_n8 bench_65536envs_c3_s20_n8 c4_n8; do python -c "
import json; d=json.load(open('gpurun_out/final/$f.json')); print('$f', '%.3e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), d.get('split_ms_per_iteration'))"; done
