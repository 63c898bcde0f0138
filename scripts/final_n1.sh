#!/bin/bash
# Round-end single-GPU record: default bench line (with CPU baseline), reference arm, the other configs, launch list, ncu capture.
set -x
mkdir -p gpurun_out/final
python bench.py > gpurun_out/final/bench_c2_n1.json 2> gpurun_out/final/bench_c2_n1.err || exit 1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final/bench_ref_n1.json 2>/dev/null
python bench.py --workload c3 --no-cpu-baseline --no-extras > gpurun_out/final/bench_c3_n1.json 2>/dev/null
python bench.py --substeps 20 --no-cpu-baseline --no-extras > gpurun_out/final/bench_s20_n1.json 2>/dev/null
python bench.py --envs-per-gpu 8192 --no-cpu-baseline --no-extras > gpurun_out/final/bench_8192_n1.json 2>/dev/null
python bench.py --no-obs --no-cpu-baseline --no-extras > gpurun_out/final/bench_c2_noobs_n1.json 2>/dev/null
python bench.py --solver newton --no-cpu-baseline --no-extras > gpurun_out/final/bench_c2_newton_n1.json 2>/dev/null
bash scripts/scaling_sweep.sh > gpurun_out/final/c5_sweep_n1.txt 2>&1
python bench_c4.py --steps 2 --warmup 1 > gpurun_out/final/c4_n1.json 2>/dev/null
# launch list of the default bench command (short run), then one --set full capture of the step kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/final/ncu_list.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'zb_kernel<\(int\)0' -s 14 -c 1 -f -o gpurun_out/final/prof_final \
  python bench.py --steps 5 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/final/ncu_full.log 2>&1
python scripts/actor_bench.py > gpurun_out/final/actor_bench.json 2>/dev/null
python scripts/refresh_traffic.py > gpurun_out/final/traffic.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_driver_cmd_n1.json 2>/dev/null
ls -la gpurun_out/final
