#!/bin/bash
# One GPU-box visit: GPU parity tests, the default bench line, then (only after both exited 0) the ncu launch list and one
# `--set full` capture of the step kernel.  usage: bash scripts/gpu_check.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; rc=$?; tail -3 gpurun_out/pytest_gpu_$tag.log
[ $rc -ne 0 ] && exit $rc
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -5 gpurun_out/bench_$tag.err; exit 1; }
python - <<PY
import json; d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("bench: %.3e env-steps/s  ms/step %.3f  kernel_ms %.3f  e2e %.3e  frac %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["e2e"]["value"], d["roofline"]["frac"]))
PY
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'zb_kernel<\(int\)0' -s 14 -c 1 -f -o gpurun_out/prof_$tag \
  python bench.py --steps 5 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1 || tail -5 gpurun_out/ncu_$tag.log
ls -la gpurun_out/prof_$tag.ncu-rep
