#!/bin/bash
# 8-GPU record (round 2): the default bench line carries 8 x 4096, 8 x 8192 (= 65 536), 8 x 16 384 (config 3), S = 20 and the actor-in-the-loop
# rollout; then the reference arm under torchrun (rank 0 only) and the config-4 stand-in.
mkdir -p gpurun_out/final
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29601 bench.py --gpus 8 --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/final/bench_c2_n8.json
$T --master-port 29604 bench_c4.py --gpus 8 --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final/c4_n8.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/final/bench_c2_n8.json"))
print("c2 n8 %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]))
for w in d.get("other_workloads", []):
    print(w["workload"], "%.3e e2e %.3e" % (w["value"], w["e2e"]["value"]))
print("rollout_actor %.3e" % d["rollout_actor"]["value"])
c = json.load(open("gpurun_out/final/c4_n8.json"))
print("c4 n8 %.3e" % c["value"], c.get("split_ms_per_iteration"))
PY
