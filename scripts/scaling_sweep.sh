#!/bin/bash
# BASELINE.json configs[4]: contact-heavy sweep over the batch size on ONE GPU (the multi-GPU points are N independent shards).
for n in 1024 4096 16384 65536 262144; do
  python bench.py --workload c5 --envs-per-gpu $n --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stats']
print('c5 envs=%7d  %.3e env-steps/s  e2e %.3e  ms/step %.3f  niter %.2f  contacts %.2f  resets/step %.1f' % ($n, d['value'], d['e2e']['value'], d['ms_per_step'], s.get('solver_niter_mean', s.get('solver_niter_mean_last_step', 0)), s['active_contacts_mean_last_step'], s['resets_per_step']))"
done
python bench.py --workload c5 --impl reference --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c5 reference arm (CPU port, %d cores): %.3e env-steps/s' % (d['cpu_baseline']['cores'], d['value']))"
