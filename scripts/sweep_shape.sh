#!/bin/bash
# CTA shape policy (pick_shape): automatic vs. the 16-warp / 128-register profile forced, over batch sizes
run() { # wpb n
  ZBPHYS_WPB=$1 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --envs-per-gpu $2 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('wpb=$1 n=$2', '%.3e'%d['value'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'], 'e2e %.3e'%d['e2e']['value'])"
}
for n in 2400 3000 4096 5000 6000 8192 10000 12000 16384 65536; do run 0 $n; run 16 $n; done
