#!/bin/bash
# tuning variants of the library (ksim-gym-zbot_b200/csrc/variants/*.so, built with -D switches) against the default build
for f in ksim-gym-zbot_b200/csrc/libzbphys.so ksim-gym-zbot_b200/csrc/variants/*.so; do for n in 4096 16384; do
  ZBPHYS_LIB=$PWD/$f python bench.py --steps 100 --warmup 10 --no-cpu-baseline --envs-per-gpu $n --e2e-shards 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $f)', $n, '%.3e'%d['value'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'], 'e2e %.3e'%d['e2e']['value'])"
done; done
