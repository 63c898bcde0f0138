#!/bin/bash
# occupancy variants: warps per CTA (= resident warps per SM) vs. registers per thread
#   build: nvcc ... -DZB_WPB=24 -o ksim-gym-zbot_b200/csrc/variants/libzbphys_wpb24.so (80 regs), =28 (72 regs), =32 (64 regs)
run() { # lib wpb n
  ZBPHYS_LIB=$1 ZBPHYS_WPB=$2 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --envs-per-gpu $3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $1) wpb=$2 n=$3', '%.3e'%d['value'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'], 'e2e %.3e'%d['e2e']['value'])"
}
B=$PWD/ksim-gym-zbot_b200/csrc
for n in 4096 8192 16384; do
  run $B/libzbphys.so 16 $n
  for w in 24 28 32; do run $B/variants/libzbphys_wpb$w.so $w $n; done
done
