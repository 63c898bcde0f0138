#!/bin/bash
# Per-launch fixed cost of the step kernel: kernel time over the number of fused substeps at a fixed dt (t(S) = a + b S).
LIB=ksim-gym-zbot_b200/csrc/libzbphys.so
for S in 1 2 4 8 16; do
  dt=$(python -c "print(0.005*$S)")
  python scripts/ab.py base=$LIB --envs ${ENVS:-4096} --substeps $S --ctrl-dt $dt --steps 40 --warmup 10 --reps 2 $EXTRA 2>&1 | grep "kernel ms"
done
