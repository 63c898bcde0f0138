for f in ksim-gym-zbot_b200/csrc/variants/*.so; do for n in 4096 8192; do
  ZBPHYS_LIB=$PWD/$f python bench.py --steps 60 --warmup 10 --no-cpu-baseline --envs-per-gpu $n 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$f', $n, '%.3e'%d['value'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'])"
done; done
