for f in ksim-gym-zbot_b200/csrc/variants/*.so; do
  ZBPHYS_LIB=$PWD/$f python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$f', '%.3e'%d['value'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'], 'e2e %.3e'%d['e2e']['value'])"
done
