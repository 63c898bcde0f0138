python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for extra in "" "--obs" "--workload c3" "--workload c3 --obs"; do
python bench.py --steps 100 --warmup 10 --no-cpu-baseline $extra 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('[$extra]', '%.3e'%d['value'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'], 'e2e %.3e'%d['e2e']['value'], 'd2h', d['e2e']['d2h_bytes_per_step'])"
done
