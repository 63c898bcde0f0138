#!/bin/bash
# end-to-end loop: how many environment blocks (streams) hide the read-back best
for sh in 1 2 4 8 16; do for extra in "" "--obs"; do
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-shards $sh $extra 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('shards=$sh [$extra]', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'd2h', d['e2e']['d2h_bytes_per_step'])"
done; done
