#!/bin/bash
# DRAM bytes ncu books on the step kernel after a bare memset flush vs. memset + read-back (clean L2), and the bench value of both.
for F in write write+read; do
  python bench.py --flush $F --steps 100 --warmup 20 --no-cpu-baseline --no-extras > gpurun_out/flush_$F.json 2>/dev/null
  python -c "
import json;d=json.loads(open('gpurun_out/flush_$F.json').read().strip().splitlines()[-1]);print('$F','value %.4e kernel_ms %.4f e2e %.4e'%(d['value'],d['roofline']['kernel_ms_per_launch'],d['e2e']['value']))"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --kernel-name-base demangled -k 'regex:zb_kernel<\(int\)0' -s 10 -c 2 --csv python bench.py --flush $F --steps 4 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | grep -E "dram__|duration|lts__" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
