#!/bin/bash
# A/B of library variants on the bench workload + one ncu capture of the long-window launch.  usage: tools/gpu_ab.sh [variant.so ...]
mkdir -p gpurun_out
one() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', '%.4g' % d['value'], round(d['ms_per_step'],2), d['e2e']['runs'])"; }
one default
for v in "$@"; do PSIM_B200_LIB=$v one $v; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 0 -c 1 -f -o gpurun_out/prof_long \
   python bench.py --phonons 100000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_long.log 2>&1
ls -la gpurun_out/prof_long.ncu-rep
