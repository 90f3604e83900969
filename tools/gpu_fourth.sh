#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 -k "variants or shard or launch or sampling or rates or linear_demo or sige or sides_trans" > gpurun_out/pytest_gpu4.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu4.log
tail -5 gpurun_out/pytest_gpu4.log
for bps in 3 4; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --blocks-per-sm $bps > gpurun_out/x_A_bps$bps.json 2> gpurun_out/x_A_bps$bps.err
done
for v in B C D E F G H; do
  PSIM_B200_LIB=$PWD/psim_b200/lib/variants/libpsim_b200_$v.so timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --blocks-per-sm 4 > gpurun_out/x_$v.json 2> gpurun_out/x_$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 600 -c 1 -o gpurun_out/prof_r01c \
   python bench.py --phonons 100000000 --steps 1 --warmup 0 --no-cpu-baseline --blocks-per-sm 4 > gpurun_out/ncu_full_c.log 2>&1
for f in gpurun_out/x_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9)"; done
