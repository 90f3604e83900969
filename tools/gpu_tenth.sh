#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 > gpurun_out/pytest_gpu10.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu10.log
tail -5 gpurun_out/pytest_gpu10.log
for spl in 1 8 16 24; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --steps-per-launch $spl > gpurun_out/d_spl${spl}.json 2> gpurun_out/d_spl${spl}.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 38 -c 1 -o gpurun_out/prof_r01i \
   python bench.py --phonons 100000000 --steps 1 --warmup 0 --no-cpu-baseline --steps-per-launch 16 > gpurun_out/ncu_full_i.log 2>&1
for f in gpurun_out/d_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9, d['stats']['warps'])"; done
