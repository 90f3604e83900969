#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 > gpurun_out/pytest_gpu5.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu5.log
tail -5 gpurun_out/pytest_gpu5.log
for bps in 2 3 4; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --blocks-per-sm $bps > gpurun_out/y_bps$bps.json 2> gpurun_out/y_bps$bps.err
done
for spl in 2 4 8 16; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --blocks-per-sm 3 --steps-per-launch $spl > gpurun_out/y_spl$spl.json 2> gpurun_out/y_spl$spl.err
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --kernel 1 > gpurun_out/y_lock.json 2> gpurun_out/y_lock.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 600 -c 1 -o gpurun_out/prof_r01d \
   python bench.py --phonons 100000000 --steps 1 --warmup 0 --no-cpu-baseline --blocks-per-sm 3 > gpurun_out/ncu_full_d.log 2>&1
for f in gpurun_out/y_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9)"; done
