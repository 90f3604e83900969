#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu3.log
tail -15 gpurun_out/pytest_gpu3.log
for k in 0 1; do for bps in 2 3 4; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --kernel $k --blocks-per-sm $bps > gpurun_out/w_k${k}_bps$bps.json 2> gpurun_out/w_k${k}_bps$bps.err
done; done
for spl in 2 4 8; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --blocks-per-sm 4 --steps-per-launch $spl > gpurun_out/w_spl$spl.json 2> gpurun_out/w_spl$spl.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 600 -c 2 -o gpurun_out/prof_r01b \
   python bench.py --phonons 10000000 --steps 1 --warmup 0 --no-cpu-baseline --blocks-per-sm 4 > gpurun_out/ncu_full_b.log 2>&1
for f in gpurun_out/w_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9)"; done
