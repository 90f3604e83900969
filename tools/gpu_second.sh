#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 -k "kinked or shard or launch or tally_paths" > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 600 -c 2 -o gpurun_out/prof_r01a \
   python bench.py --phonons 10000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
for bps in 2 3 4; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --blocks-per-sm $bps > gpurun_out/v_bps$bps.json 2> gpurun_out/v_bps$bps.err
done
for spl in 2 8 16; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --steps-per-launch $spl > gpurun_out/v_spl$spl.json 2> gpurun_out/v_spl$spl.err
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --tally-aggregate 1 > gpurun_out/v_agg1.json 2> gpurun_out/v_agg1.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --tally-shared 0 > gpurun_out/v_glob.json 2> gpurun_out/v_glob.err
tail -3 gpurun_out/pytest_gpu2.log
for f in gpurun_out/v_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9)"; done
