#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/h_base.json 2> gpurun_out/h_base.err
python -c "
import json
d=json.load(open('gpurun_out/h_base.json')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 0 -c 1 -o gpurun_out/prof_r01_kinked \
   python tools/profile_model.py kinked > gpurun_out/ncu_kinked.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 30 -c 1 -o gpurun_out/prof_r01_sides_per \
   python tools/profile_model.py sides_per > gpurun_out/ncu_sides_per.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 0 -c 1 -o gpurun_out/prof_r01_sige_long \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_sige_long.log 2>&1
tail -2 gpurun_out/ncu_kinked.log gpurun_out/ncu_sides_per.log
