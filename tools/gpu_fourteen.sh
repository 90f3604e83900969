#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_base.json 2> gpurun_out/f_base.err
for v in K3B4 K6B2 K8B2 K5B3 K2B5; do
  PSIM_B200_LIB=$PWD/psim_b200/lib/variants/libpsim_b200_$v.so timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_$v.json 2> gpurun_out/f_$v.err
done
for f in gpurun_out/f_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['stats']['warps'], d['stats']['steps_per_launch'])"; done
tail -2 gpurun_out/f_*.err
