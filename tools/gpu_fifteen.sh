#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 -k "variants or sige or kinked or shard" > gpurun_out/pytest_gpu15.log 2>&1; tail -3 gpurun_out/pytest_gpu15.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g_base.json 2> gpurun_out/g_base.err
python -c "
import json
d=json.load(open('gpurun_out/g_base.json')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'])"
timeout 900 python tools/model_walltimes.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['model'], d['kernel_ms'], round(d['drift_steps_per_s_kernel']/1e9,2))"
