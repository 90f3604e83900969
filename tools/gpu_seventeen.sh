#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 > gpurun_out/pytest_gpu17.log 2>&1; tail -3 gpurun_out/pytest_gpu17.log
timeout 900 python tools/model_walltimes.py 2>/dev/null > gpurun_out/models_r01b.jsonl; python -c "
import sys, json
for l in open('gpurun_out/models_r01b.jsonl'):
    d=json.loads(l); print(d['model'], d['kernel_ms'], round(d['drift_steps_per_s_kernel']/1e9,2), d['launches'])"
