#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 -k "variants" > gpurun_out/pytest_gpu12.log 2>&1; tail -2 gpurun_out/pytest_gpu12.log
timeout 1200 python tools/model_walltimes.py > gpurun_out/models_r01.jsonl 2> gpurun_out/models_r01.err
cat gpurun_out/models_r01.jsonl; tail -3 gpurun_out/models_r01.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
cp psim_b200/bin/psim /tmp/psim_cli; python - <<'PY'
import json, sys
sys.path.insert(0, '.')
from psim_b200 import configs
configs.save(configs.linear(num_phonons=2000000).to_dict(), 'gpurun_out/cli/linear_demo.json')
PY
PSIM_SEED=1 ./psim_b200/bin/psim gpurun_out/cli/linear_demo.json > gpurun_out/cli/stdout.txt 2>&1; cat gpurun_out/cli/stdout.txt; head -3 gpurun_out/cli/ss_linear_demo.txt
