#!/bin/bash
# One GPU call: the GPU test suite, a bench line, the per-model wall-clocks.  usage (under gpurun): tools/gpu_check.sh
out=gpurun_out/check
mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
tail -4 $out/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
cut -c1-200 $out/bench_n1.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reduce-every 48 2>/dev/null | cut -c1-200
timeout 400 python tools/model_walltimes.py > $out/models.jsonl 2> $out/models.err
cut -c1-220 $out/models.jsonl
