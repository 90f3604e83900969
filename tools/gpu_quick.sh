#!/bin/bash
# Quick GPU loop: GPU tests, then per-window kernel times of the bench workload and a short bench line.
out=gpurun_out/quick
mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
tail -6 $out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench.json 2> $out/bench.err
cut -c1-200 $out/bench.json
timeout 300 python tools/gpu_windows.py > $out/windows.jsonl 2> $out/windows.err
cat $out/windows.jsonl
timeout 300 python tools/profile_model.py linear 2>&1 | tail -1
