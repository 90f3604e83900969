#!/bin/bash
# One GPU call: the GPU test suite, the A/B of the build variants, a bench line and the per-model wall-clocks.
out=gpurun_out/check
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
tail -4 $out/pytest_gpu.log
timeout 600 python tools/gpu_ab_variants.py run > $out/ab.jsonl 2> $out/ab.err
cat $out/ab.jsonl
timeout 400 python bench.py --steps 5 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
cut -c1-200 $out/bench_n1.json
timeout 400 python tools/model_walltimes.py > $out/models.jsonl 2> $out/models.err
cut -c1-220 $out/models.jsonl
