#!/bin/bash
# first GPU contact: parity tests, a small and the full bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --phonons 10000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_1e7.json 2> gpurun_out/bench_1e7.err
for spl in 1 4; do
  timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --steps-per-launch $spl > gpurun_out/bench_1e8_spl$spl.json 2> gpurun_out/bench_1e8_spl$spl.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --phonons 10000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_1e7.json gpurun_out/bench_1e8_spl1.json gpurun_out/bench_1e8_spl4.json
