#!/bin/bash
# Everything profiles/ is built from, on one GPU: bench (both arms), launch list, ncu captures, per-model wall-clock,
# memcheck.  usage (under gpurun): tools/gpu_evidence.sh [tag]
tag=${1:-r01}
out=gpurun_out/evidence
mkdir -p $out
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_reference_cpu.json 2> $out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_bench.log 2>&1
# DRAM traffic of every launch of one job (4 launches), then the two --set full captures (long window, recorded window)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:drift_kernel -c 4 --csv \
    --log-file $out/${tag}_dram_per_launch.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $out/dram_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 0 -c 1 -f -o $out/prof_long \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $out/ncu_long.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 1 -c 1 -f -o $out/prof_rec \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $out/ncu_rec.log 2>&1
# periodic model with 1000 sensors: recorded tallies go straight to global memory in difference form (launch 3 of 8)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 2 -c 1 -f -o $out/prof_per \
    python tools/profile_model.py sides_per > $out/ncu_per.log 2>&1
timeout 900 python tools/model_walltimes.py 2> $out/models.err > $out/${tag}_models.jsonl
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --phonons 200000 --steps 1 --warmup 0 --no-cpu-baseline \
    > $out/${tag}_memcheck.log 2>&1
tail -3 $out/${tag}_memcheck.log
cut -c1-250 $out/${tag}_bench_n1.json
cut -c1-400 $out/${tag}_bench_reference_cpu.json
ls -la $out
