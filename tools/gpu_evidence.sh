#!/bin/bash
# Everything profiles/<tag>_* is built from, on one GPU.  usage (under gpurun): tools/gpu_evidence.sh [tag]
# Afterwards, here: python tools/ncu_summary.py <tag>   (raw-page CSVs, <tag>_ncu_summary.json with the hash of the sources the
# captures were taken on, <tag>_sass_histogram.txt), then the bench lines are taken with that summary committed.
tag=${1:-r02}
out=gpurun_out/evidence
mkdir -p $out
B="python bench.py --no-cpu-baseline --no-models --no-kinked"
python -c "import bench; print(bench.csrc_sha16())" > $out/${tag}_csrc_sha16.txt
python -c "import bench; print(bench.device_sha16())" > $out/${tag}_device_sha16.txt
# launch list of a short bench run: which kernels run, and their share of the GPU time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    $B --steps 2 --warmup 1 > $out/launches_bench.log 2>&1
# the four launches of ONE bench job: DRAM bytes, warp / thread instructions, time  (the job's drift-steps are in the JSON line)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum \
    --clock-control none -k regex:drift_kernel -c 4 --csv --log-file $out/${tag}_job_counters.csv $B --steps 1 --warmup 0 > $out/${tag}_job_counters_bench.json 2> $out/job_counters.err
# --set full captures: bench job long window / recorded window, periodic 1000-sensor model, kinked wire
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 0 -c 1 -f -o $out/prof_long \
    python tools/profile_model.py sige > $out/ncu_long.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 1 -c 1 -f -o $out/prof_rec \
    python tools/profile_model.py sige > $out/ncu_rec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 2 -c 1 -f -o $out/prof_per \
    python tools/profile_model.py sides_per > $out/ncu_per.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 2 -c 1 -f -o $out/prof_kinked \
    python tools/profile_model.py kinked > $out/ncu_kinked.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 6 -c 1 -f -o $out/prof_kinked_rec \
    python tools/profile_model.py kinked > $out/ncu_kinked_rec.log 2>&1
# per-launch times and instructions of every shipped model (which windows the time goes to)
for m in sige kinked sides_ss sides_per linear; do
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:drift_kernel --csv \
      --log-file $out/${tag}_launches_$m.csv python tools/profile_model.py $m > $out/launches_$m.log 2>&1
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --phonons 200000 --steps 1 --warmup 0 --no-cpu-baseline --no-models \
    > $out/${tag}_memcheck.log 2>&1
tail -3 $out/${tag}_memcheck.log
cat $out/${tag}_csrc_sha16.txt
ls -la $out
