#!/bin/bash
# ncu --set full captures of single drift-kernel launches of shipped models.
# usage (under gpurun): tools/gpu_ncu.sh <tag> <model>:<launch index> ...      e.g. tools/gpu_ncu.sh r2d sige:0 sides_per:2 kinked:2
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
for spec in "$@"; do
  model=${spec%%:*}; idx=${spec##*:}
  timeout 800 ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s $idx -c 1 -f -o $out/prof_${model}_$idx \
      python tools/profile_model.py $model > $out/ncu_${model}_$idx.log 2>&1
  tail -1 $out/ncu_${model}_$idx.log
done
ls -la $out
