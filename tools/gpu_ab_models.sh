#!/bin/bash
# A/B of library options on the shipped models (kernel ms per run): usage (under gpurun): tools/gpu_ab_models.sh [tag]
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
for model in kinked sides_per sides_trans sides_ss linear sige; do
  for slots in 64 128; do
    echo "== $model queue_slots=$slots" >> $out/ab.log
    PSIM_OPTS=queue_slots=$slots timeout 300 python tools/profile_model.py $model 2>&1 | tail -2 >> $out/ab.log
  done
done
cat $out/ab.log
