#!/bin/bash
# A/B of the "merge_cells" levels on the shipped models (kernel ms per run): usage (under gpurun): tools/gpu_ab_merge.sh [tag] ["levels"] ["models"]
tag=${1:-abm}
levels=${2:-"1 2"}
models=${3:-"kinked sige sides_ss linear sides_per"}
out=gpurun_out/$tag
mkdir -p $out
for model in $models; do
  for lv in $levels; do
    echo "== $model merge_cells=$lv" >> $out/ab.log
    PSIM_OPTS=merge_cells=$lv timeout 300 python tools/profile_model.py $model 2>&1 | tail -2 >> $out/ab.log
  done
done
cat $out/ab.log
