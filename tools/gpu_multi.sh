#!/bin/bash
# multi-device host run (one process, NCCL all-reduce of the tallies): test + timing.  usage: gpurun --gpus N -- bash tools/gpu_multi.sh
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "end_to_end" 2>&1 | tail -5
ng=$(nvidia-smi -L | wc -l)
devs=$(seq -s, 0 $((ng - 1)))
python - <<PY
import json, sys
sys.path.insert(0, '.')
from psim_b200 import configs
configs.save(configs.si_ge_grid(num_phonons=100_000_000 * $ng).to_dict(), '/tmp/sige.json')
configs.save(configs.linear_sides(sim_type=1, step_interval=4).to_dict(), '/tmp/per.json')
PY
for f in /tmp/sige.json /tmp/per.json; do
PSIM_SEED=3 PSIM_DEVICES=$devs PSIM_TIMING=1 psim_b200/bin/psim $f 2>&1 | grep -v "^psim timing.*\(prepare\|epilogue\)" | tail -8
md5sum /tmp/*_$(basename $f .json).txt | cut -c1-12
PSIM_HOST_SUM=1 PSIM_SEED=3 PSIM_DEVICES=$devs psim_b200/bin/psim $f > /dev/null 2>&1
md5sum /tmp/*_$(basename $f .json).txt | cut -c1-12
done
tail -n +2 /tmp/ss_sige.txt | head -3
