#!/bin/bash
# weak-scaling runs of bench.py on one box.  usage (under gpurun --gpus N): tools/gpu_scale.sh [tag] ["8 4 2"] [extras]
# extras = 1 also runs the reference arm under torchrun and the one-process multi-device host run on all GPUs of the box
tag=${1:-r01}
ns=${2:-"8 4 2"}
out=gpurun_out/evidence
mkdir -p $out
for n in $ns; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > $out/${tag}_bench_n$n.json 2> $out/scale_$n.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_n$n.json')); print($n, d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])"
done
if [ "${3:-0}" = "1" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $out/scale_ref2.json 2> $out/scale_ref2.err; cut -c1-200 $out/scale_ref2.json
ng=$(nvidia-smi -L | wc -l)
NG=$ng PSIM_TIMING=1 python - <<'PY'
import json, os, sys, time
sys.path.insert(0, '.')
from psim_b200 import configs, lib as psim
ng = int(os.environ["NG"])
m = psim.Model(text=json.dumps(configs.si_ge_grid(num_phonons=100_000_000 * ng).to_dict()))
t0 = time.perf_counter()
st = m.run_devices(list(range(ng)), seed=1)
print("run_devices %d GPUs, %.0e phonons: wall %.3f s, kernel_ms (max) %.1f, drift-steps %.4g -> %.4g /s (kernel)" % (
    ng, 1e8 * ng, time.perf_counter() - t0, st.kernel_ms, st.drift_steps, st.drift_steps / (st.kernel_ms * 1e-3)))
PY
fi
