#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
python -c "
import json
d=json.load(open('gpurun_out/scale_$n.json')); print($n, d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9)"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/scale_ref2.json 2> gpurun_out/scale_ref2.err; cat gpurun_out/scale_ref2.json | cut -c1-300
