#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 > gpurun_out/pytest_gpu11.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu11.log
tail -3 gpurun_out/pytest_gpu11.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/e_bench1.json 2> gpurun_out/e_bench1.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/e_ref1.json 2> gpurun_out/e_ref1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/e_bench2.json 2> gpurun_out/e_bench2.err
cat gpurun_out/e_bench1.json gpurun_out/e_ref1.json gpurun_out/e_bench2.json
tail -3 gpurun_out/e_bench2.err
