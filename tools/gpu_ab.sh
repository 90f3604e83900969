#!/bin/bash
# A/B of kernel options on the bench workload.  usage: tools/gpu_ab2.sh "<bench args>" ...
mkdir -p gpurun_out
for args in "$@"; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$args]', '%.4g' % d['value'], round(d['ms_per_step'],2), [r['ms'] for r in d['e2e']['runs']])"; done
