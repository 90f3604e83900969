#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu13.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu13.log
tail -4 gpurun_out/pytest_gpu13.log
timeout 600 compute-sanitizer --tool memcheck python - > gpurun_out/memcheck.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from tests import common as T
from tests.gpu_runner import gpu_run_case
for name in ("sige", "sides_trans"):
    m = T.load_model(T.case_model(name), num_phonons=20000)
    r = gpu_run_case(m, 1, finish=False)
    print(name, r["stats"][0]["drift_steps"])
PY
tail -5 gpurun_out/memcheck.log
