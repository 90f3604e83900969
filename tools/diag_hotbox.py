import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
from psim_b200 import configs, lib as psim
from tests import common as T
from tests.test_gpu_round2 import hot_box
from tests.gpu_runner import gpu_run_case
dt, n, steps, kernel, spl = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
model = T.load_model(hot_box(dt, n, steps))
t0 = time.time()
try:
    r = gpu_run_case(model, 3, steps_per_launch=spl, options={"kernel": kernel}, finish=False)
    print("ok", sys.argv[1:], round(time.time() - t0, 2), "s events", r["stats"][0]["events"], "ds", r["stats"][0]["drift_steps"], flush=True)
except Exception as e:
    print("exc", sys.argv[1:], round(time.time() - t0, 2), str(e)[:100], flush=True)
