"""weight_reduce_kernel alone at several sizes: device time and achieved HBM bandwidth (8 B / rollout-step + 4 B / rollout
algorithmic).  Run on the GPU box (uses ncu-free CUDA-event timing through mppi_run_resident's per-kernel pass)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autorally_b200.params import make_ellipse_costmap  # noqa: E402
from autorally_b200.scenarios import cost_params_for, make_context, straight_controls, top_state  # noqa: E402
models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
costmap = make_ellipse_costmap(); cp = cost_params_for(costmap)
for N in (1 << 20, 131072, 1920):
    with make_context("nn", models, costmap, cp, N) as ctx:
        ctx.compute_control(top_state(4.0), straight_controls(100))
        steps = 5 if N > 100000 else 50
        ms, rk = ctx.run_resident(steps, time_rollout=True)
        print("N=%d step %.4f ms, rollout kernel %.4f ms, everything else %.4f ms" % (N, ms / steps, rk / steps, (ms - rk) / steps), flush=True)
