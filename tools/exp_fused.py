"""Fused Philox (noise drawn inside the rollout kernel) against the stand-alone sampler kernel: device-resident step time
at several sizes.  Run on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autorally_b200.params import make_ellipse_costmap  # noqa: E402
from autorally_b200.scenarios import cost_params_for, make_context, straight_controls, top_state  # noqa: E402

models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
costmap = make_ellipse_costmap()
cp = cost_params_for(costmap)
state, U = top_state(4.0), straight_controls(100)
for kind, N, variant, tag in (("nn", 1 << 20, 0, "autorally_nnet"), ("nn", 131072, 0, "autorally_nnet"), ("nn", 32768, 0, "autorally_nnet"),
                              ("nn", 1920, 11, "autorally_nnet"), ("nn", 1920, 1, "autorally_nnet"), ("bf", 1 << 20, 0, "autorally_nnet"),
                              ("nn", 1 << 18, 0, "wider_deeper")):
    for fused in (0, 1):
        with make_context(kind, models, costmap, cp, N, variant=variant, tag=tag, negate_yaw_der=(tag != "wider_deeper")) as ctx:
            ctx.set_fused_noise(fused)
            ctx.compute_control(state, U)
            steps = 3 if N >= (1 << 18) else 20
            ctx.run_resident(2)
            best = min(ctx.run_resident(steps)[0] / steps for _ in range(3))
            ms, rk = ctx.run_resident(steps, time_rollout=True)
            print("%s %-14s N=%-8d variant=%-2d fused=%d  step %.4f ms  rollout kernel %.4f ms  launches/step %d" % (
                kind, tag, N, ctx.resolved_variant(), fused, best, rk / steps, ctx.last_launch_count() // steps), flush=True)
