"""Small launches of the kernels added late in round 2 (layer-pipeline, run-time layer, column-sliced tensor-core), to be
run under compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box:
    compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autorally_b200.params import make_ellipse_costmap  # noqa: E402
from autorally_b200.scenarios import cost_params_for, make_context, random_network, straight_controls, top_state  # noqa: E402

models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
costmap = make_ellipse_costmap()
cp = cost_params_for(costmap)
T = 40
state, U = top_state(4.0), straight_controls(T)
cases = [("wider_deeper", None, 12, 64), ("wider_deeper", None, 12, 192), ("wider_deeper", None, 10, 256), ("wider_deeper", None, 11, 128),
         ("autorally_nnet", None, 11, 64), (None, (6, 20, 33, 7, 4), 11, 128), (None, (6, 128, 128, 4), 11, 64), (None, (6, 48, 4), 11, 1216)]
for tag, st, variant, N in cases:
    kw = {}
    if st is not None:
        kw["theta"], kw["structure"] = random_network(st, seed=3)
    for fused in (0, 1):
        with make_context("nn", models, costmap, cp, N, variant=variant, tag=tag or "autorally_nnet", negate_yaw_der=(tag != "wider_deeper"),
                          num_timesteps=T, **kw) as ctx:
            ctx.set_fused_noise(fused)
            r = ctx.compute_control(state, U)
            print(tag or st, "variant", ctx.resolved_variant(), "N", N, "fused", fused, "normalizer %.4f" % r["normalizer"], flush=True)
