"""Work cycles per tick of the roles of the layer-pipeline kernel (library built with -DPIPE_EXP=9).  GPU box only."""
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
with make_context("nn", models, costmap, cp, N, variant=12, tag="wider_deeper", negate_yaw_der=False) as ctx:
    ctx.compute_control(top_state(4.0), straight_controls(100))
    ctx.compute_control(top_state(4.0), straight_controls(100))
    V = ctx.sampled_controls()[0][:40]
for w in range(5):
    print("warp %d work cycles, ticks 200..207:" % w, " ".join("%5.0f" % x for x in V[w * 8:(w + 1) * 8, 0]))
