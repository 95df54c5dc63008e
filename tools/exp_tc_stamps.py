"""Cycle stamps of one timestep of the tensor-core rollout kernel (library built with -DTC_EXP=9, tools/build_tc_exp.sh 9):
where the latency chain of a one-tile-per-SM launch goes.  Run on the GPU box with the experiment library in place."""
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
tag = sys.argv[1] if len(sys.argv) > 1 else "wider_deeper"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
names = {0: "step top", 1: "inputs stored", 2: "barrier", 3: "layer-1 MMAs issued", 4: "cost part A", 5: "layer-1 MMAs done",
         36: "kinematics", 37: "texels + next controls", 38: "output MMAs done", 39: "state updated"}
for l in range(1, 5):
    names.update({6 * l: "L%d chunk loaded" % l, 6 * l + 1: "L%d chunk activated" % l, 6 * l + 2: "L%d stored" % l,
                  6 * l + 3: "L%d barrier" % l, 6 * l + 4: "L%d MMAs issued" % l, 6 * l + 5: "L%d MMAs done" % l})
with make_context("nn", models, costmap, cp, N, variant=10, tag=tag, negate_yaw_der=(tag != "wider_deeper")) as ctx:
    ctx.compute_control(top_state(4.0), straight_controls(100))
    ctx.compute_control(top_state(4.0), straight_controls(100))
    V = ctx.sampled_controls()[0][:40]
prev = 0.0
for k in range(40):
    if V[k, 0] > 0 or k == 0:
        print("%2d %-26s %7.0f  (+%.0f)" % (k, names.get(k, ""), V[k, 0], V[k, 0] - prev))
        prev = V[k, 0]
