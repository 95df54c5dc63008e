"""Device-resident step time with / without (a) programmatic dependent launch and (b) the nominal trajectory on a side stream.
Run on the GPU box: MPPI_NO_PDL=1 / MPPI_NO_SPLIT_FINALIZE=1 switch the two off."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autorally_b200.params import ellipse_states, make_ellipse_costmap  # noqa: E402
from autorally_b200.scenarios import cost_params_for, make_context, straight_controls, top_state  # noqa: E402

models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
costmap = make_ellipse_costmap()
cp = cost_params_for(costmap)
tag = "pdl=%d split=%d" % (0 if os.environ.get("MPPI_NO_PDL") else 1, 0 if os.environ.get("MPPI_NO_SPLIT_FINALIZE") else 1)
for kind, N, B in (("nn", 1920, 1), ("nn", 32768, 1), ("nn", 131072, 1), ("nn", 1 << 20, 1), ("nn", 256, 4096), ("bf", 2560, 1), ("bf", 1 << 20, 1)):
    state = top_state(4.0) if B == 1 else ellipse_states(B)
    U = straight_controls(100) if B == 1 else np.broadcast_to(straight_controls(100), (B, 100, 2)).copy()
    with make_context(kind, models, costmap, cp, N, num_controllers=B) as ctx:
        ctx.compute_control(state, U)
        steps = 5 if N * B >= (1 << 18) else 50
        ctx.run_resident(3)
        best = min(ctx.run_resident(steps)[0] / steps for _ in range(5))
        print("%s  %s N=%-8d B=%-5d variant=%-2d  step %.4f ms" % (tag, kind, N, B, ctx.resolved_variant(), best), flush=True)
