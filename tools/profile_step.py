"""Runs a few complete computeControl pipelines for profiling under ncu (no timing, no oracle).

    python tools/profile_step.py --rollouts 1920 --steps 4 [--dynamics bf] [--variant 2] [--controllers B]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rollouts", type=int, default=1920)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--dynamics", default="nn")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--controllers", type=int, default=1)
    ap.add_argument("--timesteps", type=int, default=100)
    ap.add_argument("--tag", default="autorally_nnet", help="model in tests/golden/ref_models.npz (wider_deeper = 6-64-64-64-64-4)")
    a = ap.parse_args()
    from autorally_b200.params import ellipse_states, make_ellipse_costmap
    from autorally_b200.scenarios import cost_params_for, default_state, make_context, warm_controls
    models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
    costmap = make_ellipse_costmap()
    cp = cost_params_for(costmap)
    B = a.controllers
    state = default_state(5.0) if B == 1 else ellipse_states(B)
    T = a.timesteps
    U = warm_controls(T) if B == 1 else np.broadcast_to(warm_controls(T), (B, T, 2)).copy()
    with make_context(a.dynamics, models, costmap, cp, a.rollouts, tag=a.tag, negate_yaw_der=(a.tag != "wider_deeper"), variant=a.variant,
                      num_controllers=B, num_timesteps=T) as ctx:
        out = ctx.compute_control(state, U)
        ms, rk = ctx.run_resident(a.steps, time_rollout=True)
        print("variant", ctx.resolved_variant(), "ms/step", ms / a.steps, "rollout kernel ms", rk / a.steps)


if __name__ == "__main__":
    main()
