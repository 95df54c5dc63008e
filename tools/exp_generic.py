"""The run-time layer kernel (one rollout per warp) against the tensor-core and thread-per-rollout kernels on networks other
than 6-32-32-4, at controller sizes.  Run on the GPU box."""
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
state, U = top_state(4.0), straight_controls(100)
cases = [("wider_deeper", None, v, N) for N in (1920, 4096, 16384, 65536) for v in (10, 11, 1)]
cases += [(None, st, 11, N) for st in ((6, 16, 16, 4), (6, 48, 4), (6, 64, 64, 4), (6, 128, 128, 4), (6, 20, 33, 7, 4)) for N in (1920, 16384)]
cases += [("autorally_nnet", None, v, 1920) for v in (9, 11)]
if len(sys.argv) > 1 and sys.argv[1] == "tc64":
    cases = [("wider_deeper", None, 10, N) for N in (1920, 4096, 16384, 18944, 32768)]
if len(sys.argv) > 1 and sys.argv[1] == "gen":
    cases = [c for c in cases if c[2] == 11 and c[3] <= 16384]
if len(sys.argv) > 1 and sys.argv[1] == "pipe":
    cases = [("wider_deeper", None, v, N) for N in (256, 1024, 1920, 4096, 8192, 16384) for v in (12, 10)]
if len(sys.argv) > 1 and sys.argv[1] == "lat":
    cases = [("autorally_nnet", None, v, N) for N in (256, 512, 768, 1024, 1920, 2368, 2432, 4096, 8192) for v in (9, 13)]
if len(sys.argv) > 1 and sys.argv[1] == "tc32":
    cases = [("autorally_nnet", None, 10, N) for N in (1920, 16384, 32768, 131072, 1 << 20)]
if len(sys.argv) > 1 and sys.argv[1] == "bf":
    for N in (2560, 1 << 20):
        with make_context("bf", models, costmap, cp, N) as ctx:
            ctx.compute_control(state, U)
            steps = 20 if N <= 16384 else 3
            ctx.run_resident(2)
            best = min(ctx.run_resident(steps)[0] / steps for _ in range(3))
            ms, rk = ctx.run_resident(steps, time_rollout=True)
            print("basis functions N=%-8d step %.4f ms  rollout kernel %.4f ms" % (N, best, rk / steps), flush=True)
    cases = []
for tag, st, variant, N in cases:
    kw = {}
    if st is not None:
        kw["theta"], kw["structure"] = random_network(st, seed=3)
    try:
        with make_context("nn", models, costmap, cp, N, variant=variant, tag=tag or "autorally_nnet", negate_yaw_der=(tag != "wider_deeper"), **kw) as ctx:
            ctx.compute_control(state, U)
            steps = 20 if N <= 16384 else 5
            ctx.run_resident(2)
            best = min(ctx.run_resident(steps)[0] / steps for _ in range(3))
            ms, rk = ctx.run_resident(steps, time_rollout=True)
            print("%-14s %-18s N=%-6d variant=%-2d step %.4f ms  rollout kernel %.4f ms" % (
                tag or "-", st or "", N, ctx.resolved_variant(), best, rk / steps), flush=True)
    except Exception as e:  # a variant a network does not support
        print("%-14s %-18s N=%-6d variant=%-2d : %s" % (tag or "-", st or "", N, variant, e), flush=True)
