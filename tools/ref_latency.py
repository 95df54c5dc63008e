"""The reference's own computeControl (oracle/_ref) for the three single-GPU configurations: a few calls each, for an ncu
launch list of ITS kernels (rolloutKernel / normExpKernel / weightedReductionKernel / cuRAND) on this B200, plus the
harness's wall and kernel-only times.  Run on the GPU box.

    python tools/ref_latency.py [calls]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autorally_b200.params import make_ellipse_costmap  # noqa: E402
from autorally_b200.scenarios import cost_params_for, straight_controls, top_state  # noqa: E402
from oracle import reference as ref  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
costmap = make_ellipse_costmap()
state, U = top_state(4.0), straight_controls(100)
for name, kind, theta, kw, cpo in (("path_integral_nn 1920x100", ref.REF_NN_1920, models["autorally_nnet_theta"], {}, {}),
                                   ("path_integral_bf 2560x100", ref.REF_BF_2560, models["basis_function_W"], dict(init_u=(0.0, -0.01)), dict(desired_speed=6.0)),
                                   ("wider_deeper 1920x100", ref.REF_NN64_1920, models["wider_deeper_theta"], dict(negate_yaw_der=False), {})):
    cp = cost_params_for(costmap, **cpo)
    with ref.ReferenceController(kind, theta, costmap, cp, **kw) as rc:
        rc.set_controls(U, np.zeros(4, np.float32))
        for _ in range(calls):
            rc.compute_control(state, want_eps=False)
        ms = rc.time_compute_control(state, reps=calls)
        k = rc.time_kernels(state, reps=calls)
        print("%s: reference computeControl %.3f ms/call; kernels alone %.3f ms (%s)" % (name, ms, sum(k.values()), ", ".join("%s %.4f" % kv for kv in k.items())), flush=True)
