"""Generates tests/golden/ref_gpu_golden.npz from the REFERENCE's own controller running on a B200.

Run on the GPU box (the reference has no CPU path for computeControl):

    gpurun -- 'python tests/golden/make_ref_gpu_golden.py gpurun_out/ref_gpu_golden.npz'

and copy the result to tests/golden/.  It needs oracle/_ref/libautorally_ref.so, which oracle/refbuild.py
builds in the CPU container from /root/reference (rolloutKernel, normExpKernel, weightedReductionKernel,
NeuralNetModel / GeneralizedLinear / MPPICosts device code, savitskyGolay, computeNominalTraj -- the
reference's sources, unmodified).  Each case records the inputs (state, U, history, the cuRAND draws the
call consumed) and the reference's outputs, so the CPU oracle can be pinned against them without a GPU
(tests/test_oracle_golden.py) and the CUDA path on the GPU (tests/test_parity_gpu.py).
Cases use the 256-rollout instantiations to keep the fixture small (~0.5 MB).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main(out_path):
    from autorally_b200.params import make_ellipse_costmap
    from oracle import reference as ref
    from tests.common import cost_params_for, default_state, straight_controls, top_state, warm_controls
    models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
    costmap = make_ellipse_costmap()
    hist = np.array([0.1, 0.3, 0.11, 0.32], np.float32)
    out = {}
    cases = [
        ("nn_v5", ref.REF_NN_256, "autorally_nnet_theta", dict(), 5.0, 100, 1, (0.0, 0.0)),
        ("nn_v0", ref.REF_NN_256, "autorally_nnet_theta", dict(), 0.0, 100, 1, (0.0, 0.0)),
        ("nn_terms", ref.REF_NN_256, "autorally_nnet_theta",
         dict(steering_coeff=0.4, throttle_coeff=0.2, track_slop=0.05, l1_cost=True, max_slip_ang=0.4), 6.0, 60, 3, (0.0, 0.0)),
        ("bf_v5", ref.REF_BF_256, "basis_function_W", dict(desired_speed=6.0), 5.0, 100, 1, (0.0, -0.01)),
        # flat top of the ellipse: importance weights spread over many rollouts
        ("nn_top", ref.REF_NN_256, "autorally_nnet_theta", dict(), -4.0, 100, 1, (0.0, 0.0)),
        ("bf_top", ref.REF_BF_256, "basis_function_W", dict(desired_speed=6.0), -4.0, 100, 1, (0.0, -0.01)),
    ]
    for name, kind, key, over, speed, T, opt_delay, init_u in cases:
        cp = cost_params_for(costmap, **over)
        # negative speed selects the flat-top scenario (tests/common.py:top_state) at |speed|
        state, U = (default_state(speed), warm_controls(T)) if speed >= 0 else (top_state(-speed), straight_controls(T))
        with ref.ReferenceController(kind, models[key], costmap, cp, num_timesteps=T, optimization_stride=opt_delay,
                                     init_u=init_u) as rc:
            rc.set_controls(U, hist)
            r = rc.compute_control(state)
            costs, V = rc.rollout_costs(state, U, r["eps"][0])
        out[name + "/state"], out[name + "/U_in"], out[name + "/hist"] = state, U, hist
        out[name + "/eps"] = r["eps"][0]
        out[name + "/costs"], out[name + "/w"] = costs, r["w"]
        out[name + "/V_row0"], out[name + "/V_last"] = V[0], V[-1]
        out[name + "/U"], out[name + "/state_solution"], out[name + "/control_solution"] = r["U"], r["state_solution"], r["control_solution"]
        out[name + "/scalars"] = np.array([r["normalizer"], r["trajectory_cost"]], np.float32)
        out[name + "/cfg"] = np.array([T, opt_delay, int(kind in (ref.REF_BF_256, ref.REF_BF_2560))], np.int32)
        out[name + "/cost_over"] = np.array([over.get("steering_coeff", 0.0), over.get("throttle_coeff", 0.0), over.get("track_slop", 0.0),
                                             float(over.get("l1_cost", False)), over.get("max_slip_ang", 1.25),
                                             over.get("desired_speed", 8.0)], np.float32)
        print(name, "min cost %.4f  normalizer %.4f  trajectory_cost %.5f" % (costs.min(), r["normalizer"], r["trajectory_cost"]))
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_gpu_golden.npz"))
