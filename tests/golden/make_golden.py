"""Generates the committed golden fixtures from the REFERENCE's own Python model.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports, unmodified, the reference's ml_pipeline (``utils.setup_model``,
``utils.npz_to_torch_model``, ``utils.compute_state_ders``; matplotlib is not installed and is
stubbed in ``sys.modules`` -- it is only used for plots), loads the reference's shipped weights and
writes

* ``ref_models.npz``   -- the reference weight files re-packed as float32 fixtures
  (``[W1|b1|...]`` packing of PI/neural_net_model.cu:120-141) so nothing needs /root/reference at
  test time;
* ``ref_dynamics.npz`` -- float64 outputs of the reference model: single-step state derivatives for
  seeded (state, control) pairs and 100-step explicit-Euler rollouts (dt = 0.02, the BASELINE
  config-1 workload) stepping ``model(x)`` + ``compute_state_ders`` exactly as
  ``train_dynamics_model.generate_predictions`` (ML/train_dynamics_model.py:249-282) does.

The CPU oracle is pinned against these in tests/test_oracle_golden.py.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/autorally_control/src/path_integral"
ML = os.path.join(REF, "scripts", "ml_pipeline")
MODELS = os.path.join(REF, "params", "models")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def import_reference_utils():
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    sys.path.insert(0, ML)
    import utils  # noqa: the reference's ml_pipeline/utils.py
    return utils


def main():
    import torch
    from autorally_b200.params import load_bf_npz, load_nn_npz

    utils = import_reference_utils()
    out_models = {}
    for tag, fname in (("autorally_nnet", "autorally_nnet_09_12_2018.npz"),
                       ("gazebo_nnet", "gazebo_nnet_09_12_2018.npz"),
                       ("shallow", "shallow_network_08_20_2020.npz"),
                       ("wider_deeper", "wider_deeper_network_08_20_2020.npz")):
        theta, structure = load_nn_npz(os.path.join(MODELS, fname))
        out_models[tag + "_theta"] = theta
        out_models[tag + "_structure"] = structure
    out_models["basis_function_W"] = load_bf_npz(os.path.join(MODELS, "basis_function_09_12_2018.npz"))
    np.savez(os.path.join(HERE, "ref_models.npz"), **out_models)

    golden = {}
    rng = np.random.default_rng(20261018)
    for tag, fname, layers, negate in (("autorally_nnet", "autorally_nnet_09_12_2018.npz", [6, 32, 32, 4], True),
                                       ("wider_deeper", "wider_deeper_network_08_20_2020.npz", [6, 64, 64, 64, 64, 4], False)):
        model = utils.setup_model(layers=layers, verbose=False)
        model = utils.npz_to_torch_model(os.path.join(MODELS, fname), model)
        model.eval()
        # single-step derivatives
        n = 256
        states = np.zeros((n, 7))
        states[:, 0:2] = rng.uniform(-20, 20, (n, 2))
        states[:, 2] = rng.uniform(-np.pi, np.pi, n)
        states[:, 3] = rng.uniform(-0.3, 0.3, n)
        states[:, 4] = rng.uniform(0.0, 12.0, n)
        states[:, 5] = rng.uniform(-2.0, 2.0, n)
        states[:, 6] = rng.uniform(-2.0, 2.0, n)
        ctrls = np.stack([rng.uniform(-0.99, 0.99, n), rng.uniform(-0.99, 0.65, n)], 1)
        ders = np.zeros((n, 7))
        with torch.no_grad():
            for i in range(n):
                x = torch.tensor([states[i, 3], states[i, 4], states[i, 5], states[i, 6], ctrls[i, 0], ctrls[i, 1]])
                y = model(x.double()).numpy()
                ders[i] = utils.compute_state_ders(states[i], y, negate_yaw_der=negate)
        golden[tag + "_step_states"] = states
        golden[tag + "_step_ctrls"] = ctrls
        golden[tag + "_step_ders"] = ders
        # 100-step Euler rollouts, dt = 0.02
        nr, T, dt = 24, 100, 0.02
        s0 = np.array([20.0, 0.0, np.pi / 2, 0.0, 5.0, 0.0, 0.0])
        U = np.stack([0.15 * np.sin(np.arange(T) / 9.0), 0.3 + 0.1 * np.cos(np.arange(T) / 13.0)], 1)
        eps = rng.standard_normal((nr, T, 2))
        nu = np.array([0.275, 0.3])
        lo, hi = np.array([-0.99, -0.99]), np.array([0.99, 0.65])
        traj = np.zeros((nr, T + 1, 7))
        with torch.no_grad():
            for r in range(nr):
                s = s0.copy()
                traj[r, 0] = s
                for t in range(T):
                    u = np.clip(U[t] + eps[r, t] * nu, lo, hi)
                    x = torch.tensor([s[3], s[4], s[5], s[6], u[0], u[1]])
                    y = model(x.double()).numpy()
                    s = s + utils.compute_state_ders(s, y, negate_yaw_der=negate) * dt
                    traj[r, t + 1] = s
        golden[tag + "_roll_state0"] = s0
        golden[tag + "_roll_U"] = U
        golden[tag + "_roll_eps"] = eps
        golden[tag + "_roll_traj"] = traj
        golden[tag + "_negate_yaw_der"] = np.array(int(negate))
    np.savez_compressed(os.path.join(HERE, "ref_dynamics.npz"), **golden)
    print("wrote", os.path.join(HERE, "ref_models.npz"), os.path.join(HERE, "ref_dynamics.npz"))


if __name__ == "__main__":
    main()
