"""Synthetic scenarios and context builders shared by bench.py, __graft_entry__.smoke() and the tests: the launch-file
defaults of path_integral_nn / path_integral_bf (SRC/launch/path_integral_{nn,bf}.launch) on the synthetic ellipse track."""
import numpy as np

from .params import BF_DEFAULTS, NN_DEFAULTS, CostParams, ellipse_start_state


def cost_params_for(costmap, **over):
    r_c1, r_c2, trs = costmap.transform()
    cp = CostParams(r_c1=r_c1, r_c2=r_c2, trs=trs)
    for k, v in over.items():
        setattr(cp, k, v)
    return cp


def make_context(kind, models, costmap, cp, num_rollouts, tag="autorally_nnet", negate_yaw_der=True, theta=None, structure=None, **kw):
    """An MppiContext configured like the reference's main() configures its controller.  `theta` / `structure` override the
    model looked up under `tag` (any NeuralNetModel layer pack)."""
    from .capi import MppiContext
    d = NN_DEFAULTS if kind == "nn" else BF_DEFAULTS
    ctx = MppiContext(dynamics=kind, num_rollouts=num_rollouts, num_timesteps=kw.pop("num_timesteps", d["num_timesteps"]),
                      hz=d["hz"], optimization_stride=kw.pop("optimization_stride", d["optimization_stride"]),
                      gamma=kw.pop("gamma", d["gamma"]), bdim=d["bdim"], **kw)
    if kind == "nn":
        if theta is None:
            theta, structure = models[tag + "_theta"], models[tag + "_structure"]
        ctx.set_nn_params(theta, structure)
        ctx.set_negate_yaw_der(negate_yaw_der)
    else:
        ctx.set_bf_params(models["basis_function_W"])
    ctx.set_control_ranges(np.asarray(d["control_ranges"], np.float32).reshape(4))
    ctx.set_exploration_std(d["exploration_std"])
    ctx.set_cost_params(cp)
    ctx.set_costmap(costmap)
    return ctx


def random_network(structure, seed=0, scale=0.6):
    """Packed [W1|b1|W2|b2|...] (row-major out x in, PI/neural_net_model.cu:125-141) random weights for a layer pack; the
    output layer is scaled so that the synthetic car's accelerations stay plausible."""
    rng = np.random.default_rng(seed)
    parts = []
    structure = list(structure)
    for l in range(len(structure) - 1):
        nin, nout = structure[l], structure[l + 1]
        s = scale / np.sqrt(nin) * (2.0 if l + 2 == len(structure) else 1.0)
        parts.append((rng.standard_normal((nout, nin)) * s).astype(np.float32).ravel())
        parts.append((rng.standard_normal(nout) * 0.1).astype(np.float32))
    return np.concatenate(parts), np.asarray(structure, np.int32)


def warm_controls(T, kind="nn"):
    """A plausible non-trivial nominal control sequence (gentle left turn, positive throttle)."""
    t = np.arange(T)
    U = np.stack([0.12 + 0.05 * np.sin(t / 11.0), 0.35 + 0.1 * np.cos(t / 17.0)], 1).astype(np.float32)
    return U


def default_state(speed=5.0):
    return ellipse_start_state(speed=speed)


def top_state(speed=4.0):
    """On the centreline at the flat top of the ellipse, (x, y, yaw) = (0, b, pi): curvature radius a^2/b = 33 m, so
    a good share of the rollouts stays on the track for the whole horizon and the importance weights are spread over
    many rollouts (normaliser ~ 30 at gamma 0.15) instead of collapsing onto the single best one."""
    return np.array([0.0, 12.0, np.pi, 0.0, speed, 0.0, 0.0], np.float32)


def straight_controls(T, steer=0.0, throttle=0.3):
    return np.tile(np.array([steer, throttle], np.float32), (T, 1))
