"""Shared builders for the parity tests: the same configuration for the CPU oracle and the CUDA path."""
import numpy as np

from autorally_b200.params import BF_DEFAULTS, NN_DEFAULTS, CostParams, ellipse_start_state


def cost_params_for(costmap, **over):
    r_c1, r_c2, trs = costmap.transform()
    cp = CostParams(r_c1=r_c1, r_c2=r_c2, trs=trs)
    for k, v in over.items():
        setattr(cp, k, v)
    return cp


def make_oracle(kind, models, costmap, cp, tag="autorally_nnet", negate_yaw_der=True, bdim_y=None):
    from oracle.oracle import Oracle
    d = NN_DEFAULTS if kind == "nn" else BF_DEFAULTS
    if kind == "nn":
        return Oracle("nn", models[tag + "_theta"], models[tag + "_structure"], dt=1.0 / d["hz"], negate_yaw_der=negate_yaw_der,
                      control_ranges=d["control_ranges"], cost_params=cp, costmap=costmap)
    return Oracle("bf", models["basis_function_W"], None, dt=1.0 / d["hz"], control_ranges=d["control_ranges"],
                  bdim_y=bdim_y if bdim_y is not None else d["bdim"][1], cost_params=cp, costmap=costmap)


def make_context(kind, models, costmap, cp, num_rollouts, tag="autorally_nnet", negate_yaw_der=True, **kw):
    from autorally_b200.capi import MppiContext
    d = NN_DEFAULTS if kind == "nn" else BF_DEFAULTS
    ctx = MppiContext(dynamics=kind, num_rollouts=num_rollouts, num_timesteps=kw.pop("num_timesteps", d["num_timesteps"]),
                      hz=d["hz"], optimization_stride=kw.pop("optimization_stride", d["optimization_stride"]),
                      gamma=kw.pop("gamma", d["gamma"]), bdim=d["bdim"], **kw)
    if kind == "nn":
        ctx.set_nn_params(models[tag + "_theta"], models[tag + "_structure"])
        ctx.set_negate_yaw_der(negate_yaw_der)
    else:
        ctx.set_bf_params(models["basis_function_W"])
    ctx.set_control_ranges(np.asarray(d["control_ranges"], np.float32).reshape(4))
    ctx.set_exploration_std(d["exploration_std"])
    ctx.set_cost_params(cp)
    ctx.set_costmap(costmap)
    return ctx


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
