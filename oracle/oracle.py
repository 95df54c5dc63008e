"""ctypes binding of the CPU oracle (oracle/mppi_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product (autorally_b200/, include/) never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmppi_oracle.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)


class _CostParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in (
        "desired_speed", "speed_coeff", "track_coeff", "max_slip_ang", "slip_penalty", "track_slop",
        "crash_coeff", "steering_coeff", "throttle_coeff", "boundary_threshold", "discount")] + [
        ("num_timesteps", ctypes.c_int), ("grid_res", ctypes.c_int),
        ("r_c1", ctypes.c_float * 3), ("r_c2", ctypes.c_float * 3), ("trs", ctypes.c_float * 3),
        ("l1_cost", ctypes.c_int)]


class _Dynamics(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("theta", c_float_p), ("net_structure", c_int_p),
                ("num_layers", ctypes.c_int), ("bdim_y", ctypes.c_int), ("dt", ctypes.c_float),
                ("negate_yaw_der", ctypes.c_int), ("ctrl_lo", ctypes.c_float * 2), ("ctrl_hi", ctypes.c_float * 2)]


class _Costmap(ctypes.Structure):
    _fields_ = [("channel0", c_float_p), ("width", ctypes.c_int), ("height", ctypes.c_int)]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "mppi_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


class Oracle:
    """One configured reference controller (dynamics + costs + costmap), CPU float32."""

    def __init__(self, kind, theta, net_structure=None, dt=0.02, negate_yaw_der=True,
                 control_ranges=((-0.99, 0.99), (-0.99, 0.65)), bdim_y=1,
                 cost_params=None, costmap=None):
        self.kind = {"nn": 0, "bf": 1}[kind]
        self.theta = _f32(theta).reshape(-1).copy()
        self.net_structure = np.ascontiguousarray(net_structure if net_structure is not None else [0], np.int32)
        self._dyn = _Dynamics()
        self._dyn.kind = self.kind
        self._dyn.theta = _fp(self.theta)
        self._dyn.net_structure = self.net_structure.ctypes.data_as(c_int_p)
        self._dyn.num_layers = len(self.net_structure) if self.kind == 0 else 0
        self._dyn.bdim_y = int(bdim_y)
        self._dyn.dt = float(np.float32(dt))
        self._dyn.negate_yaw_der = int(bool(negate_yaw_der))
        for j in range(2):
            self._dyn.ctrl_lo[j] = float(np.float32(control_ranges[j][0]))
            self._dyn.ctrl_hi[j] = float(np.float32(control_ranges[j][1]))
        self._cp = _CostParams()
        self._map = _Costmap()
        self._map_data = None
        if cost_params is not None:
            self.set_cost_params(cost_params)
        if costmap is not None:
            self.set_costmap(costmap.channel0, costmap.width, costmap.height)

    # cost_params: autorally_b200.params.CostParams (duck-typed) -----------------------------
    def set_cost_params(self, cp):
        src = cp.to_struct()
        ctypes.memmove(ctypes.byref(self._cp), ctypes.byref(src), ctypes.sizeof(self._cp))

    def set_costmap(self, channel0, width, height):
        self._map_data = _f32(channel0).reshape(-1).copy()
        assert self._map_data.size == width * height
        self._map.channel0 = _fp(self._map_data)
        self._map.width, self._map.height = int(width), int(height)

    # ------------------------------------------------------------------------------------
    def rollouts(self, state, U, nu, eps, n_global=None, r_begin=0, opt_delay=1, threads=1):
        """rolloutKernel restatement.  eps: [n, T, 2].  Returns (V, costs, crash, final_state)."""
        eps = _f32(eps)
        n, T, _ = eps.shape
        V = eps.copy()
        costs = np.zeros(n, np.float32)
        crash = np.zeros(n, np.int32)
        fs = np.zeros((n, 7), np.float32)
        state, U, nu = _f32(state, 7), _f32(U, (T, 2)), _f32(nu, 2)
        lib().oracle_rollouts(ctypes.byref(self._dyn), ctypes.byref(self._cp), ctypes.byref(self._map),
                              int(n_global if n_global is not None else n), int(r_begin), int(n), int(T), int(opt_delay),
                              _fp(state), _fp(U), _fp(nu), _fp(V), _fp(costs), crash.ctypes.data_as(c_int_p),
                              _fp(fs), int(threads))
        return V, costs, crash, fs

    @staticmethod
    def weighting(costs, V, gamma):
        costs, V = _f32(costs), _f32(V)
        n, T, _ = V.shape
        w = np.zeros(n, np.float32)
        Unew = np.zeros((T, 2), np.float32)
        stats = np.zeros(3, np.float32)
        lib().oracle_weighting(_fp(costs), _fp(V), int(n), int(T), ctypes.c_float(gamma), _fp(w), _fp(Unew), _fp(stats))
        return w, Unew, stats

    @staticmethod
    def shard_partials(costs, V, gamma):
        costs, V = _f32(costs), _f32(V)
        n, T, _ = V.shape
        out = np.zeros(3 + 2 * T, np.float32)
        lib().oracle_shard_partials(_fp(costs), _fp(V), int(n), int(T), ctypes.c_float(gamma), _fp(out))
        return out

    @staticmethod
    def savitsky_golay(U, hist):
        U = _f32(U).copy()
        hist = _f32(hist, 4)
        lib().oracle_savitsky_golay(_fp(U), _fp(hist), int(U.shape[0]))
        return U

    def nominal_traj(self, state, U):
        U = _f32(U)
        T = U.shape[0]
        ss, cs = np.zeros((T, 7), np.float32), np.zeros((T, 2), np.float32)
        state = _f32(state, 7)
        lib().oracle_nominal_traj(ctypes.byref(self._dyn), _fp(state), _fp(U), int(T), _fp(ss), _fp(cs))
        return ss, cs

    def update_state(self, state, u):
        s, u = _f32(state, 7).copy(), _f32(u, 2).copy()
        lib().oracle_update_state(ctypes.byref(self._dyn), _fp(s), _fp(u))
        return s, u

    @staticmethod
    def slide_control_seq(U, hist, init_u, stride):
        U, hist, init_u = _f32(U).copy(), _f32(hist, 4).copy(), _f32(init_u, 2)
        lib().oracle_slide_control_seq(_fp(U), _fp(hist), _fp(init_u), int(U.shape[0]), int(stride))
        return U, hist

    @staticmethod
    def slide_state_seq(ss, stride):
        ss = _f32(ss).copy()
        lib().oracle_slide_state_seq(_fp(ss), int(ss.shape[0]), int(stride))
        return ss

    def compute_control(self, state, U, hist, nu, eps, gamma=0.15, opt_delay=1, threads=1):
        """computeControl(state) restatement.  eps: [num_iters, N, T, 2] (or [N, T, 2])."""
        eps = _f32(eps)
        if eps.ndim == 3:
            eps = eps[None]
        iters, n, T, _ = eps.shape
        U = _f32(U, (T, 2)).copy()
        V = np.zeros((n, T, 2), np.float32)
        costs = np.zeros(n, np.float32)
        crash = np.zeros(n, np.int32)
        w = np.zeros(n, np.float32)
        stats = np.zeros(3, np.float32)
        ss, cs = np.zeros((T, 7), np.float32), np.zeros((T, 2), np.float32)
        state, hist, nu = _f32(state, 7), _f32(hist, 4), _f32(nu, 2)
        lib().oracle_compute_control(ctypes.byref(self._dyn), ctypes.byref(self._cp), ctypes.byref(self._map),
                                     int(n), int(T), int(opt_delay), int(iters), ctypes.c_float(gamma),
                                     _fp(state), _fp(U), _fp(hist), _fp(nu), _fp(eps), _fp(V), _fp(costs),
                                     crash.ctypes.data_as(c_int_p), _fp(w), _fp(stats), _fp(ss), _fp(cs), int(threads))
        return dict(U=U, V=V, costs=costs, crash=crash, w=w, baseline=float(stats[0]), normalizer=float(stats[1]),
                    trajectory_cost=float(stats[2]), state_solution=ss, control_solution=cs)

    def dynamics_step(self, state, u):
        s, u = _f32(state, 7).copy(), _f32(u, 2).copy()
        sder = np.zeros(7, np.float32)
        lib().oracle_dynamics_step(ctypes.byref(self._dyn), _fp(s), _fp(u), _fp(sder))
        return s, sder

    def dynamics_rollouts(self, state, U, nu, eps):
        eps = _f32(eps)
        n, T, _ = eps.shape
        fs = np.zeros((n, 7), np.float32)
        state, U, nu = _f32(state, 7), _f32(U, (T, 2)), _f32(nu, 2)
        lib().oracle_dynamics_rollouts(ctypes.byref(self._dyn), int(n), int(T), _fp(state), _fp(U), _fp(nu), _fp(eps), _fp(fs))
        return fs


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    lib().oracle_philox4x32_10(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), out.ctypes.data_as(u32p))
    return out


def sample_noise(seed, call, r_begin, r_count, T, controller=0):
    eps = np.zeros((r_count, T, 2), np.float32)
    lib().oracle_sample_noise(ctypes.c_uint64(seed), ctypes.c_uint32(call), ctypes.c_uint32(controller), int(r_begin), int(r_count), int(T), _fp(eps))
    return eps


def make_oracle(kind, models, costmap, cp, tag="autorally_nnet", negate_yaw_der=True, bdim_y=None, theta=None, structure=None):
    """The oracle configured like the reference's main() configures its controller (launch-file defaults)."""
    from autorally_b200.params import BF_DEFAULTS, NN_DEFAULTS
    d = NN_DEFAULTS if kind == "nn" else BF_DEFAULTS
    if kind == "nn":
        if theta is None:
            theta, structure = models[tag + "_theta"], models[tag + "_structure"]
        return Oracle("nn", theta, structure, dt=1.0 / d["hz"], negate_yaw_der=negate_yaw_der,
                      control_ranges=d["control_ranges"], cost_params=cp, costmap=costmap)
    return Oracle("bf", models["basis_function_W"], None, dt=1.0 / d["hz"], control_ranges=d["control_ranges"],
                  bdim_y=bdim_y if bdim_y is not None else d["bdim"][1], cost_params=cp, costmap=costmap)
