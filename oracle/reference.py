"""ctypes binding of oracle/_ref/libautorally_ref.so: the REFERENCE's own MPPIController built from
/root/reference by oracle/refbuild.py (see oracle/ref_harness.cu for what is reference and what is shim).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_ref_gpu_golden.py and bench.py's
reference legs.  Needs a CUDA device (the reference has no CPU path for computeControl).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libautorally_ref.so")

REF_NN_1920, REF_BF_2560, REF_NN_256, REF_NN_4096, REF_BF_256, REF_NN64_1920, REF_NN16_1920, REF_NN48_1920 = 0, 1, 2, 3, 4, 5, 6, 7
KIND_ROLLOUTS = {REF_NN_1920: 1920, REF_BF_2560: 2560, REF_NN_256: 256, REF_NN_4096: 4096, REF_BF_256: 256, REF_NN64_1920: 1920,
                 REF_NN16_1920: 1920, REF_NN48_1920: 1920}

c_float_p = ctypes.POINTER(ctypes.c_float)
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        l = ctypes.CDLL(LIB_PATH)
        l.ref_version.restype = ctypes.c_char_p
        _lib = l
    return _lib


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


class ReferenceController:
    """One MPPIController<DYNAMICS, MPPICosts, ROLLOUTS, BDIM_X, BDIM_Y> of the reference with its model and costs."""

    def __init__(self, kind, theta, costmap, cost_params, negate_yaw_der=True, control_ranges=(-0.99, 0.99, -0.99, 0.65),
                 exploration_std=(0.275, 0.3), init_u=(0.0, 0.0), hz=50, num_timesteps=100, optimization_stride=1,
                 gamma=0.15, num_iters=1):
        self.kind, self.T, self.iters = kind, num_timesteps, num_iters
        self.N = KIND_ROLLOUTS[kind]
        theta = _f32(theta).reshape(-1)
        ch0 = _f32(costmap.channel0).reshape(-1)
        cp = cost_params.to_struct()
        rng, nu, iu = _f32(control_ranges, 4), _f32(exploration_std, 2), _f32(init_u, 2)
        self._h = ctypes.c_void_p()
        rc = lib().ref_create(int(kind), _fp(theta), int(bool(negate_yaw_der)), _fp(rng), _fp(ch0), int(costmap.width),
                              int(costmap.height), ctypes.byref(cp), _fp(nu), _fp(iu), int(hz), int(num_timesteps),
                              int(optimization_stride), ctypes.c_float(gamma), int(num_iters), ctypes.byref(self._h))
        if rc != 0:
            raise RuntimeError("ref_create failed: %d" % rc)
        assert lib().ref_num_rollouts(self._h) == self.N

    def close(self):
        if getattr(self, "_h", None):
            lib().ref_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_controls(self, U, hist=None):
        U = _f32(U, (self.T, 2))
        hist = _f32(hist if hist is not None else np.zeros(4), 4)
        lib().ref_set_controls(self._h, _fp(U), _fp(hist))

    def get_controls(self):
        U, hist = np.zeros((self.T, 2), np.float32), np.zeros(4, np.float32)
        lib().ref_get_controls(self._h, _fp(U), _fp(hist))
        return U, hist

    def slide(self, stride):
        lib().ref_slide(self._h, int(stride))

    def compute_control(self, state, want_eps=True):
        """The reference's computeControl(state) on its own cuRAND draws; returns them as `eps`."""
        N, T = self.N, self.T
        state = _f32(state, 7)
        eps = np.zeros((self.iters, N, T, 2), np.float32) if want_eps else None
        U, ss, cs = np.zeros((T, 2), np.float32), np.zeros((T, 7), np.float32), np.zeros((T, 2), np.float32)
        sc, w = np.zeros(2, np.float32), np.zeros(N, np.float32)
        rc = lib().ref_compute_control(self._h, _fp(state), _fp(eps) if want_eps else None, _fp(U), _fp(ss), _fp(cs), _fp(sc), _fp(w))
        if rc != 0:
            raise RuntimeError("ref_compute_control failed: %d" % rc)
        return dict(eps=eps, U=U, state_solution=ss, control_solution=cs, normalizer=float(sc[0]), trajectory_cost=float(sc[1]), w=w)

    def rollout_costs(self, state, U, eps):
        """The reference's launchRolloutKernel on the given noise: raw costs [N] and sampled controls [N, T, 2]."""
        N, T = self.N, self.T
        state, U, eps = _f32(state, 7), _f32(U, (T, 2)), _f32(eps, (N, T, 2))
        costs, V = np.zeros(N, np.float32), np.zeros((N, T, 2), np.float32)
        rc = lib().ref_rollout_costs(self._h, _fp(state), _fp(U), _fp(eps), _fp(costs), _fp(V))
        if rc != 0:
            raise RuntimeError("ref_rollout_costs failed: %d" % rc)
        return costs, V

    def feedback_gains(self, state):
        """The reference's computeFeedbackGains(state) (DDP<...>::run, DDP/ddp.h:49-157) around the controller's current
        solution (call compute_control first): gains [T, 2, 7] and feedforward terms [T, 2]."""
        g, ff = np.zeros((self.T, 2, 7), np.float32), np.zeros((self.T, 2), np.float32)
        rc = lib().ref_feedback_gains(self._h, _fp(_f32(state, 7)), _fp(g), _fp(ff))
        if rc != 0:
            raise RuntimeError("ref_feedback_gains failed: %d" % rc)
        return g, ff

    def run_control_loop(self, pose, iterations, use_feedback_gains=False):
        """The reference's own runControlLoop (PI/run_control_loop.cuh:84-321) in debug mode -- two fresh controllers sharing
        this model and cost object, the loop integrating the model as the plant -- for `iterations` iterations (20 ms each,
        real time) from pose (x, y, heading).  Returns the per-iteration record, including the N(0,1) draws consumed."""
        n, T, N = int(iterations), self.T, self.N
        out = dict(states=np.zeros((n, 7), np.float32), controls=np.zeros((n, 2), np.float32), controller_used=np.zeros(n, np.int32),
                   trajectory_costs=np.zeros((n, 2), np.float32), U_actual=np.zeros((n, T, 2), np.float32),
                   U_predicted=np.zeros((n, T, 2), np.float32), gains=np.zeros((n, T, 2, 7), np.float32),
                   eps=np.zeros((n, N, T, 2), np.float32))
        rc = lib().ref_run_control_loop(self._h, _fp(_f32(pose, 3)), n, int(bool(use_feedback_gains)), _fp(out["states"]), _fp(out["controls"]),
                                        out["controller_used"].ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(out["trajectory_costs"]),
                                        _fp(out["U_actual"]), _fp(out["U_predicted"]), _fp(out["gains"]), _fp(out["eps"]))
        if rc != 0:
            raise RuntimeError("ref_run_control_loop failed: %d" % rc)
        return out

    def time_kernels(self, state, reps=20):
        """Device time (ms) of the reference's four GPU stages per computeControl iteration, kernels alone:
        {curand, rollout, normexp, weighted_reduction} (oracle/ref_harness.cu::time_kernels)."""
        ms = np.zeros(4, np.float32)
        rc = lib().ref_time_kernels(self._h, _fp(_f32(state, 7)), int(reps), _fp(ms))
        if rc != 0:
            raise RuntimeError("ref_time_kernels failed: %d" % rc)
        return dict(curand=float(ms[0]), rollout=float(ms[1]), normexp=float(ms[2]), weighted_reduction=float(ms[3]))

    def time_compute_control(self, state, reps=20):
        ms = ctypes.c_float(0)
        rc = lib().ref_time_compute_control(self._h, _fp(_f32(state, 7)), int(reps), ctypes.byref(ms))
        if rc != 0:
            raise RuntimeError("ref_time_compute_control failed: %d" % rc)
        return ms.value
