"""ctypes binding of the C ABI in include/mppi_b200.h (libmppi_b200.so).

This is test / bench plumbing: the product's host side is the C++ template layer under
include/autorally_control/path_integral/, which calls the same C ABI.  There is no CPU fallback:
if the shared library is missing this module raises, and on a box without a CUDA device
``MppiContext`` raises ``MppiError(MPPI_ERR_NO_DEVICE)``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .params import CostParams, CostParamsStruct, Costmap

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmppi_b200.so")

MPPI_DYNAMICS_NN, MPPI_DYNAMICS_BF = 0, 1
ROLLOUT_AUTO, ROLLOUT_THREAD1, ROLLOUT_THREAD2, ROLLOUT_HALF16, ROLLOUT_TENSOR, ROLLOUT_GENERIC, ROLLOUT_LAYER_PIPE, ROLLOUT_WARP32 = 0, 1, 2, 9, 10, 11, 12, 13
MPPI_ERR_NO_DEVICE = -4

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)

# every symbol include/mppi_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "mppi_version", "mppi_error_string", "mppi_config_default", "mppi_create", "mppi_destroy",
    "mppi_set_nn_params", "mppi_set_bf_params", "mppi_set_control_ranges", "mppi_set_negate_yaw_der",
    "mppi_set_cost_params", "mppi_set_costmap", "mppi_set_exploration_std", "mppi_set_gamma",
    "mppi_set_noise", "mppi_use_sampler", "mppi_seed", "mppi_sample_noise", "mppi_compute_control",
    "mppi_get_rollout_costs", "mppi_get_rollout_crash", "mppi_get_sampled_controls",
    "mppi_get_unsmoothed_controls", "mppi_shard_floats", "mppi_shard_begin", "mppi_shard_partials_device",
    "mppi_shard_finish", "mppi_run_resident", "mppi_get_stream", "mppi_synchronize",
    "mppi_last_launch_count", "mppi_resolved_variant", "mppi_measure_fp32_peak", "mppi_measure_copy_bandwidth",
    "mppi_shard_begin_async", "mppi_shard_finish_async", "mppi_shard_result", "mppi_set_stream",
    "mppi_comm_unique_id", "mppi_comm_init", "mppi_comm_destroy", "mppi_compute_control_sharded",
    "mppi_run_resident_sharded", "mppi_compute_control_async", "mppi_compute_control_wait",
    "mppi_bench_compute_control", "mppi_p2p_export", "mppi_p2p_init", "mppi_p2p_destroy", "mppi_set_fused_noise",
    "mppi_time_stages",
]


class MppiConfig(ctypes.Structure):
    _fields_ = [("dynamics", ctypes.c_int), ("num_rollouts", ctypes.c_int), ("num_timesteps", ctypes.c_int),
                ("num_controllers", ctypes.c_int), ("rollout_begin", ctypes.c_int), ("rollout_count", ctypes.c_int),
                ("hz", ctypes.c_int), ("optimization_stride", ctypes.c_int), ("gamma", ctypes.c_float),
                ("num_iters", ctypes.c_int), ("bdim_x", ctypes.c_int), ("bdim_y", ctypes.c_int),
                ("device", ctypes.c_int), ("rollout_variant", ctypes.c_int), ("controller_begin", ctypes.c_int),
                ("seed", ctypes.c_uint64)]


class MppiResult(ctypes.Structure):
    _fields_ = [("baseline", ctypes.c_float), ("normalizer", ctypes.c_float),
                ("trajectory_cost", ctypes.c_float), ("reserved", ctypes.c_float)]


class MppiError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        msg = load_library().mppi_error_string(code).decode()
        super().__init__("%s failed: %s (%d)" % (what, msg, code))


_lib = None


def load_library():
    """Loads libmppi_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.mppi_version.restype = ctypes.c_char_p
        lib.mppi_error_string.restype = ctypes.c_char_p
        lib.mppi_error_string.argtypes = [ctypes.c_int]
        lib.mppi_set_noise.argtypes = [ctypes.c_void_p, c_float_p, ctypes.c_size_t]
        lib.mppi_seed.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32]
        lib.mppi_measure_copy_bandwidth.argtypes = [ctypes.c_int, ctypes.c_size_t, c_float_p]
        lib.mppi_set_gamma.argtypes = [ctypes.c_void_p, ctypes.c_float]
        lib.mppi_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.mppi_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        lib.mppi_p2p_export.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p]
        lib.mppi_p2p_init.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


class MppiContext:
    """One mppi_ctx: an MPPIController's device side (B controllers in batched mode)."""

    def __init__(self, dynamics="nn", num_rollouts=1920, num_timesteps=100, num_controllers=1, rollout_begin=0,
                 rollout_count=0, hz=50, optimization_stride=1, gamma=0.15, num_iters=1, bdim=(8, 16), device=-1,
                 variant=ROLLOUT_AUTO, seed=1234, controller_begin=0):
        self.lib = load_library()
        cfg = MppiConfig()
        self.lib.mppi_config_default(ctypes.byref(cfg))
        cfg.dynamics = {"nn": MPPI_DYNAMICS_NN, "bf": MPPI_DYNAMICS_BF}[dynamics]
        cfg.num_rollouts, cfg.num_timesteps, cfg.num_controllers = num_rollouts, num_timesteps, num_controllers
        cfg.rollout_begin, cfg.rollout_count = rollout_begin, rollout_count
        cfg.hz, cfg.optimization_stride, cfg.gamma, cfg.num_iters = hz, optimization_stride, gamma, num_iters
        cfg.bdim_x, cfg.bdim_y, cfg.device, cfg.rollout_variant, cfg.seed = bdim[0], bdim[1], device, variant, seed
        cfg.controller_begin = controller_begin
        self.cfg = cfg
        self.B, self.T = num_controllers, num_timesteps
        self.n_local = rollout_count if rollout_count else num_rollouts - rollout_begin
        self.num_iters = num_iters
        self._ctx = ctypes.c_void_p()
        self._ck(self.lib.mppi_create(ctypes.byref(cfg), ctypes.byref(self._ctx)), "mppi_create")

    def _ck(self, code, what):
        if code != 0:
            raise MppiError(code, what)

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.mppi_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- parameters ------------------------------------------------------------------
    def set_nn_params(self, theta, structure):
        theta = _f32(theta).reshape(-1)
        structure = np.ascontiguousarray(structure, np.int32)
        self._ck(self.lib.mppi_set_nn_params(self._ctx, _fp(theta), structure.ctypes.data_as(c_int_p), len(structure)), "mppi_set_nn_params")

    def set_bf_params(self, theta):
        theta = _f32(theta).reshape(-1)
        assert theta.size == 100
        self._ck(self.lib.mppi_set_bf_params(self._ctx, _fp(theta)), "mppi_set_bf_params")

    def set_control_ranges(self, ranges):
        r = _f32(ranges).reshape(4)
        self._ck(self.lib.mppi_set_control_ranges(self._ctx, _fp(r)), "mppi_set_control_ranges")

    def set_negate_yaw_der(self, negate):
        self._ck(self.lib.mppi_set_negate_yaw_der(self._ctx, int(bool(negate))), "mppi_set_negate_yaw_der")

    def set_cost_params(self, cp: CostParams):
        s = cp.to_struct()
        self._ck(self.lib.mppi_set_cost_params(self._ctx, ctypes.byref(s)), "mppi_set_cost_params")

    def set_costmap(self, costmap: Costmap):
        ch = _f32(costmap.channel0)
        self._ck(self.lib.mppi_set_costmap(self._ctx, _fp(ch), costmap.width, costmap.height, 1), "mppi_set_costmap")

    def set_exploration_std(self, std):
        s = _f32(std).reshape(2)
        self._ck(self.lib.mppi_set_exploration_std(self._ctx, _fp(s)), "mppi_set_exploration_std")

    def set_gamma(self, gamma):
        self._ck(self.lib.mppi_set_gamma(self._ctx, float(gamma)), "mppi_set_gamma")

    # ---- noise -----------------------------------------------------------------------
    def set_noise(self, eps):
        eps = _f32(eps).reshape(-1)
        self._ck(self.lib.mppi_set_noise(self._ctx, _fp(eps), eps.size), "mppi_set_noise")

    def use_sampler(self):
        self._ck(self.lib.mppi_use_sampler(self._ctx), "mppi_use_sampler")

    def seed(self, seed, call_counter=0):
        self._ck(self.lib.mppi_seed(self._ctx, seed, call_counter), "mppi_seed")

    def set_fused_noise(self, mode):
        """1: the rollout kernel draws its own Philox noise, 0: stand-alone sampler kernel, -1: automatic."""
        self._ck(self.lib.mppi_set_fused_noise(self._ctx, int(mode)), "mppi_set_fused_noise")

    def sample_noise(self):
        eps = np.zeros((self.B, self.n_local, self.T, 2), np.float32)
        self._ck(self.lib.mppi_sample_noise(self._ctx, _fp(eps)), "mppi_sample_noise")
        return eps

    # ---- compute ---------------------------------------------------------------------
    def compute_control(self, state, U, hist=None):
        B, T = self.B, self.T
        state = _f32(state).reshape(B, 7)
        U = _f32(U).reshape(B, T, 2).copy()
        hist = _f32(hist if hist is not None else np.zeros((B, 4))).reshape(B, 4)
        ss, cs = np.zeros((B, T, 7), np.float32), np.zeros((B, T, 2), np.float32)
        res = (MppiResult * B)()
        self._ck(self.lib.mppi_compute_control(self._ctx, _fp(state), _fp(U), _fp(hist), _fp(ss), _fp(cs), res), "mppi_compute_control")
        out = dict(U=U, state_solution=ss, control_solution=cs,
                   baseline=np.array([r.baseline for r in res], np.float32),
                   normalizer=np.array([r.normalizer for r in res], np.float32),
                   trajectory_cost=np.array([r.trajectory_cost for r in res], np.float32))
        if B == 1:
            out = {k: v[0] for k, v in out.items()}
        return out

    def compute_control_async(self, state, U, hist=None):
        B, T = self.B, self.T
        state, U = _f32(state).reshape(B, 7), _f32(U).reshape(B, T, 2)
        hist = _f32(hist if hist is not None else np.zeros((B, 4))).reshape(B, 4)
        self._ck(self.lib.mppi_compute_control_async(self._ctx, _fp(state), _fp(U), _fp(hist)), "mppi_compute_control_async")

    def compute_control_wait(self):
        B, T = self.B, self.T
        U, ss, cs = np.zeros((B, T, 2), np.float32), np.zeros((B, T, 7), np.float32), np.zeros((B, T, 2), np.float32)
        res = (MppiResult * B)()
        self._ck(self.lib.mppi_compute_control_wait(self._ctx, _fp(U), _fp(ss), _fp(cs), res), "mppi_compute_control_wait")
        return self._result(U, ss, cs, res)

    def bench_compute_control(self, state, U, hist=None, reps=100):
        """Per-call host latencies (ms) of `reps` C-side mppi_compute_control calls; returns (latencies, final U)."""
        B, T = self.B, self.T
        state = _f32(state).reshape(B, 7)
        U = _f32(U).reshape(B, T, 2).copy()
        hist = _f32(hist if hist is not None else np.zeros((B, 4))).reshape(B, 4)
        lat = np.zeros(reps, np.float32)
        self._ck(self.lib.mppi_bench_compute_control(self._ctx, _fp(state), _fp(U), _fp(hist), int(reps), _fp(lat)), "mppi_bench_compute_control")
        return lat, U

    def rollout_costs(self):
        c = np.zeros((self.B, self.n_local), np.float32)
        self._ck(self.lib.mppi_get_rollout_costs(self._ctx, _fp(c)), "mppi_get_rollout_costs")
        return c[0] if self.B == 1 else c

    def rollout_crash(self):
        c = np.zeros((self.B, self.n_local), np.int32)
        self._ck(self.lib.mppi_get_rollout_crash(self._ctx, c.ctypes.data_as(c_int_p)), "mppi_get_rollout_crash")
        return c[0] if self.B == 1 else c

    def sampled_controls(self):
        v = np.zeros((self.B, self.n_local, self.T, 2), np.float32)
        self._ck(self.lib.mppi_get_sampled_controls(self._ctx, _fp(v)), "mppi_get_sampled_controls")
        return v[0] if self.B == 1 else v

    def unsmoothed_controls(self):
        u = np.zeros((self.B, self.T, 2), np.float32)
        self._ck(self.lib.mppi_get_unsmoothed_controls(self._ctx, _fp(u)), "mppi_get_unsmoothed_controls")
        return u[0] if self.B == 1 else u

    # ---- multi-GPU -------------------------------------------------------------------
    def shard_floats(self):
        return self.lib.mppi_shard_floats(self._ctx)

    def shard_begin(self, state, U, hist=None):
        B, T = self.B, self.T
        state = _f32(state).reshape(B, 7)
        U = _f32(U).reshape(B, T, 2)
        hist = _f32(hist if hist is not None else np.zeros((B, 4))).reshape(B, 4)
        self._ck(self.lib.mppi_shard_begin(self._ctx, _fp(state), _fp(U), _fp(hist)), "mppi_shard_begin")

    def shard_partials_ptr(self):
        p = c_float_p()
        self._ck(self.lib.mppi_shard_partials_device(self._ctx, ctypes.byref(p)), "mppi_shard_partials_device")
        return ctypes.cast(p, ctypes.c_void_p).value

    def shard_finish(self, gathered_dev_ptr, num_shards):
        B, T = self.B, self.T
        U = np.zeros((B, T, 2), np.float32)
        ss, cs = np.zeros((B, T, 7), np.float32), np.zeros((B, T, 2), np.float32)
        res = (MppiResult * B)()
        self._ck(self.lib.mppi_shard_finish(self._ctx, ctypes.cast(ctypes.c_void_p(gathered_dev_ptr), c_float_p), int(num_shards),
                                            _fp(U), _fp(ss), _fp(cs), res), "mppi_shard_finish")
        out = dict(U=U, state_solution=ss, control_solution=cs,
                   baseline=np.array([r.baseline for r in res], np.float32),
                   normalizer=np.array([r.normalizer for r in res], np.float32),
                   trajectory_cost=np.array([r.trajectory_cost for r in res], np.float32))
        if B == 1:
            out = {k: v[0] for k, v in out.items()}
        return out

    def _result(self, U, ss, cs, res):
        out = dict(U=U, state_solution=ss, control_solution=cs,
                   baseline=np.array([r.baseline for r in res], np.float32),
                   normalizer=np.array([r.normalizer for r in res], np.float32),
                   trajectory_cost=np.array([r.trajectory_cost for r in res], np.float32))
        return {k: v[0] for k, v in out.items()} if self.B == 1 else out

    def set_stream(self, cuda_stream):
        """Enqueue on the caller's stream (an int handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._ck(self.lib.mppi_set_stream(self._ctx, ctypes.c_void_p(cuda_stream)), "mppi_set_stream")

    def shard_begin_async(self, state=None, U=None, hist=None):
        B, T = self.B, self.T
        if state is None:
            self._ck(self.lib.mppi_shard_begin_async(self._ctx, None, None, None), "mppi_shard_begin_async")
            return
        state, U = _f32(state).reshape(B, 7), _f32(U).reshape(B, T, 2)
        hist = _f32(hist if hist is not None else np.zeros((B, 4))).reshape(B, 4)
        self._ck(self.lib.mppi_shard_begin_async(self._ctx, _fp(state), _fp(U), _fp(hist)), "mppi_shard_begin_async")

    def shard_finish_async(self, gathered_dev_ptr, num_shards, feed_back=False):
        self._ck(self.lib.mppi_shard_finish_async(self._ctx, ctypes.cast(ctypes.c_void_p(gathered_dev_ptr), c_float_p), int(num_shards),
                                                  int(bool(feed_back))), "mppi_shard_finish_async")

    def shard_result(self):
        B, T = self.B, self.T
        U, ss, cs = np.zeros((B, T, 2), np.float32), np.zeros((B, T, 7), np.float32), np.zeros((B, T, 2), np.float32)
        res = (MppiResult * B)()
        self._ck(self.lib.mppi_shard_result(self._ctx, _fp(U), _fp(ss), _fp(cs), res), "mppi_shard_result")
        return self._result(U, ss, cs, res)

    # ---- in-library NCCL exchange ----------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = ctypes.create_string_buffer(128)
        code = load_library().mppi_comm_unique_id(buf)
        if code:
            raise MppiError(code, "mppi_comm_unique_id")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, num_ranks: int):
        assert len(unique_id) == 128
        self._ck(self.lib.mppi_comm_init(self._ctx, unique_id, int(rank), int(num_ranks)), "mppi_comm_init")

    # ---- peer-memory exchange (no collective) ---------------------------------------
    def p2p_export(self, num_ranks: int) -> bytes:
        buf = ctypes.create_string_buffer(128)
        self._ck(self.lib.mppi_p2p_export(self._ctx, int(num_ranks), buf), "mppi_p2p_export")
        return buf.raw

    def p2p_init(self, all_handles: bytes, rank: int, num_ranks: int):
        assert len(all_handles) == 128 * num_ranks
        self._ck(self.lib.mppi_p2p_init(self._ctx, all_handles, int(rank), int(num_ranks)), "mppi_p2p_init")

    def compute_control_sharded(self, state, U, hist=None):
        B, T = self.B, self.T
        state = _f32(state).reshape(B, 7)
        U = _f32(U).reshape(B, T, 2).copy()
        hist = _f32(hist if hist is not None else np.zeros((B, 4))).reshape(B, 4)
        ss, cs = np.zeros((B, T, 7), np.float32), np.zeros((B, T, 2), np.float32)
        res = (MppiResult * B)()
        self._ck(self.lib.mppi_compute_control_sharded(self._ctx, _fp(state), _fp(U), _fp(hist), _fp(ss), _fp(cs), res),
                 "mppi_compute_control_sharded")
        return self._result(U, ss, cs, res)

    def run_resident_sharded(self, steps):
        el = ctypes.c_float(0)
        self._ck(self.lib.mppi_run_resident_sharded(self._ctx, int(steps), ctypes.byref(el)), "mppi_run_resident_sharded")
        return el.value

    # ---- measurement -----------------------------------------------------------------
    def run_resident(self, steps, time_rollout=False, flush_l2=False):
        el, rk = ctypes.c_float(0), ctypes.c_float(0)
        self._ck(self.lib.mppi_run_resident(self._ctx, int(steps), int(bool(flush_l2)), ctypes.byref(el),
                                            ctypes.byref(rk) if time_rollout else None), "mppi_run_resident")
        return el.value, (rk.value if time_rollout else None)

    def time_stages(self, reps=5):
        """Device ms of {sampler kernel, rollout kernel, weighting kernel, finalize kernel}, each timed on its own."""
        ms = (ctypes.c_float * 4)()
        self._ck(self.lib.mppi_time_stages(self._ctx, int(reps), ms), "mppi_time_stages")
        return dict(sampler=ms[0], rollout=ms[1], weighting=ms[2], finalize=ms[3])

    def last_launch_count(self):
        return self.lib.mppi_last_launch_count(self._ctx)

    def resolved_variant(self):
        return self.lib.mppi_resolved_variant(self._ctx)

    def synchronize(self):
        self._ck(self.lib.mppi_synchronize(self._ctx), "mppi_synchronize")


def measure_fp32_peak(device=-1):
    v = ctypes.c_float(0)
    code = load_library().mppi_measure_fp32_peak(device, ctypes.byref(v))
    if code:
        raise MppiError(code, "mppi_measure_fp32_peak")
    return v.value


def measure_copy_bandwidth(nbytes=1 << 30, device=-1):
    v = ctypes.c_float(0)
    code = load_library().mppi_measure_copy_bandwidth(device, nbytes, ctypes.byref(v))
    if code:
        raise MppiError(code, "mppi_measure_copy_bandwidth")
    return v.value
