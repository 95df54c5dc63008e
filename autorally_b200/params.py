"""Host-side parameter plumbing for the MPPI hot path (no GPU, no oracle).

Mirrors the reference's parameter sources so tests and the bench read like the reference's own
configuration:

* cost parameters: ``MPPICosts::CostParams`` (PI/costs.cuh:67-85) with the defaults of
  ``autorally_control/launch/path_integral_nn.launch:51-62``;
* costmap npz schema: ``SRC/scripts/track_generator.py:34-40`` and ``MPPICosts::loadTrackData``
  (PI/costs.cu:190-232) for the world -> texture transform;
* neural-net weights: ``NeuralNetModel::loadParams`` / ``paramsToDevice``
  (PI/neural_net_model.cu:73-150): npz keys ``dynamics_W{i}`` / ``dynamics_b{i}`` (float64), packed
  as ``[W1|b1|W2|b2|...]`` row-major float32;
* basis-function weights: ``GeneralizedLinear::loadParams`` (PI/generalized_linear.cu:92-108), key ``W``.

(PI/ = autorally_control/include/autorally_control/path_integral/ of the reference.)
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

STATE_DIM = 7
CONTROL_DIM = 2


class CostParamsStruct(ctypes.Structure):
    """C layout shared by the C-ABI (include/mppi_b200.h: mppi_cost_params) and the oracle."""

    _fields_ = [
        ("desired_speed", ctypes.c_float),
        ("speed_coeff", ctypes.c_float),
        ("track_coeff", ctypes.c_float),
        ("max_slip_ang", ctypes.c_float),
        ("slip_penalty", ctypes.c_float),
        ("track_slop", ctypes.c_float),
        ("crash_coeff", ctypes.c_float),
        ("steering_coeff", ctypes.c_float),
        ("throttle_coeff", ctypes.c_float),
        ("boundary_threshold", ctypes.c_float),
        ("discount", ctypes.c_float),
        ("num_timesteps", ctypes.c_int),
        ("grid_res", ctypes.c_int),
        ("r_c1", ctypes.c_float * 3),
        ("r_c2", ctypes.c_float * 3),
        ("trs", ctypes.c_float * 3),
        ("l1_cost", ctypes.c_int),
    ]


@dataclass
class CostParams:
    """Defaults: launch/path_integral_nn.launch:51-62 (BF launch file differs in desired_speed=6.0)."""

    desired_speed: float = 8.0
    speed_coeff: float = 4.25
    track_coeff: float = 200.0
    max_slip_ang: float = 1.25
    slip_penalty: float = 10.0
    track_slop: float = 0.0
    crash_coeff: float = 10000.0
    steering_coeff: float = 0.0
    throttle_coeff: float = 0.0
    boundary_threshold: float = 0.65
    discount: float = 0.1
    num_timesteps: int = 100
    grid_res: int = 0
    r_c1: tuple = (1.0, 0.0, 0.0)
    r_c2: tuple = (0.0, 1.0, 0.0)
    trs: tuple = (0.0, 0.0, 1.0)
    l1_cost: bool = False

    def to_struct(self) -> CostParamsStruct:
        s = CostParamsStruct()
        for name in ("desired_speed", "speed_coeff", "track_coeff", "max_slip_ang", "slip_penalty",
                     "track_slop", "crash_coeff", "steering_coeff", "throttle_coeff",
                     "boundary_threshold", "discount"):
            setattr(s, name, float(getattr(self, name)))
        s.num_timesteps = int(self.num_timesteps)
        s.grid_res = int(self.grid_res)
        for i in range(3):
            s.r_c1[i] = float(self.r_c1[i])
            s.r_c2[i] = float(self.r_c2[i])
            s.trs[i] = float(self.trs[i])
        s.l1_cost = int(bool(self.l1_cost))
        return s


@dataclass
class Costmap:
    """A costmap in the reference's npz schema (SRC/params/maps/README.md)."""

    x_bounds: np.ndarray
    y_bounds: np.ndarray
    pixels_per_meter: float
    channel0: np.ndarray  # float32 [H*W] row-major, row = y, col = x
    width: int = field(init=False)
    height: int = field(init=False)

    def __post_init__(self):
        # MPPICosts::loadTrackData, PI/costs.cu:206-207 (float arithmetic, truncation to int)
        xb = np.asarray(self.x_bounds, np.float32)
        yb = np.asarray(self.y_bounds, np.float32)
        ppm = np.float32(self.pixels_per_meter)
        self.width = int(np.float32(xb[1] - xb[0]) * ppm)
        self.height = int(np.float32(yb[1] - yb[0]) * ppm)
        self.channel0 = np.ascontiguousarray(self.channel0, np.float32).reshape(-1)
        if self.channel0.size != self.width * self.height:
            raise ValueError("channel0 has %d texels, expected %d x %d" % (self.channel0.size, self.width, self.height))

    def transform(self):
        """World -> normalised texture coordinates: R and trs of PI/costs.cu:225-229 (float32)."""
        xb = np.asarray(self.x_bounds, np.float32)
        yb = np.asarray(self.y_bounds, np.float32)
        one = np.float32(1.0)
        r_c1 = (float(one / (xb[1] - xb[0])), 0.0, 0.0)
        r_c2 = (0.0, float(one / (yb[1] - yb[0])), 0.0)
        trs = (float(-xb[0] / (xb[1] - xb[0])), float(-yb[0] / (yb[1] - yb[0])), 1.0)
        return r_c1, r_c2, trs

    def to_npz_dict(self):
        zeros = np.zeros_like(self.channel0)
        return {"xBounds": np.asarray(self.x_bounds, np.float32), "yBounds": np.asarray(self.y_bounds, np.float32),
                "pixelsPerMeter": np.asarray(self.pixels_per_meter, np.float32),
                "channel0": self.channel0, "channel1": zeros, "channel2": zeros, "channel3": zeros}

    @staticmethod
    def from_npz(path):
        z = np.load(path)
        return Costmap(z["xBounds"], z["yBounds"], float(np.asarray(z["pixelsPerMeter"]).reshape(-1)[0]), z["channel0"])


def make_ellipse_costmap(a=20.0, b=12.0, half_width=2.0, x_bounds=(-30.0, 30.0), y_bounds=(-20.0, 20.0),
                         pixels_per_meter=20.0) -> Costmap:
    """The synthetic ellipse-track costmap of SURVEY.md section 8(d) (the reference ships no maps).

    channel0 = |r - R(theta)| / half_width at texel centres: 0 on the centreline, 1.0 on the boundary.
    """
    w = int(np.float32(np.float32(x_bounds[1]) - np.float32(x_bounds[0])) * np.float32(pixels_per_meter))
    h = int(np.float32(np.float32(y_bounds[1]) - np.float32(y_bounds[0])) * np.float32(pixels_per_meter))
    x = x_bounds[0] + (np.arange(w, dtype=np.float64) + 0.5) / pixels_per_meter
    y = y_bounds[0] + (np.arange(h, dtype=np.float64) + 0.5) / pixels_per_meter
    xx, yy = np.meshgrid(x, y)  # [h, w]
    r = np.hypot(xx, yy)
    th = np.arctan2(yy, xx)
    big_r = a * b / np.sqrt((b * np.cos(th)) ** 2 + (a * np.sin(th)) ** 2)
    ch0 = (np.abs(r - big_r) / half_width).astype(np.float32)
    return Costmap(np.asarray(x_bounds, np.float32), np.asarray(y_bounds, np.float32), pixels_per_meter, ch0)


def ellipse_start_state(a=20.0, speed=0.0):
    """On-centreline pose, heading tangent: (x, y, yaw) = (a, 0, pi/2)."""
    return np.array([a, 0.0, np.pi / 2.0, 0.0, speed, 0.0, 0.0], np.float32)


def ellipse_states(n, a=20.0, b=12.0, speed=6.0, seed=7):
    """n seeded poses along the ellipse centreline, heading tangent (counter-clockwise)."""
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0.0, 2.0 * np.pi, n)
    x, y = a * np.cos(ang), b * np.sin(ang)
    yaw = np.arctan2(b * np.cos(ang), -a * np.sin(ang))
    st = np.zeros((n, STATE_DIM), np.float32)
    st[:, 0], st[:, 1], st[:, 2], st[:, 4] = x, y, yaw, speed
    return st


def pack_nn_params(weights, biases):
    """[W1|b1|W2|b2|...] row-major float32 + net structure (PI/neural_net_model.cu:120-141)."""
    parts, structure = [], [int(weights[0].shape[1])]
    for w, b in zip(weights, biases):
        w = np.asarray(w)
        b = np.asarray(b).reshape(-1)
        if w.shape[1] != structure[-1] or w.shape[0] != b.size:
            raise ValueError("inconsistent layer shapes")
        parts.append(w.astype(np.float32).reshape(-1))
        parts.append(b.astype(np.float32))
        structure.append(int(w.shape[0]))
    return np.concatenate(parts), np.asarray(structure, np.int32)


def load_nn_npz(path):
    """NeuralNetModel::loadParams (PI/neural_net_model.cu:73-106): float64 npz -> float32 theta."""
    z = np.load(path)
    n = sum(1 for k in z.files if k.startswith("dynamics_W"))
    ws = [z["dynamics_W%d" % i] for i in range(1, n + 1)]
    bs = [z["dynamics_b%d" % i] for i in range(1, n + 1)]
    return pack_nn_params(ws, bs)


def unpack_nn_params(theta, structure):
    ws, bs, off = [], [], 0
    for i in range(len(structure) - 1):
        nin, nout = int(structure[i]), int(structure[i + 1])
        ws.append(np.asarray(theta[off:off + nin * nout]).reshape(nout, nin))
        off += nin * nout
        bs.append(np.asarray(theta[off:off + nout]))
        off += nout
    return ws, bs


def load_bf_npz(path):
    """GeneralizedLinear::loadParams (PI/generalized_linear.cu:92-108): key 'W', 4x25, float64 -> float32."""
    z = np.load(path)
    return np.ascontiguousarray(z["W"], np.float64).astype(np.float32).reshape(4, 25)


def pure_noise_threshold(n_global: int) -> int:
    """Smallest r with (double)r >= .99 * N (PI/mppi_controller.cu:141)."""
    lim = 0.99 * float(n_global)
    r = int(np.floor(lim))
    while float(r) < lim:
        r += 1
    return r


# launch/path_integral_nn.launch:36-48 and SRC/path_integral_main.cu:98
NN_DEFAULTS = dict(hz=50, num_timesteps=100, optimization_stride=1, gamma=0.15, num_iters=1,
                   init_u=(0.0, 0.0), exploration_std=(0.275, 0.3),
                   control_ranges=((-0.99, 0.99), (-0.99, 0.65)), num_rollouts=1920, bdim=(8, 16))
# launch/path_integral_bf.launch and SRC/path_integral_main.cu:71-74
BF_DEFAULTS = dict(hz=50, num_timesteps=100, optimization_stride=1, gamma=0.15, num_iters=1,
                   init_u=(0.0, -0.01), exploration_std=(0.275, 0.3),
                   control_ranges=((-0.99, 0.99), (-0.99, 0.65)), num_rollouts=2560, bdim=(16, 4),
                   desired_speed=6.0)
