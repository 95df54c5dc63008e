"""CPU: host-side logic and the C-ABI library surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from autorally_b200 import capi, params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.load_library()
    header = open(os.path.join(ROOT, "include", "mppi_b200.h")).read()
    declared = set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", header))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.mppi_version()


def test_error_strings_and_default_config():
    lib = capi.load_library()
    assert lib.mppi_error_string(-4).decode().startswith("no CUDA device")
    cfg = capi.MppiConfig()
    lib.mppi_config_default(ctypes.byref(cfg))
    # launch/path_integral_nn.launch + SRC/path_integral_main.cu:66-69
    assert (cfg.num_rollouts, cfg.num_timesteps, cfg.hz, cfg.optimization_stride, cfg.num_iters) == (1920, 100, 50, 1, 1)
    assert abs(cfg.gamma - 0.15) < 1e-7 and (cfg.bdim_x, cfg.bdim_y) == (8, 16) and cfg.seed == 1234


def test_create_validates_arguments_before_touching_the_device():
    lib = capi.load_library()
    cfg = capi.MppiConfig()
    lib.mppi_config_default(ctypes.byref(cfg))
    ctx = ctypes.c_void_p()
    cfg.num_rollouts = 1000  # not a multiple of 64 (NUM_ROLLOUTS, PI/mppi_controller.cuh:58-60)
    assert lib.mppi_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1
    cfg.num_rollouts = 1920
    cfg.rollout_begin, cfg.rollout_count = 64, 100
    assert lib.mppi_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1
    assert lib.mppi_create(None, ctypes.byref(ctx)) == -1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.MppiError) as e:
        capi.MppiContext()
    assert e.value.code == capi.MPPI_ERR_NO_DEVICE


def test_pure_noise_threshold_matches_the_double_compare():
    # r >= .99 * N in double (PI/mppi_controller.cu:141); values from SURVEY.md section 8a R2
    assert params.pure_noise_threshold(1920) == 1901
    assert params.pure_noise_threshold(2560) == 2535
    assert params.pure_noise_threshold(256) == 254
    assert params.pure_noise_threshold(1000000) == 990000


def test_costmap_schema_and_transform(costmap):
    assert (costmap.width, costmap.height) == (1200, 800)
    d = costmap.to_npz_dict()
    assert set(d) == {"xBounds", "yBounds", "pixelsPerMeter", "channel0", "channel1", "channel2", "channel3"}
    assert all(v.dtype == np.float32 for v in d.values())
    r_c1, r_c2, trs = costmap.transform()
    np.testing.assert_allclose([r_c1[0], r_c2[1]], [1 / 60.0, 1 / 40.0], rtol=1e-6)
    np.testing.assert_allclose(trs, [0.5, 0.5, 1.0], rtol=1e-6)
    # centreline texel ~0, boundary (half width 2 m) ~1
    ch = costmap.channel0.reshape(800, 1200)
    col = int((20.0 + 30.0) * 20)
    assert ch[400, col] < 0.03 and abs(ch[400, col + 40] - 1.0) < 0.03


def test_costmap_npz_round_trip(tmp_path, small_costmap):
    p = tmp_path / "map.npz"
    np.savez(p, **small_costmap.to_npz_dict())
    back = params.Costmap.from_npz(str(p))
    assert (back.width, back.height) == (small_costmap.width, small_costmap.height)
    np.testing.assert_array_equal(back.channel0, small_costmap.channel0)


def test_nn_param_packing_round_trip(models):
    theta, st = models["autorally_nnet_theta"], models["autorally_nnet_structure"]
    ws, bs = params.unpack_nn_params(theta, st)
    assert [w.shape for w in ws] == [(32, 6), (32, 32), (4, 32)]
    theta2, st2 = params.pack_nn_params(ws, bs)
    np.testing.assert_array_equal(theta, theta2)
    np.testing.assert_array_equal(st, st2)


def test_cost_param_struct_layout_matches_the_header():
    # 11 floats + 2 ints + 9 floats + 1 int
    assert ctypes.sizeof(params.CostParamsStruct) == 4 * (11 + 2 + 9 + 1)
    s = params.CostParams().to_struct()
    assert abs(s.desired_speed - 8.0) < 1e-7 and abs(s.boundary_threshold - 0.65) < 1e-7 and s.l1_cost == 0


def test_sampler_fast_division_constants_are_exact():
    """weighting.cuh FastDiv: with s = ceil(log2 d) and M = ceil(2^(31+s) / d), (n * M) >> (31 + s) == n // d for every
    n < 2^31 (the sampler divides its flat float4 index by T/2 and by the rollout count this way)."""
    import random

    def make(d):
        if d <= 1:
            return 0, 0
        s = 0
        while (1 << s) < d:
            s += 1
        return ((1 << (31 + s)) + d - 1) // d, s - 1

    rng = random.Random(5)
    divisors = list(range(1, 130)) + [50, 1000, 1920, 2560, 4096, 65536, 524288, 1048576, 999983, (1 << 20) + 1, 1 << 30, (1 << 31) - 1]
    for d in divisors:
        mul, shift = make(d)
        assert mul < 1 << 32
        probes = [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1, (1 << 31) - 2] + [rng.randrange(1 << 31) for _ in range(500)]
        for n in probes:
            if 0 <= n < 1 << 31:
                q = (((n * mul) >> 32) >> shift) if mul else n
                assert q == n // d, (d, n)
