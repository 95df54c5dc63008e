"""GPU: the caller of the hot path -- runControlLoop with two controllers sharing model and costs, in debug mode (the
host dynamics model is the plant), as the reference's main() sets it up (SRC/path_integral_main.cu:80-153,
PI/run_control_loop.cuh:84-321).  Closed-loop validation of the whole drop-in stack: launch file -> params ->
MPPICosts / model loaders -> MPPIController (Philox sampler, CUDA graph) -> slide / arbitration -> plant.
The car must drive counter-clockwise around the synthetic ellipse track without leaving the drivable band."""
import os
import subprocess

import numpy as np
import pytest

from tests.test_host_cpp import LIB, write_inputs

pytestmark = pytest.mark.gpu


def run_loop(tmp_path, models, costmap, kind, iterations, pose, double_step=0, swap=None):
    exe = os.path.join(LIB, "control_loop_driver")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    launch, env = write_inputs(tmp_path, models, costmap, kind)
    out = tmp_path / "loop.npz"
    extra = []
    if swap is not None:  # (theta, structure, iteration): hot swap through the flattened update-model message
        from autorally_b200.model_io import flatten_for_update_model
        description, data = flatten_for_update_model(swap[0], swap[1])
        np.savez(tmp_path / "swap.npz", description=description.astype(np.int32), data=data.astype(np.float32))
        extra = [str(tmp_path / "swap.npz"), str(swap[2])]
    subprocess.check_call([exe, kind, launch, str(out), str(iterations)] + [repr(float(v)) for v in pose] + [str(double_step)] + extra, env=env)
    return np.load(out)


def track_value(costmap, x, y):
    """Costmap channel 0 at world (x, y): 0 on the centreline, 1 at the track edge (boundary_threshold 0.65)."""
    u = (x - costmap.x_bounds[0]) / (costmap.x_bounds[1] - costmap.x_bounds[0])
    v = (y - costmap.y_bounds[0]) / (costmap.y_bounds[1] - costmap.y_bounds[0])
    col = np.clip((u * costmap.width).astype(int), 0, costmap.width - 1)
    row = np.clip((v * costmap.height).astype(int), 0, costmap.height - 1)
    return costmap.channel0.reshape(costmap.height, costmap.width)[row, col]


@pytest.mark.parametrize("kind", ["nn", "bf"])
def test_car_laps_the_ellipse_in_closed_loop(tmp_path, models, costmap, kind):
    n = 600  # 12 s at 50 Hz
    got = run_loop(tmp_path, models, costmap, kind, n, (0.0, 12.0, np.pi))
    s, u = got["states"], got["controls"]
    assert s.shape == (n, 7) and u.shape == (n, 2) and np.all(np.isfinite(s)) and np.all(np.isfinite(u))
    # stays inside the drivable band for the whole run
    tv = track_value(costmap, s[:, 0], s[:, 1])
    assert tv.max() < 0.65, "left the track: max costmap value %.3f at iteration %d" % (tv.max(), tv.argmax())
    # accelerates from rest and keeps moving forward, counter-clockwise
    assert s[0, 4] == 0.0 and s[100:, 4].min() > 1.0 and s[:, 4].max() < 12.0
    ang = np.unwrap(np.arctan2(s[:, 1] / 12.0, s[:, 0] / 20.0))
    assert ang[-1] - ang[0] > 1.0 and np.all(np.diff(ang[50:]) > -1e-3)
    # controls respect the constraints of main() (steering +-0.99, throttle [-0.99, max_throttle])
    assert np.abs(u[:, 0]).max() <= 0.99 + 1e-6 and u[:, 1].max() <= 0.65 + 1e-6 and u[:, 1].min() >= -0.99 - 1e-6
    # both controllers take part in the arbitration
    used = got["controller_used"]
    assert set(np.unique(used)) <= {0.0, 1.0} and len(used) == n
    print("%s: %d iterations, mean speed %.2f m/s, %.2f rad around the track, avg tick %.3f ms (2 x computeControl), "
          "actual-state controller used %.0f%%" % (kind, n, s[100:, 4].mean(), ang[-1] - ang[0], float(got["avg_tick_ms"][0]),
                                                   100 * (used == 0).mean()))
    assert float(got["avg_tick_ms"][0]) < 20.0  # the reference's budget: one 50 Hz period for both controllers


def test_reference_debug_double_step_quirk_is_optional(tmp_path, models, costmap):
    """With the reference's debug-plant quirk (state advanced through both controllers' shared model) the simulated car
    covers about twice the distance per iteration."""
    a = run_loop(tmp_path, models, costmap, "nn", 150, (0.0, 12.0, np.pi), double_step=0)["states"]
    b = run_loop(tmp_path, models, costmap, "nn", 150, (0.0, 12.0, np.pi), double_step=1)["states"]
    da = np.abs(np.diff(a[:, 0])).sum() + np.abs(np.diff(a[:, 1])).sum()
    db = np.abs(np.diff(b[:, 0])).sum() + np.abs(np.diff(b[:, 1])).sum()
    assert db > 1.5 * da


def test_model_hot_swap_through_the_update_model_message(tmp_path, models, costmap):
    """updateModel(description, data) (PI/neural_net_model.cu:152-180; all weights then all biases) takes effect on the
    next computeControl: swapping in the SAME weights changes nothing, swapping in another model changes the drive from
    the swap iteration on and only from there (the Philox stream is deterministic)."""
    pose, n, k = (0.0, 12.0, np.pi), 160, 80
    base = run_loop(tmp_path, models, costmap, "nn", n, pose)["states"]
    same = run_loop(tmp_path, models, costmap, "nn", n, pose, swap=(models["autorally_nnet_theta"], models["autorally_nnet_structure"], k))["states"]
    np.testing.assert_array_equal(same, base)
    other = run_loop(tmp_path, models, costmap, "nn", n, pose, swap=(models["gazebo_nnet_theta"], models["gazebo_nnet_structure"], k))["states"]
    # the swap is applied before computeControl of iteration k: rows < k are untouched; row k is still the pre-swap state, but
    # the arbitration of iteration k (run on the NEW model, on the device) may hand over the other controller's copy of it
    # (measured state vs predicted state: equal to float rounding)
    np.testing.assert_array_equal(other[:k], base[:k])
    np.testing.assert_allclose(other[k], base[k], rtol=1e-5, atol=1e-6)
    assert np.abs(other[k + 5:] - base[k + 5:]).max() > 1e-3
