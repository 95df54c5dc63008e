"""CPU: pins the oracle against fixtures produced by the reference's own Python model
(tests/golden/make_golden.py imports ml_pipeline/utils.py unmodified) and against the published
Philox4x32-10 known-answer vectors; plus self-consistency properties of the restated pipeline."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.common import cost_params_for, default_state, make_oracle, warm_controls


@pytest.mark.parametrize("tag", ["autorally_nnet", "wider_deeper"])
def test_single_step_derivatives_match_reference_model(models, ref_dynamics, tag):
    g = ref_dynamics
    o = orc.Oracle("nn", models[tag + "_theta"], models[tag + "_structure"], negate_yaw_der=bool(g[tag + "_negate_yaw_der"]))
    st, ct, de = g[tag + "_step_states"], g[tag + "_step_ctrls"], g[tag + "_step_ders"]
    for i in range(len(st)):
        _, sd = o.dynamics_step(st[i], ct[i])
        np.testing.assert_allclose(sd, de[i], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("tag,tol", [("autorally_nnet", 1e-4), ("wider_deeper", 5e-4)])
def test_100_step_rollouts_match_reference_model(models, ref_dynamics, tag, tol):
    """BASELINE config 1: the ml_pipeline model rolled out 100 steps at dt = 0.02 (float64) vs the
    float32 oracle; tolerance is relative to 1 + |x| after 100 recurrent steps."""
    g = ref_dynamics
    o = orc.Oracle("nn", models[tag + "_theta"], models[tag + "_structure"], negate_yaw_der=bool(g[tag + "_negate_yaw_der"]))
    fs = o.dynamics_rollouts(g[tag + "_roll_state0"], g[tag + "_roll_U"], [0.275, 0.3], g[tag + "_roll_eps"])
    ref = g[tag + "_roll_traj"][:, -1]
    assert np.max(np.abs(fs - ref) / (1 + np.abs(ref))) < tol


def test_param_packing_offsets(models):
    # W1=0, b1=192, W2=224, b2=1248, W3=1280, b3=1408, end=1412 (PI/neural_net_model.cu:125-141)
    assert models["autorally_nnet_theta"].size == 1412
    assert list(models["autorally_nnet_structure"]) == [6, 32, 32, 4]
    assert models["wider_deeper_theta"].size == 7 * 64 + 3 * 65 * 64 + 65 * 4
    assert models["basis_function_W"].shape == (4, 25)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        got = orc.philox4x32_10(ctr, key)
        assert tuple(int(x) for x in got) == want


def test_sampler_oracle_statistics():
    eps = orc.sample_noise(1234, 0, 0, 2048, 100)
    assert abs(eps.mean()) < 5e-3 and abs(eps.std() - 1.0) < 5e-3
    # stream definition: a shard reproduces the same rows of the global stream
    part = orc.sample_noise(1234, 0, 512, 64, 100)
    np.testing.assert_array_equal(part, eps[512:576])
    assert not np.array_equal(orc.sample_noise(1234, 1, 0, 64, 100), eps[:64])


def test_bookkeeping_and_unclamped_writeback(models, small_costmap):
    cp = cost_params_for(small_costmap)
    o = make_oracle("nn", models, small_costmap, cp)
    N, T = 128, 20
    rng = np.random.default_rng(0)
    eps = (3.0 * rng.standard_normal((N, T, 2))).astype(np.float32)  # large noise -> clamping happens
    U = warm_controls(T)
    nu = np.array([0.275, 0.3], np.float32)
    V, costs, crash, _ = o.rollouts(default_state(), U, nu, eps, opt_delay=1)
    thr = 127  # smallest r with r >= .99*128 = 126.72
    np.testing.assert_array_equal(V[0], U)                                  # rollout 0 is noise free
    np.testing.assert_array_equal(V[:, 0], np.broadcast_to(U[0], (N, 2)))   # t < opt_delay untouched
    np.testing.assert_array_equal(V[1:thr, 1:], U[None, 1:] + eps[1:thr, 1:] * nu)   # un-clamped U + eps*nu
    np.testing.assert_array_equal(V[thr:, 1:], eps[thr:, 1:] * nu)          # pure-noise tail
    assert np.abs(V).max() > 0.99                                            # write-back is NOT clamped
    assert np.all(np.isfinite(costs)) and costs.min() >= 0


def test_running_cost_is_mean_of_steps_1_to_Tm1(models, small_costmap):
    """With zero noise every rollout equals the nominal one; a constant per-step cost c gives running mean c."""
    cp = cost_params_for(small_costmap, track_coeff=0.0, slip_penalty=0.0, crash_coeff=0.0, speed_coeff=0.0,
                         steering_coeff=0.0, throttle_coeff=0.0)
    o = make_oracle("nn", models, small_costmap, cp)
    V, costs, crash, _ = o.rollouts(default_state(), warm_controls(10), [0.275, 0.3], np.zeros((64, 10, 2), np.float32))
    np.testing.assert_array_equal(costs, 0.0)


def test_weighting_matches_float64_definition():
    rng = np.random.default_rng(3)
    N, T, gamma = 256, 12, 0.15
    costs = (50 + 30 * rng.random(N)).astype(np.float32)
    V = rng.standard_normal((N, T, 2)).astype(np.float32)
    w, Unew, stats = orc.Oracle.weighting(costs, V, gamma)
    w64 = np.exp(-gamma * (costs.astype(np.float64) - costs.min()))
    np.testing.assert_allclose(w, w64, rtol=2e-6)
    np.testing.assert_allclose(stats[1], w64.sum(), rtol=1e-5)
    np.testing.assert_allclose(stats[2], (w64 ** 2).sum() / w64.sum(), rtol=1e-5)
    np.testing.assert_allclose(Unew, np.einsum("r,rtj->tj", w64 / w64.sum(), V), rtol=1e-4, atol=1e-6)
    # shard partials recombine to the same answer
    parts = [orc.Oracle.shard_partials(costs[i:i + 64], V[i:i + 64], gamma) for i in range(0, N, 64)]
    b = min(p[0] for p in parts)
    s = [np.exp(-gamma * (p[0] - b)) for p in parts]
    Z = sum(si * p[1] for si, p in zip(s, parts))
    W = sum(si * p[3:3 + 2 * T] for si, p in zip(s, parts))
    np.testing.assert_allclose((W / Z).reshape(T, 2), Unew, rtol=1e-4, atol=1e-6)


def test_savitsky_golay_and_slides():
    T = 16
    U = np.arange(2 * T, dtype=np.float32).reshape(T, 2)
    hist = np.array([-4, -3, -2, -1], np.float32)
    sm = orc.Oracle.savitsky_golay(U, hist)
    # a straight line is a fixed point of the quadratic/cubic SG filter away from the padded tail
    np.testing.assert_allclose(sm[: T - 2], U[: T - 2], rtol=1e-5, atol=1e-5)
    U1, h1 = orc.Oracle.slide_control_seq(U, hist, [9, 9], 1)
    np.testing.assert_array_equal(h1, [-2, -1, 0, 1])
    np.testing.assert_array_equal(U1[:-1], U[1:])
    np.testing.assert_array_equal(U1[-1], [9, 9])
    U3, h3 = orc.Oracle.slide_control_seq(U, hist, [9, 9], 3)
    np.testing.assert_array_equal(h3, U.reshape(-1)[1:5])   # reference quirk: flat index t = stride-2
    np.testing.assert_array_equal(U3[:-3], U[3:])
    np.testing.assert_array_equal(U3[-3:], 9)


def test_compute_control_end_to_end_properties(models, small_costmap):
    cp = cost_params_for(small_costmap)
    o = make_oracle("nn", models, small_costmap, cp)
    N, T = 256, 30
    eps = np.random.default_rng(5).standard_normal((1, N, T, 2)).astype(np.float32)
    U = warm_controls(T)
    out = o.compute_control(default_state(), U, np.zeros(4), [0.275, 0.3], eps)
    assert out["normalizer"] >= 1.0 and 0 < out["trajectory_cost"] <= 1.0 + 1e-6
    assert out["baseline"] == out["costs"].min()
    # nominal trajectory starts at the state and its controls are the clamped smoothed controls
    np.testing.assert_array_equal(out["state_solution"][0], default_state())
    np.testing.assert_array_equal(out["control_solution"], np.clip(out["U"], [-0.99, -0.99], [0.99, 0.65]))


def test_bf_oracle_runs_and_is_finite(models, small_costmap):
    cp = cost_params_for(small_costmap, desired_speed=6.0)
    o = make_oracle("bf", models, small_costmap, cp)
    eps = np.random.default_rng(6).standard_normal((128, 40, 2)).astype(np.float32)
    V, costs, crash, fs = o.rollouts(default_state(), warm_controls(40), [0.275, 0.3], eps)
    assert np.all(np.isfinite(costs)) and np.all(np.isfinite(fs))


# ---- pinned against the reference's own kernels (tests/golden/make_ref_gpu_golden.py, run on a B200) --------------
def _ref_gpu_cases():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_gpu_golden.npz")
    if not os.path.exists(path):
        return None, []
    z = np.load(path)
    return z, sorted({k.split("/")[0] for k in z.files})


_REF_GPU, _REF_GPU_CASES = _ref_gpu_cases()


@pytest.mark.skipif(not _REF_GPU_CASES, reason="tests/golden/ref_gpu_golden.npz not generated yet")
@pytest.mark.parametrize("case", _REF_GPU_CASES)
def test_oracle_matches_reference_gpu_outputs(models, costmap, case):
    """The CPU oracle against outputs of the REFERENCE's own MPPIController (rolloutKernel, normExpKernel,
    weightedReductionKernel, device dynamics and costs, host smoothing and nominal rollout) recorded on a B200 with the
    noise its cuRAND generator drew.  This is what makes the oracle "pinned": bookkeeping bit-exact, costs within
    1e-4 relative for >= 99% of the rollouts (the rest are discrete threshold flips), controls within 1e-4."""
    from tests.common import cost_params_for, make_oracle
    z = _REF_GPU
    g = lambda k: z[case + "/" + k]
    T, opt_delay, is_bf = (int(v) for v in g("cfg"))
    sc, tc, slop, l1, slip, vdes = (float(v) for v in g("cost_over"))
    cp = cost_params_for(costmap, steering_coeff=sc, throttle_coeff=tc, track_slop=slop, l1_cost=bool(l1), max_slip_ang=slip,
                         desired_speed=vdes)
    o = make_oracle("bf" if is_bf else "nn", models, costmap, cp)
    got = o.compute_control(g("state"), g("U_in"), g("hist"), [0.275, 0.3], g("eps"), opt_delay=opt_delay, threads=4)
    np.testing.assert_array_equal(got["V"][0], g("V_row0"))    # noise-free rollout 0 (PI/mppi_controller.cu:136-140)
    np.testing.assert_array_equal(got["V"][-1], g("V_last"))   # pure-noise tail (:141-145)
    want_c = g("costs")
    err = np.abs(got["costs"] - want_c) / (1 + np.abs(want_c))
    tol = 2e-4 if is_bf else 1e-4
    assert (err < tol).mean() >= 0.99, "only %.2f%% of costs within %g" % (100 * (err < tol).mean(), tol)
    assert abs(got["costs"].min() - want_c.min()) <= tol * (1 + abs(want_c.min()))
    np.testing.assert_allclose(got["w"], g("w"), rtol=5e-3, atol=1e-6)
    normalizer, traj_cost = g("scalars")
    assert abs(got["normalizer"] - normalizer) <= 1e-3 * normalizer
    assert abs(got["trajectory_cost"] - traj_cost) <= 1e-3 * traj_cost
    for k in ("U", "state_solution", "control_solution"):
        e = np.abs(got[k] - g(k)) / (1 + np.abs(g(k)))
        assert e.max() < 1e-4, "%s: max rel err %.3g" % (k, e.max())
