"""GPU: three-way parity -- the REFERENCE's own MPPIController (oracle/_ref/libautorally_ref.so, built from
/root/reference by oracle/refbuild.py and run here on the B200), the CUDA product path (through the C ABI)
and the CPU oracle, all on the SAME noise: the N(0,1) draws the reference's cuRAND generator produced.

This is what pins the oracle: rdesc/autorally ships no tests or golden vectors for this path, so its own
kernels, compiled unmodified and executed on the GPU box, are the authority.  Tolerances as in
tests/test_parity_gpu.py: bookkeeping bit-exact; costs / controls within 1e-4 relative, with <= 1% of the
rollouts allowed to differ by discrete events (crash / slip-kill flips at a threshold).
"""
import numpy as np
import pytest

from oracle import reference as ref
from tests.common import (cost_params_for, default_state, make_context, make_oracle, random_network, straight_controls, top_state,
                          warm_controls)
from tests.test_parity_gpu import TRUE_REL_BOUND, check_costs, rel_err, true_rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libautorally_ref.so not built")]

NU = np.array([0.275, 0.3], np.float32)
HIST = np.array([0.1, 0.3, 0.11, 0.32], np.float32)


def reference_run(kind, theta, costmap, cp, state, U, T=100, init_u=(0.0, 0.0), opt_delay=1, iters=1, gamma=0.15):
    with ref.ReferenceController(kind, theta, costmap, cp, num_timesteps=T, init_u=init_u, optimization_stride=opt_delay,
                                 num_iters=iters, gamma=gamma) as rc:
        rc.set_controls(U, HIST)
        out = rc.compute_control(state)
        out["costs"], out["V"] = rc.rollout_costs(state, U, out["eps"][0])
    return out


def compare(got, want, T, what, cost_tol=1e-4, u_tol=1e-4):
    np.testing.assert_array_equal(got["V"], want["V"], err_msg=what + ": sampled-control bookkeeping")
    check_costs(got["costs"], want["costs"], T, cost_tol)
    assert abs(got["costs"].min() - want["costs"].min()) <= cost_tol * (1 + abs(want["costs"].min())), what
    # truly relative (PI/mppi_controller.cu:641-652); widened by 4 gamma ulp(baseline): see tests/test_parity_gpu.py:check_scalars
    wide = cost_tol + 4.0 * 0.5 * float(np.spacing(np.float32(abs(want["costs"].min()))))
    for k in ("normalizer", "trajectory_cost"):
        g, w = float(got[k]), float(want[k])
        assert abs(g - w) <= wide * abs(w), "%s: %s got %.9g want %.9g (bound %.3g)" % (what, k, g, w, wide)
    e = rel_err(got["U"], want["U"]).max()
    assert e < u_tol, "%s: max control rel err %.3g" % (what, e)
    tr = true_rel_err(got["U"], want["U"])   # reported; see tests/test_parity_gpu.py:TRUE_REL_BOUND for why it is not held to 1e-4
    print("%s: max |dU| / (1 + |U|) = %.3g, true relative over |u| > 1e-2 = %.3g" % (what, e, tr))
    assert tr < TRUE_REL_BOUND, "%s: true relative control error %.3g over |u| > 1e-2" % (what, tr)
    assert rel_err(got["state_solution"], want["state_solution"]).max() < 1e-4, what
    assert rel_err(got["control_solution"], want["control_solution"]).max() < u_tol, what


def cuda_run(kind, models, costmap, cp, N, state, U, eps, T=100, opt_delay=1, iters=1, gamma=0.15, variant=0):
    with make_context(kind, models, costmap, cp, N, num_timesteps=T, optimization_stride=opt_delay, num_iters=iters,
                      gamma=gamma, variant=variant) as ctx:
        ctx.set_noise(eps)
        got = ctx.compute_control(state, U, HIST)
        if iters == 1:
            got["costs"], got["V"] = ctx.rollout_costs(), ctx.sampled_controls()
    return got


@pytest.mark.parametrize("speed", [0.0, 5.0, 8.0])
def test_nn_1920x100_three_way(models, costmap, speed):
    """BASELINE config 2 against MPPIController<NeuralNetModel<7,2,3,6,32,32,4>, MPPICosts, 1920, 8, 16>."""
    cp = cost_params_for(costmap)
    state, U = default_state(speed), warm_controls(100)
    want = reference_run(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp, state, U)
    # the twin cuRAND generator really produced the reference's draws: V = U + eps * nu for ordinary rollouts
    r = 7
    np.testing.assert_array_equal(want["V"][r, 1:], (U[1:] + want["eps"][0, r, 1:] * NU).astype(np.float32))
    np.testing.assert_allclose(want["w"], np.exp(-0.15 * (want["costs"] - want["costs"].min())), rtol=2e-5, atol=1e-30)
    got = cuda_run("nn", models, costmap, cp, 1920, state, U, want["eps"])
    compare(got, want, 100, "cuda vs reference")
    o = make_oracle("nn", models, costmap, cp).compute_control(state, U, HIST, NU, want["eps"], threads=8)
    compare(o, want, 100, "oracle vs reference")


@pytest.mark.parametrize("gamma", [0.15, 0.01])
def test_nn_spread_weights_three_way(models, costmap, gamma):
    """Flat top of the ellipse at 4 m/s: weights spread over many rollouts (see tests/common.py:top_state)."""
    cp = cost_params_for(costmap)
    state, U = top_state(4.0), straight_controls(100)
    want = reference_run(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp, state, U, gamma=gamma)
    assert want["normalizer"] > (15 if gamma > 0.1 else 150)
    for variant in (2, 9, 10):
        got = cuda_run("nn", models, costmap, cp, 1920, state, U, want["eps"], gamma=gamma, variant=variant)
        compare(got, want, 100, "cuda (variant %d) vs reference" % variant)
    o = make_oracle("nn", models, costmap, cp).compute_control(state, U, HIST, NU, want["eps"], gamma=gamma, threads=8)
    compare(o, want, 100, "oracle vs reference")


def test_bf_spread_weights_three_way(models, costmap):
    cp = cost_params_for(costmap, desired_speed=6.0)
    state, U = top_state(4.0), straight_controls(100)
    want = reference_run(ref.REF_BF_2560, models["basis_function_W"], costmap, cp, state, U, init_u=(0.0, -0.01))
    got = cuda_run("bf", models, costmap, cp, 2560, state, U, want["eps"])
    compare(got, want, 100, "cuda vs reference", cost_tol=2e-4)
    o = make_oracle("bf", models, costmap, cp).compute_control(state, U, HIST, NU, want["eps"], threads=8)
    compare(o, want, 100, "oracle vs reference", cost_tol=2e-4)


def test_bf_2560x100_three_way(models, costmap):
    """BASELINE config 3 against MPPIController<GeneralizedLinear<CarBasisFuncs,7,2,25,CarKinematics,3>, MPPICosts, 2560, 16, 4>.
    The reference sums the 25 basis-function products with shared-memory atomicAdd (PI/generalized_linear.cu:242-244),
    so its own result is order-dependent at the 1e-7 level; tolerance 2e-4 as in test_parity_gpu."""
    cp = cost_params_for(costmap, desired_speed=6.0)
    state, U = default_state(5.0), warm_controls(100)
    want = reference_run(ref.REF_BF_2560, models["basis_function_W"], costmap, cp, state, U, init_u=(0.0, -0.01))
    got = cuda_run("bf", models, costmap, cp, 2560, state, U, want["eps"])
    compare(got, want, 100, "cuda vs reference", cost_tol=2e-4)
    o = make_oracle("bf", models, costmap, cp).compute_control(state, U, HIST, NU, want["eps"], threads=8)
    compare(o, want, 100, "oracle vs reference", cost_tol=2e-4)


def test_cost_terms_and_opt_delay_three_way(models, costmap):
    over = dict(steering_coeff=0.4, throttle_coeff=0.2, track_slop=0.05, l1_cost=True, max_slip_ang=0.4)
    cp = cost_params_for(costmap, **over)
    state, U = default_state(6.0), warm_controls(60)
    want = reference_run(ref.REF_NN_256, models["autorally_nnet_theta"], costmap, cp, state, U, T=60, opt_delay=3)
    got = cuda_run("nn", models, costmap, cp, 256, state, U, want["eps"], T=60, opt_delay=3)
    compare(got, want, 60, "cuda vs reference")
    o = make_oracle("nn", models, costmap, cp).compute_control(state, U, HIST, NU, want["eps"], opt_delay=3, threads=8)
    compare(o, want, 60, "oracle vs reference")


def test_multiple_iterations_three_way(models, costmap):
    cp = cost_params_for(costmap)
    state, U = default_state(5.0), warm_controls(100)
    want = reference_run(ref.REF_NN_4096, models["autorally_nnet_theta"], costmap, cp, state, U, iters=2)
    got = cuda_run("nn", models, costmap, cp, 4096, state, U, want["eps"], iters=2)
    assert rel_err(got["U"], want["U"]).max() < 2e-4
    assert rel_err(got["state_solution"], want["state_solution"]).max() < 2e-4
    o = make_oracle("nn", models, costmap, cp).compute_control(state, U, HIST, NU, want["eps"], threads=8)
    assert rel_err(o["U"], want["U"]).max() < 2e-4


def test_closed_loop_with_slides_tracks_the_reference(models, costmap):
    """Five control-loop iterations (slideControlAndStateSeq(1) + computeControl, PI/run_control_loop.cuh:212-219):
    the CUDA path is fed the reference's U / history / noise each iteration and must reproduce its output; the host
    sliding logic is checked against the reference's own slideControlSeq."""
    from oracle.oracle import Oracle
    cp = cost_params_for(costmap)
    state = default_state(5.0)
    o = make_oracle("nn", models, costmap, cp)
    with ref.ReferenceController(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp) as rc, \
            make_context("nn", models, costmap, cp, 1920) as ctx:
        rc.set_controls(warm_controls(100), HIST)
        for it in range(5):
            U_in, hist_in = rc.get_controls()
            want = rc.compute_control(state)
            ctx.set_noise(want["eps"])
            got = ctx.compute_control(state, U_in, hist_in)
            assert rel_err(got["U"], want["U"]).max() < 1e-4, "iteration %d" % it
            assert rel_err(got["state_solution"], want["state_solution"]).max() < 1e-4
            U_sl, hist_sl = Oracle.slide_control_seq(want["U"], hist_in, (0.0, 0.0), 1)
            rc.slide(1)
            U_ref, hist_ref = rc.get_controls()
            np.testing.assert_array_equal(U_sl, U_ref)
            np.testing.assert_array_equal(hist_sl, hist_ref)
            state, _ = o.update_state(state, want["control_solution"][0])  # plant = host model, as in debug mode


def test_reference_latency_is_reported(models, costmap):
    """The reference's own computeControl wall time on this GPU (printed; bench.py reports it beside ours)."""
    cp = cost_params_for(costmap)
    with ref.ReferenceController(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp) as rc:
        rc.set_controls(warm_controls(100), HIST)
        ms = rc.time_compute_control(default_state(5.0), reps=10)
    print("reference computeControl (1920x100, sm_100a build): %.3f ms/call" % ms)
    assert ms > 0


@pytest.mark.parametrize("seed", range(8))
def test_randomized_scenarios_three_way(models, costmap, seed):
    """Random poses along the track (random lateral offset, heading error, speed, slip), random smooth nominal controls,
    random cost parameters and exploration: the reference's kernels, the CUDA path and the oracle on the same draws."""
    rng = np.random.default_rng(1000 + seed)
    ang = rng.uniform(0, 2 * np.pi)
    a, b = 20.0, 12.0
    off = rng.uniform(-0.8, 0.8)
    x, y = (a + off) * np.cos(ang), (b + off) * np.sin(ang)
    yaw = np.arctan2(b * np.cos(ang), -a * np.sin(ang)) + rng.uniform(-0.3, 0.3)
    state = np.array([x, y, yaw, rng.uniform(-0.05, 0.05), rng.uniform(0.0, 9.0), rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5)], np.float32)
    t = np.arange(100)
    U = np.stack([rng.uniform(-0.3, 0.3) + 0.1 * np.sin(t / rng.uniform(5, 30)),
                  rng.uniform(0.0, 0.6) + 0.1 * np.cos(t / rng.uniform(5, 30))], 1).astype(np.float32)
    over = dict(desired_speed=float(rng.uniform(3, 10)), speed_coeff=float(rng.uniform(1, 8)), track_coeff=float(rng.uniform(50, 400)),
                max_slip_ang=float(rng.uniform(0.5, 1.5)), slip_penalty=float(rng.uniform(1, 20)), track_slop=float(rng.choice([0.0, 0.1])),
                steering_coeff=float(rng.choice([0.0, 0.3])), throttle_coeff=float(rng.choice([0.0, 0.2])),
                boundary_threshold=float(rng.uniform(0.5, 0.9)), discount=float(rng.uniform(0.0, 0.3)), l1_cost=bool(rng.integers(2)))
    cp = cost_params_for(costmap, **over)
    gamma = float(rng.choice([0.15, 0.05, 0.5]))
    want = reference_run(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp, state, U, gamma=gamma)
    got = cuda_run("nn", models, costmap, cp, 1920, state, U, want["eps"], gamma=gamma)
    compare(got, want, 100, "cuda vs reference (seed %d)" % seed)
    o = make_oracle("nn", models, costmap, cp).compute_control(state, U, HIST, NU, want["eps"], gamma=gamma, threads=8)
    compare(o, want, 100, "oracle vs reference (seed %d)" % seed)


def test_wider_deeper_network_three_way(models, costmap):
    """The fork's 6-64-64-64-64-4 network (SRC/params/models/wider_deeper_network_08_20_2020.npz) against the reference's own
    MPPIController<NeuralNetModel<7,2,3,6,64,64,64,64,4>, MPPICosts, 1920, 8, 16> compiled from its sources: the
    layer-pipeline kernel (AUTO at this size), the tensor-core kernel rollout_tc_kernel<64,4>, the one-rollout-per-thread FP32
    kernel, and the CPU oracle."""
    cp = cost_params_for(costmap)
    state, U = top_state(4.0), straight_controls(100)
    theta = models["wider_deeper_theta"]
    with ref.ReferenceController(ref.REF_NN64_1920, theta, costmap, cp, negate_yaw_der=False) as rc:
        rc.set_controls(U, HIST)
        want = rc.compute_control(state)
        want["costs"], want["V"] = rc.rollout_costs(state, U, want["eps"][0])
    for variant in (0, 10, 1):
        with make_context("nn", models, costmap, cp, 1920, tag="wider_deeper", negate_yaw_der=False, variant=variant) as ctx:
            ctx.set_noise(want["eps"])
            got = ctx.compute_control(state, U, HIST)
            got["costs"], got["V"] = ctx.rollout_costs(), ctx.sampled_controls()
            assert ctx.resolved_variant() == (12 if variant == 0 else variant)
        compare(got, want, 100, "cuda (variant %d) vs reference, 64-wide network" % variant, cost_tol=3e-4)
    o = make_oracle("nn", models, costmap, cp, tag="wider_deeper", negate_yaw_der=False).compute_control(state, U, HIST, NU, want["eps"], threads=8)
    compare(o, want, 100, "oracle vs reference, 64-wide network", cost_tol=3e-4)


@pytest.mark.parametrize("stride", [1, 3])
def test_slide_with_stride_three_way(models, costmap, stride):
    """slideControlAndStateSeq(stride) (PI/mppi_controller.cu:527-568), including the stride != 1 branch that fills
    control_hist_ from the flat U_ at index stride - 2: the reference's own slide against the oracle's and the drop-in
    C++ template's (tests/cpp/host_cpu_driver.cpp `slide`), then a computeControl from the slid sequences with
    optimization_stride = stride on all three."""
    import os
    import subprocess
    from oracle.oracle import Oracle
    from tests.test_host_cpp import cpu_driver
    cp = cost_params_for(costmap)
    state, init_u = top_state(4.0), (0.05, -0.01)
    o = make_oracle("nn", models, costmap, cp)
    with ref.ReferenceController(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp, init_u=init_u, optimization_stride=stride) as rc, \
            make_context("nn", models, costmap, cp, 1920, optimization_stride=stride) as ctx:
        rc.set_controls(warm_controls(100), HIST)
        first = rc.compute_control(state)
        U0, hist0 = rc.get_controls()
        rc.slide(stride)
        U_ref, hist_ref = rc.get_controls()
        U_sl, hist_sl = Oracle.slide_control_seq(U0, hist0, init_u, stride)
        np.testing.assert_array_equal(U_sl, U_ref)
        np.testing.assert_array_equal(hist_sl, hist_ref)
        # the drop-in MPPIController template's slide (host C++, no GPU involved)
        out = subprocess.check_output([cpu_driver(), "slide", str(stride), repr(float(init_u[0])), repr(float(init_u[1]))] +
                                      [repr(float(v)) for v in hist0] + [repr(float(v)) for v in U0.reshape(-1)], text=True).split()
        tmpl = np.array([float(v) for v in out], np.float32)
        np.testing.assert_array_equal(tmpl[:4], hist_ref)
        np.testing.assert_array_equal(tmpl[4:].reshape(100, 2), U_ref)
        # plan again from the slid sequence (the predicted state = head of the slid state sequence)
        state2 = first["state_solution"][stride]
        want = rc.compute_control(state2)
        want["costs"], want["V"] = rc.rollout_costs(state2, U_ref, want["eps"][0])
        ctx.set_noise(want["eps"])
        got = ctx.compute_control(state2, U_ref, hist_ref)
        got["costs"], got["V"] = ctx.rollout_costs(), ctx.sampled_controls()
        compare(got, want, 100, "cuda vs reference after slide(%d)" % stride)
        oo = o.compute_control(state2, U_ref, hist_ref, NU, want["eps"], opt_delay=stride, threads=8)
        compare(oo, want, 100, "oracle vs reference after slide(%d)" % stride)


@pytest.mark.parametrize("kind,structure", [(ref.REF_NN16_1920, (6, 16, 16, 4)), (ref.REF_NN48_1920, (6, 48, 4))])
def test_arbitrary_layer_packs_three_way(models, costmap, kind, structure):
    """Layer packs without a dedicated kernel -- NeuralNetModel<7,2,3,6,16,16,4> and <7,2,3,6,48,4>, instantiated from the
    reference's variadic template in oracle/ref_harness.cu, random weights -- on the run-time layer kernel
    (rollout_generic.cu, finalize_kernel<0>) and the CPU oracle, against the reference's own kernels."""
    cp = cost_params_for(costmap)
    theta, st = random_network(structure, seed=17)
    state, U = top_state(4.0), straight_controls(100)
    with ref.ReferenceController(kind, theta, costmap, cp) as rc:
        rc.set_controls(U, HIST)
        want = rc.compute_control(state)
        want["costs"], want["V"] = rc.rollout_costs(state, U, want["eps"][0])
    with make_context("nn", models, costmap, cp, 1920, theta=theta, structure=st) as ctx:
        assert ctx.resolved_variant() == 11
        ctx.set_noise(want["eps"])
        got = ctx.compute_control(state, U, HIST)
        got["costs"], got["V"] = ctx.rollout_costs(), ctx.sampled_controls()
    compare(got, want, 100, "cuda (run-time layer kernel) vs reference %s" % (structure,))
    o = make_oracle("nn", models, costmap, cp, theta=theta, structure=st).compute_control(state, U, HIST, NU, want["eps"], threads=8)
    compare(o, want, 100, "oracle vs reference %s" % (structure,))


def test_reference_kernel_times_are_reported(models, costmap):
    """Device time of the reference's own kernels on this GPU (bench.py --impl reference reports the same split)."""
    cp = cost_params_for(costmap)
    with ref.ReferenceController(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp) as rc:
        rc.set_controls(straight_controls(100), HIST)
        k = rc.time_kernels(top_state(4.0), reps=10)
        e2e = rc.time_compute_control(top_state(4.0), reps=10)
        want = rc.compute_control(top_state(4.0))   # the twin generator is still in step with the controller's
        costs, V = rc.rollout_costs(top_state(4.0), straight_controls(100), want["eps"][0])
    print("reference kernels (ms): %s; computeControl %.3f ms" % (k, e2e))
    assert all(v > 0 for v in k.values()) and sum(k.values()) < e2e
    np.testing.assert_array_equal(V[7, 1:], (straight_controls(100)[1:] + want["eps"][0, 7, 1:] * NU).astype(np.float32))


# ---- SURVEY section 8 f-2 / f-1 pinned on the reference's own DDP and runControlLoop (oracle/ref_harness.cu) ----

def _gain_errors(got, want):
    """Per timestep ||K_k - K_k_ref||_F / ||K_k_ref||_F over the steps that carry a gain (the last one is zero by construction)."""
    num = np.sqrt(((got - want) ** 2).sum(axis=(-1, -2)))
    den = np.sqrt((want ** 2).sum(axis=(-1, -2)))
    m = den > 1e-6 * den.max()
    return (num[m] / den[m])


@pytest.mark.parametrize("kind", ["nn", "bf"])
def test_feedback_gains_match_reference_ddp(tmp_path, models, costmap, kind):
    """computeFeedbackGains: the reference's DDP<ModelWrapperDDP<MODEL>>::run with TrackingCostDDP (DDP/ddp.h:49-157,
    DDP/ddp_tracking_costs.h:7-117, analytic NeuralNetModel::computeGrad or Eigen::NumericalDiff for the basis functions), compiled
    from its sources, against include/autorally_control/ddp/ddp_feedback.h through the drop-in MPPIController template.  Both
    controllers plan from the same state on the same noise (the reference's cuRAND draws), then linearise around their own
    solutions."""
    import os
    import subprocess
    from autorally_b200.params import BF_DEFAULTS, NN_DEFAULTS
    from tests.test_host_cpp import LIB, write_inputs
    dflt = NN_DEFAULTS if kind == "nn" else BF_DEFAULTS
    cp = cost_params_for(costmap, desired_speed=dflt.get("desired_speed", 8.0))
    state = default_state(5.0)
    T = 100
    init_u = tuple(float(v) for v in dflt["init_u"])
    U0 = np.broadcast_to(np.asarray(init_u, np.float32), (T, 2)).copy()
    theta = models["autorally_nnet_theta"] if kind == "nn" else models["basis_function_W"]
    with ref.ReferenceController(ref.REF_NN_1920 if kind == "nn" else ref.REF_BF_2560, theta, costmap, cp, init_u=init_u) as rc:
        rc.set_controls(U0, np.zeros(4, np.float32))
        want = rc.compute_control(state)
        g_ref, ff_ref = rc.feedback_gains(state)
    np.concatenate([want["eps"][0].ravel(), want["eps"][0].ravel()]).astype(np.float32).tofile(tmp_path / "noise.bin")
    launch, env = write_inputs(tmp_path, models, costmap, kind)
    subprocess.check_call([os.path.join(LIB, "host_api_driver"), kind, launch, str(tmp_path / "noise.bin"), str(tmp_path / "out.npz")] +
                          [repr(float(v)) for v in state], env=env)
    got = np.load(tmp_path / "out.npz")
    assert rel_err(got["state_solution1"], want["state_solution"]).max() < 2e-4   # the trajectories the gains linearise around agree
    err = _gain_errors(got["feedback_gain"], g_ref)
    print("%s feedback gains vs the reference's DDP: max relative (Frobenius, per step) %.3g, median %.3g, |K| max %.3g" % (
        kind, err.max(), np.median(err), np.abs(g_ref).max()))
    assert np.abs(g_ref[:50]).max() > 1e-3 and np.all(g_ref[-1] == 0)
    # NN: analytic Jacobians on both sides.  BF: both sides difference the model numerically with h = sqrt(eps) |x| in float32
    # (DDP/ddp_dynamics.h:79-83), i.e. steps of ~1e-7 on the small state components; the two Jacobians are equally noisy
    # and differ in their noise.
    assert err.max() < (1e-3 if kind == "nn" else 0.5), err.max()


def test_run_control_loop_tracks_the_reference_loop(tmp_path, models, costmap):
    """The reference's own runControlLoop (PI/run_control_loop.cuh:84-321; debug mode, two controllers sharing model and costs,
    arbitration on getComputedTrajectoryCost :250-285, feedback gains on) against the drop-in loop of
    include/autorally_control/path_integral/run_control_loop.cuh, free-running for 50 iterations from the same pose on the same
    noise.  The reference's two controllers both seed cuRAND with 1234, so they draw the SAME sequence; the drop-in loop is
    given those draws for both of its controllers, and the reference's debug-plant double step."""
    import os
    import subprocess
    from tests.test_host_cpp import LIB, write_inputs
    cp = cost_params_for(costmap)
    n, pose = 50, (0.0, 12.0, float(np.pi))
    with ref.ReferenceController(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp) as rc:
        want = rc.run_control_loop(pose, n, use_feedback_gains=True)
    assert set(np.unique(want["controller_used"])) == {0, 1}, "the reference trace must exercise both branches of the arbitration"
    want["eps"].tofile(tmp_path / "noise.bin")
    launch, env = write_inputs(tmp_path, models, costmap, "nn")
    out = tmp_path / "loop.npz"
    subprocess.check_call([os.path.join(LIB, "control_loop_driver"), "nn", launch, str(out), str(n)] + [repr(float(v)) for v in pose] +
                          ["1", "-", "0", str(tmp_path / "noise.bin"), "1"], env=env)
    got = np.load(out)
    used = got["controller_used"].astype(np.int32)
    tc_w, tc_g = want["trajectory_costs"], got["trajectory_costs"]
    gap = np.abs(tc_w[:, 0] - tc_w[:, 1]) / np.maximum(np.abs(tc_w).max(axis=1), 1e-30)
    print("controller_used reference %s\n                drop-in   %s" % ("".join(map(str, want["controller_used"])), "".join(map(str, used))))
    print("smallest relative gap between the two trajectory costs in the reference run (iterations with a gap): %.3g; max state diff %.3g; "
          "max trajectory-cost rel diff %.3g" % (gap[gap > 0].min() if (gap > 0).any() else 0.0, np.abs(got["states"] - want["states"]).max(),
                                               (np.abs(tc_g - tc_w) / np.abs(tc_w)).max()))
    # A free-running closed loop can only be compared up to the first near-tie of the arbitration: where the two trajectory costs
    # are within 5e-4 of each other the parity of one computeControl no longer fixes the sign of their difference, the two
    # loops may pick different controllers and are different experiments from there on.
    close = np.nonzero(gap < 5e-4)[0]
    m = int(close[0]) if len(close) else n
    print("comparing the first %d iterations (first near-tie of the reference's two trajectory costs at iteration %d, gap %.3g)" % (
        m, m, gap[m] if m < n else float("nan")))
    assert m >= 12 and set(np.unique(want["controller_used"][:m])) == {0, 1}
    # the arbitration rule itself (:250-251: the measured-state plan wins only with the strictly smaller cost), all 50 iterations, both loops
    np.testing.assert_array_equal(want["controller_used"], np.where(tc_w[:, 0] < tc_w[:, 1], 0, 1))
    np.testing.assert_array_equal(used, np.where(tc_g[:, 0] < tc_g[:, 1], 0, 1))
    np.testing.assert_array_equal(used[:m], want["controller_used"][:m])
    # differences of ~1e-6 per call accumulate along the closed loop (each iteration plans from the previous one's output):
    # tight over the first 10 iterations, looser up to the near-tie
    for hi, tol in ((min(m, 10), 1.0), (m, 10.0)):
        np.testing.assert_allclose(got["states"][:hi], want["states"][:hi], rtol=1e-3 * tol, atol=1e-3 * tol)
        np.testing.assert_allclose(got["controls"][:hi], want["controls"][:hi], rtol=0, atol=1e-3 * tol)
        np.testing.assert_allclose(tc_g[:hi], tc_w[:hi], rtol=2e-3 * tol)
        # the sequences each controller holds after the iteration (the predicted-state controller adopts the winner's when the
        # measured-state plan wins, PI/run_control_loop.cuh:259-260)
        np.testing.assert_allclose(got["U_actual"][:hi], want["U_actual"][:hi], rtol=0, atol=2e-3 * tol)
        np.testing.assert_allclose(got["U_predicted"][:hi], want["U_predicted"][:hi], rtol=0, atol=2e-3 * tol)
    # the gains handed to the plant every iteration (the chosen controller's), per step of the horizon
    worst = max(_gain_errors(got["gains"][i], want["gains"][i]).max() for i in range(m))
    print("feedback gains handed over, worst relative (Frobenius, per step) over the first %d iterations: %.3g" % (m, worst))
    assert worst < 5e-2, worst
