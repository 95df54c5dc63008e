#!/usr/bin/env python
"""bench.py -- rollout-steps/s of the MPPI hot path (computeControl) on B200.

    python bench.py --gpus 1 --steps 200 --warmup 10            # our CUDA path
    python bench.py --impl reference --steps 5 --warmup 1       # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete computeControl pipeline (Philox noise -> fused rollouts -> importance
weighting -> Savitzky-Golay -> nominal trajectory) of BASELINE.json configs[1]: path_integral_nn,
1920 rollouts x 100 timesteps on the synthetic ellipse costmap.  At N > 1 every rank runs one such
controller (batched-MPC sharding: independent controllers, no communication) for `value`, and the
1M-rollout configuration sharded over the ranks is reported under "sharded_large", once with the 816-byte exchange
fused into the weighting / finalize kernels over NVLink peer memory and once with one NCCL all-gather per step.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_ROLLOUT_STEP_NN = 2756.0  # 2*(6*32+32*32+32*4) + 68 bias adds
N_ROLLOUTS, T_STEPS = 1920, 100
# rollout_tc.cu: per 128-rollout tile and timestep 2 MMAs 128x32x16 (layer 1), 6 of 128x32x16, 6 of 128x16x16
TENSOR_FLOP_ISSUED_PER_ROLLOUT_STEP = 2.0 * (2 * 32 * 16 + 6 * 32 * 16 + 6 * 16 * 16)
LARGE_ROLLOUTS = 1 << 20           # "large-sample MPPI": 1M rollouts (16384 x 64)


WORKLOAD = "path_integral_nn: NeuralNetModel<7,2,3,6,32,32,4>, 1920 rollouts x 100 steps, synthetic ellipse costmap"


def bench_config(world):
    return {"workload": WORKLOAD, "rollouts": N_ROLLOUTS, "timesteps": T_STEPS,
            "parallelism": "one controller per GPU x%d (independent controllers, no communication)" % world}


def load_setup():
    from autorally_b200.params import make_ellipse_costmap
    from tests.common import cost_params_for, straight_controls, top_state
    models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
    costmap = make_ellipse_costmap()
    cp = cost_params_for(costmap)
    # flat top of the ellipse at 4 m/s: ~30% of the rollouts survive, importance weights spread over many rollouts
    return models, costmap, cp, top_state(4.0), straight_controls(T_STEPS)


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.index, self.period = index, period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(models, costmap, cp, state, U, budget_s=12.0):
    """The CPU restatement (oracle port) on the host's cores: full computeControl, all threads."""
    from tests.common import make_oracle
    o = make_oracle("nn", models, costmap, cp)
    cores = os.cpu_count() or 1
    eps = np.random.default_rng(0).standard_normal((1, N_ROLLOUTS, T_STEPS, 2)).astype(np.float32)
    o.compute_control(state, U, np.zeros(4), [0.275, 0.3], eps, threads=cores)
    times = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end and len(times) < 30:
        t0 = time.perf_counter()
        o.compute_control(state, U, np.zeros(4), [0.275, 0.3], eps, threads=cores)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    out = {"value": N_ROLLOUTS * T_STEPS / med, "unit": "rollout-steps/s", "cores": cores, "kind": "port",
           "sample": "%d x computeControl(1920x100), oracle/mppi_oracle.c, %d pthreads, median" % (len(times), cores)}
    # single-thread straight loop and the ml_pipeline-style torch float64 model (dynamics only)
    t0 = time.perf_counter()
    o.dynamics_rollouts(state, U, [0.275, 0.3], eps[0][:480])
    out["dynamics_only_1thread"] = 480 * T_STEPS / (time.perf_counter() - t0)
    try:
        import torch
        from autorally_b200.params import unpack_nn_params
        ws, bs = unpack_nn_params(models["autorally_nnet_theta"], models["autorally_nnet_structure"])
        torch.set_num_threads(cores)
        layers = []
        for i, (w, b) in enumerate(zip(ws, bs)):
            lin = torch.nn.Linear(w.shape[1], w.shape[0]).double()
            lin.weight.data = torch.from_numpy(np.asarray(w, np.float64))
            lin.bias.data = torch.from_numpy(np.asarray(b, np.float64))
            layers.append(lin)
            if i < len(ws) - 1:
                layers.append(torch.nn.Tanh())
        net = torch.nn.Sequential(*layers)
        s = torch.from_numpy(np.broadcast_to(state.astype(np.float64), (N_ROLLOUTS, 7)).copy())
        u = torch.from_numpy(np.broadcast_to(U.astype(np.float64), (N_ROLLOUTS, T_STEPS, 2)).copy())

        def roll():
            x = s.clone()
            with torch.no_grad():
                for t in range(T_STEPS):
                    y = net(torch.cat([x[:, 3:7], u[:, t]], 1))
                    der = torch.stack([torch.cos(x[:, 2]) * x[:, 4] - torch.sin(x[:, 2]) * x[:, 5],
                                       torch.sin(x[:, 2]) * x[:, 4] + torch.cos(x[:, 2]) * x[:, 5], -x[:, 6]], 1)
                    x = x + torch.cat([der, y], 1) * 0.02
            return x
        roll()
        tt = []
        for _ in range(5):
            t0 = time.perf_counter()
            roll()
            tt.append(time.perf_counter() - t0)
        out["torch_f64_dynamics_only"] = N_ROLLOUTS * T_STEPS / statistics.median(tt)
    except Exception as e:  # pragma: no cover
        out["torch_f64_dynamics_only"] = "unavailable: %r" % (e,)
    return out


def run_reference(args):
    """--impl reference.  The reference implements computeControl only in CUDA (there is no CPU path), so when
    oracle/_ref/libautorally_ref.so (the reference's own sources built by oracle/refbuild.py) and a GPU are present this
    arm runs the UNMODIFIED reference controller -- its kernels, cuRAND noise, host syncs and copies -- through its public
    computeControl(state) on the same B200.  Otherwise it times the CPU port (oracle/mppi_oracle.c) on all host cores.
    Under torchrun only rank 0 works; the other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    models, costmap, cp, state, U = load_setup()
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "rollout-steps/sec", "unit": "rollout-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    gpu_ok = False
    try:
        import torch
        from oracle import reference as ref
        gpu_ok = ref.available() and torch.cuda.is_available()
    except Exception:
        gpu_ok = False
    if gpu_ok:
        with ref.ReferenceController(ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp) as rc:
            rc.set_controls(U, np.zeros(4, np.float32))
            rc.time_compute_control(state, reps=max(args.warmup, 1))
            # K consecutive computeControl(state) calls timed inside the harness (oracle/ref_harness.cu: CUDA events
            # around the loop; the reference's own host syncs, copies and host-side smoothing / nominal rollout included)
            ms_per_call = rc.time_compute_control(state, reps=args.steps)
        dt = ms_per_call * 1e-3 * args.steps
        val = N_ROLLOUTS * T_STEPS * args.steps / dt
        line.update(value=val, ms_per_step=1e3 * dt / args.steps,
                    config=bench_config(1),
                    reference_impl="rdesc/autorally MPPIController<NeuralNetModel<7,2,3,6,32,32,4>,MPPICosts,1920,8,16>::computeControl, "
                                   "its own CUDA kernels and cuRAND noise, sources compiled unmodified for sm_100a (oracle/refbuild.py), "
                                   "on this B200",
                    ms_per_call_mean=ms_per_call,
                    cpu_baseline={"value": val, "unit": "rollout-steps/s", "cores": 1, "kind": "reference",
                                  "sample": "%d x reference computeControl(1920x100) from oracle/_ref (GPU kernels + 1 host thread; "
                                            "the reference has no CPU implementation of this path)" % args.steps},
                    e2e={"value": val, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    else:
        from tests.common import make_oracle
        o = make_oracle("nn", models, costmap, cp)
        eps = np.random.default_rng(0).standard_normal((1, N_ROLLOUTS, T_STEPS, 2)).astype(np.float32)
        Uc = U.copy()
        for _ in range(args.warmup):
            Uc = o.compute_control(state, Uc, np.zeros(4), [0.275, 0.3], eps, threads=cores)["U"]
        t0 = time.perf_counter()
        for _ in range(args.steps):
            Uc = o.compute_control(state, Uc, np.zeros(4), [0.275, 0.3], eps, threads=cores)["U"]
        dt = time.perf_counter() - t0
        val = N_ROLLOUTS * T_STEPS * args.steps / dt
        line.update(value=val, ms_per_step=1e3 * dt / args.steps, config=bench_config(1),
                    reference_impl="CPU port of the reference (oracle/mppi_oracle.c) on all host cores: oracle/_ref or a GPU is missing",
                    cpu_baseline={"value": val, "unit": "rollout-steps/s", "cores": cores, "kind": "port",
                                  "sample": "%d x full computeControl, oracle/mppi_oracle.c on %d pthreads" % (args.steps, cores)},
                    e2e={"value": val, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written on this pool's B200s) or the profiling guide's fallbacks."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def measure_large(models, costmap, cp, state, U, n_rollouts, r_begin=0, r_count=0, steps=3, fp32_peak=None, hbm_peak=None):
    from tests.common import make_context
    out = {}
    with make_context("nn", models, costmap, cp, n_rollouts, rollout_begin=r_begin, rollout_count=r_count) as ctx:
        ctx.compute_control(state, U)
        ctx.run_resident(1)
        ms, rk = ctx.run_resident(steps, time_rollout=True)
        n_local = ctx.n_local
        out.update(rollouts=n_local, steps=steps, ms_per_step=ms / steps, rollout_kernel_ms=rk / steps,
                   value=n_local * T_STEPS * steps / (ms * 1e-3), variant=ctx.resolved_variant())
        tf = FLOP_PER_ROLLOUT_STEP_NN * n_local * T_STEPS / (rk / steps * 1e-3) / 1e12
        out["rollout_tflops"] = tf
        if fp32_peak:
            out["rollout_frac_of_fp32_peak"] = tf / fp32_peak
        if out["variant"] == 10:
            # rollout_tc_kernel: the contraction runs on the tensor pipe (tcgen05, FP16 hi/lo split = 3 passes, K and N padded
            # to the MMA shapes), so the algorithmic FP32 rate above may exceed the CUDA-core FFMA peak; what the kernel is
            # bound by is the tanh / split / cost epilogue on the CUDA cores (profiles/ncu_1m_r01t.txt).
            issued = TENSOR_FLOP_ISSUED_PER_ROLLOUT_STEP * n_local * T_STEPS / (rk / steps * 1e-3) / 1e12
            peaks = measured_peaks()
            out.update(kernel="rollout_tc_kernel", tensor_tflops_issued=issued,
                       tensor_frac_of_measured_dense_16bit_peak=issued / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
                       note="layer contractions on tcgen05 (A in tensor memory); rollout_tflops counts the 2756 algorithmic "
                            "FLOP per rollout-step, tensor_tflops_issued the 11264 FLOP the 14 MMAs per tile-step execute")
    return out


def measure_other_configs(models, costmap, cp, world, rank, local_rank, barrier, max_over_ranks):
    """BASELINE configs[2] (path_integral_bf, 2560 x 100) and configs[4] (batched MPC: 4096 independent controllers x
    256 rollouts x 100, controllers sharded over the ranks with no communication), device-resident, CUDA events."""
    from autorally_b200.params import ellipse_states
    from autorally_b200.sharding import controller_shard
    from tests.common import cost_params_for, make_context, straight_controls, top_state
    out = {}
    cp_bf = cost_params_for(costmap, desired_speed=6.0)
    with make_context("bf", models, costmap, cp_bf, 2560, device=local_rank) as ctx:
        ctx.compute_control(top_state(4.0), straight_controls(T_STEPS))
        ctx.run_resident(3)
        ms, rk = ctx.run_resident(20, time_rollout=True)
        out["bf_2560x100"] = {"ms_per_step": ms / 20, "rollout_kernel_ms": rk / 20, "value": 2560 * T_STEPS * 20 / (ms * 1e-3),
                              "unit": "rollout-steps/s", "note": "per GPU, one controller"}
    # the fork's wider / deeper dynamics network 6-64-64-64-64-4 (SRC/params/models/wider_deeper_network_08_20_2020.npz)
    with make_context("nn", models, costmap, cp, N_ROLLOUTS, tag="wider_deeper", negate_yaw_der=False, device=local_rank) as ctx:
        ctx.compute_control(top_state(4.0), straight_controls(T_STEPS))
        ctx.run_resident(3)
        ms, rk = ctx.run_resident(20, time_rollout=True)
        out["wider_deeper_1920x100"] = {"ms_per_step": ms / 20, "rollout_kernel_ms": rk / 20, "value": N_ROLLOUTS * T_STEPS * 20 / (ms * 1e-3),
                                        "unit": "rollout-steps/s", "variant": ctx.resolved_variant(),
                                        "note": "per GPU, one controller; rollout_tc_kernel<64,4> (tcgen05)"}
    B_total, n = 4096, 256
    b0, B = controller_shard(rank, world, B_total)
    states = ellipse_states(B_total)[b0:b0 + B]
    Ub = np.broadcast_to(straight_controls(T_STEPS), (B, T_STEPS, 2)).copy()
    with make_context("nn", models, costmap, cp, n, num_controllers=B, device=local_rank) as ctx:
        ctx.compute_control(states, Ub)
        ctx.run_resident(1)
        barrier()
        ms, rk = ctx.run_resident(3, time_rollout=True)
        barrier()
        ms = max_over_ranks(ms)
        out["batched_4096x256x100"] = {"controllers_per_gpu": B, "ms_per_step": ms / 3, "rollout_kernel_ms": rk / 3,
                                       "value": B_total * n * T_STEPS * 3 / (ms * 1e-3), "unit": "rollout-steps/s",
                                       "controllers_per_s": B_total * 3 / (ms * 1e-3),
                                       "scaling": "strong (4096 controllers split over the ranks, no communication)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-large", action="store_true", help="skip the 1M-rollout section")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from autorally_b200 import capi
    from tests.common import make_context
    models, costmap, cp, state, U = load_setup()
    fp32_peak = capi.measure_fp32_peak()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        hbm_peak, hbm_src = json.load(open(peaks_path)).get("hbm_gbs", 6650.0), "measured"

    ctx = make_context("nn", models, costmap, cp, N_ROLLOUTS, device=local_rank)
    out = ctx.compute_control(state, U)  # initialises the device-resident state / U
    # ---- device-resident throughput (value): L2 flushed between timed steps ----
    ctx.run_resident(args.warmup, flush_l2=True)
    barrier()
    with ClockSampler(local_rank) as clk:
        ms, _ = ctx.run_resident(args.steps, flush_l2=True)
    barrier()
    # the dominant kernel's own duration: a separate pass with CUDA events around the rollout launch (events inside the
    # pipeline serialise it, so this pass is not the one `value` comes from)
    _, rk = ctx.run_resident(args.steps, time_rollout=True, flush_l2=True)
    launches = ctx.last_launch_count()
    ms = max_over_ranks(ms)
    value = world * N_ROLLOUTS * T_STEPS * args.steps / (ms * 1e-3)
    # warm-L2 figure (how the controller actually runs: same buffers every call)
    ctx.run_resident(args.warmup)
    ms_warm, _ = ctx.run_resident(args.steps)
    # ---- end to end through the C ABI with host buffers (e2e) + latency percentiles ----
    # Every call stages state / U / history into pinned memory, launches the graph (H2D 848 B -> 4 kernels -> D2H
    # 5216 B), waits and unpacks into the caller's host arrays.  The loop runs on the C side of the ABI (what a C++
    # control loop sees); the same through the ctypes binding, which adds Python argument marshalling, is reported too.
    hist = np.zeros(4, np.float32)
    ctx.bench_compute_control(state, U, hist, reps=args.warmup)
    barrier()
    t_all = time.perf_counter()
    lat_c, Uc = ctx.bench_compute_control(state, U, hist, reps=args.steps)
    e2e_s = max_over_ranks(time.perf_counter() - t_all)
    e2e_launches = ctx.last_launch_count() * args.steps
    lat_c = np.sort(lat_c)
    lat = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        Uc = ctx.compute_control(state, Uc, hist)["U"]
        lat.append(time.perf_counter() - t0)
    lat.sort()
    e2e = {"value": world * N_ROLLOUTS * T_STEPS * args.steps / e2e_s, "unit": "rollout-steps/s",
           "h2d_bytes_per_step": int(4 * (12 + 2 * T_STEPS)), "d2h_bytes_per_step": int(4 * (4 + 13 * T_STEPS)),
           "p50_ms": float(lat_c[len(lat_c) // 2]), "p99_ms": float(lat_c[min(len(lat_c) - 1, int(0.99 * len(lat_c)))]),
           "caller": "C loop over mppi_compute_control (mppi_bench_compute_control)",
           "ctypes_binding_p50_ms": 1e3 * lat[len(lat) // 2]}
    rollout_ms = rk / args.steps
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")
    if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of the rollout kernel, one ncu --set full capture
        t = json.load(open(tpath))["configs"]
        k = t["1920"].get("rollout_half_kernel") or next(iter(t["1920"].values()))
        traffic, traffic_src = k["dram_read_bytes"] + k["dram_write_bytes"], "profiles/ncu_traffic_r01.json (cold-cache ncu replay)"
    achieved = FLOP_PER_ROLLOUT_STEP_NN * N_ROLLOUTS * T_STEPS / (rollout_ms * 1e-3) / 1e12
    roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": 16.0 * N_ROLLOUTS * T_STEPS,
                "kernel": "rollout_half_kernel" if ctx.resolved_variant() == 9 else "rollout kernel (variant %d)" % ctx.resolved_variant(),
                "kernel_ms": rollout_ms,
                "note": "FP32 FFMA issue bound (CUDA cores; neither HBM nor tensor); peak = FFMA microbenchmark measured in this run; "
                        "1920 rollouts occupy <2% of the machine, see 'large' for the filled-GPU fraction"}
    line = {"metric": "rollout-steps/sec", "value": value, "unit": "rollout-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(world), "rollout_variant": ctx.resolved_variant(),
            "l2": "flushed between timed steps (256 MiB memset outside the timed intervals)",
            "per_gpu": value / world,
            "ms_per_step_warm_l2": ms_warm / args.steps, "e2e": e2e, "gpu_launches": launches + e2e_launches,
            "roofline": roofline, "clocks": clk.summary(), "fp32_peak_tflops_measured": fp32_peak,
            "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
            "trajectory_cost": float(out["trajectory_cost"])}
    ctx.close()
    if not args.no_large:
        if world == 1:
            line["large"] = measure_large(models, costmap, cp, state, U, LARGE_ROLLOUTS, fp32_peak=fp32_peak)
        else:
            line["sharded_large"] = run_sharded_large(models, costmap, cp, state, U, world, rank, local_rank, barrier, max_over_ranks)
    if not args.no_large:
        line["other_configs"] = measure_other_configs(models, costmap, cp, world, rank, local_rank, barrier, max_over_ranks)
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(models, costmap, cp, state, U)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_sharded_large(models, costmap, cp, state, U, world, rank, local_rank, barrier, max_over_ranks, steps=20):
    """configs[3]: 1M rollouts x 100 steps sharded over the ranks; per step ONE exchange of the (4 + 2T)-float shard
    record per rank.  Measured twice: with the peer-memory exchange fused into the weighting /
    finalize kernels (mppi_p2p_init: NVLink stores + flags, no collective launch) and with one ncclAllGather issued by
    the library on the context's stream between those kernels.  Device-resident, timed with CUDA events on that
    stream, max over ranks; one untimed sharded step right before the timed ones aligns the ranks on the device."""
    import torch.distributed as dist
    from autorally_b200.capi import MppiContext
    from autorally_b200.sharding import rollout_shard
    from tests.common import make_context
    lo, n = rollout_shard(rank, world, LARGE_ROLLOUTS)
    ids = [MppiContext.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = make_context("nn", models, costmap, cp, LARGE_ROLLOUTS, rollout_begin=lo, rollout_count=n, device=local_rank)
    ctx.comm_init(ids[0], rank, world)
    out = ctx.compute_control_sharded(state, U)      # initialises the device-resident state / U

    def timed():
        ctx.run_resident_sharded(2)
        barrier()
        ctx.run_resident_sharded(1)                  # its exchange lines the ranks up on the device
        ms = ctx.run_resident_sharded(steps)
        barrier()
        return max_over_ranks(ms)

    ms_nccl = timed()
    handles = [None] * world
    dist.all_gather_object(handles, ctx.p2p_export(world))
    ctx.p2p_init(b"".join(handles), rank, world)
    out_p2p = ctx.compute_control_sharded(state, U)
    ms = timed()
    sf = ctx.shard_floats()
    ctx.close()
    # the latency end of the same protocol: the 1920-rollout controller sharded over the ranks (the rollout kernel is a
    # dependent chain of 100 timesteps, so sharding cannot shorten it; what shows here is the cost of the exchange itself)
    lo_s, n_s = rollout_shard(rank, world, N_ROLLOUTS)
    small = {}
    cs = make_context("nn", models, costmap, cp, N_ROLLOUTS, rollout_begin=lo_s, rollout_count=n_s, device=local_rank)
    cs.comm_init(_fresh_id(dist, rank), rank, world)
    cs.compute_control_sharded(state, U)

    def timed_small(k=200):
        cs.run_resident_sharded(20)
        barrier()
        cs.run_resident_sharded(1)
        t = cs.run_resident_sharded(k)
        barrier()
        return max_over_ranks(t) / k
    small["ms_per_step_nccl_allgather"] = timed_small()
    hs = [None] * world
    dist.all_gather_object(hs, cs.p2p_export(world))
    cs.p2p_init(b"".join(hs), rank, world)
    cs.compute_control_sharded(state, U)
    small["ms_per_step"] = timed_small()
    small["rollouts_per_gpu"] = n_s
    cs.close()
    return {"rollouts": LARGE_ROLLOUTS, "rollouts_per_gpu": n, "steps": steps, "ms_per_step": ms / steps,
            "value": LARGE_ROLLOUTS * T_STEPS * steps / (ms * 1e-3), "unit": "rollout-steps/s", "scaling": "strong",
            "exchange": "peer-memory: the weighting kernel stores each rank's %d-float record into every GPU's mailbox over "
                        "NVLink and raises a flag, finalize waits on the flags; no collective launch" % sf,
            "ms_per_step_nccl_allgather": ms_nccl / steps,
            "value_nccl_allgather": LARGE_ROLLOUTS * T_STEPS * steps / (ms_nccl * 1e-3),
            "timing": "CUDA events on the context's stream, max over ranks", "normalizer": float(out_p2p["normalizer"]),
            "normalizer_nccl": float(out["normalizer"]), "sharded_1920": small}


def _fresh_id(dist, rank):
    from autorally_b200.capi import MppiContext
    ids = [MppiContext.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ids[0]


if __name__ == "__main__":
    main()
