#!/usr/bin/env python
"""bench.py -- rollout-steps/s of the MPPI hot path (computeControl) on B200.

    python bench.py --gpus 1 --steps 200 --warmup 10            # our CUDA path
    python bench.py --impl reference --steps 20 --warmup 3      # the reference's own computeControl (oracle/_ref) on the same GPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete computeControl pipeline (Philox noise -> fused rollouts -> importance weighting ->
Savitzky-Golay -> nominal trajectory) of BASELINE.json configs[1]: path_integral_nn, 1920 rollouts x 100 timesteps on the
synthetic ellipse costmap.  A timed batch is exactly --steps steps (CUDA events on the context's stream, L2 flushed between
steps); batches are repeated until >= 250 ms have been spent under load and the MEDIAN batch is reported, with the clock
sampler running over the whole interval.  At N > 1 every rank runs one such controller (batched-MPC sharding: independent
controllers, no communication) for `value`; the communicating workloads are reported beside it with their N = 1
denominators measured in the same run (rank 0 alone): 1M rollouts sharded over the ranks (strong scaling), 1M rollouts per
rank (weak scaling), 4096 controllers x 256 rollouts split over the ranks, and a sharded-vs-unsharded parity figure.
Prints ONE JSON line on rank 0.
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
FLOP_PER_ROLLOUT_STEP_BF = 200.0   # 2*(4*25); the 25 basis functions are excluded (MUFU utilisation reported by ncu)
N_ROLLOUTS, T_STEPS = 1920, 100
# rollout_tc.cu: per 128-rollout tile and timestep 2 MMAs 128x32x16 (layer 1), 6 of 128x32x16, 6 of 128x16x16
TENSOR_FLOP_ISSUED_PER_ROLLOUT_STEP = 2.0 * (2 * 32 * 16 + 6 * 32 * 16 + 6 * 16 * 16)
LARGE_ROLLOUTS = 1 << 20           # "large-sample MPPI": 1M rollouts (16384 x 64)
PARITY_ROLLOUTS = 65536            # sharded-vs-unsharded parity check at N > 1
MIN_TIMED_MS = 250.0               # every headline figure is the median of batches covering at least this much load
SEED = 1234

WORKLOAD = "path_integral_nn: NeuralNetModel<7,2,3,6,32,32,4>, 1920 rollouts x 100 steps, synthetic ellipse costmap"


def bench_config(world):
    return {"workload": WORKLOAD, "rollouts": N_ROLLOUTS, "timesteps": T_STEPS,
            "parallelism": "one controller per GPU x%d (independent controllers, no communication)" % world}


def load_setup():
    from autorally_b200.params import make_ellipse_costmap
    from autorally_b200.scenarios import cost_params_for, straight_controls, top_state
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


def timed_batches(run_batch, min_ms=MIN_TIMED_MS, min_batches=3, max_batches=4000):
    """Repeats run_batch() -> device ms until min_ms of load have been timed; returns the list of batch times."""
    out, total = [], 0.0
    while (total < min_ms or len(out) < min_batches) and len(out) < max_batches:
        ms = run_batch()
        out.append(ms)
        total += ms
    return out


def cpu_baseline(models, costmap, cp, state, U, budget_s=12.0):
    """The CPU restatement (oracle port) on the host's cores: full computeControl, all threads."""
    from oracle.oracle import make_oracle
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


def _reference_arm(ref, kind, theta, costmap, cp, state, U, steps, warmup, n_rollouts, **kw):
    """One reference controller: K consecutive computeControl(state) calls timed inside the harness (CUDA events around the
    loop; the reference's own host syncs, copies and host-side smoothing / nominal rollout included), plus the device time
    of its four GPU stages alone (oracle/ref_harness.cu::time_kernels)."""
    with ref.ReferenceController(kind, theta, costmap, cp, **kw) as rc:
        rc.set_controls(U, np.zeros(4, np.float32))
        rc.time_compute_control(state, reps=max(warmup, 1))
        calls = []
        while sum(calls) < MIN_TIMED_MS or len(calls) < 3:
            calls.append(rc.time_compute_control(state, reps=steps) * steps)
        ms_per_call = statistics.median(calls) / steps
        k = rc.time_kernels(state, reps=max(steps, 10))
    kernel_ms = sum(k.values())
    return {"ms_per_call": ms_per_call, "value": n_rollouts * T_STEPS / (ms_per_call * 1e-3), "unit": "rollout-steps/s",
            "kernel_ms": kernel_ms, "kernels_ms": k, "host_overhead_ms": ms_per_call - kernel_ms, "batches": len(calls)}


def run_reference(args):
    """--impl reference.  The reference implements computeControl only in CUDA (it has no CPU path), so when
    oracle/_ref/libautorally_ref.so (the reference's own sources built by oracle/refbuild.py) and a GPU are present this
    arm runs the UNMODIFIED reference controller -- its kernels, cuRAND noise, host syncs and copies -- through its public
    computeControl(state) on the same B200.  Otherwise it times the CPU port (oracle/mppi_oracle.c) on all host cores.
    Under torchrun only rank 0 works; the other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from autorally_b200.scenarios import cost_params_for
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
        with ClockSampler(0) as clk:
            nn = _reference_arm(ref, ref.REF_NN_1920, models["autorally_nnet_theta"], costmap, cp, state, U, args.steps, args.warmup, N_ROLLOUTS)
        val = nn["value"]
        line.update(value=val, ms_per_step=nn["ms_per_call"], config=bench_config(1), clocks=clk.summary(),
                    reference_impl="rdesc/autorally MPPIController<NeuralNetModel<7,2,3,6,32,32,4>,MPPICosts,1920,8,16>::computeControl, "
                                   "its own CUDA kernels and cuRAND noise, sources compiled unmodified for sm_100a (oracle/refbuild.py), "
                                   "on this B200",
                    ms_per_call_mean=nn["ms_per_call"],
                    reference={"kernel_ms": nn["kernel_ms"], "kernels_ms": nn["kernels_ms"], "host_overhead_ms": nn["host_overhead_ms"],
                               "note": "kernel_ms = device time of curandGenerateNormal + rolloutKernel + normExpKernel + "
                                       "weightedReductionKernel, each launched as computeControl launches it but without the host-side "
                                       "deep copy of the by-value MPPICosts argument; host_overhead_ms = ms_per_call - kernel_ms "
                                       "(param memcpys, 3 stream syncs, 2 D2H cost copies, host min / normaliser loops, smoothing, "
                                       "nominal rollout, the costmap vector copy at launch)"},
                    cpu_baseline={"value": val, "unit": "rollout-steps/s", "cores": 1, "kind": "reference",
                                  "sample": "median of %d batches of %d x reference computeControl(1920x100) from oracle/_ref (GPU kernels + 1 "
                                            "host thread; the reference has no CPU implementation of this path)" % (nn["batches"], args.steps)},
                    e2e={"value": val, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        # the other single-GPU configurations, so that every config has a same-box reference number
        other = {}
        try:
            cp_bf = cost_params_for(costmap, desired_speed=6.0)
            other["bf_2560x100"] = _reference_arm(ref, ref.REF_BF_2560, models["basis_function_W"], costmap, cp_bf, state, U, args.steps, args.warmup,
                                                  2560, init_u=(0.0, -0.01))
            other["wider_deeper_1920x100"] = _reference_arm(ref, ref.REF_NN64_1920, models["wider_deeper_theta"], costmap, cp, state, U,
                                                            max(args.steps // 2, 5), args.warmup, N_ROLLOUTS, negate_yaw_der=False)
        except Exception as e:  # pragma: no cover
            other["error"] = repr(e)
        line["other_configs"] = other
    else:
        from oracle.oracle import make_oracle
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
        d = json.load(open(path))
        d["source"] = "MEASURED_PEAKS.json"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def resident_median(ctx, steps, flush_l2=False, min_ms=MIN_TIMED_MS):
    """Median device ms per step over batches of `steps` resident steps covering >= min_ms."""
    b = timed_batches(lambda: ctx.run_resident(steps, flush_l2=flush_l2)[0], min_ms=min_ms)
    return statistics.median(b) / steps, len(b)


def measure_large(kind, models, costmap, cp, state, U, n_rollouts, device, fp32_peak, steps=5, **ctx_kw):
    """One controller of n_rollouts on one GPU, device-resident: step time (median of batches), the rollout kernel's own
    duration, roofline fractions with both denominators, and the normaliser for the fixed Philox (seed, call 0)."""
    from autorally_b200.scenarios import make_context
    out = {}
    peaks = measured_peaks()
    with make_context(kind, models, costmap, cp, n_rollouts, device=device, seed=SEED, **ctx_kw) as ctx:
        ctx.seed(SEED, 0)
        res = ctx.compute_control(state, U)
        out["normalizer"] = float(res["normalizer"])
        out["normalizer_note"] = "Philox seed %d, call 0: the sharded runs at N > 1 report the same figure" % SEED
        ctx.run_resident(2)
        ms_step, nb = resident_median(ctx, steps)
        rk = statistics.median(ctx.run_resident(steps, time_rollout=True)[1] / steps for _ in range(3))
        n_local = ctx.n_local
        # the HBM-bound stages on their own (north_star: "achieved HBM GB/s for the noise and reduction stages ... against the
        # B200 peak"): algorithmic bytes = 8 B per rollout-step written by the sampler; 8 B per rollout-step + 4 B per rollout
        # read by the weighting kernel
        st = ctx.time_stages(5)
        hbm = peaks.get("hbm_gbs") or 6650.0
        samp_gbs = 8.0 * n_local * T_STEPS / (st["sampler"] * 1e-3) / 1e9
        wred_gbs = (8.0 * n_local * T_STEPS + 4.0 * n_local) / (st["weighting"] * 1e-3) / 1e9
        out["stages"] = {"ms": st, "sampler": {"bound": "hbm", "achieved": samp_gbs, "peak": hbm, "unit": "GB/s", "frac": samp_gbs / hbm,
                                               "note": "stand-alone sample_noise_kernel (the pipeline above draws its noise inside the rollout "
                                                       "kernel when launches_per_step is 3); Philox-multiply-bound"},
                         "weighting": {"bound": "hbm", "achieved": wred_gbs, "peak": hbm, "unit": "GB/s", "frac": wred_gbs / hbm},
                         "peak_source": peaks["source"]}
        out.update(rollouts=n_local, steps=steps, batches=nb, ms_per_step=ms_step, rollout_kernel_ms=rk,
                   value=n_local * T_STEPS / (ms_step * 1e-3), variant=ctx.resolved_variant(),
                   launches_per_step=ctx.last_launch_count() // steps)
        flop = FLOP_PER_ROLLOUT_STEP_BF if kind == "bf" else FLOP_PER_ROLLOUT_STEP_NN
        tf = flop * n_local * T_STEPS / (rk * 1e-3) / 1e12
        roof = {"kernel_ms": rk, "algorithmic_flop_per_rollout_step": flop, "achieved_tflops": tf,
                "fp32_ffma_peak_tflops": fp32_peak, "frac_of_fp32_ffma_peak": tf / fp32_peak if fp32_peak else None,
                "algorithmic_bytes_per_launch": (8.0 if out["launches_per_step"] == 3 else 16.0) * n_local * T_STEPS,
                "hbm_peak_gbs": peaks.get("hbm_gbs")}
        if kind == "nn" and out["variant"] == 10 and ctx_kw.get("tag", "autorally_nnet") == "autorally_nnet":
            # rollout_tc_kernel: the contraction runs on the tensor pipe (tcgen05, FP16 hi/lo split = 3 passes, K and N padded
            # to the MMA shapes), so the algorithmic FP32 rate may exceed the CUDA-core FFMA peak; what the kernel is
            # bound by is the tanh / split / cost epilogue on the CUDA cores (profiles/ncu_1m_*.txt).
            issued = TENSOR_FLOP_ISSUED_PER_ROLLOUT_STEP * n_local * T_STEPS / (rk * 1e-3) / 1e12
            roof.update(kernel="rollout_tc_kernel", tensor_tflops_issued=issued, bf16_burst_peak_tflops=peaks.get("bf16_tflops"),
                        frac_of_bf16_burst_peak_issued=issued / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
                        frac_of_bf16_burst_peak_algorithmic=tf / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
                        note="layer contractions on tcgen05 (A in tensor memory); achieved_tflops counts the 2756 algorithmic FLOP "
                             "per rollout-step, tensor_tflops_issued the 11264 FLOP the 14 MMAs per tile-step execute; the binding "
                             "unit is instruction issue of the CUDA-core epilogue (ncu: profiles/)")
        elif kind == "bf":
            roof.update(kernel="rollout_kernel<CarBasisDyn>", note="25 basis functions (tanf / atanf / sinf, 22 divisions) per rollout-step are "
                        "excluded from the 200 algorithmic FLOP: the kernel is transcendental-bound, MUFU utilisation in profiles/ncu_bf_1m_*.txt")
        out["roofline"] = roof
    return out


def measure_other_configs(models, costmap, cp, local_rank):
    """BASELINE configs[2] (path_integral_bf, 2560 x 100) and the fork's wider / deeper network at 1920 x 100: one controller
    per GPU, device-resident, CUDA events, median of batches covering >= 250 ms."""
    from autorally_b200.scenarios import cost_params_for, make_context, straight_controls, top_state
    out = {}
    cp_bf = cost_params_for(costmap, desired_speed=6.0)
    for name, kind, n, kw, c in (("bf_2560x100", "bf", 2560, {}, cp_bf),
                                 ("wider_deeper_1920x100", "nn", N_ROLLOUTS, dict(tag="wider_deeper", negate_yaw_der=False), cp)):
        with make_context(kind, models, costmap, c, n, device=local_rank, **kw) as ctx:
            ctx.compute_control(top_state(4.0), straight_controls(T_STEPS))
            ctx.run_resident(3)
            ms_step, nb = resident_median(ctx, 20)
            rk = ctx.run_resident(20, time_rollout=True)[1] / 20
            lat, _ = ctx.bench_compute_control(top_state(4.0), straight_controls(T_STEPS), np.zeros(4, np.float32), reps=200)
            out[name] = {"ms_per_step": ms_step, "rollout_kernel_ms": rk, "value": n * T_STEPS / (ms_step * 1e-3), "unit": "rollout-steps/s",
                         "e2e_p50_ms": float(np.median(lat)), "variant": ctx.resolved_variant(), "batches": nb,
                         "note": "per GPU, one controller; e2e_p50_ms = host-observed mppi_compute_control latency (C loop, 200 calls)"}
    return out


class Dist:
    """The torch.distributed plumbing bench.py needs (barrier + max over ranks + object exchange); trivial at N = 1."""

    def __init__(self, world, rank, local_rank):
        import torch
        self.torch, self.world, self.rank, self.local_rank = torch, world, rank, local_rank
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def bcast(self, obj):
        if not self.dist:
            return obj
        box = [obj if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather(self, obj):
        if not self.dist:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


class ShardedRun:
    """One controller whose rollouts are sharded over the ranks: this rank's context, connected
    through NCCL (one ncclAllGather of the shard records per step) and, after connect_p2p(), through the peer-memory
    exchange fused into the weighting / finalize kernels."""

    def __init__(self, d, kind, models, costmap, cp, n_global, **kw):
        from autorally_b200.capi import MppiContext
        from autorally_b200.scenarios import make_context
        from autorally_b200.sharding import rollout_shard
        self.d = d
        lo, n = rollout_shard(d.rank, d.world, n_global)
        self.n_local, self.n_global = n, n_global
        self.ctx = make_context(kind, models, costmap, cp, n_global, rollout_begin=lo, rollout_count=n, device=d.local_rank, seed=SEED, **kw)
        self.ctx.comm_init(d.bcast(MppiContext.comm_unique_id() if d.rank == 0 else None), d.rank, d.world)

    def connect_p2p(self):
        h = self.d.gather(self.ctx.p2p_export(self.d.world))
        self.ctx.p2p_init(b"".join(h), self.d.rank, self.d.world)

    def compute(self, state, U, call=0):
        self.ctx.seed(SEED, call)
        return self.ctx.compute_control_sharded(state, U)

    def timed(self, steps, min_ms=MIN_TIMED_MS):
        """Median over batches of `steps` resident sharded steps; every batch is bracketed by barriers, preceded by one
        untimed sharded step whose exchange lines the ranks up on the device, timed with CUDA events on the context's
        stream, max over ranks."""
        c, d = self.ctx, self.d
        c.run_resident_sharded(2)

        def batch():
            d.barrier()
            c.run_resident_sharded(1)
            ms = c.run_resident_sharded(steps)
            d.barrier()
            return d.max(ms)
        b = timed_batches(batch, min_ms=min_ms)
        return statistics.median(b) / steps, len(b)

    def close(self):
        self.ctx.close()


def max_rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / (1.0 + np.abs(b))))


def run_multi_gpu(d, models, costmap, cp, state, U, fp32_peak):
    """Everything that communicates, with the N = 1 denominators measured in the same run by rank 0 alone."""
    from autorally_b200.params import ellipse_states
    from autorally_b200.scenarios import make_context, straight_controls
    from autorally_b200.sharding import controller_shard
    out = {}
    world = d.world
    # ---- N = 1 denominators (rank 0 alone; the other ranks wait at the barrier) ----
    base = None
    if d.rank == 0:
        base = {"large": measure_large("nn", models, costmap, cp, state, U, LARGE_ROLLOUTS, d.local_rank, fp32_peak)}
        with make_context("nn", models, costmap, cp, PARITY_ROLLOUTS, device=d.local_rank, seed=SEED) as ctx:
            ctx.seed(SEED, 0)
            base["parity"] = ctx.compute_control(state, U)
    d.barrier()
    base = d.bcast(base)
    t1 = base["large"]["ms_per_step"]
    # ---- sharded-vs-unsharded parity: 65536 rollouts, same Philox seed and call, NCCL and peer-memory exchanges ----
    par = ShardedRun(d, "nn", models, costmap, cp, PARITY_ROLLOUTS)
    got_nccl = par.compute(state, U)
    par.connect_p2p()
    got_p2p = par.compute(state, U)
    par.close()
    errs = []
    for got in (got_nccl, got_p2p):
        errs.append(max(max_rel(got["U"], base["parity"]["U"]), max_rel(got["normalizer"], base["parity"]["normalizer"]),
                        max_rel(got["baseline"], base["parity"]["baseline"]), max_rel(got["state_solution"], base["parity"]["state_solution"])))
    out["sharded_parity_max_rel"] = d.max(max(errs))
    out["sharded_parity"] = {"rollouts": PARITY_ROLLOUTS, "max_rel_nccl_allgather": d.max(errs[0]), "max_rel_peer_memory": d.max(errs[1]),
                             "normalizer_sharded": float(got_p2p["normalizer"]), "normalizer_unsharded": float(base["parity"]["normalizer"]),
                             "what": "max over U, normalizer, baseline, state_solution of |sharded - unsharded| / (1 + |unsharded|), max over "
                                     "ranks; unsharded = rank 0's single-GPU controller, same Philox seed and call"}
    # ---- configs[3]: 1M rollouts sharded over the ranks (strong scaling) ----
    sr = ShardedRun(d, "nn", models, costmap, cp, LARGE_ROLLOUTS)
    res_nccl = sr.compute(state, U)
    ms_nccl, _ = sr.timed(10)
    sr.connect_p2p()
    res = sr.compute(state, U)
    ms, nb = sr.timed(10)
    sf = sr.ctx.shard_floats()
    sr.close()
    out["sharded_large"] = {
        "rollouts": LARGE_ROLLOUTS, "rollouts_per_gpu": sr.n_local, "steps": 10, "batches": nb, "ms_per_step": ms,
        "value": LARGE_ROLLOUTS * T_STEPS / (ms * 1e-3), "unit": "rollout-steps/s", "scaling": "strong",
        "ms_per_step_1gpu_same_run": t1, "strong_efficiency": t1 / (world * ms),
        "exchange": "peer-memory: the weighting kernel stores each rank's %d-float record into every GPU's mailbox over NVLink and "
                    "raises a flag, finalize waits on the flags; no collective launch" % sf,
        "ms_per_step_nccl_allgather": ms_nccl, "strong_efficiency_nccl_allgather": t1 / (world * ms_nccl),
        "timing": "CUDA events on the context's stream, median of batches covering >= %d ms, max over ranks" % MIN_TIMED_MS,
        "normalizer": float(res["normalizer"]), "normalizer_nccl": float(res_nccl["normalizer"]),
        "normalizer_1gpu_same_run": base["large"]["normalizer"]}
    out["sharded_large_strong_efficiency"] = out["sharded_large"]["strong_efficiency"]
    # ---- weak scaling: 1M rollouts PER GPU, one controller of world x 1M rollouts ----
    wr = ShardedRun(d, "nn", models, costmap, cp, LARGE_ROLLOUTS * world)
    wr.connect_p2p()
    wres = wr.compute(state, U)
    wms, wnb = wr.timed(5)
    wr.close()
    out["weak_large"] = {"rollouts": LARGE_ROLLOUTS * world, "rollouts_per_gpu": LARGE_ROLLOUTS, "steps": 5, "batches": wnb, "ms_per_step": wms,
                         "value": LARGE_ROLLOUTS * world * T_STEPS / (wms * 1e-3), "unit": "rollout-steps/s", "scaling": "weak",
                         "ms_per_step_1gpu_same_run": t1, "weak_efficiency": t1 / wms, "normalizer": float(wres["normalizer"]),
                         "exchange": "peer-memory (fused into the weighting / finalize kernels)"}
    out["weak_large_efficiency"] = out["weak_large"]["weak_efficiency"]
    # ---- the latency end of the same protocol: the 1920-rollout controller sharded over the ranks (the rollout kernel is a
    # dependent chain of 100 timesteps, so sharding cannot shorten it; what shows here is the cost of the exchange itself) ----
    ss = ShardedRun(d, "nn", models, costmap, cp, N_ROLLOUTS)
    ss.compute(state, U)
    small = {"ms_per_step_nccl_allgather": ss.timed(200, min_ms=100.0)[0]}
    ss.connect_p2p()
    ss.compute(state, U)
    small["ms_per_step"] = ss.timed(200, min_ms=100.0)[0]
    small["rollouts_per_gpu"] = ss.n_local
    ss.close()
    out["sharded_1920"] = small
    # ---- configs[4]: batched MPC, 4096 controllers x 256 rollouts split over the ranks, no communication ----
    B_total, n = 4096, 256
    states_all = ellipse_states(B_total)
    tb1 = None
    if d.rank == 0:
        Ub = np.broadcast_to(straight_controls(T_STEPS), (B_total, T_STEPS, 2)).copy()
        with make_context("nn", models, costmap, cp, n, num_controllers=B_total, device=d.local_rank, seed=SEED) as ctx:
            ctx.compute_control(states_all, Ub)
            ctx.run_resident(1)
            tb1, _ = resident_median(ctx, 3)
    d.barrier()
    tb1 = d.bcast(tb1)
    b0, B = controller_shard(d.rank, world, B_total)
    Ub = np.broadcast_to(straight_controls(T_STEPS), (B, T_STEPS, 2)).copy()
    with make_context("nn", models, costmap, cp, n, num_controllers=B, controller_begin=b0, device=d.local_rank, seed=SEED) as ctx:
        ctx.compute_control(states_all[b0:b0 + B], Ub)
        ctx.run_resident(1)

        def batch():
            d.barrier()
            ms = ctx.run_resident(3)[0]
            d.barrier()
            return d.max(ms)
        bb = timed_batches(batch)
        tb = statistics.median(bb) / 3
    out["batched_4096x256x100"] = {"controllers_per_gpu": B, "ms_per_step": tb, "ms_per_step_1gpu_same_run": tb1, "batches": len(bb),
                                   "value": B_total * n * T_STEPS / (tb * 1e-3), "unit": "rollout-steps/s",
                                   "controllers_per_s": B_total / (tb * 1e-3), "strong_efficiency": tb1 / (world * tb),
                                   "scaling": "strong (4096 controllers split over the ranks, no communication)"}
    out["batched_efficiency"] = out["batched_4096x256x100"]["strong_efficiency"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-large", action="store_true", help="skip the 1M-rollout / multi-GPU / other-config sections")
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
    d = Dist(world, rank, local_rank)

    from autorally_b200 import capi
    from autorally_b200.scenarios import make_context
    models, costmap, cp, state, U = load_setup()
    fp32_peak = capi.measure_fp32_peak()
    peaks = measured_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    K = args.steps
    ctx = make_context("nn", models, costmap, cp, N_ROLLOUTS, device=local_rank, seed=SEED)
    out = ctx.compute_control(state, U)  # initialises the device-resident state / U
    hist = np.zeros(4, np.float32)
    ctx.run_resident(args.warmup, flush_l2=True)
    ctx.bench_compute_control(state, U, hist, reps=max(args.warmup, 10))
    launches = 0
    with ClockSampler(local_rank) as clk:
        # ---- device-resident throughput (value): batches of exactly K steps, L2 flushed between timed steps, repeated until
        # >= 250 ms have been timed; every batch bracketed by a barrier + synchronize; max over ranks per batch; median batch ----
        def value_batch():
            nonlocal launches
            d.barrier()
            ms = ctx.run_resident(K, flush_l2=True)[0]
            launches += ctx.last_launch_count()
            d.barrier()
            return d.max(ms)
        batches = timed_batches(value_batch)
        ms_batch = statistics.median(batches)
        # ---- end to end through the C ABI with host buffers (e2e): every call stages state / U / history into pinned
        # memory, launches the graph (H2D 848 B -> kernels -> D2H 5216 B), waits and unpacks into the caller's host arrays.
        # The loop runs on the C side of the ABI (what a C++ control loop sees).  Median of >= 200 calls per rank. ----
        e2e_calls = max(200, K)
        d.barrier()
        t_all = time.perf_counter()
        lat_c, Uc = ctx.bench_compute_control(state, U, hist, reps=e2e_calls)
        wall = time.perf_counter() - t_all
        while wall < MIN_TIMED_MS * 1e-3:   # keep going until the e2e region also covers >= 250 ms
            more, Uc = ctx.bench_compute_control(state, Uc, hist, reps=e2e_calls)
            lat_c = np.concatenate([lat_c, more])
            wall = time.perf_counter() - t_all
        e2e_launches = ctx.last_launch_count() * len(lat_c)
        d.barrier()
    value = world * N_ROLLOUTS * T_STEPS * K / (ms_batch * 1e-3)
    p50 = d.max(float(np.median(lat_c)))
    p99 = d.max(float(np.sort(lat_c)[min(len(lat_c) - 1, int(0.99 * len(lat_c)))]))
    # the dominant kernel's own duration: a separate pass with CUDA events around the rollout launch (events inside the
    # pipeline serialise it, so this pass is not the one `value` comes from)
    rk_b = timed_batches(lambda: ctx.run_resident(K, time_rollout=True, flush_l2=True)[1], min_ms=20.0)
    rollout_ms = statistics.median(rk_b) / K
    # warm-L2 figure (how the controller actually runs: same buffers every call)
    ctx.run_resident(args.warmup)
    ms_warm, _ = resident_median(ctx, K, min_ms=50.0)
    lat = []
    for _ in range(100):   # the same call through the ctypes binding (adds Python argument marshalling)
        t0 = time.perf_counter()
        Uc = ctx.compute_control(state, Uc, hist)["U"]
        lat.append(time.perf_counter() - t0)
    e2e = {"value": world * N_ROLLOUTS * T_STEPS / (p50 * 1e-3), "unit": "rollout-steps/s",
           "h2d_bytes_per_step": int(4 * (12 + 2 * T_STEPS)), "d2h_bytes_per_step": int(4 * (4 + 13 * T_STEPS)),
           "p50_ms": p50, "p99_ms": p99, "calls_per_rank": int(len(lat_c)),
           "value_definition": "n_gpus x 1920 x 100 / (median host-observed latency of one mppi_compute_control call, max over ranks)",
           "wall_value": world * N_ROLLOUTS * T_STEPS * len(lat_c) / d.max(wall),
           "caller": "C loop over mppi_compute_control (mppi_bench_compute_control)",
           "ctypes_binding_p50_ms": 1e3 * statistics.median(lat)}
    traffic, traffic_src = None, None
    for tname in ("ncu_traffic_r02.json", "ncu_traffic_r01.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of the rollout kernel, one ncu --set full capture
            t = json.load(open(tpath))["configs"]
            k = t["1920"].get("rollout_half_kernel") or next(iter(t["1920"].values()))
            traffic, traffic_src = k["dram_read_bytes"] + k["dram_write_bytes"], "profiles/%s (cold-cache ncu replay)" % tname
            break
    achieved = FLOP_PER_ROLLOUT_STEP_NN * N_ROLLOUTS * T_STEPS / (rollout_ms * 1e-3) / 1e12
    roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": 16.0 * N_ROLLOUTS * T_STEPS,
                "kernel": "rollout_half_kernel" if ctx.resolved_variant() == 9 else "rollout kernel (variant %d)" % ctx.resolved_variant(),
                "kernel_ms": rollout_ms,
                "note": "FP32 FFMA issue bound (CUDA cores; neither HBM nor tensor); peak = FFMA microbenchmark measured in this run "
                        "(MEASURED_PEAKS.json has no FP32 figure); 1920 rollouts occupy <2% of the machine and the kernel is a 100-step "
                        "dependent chain: see roofline_large for the filled-GPU fractions"}
    line = {"metric": "rollout-steps/sec", "value": value, "unit": "rollout-steps/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms_batch / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(world), "rollout_variant": ctx.resolved_variant(),
            "l2": "flushed between timed steps (256 MiB memset outside the timed intervals)",
            "timed_batches": len(batches), "timed_ms_total": sum(batches),
            "timing": "median over batches of exactly --steps steps (CUDA events, barrier + synchronize around every batch, max over "
                      "ranks per batch), repeated until >= %d ms were timed" % MIN_TIMED_MS,
            "per_gpu": value / world,
            "ms_per_step_warm_l2": ms_warm, "e2e": e2e, "gpu_launches": launches + e2e_launches,
            "roofline": roofline, "clocks": clk.summary(), "fp32_peak_tflops_measured": fp32_peak,
            "hbm_peak_gbs": hbm_peak, "hbm_peak_source": peaks["source"],
            "trajectory_cost": float(out["trajectory_cost"])}
    ctx.close()
    if not args.no_large:
        if world == 1:
            line["large"] = measure_large("nn", models, costmap, cp, state, U, LARGE_ROLLOUTS, local_rank, fp32_peak)
            line["roofline_large"] = line["large"]["roofline"]
            from autorally_b200.scenarios import cost_params_for
            line["bf_large"] = measure_large("bf", models, costmap, cost_params_for(costmap, desired_speed=6.0), state, U, LARGE_ROLLOUTS,
                                             local_rank, fp32_peak, steps=3)
        else:
            line.update(run_multi_gpu(d, models, costmap, cp, state, U, fp32_peak))
        line["other_configs"] = measure_other_configs(models, costmap, cp, local_rank)
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(models, costmap, cp, state, U)
    if rank == 0:
        print(json.dumps(line))
    d.close()


if __name__ == "__main__":
    main()
