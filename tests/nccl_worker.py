"""One rank of the multi-GPU parity check (launched by tests/test_multi_gpu_nccl.py and usable by hand):

    RANK=r WORLD_SIZE=G MPPI_ID_FILE=/tmp/id python tests/nccl_worker.py

Each rank owns GPU `RANK`, takes its shard of the rollouts, and runs mppi_compute_control_sharded (sampler or injected
noise -> rollouts -> local record -> ONE ncclAllGather -> finalize), then the same with the peer-memory exchange.  Every rank also computes the unsharded answer on
its own GPU and asserts the sharded result equals it (<= 1e-5 relative): results must not depend on the number of GPUs.
The NCCL unique id travels through a file, so no torch.distributed is involved.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    id_file = os.environ["MPPI_ID_FILE"]
    from autorally_b200.capi import MppiContext
    from autorally_b200.params import make_ellipse_costmap
    from autorally_b200.sharding import rollout_shard
    from tests.common import cost_params_for, make_context, straight_controls, top_state
    models = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
    costmap = make_ellipse_costmap()
    cp = cost_params_for(costmap)
    N, T = int(os.environ.get("MPPI_ROLLOUTS", "4096")), 100
    state, U = top_state(4.0), straight_controls(T)
    hist = np.array([0.1, 0.3, 0.11, 0.32], np.float32)
    if rank == 0:
        uid = MppiContext.comm_unique_id()
        with open(id_file + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            if time.time() - t0 > 120:
                raise TimeoutError("no NCCL id from rank 0")
            time.sleep(0.05)
        uid = open(id_file, "rb").read()
    lo, n = rollout_shard(rank, world, N)
    eps = np.random.default_rng(21).standard_normal((N, T, 2)).astype(np.float32)

    def rel(a, b):
        return float(np.max(np.abs(a - b) / (1 + np.abs(b))))

    with make_context("nn", models, costmap, cp, N, device=rank) as full, \
            make_context("nn", models, costmap, cp, N, rollout_begin=lo, rollout_count=n, device=rank) as ctx:
        ctx.comm_init(uid, rank, world)
        # (1) injected noise
        full.set_noise(eps)
        want = full.compute_control(state, U, hist)
        ctx.set_noise(eps[lo:lo + n])
        got = ctx.compute_control_sharded(state, U, hist)
        assert got["baseline"] == want["baseline"], (got["baseline"], want["baseline"])
        for k in ("U", "state_solution", "control_solution", "normalizer", "trajectory_cost"):
            assert rel(got[k], want[k]) < 1e-5, (k, rel(got[k], want[k]))
        # (2) Philox sampler: the counter uses the GLOBAL rollout index, so shards draw exactly the unsharded noise
        full.use_sampler(); ctx.use_sampler()
        full.seed(77, 3); ctx.seed(77, 3)
        want = full.compute_control(state, U, hist)
        got = ctx.compute_control_sharded(state, U, hist)
        for k in ("U", "state_solution", "normalizer"):
            assert rel(got[k], want[k]) < 1e-5, ("sampler", k, rel(got[k], want[k]))
        # (3) device-resident sharded stepping runs and is timed on the device
        ms = ctx.run_resident_sharded(5)
        assert ms > 0
        # (4) the same exchange without a collective: peer-memory stores fused into the weighting kernel, flag wait in
        # finalize (mppi_p2p_export / mppi_p2p_init; CUDA IPC handles travel through files, in rank order)
        mine = ctx.p2p_export(world)
        with open("%s.p2p.%d.tmp" % (id_file, rank), "wb") as f:
            f.write(mine)
        os.replace("%s.p2p.%d.tmp" % (id_file, rank), "%s.p2p.%d" % (id_file, rank))
        handles = b""
        for r in range(world):
            t0 = time.time()
            while not os.path.exists("%s.p2p.%d" % (id_file, r)):
                if time.time() - t0 > 120:
                    raise TimeoutError("no peer-memory handle from rank %d" % r)
                time.sleep(0.05)
            handles += open("%s.p2p.%d" % (id_file, r), "rb").read()
        ctx.p2p_init(handles, rank, world)
        full.set_noise(eps); ctx.set_noise(eps[lo:lo + n])
        for rep in range(3):   # both mailbox halves and the sequence numbers
            want = full.compute_control(state, U, hist)
            got = ctx.compute_control_sharded(state, U, hist)
            assert got["baseline"] == want["baseline"], ("p2p", got["baseline"], want["baseline"])
            for k in ("U", "state_solution", "control_solution", "normalizer", "trajectory_cost"):
                assert rel(got[k], want[k]) < 1e-5, ("p2p", rep, k, rel(got[k], want[k]))
        full.use_sampler(); ctx.use_sampler()
        full.seed(78, 5); ctx.seed(78, 5)
        want = full.compute_control(state, U, hist)
        got = ctx.compute_control_sharded(state, U, hist)
        for k in ("U", "state_solution", "normalizer"):
            assert rel(got[k], want[k]) < 1e-5, ("p2p sampler", k, rel(got[k], want[k]))
        ctx.run_resident_sharded(5)
        ms_p2p = ctx.run_resident_sharded(20)
        assert ms_p2p > 0
        print("rank %d/%d ok: shard [%d, %d), sharded step %.3f ms (NCCL all-gather), %.3f ms (peer-memory exchange)"
              % (rank, world, lo, lo + n, ms / 5, ms_p2p / 20))


if __name__ == "__main__":
    main()
