"""CPU, world_size 2, gloo: the N > 1 host logic -- rollout sharding with GLOBAL bookkeeping indices, the exchange of
one (3 + 2T)-float record per rank, and the log-sum-exp combine -- against the unsharded oracle.  The CUDA side of
the same protocol (mppi_shard_begin / finish) is covered on the GPU by
tests/test_parity_gpu.py::test_sharded_rollouts_reproduce_single_gpu_answer."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from autorally_b200.sharding import controller_shard, pure_noise_threshold, rollout_shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def combine_records(records, gamma):
    """What finalize_kernel does with the gathered records [G][3 + 2T] (autorally_b200/csrc/weighting.cuh)."""
    records = np.asarray(records, np.float64)
    b = records[:, 0].min()
    s = np.exp(-gamma * (records[:, 0] - b))
    Z = (s * records[:, 1]).sum()
    Q = (s * s * records[:, 2]).sum()
    W = (s[:, None] * records[:, 3:]).sum(0)
    return b, Z, Q / Z, (W / Z).reshape(-1, 2)


def test_shard_ranges_tile_the_rollouts():
    for N, G in [(1920, 2), (1920, 4), (1 << 20, 8), (1000000 // 64 * 64, 8), (2560, 3), (64 * 8, 8)]:
        pos = 0
        for g in range(G):
            lo, n = rollout_shard(g, G, N)
            assert lo == pos and n > 0 and lo % 64 == 0 and n % 64 == 0
            pos += n
        assert pos == N
    with pytest.raises(ValueError):
        rollout_shard(0, 4, 128)
    assert [controller_shard(g, 8, 4096) for g in (0, 7)] == [(0, 512), (3584, 512)]
    assert pure_noise_threshold(1920) == 1901 and pure_noise_threshold(2560) == 2535 and pure_noise_threshold(256) == 254


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["MPPI_ROOT"])
from autorally_b200.params import make_ellipse_costmap
from autorally_b200.sharding import rollout_shard
from tests.common import cost_params_for, make_oracle, straight_controls, top_state
from tests.test_multi_gpu_gloo import combine_records

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MPPI_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
N, T, gamma = 512, 40, 0.15
models = np.load(os.path.join(os.environ["MPPI_ROOT"], "tests", "golden", "ref_models.npz"))
costmap = make_ellipse_costmap(pixels_per_meter=5.0)
cp = cost_params_for(costmap)
o = make_oracle("nn", models, costmap, cp)
eps = np.random.default_rng(3).standard_normal((N, T, 2)).astype(np.float32)   # same seed on every rank
state, U, nu = top_state(4.0), straight_controls(T), [0.275, 0.3]
lo, n = rollout_shard(rank, world, N)
V, costs, crash, _ = o.rollouts(state, U, nu, eps[lo:lo + n], n_global=N, r_begin=lo)
rec = torch.from_numpy(o.shard_partials(costs, V, gamma).astype(np.float32))
gathered = [torch.empty_like(rec) for _ in range(world)]
dist.all_gather(gathered, rec)                                    # the ONE exchange: 3 + 2T floats per rank
b, Z, tc, Unew = combine_records(torch.stack(gathered).numpy(), gamma)
# unsharded answer
Vf, cf, _, _ = o.rollouts(state, U, nu, eps, n_global=N, r_begin=0)
w, Uref, stats = o.weighting(cf, Vf, gamma)
np.testing.assert_array_equal(V, Vf[lo:lo + n])                   # global-index bookkeeping: shards see the same branches
np.testing.assert_array_equal(costs, cf[lo:lo + n])
assert b == stats[0]
np.testing.assert_allclose(Z, stats[1], rtol=1e-6)
np.testing.assert_allclose(tc, stats[2], rtol=1e-6)
np.testing.assert_allclose(Unew, Uref, rtol=2e-6, atol=1e-7)
if rank == 0:
    assert np.array_equal(V[0], U)                                # the noise-free rollout lives on rank 0
if rank == world - 1:
    assert np.array_equal(V[-1][1:], (eps[-1][1:] * np.asarray(nu, np.float32)).astype(np.float32))  # pure-noise tail
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
'''


def test_two_rank_gloo_exchange_reproduces_the_unsharded_controller(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MPPI_PORT=str(port), MPPI_ROOT=ROOT, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for rank, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out[-3000:])
        assert "rank %d ok" % rank in out
