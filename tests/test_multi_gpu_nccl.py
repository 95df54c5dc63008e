"""GPU (>= 2 devices): rollouts sharded over G GPUs with the in-library NCCL exchange reproduce the single-GPU
controller.  One process per GPU; skipped on a single-GPU box (the protocol is also covered with G shards on one GPU
in test_parity_gpu.py and with gloo on CPU in test_multi_gpu_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_controller_matches_single_gpu(tmp_path, world):
    if _device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    id_file = str(tmp_path / "nccl_id")
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MPPI_ID_FILE=id_file)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "nccl_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for rank, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out[-3000:])
        assert "ok" in out
