"""Host-side sharding arithmetic for the multi-GPU modes (nothing like it in the reference).

Rollout sharding: GPU g of G owns the global rollouts [begin, begin + count) with both multiples of 64
(BLOCKSIZE_WRX, PI/mppi_controller.cuh:58-60); bookkeeping and the Philox counter use the global index, so results do
not depend on G.  Controller sharding (batched MPC): disjoint controller ranges, no communication.
"""
from __future__ import annotations


def rollout_shard(rank: int, world: int, num_rollouts: int) -> tuple[int, int]:
    """(rollout_begin, rollout_count) of `rank`; chunks of 64 rollouts dealt as evenly as possible."""
    if num_rollouts <= 0 or num_rollouts % 64:
        raise ValueError("num_rollouts must be a positive multiple of 64")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    chunks = num_rollouts // 64
    if chunks < world:
        raise ValueError("fewer 64-rollout chunks (%d) than ranks (%d)" % (chunks, world))
    lo, hi = chunks * rank // world * 64, chunks * (rank + 1) // world * 64
    return lo, hi - lo


def controller_shard(rank: int, world: int, num_controllers: int) -> tuple[int, int]:
    """(first controller, count) of `rank` in batched-MPC mode."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    lo, hi = num_controllers * rank // world, num_controllers * (rank + 1) // world
    return lo, hi - lo


def pure_noise_threshold(num_rollouts: int) -> int:
    """Smallest global rollout index r with (double)r >= .99 * N (PI/mppi_controller.cu:141)."""
    lim = .99 * float(num_rollouts)
    r = int(lim)
    while float(r) < lim:
        r += 1
    return r
