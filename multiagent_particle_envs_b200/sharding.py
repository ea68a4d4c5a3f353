"""Data-parallel sharding of a batch of worlds over ranks (one process per GPU).

Worlds are independent (no cross-world term anywhere in core.py:117-196 or the scenario
callbacks), so the step path needs NO collective: rank r simply owns the contiguous range
shard_range(n_env, r, world_size).  Philox streams are keyed by the *global* world index, hence
trajectories do not depend on the number of ranks.  The only exchange is the throughput counter
(`aggregate_counters`): one all-gather of (env_steps, seconds) per timing window.
"""


def shard_range(n_env, rank, world_size):
    """[start, stop) of the worlds owned by `rank`; the first n_env % world_size ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of size %d" % (rank, world_size))
    base, extra = divmod(int(n_env), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def aggregate_counters(env_steps, seconds, group=None):
    """All-gather each rank's (env_steps, seconds); returns (total env steps, max seconds,
    per-rank list).  Uses torch.distributed when initialised (NCCL on GPUs, gloo on CPU);
    degenerates to the local values in a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(env_steps), float(seconds), [(float(env_steps), float(seconds))]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([float(env_steps), float(seconds)], dtype=torch.float64, device=dev)
    world = dist.get_world_size(group)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    per_rank = [(float(g[0]), float(g[1])) for g in gathered]
    return sum(p[0] for p in per_rank), max(p[1] for p in per_rank), per_rank
