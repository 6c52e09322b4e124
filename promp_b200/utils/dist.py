"""torch.distributed plumbing: the path shards by task, the only data-path collective is the sum
all-reduce of the flat meta-gradient [P] (plus a packed vector of logged scalars)."""


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def allreduce_sum_(t):
    d = _dist()
    if d is not None:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def world_size():
    d = _dist()
    return d.get_world_size() if d is not None else 1


def rank():
    d = _dist()
    return d.get_rank() if d is not None else 0


def shard_tasks(all_tasks, rank, world):
    """Rank `rank` of `world` owns the contiguous slice of the global task list (every rank draws the same
    list from the same numpy seed, so results do not depend on the number of GPUs)."""
    n = len(all_tasks)
    assert n % world == 0, "global meta batch must be divisible by the number of ranks"
    per = n // world
    return list(all_tasks[rank * per:(rank + 1) * per])
