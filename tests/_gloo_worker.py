"""world_size-2 gloo worker for tests/test_host_logic.py (CPU): task sharding + the meta-gradient all-reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promp_b200.utils.dist import allreduce_sum_, world_size, rank, shard_tasks  # noqa: E402


def main():
    dist.init_process_group('gloo')
    r, w = rank(), world_size()
    assert w == 2
    np.random.seed(7)                                   # same seed on every rank -> same global task list
    corners = [(-2, -2), (2, -2), (-2, 2), (2, 2)]
    all_tasks = [corners[i] for i in np.random.choice(range(4), size=8)]
    mine = shard_tasks(all_tasks, r, w)
    assert mine == all_tasks[r * 4:(r + 1) * 4]
    # per-rank partial meta-gradient: (1/M_global) * sum over local tasks; all-reduce(SUM) = global mean
    P, M_local = 4484, 4
    g = torch.Generator().manual_seed(100)
    per_task = torch.randn(w * M_local, P, generator=g)            # identical on both ranks
    local = per_task[r * M_local:(r + 1) * M_local].sum(0) / (w * M_local)
    allreduce_sum_(local)
    assert torch.allclose(local, per_task.mean(0), atol=1e-6)
    # packed scalar stats
    vec = torch.tensor([float(r + 1), 2.0])
    allreduce_sum_(vec)
    assert vec.tolist() == [3.0, 4.0]
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % r)


if __name__ == '__main__':
    main()
