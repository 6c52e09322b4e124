"""2-rank GPU worker: promp_allreduce_p2p vs NCCL, eager and inside a CUDA graph."""
import datetime
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promp_b200.utils.dist import enable_p2p_allreduce, allreduce_sum_  # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()),
                            timeout=datetime.timedelta(seconds=60))
    g = torch.Generator(device='cuda').manual_seed(100 + rank)
    x = torch.randn(4484, generator=g, device='cuda')
    want = x.clone()
    dist.all_reduce(want)
    comm = enable_p2p_allreduce(8192)
    for i in range(5):                      # eager, several epochs (both slots)
        y = (x * (i + 1)).clone()
        allreduce_sum_(y)
        assert torch.allclose(y, want * (i + 1), rtol=1e-6, atol=1e-6), i
    # rank-ordered sum: bitwise identical on all ranks
    y = x.clone()
    allreduce_sum_(y)
    both = [torch.empty_like(y) for _ in range(dist.get_world_size())]
    dist.all_gather(both, y)
    assert all(torch.equal(both[0], b) for b in both)
    # inside a CUDA graph, replayed
    static = x.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            allreduce_sum_(static.copy_(x))
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static.copy_(x)
        allreduce_sum_(static)
    for _ in range(4):
        graph.replay()
        torch.cuda.synchronize()
        assert torch.allclose(static, want, rtol=1e-6, atol=1e-6)
    comm.check()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d p2p ok" % rank)


if __name__ == '__main__':
    main()
