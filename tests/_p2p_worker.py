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
    # ---- fused task-mean + all-reduce + TF1 Adam (promp_meta_update): vs NCCL + the Adam formula, eager and in a graph
    from promp_b200 import _lib
    import math
    W, M, P = dist.get_world_size(), 7, 4484
    v = torch.randn(M, P, generator=g, device='cuda')
    want_g = v.sum(0) / (M * W)
    dist.all_reduce(want_g)
    theta0 = torch.linspace(-1, 1, P, device='cuda')
    theta, m_, v_ = theta0.clone(), torch.zeros(P, device='cuda'), torch.zeros(P, device='cuda')
    step = torch.zeros(1, dtype=torch.int32, device='cuda')
    ticket = torch.zeros(1, dtype=torch.int32, device='cuda')
    gout = torch.empty(P, device='cuda')

    def update():
        _lib.call('promp_meta_update', M, P, _lib.ptr(v), 1.0 / (M * W), _lib.ptr(gout), _lib.ptr(theta), _lib.ptr(m_), _lib.ptr(v_),
                  _lib.ptr(step), 1e-3, 0.9, 0.999, 1e-8, comm.world, comm.rank, comm.cap, _lib.ptr(comm.peers), _lib.ptr(comm.epoch),
                  _lib.ptr(comm.error), _lib.ptr(ticket), _lib.stream())
    rm, rv, rt = torch.zeros(P, device='cuda'), torch.zeros(P, device='cuda'), theta0.clone()
    for t in range(1, 4):
        update()
        rm = 0.9 * rm + 0.1 * want_g
        rv = 0.999 * rv + 0.001 * want_g * want_g
        rt = rt - 1e-3 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * rm / (rv.sqrt() + 1e-8)
        assert torch.allclose(gout, want_g, rtol=1e-5, atol=1e-7), t
        assert torch.allclose(theta, rt, rtol=0, atol=2e-6), (t, float((theta - rt).abs().max()))
    assert int(step.item()) == 3 and int(ticket.item()) == 0
    both = [torch.empty_like(theta) for _ in range(W)]
    dist.all_gather(both, theta)
    assert all(torch.equal(both[0], b) for b in both)            # replicas stay bitwise identical
    dist.barrier()
    graph2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph2):
        update()
    for t in range(4, 7):
        graph2.replay()
        torch.cuda.synchronize()
        assert torch.allclose(gout, want_g, rtol=1e-5, atol=1e-7)
    assert int(step.item()) == 6
    # ---- fused loss terms: [loss, inner kl, outer kl] means over all ranks' tasks + KL penalty
    S, Mt = 2, 5
    st = torch.rand(S, Mt, 4, generator=g, device='cuda')
    coeff = torch.tensor([5e-4], device='cuda')
    out = torch.empty(3, device='cuda')
    _lib.call('promp_meta_loss_terms_p2p', S, Mt, _lib.ptr(st), 1.0 / (Mt * W), _lib.ptr(coeff), 3, _lib.ptr(out), comm.world,
              comm.rank, comm.cap, _lib.ptr(comm.peers), _lib.ptr(comm.epoch), _lib.ptr(comm.error), _lib.stream())
    loc = torch.stack([st[1, :, 0].sum(), st[0, :, 1].sum(), st[1, :, 1].sum()]) / (Mt * W)
    dist.all_reduce(loc)
    loc[0] += 5e-4 * loc[1]
    assert torch.allclose(out, loc, rtol=1e-5, atol=1e-7), (out, loc)
    # the one-shot all-reduce still interleaves with the fused kernels (shared epoch counter)
    y = x.clone()
    allreduce_sum_(y)
    assert torch.allclose(y, want, rtol=1e-6, atol=1e-6)
    comm.check()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d p2p ok" % rank)


if __name__ == '__main__':
    main()
