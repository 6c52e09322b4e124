"""cProfile of the eager (non-graph) Trainer.train_iteration on the GPU box: where the host time of an iteration goes."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from promp_b200.utils import logger
    logger.set_quiet(True)
    np.random.seed(1)
    for log in (False, True):
        tr = bench.build_stack(bench.WORKLOADS['point'], 'numpy')
        for i in range(5):
            tr.train_iteration(i, log=log)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for i in range(20):
            tr.train_iteration(i, log=log)
        torch.cuda.synchronize()
        print('eager log=%s: %.3f ms / iteration' % (log, (time.perf_counter() - t) / 20 * 1e3))
        pr = cProfile.Profile()
        pr.enable()
        for i in range(20):
            tr.train_iteration(i, log=log)
        torch.cuda.synchronize()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats('tottime').print_stats(22)


if __name__ == '__main__' and len(sys.argv) == 1:
    main()


def after_graph():
    """bench.py order: a graph-mode Trainer.train() first, then an eager Trainer with logging."""
    import torch
    import bench
    from promp_b200.utils import logger
    logger.set_quiet(True)
    np.random.seed(1)
    tr = bench.build_stack(bench.WORKLOADS['point'], 'numpy')
    tr.start_itr, tr.n_itr = 0, 10
    tr.train()
    torch.cuda.synchronize()
    tr2 = bench.build_stack(bench.WORKLOADS['point'], 'numpy', use_cuda_graph=False)
    for i in range(3):
        tr2.train_iteration(i, log=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(10):
        tr2.train_iteration(i, log=True)
    torch.cuda.synchronize()
    print('eager log=True after a graph-mode train(): %.3f ms / iteration' % ((time.perf_counter() - t) / 10 * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    for i in range(10):
        tr2.train_iteration(i, log=True)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(14)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'after_graph':
    after_graph()
