"""Run a few device-resident meta-iterations (profiling target for ncu): python tools/run_iters.py [point|cheetah] [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    from promp_b200.utils import logger
    logger.set_quiet(True)
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'point']
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    np.random.seed(1)
    tr = bench.build_stack(wl, 'device')
    for i in range(n):
        tr.train_iteration(i, log=False)
    torch.cuda.synchronize()
    print('done', n)


if __name__ == '__main__':
    main()
