"""Train ProMP on MetaPointEnvCorner for a few hundred meta-iterations (CUDA-graph mode) and print the learning curve."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    from promp_b200.utils import logger
    logger.set_quiet(True)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    wl = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else 'point']
    np.random.seed(1)
    tr = bench.build_stack(wl, 'numpy')
    step = tr.capture_graph(warmup=2, log=True)
    t0 = time.time()
    for itr in range(n):
        step(itr)
        if itr % max(n // 10, 1) == 0 or itr == n - 1:
            kv = logger.getkvs()
            print('itr %4d  pre-update return %8.3f  post-update return %8.3f  loss %.5f -> %.5f  KLInner %.5f  std %.3f' % (
                itr, kv['Step_0-AverageReturn'], kv['Step_1-AverageReturn'], kv['LossBefore'], kv['LossAfter'], kv['KLInner'],
                kv['Step_1-AveragePolicyStd']))
    torch.cuda.synchronize()
    print('%d meta-iterations in %.2f s (%.1f M env-steps)' % (n, time.time() - t0, n * wl['M'] * wl['E'] * wl['H'] * 2 / 1e6))


if __name__ == '__main__':
    main()
