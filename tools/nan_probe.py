"""Investigate the late-training divergence of ProMP on PointEnv (profiles/r01_learning_curves.txt: LossAfter > LossBefore
from ~itr 1000, NaN at ~1400 with the reference's default hyper-parameters).

Runs the CUDA-graph Trainer (BASELINE.json configs[1], seed 1) and keeps, every iteration, a host copy of what the
optimisation step started from (theta, Adam m / v / step).  Dumps to gpurun_out/nan_probe_*.npz:
  * 'first_bad'  - the first iteration >= --after whose 5 Adam epochs INCREASE the meta objective by > --thresh
  * 'pre_nan'    - the last iteration with finite parameters before the first non-finite one
each with theta_before, adam_m, adam_v, adam_step, theta_after (device result) and both phases' obs / act / adv / mean /
log_std, so tests/test_nan_replay.py can replay exactly these inputs through oracle/tf_half.promp_optimize (float32 and
float64) on the CPU.  Usage: python tools/nan_probe.py [--iters 2000] [--after 600] [--thresh 1e-4] [--guard 0|1]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=2000)
    ap.add_argument('--after', type=int, default=600)
    ap.add_argument('--thresh', type=float, default=1e-4)
    ap.add_argument('--out', default='gpurun_out')
    ap.add_argument('--tasks', type=int, default=8, help="how many tasks' trajectories to keep in the dump")
    args = ap.parse_args()
    import torch
    import bench
    from promp_b200.utils import logger
    logger.set_quiet(True)
    np.random.seed(1)
    tr = bench.build_stack(bench.WORKLOADS['point'], 'numpy')
    step = tr.capture_graph(warmup=2, log=True)
    pol, algo = tr.policy, tr.algo
    opt = algo.optimizer
    os.makedirs(args.out, exist_ok=True)

    def snapshot():
        return dict(theta=pol.theta.cpu().numpy().copy(), m=opt.m.cpu().numpy().copy(), v=opt.v.cpu().numpy().copy(),
                    step=int(opt.step.item()))

    def dump(name, itr, before, phases, kv):
        d = dict(itr=itr, theta_before=before['theta'], adam_m=before['m'], adam_v=before['v'], adam_step=before['step'],
                 theta_after=pol.theta.cpu().numpy(), loss_before=kv['LossBefore'], loss_after=kv['LossAfter'],
                 kl_inner=kv['KLInner'])
        for s, ph in enumerate(phases):
            for k in ('obs', 'act', 'adv', 'mean', 'log_std'):
                d['p%d_%s' % (s, k)] = getattr(ph, k).cpu().numpy()
        np.savez_compressed(os.path.join(args.out, 'nan_probe_%s.npz' % name), **d)
        print('dumped', name, 'itr', itr, 'loss %.6g -> %.6g' % (kv['LossBefore'], kv['LossAfter']), flush=True)

    have_bad = False
    curve = []
    prev_dump = None
    for itr in range(args.iters):
        before = snapshot()
        phases = step(itr)
        kv = dict(logger.getkvs())
        finite = bool(torch.isfinite(pol.theta).all()) and np.isfinite(kv['LossAfter'])
        curve.append((itr, kv['Step_0-AverageReturn'], kv['Step_1-AverageReturn'], kv['LossBefore'], kv['LossAfter'],
                      kv['KLInner'], kv['Step_1-AveragePolicyStd'], float(np.abs(before['theta']).max())))
        if itr % 100 == 0 or not finite:
            print('itr %4d ret %.3f/%.3f loss %.5g -> %.5g klin %.4g std %.3f |th|max %.3g' % curve[-1], flush=True)
        if not finite:
            if prev_dump is not None:
                np.savez_compressed(os.path.join(args.out, 'nan_probe_pre_nan.npz'), **prev_dump)
                print('dumped pre_nan (itr %d)' % prev_dump['itr'])
            d = dict(itr=itr, theta_before=before['theta'], adam_m=before['m'], adam_v=before['v'], adam_step=before['step'])
            for s, ph in enumerate(phases):
                for k in ('obs', 'act', 'adv', 'mean', 'log_std'):
                    d['p%d_%s' % (s, k)] = getattr(ph, k).cpu().numpy()
            np.savez_compressed(os.path.join(args.out, 'nan_probe_nan_itr.npz'), **d)
            print('first non-finite iteration:', itr)
            break
        if itr >= args.after and not have_bad and kv['LossAfter'] > kv['LossBefore'] + args.thresh:
            dump('first_bad', itr, before, phases, kv)
            have_bad = True
        if itr >= args.after and itr % 1 == 0:
            # keep the previous iteration's full inputs (host copy) in case the next one goes non-finite
            prev_dump = dict(itr=itr, theta_before=before['theta'], adam_m=before['m'], adam_v=before['v'],
                             adam_step=before['step'], theta_after=pol.theta.cpu().numpy(),
                             loss_before=kv['LossBefore'], loss_after=kv['LossAfter'], kl_inner=kv['KLInner'])
            for s, ph in enumerate(phases):
                for k in ('obs', 'act', 'adv', 'mean', 'log_std'):
                    prev_dump['p%d_%s' % (s, k)] = getattr(ph, k).cpu().numpy()
    np.save(os.path.join(args.out, 'nan_probe_curve.npy'), np.asarray(curve))
    print('done: %d iterations, finite=%s' % (len(curve), finite))


if __name__ == '__main__':
    main()
