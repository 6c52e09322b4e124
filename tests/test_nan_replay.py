"""The late-training divergence of ProMP on PointEnv is the ALGORITHM's, not the kernels'.

Observed (profiles/r02_nan_probe.txt): with the reference's default hyper-parameters (run_scripts/pro-mp_run_point_mass.py:
lr 1e-3, 5 Adam epochs per iteration on the same data, clip 0.3, KL penalty 5e-4, n_itr 1001) the learned policy std
shrinks to ~0.1 after ~1000 iterations; from ~itr 600 on single iterations INCREASE the meta objective (LossAfter >
LossBefore: Adam's momentum + 5 fixed-size steps overshoot once exp(logp_new - logp_old) is this steep), and around itr
1400-1750 one iteration's likelihood ratios overflow float32 (outer KL ~1e9 in float64) and the parameters become NaN.

This test re-creates that run on the device (BASELINE.json configs[1], seed 1) and replays the interesting iterations'
exact inputs (theta, Adam slots, both phases) through oracle/tf_half.promp_optimize - the restatement that
tests/test_tf_golden.py pins to the reference's unmodified graph code:
  * first iteration (>= 500) whose 5 epochs increase the objective: the float32 AND float64 oracle increase it by the
    same amount and produce the same parameter update as the device (1e-3);
  * first non-finite iteration (if one occurs within 2000 iterations): the float32 oracle is non-finite on the same
    inputs too (the float64 oracle survives with an absurd KL), i.e. TF1-float32 would have diverged identically."""
import numpy as np
import pytest


def _replay(torch, th, before, phases, dt):
    N = phases[0].N
    data = [dict(obs=ph.obs.cpu().to(dt), act=ph.act.cpu().to(dt), adv=ph.adv.cpu().to(dt), mean=ph.mean.cpu().to(dt),
                 log_std=ph.log_std.cpu().to(dt)[:, None, :].expand(-1, N, -1)) for ph in phases]
    theta = torch.tensor(before['theta'], dtype=dt)
    adam = th.TF1Adam(theta.numel(), dtype=dt)
    adam.m, adam.v, adam.t = torch.tensor(before['m'], dtype=dt), torch.tensor(before['v'], dtype=dt), before['step']
    new, st = th.promp_optimize(theta, data, (2, 2, (64, 64)), adam, 0.1, 0.3, [5e-4], 5)
    return new.numpy(), st


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_late_training_divergence_is_reproduced_by_the_oracle():
    import torch
    from promp_b200 import _lib
    _lib.require_cuda()
    import bench
    from oracle import tf_half as th
    from promp_b200.utils import logger
    logger.set_quiet(True)
    np.random.seed(1)
    tr = bench.build_stack(bench.WORKLOADS['point'], 'numpy')
    step = tr.capture_graph(warmup=2, log=True)
    pol, opt = tr.policy, tr.algo.optimizer
    checked_bad = False
    nan_itr = None
    for itr in range(2000):
        before = dict(theta=pol.theta.cpu().numpy().copy(), m=opt.m.cpu().numpy().copy(), v=opt.v.cpu().numpy().copy(),
                      step=int(opt.step.item()))
        phases = step(itr)
        kv = logger.getkvs()
        lb, la = float(kv['LossBefore']), float(kv['LossAfter'])
        finite = bool(torch.isfinite(pol.theta).all()) and np.isfinite(la)
        if not finite:
            nan_itr = itr
            _, st32 = _replay(torch, th, before, phases, torch.float32)
            assert not np.isfinite(st32['loss_after']), \
                "the device went non-finite at itr %d but the float32 oracle did not (%r)" % (itr, st32['loss_after'])
            _, st64 = _replay(torch, th, before, phases, torch.float64)
            print("itr %d: device and float32 oracle non-finite; float64 oracle loss_after %.4g outer KL %.4g"
                  % (itr, st64['loss_after'], st64['outer_kl']))
            break
        if not checked_bad and itr >= 500 and la > lb + 1e-4:
            got = pol.theta.cpu().numpy().astype(np.float64)
            for dt in (torch.float32, torch.float64):
                new, st = _replay(torch, th, before, phases, dt)
                assert st['loss_after'] > st['loss_before'] + 1e-4          # the oracle's 5 epochs increase the objective too
                assert abs(st['loss_after'] - la) < 1e-3 * max(abs(la), 1e-3), (st['loss_after'], la)
                upd, upd_want = got - before['theta'], new.astype(np.float64) - before['theta']
                assert np.linalg.norm(upd - upd_want) / np.linalg.norm(upd_want) < 1e-3
            checked_bad = True
            print("itr %d: LossAfter %.6g > LossBefore %.3g reproduced by the float32 and float64 oracle" % (itr, la, lb))
    assert checked_bad, "no objective-increasing iteration found: the premise of this test changed"
    print("first non-finite iteration:", nan_itr)
