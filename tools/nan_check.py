import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from promp_b200.utils import logger
logger.set_quiet(True)
np.random.seed(1)
tr = bench.build_stack(bench.WORKLOADS['point'], 'numpy')
step = tr.capture_graph(warmup=2, log=True)
pol, algo = tr.policy, tr.algo
for itr in range(1600):
    phases = step(itr)
    if itr >= 1300 and (itr % 10 == 0):
        kv = logger.getkvs()
        th = pol.theta
        tt = pol.theta_tasks
        g = algo.optimizer.last_grad
        r = phases[1]
        print(itr, 'ret %.3f/%.3f loss %.4g->%.4g klin %.4g std %.3f |th|max %.3g |th_t-th|max %.3g gnorm %.3g act|max| %.3g advmax %.3g' % (
            kv['Step_0-AverageReturn'], kv['Step_1-AverageReturn'], kv['LossBefore'], kv['LossAfter'], kv['KLInner'], kv['Step_1-AveragePolicyStd'],
            float(th.abs().max()), float((tt - th).abs().max()), float(g.norm()), float(r.act.abs().max()), float(phases[0].adv.abs().max())))
        if not np.isfinite(kv['LossAfter']): break
