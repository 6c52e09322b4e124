"""TRPOMAML (ref: meta_policy_search/meta_algos/trpo_maml.py:8-192) on the GPU."""
import numpy as np

from promp_b200 import _lib
from promp_b200.meta_algos.base import MAMLAlgo
from promp_b200.optimizers.conjugate_gradient_optimizer import ConjugateGradientOptimizer
from promp_b200.utils import logger
from promp_b200.utils.dist import allreduce_sum_, world_size


class TRPOMAML(MAMLAlgo):
    """Same constructor arguments as the reference (trpo_maml.py:23-47)."""

    def __init__(self, *args, name="trpo_maml", step_size=0.01, inner_type='likelihood_ratio', exploration=False,
                 **kwargs):
        super(TRPOMAML, self).__init__(*args, **kwargs)
        assert inner_type in ["log_likelihood", "likelihood_ratio", "dice"]
        if inner_type == 'dice':
            raise NotImplementedError("inner_type='dice' (reference raises NotImplementedError too, trpo_maml.py:64)")
        if exploration:
            raise NotImplementedError("E-MAML (exploration=True) is a 'next' row (SURVEY.md section 8f item 1)")
        self.step_size = step_size
        self.inner_type = inner_type
        self.name = name
        self._optimization_keys = ['observations', 'actions', 'advantages', 'agent_infos']
        self.exploration = exploration
        self.inner_obj_kind = _lib.OBJ_RATIO if inner_type == 'likelihood_ratio' else _lib.OBJ_LOGLIK
        self.optimizer = ConjugateGradientOptimizer()
        self.optimizer.build(self, step_size)

    # meta objective = mean_i -mean(ratio*adv) (:135,152); constraint = mean_i mean KL(old || theta_i') (:133,149)
    def eval_scalars(self, theta, phases):
        import torch
        res = self._meta_pass(theta, phases, _lib.OBJ_RATIO, 0.0, [0.0] * self.num_inner_grad_steps, want_grad=False)
        vec = torch.stack([res['surr'].sum(), res['outer_kl'].sum()]) / (self.meta_batch_size * world_size())
        allreduce_sum_(vec)
        host = vec.cpu().numpy()
        return float(host[0]), float(host[1])

    def eval_gradient(self, theta, phases, which):
        zeros = [0.0] * self.num_inner_grad_steps
        if which == 'loss':
            res = self._meta_pass(theta, phases, _lib.OBJ_RATIO, 0.0, zeros, want_grad=True)
        else:
            res = self._meta_pass(theta, phases, _lib.OBJ_NONE, 0.0, zeros, want_grad=True, outer_kl_coeff=1.0)
        allreduce_sum_(res['grad'])
        return res['grad'].cpu().numpy().astype(np.float32)

    def optimize_policy(self, all_samples_data, log=True):
        """trpo_maml.py:161-192."""
        assert len(all_samples_data) == self.num_inner_grad_steps + 1
        phases = [self._phase_of(s) for s in all_samples_data]
        theta = self.policy.theta
        logger.log("Computing KL before")
        loss_before, mean_kl_before = self.eval_scalars(theta, phases)
        logger.log("Optimizing")
        self.optimizer.optimize(phases)
        loss_after, mean_kl = self.eval_scalars(self.policy.theta, phases)
        self.last_stats = dict(loss_before=loss_before, loss_after=loss_after, kl_before=mean_kl_before, kl=mean_kl)
        if log:
            logger.logkv('MeanKLBefore', mean_kl_before)
            logger.logkv('MeanKL', mean_kl)
            logger.logkv('LossBefore', loss_before)
            logger.logkv('LossAfter', loss_after)
            logger.logkv('dLoss', loss_before - loss_after)
