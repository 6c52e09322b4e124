"""ConjugateGradientOptimizer with FiniteDifferenceHvp
(ref: meta_policy_search/optimizers/conjugate_gradient_optimizer.py:8-354).

As in the reference the CG iteration and the backtracking line search run on the host on flat float32
numpy vectors; every loss / KL / gradient evaluation is a device pass of the meta objective
(MAMLAlgo._meta_pass) at a candidate theta.  Scalars and gradients are all-reduced over ranks before
any host decision so all ranks take identical steps.
"""
import numpy as np

from promp_b200.utils import logger
from promp_b200.utils.dist import allreduce_sum_


def conjugate_gradients(f_Ax, b, cg_iters=10, residual_tol=1e-10):
    """(:325-354, Demmel p 312)."""
    p = b.copy()
    r = b.copy()
    x = np.zeros_like(b, dtype=np.float32)
    rdotr = r.dot(r)
    for _ in range(cg_iters):
        z = f_Ax(p)
        v = rdotr / p.dot(z)
        x += v * p
        r -= v * z
        newrdotr = r.dot(r)
        mu = newrdotr / rdotr
        p = r + mu * p
        rdotr = newrdotr
        if rdotr < residual_tol:
            break
    return x


class ConjugateGradientOptimizer(object):
    def __init__(self, cg_iters=10, reg_coeff=0, subsample_factor=1., backtrack_ratio=0.8, max_backtracks=15,
                 accept_violation=False, base_eps=1e-5, symmetric=True):
        self._cg_iters, self._reg_coeff = cg_iters, reg_coeff
        self._backtrack_ratio, self._max_backtracks = backtrack_ratio, max_backtracks
        self._accept_violation = accept_violation
        self.base_eps = np.float32(base_eps)
        self.symmetric = symmetric
        self._max_constraint_val = None
        self._algo = None

    def build(self, algo, max_constraint_val):
        self._algo, self._max_constraint_val = algo, max_constraint_val

    # -- evaluations at a flat parameter vector (host float32) ------------------------------------
    def _theta(self, flat):
        import torch
        return torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float32)).to(self._algo.policy.device)

    def loss(self, flat, phases):
        return self._algo.eval_scalars(self._theta(flat), phases)[0]

    def constraint_val(self, flat, phases):
        return self._algo.eval_scalars(self._theta(flat), phases)[1]

    def loss_and_constraint(self, flat, phases):
        return self._algo.eval_scalars(self._theta(flat), phases)

    def gradient(self, flat, phases):
        return self._algo.eval_gradient(self._theta(flat), phases, 'loss')

    def constraint_gradient(self, flat, phases):
        return self._algo.eval_gradient(self._theta(flat), phases, 'kl')

    def Hx(self, flat, phases, x):
        """FiniteDifferenceHvp.Hx (:59-89)."""
        eps = self.base_eps
        gp = self.constraint_gradient(flat + eps * x, phases)
        if self.symmetric:
            gm = self.constraint_gradient(flat - eps * x, phases)
            return (gp - gm) / (2 * eps)
        return (gp - self.constraint_gradient(flat, phases)) / eps

    def optimize(self, phases):
        """(:239-307).  Updates policy.theta in place (or restores it when the step is rejected)."""
        policy = self._algo.policy
        prev = policy.theta.detach().cpu().numpy().astype(np.float32)
        logger.log("Start CG optimization")
        loss_before = self.loss(prev, phases)
        gradient = self.gradient(prev, phases)
        Hx = lambda x: self.Hx(prev, phases, x) + self._reg_coeff * x
        descent_direction = conjugate_gradients(Hx, gradient, cg_iters=self._cg_iters)
        initial_step_size = np.sqrt(2.0 * self._max_constraint_val *
                                    (1. / (descent_direction.dot(Hx(descent_direction)) + 1e-8)))
        self.last = dict(loss_before=loss_before, gradient=gradient, direction=descent_direction,
                         init_step=initial_step_size)
        if np.isnan(initial_step_size):
            logger.log("Initial step size is NaN! Rejecting the step!")
            return
        initial_descent_step = initial_step_size * descent_direction
        loss, constraint_val, n_iter, violated = 0, 0, 0, False
        cur = prev
        for n_iter, ratio in enumerate(self._backtrack_ratio ** np.arange(self._max_backtracks)):
            cur = (prev - ratio * initial_descent_step).astype(np.float32)
            loss, constraint_val = self.loss_and_constraint(cur, phases)
            if loss < loss_before and constraint_val <= self._max_constraint_val:
                break
        if np.isnan(loss) or np.isnan(constraint_val) or loss >= loss_before or \
                constraint_val >= self._max_constraint_val:
            violated = True
        if violated and not self._accept_violation:
            logger.log("Line search condition violated. Rejecting the step!")
            cur = prev
        policy.set_params(cur)
        self.last.update(loss=loss, kl=constraint_val, backtracks=n_iter, rejected=violated)
        logger.log("backtrack iters: %d" % n_iter)
