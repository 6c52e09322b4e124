"""ConjugateGradientOptimizer with FiniteDifferenceHvp, device-resident
(ref: meta_policy_search/optimizers/conjugate_gradient_optimizer.py:8-354).

The reference keeps p, r, x, z as host numpy vectors and pays one `sess.run` + `set_params` round trip for each of the
~1 + 22 gradient and <= 30 scalar evaluations of a TRPO step.  Here every vector lives on the device:

    g  = d loss / d theta                                        1 meta-gradient pass
    CG: 10 x [ theta +- eps p -> 2 KL-gradient passes -> promp_cg_step ]   (Hx(p) fused into the CG update)
    step = sqrt(2 delta / (x.Hx(x) + 1e-8)) x                    2 passes + promp_trpo_step
    line search: candidates theta - 0.8^k step are evaluated speculatively in groups of `group` (default 4) and
                 promp_trpo_select applies the reference's accept / violate / restore rule on the device

so the host reads back ONE small vector per group (normally a single read per TRPO step) and the whole sequence up to the
first verdict is CUDA-graph capturable (TRPOMAML.optimize_phases).  Gradients and scalars are all-reduced over ranks on the
device before they are used, so every rank takes the identical step.
"""
import numpy as np

from promp_b200 import _lib
from promp_b200.utils import logger


class ConjugateGradientOptimizer(object):
    def __init__(self, cg_iters=10, reg_coeff=0, subsample_factor=1., backtrack_ratio=0.8, max_backtracks=15,
                 accept_violation=False, base_eps=1e-5, symmetric=True, residual_tol=1e-10, group=4):
        if accept_violation:
            raise NotImplementedError("accept_violation=True is not implemented (no shipped reference config uses it)")
        self._cg_iters, self._reg_coeff = int(cg_iters), float(reg_coeff)
        self._backtrack_ratio, self._max_backtracks = float(backtrack_ratio), int(max_backtracks)
        self.base_eps = np.float32(base_eps)
        self.symmetric = symmetric
        self._residual_tol = float(residual_tol)
        self._group = int(group)
        self._max_constraint_val = None
        self._algo = None
        self._buf = None
        self.last = {}

    def build(self, algo, max_constraint_val):
        self._algo, self._max_constraint_val = algo, float(max_constraint_val)

    # ------------------------------------------------------------------------------------------------ buffers
    def _buffers(self):
        import torch
        algo = self._algo
        P, dev = algo.policy.num_params, algo.policy.device
        T = algo.num_inner_grad_steps + 2
        if self._buf is None or self._buf['P'] != P:
            f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
            self._buf = dict(P=P, T=T, theta_prev=f(P), th_a=f(P), th_b=f(P), p=f(P), r=f(P), x=f(P), step=f(P),
                             scal=torch.zeros(4, dtype=torch.float32, device=dev), base=f(T), cands=f(self._group, P),
                             terms=f(self._group, T), result=torch.zeros(8, dtype=torch.float32, device=dev))
        return self._buf

    def _axpy(self, a, x, y, out):
        _lib.call('promp_vec_axpy', x.numel(), float(a), _lib.ptr(x), _lib.ptr(y), _lib.ptr(out), _lib.stream())
        return out

    def _kl_grad_pair(self, theta, direction, phases):
        """d KL / d theta at theta + eps*direction and theta - eps*direction (FiniteDifferenceHvp.Hx, :59-89); the
        non-symmetric variant differences against theta itself.  Returns (grad_plus, grad_minus, divisor)."""
        b, eps, algo = self._buffers(), float(self.base_eps), self._algo
        gp = algo.eval_gradient_dev(self._axpy(eps, direction, theta, b['th_a']), phases, 'kl')
        if self.symmetric:
            gm = algo.eval_gradient_dev(self._axpy(-eps, direction, theta, b['th_b']), phases, 'kl')
            return gp, gm, float(np.float32(2.0) * self.base_eps)
        return gp, algo.eval_gradient_dev(theta, phases, 'kl'), eps

    # ------------------------------------------------------------------------------------------------ pieces
    def Hx(self, flat, phases, x):
        """Finite-difference Hessian-vector product as a host vector (diagnostics / tests)."""
        import torch
        dev = self._algo.policy.device
        theta = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float32)).to(dev)
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
        gp, gm, div = self._kl_grad_pair(theta, xd, phases)
        return ((gp - gm) / div).cpu().numpy() + np.float32(self._reg_coeff) * np.asarray(x, dtype=np.float32)

    def descent_direction(self, theta, phases, g):
        """x ~ H^-1 g by `cg_iters` conjugate-gradient iterations (:325-354), all on the device.  Returns x (device)."""
        b = self._buffers()
        n = b['P']
        _lib.call('promp_cg_init', n, _lib.ptr(g), _lib.ptr(b['p']), _lib.ptr(b['r']), _lib.ptr(b['x']), _lib.ptr(b['scal']),
                  _lib.stream())
        for _ in range(self._cg_iters):
            gp, gm, div = self._kl_grad_pair(theta, b['p'], phases)
            _lib.call('promp_cg_step', n, _lib.ptr(gp), _lib.ptr(gm), div, self._reg_coeff, _lib.ptr(b['p']), _lib.ptr(b['r']),
                      _lib.ptr(b['x']), _lib.ptr(b['scal']), self._residual_tol, _lib.stream())
        return b['x']

    def _evaluate_group(self, phases, k0):
        """Candidates k0..k0+K-1 of the backtracking line search (:274-282) + the verdict kernel; no host interaction."""
        b, algo = self._buffers(), self._algo
        K = min(self._group, self._max_backtracks - k0)
        for k in range(K):
            ratio = np.float32(self._backtrack_ratio ** (k0 + k))
            self._axpy(-float(ratio), b['step'], b['theta_prev'], b['cands'][k])
            algo.loss_terms_dev(b['cands'][k], phases, b['terms'][k])
        _lib.call('promp_trpo_select', b['P'], K, b['T'], k0, self._max_backtracks, _lib.ptr(b['terms']), _lib.ptr(b['base']),
                  self._max_constraint_val, _lib.ptr(b['theta_prev']), _lib.ptr(b['cands']), _lib.ptr(b['scal']),
                  _lib.ptr(algo.policy.theta), _lib.ptr(b['result']), _lib.stream())
        return K

    def optimize_device(self, phases):
        """Everything up to the verdict on the first candidate group, without touching the host (CUDA-graph capturable).
        Returns the device result vector (promp_trpo_select layout)."""
        b, algo = self._buffers(), self._algo
        theta = algo.policy.theta
        b['theta_prev'].copy_(theta)
        algo.loss_terms_dev(b['theta_prev'], phases, b['base'])                     # loss_before, KL before (:249, trpo_maml.py:172-175)
        g = algo.eval_gradient_dev(b['theta_prev'], phases, 'loss')                 # (:253)
        self.last_gradient = g
        x = self.descent_direction(b['theta_prev'], phases, g)                      # (:259-260)
        gp, gm, div = self._kl_grad_pair(b['theta_prev'], x, phases)
        _lib.call('promp_trpo_step', b['P'], _lib.ptr(gp), _lib.ptr(gm), div, self._reg_coeff, _lib.ptr(x),
                  self._max_constraint_val, _lib.ptr(b['step']), _lib.ptr(b['scal']), _lib.stream())      # (:262-269)
        self._evaluate_group(phases, 0)
        return b['result']

    def continue_line_search(self, phases, result_host):
        """Host continuation when no candidate of the speculative group was acceptable: further groups, one small
        device->host read each.  Returns the final result vector (host)."""
        k0 = self._group
        while result_host[6] != 0.0:
            self._evaluate_group(phases, k0)
            result_host = self._buffers()['result'].cpu().numpy()
            k0 += self._group
        return result_host

    def optimize(self, phases):
        """(:239-307).  Updates policy.theta in place (or restores it when the step is rejected)."""
        logger.log("Start CG optimization")
        res = self.optimize_device(phases).cpu().numpy()           # the one host read of a typical TRPO step
        res = self.continue_line_search(phases, res)
        self._record(res)
        return res

    def _record(self, res):
        self.last = dict(loss_before=float(res[0]), kl_before=float(res[1]), loss=float(res[2]), kl=float(res[3]),
                         backtracks=int(res[4]), rejected=bool(res[5]), init_step_scale=float(res[7]))
        if np.isnan(res[7]):
            logger.log("Initial step size is NaN! Rejecting the step!")
        elif res[5]:
            logger.log("Line search condition violated. Rejecting the step!")
        logger.log("backtrack iters: %d" % int(res[4]))
