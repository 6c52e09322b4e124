"""MAMLPPOOptimizer (ref: meta_policy_search/optimizers/maml_first_order_optimizer.py:6-163):
`max_epochs` full-batch steps of tf.train.AdamOptimizer on the meta objective with persistent slot
state, then compute_stats.  The Adam update is promp_adam_tf1 (TF1 formula, device step counter)."""
from promp_b200 import _lib
from promp_b200.utils.dist import allreduce_sum_


class MAMLPPOOptimizer(object):
    def __init__(self, learning_rate=1e-3, max_epochs=1, tolerance=1e-6, num_minibatches=1, verbose=False,
                 beta1=0.9, beta2=0.999, epsilon=1e-8):
        self._lr, self._max_epochs = float(learning_rate), int(max_epochs)
        self._tolerance, self._num_minibatches, self._verbose = tolerance, num_minibatches, verbose
        self._b1, self._b2, self._eps = beta1, beta2, epsilon
        self._target = None

    def build(self, policy):
        """build_graph (:48-64): the Adam slots (m, v, step) are created once and persist across iterations."""
        import torch
        self._target = policy
        P = policy.num_params
        self.m = torch.zeros(P, dtype=torch.float32, device=policy.device)
        self.v = torch.zeros(P, dtype=torch.float32, device=policy.device)
        self.step = torch.zeros(1, dtype=torch.int32, device=policy.device)

    def get_state(self):
        """Adam slots as numpy (snapshots; the reference's tf.train.Saver-less snapshot drops them, we keep them so that a
        resumed run continues bit-identically)."""
        return dict(m=self.m.cpu().numpy(), v=self.v.cpu().numpy(), step=int(self.step.item()))

    def set_state(self, st):
        import torch
        self.m.copy_(torch.from_numpy(st['m']))
        self.v.copy_(torch.from_numpy(st['v']))
        self.step.fill_(int(st['step']))

    def apply_gradient(self, grad):
        p = self._target
        _lib.call('promp_adam_tf1', p.num_params, _lib.ptr(p.theta), _lib.ptr(grad), _lib.ptr(self.m), _lib.ptr(self.v),
                  _lib.ptr(self.step), self._lr, self._b1, self._b2, self._eps, _lib.stream())

    def optimize(self, algo, phases):
        """optimize (:82-115) + compute_stats (:146-163).  Returns a device vector
        [loss_before, loss_after, inner_kl_0.., outer_kl] without synchronising the host."""
        import torch
        S1 = algo.num_inner_grad_steps
        fused = getattr(algo, 'FUSED_LOSS_TERMS', False)
        if fused:
            # [loss_before | loss_after, inner_kls, outer_kl] written in place by promp_meta_loss_terms: no cat / slicing kernels
            final = torch.empty(S1 + 3, dtype=torch.float32, device=algo.policy.device)
        loss_before = None
        for epoch in range(self._max_epochs):
            res = algo._objective_pass(phases, want_grad=True)
            allreduce_sum_(res['grad'])                 # the ONE collective of the data path: [P] floats over NVLink
            if loss_before is None:
                loss_before = algo.loss_terms(res, out=final[0:], n_out=1)[0:1] if fused else algo.loss_terms(res)[0:1]
            self.apply_gradient(res['grad'])
            self.last_grad = res['grad']
        res = algo._objective_pass(phases, want_grad=False)
        if fused and loss_before is not None:
            algo.loss_terms(res, out=final[1:])
            return final
        terms = algo.loss_terms(res)
        if loss_before is None:
            loss_before = terms[0:1]
        return torch.cat([loss_before, terms])
