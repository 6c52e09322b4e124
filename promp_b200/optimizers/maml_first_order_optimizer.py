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
        self._ticket = torch.zeros(1, dtype=torch.int32, device=policy.device)      # completion ticket of promp_meta_update
        self.last_grad = torch.zeros(P, dtype=torch.float32, device=policy.device)

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

    def apply_task_gradients(self, task_grads):
        """Per-task meta-gradients [M, P] -> task mean -> sum over ranks (NVLink peer memory) -> TF1 Adam, one launch
        (promp_meta_update).  Returns False when the multi-rank case has no peer-memory communicator (NCCL fallback)."""
        from promp_b200.utils import dist as _dist
        p = self._target
        W = _dist.world_size()
        p2p = _dist._p2p
        if W > 1 and (p2p is None or p.num_params > p2p.cap):
            return False
        M = task_grads.shape[0]
        comm = (p2p.world, p2p.rank, p2p.cap, _lib.ptr(p2p.peers), _lib.ptr(p2p.epoch), _lib.ptr(p2p.error)) if W > 1 else \
            (1, 0, 0, None, None, None)
        _lib.call('promp_meta_update', M, p.num_params, _lib.ptr(task_grads), 1.0 / (M * W), _lib.ptr(self.last_grad),
                  _lib.ptr(p.theta), _lib.ptr(self.m), _lib.ptr(self.v), _lib.ptr(self.step), self._lr, self._b1, self._b2,
                  self._eps, *comm, _lib.ptr(self._ticket), _lib.stream())
        return True

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
        fused_update = getattr(algo, 'FUSED_META_UPDATE', False)
        for epoch in range(self._max_epochs):
            res = algo._objective_pass(phases, want_grad=True, reduce=False) if fused_update else \
                algo._objective_pass(phases, want_grad=True)
            if loss_before is None:
                loss_before = algo.loss_terms(res, out=final[0:], n_out=1)[0:1] if fused else algo.loss_terms(res)[0:1]
            if fused_update and self.apply_task_gradients(res['grad_tasks']):
                continue                                # task mean + all-reduce + Adam happened in one launch
            if fused_update:                            # no peer-memory communicator: reduce here, all-reduce through NCCL
                flat = torch.empty(algo.policy.num_params, dtype=torch.float32, device=algo.policy.device)
                M = res['grad_tasks'].shape[0]
                from promp_b200.utils.dist import world_size
                _lib.call('promp_reduce_tasks', M, algo.policy.num_params, _lib.ptr(res['grad_tasks']), 1.0 / (M * world_size()),
                          _lib.ptr(flat), _lib.stream())
                res['grad'] = flat
            allreduce_sum_(res['grad'])                 # the ONE collective of the data path: [P] floats over NVLink
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
