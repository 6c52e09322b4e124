"""PyTorch-CPU restatement of the reference's TF1 graph half.  TEST INFRASTRUCTURE ONLY.

TensorFlow 1.x is not installable here (no network), so the graph the reference builds in
policies/*, meta_algos/* and optimizers/* is restated on torch autograd (float32 or float64).
PARITY UNPINNED by reference outputs except the distribution math and the CG solver (tests/golden/tf_half_known.npz)
- see oracle/__init__.py for what pins the rest.
`ref:` citations are relative to /root/reference/meta_policy_search/.

Conventions
-----------
* A parameter set is a flat vector theta[P] in the reference's creation order
  (ref: policies/gaussian_mlp_policy.py:55-80): hidden_0/kernel [Do,H] (row-major, [in,out] as
  ref policies/networks/mlp.py:100), hidden_0/bias [H], hidden_1/kernel [H,H], hidden_1/bias [H],
  output/kernel [H,Da], output/bias [Da], log_std_var [1,Da].
* Per-task quantities are batched along a leading M axis (same math as the reference's M replicated
  sub-graphs, ref: meta_algos/pro_mp.py:88-149; batched so the CPU timing is not handicapped).
* `data` for one sampling phase is a dict of tensors: obs [M,N,Do], act [M,N,Da], adv [M,N],
  mean [M,N,Da], log_std [M,N,Da] (the agent_infos of the sampling policy).
"""
import math
from collections import OrderedDict

import numpy as np
import torch

PARAM_NAMES = ('mean_network/hidden_0/kernel', 'mean_network/hidden_0/bias',
               'mean_network/hidden_1/kernel', 'mean_network/hidden_1/bias',
               'mean_network/output/kernel', 'mean_network/output/bias',
               'log_std_network/log_std_var')


def param_shapes(obs_dim, act_dim, hidden=(64, 64)):
    h0, h1 = hidden
    return OrderedDict(zip(PARAM_NAMES, ((obs_dim, h0), (h0,), (h0, h1), (h1,), (h1, act_dim), (act_dim,),
                                         (1, act_dim))))


def num_params(obs_dim, act_dim, hidden=(64, 64)):
    return sum(int(np.prod(s)) for s in param_shapes(obs_dim, act_dim, hidden).values())


def init_params(obs_dim, act_dim, hidden=(64, 64), init_std=1.0, rng=None, dtype=np.float32):
    """Xavier-uniform kernels, zero biases, log_std = log(init_std).
    ref: policies/networks/mlp.py:12-13, policies/gaussian_mlp_policy.py:64-69."""
    rng = np.random if rng is None else rng
    out = []
    for name, shape in param_shapes(obs_dim, act_dim, hidden).items():
        if name.endswith('kernel'):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            out.append(rng.uniform(-lim, lim, size=shape).reshape(-1))
        elif name.endswith('bias'):
            out.append(np.zeros(shape).reshape(-1))
        else:
            out.append(np.full(shape, math.log(init_std)).reshape(-1))
    return np.concatenate(out).astype(dtype)


def split_params(theta, obs_dim, act_dim, hidden=(64, 64)):
    """theta [..., P] -> list of 7 tensors with leading batch dims preserved."""
    out, off = [], 0
    lead = theta.shape[:-1]
    for shape in param_shapes(obs_dim, act_dim, hidden).values():
        n = int(np.prod(shape))
        out.append(theta[..., off:off + n].reshape(*lead, *shape))
        off += n
    return out


def dist_info(theta, obs, dims, min_log_std=None):
    """Policy forward.  theta [M,P], obs [M,N,Do] -> mean [M,N,Da], log_std [M,1,Da].
    ref: policies/networks/mlp.py:65-119 (forward_mlp), policies/gaussian_mlp_policy.py:142-184.
    `min_log_std` is applied only on the params=None path (ref gaussian_mlp_policy.py:71,161)."""
    W0, b0, W1, b1, W2, b2, ls = split_params(theta, *dims)
    h = torch.tanh(torch.matmul(obs, W0) + b0.unsqueeze(-2))
    h = torch.tanh(torch.matmul(h, W1) + b1.unsqueeze(-2))
    mean = torch.matmul(h, W2) + b2.unsqueeze(-2)
    if min_log_std is not None:
        ls = torch.clamp(ls, min=min_log_std)   # tf.maximum
    return mean, ls


def log_likelihood(x, mean, log_std):
    """ref: policies/distributions/diagonal_gaussian.py:89-109."""
    zs = (x - mean) / torch.exp(log_std)
    d = x.shape[-1]
    return -torch.sum(log_std.expand_as(mean), -1) - 0.5 * torch.sum(zs ** 2, -1) - 0.5 * d * math.log(2 * math.pi)


def likelihood_ratio(x, old_mean, old_log_std, new_mean, new_log_std):
    """ref: policies/distributions/diagonal_gaussian.py:71-87."""
    return torch.exp(log_likelihood(x, new_mean, new_log_std) - log_likelihood(x, old_mean, old_log_std))


def kl(old_mean, old_log_std, new_mean, new_log_std):
    """KL(old || new).  ref: policies/distributions/diagonal_gaussian.py:16-44."""
    old_std, new_std = torch.exp(old_log_std), torch.exp(new_log_std)
    num = (old_mean - new_mean) ** 2 + old_std ** 2 - new_std ** 2
    den = 2 * new_std ** 2 + 1e-8
    return torch.sum(num / den + new_log_std - old_log_std, -1)


def inner_surrogate(theta, d, dims, inner_type='likelihood_ratio', min_log_std=None):
    """Per-task inner objective [M].  ref: meta_algos/pro_mp.py:59-65, meta_algos/trpo_maml.py:49-67."""
    mean, ls = dist_info(theta, d['obs'], dims, min_log_std)
    if inner_type == 'likelihood_ratio':
        lr = likelihood_ratio(d['act'], d['mean'], d['log_std'], mean, ls)
        return -torch.mean(lr * d['adv'], -1), (mean, ls)
    if inner_type == 'log_likelihood':
        return -torch.mean(log_likelihood(d['act'], mean, ls) * d['adv'], -1), (mean, ls)
    raise NotImplementedError(inner_type)


def adapt_sym(theta, d, dims, inner_lr, inner_type='likelihood_ratio', min_log_std=None, create_graph=False):
    """theta' = theta - inner_lr * d surr / d theta, per task.  ref: meta_algos/base.py:192-215."""
    surr, dist = inner_surrogate(theta, d, dims, inner_type, min_log_std)
    (g,) = torch.autograd.grad(surr.sum(), theta, create_graph=create_graph)
    return theta - inner_lr * g, surr, dist


def adapt(theta_tasks, d, dims, inner_lr, inner_type='likelihood_ratio'):
    """MAMLAlgo._adapt: numeric inner step on the current per-task parameters (no log_std clip:
    the adapt graph is fed parameter placeholders).  ref: meta_algos/base.py:158-190, 217-242."""
    th = theta_tasks.detach().clone().requires_grad_(True)
    new, _, _ = adapt_sym(th, d, dims, inner_lr, inner_type, None, False)
    return new.detach()


def meta_objective(theta, all_data, dims, inner_lr, algo='promp', clip_eps=0.3, inner_kl_coeff=None,
                   inner_type='likelihood_ratio', min_log_std=math.log(1e-6), exploration=False):
    """The outer objective with the inner steps kept symbolic (second order).
    ProMP: ref meta_algos/pro_mp.py:88-163.  TRPO-MAML: ref meta_algos/trpo_maml.py:100-159.
    theta [P] (requires_grad).  all_data: list (len S = num_inner_grad_steps+1) of phase dicts.
    Returns (objective, inner_kls [S-1], outer_kl)."""
    M = all_data[0]['obs'].shape[0]
    if not theta.requires_grad:
        theta = theta.detach().clone().requires_grad_(True)
    cur = theta.unsqueeze(0).expand(M, -1)
    inner_kls = []
    clip0 = min_log_std     # step 0 runs distribution_info_sym(params=None): clipped log_std
    for s in range(len(all_data) - 1):
        d = all_data[s]
        new, surr, (mean, ls) = adapt_sym(cur, d, dims, inner_lr, inner_type, clip0, create_graph=True)
        inner_kls.append(torch.mean(torch.mean(kl(d['mean'], d['log_std'], mean, ls), -1)))
        cur, clip0 = new, None
    d = all_data[-1]
    mean, ls = dist_info(cur, d['obs'], dims, clip0)
    lr = likelihood_ratio(d['act'], d['mean'], d['log_std'], mean, ls)
    outer_kl = torch.mean(torch.mean(kl(d['mean'], d['log_std'], mean, ls), -1))
    if algo == 'promp':
        clipped = torch.minimum(lr * d['adv'], torch.clamp(lr, 1 - clip_eps, 1 + clip_eps) * d['adv'])
        surr = -torch.mean(clipped, -1)
        coeff = torch.as_tensor(inner_kl_coeff, dtype=theta.dtype)
        inner_kls_t = torch.stack(inner_kls) if inner_kls else torch.zeros(0, dtype=theta.dtype)
        penalty = torch.mean(coeff * inner_kls_t) if inner_kls else torch.zeros((), dtype=theta.dtype)
        obj = torch.mean(surr) + penalty
    elif algo in ('trpo', 'vpg'):
        if algo == 'trpo':
            surr = -torch.mean(lr * d['adv'], -1)
        else:   # VPG-MAML outer objective (ref meta_algos/vpg_maml.py:128-130)
            surr = -torch.mean(log_likelihood(d['act'], mean, ls) * d['adv'], -1)
        if exploration:
            # E-MAML (ref meta_algos/trpo_maml.py:137-144): - mean(adj_avg_rewards of the last phase) *
            # mean(log-likelihood of the INITIAL actions under the pre-update policy)
            d0 = all_data[0]
            mean0, ls0 = dist_info(theta.unsqueeze(0).expand(M, -1), d0['obs'], dims, min_log_std)
            surr = surr - torch.mean(d['adj_avg_rewards'], -1) * torch.mean(log_likelihood(d0['act'], mean0, ls0), -1)
        obj = torch.mean(surr)
        inner_kls_t = torch.stack(inner_kls) if inner_kls else torch.zeros(0, dtype=theta.dtype)
    else:
        raise NotImplementedError(algo)
    return obj, inner_kls_t, outer_kl


class TF1Adam(object):
    """tf.train.AdamOptimizer update rule (persistent slots).
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, n, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, dtype=torch.float32):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = torch.zeros(n, dtype=dtype)
        self.v = torch.zeros(n, dtype=dtype)
        self.t = 0

    def step(self, theta, grad):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        self.m = self.b1 * self.m + (1 - self.b1) * grad
        self.v = self.b2 * self.v + (1 - self.b2) * grad * grad
        return theta - lr_t * self.m / (torch.sqrt(self.v) + self.eps)


def promp_optimize(theta, all_data, dims, adam, inner_lr, clip_eps, inner_kl_coeff, num_ppo_steps=5):
    """ProMP.optimize_policy: K full-batch Adam epochs, then one stats pass.
    ref: meta_algos/pro_mp.py:165-199, optimizers/maml_first_order_optimizer.py:82-115, 146-163.
    Returns (theta_new, dict(loss_before, loss_after, inner_kls, outer_kl, grads=[...]))."""
    loss_before, grads = None, []
    for _ in range(num_ppo_steps):
        th = theta.detach().clone().requires_grad_(True)
        obj, _, _ = meta_objective(th, all_data, dims, inner_lr, 'promp', clip_eps, inner_kl_coeff)
        (g,) = torch.autograd.grad(obj, th)
        if not loss_before:                     # ref :104 (`if not loss_before_opt`)
            loss_before = float(obj)
        grads.append(g.detach().clone())
        theta = adam.step(theta.detach(), g.detach())
    th = theta.detach().clone().requires_grad_(True)
    obj, inner_kls, outer_kl = meta_objective(th, all_data, dims, inner_lr, 'promp', clip_eps, inner_kl_coeff)
    return theta.detach(), dict(loss_before=loss_before, loss_after=float(obj),
                                inner_kls=inner_kls.detach().numpy(), outer_kl=float(outer_kl), grads=grads)


def adapt_kl_coeff(kl_coeff, kl_values, kl_target):
    """ref: meta_algos/pro_mp.py:201-214."""
    out = []
    for c, k in zip(kl_coeff, kl_values):
        if k < kl_target / 1.5:
            c = c / 2
        elif k > kl_target * 1.5:
            c = c * 2
        out.append(c)
    return np.array(out)


# ------------------------------------------------------------------------------- TRPO-MAML
def conjugate_gradients(f_Ax, b, cg_iters=10, residual_tol=1e-10):
    """ref: optimizers/conjugate_gradient_optimizer.py:325-354 (numpy float32 vectors)."""
    p, r = b.copy(), b.copy()
    x = np.zeros_like(b, dtype=np.float32)
    rdotr = r.dot(r)
    for _ in range(cg_iters):
        z = f_Ax(p)
        v = rdotr / p.dot(z)
        x += v * p
        r -= v * z
        newrdotr = r.dot(r)
        p = r + (newrdotr / rdotr) * p
        rdotr = newrdotr
        if rdotr < residual_tol:
            break
    return x


class TRPOMAMLOracle(object):
    """TRPOMAML.optimize_policy + ConjugateGradientOptimizer.optimize with FiniteDifferenceHvp.
    ref: meta_algos/trpo_maml.py:161-192; optimizers/conjugate_gradient_optimizer.py:59-89, 239-307."""

    def __init__(self, dims, inner_lr, step_size=0.01, inner_type='likelihood_ratio', cg_iters=10,
                 backtrack_ratio=0.8, max_backtracks=15, base_eps=1e-5, dtype=torch.float32):
        self.dims, self.inner_lr, self.step_size, self.inner_type = dims, inner_lr, step_size, inner_type
        self.cg_iters, self.backtrack_ratio, self.max_backtracks = cg_iters, backtrack_ratio, max_backtracks
        self.base_eps, self.dtype = np.float32(base_eps), dtype

    def _eval(self, theta_np, all_data, want):
        th = torch.as_tensor(theta_np, dtype=self.dtype).clone().requires_grad_(True)
        obj, _, okl = meta_objective(th, all_data, self.dims, self.inner_lr, 'trpo', inner_type=self.inner_type)
        if want == 'loss':
            return float(obj)
        if want == 'kl':
            return float(okl)
        target = obj if want == 'grad' else okl
        (g,) = torch.autograd.grad(target, th)
        return g.detach().numpy().astype(np.float32)

    def loss(self, theta, all_data):
        return self._eval(theta, all_data, 'loss')

    def constraint_val(self, theta, all_data):
        return self._eval(theta, all_data, 'kl')

    def gradient(self, theta, all_data):
        return self._eval(theta, all_data, 'grad')

    def constraint_gradient(self, theta, all_data):
        return self._eval(theta, all_data, 'klgrad')

    def Hx(self, theta, all_data, x):
        eps = self.base_eps
        gp = self.constraint_gradient(theta + eps * x, all_data)
        gm = self.constraint_gradient(theta - eps * x, all_data)
        return (gp - gm) / (2 * eps)

    def optimize(self, theta, all_data):
        theta = np.asarray(theta, dtype=np.float32)
        loss_before = self.loss(theta, all_data)
        g = self.gradient(theta, all_data)
        Hx = lambda x: self.Hx(theta, all_data, x)
        direction = conjugate_gradients(Hx, g, cg_iters=self.cg_iters)
        init_step = np.sqrt(2.0 * self.step_size * (1. / (direction.dot(Hx(direction)) + 1e-8)))
        if np.isnan(init_step):
            return theta, dict(loss_before=loss_before, rejected=True)
        full = init_step * direction
        loss = kl_ = 0.0
        n_iter = 0
        cur = theta
        for n_iter, ratio in enumerate(self.backtrack_ratio ** np.arange(self.max_backtracks)):
            cur = (theta - ratio * full).astype(np.float32)
            loss, kl_ = self.loss(cur, all_data), self.constraint_val(cur, all_data)
            if loss < loss_before and kl_ <= self.step_size:
                break
        violated = bool(np.isnan(loss) or np.isnan(kl_) or loss >= loss_before or kl_ >= self.step_size)
        if violated:
            cur = theta
        return cur, dict(loss_before=loss_before, loss=loss, kl=kl_, backtracks=n_iter, rejected=violated,
                         gradient=g, direction=direction, init_step=init_step)


# ------------------------------------------------------------------------- numpy-facing policy
class OraclePolicy(object):
    """MetaGaussianMLPPolicy as seen by the sampler (get_actions / pre & post update modes), with
    injectable action noise so rollouts are reproducible.
    ref: policies/meta_gaussian_mlp_policy.py:84-157, policies/base.py:218-286."""

    def __init__(self, meta_batch_size, obs_dim, action_dim, hidden_sizes=(64, 64), init_std=1., min_std=1e-6,
                 theta=None, noise=None, dtype=np.float32):
        self.meta_batch_size, self.obs_dim, self.action_dim = meta_batch_size, obs_dim, action_dim
        self.dims = (obs_dim, action_dim, tuple(hidden_sizes))
        self.min_log_std = math.log(min_std)
        self.dtype = dtype
        self.theta = init_params(*self.dims, init_std=init_std, dtype=dtype) if theta is None else np.asarray(theta, dtype)
        self.theta_tasks = None
        self._pre_update_mode = True
        self.noise = noise          # callable (step_index, shape) -> eps, or None for np.random.normal
        self._t = 0

    def switch_to_pre_update(self):
        self._pre_update_mode = True
        self.theta_tasks = np.tile(self.theta, (self.meta_batch_size, 1))
        self._t = 0

    def update_task_parameters(self, theta_tasks):
        self.theta_tasks = np.asarray(theta_tasks, dtype=self.dtype)
        self._pre_update_mode = False
        self._t = 0

    def reset_noise_clock(self):
        self._t = 0

    def get_actions(self, observations):
        obs = np.asarray(observations).astype(self.dtype)                 # f64 -> f32 at the placeholder
        W0, b0, W1, b1, W2, b2, ls = [p.numpy() if hasattr(p, 'numpy') else p for p in
                                      split_params(torch.as_tensor(self.theta_tasks), *self.dims)]
        h = np.tanh(np.matmul(obs, W0) + b0[:, None, :])
        h = np.tanh(np.matmul(h, W1) + b1[:, None, :])
        mean = np.matmul(h, W2) + b2[:, None, :]
        # sampled with the raw log_std; reported log_std is clipped only pre-update
        # (ref gaussian_mlp_policy.py:71-74, meta_gaussian_mlp_policy.py:45-47, 66-70)
        eps = (np.random.normal(size=mean.shape) if self.noise is None else self.noise(self._t, mean.shape))
        self._t += 1
        actions = mean + eps.astype(self.dtype) * np.exp(ls)
        rep = np.maximum(ls, self.min_log_std) if self._pre_update_mode else ls
        infos = [[dict(mean=mean[m, e], log_std=rep[m, 0]) for e in range(mean.shape[1])]
                 for m in range(self.meta_batch_size)]
        return list(actions), infos
