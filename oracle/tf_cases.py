"""Seeded INPUTS of the TF-half golden cases.  TEST INFRASTRUCTURE ONLY.

`oracle/make_golden.py::gen_tf_half_graph` feeds these inputs to the UNMODIFIED reference graph code
(meta_algos/pro_mp.py, trpo_maml.py, vpg_maml.py, base.py, optimizers/*, policies/* running on the torch-backed
`tensorflow` stand-in of oracle/stubs_tf) and stores only the OUTPUTS in tests/golden/tf_half_graph.npz; the tests
regenerate the same inputs from the same seeds (numpy RandomState streams are version-stable), so the full-size
BASELINE.json configs (40 x 2000 / 40 x 4000 samples) need no multi-megabyte fixtures.

A case = policy parameters theta [P] (reference creation order) + S = num_inner_grad_steps+1 sampling phases, each
{obs [M,N,Do], act [M,N,Da], adv [M,N], mean [M,N,Da], log_std [M,Da] (agent_infos of the sampling policy, constant
over a phase per task), adj_avg_rewards [M,N]}, all float32.
"""
import math
from collections import OrderedDict

import numpy as np

CASES = OrderedDict([
    # name                algo     M   N     Do  Da hidden S  drift  extra
    ('promp_small',   dict(algo='promp', M=4, N=300, Do=2, Da=2, hidden=64, S=2, drift=0.01)),
    ('promp_iter0',   dict(algo='promp', M=4, N=256, Do=2, Da=2, hidden=64, S=2, drift=0.0)),
    ('promp_cheetah', dict(algo='promp', M=3, N=200, Do=17, Da=6, hidden=64, S=2, drift=0.01)),
    ('promp_s3',      dict(algo='promp', M=3, N=150, Do=2, Da=2, hidden=64, S=3, drift=0.01)),
    ('promp_h32',     dict(algo='promp', M=3, N=130, Do=17, Da=6, hidden=32, S=2, drift=0.01)),
    ('promp_obs4',    dict(algo='promp', M=3, N=140, Do=4, Da=2, hidden=64, S=2, drift=0.01)),
    ('promp_cfg2',    dict(algo='promp', M=40, N=2000, Do=2, Da=2, hidden=64, S=2, drift=0.005)),      # BASELINE configs[1]
    ('promp_cfg3',    dict(algo='promp', M=40, N=4000, Do=17, Da=6, hidden=64, S=2, drift=0.005)),     # BASELINE configs[2]
    ('trpo_small',    dict(algo='trpo', M=4, N=300, Do=2, Da=2, hidden=64, S=2, drift=0.01, inner_type='likelihood_ratio')),
    ('trpo_loglik',   dict(algo='trpo', M=3, N=200, Do=17, Da=6, hidden=64, S=2, drift=0.01, inner_type='log_likelihood')),
    ('trpo_cfg4',     dict(algo='trpo', M=40, N=2000, Do=2, Da=2, hidden=64, S=2, drift=0.005, inner_type='likelihood_ratio')),  # configs[3] (all 40 tasks)
    ('emaml_small',   dict(algo='trpo', M=4, N=220, Do=2, Da=2, hidden=64, S=2, drift=0.01, inner_type='log_likelihood', exploration=True)),
    ('vpg_small',     dict(algo='vpg', M=4, N=240, Do=2, Da=2, hidden=64, S=2, drift=0.01, inner_type='likelihood_ratio')),
    ('vpg_explore',   dict(algo='vpg', M=3, N=180, Do=17, Da=6, hidden=64, S=2, drift=0.01, inner_type='log_likelihood', exploration=True)),
])

HYPER = dict(inner_lr=0.1, learning_rate=1e-3, num_ppo_steps=5, clip_eps=0.3, init_inner_kl_penalty=5e-4, step_size=0.01)


def param_layout(Do, Da, hidden):
    names = ('mean_network/hidden_0/kernel', 'mean_network/hidden_0/bias', 'mean_network/hidden_1/kernel',
             'mean_network/hidden_1/bias', 'mean_network/output/kernel', 'mean_network/output/bias',
             'log_std_network/log_std_var')
    shapes = ((Do, hidden), (hidden,), (hidden, hidden), (hidden,), (hidden, Da), (Da,), (1, Da))
    return OrderedDict(zip(names, shapes))


def unflatten(theta, Do, Da, hidden):
    out, off = OrderedDict(), 0
    for k, shp in param_layout(Do, Da, hidden).items():
        n = int(np.prod(shp))
        out[k] = np.asarray(theta[off:off + n]).reshape(shp)
        off += n
    return out


def flatten(params):
    return np.concatenate([np.asarray(v).reshape(-1) for v in params.values()])


def _forward(theta, obs, Do, Da, hidden):
    """float32 numpy tanh-MLP, only used to make the stored old means plausible (any values would do)."""
    p = list(unflatten(theta, Do, Da, hidden).values())
    h = np.tanh(obs @ p[0] + p[1])
    h = np.tanh(h @ p[2] + p[3])
    return (h @ p[4] + p[5]).astype(np.float32), p[6].reshape(-1)


def make_case(name):
    c = dict(CASES[name])
    M, N, Do, Da, hid, S = c['M'], c['N'], c['Do'], c['Da'], c['hidden'], c['S']
    rng = np.random.RandomState(sum(map(ord, name)) * 7 + 1)
    theta = []
    for k, shp in param_layout(Do, Da, hid).items():
        if k.endswith('kernel'):
            lim = math.sqrt(6.0 / (shp[0] + shp[1]))
            theta.append(rng.uniform(-lim, lim, size=shp).reshape(-1))
        elif k.endswith('bias'):
            theta.append(0.1 * rng.randn(*shp).reshape(-1))
        else:
            theta.append(-0.3 + 0.2 * rng.randn(*shp).reshape(-1))
    theta = np.concatenate(theta).astype(np.float32)
    P = theta.size
    phases = []
    for s in range(S):
        scale = c['drift'] if s == 0 else 0.02
        obs = (rng.randn(M, N, Do) * (1.5 if Do <= 4 else 0.8)).astype(np.float32)
        mean = np.zeros((M, N, Da), np.float32)
        log_std = np.zeros((M, Da), np.float32)
        for m in range(M):
            th_s = theta + (scale * rng.randn(P)).astype(np.float32) if scale > 0 else theta
            mean[m], log_std[m] = _forward(th_s, obs[m], Do, Da, hid)
        act = (mean + np.exp(log_std)[:, None, :] * rng.randn(M, N, Da)).astype(np.float32)
        adv = rng.randn(M, N)
        adv = ((adv - adv.mean(1, keepdims=True)) / (adv.std(1, keepdims=True) + 1e-8)).astype(np.float32)
        adj = (0.5 * rng.randn(M, N)).astype(np.float32)
        phases.append(dict(obs=obs, act=act, adv=adv, mean=mean, log_std=log_std, adj_avg_rewards=adj))
    c.update(name=name, theta=theta, phases=phases, P=P)
    return c


def reference_samples(case):
    """The phases as the reference's processed-sample dicts: list (S) of lists (M) of dicts (meta_algos/base.py:245-283)."""
    out = []
    for ph in case['phases']:
        N = ph['obs'].shape[1]
        out.append([dict(observations=ph['obs'][m], actions=ph['act'][m], advantages=ph['adv'][m],
                         adj_avg_rewards=ph['adj_avg_rewards'][m],
                         agent_infos=dict(mean=ph['mean'][m], log_std=np.tile(ph['log_std'][m][None], (N, 1))))
                    for m in range(case['M'])])
    return out
