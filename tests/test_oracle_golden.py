"""Pin the CPU oracle (oracle/numpy_half.py) to the UNMODIFIED reference's outputs stored in
tests/golden/ (written by oracle/make_golden.py) and to the known-answer vectors of the reference's
own tests (tests/test_samplers.py).  CPU only."""
import os
from collections import OrderedDict

import numpy as np
import pytest

from oracle import numpy_half as nh
from oracle.tf_half import OraclePolicy


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize('rtype', ['sparse', 'dense', 'dense_squared'])
def test_point_corner_env_matches_reference(golden_dir, rtype):
    g = _load(golden_dir, 'point_corner_steps.npz')
    actions, goals, obs0 = g['actions'], g['goals'], g['obs0']
    T, n_env, _ = actions.shape
    envs = []
    for i in range(n_env):
        e = nh.NormalizedEnv(nh.PointEnvCorner(reward_type=rtype))
        e.set_task(goals[i])
        e._wrapped_env._state = obs0[i].copy()
        envs.append(e)
    for t in range(T):
        for i, e in enumerate(envs):
            o, r, d, info = e.step(actions[t, i])
            assert d is False and info == {}
            assert np.array_equal(o, g['next_obs_' + rtype][t, i])      # bit-exact (float64 both sides)
            assert r == g['rewards_' + rtype][t, i]
    if rtype == 'sparse':   # all three branches of the sparse reward occurred in the fixture
        r = g['rewards_sparse']
        assert (r == 0).any() and (r > 0).any() and (r < 0).any()


def test_point_env_matches_reference(golden_dir):
    g = _load(golden_dir, 'point_env_steps.npz')
    T, n_env, _ = g['actions'].shape
    envs = []
    for i in range(n_env):
        e = nh.NormalizedEnv(nh.PointEnv())
        e._wrapped_env._state = g['obs0'][i].copy()
        envs.append(e)
    for t in range(T):
        for i, e in enumerate(envs):
            o, r, d, _ = e.step(g['actions'][t, i])
            assert np.array_equal(o, g['next_obs'][t, i])
            assert r == g['rewards'][t, i]
            assert d == g['dones'][t, i]
    assert g['dones'].any()


def test_utils_known_answers(golden_dir):
    g = _load(golden_dir, 'utils_known.npz')
    assert np.array_equal(nh.discount_cumsum(g['x'], 0.99), g['dc_099'])
    assert np.array_equal(nh.discount_cumsum(g['x'], 0.5), g['dc_05'])
    assert np.array_equal(nh.baseline_features(g['obs']), g['feats'])
    # plain recurrence == lfilter to rounding
    y = np.zeros(len(g['x']) + 1)
    for t in reversed(range(len(g['x']))):
        y[t] = g['x'][t] + 0.99 * y[t + 1]
    np.testing.assert_allclose(y[:-1], g['dc_099'], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize('case', list('abcde'))
def test_process_samples_matches_reference(golden_dir, case):
    g = _load(golden_dir, 'process_samples.npz')
    pre = 'case_%s_' % case
    cfg = {k: g[pre + 'cfg_' + k].item() for k in ('M', 'E', 'H', 'Do', 'Da', 'discount', 'gae_lambda',
                                                    'normalize_adv', 'positive_adv')}
    obs, act, rew = (g[pre + k].astype(np.float64) for k in ('obs', 'act', 'rew'))
    paths = OrderedDict()
    for m in range(cfg['M']):
        paths[m] = [dict(observations=obs[m, e], actions=act[m, e], rewards=rew[m, e], env_infos={}, agent_infos={})
                    for e in range(cfg['E'])]
    base = nh.LinearFeatureBaseline()
    proc = nh.SampleProcessor(base, cfg['discount'], cfg['gae_lambda'], bool(cfg['normalize_adv']), bool(cfg['positive_adv']))
    coeffs = []
    fit = base.fit
    base.fit = lambda p, target_key='returns': (fit(p, target_key), coeffs.append(base._coeffs.copy()))
    data = proc.process_samples(paths)
    assert np.array_equal(np.stack([d['returns'] for d in data]), g[pre + 'returns'])
    np.testing.assert_allclose(np.stack(coeffs), g[pre + 'coeffs'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.stack([d['advantages'] for d in data]), g[pre + 'advantages'], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.stack([d['adj_avg_rewards'] for d in data]), g[pre + 'adj_avg_rewards'], rtol=1e-12)
    if case == 'a':   # index contract: flat sample n = e*H + t
        np.testing.assert_array_equal(np.stack([d['observations'] for d in data]).astype(np.float32),
                                      g[pre + 'observations_stacked'])
        np.testing.assert_array_equal(data[0]['observations'], obs[0].reshape(-1, cfg['Do']))


def test_point_variants_match_reference(golden_dir):
    """MetaPointEnvWalls / MetaPointEnvMomentum restatements vs the unmodified reference (point_variants_steps.npz):
    task draws (RNG order), resets and 150 / 60 closed-loop steps, bit for bit."""
    import copy
    g = _load(golden_dir, 'point_variants_steps.npz')
    T, n_env, _ = g['walls_actions'].shape
    for rtype in ('dense', 'dense_squared'):
        env = nh.NormalizedEnv(nh.PointEnvWalls(reward_type=rtype))
        np.random.seed(17)
        tasks = env.sample_tasks(n_env)
        assert np.array_equal(np.stack([np.concatenate([t['goal'], t['gap_1'], t['gap_2']]) for t in tasks]), g['walls_tasks'])
        envs = [copy.deepcopy(env) for _ in range(n_env)]
        for i, e in enumerate(envs):
            e.set_task(tasks[i])
            assert np.array_equal(e.reset(), g['walls_obs0'][i])
        for t in range(T):
            for i, e in enumerate(envs):
                o, r, d, info = e.step(g['walls_actions'][t, i])
                assert np.array_equal(o, g['walls_next_obs_' + rtype][t, i]) and r == g['walls_rewards_' + rtype][t, i]
    assert np.array_equal(np.random.uniform(size=3), g['walls_rng_probe_after'])
    T, n_env, _ = g['momentum_actions'].shape
    for rtype in ('sparse', 'dense', 'dense_squared'):
        env = nh.NormalizedEnv(nh.PointEnvMomentum(reward_type=rtype))
        np.random.seed(19)
        tasks = env.sample_tasks(n_env)
        assert np.array_equal(np.asarray(tasks, dtype=np.float64), g['momentum_goals'])
        envs = [copy.deepcopy(env) for _ in range(n_env)]
        for i, e in enumerate(envs):
            e.set_task(tasks[i])
            assert np.array_equal(e.reset(), g['momentum_obs0'][i])
        for t in range(T):
            for i, e in enumerate(envs):
                o, r, d, info = e.step(g['momentum_actions'][t, i])
                assert np.array_equal(o, g['momentum_next_obs_' + rtype][t, i]) and r == g['momentum_rewards_' + rtype][t, i]


def ragged_paths_from_golden(g, pre):
    """Rebuild the per-task path lists of a process_samples_ragged.npz case (float64 like the reference's inputs)."""
    M = int(g[pre + 'cfg_M'])
    n_paths, path_len = g[pre + 'n_paths'], g[pre + 'path_len']
    obs, act, rew, mean = (g[pre + k].astype(np.float64) for k in ('obs', 'act', 'rew', 'mean'))
    paths, off, pi = OrderedDict(), 0, 0
    for m in range(M):
        paths[m] = []
        for _ in range(int(n_paths[m])):
            L = int(path_len[pi]); pi += 1
            sl = slice(off, off + L); off += L
            paths[m].append(dict(observations=obs[sl], actions=act[sl], rewards=rew[sl], env_infos={},
                                 agent_infos=dict(mean=mean[sl], log_std=np.tile(g[pre + 'log_std'][m].astype(np.float64), (L, 1)))))
    return paths


@pytest.mark.parametrize('case', ['r1', 'r2', 'r3'])
def test_process_samples_ragged_matches_reference(golden_dir, case):
    """Variable-length paths (early termination): the oracle's per-path scans, feature time index and ragged fit
    reproduce the reference MetaSampleProcessor."""
    g = _load(golden_dir, 'process_samples_ragged.npz')
    pre = 'case_%s_' % case
    paths = ragged_paths_from_golden(g, pre)
    base = nh.LinearFeatureBaseline()
    proc = nh.SampleProcessor(base, float(g[pre + 'cfg_discount']), float(g[pre + 'cfg_gae_lambda']),
                              bool(g[pre + 'cfg_normalize_adv']), bool(g[pre + 'cfg_positive_adv']))
    coeffs = []
    fit = base.fit
    base.fit = lambda p, target_key='returns': (fit(p, target_key), coeffs.append(base._coeffs.copy()))
    data = proc.process_samples(paths)
    assert np.array_equal(np.concatenate([d['returns'] for d in data]), g[pre + 'returns'])
    np.testing.assert_allclose(np.stack(coeffs), g[pre + 'coeffs'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.concatenate([d['advantages'] for d in data]), g[pre + 'advantages'], rtol=1e-9, atol=1e-10)
    np.testing.assert_array_equal(np.concatenate([d['observations'] for d in data]).astype(np.float32),
                                  g[pre + 'observations_stacked'])


def test_sampler_rollout_matches_reference(golden_dir):
    """Oracle sampler + envs reproduce the reference MetaSampler draw-for-draw (seed 1, configs[0])."""
    g = _load(golden_dir, 'sampler_rollout.npz')
    M, E, H = 5, 4, 100
    noise = g['noise']
    phase = [0]
    policy = OraclePolicy(M, 2, 2, theta=g['theta'], noise=lambda t, shape: noise[phase[0], t])
    np.random.seed(1)
    env = nh.NormalizedEnv(nh.PointEnvCorner())
    sampler = nh.Sampler(env, policy, E, M, H)
    for it in range(2):
        tasks = sampler.update_tasks()
        assert np.array_equal(np.asarray(tasks, dtype=np.float64), g['it%d_goals' % it])
        policy.switch_to_pre_update()
        phase[0] = it
        paths = sampler.obtain_samples()
        for key, name in (('observations', 'obs'), ('actions', 'act'), ('rewards', 'rew')):
            got = np.stack([np.stack([p[key] for p in paths[m]]) for m in range(M)])
            assert np.array_equal(got, g['it%d_%s' % (it, name)]), (it, key)
    assert np.array_equal(np.random.uniform(size=4), g['rng_probe_after'])


# ---- known-answer vectors of the reference's own tests (tests/test_samplers.py) ----
class _IntegratorEnv(object):
    """The reference test double (tests/test_samplers.py:13-45): 1-D integrator, obs = 100*state+goal.
    (The reference runs this known-answer check through the pickling parallel executor only, because
    its reset() returns the live state array; here reset() hands out a copy instead.)"""
    obs_dim = act_dim = 1

    def __init__(self):
        self.state, self.goal = np.zeros(1), 0

    def sample_tasks(self, n):
        return np.random.choice(100, n, replace=False)

    def set_task(self, task):
        self.goal = task

    def step(self, action):
        self.state += self.goal - action
        return self.state * 100 + self.goal, (self.goal - action)[0], 0, {'e': self.state}

    def reset(self):
        self.state = np.zeros(1)
        return self.state.copy()


class _OnesPolicy(object):
    def get_actions(self, observations):
        return [[np.ones(1) for _ in task] for task in observations], None


def test_reference_known_answer_rollout():
    """tests/test_samplers.py:84-97: constant action 1 => obs 0,-100,-200,...; 3 tasks x 4 paths x 5."""
    sampler = nh.Sampler(_IntegratorEnv(), _OnesPolicy(), 4, 3, 5)
    paths = sampler.obtain_samples()
    assert len(paths) == 3
    for task in paths.values():
        assert len(task) == 4
        for path in task:
            assert all(a == 1 for a in path['actions'])
            assert [float(o[0]) for o in path['observations']] == [0., -100., -200., -300., -400.]


def test_reference_advantage_identity():
    """tests/test_samplers.py:326-342: gamma=lambda=1, zero baseline => adv[t] = sum_{k>=t} r[k]."""
    rng = np.random.RandomState(0)
    paths = OrderedDict((m, [dict(observations=rng.randn(L, 2), actions=rng.randn(L, 2), rewards=rng.randn(L),
                                  env_infos={}, agent_infos={}) for L in (7, 3, 11)]) for m in range(2))
    proc = nh.SampleProcessor(nh.ZeroBaseline(), discount=1.0, gae_lambda=1.0)
    data = proc.process_samples(paths)
    for m in range(2):
        off = 0
        for p in paths[m]:
            L = len(p['rewards'])
            np.testing.assert_allclose(data[m]['advantages'][off:off + L], np.cumsum(p['rewards'][::-1])[::-1], atol=1e-10)
            off += L
        assert len(data[m].keys()) == 8


@pytest.mark.parametrize('Da', [2, 6])
def test_tf_half_distribution_math_matches_reference_numpy(golden_dir, Da):
    """The numpy-executable part of the reference's TF1 half pins the oracle's distribution math: DiagonalGaussian.kl /
    log_likelihood (policies/distributions/diagonal_gaussian.py:46-69, 111-127; the *_sym graph versions are the same
    expressions in TF ops) for random distributions incl. log_std at the 1e-6 clip floor.  float64, 1e-12."""
    import torch
    from oracle import tf_half as th
    g = _load(golden_dir, 'tf_half_known.npz')
    pre = 'dist%d_' % Da
    t = {k: torch.from_numpy(g[pre + k]) for k in ('old_mean', 'old_ls', 'new_mean', 'new_ls', 'x')}
    np.testing.assert_allclose(th.kl(t['old_mean'], t['old_ls'], t['new_mean'], t['new_ls']).numpy(), g[pre + 'kl'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(th.log_likelihood(t['x'], t['old_mean'], t['old_ls']).numpy(), g[pre + 'll_old'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(th.log_likelihood(t['x'], t['new_mean'], t['new_ls']).numpy(), g[pre + 'll_new'], rtol=1e-12, atol=1e-12)
    ratio = th.likelihood_ratio(t['x'], t['old_mean'], t['old_ls'], t['new_mean'], t['new_ls']).numpy()
    np.testing.assert_allclose(ratio, np.exp(g[pre + 'll_new'] - g[pre + 'll_old']), rtol=1e-12, atol=0)
    # entropy (diagonal_gaussian.py:142-153) = sum(log_std + log sqrt(2 pi e)): consistency of the fixture itself
    np.testing.assert_allclose(g[pre + 'entropy'], np.sum(g[pre + 'new_ls'] + np.log(np.sqrt(2 * np.pi * np.e)), axis=-1), rtol=1e-12)


def test_conjugate_gradients_match_reference(golden_dir):
    """conjugate_gradients (optimizers/conjugate_gradient_optimizer.py:325-354) run from the unmodified reference: the
    oracle's restatement reproduces it bit for bit (float32).  The product's CG runs on the device (promp_cg_init /
    promp_cg_step) and is checked against the same vectors in tests/test_gpu_parity.py::test_device_cg_and_line_search_kernels."""
    from oracle import tf_half as th
    g = _load(golden_dir, 'tf_half_known.npz')
    A, b = g['cg_A'], g['cg_b']
    for fn in (th.conjugate_gradients,):
        assert np.array_equal(fn(lambda p: A.dot(p), b, cg_iters=10), g['cg_x10'])
        assert np.array_equal(fn(lambda p: A.dot(p), b, cg_iters=3), g['cg_x3'])
        assert np.array_equal(fn(lambda p: A.dot(p), b, cg_iters=200, residual_tol=1e-6), g['cg_x_tol'])
    assert np.linalg.norm(A.dot(g['cg_x_tol']) - b) < 1e-2 * np.linalg.norm(b)


def test_adaptive_kl_coefficient_rule_matches_reference(golden_dir):
    """_adapt_kl_coeff (meta_algos/pro_mp.py:201-214) from the unmodified reference, incl. the exact /1.5 and *1.5
    thresholds: the oracle's and the product's restatements give the same coefficients."""
    import ast
    from oracle import tf_half as th
    g = _load(golden_dir, 'tf_half_known.npz')
    target = float(g['klc_target'])
    np.testing.assert_array_equal(th.adapt_kl_coeff(g['klc_in'], g['klc_kl'], target), g['klc_out'])
    # the product's rule lives in promp_b200/meta_algos/pro_mp.py (module import needs the CUDA library: evaluate the
    # function's source alone)
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'promp_b200', 'meta_algos', 'pro_mp.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == '_adapt_kl_coeff'][0]
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'pro_mp._adapt_kl_coeff', 'exec'), ns)
    got = np.asarray([ns['_adapt_kl_coeff'](float(c), float(k), target) for c, k in zip(g['klc_in'], g['klc_kl'])])
    np.testing.assert_array_equal(got, g['klc_out'])
