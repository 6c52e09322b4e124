"""The reference's own sampler / baseline tests, restated 1:1 against promp_b200's classes with the same duck-typed
fakes (ref: tests/test_samplers.py:13-190, 283-344; tests/test_baselines.py:11-98).  The fake envs / policies are
arbitrary host Python objects, so MetaSampler steps them through MetaHostEnvExecutor (no CUDA needed: those tests run in
the CPU suite); everything that touches sample processing or the baseline runs the device kernels (-m gpu)."""
import pickle

import numpy as np
import pytest

from promp_b200.samplers import MetaSampler, MetaSampleProcessor
from promp_b200.baselines import LinearFeatureBaseline, ZeroBaseline


# ---- test doubles with the behaviour of the reference's (tests/test_samplers.py:13-67) --------------------------------
class FakeEnv(object):
    """1-d integrator: state += goal - action; obs = 100*state + goal; reward = goal - action."""

    def __init__(self):
        self.state, self.goal = np.zeros(1), 0

    def sample_tasks(self, n_tasks):
        return np.random.choice(100, n_tasks, replace=False)      # distinct goals

    def set_task(self, task):
        self.goal = task

    def get_task(self):
        return self.goal

    def step(self, action):
        self.state += self.goal - action
        return self.state * 100 + self.goal, (self.goal - action)[0], 0, {'e': self.state}

    def reset(self):
        self.state = np.zeros(1)
        return self.state


class FakeRandomEnv(FakeEnv):
    def step(self, action):
        self.state += (self.goal - action) * np.random.random()
        return self.state * 100 + self.goal, (self.goal - action)[0], 0, {'e': self.state}


class OnesPolicy(object):
    def get_actions(self, observations):
        return [[np.ones(1) for _ in task] for task in observations], None


class FakeRandomPolicy(object):
    def get_actions(self, observations):
        return ([[np.random.random() * o for o in task] for task in observations],
                [[{'a': 1, 'b': 2} for _ in task] for task in observations])


M, E, H = 3, 4, 5


def _samplers(env, policy):
    return [MetaSampler(env, policy, E, M, H, parallel=p) for p in (False, True)]


def test_single():
    """ref tests/test_samplers.py:84-98."""
    for sampler in _samplers(FakeEnv(), OnesPolicy()):
        paths = sampler.obtain_samples()
        assert len(paths) == M
        for task in paths.values():
            assert len(task) == E
            for path in task:
                assert len(path) == H                 # 5 keys, like the reference's path dict
                assert all(a == 1 for a in path['actions'])
                expect = 0
                for obs in path['observations']:
                    assert obs == expect
                    expect += -100


def test_goal_set():
    """ref :100-117: same goal within a task, different goals across tasks."""
    for sampler in _samplers(FakeEnv(), OnesPolicy()):
        sampler.update_tasks()
        paths = sampler.obtain_samples()
        assert len(paths) == M
        for task in paths.values():
            for j in range(H):
                assert all(path["observations"][j] == task[0]["observations"][j] for path in task)
        for j in range(1, H):
            for i in range(E):
                for h in range(1, M):
                    assert paths[h][i]['observations'][j] != paths[0][i]['observations'][j]


def test_random_seeds():
    """ref :116-151: a fixed numpy seed reproduces the rollouts; envs of a task see different noise."""
    for parallel in (True, False):
        runs = []
        for _ in range(2):
            np.random.seed(22)
            sampler = MetaSampler(FakeRandomEnv(), FakeRandomPolicy(), E, M, H, parallel=parallel)
            sampler.update_tasks()
            runs.append(sampler.obtain_samples())
        for t1, t2 in zip(runs[0].values(), runs[1].values()):
            for j in range(E):
                for k in range(H):
                    assert t1[j]["observations"][k] == t2[j]["observations"][k]
        np.random.seed(22)
        sampler = MetaSampler(FakeRandomEnv(), OnesPolicy(), E, M, H, parallel=parallel)
        sampler.update_tasks()
        paths = sampler.obtain_samples()
        for task in paths.values():
            for j in range(1, H):
                for h in range(1, E):
                    assert task[h]["observations"][j] != task[0]["observations"][j]
                    assert task[h]['rewards'][j] == task[0]['rewards'][j]


def test_info_dicts():
    """ref :153-170."""
    for sampler in _samplers(FakeRandomEnv(), FakeRandomPolicy()):
        sampler.update_tasks()
        paths = sampler.obtain_samples()
        assert len(paths) == M
        for task in paths.values():
            for h in range(1, E):
                assert type(task[h]["agent_infos"]) == dict and type(task[h]["env_infos"]) == dict
                assert len(task[h]["agent_infos"].keys()) == 2
                assert len(task[h]["env_infos"].keys()) == 1
    assert sampler.total_timesteps_sampled == M * E * H


@pytest.mark.gpu
def test_meta_sample_processor_on_fake_paths():
    """ref :172-180: 8 keys per task, advantages of size path_length * batch_size."""
    proc = MetaSampleProcessor(baseline=LinearFeatureBaseline())
    for sampler in _samplers(FakeEnv(), OnesPolicy()):
        sampler.update_tasks()
        paths = sampler.obtain_samples()
        samples = proc.process_samples(paths)
        assert len(samples) == M
        for sd in samples:
            assert len(sd.keys()) == 8
            assert sd['advantages'].size == H * E


# ---- variable-length paths: advantages == reverse cumulative sum of the rewards at gamma = lambda = 1, zero baseline -------
class FakePointEnv(FakeEnv):
    """ref tests/test_samplers.py:283-300: 2-d point, early `done` near the origin."""

    def __init__(self):
        self.reset()
        self.goal = np.array([0, 0])

    def sample_tasks(self, n_tasks):
        return [np.array([0, 0]) for _ in range(n_tasks)]

    def step(self, action):
        self.state += np.clip(action, -0.1, 0.1)
        dist = np.linalg.norm(self.goal - self.state)
        return self.state, -dist, dist < 0.1, {'e': self.state}

    def reset(self):
        self.state = np.random.uniform(-2, 2, size=(2,))
        return self.state


class FakePointPolicy(object):
    def get_actions(self, observations):
        return [-np.clip(obs, -0.1, 0.1) + np.random.normal(0, scale=0.03, size=2) for obs in observations], None


@pytest.mark.gpu
def test_advantages_match_reward_to_go_on_variable_length_paths():
    """ref :302-344 (SampleProcConsistency): with ZeroBaseline, discount = gae_lambda = 1 and no normalisation the advantage
    of every step is the sum of the remaining rewards of its path - through the device processing kernel, on paths of
    different lengths collected from a host env by the collect-until-enough loop."""
    np.random.seed(5)
    sampler = MetaSampler(FakePointEnv(), FakePointPolicy(), 20, 3, 25, parallel=False)
    paths = sampler.obtain_samples()
    lens = [len(p['rewards']) for task in paths.values() for p in task]
    proc = MetaSampleProcessor(baseline=ZeroBaseline(), discount=1.0, gae_lambda=1.0, normalize_adv=False)
    samples = proc.process_samples(paths)
    for m, task in paths.items():
        adv = np.asarray(samples[m]['advantages'])
        off = 0
        for path in task:
            L = len(path['rewards'])
            want = np.cumsum(np.asarray(path['rewards'])[::-1])[::-1]
            np.testing.assert_allclose(adv[off:off + L], want, rtol=1e-5, atol=1e-5)
            off += L
        assert adv.size == off
    assert len(set(lens)) >= 1


# ---- LinearFeatureBaseline standalone (ref tests/test_baselines.py:11-98) ---------------------------------------------
class BaselineRandomEnv(FakeEnv):
    def step(self, action):
        self.state += (self.goal - action) * np.random.random()
        return self.state * 100 + self.goal, (self.goal - action)[0], 0, {}


class BaselineRandomPolicy(object):
    def get_actions(self, observations):
        return [[np.random.random() + obs / 100 for obs in task] for task in observations], None


def _discount_cumsum(x, discount):
    out, run = np.zeros(len(x)), 0.0
    for t in range(len(x) - 1, -1, -1):
        run = x[t] + discount * run
        out[t] = run
    return out


def _sq_err(baseline, task):
    return sum(float(np.sum(np.square(baseline.predict(path) - path['returns']))) for path in task)


@pytest.mark.gpu
def test_linear_feature_baseline_fit_improves_error_and_pickles():
    """ref tests/test_baselines.py:67-98: fit(paths) reduces the squared error of predict(path) and survives pickling
    bit-identically; plus the coefficients against numpy's lstsq on the same ridge system (reference formula, :66-77)."""
    np.random.seed(3)
    linear = LinearFeatureBaseline()
    sampler = MetaSampler(BaselineRandomEnv(), BaselineRandomPolicy(), 10, 2, 100, parallel=True)
    paths = sampler.obtain_samples()
    for task in paths.values():
        for path in task:
            path["returns"] = _discount_cumsum(path["rewards"], 0.99)
        fresh = LinearFeatureBaseline()
        unfit_error = _sq_err(fresh, task)                 # zeros before the first fit
        linear.fit(task)
        fit_error = _sq_err(linear, task)
        assert fit_error < unfit_error
        # coefficients == the reference's damped least squares (host numpy, float64)
        feat = np.concatenate([linear._features(p) for p in task])
        target = np.concatenate([p['returns'] for p in task])
        want = np.linalg.lstsq(feat.T.dot(feat) + 1e-5 * np.identity(feat.shape[1]), feat.T.dot(target), rcond=-1)[0]
        pred_got = feat.dot(np.asarray(linear.get_param_values()))
        np.testing.assert_allclose(pred_got, feat.dot(want), rtol=1e-6, atol=1e-6 * np.abs(target).max())
        clone = pickle.loads(pickle.dumps(linear))
        assert _sq_err(clone, task) == fit_error
        # target_key other than 'returns'
        for path in task:
            path['other'] = np.asarray(path['rewards'], dtype=np.float64) * 2.0 + 1.0
        other = LinearFeatureBaseline()
        other.fit(task, target_key='other')
        assert sum(float(np.sum(np.square(other.predict(p) - p['other']))) for p in task) < \
            sum(float(np.sum(np.square(p['other']))) for p in task)


@pytest.mark.gpu
def test_baseline_rank_deficient_gram_gives_finite_lstsq_like_fit():
    """with reg_coeff = 0 and collinear features (a constant observation column duplicates the bias column) the
    reference's lstsq returns finite minimum-norm coefficients.  The device solve marks the vanishing pivots rank-deficient
    instead of failing, and its FITTED VALUES (the only thing advantages depend on) equal the lstsq ones."""
    rng = np.random.RandomState(0)
    paths = []
    for _ in range(6):
        obs = np.concatenate([rng.randn(50, 1), np.full((50, 1), 2.0)], axis=1)      # 2nd column constant: c, c^2, 1 collinear
        paths.append(dict(observations=obs, returns=rng.randn(50) + obs[:, 0]))
    b = LinearFeatureBaseline(reg_coeff=0.0)
    b.fit(paths)
    coeffs = np.asarray(b.get_param_values())
    assert np.all(np.isfinite(coeffs))
    feat = np.concatenate([b._features(p) for p in paths])
    target = np.concatenate([p['returns'] for p in paths])
    want = np.linalg.lstsq(feat.T.dot(feat), feat.T.dot(target), rcond=-1)[0]
    np.testing.assert_allclose(feat.dot(coeffs), feat.dot(want), rtol=1e-6, atol=1e-7)
