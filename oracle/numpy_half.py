"""CPU restatement of the reference's numpy half.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Pinned against the unmodified reference by tests/test_oracle_golden.py (fixtures written by
oracle/make_golden.py).  All `ref:` citations are relative to /root/reference/meta_policy_search/.
The granularity (one Python object per env, per-sample appends) deliberately mirrors the
reference so that timing this module is a fair stand-in for timing the reference on the GPU box,
where /root/reference does not exist.
"""
from collections import OrderedDict

import numpy as np
import scipy.signal

CORNERS = ((-2, -2), (2, -2), (-2, 2), (2, 2))   # ref: envs/point_envs/point_env_2d_corner.py:18


# --------------------------------------------------------------------------------------- envs
class PointEnvCorner(object):
    """ref: envs/point_envs/point_env_2d_corner.py:7-93 (MetaPointEnvCorner)."""
    obs_dim = 2
    act_dim = 2
    # gym 0.10.5 Box defaults to float32 bounds (ref :20 passes no dtype)
    action_low = np.full(2, -0.2, dtype=np.float32)
    action_high = np.full(2, 0.2, dtype=np.float32)

    def __init__(self, reward_type='sparse', sparse_reward_radius=0.5):
        assert reward_type in ('dense', 'dense_squared', 'sparse')
        self.reward_type = reward_type
        self.sparse_reward_radius = sparse_reward_radius
        self.corners = [np.array(c) for c in CORNERS]
        self.goal = None
        self._state = None

    def sample_tasks(self, n_tasks):                       # ref :86-87
        return [self.corners[i] for i in np.random.choice(range(4), size=n_tasks)]

    def set_task(self, task):
        self.goal = task

    def get_task(self):
        return self.goal

    def reset(self):                                       # ref :43-52
        self._state = np.random.uniform(-0.2, 0.2, size=(2,))
        return self._state.copy()

    def reward(self, prev, nxt):                           # ref :61-81
        g = np.sqrt(np.sum((nxt - self.goal) ** 2))
        if self.reward_type == 'dense':
            return -g
        if self.reward_type == 'dense_squared':
            return -g ** 2
        if np.sum(np.abs(nxt)) < self.sparse_reward_radius:
            return 0
        nearest = min(np.sqrt(np.sum((nxt - c) ** 2)) for c in self.corners)
        if g == nearest:
            return np.sqrt(np.sum((prev - self.goal) ** 2)) - g
        return 0

    def step(self, action):                                # ref :22-41
        prev = self._state
        self._state = prev + np.clip(action, -0.2, 0.2)
        return self._state.copy(), self.reward(prev, self._state), False, {}

    def log_diagnostics(self, *a, **k):
        pass


class PointEnvWalls(PointEnvCorner):
    """ref: envs/point_envs/point_env_2d_walls.py:7-117 (MetaPointEnvWalls): two circular walls (radius 1 and 2), each
    with one gap of radius 1 around gap_k; tasks = {goal corner, gap_1 on the unit circle, gap_2 on the radius-2 circle}."""

    def __init__(self, reward_type='dense', sparse_reward_radius=2):
        PointEnvCorner.__init__(self, reward_type, sparse_reward_radius)
        self.gap_1 = self.gap_2 = None

    def sample_tasks(self, n_tasks):                       # ref :102-108 (one choice draw, then two normal draws)
        goals = [self.corners[i] for i in np.random.choice(range(4), size=n_tasks)]
        g1 = np.random.normal(size=(n_tasks, 2))
        g1 /= np.linalg.norm(g1, axis=1)[..., np.newaxis]
        g2 = np.random.normal(size=(n_tasks, 2))
        g2 /= (np.linalg.norm(g2, axis=1) / 2)[..., np.newaxis]
        return [dict(goal=a, gap_1=b, gap_2=c) for a, b, c in zip(goals, g1, g2)]

    def set_task(self, task):
        self.goal, self.gap_1, self.gap_2 = task['goal'], task['gap_1'], task['gap_2']

    def get_task(self):
        return dict(goal=self.goal, gap_1=self.gap_1, gap_2=self.gap_2)

    def reward(self, prev, nxt):                           # ref :75-90 (the 'sparse' branch yields None outside the radius)
        g = np.sqrt(np.sum((nxt - self.goal) ** 2))
        if self.reward_type == 'dense':
            return -g
        if self.reward_type == 'dense_squared':
            return -g ** 2
        return (np.sqrt(np.sum((prev - self.goal) ** 2)) - g) if g < self.sparse_reward_radius else None

    def step(self, action):                                # ref :22-51: reward first, then the wall projection
        prev = self._state
        nxt = prev + np.clip(action, -0.2, 0.2)
        r = self.reward(prev, nxt)
        pn, nn = np.linalg.norm(prev), np.linalg.norm(nxt)
        if pn < 1 and nn > 1:
            if np.linalg.norm(nxt - self.gap_1) > 1:
                nxt = nxt / (nn + 1e-6)
        elif pn < 2 and nn > 2:
            if np.linalg.norm(nxt - self.gap_2) > 1:
                nxt = nxt / (nn * 0.5 + 1e-6)
        self._state = nxt
        return self._state.copy(), r, False, {}


class PointEnvMomentum(PointEnvCorner):
    """ref: envs/point_envs/point_env_2d_momentum.py:7-88 (MetaPointEnvMomentum): the action accelerates the point;
    obs = (position, velocity); sparse reward = max(radius - goal distance, 0)."""
    obs_dim = 4
    action_low = np.full(2, -0.1, dtype=np.float32)
    action_high = np.full(2, 0.1, dtype=np.float32)

    def __init__(self, reward_type='sparse', sparse_reward_radius=2):
        PointEnvCorner.__init__(self, reward_type, sparse_reward_radius)
        self._velocity = None

    def reset(self):                                       # ref :44-54: position draw, then velocity draw
        self._state = np.random.uniform(-0.2, 0.2, size=(2,))
        self._velocity = np.random.uniform(-0.1, 0.1, size=(2,))
        return np.hstack((self._state, self._velocity))

    def reward(self, prev, nxt):                           # ref :63-72
        g = np.sqrt(np.sum((nxt - self.goal) ** 2))
        if self.reward_type == 'dense':
            return -g
        if self.reward_type == 'dense_squared':
            return -g ** 2
        return np.maximum(self.sparse_reward_radius - g, 0)

    def step(self, action):                                # ref :22-42
        prev = self._state
        self._velocity = self._velocity + np.clip(action, -0.1, 0.1)
        self._state = prev + self._velocity
        return np.hstack((self._state, self._velocity)), self.reward(prev, self._state), False, {}


class PointEnv(object):
    """ref: envs/point_envs/point_env_2d.py:7-71 (MetaPointEnv; early `done`, no task)."""
    obs_dim = 2
    act_dim = 2
    action_low = np.full(2, -0.1, dtype=np.float32)
    action_high = np.full(2, 0.1, dtype=np.float32)

    def sample_tasks(self, n_tasks):
        return [{}] * n_tasks

    def set_task(self, task):
        pass

    def get_task(self):
        return {}

    def reset(self):                                       # ref :27-36
        self._state = np.random.uniform(-2, 2, size=(2,))
        return self._state.copy()

    def step(self, action):                                # ref :9-25, :46-59
        self._state = self._state + np.clip(action, -0.1, 0.1)
        s = self._state
        reward = -np.sqrt(s[0] ** 2 + s[1] ** 2)
        done = bool(abs(s[0]) < 0.01 and abs(s[1]) < 0.01)
        return s.copy(), reward, done, {}

    def log_diagnostics(self, *a, **k):
        pass


class NormalizedEnv(object):
    """ref: envs/normalized_env.py:6-126 with the defaults (no obs / reward normalisation)."""

    def __init__(self, env, normalization_scale=10.):
        self._wrapped_env = env
        self._scale = normalization_scale
        self.obs_dim = env.obs_dim
        self.act_dim = env.act_dim

    def __getattr__(self, name):
        if name.startswith('__') or name == '_wrapped_env':
            raise AttributeError(name)
        return getattr(self._wrapped_env, name)

    def reset(self):
        return self._wrapped_env.reset()

    def rescale(self, action):                             # ref :109-117
        lb, ub = self._wrapped_env.action_low, self._wrapped_env.action_high
        scaled = lb + (action + self._scale) * (ub - lb) / (2 * self._scale)
        return np.clip(scaled, lb, ub)

    def step(self, action):
        return self._wrapped_env.step(self.rescale(action))


# ------------------------------------------------------------------------------------ sampler
class IterativeEnvExecutor(object):
    """ref: samplers/vectorized_env_executor.py:7-85 (MetaIterativeEnvExecutor)."""

    def __init__(self, env, meta_batch_size, envs_per_task, max_path_length):
        import copy
        self.envs = [copy.deepcopy(env) for _ in range(meta_batch_size * envs_per_task)]
        self.ts = np.zeros(len(self.envs), dtype=int)
        self.max_path_length = max_path_length
        self.meta_batch_size = meta_batch_size

    @property
    def num_envs(self):
        return len(self.envs)

    def set_tasks(self, tasks):                            # ref :54-64
        per = len(self.envs) // len(tasks)
        for i, env in enumerate(self.envs):
            env.set_task(tasks[i // per])

    def reset(self):                                       # ref :66-75
        obs = [env.reset() for env in self.envs]
        self.ts[:] = 0
        return obs

    def step(self, actions):                               # ref :25-52
        results = [env.step(a) for a, env in zip(actions, self.envs)]
        obs, rewards, dones, infos = [list(x) for x in zip(*results)]
        self.ts += 1
        dones = np.logical_or(self.ts >= self.max_path_length, np.asarray(dones))
        for i in np.flatnonzero(dones):
            obs[i] = self.envs[i].reset()
            self.ts[i] = 0
        return obs, rewards, dones, infos


def _stack_dicts(dict_list):
    """ref: utils/utils.py:144-159 (stack_tensor_dict_list)."""
    if not dict_list or not dict_list[0]:
        return {}
    return {k: (_stack_dicts([d[k] for d in dict_list]) if isinstance(dict_list[0][k], dict)
                else np.asarray([d[k] for d in dict_list])) for k in dict_list[0]}


def _concat_dicts(dict_list):
    """ref: utils/utils.py:104-141 (concat_tensor_dict_list)."""
    if not dict_list or not dict_list[0]:
        return {}
    return {k: (_concat_dicts([d[k] for d in dict_list]) if isinstance(dict_list[0][k], dict)
                else np.concatenate([d[k] for d in dict_list])) for k in dict_list[0]}


class Sampler(object):
    """ref: samplers/meta_sampler.py:12-150 (MetaSampler, iterative executor only)."""

    def __init__(self, env, policy, rollouts_per_meta_task, meta_batch_size, max_path_length, envs_per_task=None):
        self.env, self.policy = env, policy
        self.envs_per_task = rollouts_per_meta_task if envs_per_task is None else envs_per_task
        self.meta_batch_size = meta_batch_size
        self.max_path_length = max_path_length
        self.total_samples = meta_batch_size * rollouts_per_meta_task * max_path_length
        self.total_timesteps_sampled = 0
        self.vec_env = IterativeEnvExecutor(env, meta_batch_size, self.envs_per_task, max_path_length)

    def update_tasks(self):                                # ref :51-57
        tasks = self.env.sample_tasks(self.meta_batch_size)
        self.vec_env.set_tasks(tasks)
        return tasks

    def obtain_samples(self):                              # ref :59-137
        M, E = self.meta_batch_size, self.envs_per_task
        paths = OrderedDict((i, []) for i in range(M))
        running = [dict(observations=[], actions=[], rewards=[], env_infos=[], agent_infos=[])
                   for _ in range(M * E)]
        n_samples = 0
        obses = self.vec_env.reset()
        while n_samples < self.total_samples:
            actions, agent_infos = self.policy.get_actions(np.split(np.asarray(obses), M))
            actions = np.concatenate(actions)
            next_obses, rewards, dones, env_infos = self.vec_env.step(actions)
            agent_infos = sum(agent_infos, []) if agent_infos else [dict() for _ in range(M * E)]
            for idx in range(M * E):
                r = running[idx]
                r["observations"].append(obses[idx])
                r["actions"].append(actions[idx])
                r["rewards"].append(rewards[idx])
                r["env_infos"].append(env_infos[idx])
                r["agent_infos"].append(agent_infos[idx])
                if dones[idx]:
                    paths[idx // E].append(dict(
                        observations=np.asarray(r["observations"]),
                        actions=np.asarray(r["actions"]),
                        rewards=np.asarray(r["rewards"]),
                        env_infos=_stack_dicts(r["env_infos"]),
                        agent_infos=_stack_dicts(r["agent_infos"])))
                    n_samples += len(r["rewards"])
                    running[idx] = dict(observations=[], actions=[], rewards=[], env_infos=[], agent_infos=[])
            obses = next_obses
        self.total_timesteps_sampled += self.total_samples
        return paths


# ---------------------------------------------------------------------------- sample processing
def discount_cumsum(x, discount):
    """ref: utils/utils.py:74-81.  y[t] = x[t] + discount*y[t+1]."""
    return scipy.signal.lfilter([1], [1, float(-discount)], x[::-1], axis=0)[::-1]


def baseline_features(observations):
    """ref: baselines/linear_baseline.py:101-106 (LinearFeatureBaseline._features)."""
    obs = np.clip(observations, -10, 10)
    n = len(observations)
    t = np.arange(n).reshape(-1, 1) / 100.0
    return np.concatenate([obs, obs ** 2, t, t ** 2, t ** 3, np.ones((n, 1))], axis=1)


class LinearFeatureBaseline(object):
    """ref: baselines/linear_baseline.py:6-106."""

    def __init__(self, reg_coeff=1e-5):
        self._coeffs = None
        self._reg_coeff = reg_coeff

    def fit(self, paths, target_key='returns'):            # ref :55-77
        featmat = np.concatenate([baseline_features(p["observations"]) for p in paths], axis=0)
        target = np.concatenate([p[target_key] for p in paths], axis=0)
        reg = self._reg_coeff
        for _ in range(5):
            self._coeffs = np.linalg.lstsq(featmat.T.dot(featmat) + reg * np.identity(featmat.shape[1]),
                                           featmat.T.dot(target), rcond=-1)[0]
            if not np.any(np.isnan(self._coeffs)):
                break
            reg *= 10

    def predict(self, path):                               # ref :17-33
        if self._coeffs is None:
            return np.zeros(len(path["observations"]))
        return baseline_features(path["observations"]).dot(self._coeffs)


class ZeroBaseline(object):
    """ref: baselines/zero_baseline.py:5-55."""

    def fit(self, paths, **kw):
        pass

    def predict(self, path):
        return np.zeros_like(path["rewards"])


class SampleProcessor(object):
    """ref: samplers/base.py:33-173 + samplers/meta_sample_processor.py:6-49."""

    def __init__(self, baseline, discount=0.99, gae_lambda=1, normalize_adv=False, positive_adv=False):
        self.baseline, self.discount, self.gae_lambda = baseline, discount, gae_lambda
        self.normalize_adv, self.positive_adv = normalize_adv, positive_adv

    def compute_samples_data(self, paths):                 # ref samplers/base.py:99-133
        for p in paths:
            p["returns"] = discount_cumsum(p["rewards"], self.discount)
        self.baseline.fit(paths, target_key="returns")
        for p in paths:                                    # ref :151-162
            b = np.append(self.baseline.predict(p), 0)
            deltas = p["rewards"] + self.discount * b[1:] - b[:-1]
            p["advantages"] = discount_cumsum(deltas, self.discount * self.gae_lambda)
        data = dict(                                       # ref :165-173
            observations=np.concatenate([p["observations"] for p in paths]),
            actions=np.concatenate([p["actions"] for p in paths]),
            rewards=np.concatenate([p["rewards"] for p in paths]),
            returns=np.concatenate([p["returns"] for p in paths]),
            advantages=np.concatenate([p["advantages"] for p in paths]),
            env_infos=_concat_dicts([p["env_infos"] for p in paths]),
            agent_infos=_concat_dicts([p["agent_infos"] for p in paths]))
        adv = data["advantages"]
        if self.normalize_adv:                             # ref utils/utils.py:59-67
            adv = (adv - np.mean(adv)) / (adv.std() + 1e-8)
        if self.positive_adv:                              # ref utils/utils.py:70-71
            adv = (adv - np.min(adv)) + 1e-8
        data["advantages"] = adv
        return data, paths

    def process_samples(self, paths_meta_batch):           # ref samplers/meta_sample_processor.py:8-49
        out, all_paths = [], []
        for _, paths in paths_meta_batch.items():
            data, paths = self.compute_samples_data(paths)
            out.append(data)
            all_paths.extend(paths)
        all_r = np.concatenate([d['rewards'] for d in out])
        mean, std = np.mean(all_r), np.std(all_r)
        for d in out:
            d['adj_avg_rewards'] = (d['rewards'] - mean) / (std + 1e-8)
        self.last_stats = path_stats(all_paths)
        return out


def path_stats(paths):
    """ref: samplers/base.py:135-149 (_log_path_stats) as a dict instead of logger calls."""
    undisc = [sum(p["rewards"]) for p in paths]
    return dict(AverageDiscountedReturn=np.mean([p["returns"][0] for p in paths]),
                AverageReturn=np.mean(undisc), NumTrajs=len(paths), StdReturn=np.std(undisc),
                MaxReturn=np.max(undisc), MinReturn=np.min(undisc))
