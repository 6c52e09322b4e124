"""MetaHostEnvExecutor: vec-env interface (reset / step / set_tasks / num_envs) for environments that are arbitrary
HOST Python objects - the duck-typed fakes of the reference's tests (tests/test_samplers.py:13-67: `TestEnv`,
`RandomEnv` with sample_tasks / set_task / get_task / step / reset) or any user env without a `device_spec()`.

This is NOT a CPU fallback of the product: user env code can only run where Python runs.  promp_b200's own envs are
device descriptions stepped by promp_rollout / promp_env_step; this executor exists so that the reference-shaped
MetaSampler accepts the same inputs the reference's sampler accepts (one env.step call per env per step: slow, like
the reference's MetaIterativeEnvExecutor, vectorized_env_executor.py:7-85).  Everything downstream of the paths
(sample processing, adapt, meta-gradient) still runs on the device.
"""
import copy

import numpy as np


class MetaHostEnvExecutor(object):
    def __init__(self, env, meta_batch_size, envs_per_task, max_path_length):
        self.meta_batch_size, self.envs_per_task = meta_batch_size, envs_per_task
        self.max_path_length = max_path_length
        self.envs = [copy.deepcopy(env) for _ in range(meta_batch_size * envs_per_task)]
        self.steps_taken = np.zeros(len(self.envs), dtype=np.int64)
        self.tasks = None

    @property
    def num_envs(self):
        return len(self.envs)

    def set_tasks(self, tasks):
        """Slot i runs task i // envs_per_task (vectorized_env_executor.py:54-64)."""
        assert len(tasks) == self.meta_batch_size
        self.tasks = list(tasks)
        for slot, env in enumerate(self.envs):
            env.set_task(tasks[slot // self.envs_per_task])

    def reset(self):
        """Every env is reset in slot order (the order in which a shared numpy RNG is consumed, :66-75)."""
        self.steps_taken[:] = 0
        return [np.array(env.reset(), copy=True) for env in self.envs]      # copies: envs may mutate their state in place

    def step(self, actions):
        """One env.step per slot; a slot whose env reports done, or that reached the horizon, is reset right away and
        returns the reset observation (:25-52)."""
        assert len(actions) == self.num_envs
        obs, rewards, dones, infos = [], [], [], []
        for env, action in zip(self.envs, actions):
            o, r, d, info = env.step(action)
            obs.append(np.array(o, copy=True)); rewards.append(r); dones.append(d); infos.append(copy.deepcopy(info))
        self.steps_taken += 1
        finished = np.logical_or(np.asarray(dones, dtype=bool), self.steps_taken >= self.max_path_length)
        for slot in np.flatnonzero(finished):
            obs[slot] = np.array(self.envs[slot].reset(), copy=True)
            self.steps_taken[slot] = 0
        return obs, rewards, finished, infos
