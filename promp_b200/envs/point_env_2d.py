"""MetaPointEnv (ref: meta_policy_search/envs/point_envs/point_env_2d.py:7-71): origin goal, early
`done`, no task.  Early termination gives variable-length paths, so this env is served by the
stepwise device executor (promp_env_step), not by the fused fixed-horizon rollout kernel."""
import numpy as np

from promp_b200 import _lib
from promp_b200.envs.base import MetaEnv, Box


class MetaPointEnv(MetaEnv):
    env_kind = _lib.ENV_POINT
    obs_dim = 2
    act_dim = 2

    @property
    def observation_space(self):
        return Box(low=-np.inf, high=np.inf, shape=(2,))

    @property
    def action_space(self):
        return Box(low=-0.1, high=0.1, shape=(2,))

    def sample_tasks(self, n_tasks):
        return [{}] * n_tasks

    def set_task(self, task):
        pass

    def get_task(self):
        return {}

    def task_vector(self, task):
        return np.zeros(1, dtype=np.float32)

    def host_reset_states(self, n):
        return np.random.uniform(-2, 2, size=(n, 2))       # reset (:27-36)

    def log_diagnostics(self, *args, **kwargs):
        pass
