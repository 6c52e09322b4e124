"""HalfCheetahRandDirecEnv - MuJoCo-free analytic surrogate.

Follows the reference (meta_policy_search/envs/mujoco_envs/half_cheetah_rand_direc.py:8-68) for the
interface: obs 17 (qpos[1:] ++ qvel), action 6 in [-1,1], reward = reward_ctrl + reward_run with
reward_ctrl = -0.05*sum(a^2), reward_run = direction*(x_after-x_before)/dt, dt = 0.05, done = False,
env_infos {reward_run, reward_ctrl}, tasks = direction in {-1,+1}, reset noise U(-.1,.1)^9 / .1*N(0,1)^9.
The dynamics are this repo's analytic model (CUDA: promp_b200/csrc/envs.cuh cheetah::).
"""
import numpy as np

from promp_b200 import _lib
from promp_b200.envs.base import MetaEnv, Box
from promp_b200.utils import logger


class HalfCheetahRandDirecEnv(MetaEnv):
    env_kind = _lib.ENV_CHEETAH_DIR
    reward_type = 0                      # device reward mode: direction * forward_vel
    info_keys = ('reward_run', 'reward_ctrl')
    obs_dim = 17
    act_dim = 6

    def __init__(self, goal_direction=None):
        self.goal_direction = goal_direction if goal_direction else 1.0
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(17,))
        self.action_space = Box(low=-1.0, high=1.0, shape=(6,))

    def sample_tasks(self, n_tasks):
        return np.random.choice((-1.0, 1.0), (n_tasks,))       # (:14-16)

    def set_task(self, task):
        self.goal_direction = task

    def get_task(self):
        return self.goal_direction

    def task_vector(self, task):
        return np.asarray([task], dtype=np.float32)

    def host_reset_states(self, n):
        """reset_model (:49-53): qpos = init + U(-.1,.1)^9, qvel = init + .1*N(0,1)^9 for each env.  The reference
        draws these from each env's own gym `np_random` stream (not the global numpy RNG), so there is no global
        draw order to reproduce: the n envs are drawn vectorised from the global RNG."""
        out = np.empty((n, 18))
        out[:, :9] = np.random.uniform(low=-.1, high=.1, size=(n, 9))
        out[:, 9:] = np.random.randn(n, 9) * .1
        return out

    def log_diagnostics(self, paths, prefix=''):
        """(:58-65) incl. the reference's quirk of logging std of ctrl cost as AvgCtrlCost."""
        phase = getattr(paths[0], 'phase', None) if len(paths) else None
        if phase is not None and phase.info is not None:
            import torch
            run, ctrl = phase.info[0], phase.info[1]
            H = phase.H
            logger.logkv(prefix + 'AvgForwardVel', float(run.mean()))
            logger.logkv(prefix + 'AvgFinalForwardVel', float(run.reshape(-1, H)[:, -1].mean()))
            logger.logkv(prefix + 'AvgCtrlCost', float(torch.std(-ctrl, unbiased=False)))
            return
        fwrd_vel = [path["env_infos"]['reward_run'] for path in paths]
        final_fwrd_vel = [path["env_infos"]['reward_run'][-1] for path in paths]
        ctrl_cost = [-path["env_infos"]['reward_ctrl'] for path in paths]
        logger.logkv(prefix + 'AvgForwardVel', np.mean(fwrd_vel))
        logger.logkv(prefix + 'AvgFinalForwardVel', np.mean(final_fwrd_vel))
        logger.logkv(prefix + 'AvgCtrlCost', np.std(ctrl_cost))

    DEVICE_LOG_KEYS = ('AvgForwardVel', 'AvgFinalForwardVel', 'AvgCtrlCost')

    def device_log_terms(self, phase):
        import torch
        run, ctrl = phase.info[0], phase.info[1]
        return torch.stack([run.mean(), run.reshape(-1, phase.H)[:, -1].mean(), torch.std(-ctrl, unbiased=False)]).double()

    def __str__(self):
        return 'HalfCheetahRandDirecEnv'
