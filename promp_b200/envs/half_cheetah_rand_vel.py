"""HalfCheetahRandVelEnv - MuJoCo-free analytic surrogate (same dynamics as HalfCheetahRandDirecEnv).

Follows the reference (meta_policy_search/envs/mujoco_envs/half_cheetah_rand_vel.py:7-62) for the interface: tasks =
goal velocity ~ U(0, 3) (:13-14), reward_run = -|forward_vel - goal_velocity|, reward_ctrl = -0.05*sum(a^2) (:30-40),
env_infos {forward_vel, reward_run, reward_ctrl}, obs / reset as in the RandDirec env.  On the device it is reward
mode 1 of the cheetah step functor (promp_b200/csrc/envs.cuh cheetah::step_*).
"""
import numpy as np

from promp_b200.envs.half_cheetah_rand_direc import HalfCheetahRandDirecEnv
from promp_b200.utils import logger


class HalfCheetahRandVelEnv(HalfCheetahRandDirecEnv):
    reward_type = 1                      # device reward mode: -|v - goal|
    info_keys = ('reward_run', 'reward_ctrl', 'forward_vel')

    def __init__(self, goal_velocity=None):
        HalfCheetahRandDirecEnv.__init__(self)
        self.goal_velocity = float(goal_velocity) if goal_velocity is not None else float(self.sample_tasks(1)[0])

    def sample_tasks(self, n_tasks):
        return np.random.uniform(0.0, 3.0, (n_tasks,))         # (:13-14)

    def set_task(self, task):
        self.goal_velocity = task

    def get_task(self):
        return self.goal_velocity

    def log_diagnostics(self, paths, prefix=''):
        """(:55-62) incl. the reference's quirk of logging std of ctrl cost as AvgCtrlCost."""
        phase = getattr(paths[0], 'phase', None) if len(paths) else None
        if phase is not None and phase.info is not None:
            for k, v in zip(self.DEVICE_LOG_KEYS, self.device_log_terms(phase).cpu().numpy()):
                logger.logkv(prefix + k, float(v))
            return
        fwrd_vel = [path["env_infos"]['forward_vel'] for path in paths]
        final_fwrd_vel = [path["env_infos"]['forward_vel'][-1] for path in paths]
        ctrl_cost = [-path["env_infos"]['reward_ctrl'] for path in paths]
        logger.logkv(prefix + 'AvgForwardVel', np.mean(fwrd_vel))
        logger.logkv(prefix + 'AvgFinalForwardVel', np.mean(final_fwrd_vel))
        logger.logkv(prefix + 'AvgCtrlCost', np.std(ctrl_cost))

    def device_log_terms(self, phase):
        import torch
        vel, ctrl = phase.info[2], phase.info[1]
        return torch.stack([vel.mean(), vel.reshape(-1, phase.H)[:, -1].mean(), torch.std(-ctrl, unbiased=False)]).double()

    def __str__(self):
        return 'HalfCheetahRandVelEnv'
