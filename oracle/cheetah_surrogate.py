"""Specification + CPU restatement of the MuJoCo-free HalfCheetahRandDirec surrogate.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

What follows the reference (meta_policy_search/envs/mujoco_envs/half_cheetah_rand_direc.py):
  * obs = qpos[1:] (8) ++ qvel (9) = 17 floats, action = 6 floats          (:43-47)
  * reward_ctrl = -0.5*0.1*sum(a^2); reward_run = dir*(x_after-x_before)/dt;
    reward = ctrl + run; done = False; env_infos = {reward_run, reward_ctrl} (:32-41)
  * frame_skip = 5 (:11) with MuJoCo's half_cheetah timestep 0.01 => dt = 0.05
  * task = direction in {-1, +1} (:14-16)
  * reset: qpos = init + U(-.1,.1)^9, qvel = init + 0.1*N(0,1)^9           (:49-53)
  * ctrlrange of half_cheetah.xml is [-1, 1]: the NormalizedEnv wrapper maps the policy action
    a -> clip(0.1*a, -1, 1) (envs/normalized_env.py:109-117).

What is NEW (MuJoCo is absent and the north-star asks for an analytic model): the dynamics.
Six actuated joints are damped springs driven by the torques; the root body is pushed by a
smooth "paddling" thrust that depends on joint velocities and the leg angle; height and pitch
are damped springs excited by the legs.  Semi-implicit Euler, 5 sub-steps of h = 0.01.

  for each sub-step:
      for j in 0..5:   acc = G[j]*u[j] - K[j]*q[j] - D[j]*qd[j];  qd[j] += h*acc;  q[j] += h*qd[j]
                        s, c = sin/cos(q[j] + pitch + PH[j])
                        thrust += C[j]*qd[j]*s ;  lift += C[j]*qd[j]*c ;  twist += P[j]*u[j]
      xd  += h*(thrust - BX*xd);                    x     += h*xd
      zd  += h*(LZ*lift - KZ*z - DZ*zd);            z     += h*zd
      pd  += h*(twist - KP*pitch - DP*pd);          pitch += h*pd
(the root update uses the joint state *after* this sub-step's joint update, and pitch from the
previous sub-step inside the sin/cos).

All arithmetic is float32 in the CUDA kernel; this restatement runs in the dtype of its inputs so
tests can evaluate it in float32 (same rounding model) or float64 (reference-style).
"""
import numpy as np

NQ = 9
OBS_DIM = 17
ACT_DIM = 6
FRAME_SKIP = 5
H_SIM = 0.01
DT = FRAME_SKIP * H_SIM

G = (12.0, 9.0, 6.0, 12.0, 6.0, 3.0)          # torque gain
K = (24.0, 18.0, 12.0, 18.0, 12.0, 6.0)       # joint spring
D = (4.5, 3.0, 1.5, 3.0, 1.5, 0.75)           # joint damping
C = (0.9, 0.6, 0.3, -0.8, -0.5, -0.25)        # thrust coupling
PH = (0.3, -0.4, 0.8, -0.3, 0.5, -0.9)        # leg phase offsets
P = (0.6, 0.4, 0.2, -0.6, -0.4, -0.2)         # pitch torque coupling
BX = 1.5
LZ = 0.1
KZ = 40.0
DZ = 6.0
KP = 30.0
DP = 5.0


def reset_state(rng, dtype=np.float64):
    """half_cheetah_rand_direc.py:49-53: uniform qpos noise, then gaussian qvel noise."""
    qpos = rng.uniform(low=-.1, high=.1, size=NQ)
    qvel = rng.randn(NQ) * .1
    return qpos.astype(dtype), qvel.astype(dtype)


def get_obs(qpos, qvel):
    """half_cheetah_rand_direc.py:43-47."""
    return np.concatenate([qpos[..., 1:], qvel], axis=-1)


def step(qpos, qvel, u, direction, goal_velocity=None):
    """One env step on (..., 9) state arrays; `u` is the already-rescaled/clipped action (..., 6).
    Returns (qpos', qvel', reward, reward_run, reward_ctrl).  goal_velocity given: HalfCheetahRandVel reward
    reward_run = -|forward_vel - goal_velocity| (half_cheetah_rand_vel.py:30-40) instead of direction * forward_vel."""
    dt_ = qpos.dtype.type
    qpos = qpos.copy()
    qvel = qvel.copy()
    u = u.astype(qpos.dtype)
    h = dt_(H_SIM)
    x_before = qpos[..., 0].copy()
    for _ in range(FRAME_SKIP):
        thrust = np.zeros_like(x_before)
        lift = np.zeros_like(x_before)
        twist = np.zeros_like(x_before)
        pitch = qpos[..., 2]
        for j in range(6):
            q = qpos[..., 3 + j]
            qd = qvel[..., 3 + j]
            acc = dt_(G[j]) * u[..., j] - dt_(K[j]) * q - dt_(D[j]) * qd
            qd = qd + h * acc
            q = q + h * qd
            qvel[..., 3 + j] = qd
            qpos[..., 3 + j] = q
            ang = q + pitch + dt_(PH[j])
            thrust = thrust + dt_(C[j]) * qd * np.sin(ang)
            lift = lift + dt_(C[j]) * qd * np.cos(ang)
            twist = twist + dt_(P[j]) * u[..., j]
        xd = qvel[..., 0] + h * (thrust - dt_(BX) * qvel[..., 0])
        qvel[..., 0] = xd
        qpos[..., 0] = qpos[..., 0] + h * xd
        zd = qvel[..., 1] + h * (dt_(LZ) * lift - dt_(KZ) * qpos[..., 1] - dt_(DZ) * qvel[..., 1])
        qvel[..., 1] = zd
        qpos[..., 1] = qpos[..., 1] + h * zd
        pd = qvel[..., 2] + h * (twist - dt_(KP) * qpos[..., 2] - dt_(DP) * qvel[..., 2])
        qvel[..., 2] = pd
        qpos[..., 2] = qpos[..., 2] + h * pd
    reward_ctrl = -dt_(0.05) * np.sum(np.square(u), axis=-1)
    forward_vel = (qpos[..., 0] - x_before) / dt_(DT)
    if goal_velocity is None:
        reward_run = np.asarray(direction, dtype=qpos.dtype) * forward_vel
    else:
        reward_run = -np.abs(forward_vel - np.asarray(goal_velocity, dtype=qpos.dtype))
    return qpos, qvel, reward_ctrl + reward_run, reward_run, reward_ctrl


class HalfCheetahRandDirecSurrogate(object):
    """Per-env object with the reference MetaEnv surface (sample_tasks/set_task/get_task/reset/step)."""
    obs_dim = OBS_DIM
    act_dim = ACT_DIM
    action_low = -np.ones(ACT_DIM, dtype=np.float32)
    action_high = np.ones(ACT_DIM, dtype=np.float32)

    def __init__(self, goal_direction=None):
        self.goal_direction = goal_direction if goal_direction else 1.0
        self.qpos = np.zeros(NQ)
        self.qvel = np.zeros(NQ)

    def sample_tasks(self, n_tasks):
        return np.random.choice((-1.0, 1.0), (n_tasks,))

    def set_task(self, task):
        self.goal_direction = task

    def get_task(self):
        return self.goal_direction

    def reset(self):
        self.qpos, self.qvel = reset_state(np.random)
        return get_obs(self.qpos, self.qvel)

    def step(self, action):
        self.qpos, self.qvel, r, rr, rc = step(self.qpos, self.qvel, np.asarray(action, dtype=np.float64),
                                               self.goal_direction)
        return get_obs(self.qpos, self.qvel), float(r), False, dict(reward_run=float(rr), reward_ctrl=float(rc))

    def log_diagnostics(self, paths, prefix=''):
        pass
