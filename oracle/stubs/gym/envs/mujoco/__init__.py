class MujocoEnv(object):
    """Placeholder: MuJoCo is absent; only lets envs/base.py import."""
    def __init__(self, *a, **k):
        raise RuntimeError("MuJoCo is not available in this container")
