from . import MujocoEnv  # noqa: F401
