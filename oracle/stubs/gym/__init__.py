"""Minimal stand-in for gym==0.10.5 (absent here; no network): just enough surface for the
reference's numpy half (envs/point_envs, envs/normalized_env.py, envs/base.py) to import
unmodified from /root/reference when oracle/make_golden.py generates fixtures.
TEST INFRASTRUCTURE ONLY - never imported by promp_b200."""
from . import spaces, core, utils  # noqa: F401
from .core import Env  # noqa: F401
