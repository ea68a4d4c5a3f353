"""Top-level `make_env` module, as in the reference repository root (make_env.py:15):
`from make_env import make_env; env = make_env('simple_spread')`."""
from multiagent_particle_envs_b200.make_env import make_env  # noqa: F401
