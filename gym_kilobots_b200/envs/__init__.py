from .vec_env import KilobotsVecEnv  # noqa: F401
