from .vec_env import VecEvacuationEnv  # noqa: F401
