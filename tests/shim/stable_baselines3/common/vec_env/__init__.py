from .dummy_vec_env import DummyVecEnv  # noqa: F401


class VecNormalize:
    @staticmethod
    def load(path, venv):
        raise NotImplementedError("shim: VecNormalize is import-only here")
