class DummyVecEnv:
    def __init__(self, env_fns):
        self.envs = [f() for f in env_fns]
