class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    observation_space = None
    action_space = None

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError("accessing private attribute '%s' is prohibited" % name)
        return getattr(self.env, name)
