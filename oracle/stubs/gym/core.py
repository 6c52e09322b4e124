class Env(object):
    metadata = {}
    reward_range = (-float('inf'), float('inf'))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError
