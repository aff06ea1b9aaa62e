"""actorcritic/multi_env.py: `MultiEnv` - the vectorised environment the agent steps.  API shape and auto-reset
semantics only; the reference's subprocess / pipe hosting of emulators (multi_env.py:140-362) is
outside the accelerated path, so `create_subprocess_envs` is not provided."""
import concurrent.futures


class _AutoResetWrapper:
    """multi_env.py:121-137: an environment that was terminal at the previous step is reset first; the reset
    observation is discarded and the step is taken with the action chosen from the terminal observation."""

    def __init__(self, env):
        self.env = env
        self._terminated = False

    def __getattr__(self, name):
        return getattr(self.env, name)

    def step(self, action):
        if self._terminated:
            self.env.reset()
        observation, reward, terminal, info = self.env.step(action)
        self._terminated = terminal
        return observation, reward, terminal, info

    def reset(self, **kwargs):
        observation = self.env.reset(**kwargs)
        self._terminated = False
        return observation


class MultiEnv:
    """multi_env.py:11-89."""

    def __init__(self, envs):
        self._envs = [_AutoResetWrapper(env) for env in envs]
        self._executor = concurrent.futures.ThreadPoolExecutor(len(self._envs))

    envs = property(lambda self: self._envs)
    observation_space = property(lambda self: self._envs[0].observation_space)
    action_space = property(lambda self: self._envs[0].action_space)

    def reset(self):
        return list(self._executor.map(lambda env: env.reset(), self._envs))

    def step(self, actions):
        """multi_env.py:59-81: an action of None skips that environment and yields (None, None, None, None)."""
        def call_step(env_action):
            env, action = env_action
            if action is None:
                return None, None, None, None
            return env.step(action)
        results = list(self._executor.map(call_step, zip(self._envs, actions)))
        observations, rewards, terminals, infos = (list(x) for x in zip(*results))
        return observations, rewards, terminals, infos

    def close(self):
        for env in self._envs:
            if hasattr(env.env, "close"):
                env.env.close()
        self._executor.shutdown()


def create_subprocess_envs(env_fns):
    raise NotImplementedError("hosting emulators in subprocesses (multi_env.py:92-362) is outside this framework's scope; "
                              "build the environments in-process or use envs.atari.device_env.DeviceAtariMultiEnv")
