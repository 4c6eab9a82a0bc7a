"""``gym`` is imported by the reference for a type hint (``gym.Env``) and ``env.action_space.n``."""


class Env:
    pass
