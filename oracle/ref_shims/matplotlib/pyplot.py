"""No-op stand-in for the plotting calls of ``Agent._plot`` (General/QLearning/q_agent.py:233-246)."""


class _Figure:
    def show(self):
        pass


def figure(*a, **k):
    return _Figure()


def plot(*a, **k):
    pass


def xlabel(*a, **k):
    pass


def ylabel(*a, **k):
    pass


def show(*a, **k):
    pass
