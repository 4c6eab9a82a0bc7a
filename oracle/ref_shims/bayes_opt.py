"""``bayes_opt`` is imported at module level by the reference's hyperparameter_optimization.py; only ``ParamAgent`` (which
does not use it) is exercised over these shims."""


class BayesianOptimization:
    def __init__(self, *a, **k):
        raise NotImplementedError("bayes_opt is a control-plane dependency outside the hot path")


class UtilityFunction:
    def __init__(self, *a, **k):
        raise NotImplementedError("bayes_opt is a control-plane dependency outside the hot path")
