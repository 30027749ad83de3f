class LearningRate:
    """Step-decay schedule: every `decay_steps` calls the rate is multiplied by `decay_rate`; a call returns the
    rate BEFORE that update (reference: struct/LearningRate.py)."""

    def __init__(self, initial_lr: float, decay_rate: float, decay_steps: int):
        self.lr = initial_lr
        self.decay_rate = decay_rate
        self.decay_steps = decay_steps
        self._calls = 0

    def __call__(self) -> float:
        if self.decay_rate == 0 or self.decay_steps <= 0:
            return self.lr
        current = self.lr
        self._calls += 1
        if self._calls >= self.decay_steps:
            self.lr *= self.decay_rate
            self._calls = 0
        return current
