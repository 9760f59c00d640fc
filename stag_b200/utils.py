"""Training utilities.  ``EarlyStopping`` keeps the call contract of the reference
(stag/utils.py:1-26): ``stop = early_stopping(losses, model)`` with ``losses`` a list of tracked
values (lower is better); a snapshot of ``model.state_dict()`` is kept in ``best_state`` whenever
EVERY tracked value is at least as good as its best so far, the patience counter restarts whenever
ANY of them is, and the call returns True once ``patience`` consecutive calls improved nothing."""
import copy


class EarlyStopping:
    best_losses = None   # per-value running minima
    best_state = None    # deep copy of the state_dict at the last all-improve call
    counter = 0          # consecutive calls without any improvement

    def __init__(self, patience=10):
        self.patience = patience

    def __call__(self, losses, model):
        losses = list(losses)
        if self.best_losses is None:          # first call only records the values
            self.best_losses, self.counter = losses, 0
            return False
        better = [new <= old for new, old in zip(losses, self.best_losses)]
        if not any(better):
            self.counter += 1
            return self.counter == self.patience
        if all(better):
            self.best_state = copy.deepcopy(model.state_dict())
        self.best_losses = [min(pair) for pair in zip(losses, self.best_losses)]
        self.counter = 0
        return False
