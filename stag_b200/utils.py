"""Training utilities -- ``EarlyStopping`` with the reference's call contract
(stag/utils.py:1-26): ``stop = early_stopping(losses, model)``; the best
``state_dict`` is deep-copied when every tracked loss improves."""
import copy


class EarlyStopping(object):
    best_losses = None
    best_state = None
    counter = 0

    def __init__(self, patience=10):
        self.patience = patience

    def __call__(self, losses, model):
        if self.best_losses is None:
            self.best_losses = losses
            self.counter = 0
            return False
        improved = [loss <= best for loss, best in zip(losses, self.best_losses)]
        if any(improved):
            if all(improved):
                self.best_state = copy.deepcopy(model.state_dict())
            self.best_losses = [min(loss, best) for loss, best in zip(losses, self.best_losses)]
            self.counter = 0
            return False
        self.counter += 1
        return self.counter == self.patience
