"""Observation models with the interface of the reference's ``stag.likelihoods``
(stag/likelihoods.py:4-38): ``Likelihood(distribution)`` with ``condition(feat)`` -> a torch
distribution over the labels and ``log_prob(feat, y)``; the network output is read as ``probs`` of a
Categorical (node classification) or of independent Bernoullis (multi-label, PPI).

The two concrete likelihoods are produced by one factory, since they differ only in the distribution
family; ``nll`` adds the masked mean the models take of ``-log_prob`` (stag/models.py:69-72).
"""
import abc

import torch


class Likelihood(torch.nn.Module, abc.ABC):
    """``distribution``: the torch.distributions class instantiated by :meth:`condition`."""

    def __init__(self, distribution):
        torch.nn.Module.__init__(self)
        self.distribution = distribution

    @abc.abstractmethod
    def condition(self, feat):
        """The label distribution given the network output ``feat``."""

    def log_prob(self, feat, y):
        return self.condition(feat).log_prob(y)

    def nll(self, feat, y, mask=None):
        """Mean negative log-likelihood over the (masked) nodes."""
        values = -self.log_prob(feat, y)
        return (values if mask is None else values[mask]).mean()


def _probs_likelihood(name, family, doc):
    """A Likelihood whose network output parameterises ``family(probs=...)``."""

    def __init__(self):
        Likelihood.__init__(self, distribution=family)

    def condition(self, feat):
        return self.distribution(probs=feat)

    return type(name, (Likelihood,), {"__init__": __init__, "condition": condition, "__doc__": doc,
                                      "__module__": __name__})


CategoricalLikelihood = _probs_likelihood(
    "CategoricalLikelihood", torch.distributions.Categorical,
    "feat [N,C] = class probabilities (torch renormalises them), y [N] int64 (stag/likelihoods.py:18-27).")
BernoulliLikelihood = _probs_likelihood(
    "BernoulliLikelihood", torch.distributions.Bernoulli,
    "feat [N,C] = per-label probabilities, y [N,C] in {0,1} (stag/likelihoods.py:29-38).")
