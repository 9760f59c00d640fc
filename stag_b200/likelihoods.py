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

    # kind of the fused epilogue kernel (stag_nll) that evaluates this likelihood, or None
    fused_kind = None

    def nll_samples(self, feats, y, mask=None):
        """``[S]`` masked mean NLLs of the sample outputs ``feats [S,N,C]`` (stag/models.py:69-72 per sample).
        CUDA tensors of the two probs-likelihoods take one fused kernel pass (forward + gradient); anything else
        is evaluated through torch.distributions like the reference."""
        usable = (self.fused_kind is not None and feats.is_cuda and feats.dim() == 3
                  and (mask is None or (mask.dtype == torch.bool and mask.dim() == 1))
                  and ((self.fused_kind == "categorical" and y.dim() == 1)
                       or (self.fused_kind == "bernoulli" and y.dim() == 2)))
        if usable:
            from . import ops
            return ops.fused_nll(feats, y, mask, self.fused_kind)
        return torch.stack([self.nll(f, y, mask) for f in feats.unbind(0)])


def _probs_likelihood(name, family, doc, fused_kind=None):
    """A Likelihood whose network output parameterises ``family(probs=...)``."""

    def __init__(self):
        Likelihood.__init__(self, distribution=family)

    def condition(self, feat):
        return self.distribution(probs=feat)

    return type(name, (Likelihood,), {"__init__": __init__, "condition": condition, "__doc__": doc,
                                      "__module__": __name__, "fused_kind": fused_kind})


CategoricalLikelihood = _probs_likelihood(
    "CategoricalLikelihood", torch.distributions.Categorical,
    "feat [N,C] = class probabilities (torch renormalises them), y [N] int64 (stag/likelihoods.py:18-27).",
    fused_kind="categorical")
BernoulliLikelihood = _probs_likelihood(
    "BernoulliLikelihood", torch.distributions.Bernoulli,
    "feat [N,C] = per-label probabilities, y [N,C] in {0,1} (stag/likelihoods.py:29-38).",
    fused_kind="bernoulli")
