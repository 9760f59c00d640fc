"""Observation models -- same interface as the reference's ``stag.likelihoods``
(stag/likelihoods.py:1-38): the network output is read as ``probs`` of a Categorical
(node classification) or Bernoulli (multi-label, PPI) distribution."""
import abc

import torch


class Likelihood(torch.nn.Module, abc.ABC):
    def __init__(self, distribution):
        super().__init__()
        self.distribution = distribution

    @abc.abstractmethod
    def condition(self, feat):
        raise NotImplementedError

    def log_prob(self, feat, y):
        return self.condition(feat).log_prob(y)


class CategoricalLikelihood(Likelihood):
    def __init__(self):
        super().__init__(distribution=torch.distributions.Categorical)

    def condition(self, feat):
        return self.distribution(probs=feat)


class BernoulliLikelihood(Likelihood):
    def __init__(self):
        super().__init__(distribution=torch.distributions.Bernoulli)

    def condition(self, feat):
        return self.distribution(probs=feat)
