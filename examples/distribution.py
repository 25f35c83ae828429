#!/usr/bin/env python
"""BASELINE config 1 on the B200 engine: the reference's Gaussian-distribution example
(example/distribution.py: 10 groups x 10 responses, parameters a, b, c, 2 chains x 1,000 iterations).

    python examples/distribution.py [partial|none|complete]

Every parameter j has a group mean mu_j[g] ~ N(0, 1) and one sd_j ~ Gamma(1); an observation of group g
contributes sum_j log N(theta_j | mu_j[g], sd_j).  The numbers come off numpy's global stream seeded
with 12345 in the reference's order (per name: the 10 means, then the sd), so the data are the
reference's; the objective is the registry's device function instead of a Python closure."""

import numpy
import scipy.stats

from common import Workload, poolingFromCommandLine
from objectives import Objective

numpy.random.seed(12345)

CONFIG = Workload("Example MCMC to sample from Gaussian distribution.", "./example/sample/distribution/",
                  names=("a", "b", "c"), groups=10, responses=10, chains=2, iterations=1000, retained=100,
                  nProcesses=1)


def mockData(config):
    means = numpy.empty((len(config.names), config.groups))
    sds = numpy.empty(len(config.names))
    for j in range(len(config.names)):
        means[j] = numpy.random.normal(0.0, 1.0, config.groups)
        sds[j] = numpy.random.gamma(1)
    truth = "\nTrue value:\n" + "".join("\t%s: {mean: %.2f, var: %.2f}\n" % (name, means[j].mean(), means[j].var())
                                          for j, name in enumerate(config.names))
    return means, sds, truth


if __name__ == "__main__":
    pooling = poolingFromCommandLine(CONFIG.title)
    means, sds, truth = mockData(CONFIG)
    CONFIG.samplerOptions["priorDistribution"] = [scipy.stats.norm(loc=0, scale=1) for _ in CONFIG.names]
    CONFIG.run(pooling, Objective.gaussian_distribution(means, sds, CONFIG.responses), truth)
