#!/usr/bin/env python
"""The reference's Gaussian-distribution example (example/distribution.py) on the B200 engine.

    python examples/distribution.py [partial|none|complete]

Same model, data (numpy.random.seed(12345)), chain counts and priors as the reference; the only
change is that the objective is a device-function handle instead of a Python closure."""

import argparse
import os
import sys

import numpy
import scipy.stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mcmc-for-nested-data_b200"))

from posteriorSampling import samplePosterior  # noqa: E402
from sampleDiagnosis import diagnoseSamples  # noqa: E402
from objectives import Objective  # noqa: E402

numpy.random.seed(12345)


def getFunction(parameterName, nGroups, nResponsesPerGroup):
    """example/distribution.py:16-47: mu_j ~ N(0,1) per group, sd_j ~ Gamma(1) per name."""
    trueValueString = "\nTrue value:\n"
    mu, sd = [], []
    for i, name in enumerate(parameterName):
        m = numpy.random.normal(loc=0, scale=1, size=nGroups)
        s = numpy.random.gamma(1)
        mu.append(m)
        sd.append(s)
        trueValueString += "\t%s: {mean: %.2f, var: %.2f}\n" % (name, numpy.mean(m), numpy.var(m))
    objective = Objective.gaussian_distribution(numpy.array(mu), numpy.array(sd), nResponsesPerGroup)
    prior = [scipy.stats.norm(loc=0, scale=1) for name in parameterName]
    return objective, prior, trueValueString


def main(pooling):
    nChains, nIter, nSamples = 2, 1000, 100
    outputDirectory = "./example/sample/distribution/"
    parameterName = ("a", "b", "c")
    nGroups, nResponsesPerGroup = 10, 10
    objective, prior, trueValueString = getFunction(parameterName, nGroups, nResponsesPerGroup)
    samplePosterior(nChains, nIter, nSamples, parameterName, nGroups, nResponsesPerGroup,
                    pooling, objective, outputDirectory, priorDistribution=prior, nProcesses=1)
    print(trueValueString)
    diagnoseSamples(outputDirectory)


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Example MCMC to sample from Gaussian distribution.")
    parser.add_argument("pooling", nargs="?", default="partial",
                        help="Pooling method (optional) : partial, complete or none. Default is partial.")
    main(parser.parse_args().pooling)
