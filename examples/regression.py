#!/usr/bin/env python
"""The reference's linear-regression example (example/regression.py) on the B200 engine.

    python examples/regression.py [partial|none|complete]

Same data generator (numpy.random.seed(12345)), priors, ranges, MLE start and chain counts as the
reference; the objective is the registry's linear_regression device function."""

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


def generateData(nGroups, nResponsesPerGroup):
    """example/regression.py:16-50."""
    n = nGroups * nResponsesPerGroup
    x = numpy.hstack([numpy.tile([1], (n))[numpy.newaxis].T, numpy.random.normal(size=(n, 1))])
    beta = numpy.hstack([
        numpy.repeat(numpy.random.normal(loc=0, scale=1, size=nGroups),
                     [nResponsesPerGroup] * nGroups)[numpy.newaxis].T,
        numpy.repeat(numpy.random.normal(loc=100, scale=100, size=nGroups),
                     [nResponsesPerGroup] * nGroups)[numpy.newaxis].T])
    y = numpy.sum(x * beta, axis=1) + numpy.random.normal(size=n)
    trueValueString = "\nTrue value:\n"
    for i in range(beta.shape[1]):
        trueValueString += "\tbeta%i: {mean: %.2f, sd: %.2f}\n" % (i, numpy.mean(beta[:, i]), numpy.std(beta[:, 1]))
    return {"X": x, "y": y}, trueValueString


def main(pooling):
    nChains, nIter, nSamples = 4, 2000, 1000
    outputDirectory = "./example/sample/regression/"
    parameterName = ("b0", "b1", "sigma")
    startingPointValueRange = {"b0": [-100, 100], "b1": [0, 200], "sigma": [0.00, 100.]}
    prior = [scipy.stats.norm(loc=0, scale=10), scipy.stats.norm(loc=100, scale=10), scipy.stats.gamma(10)]
    nGroups, nResponsesPerGroup = 10, 10
    data, trueValueString = generateData(nGroups, nResponsesPerGroup)
    objective = Objective.linear_regression(data["X"], data["y"])
    samplePosterior(nChains, nIter, nSamples, parameterName, nGroups, nResponsesPerGroup,
                    pooling, objective, outputDirectory, priorDistribution=prior,
                    startWithMLE=True, startingPointValueRange=startingPointValueRange, nProcesses=0)
    print(trueValueString)
    diagnoseSamples(outputDirectory)


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Example MCMC for a linear regression.")
    parser.add_argument("pooling", nargs="?", default="partial",
                        help="Pooling method (optional) : partial, complete or none. Default is partial.")
    main(parser.parse_args().pooling)
