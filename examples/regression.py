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
    """Mock data of the reference's example (example/regression.py:16-50): a constant and one
    standard-normal predictor, group intercepts ~ N(0, 1), group slopes ~ N(100, 100), unit noise.
    The global numpy stream (seeded above) is consumed in the reference's order -- predictor,
    intercepts, slopes, noise -- so the data are the same numbers."""
    nObservations = nGroups * nResponsesPerGroup
    predictor = numpy.random.normal(size=nObservations)
    intercept = numpy.random.normal(0.0, 1.0, nGroups)
    slope = numpy.random.normal(100.0, 100.0, nGroups)
    noise = numpy.random.normal(size=nObservations)
    member = numpy.arange(nObservations) // nResponsesPerGroup      # group of each observation
    X = numpy.column_stack([numpy.ones(nObservations), predictor])
    y = intercept[member] + slope[member] * predictor + noise
    # the reference reports the slope's sd for both coefficients (example/regression.py:45); kept, it is only a printout
    sdShown = numpy.std(slope[member])
    lines = ["\tbeta%i: {mean: %.2f, sd: %.2f}\n" % (i, numpy.mean(b[member]), sdShown)
             for i, b in enumerate((intercept, slope))]
    return {"X": X, "y": y}, "\nTrue value:\n" + "".join(lines)


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
