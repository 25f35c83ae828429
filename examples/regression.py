#!/usr/bin/env python
"""BASELINE config 2 on the B200 engine: the reference's linear-regression example
(example/regression.py: 10 groups x 10 responses, parameters b0, b1, sigma, 4 chains x 2,000 iterations,
MLE start).

    python examples/regression.py [partial|none|complete]

Mock data as in the reference: a constant and one standard-normal predictor, group intercepts ~ N(0, 1),
group slopes ~ N(100, 100), unit noise, drawn off numpy's global stream (seed 12345) in the reference's
order -- predictor, intercepts, slopes, noise -- so the data are the same numbers.  The objective is the
registry's linear_regression device function."""

import numpy
import scipy.stats

from common import Workload, poolingFromCommandLine
from objectives import Objective

numpy.random.seed(12345)

CONFIG = Workload("Example MCMC for a linear regression.", "./example/sample/regression/",
                  names=("b0", "b1", "sigma"), groups=10, responses=10, chains=4, iterations=2000, retained=1000,
                  priorDistribution=[scipy.stats.norm(loc=0, scale=10), scipy.stats.norm(loc=100, scale=10),
                                     scipy.stats.gamma(10)],
                  startWithMLE=True,
                  startingPointValueRange={"b0": [-100, 100], "b1": [0, 200], "sigma": [0.00, 100.]},
                  nProcesses=0)


def generateData(nGroups, nResponsesPerGroup):
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


if __name__ == "__main__":
    pooling = poolingFromCommandLine(CONFIG.title)
    data, truth = generateData(CONFIG.groups, CONFIG.responses)
    CONFIG.run(pooling, Objective.linear_regression(data["X"], data["y"]), truth)
