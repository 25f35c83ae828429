import json
import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mcmc-for-nested-data_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Like the reference, the product modules (posteriorSampling, sampleDiagnosis)
# are top-level modules of the package directory.
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def loadGolden(case):
    with open(os.path.join(GOLDEN, case, "meta.json")) as h:
        meta = json.load(h)
    return meta


def goldenPath(case, name):
    return os.path.join(GOLDEN, case, name)


def oracleObjectiveFromMeta(meta):
    """Rebuild the numpy objective + scipy priors of a golden case (oracle side)."""
    import scipy.stats
    from oracle import posterior_oracle as po
    d = meta["data"]
    nResp = meta["nResponsesPerGroup"]
    if d["objective"] == "gaussian_distribution":
        obj = po.GaussianDistributionObjective(d["mu"], d["sd"], nResp)
    elif d["objective"] == "linear_regression":
        X = numpy.array([[float(v) for v in row] for row in d["X"]])
        y = numpy.array([float(v) for v in d["y"]])
        obj = po.LinearRegressionObjective(X, y)
    else:
        raise KeyError(d["objective"])
    prior = None
    if meta["prior"] is not None:
        prior = [getattr(scipy.stats, s["dist"])(*s["args"], **s["kwds"]) for s in meta["prior"]]
    return obj, prior


@pytest.fixture
def golden():
    return loadGolden
