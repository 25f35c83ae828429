#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED
reference from /root/reference (read-only; imported in place, nothing copied).

Run once in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [case ...]

Shims: ``matplotlib`` is stubbed because it is not
installed and figures are out of scope, and ``diagnoseSamples`` is always
called with ``nFigures=0``.  The reference's source is not modified.

Each case directory holds ``meta.json`` (the inputs, incl. the synthetic
data the example drew, and sha256 of every sample file), the sample CSVs for
the small cases, and ``diag.json`` + the diagnostic CSV texts produced by the
reference's ``Diagnostic`` / ``Summary``.
"""

import contextlib
import hashlib
import io
import json
import os
import runpy
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

# --- shims -----------------------------------------------------------------
_mpl = types.ModuleType("matplotlib")
_plt = types.ModuleType("matplotlib.pyplot")
_plt.close = lambda *a, **k: None
_mpl.pyplot = _plt
sys.modules.setdefault("matplotlib", _mpl)
sys.modules.setdefault("matplotlib.pyplot", _plt)
sys.path.insert(0, REF)

import numpy  # noqa: E402
import scipy.stats  # noqa: E402
import posteriorSampling as refPS  # noqa: E402
import sampleDiagnosis as refSD  # noqa: E402


def sha256(path):
    with open(path, "rb") as h:
        return hashlib.sha256(h.read()).hexdigest()


def priorSpec(prior):
    if prior is None:
        return None
    out = []
    for d in prior:
        out.append({"dist": d.dist.name, "args": list(map(float, d.args)),
                    "kwds": dict((k, float(v)) for k, v in d.kwds.items())})
    return out


def diagnose(outDir, caseDir):
    """Run the reference's Diagnostic + Summary; store results at full precision."""
    sampleDir = outDir + "/sample/"
    d = refSD.Diagnostic(sampleDir)
    diag = {"m": d._m, "n": d._n,
            "partiallyPooled": bool(d.partiallyPooled),
            "completelyPooled": bool(d.completelyPooled),
            "fileOrder": [os.path.basename(f) for f in
                          __import__("glob").glob(sampleDir + "/sample*.csv")],
            "rhat": dict((k, repr(float(v))) for k, v in d.rhat.items()),
            "effectiveN": dict((k, repr(float(v))) for k, v in d.effectiveN.items()),
            "median": dict((k, repr(float(v))) for k, v in d.median.items()),
            "hdi": dict((k, [repr(float(v[0])), repr(float(v[1]))]) for k, v in d.hdi.items())}
    with open(os.path.join(caseDir, "diag.json"), "w") as h:
        json.dump(diag, h, indent=1, sort_keys=True)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        refSD.diagnoseSamples(outDir, nFigures=0)
    with open(os.path.join(caseDir, "diagnose.stdout.txt"), "w") as h:
        h.write(buf.getvalue())
    for name in ("diagnosticAssessment.csv", "diagnosticAssessmentHyperOnly.csv",
                 "diagnosticAssessmentIndividual.csv"):
        src = os.path.join(outDir, "diagnostic", name)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(caseDir, name))
    shutil.copy(os.path.join(outDir, "sample", "summary.csv"),
                os.path.join(caseDir, "summary.csv"))


def runCase(case, objective, dataMeta, nChains, nIter, nSamples, names, nGroups,
            nResp, pooling, prior, startWithMLE, valueRange, keepCsv,
            saveLogLikelihood=False, keepLL=False, doDiag=True):
    caseDir = os.path.join(HERE, case)
    shutil.rmtree(caseDir, ignore_errors=True)
    os.makedirs(caseDir)
    tmp = tempfile.mkdtemp(prefix="golden_")
    out = os.path.join(tmp, "out")
    with contextlib.redirect_stdout(io.StringIO()):
        refPS.samplePosterior(nChains, nIter, nSamples, names, nGroups, nResp,
                              pooling, objective, out,
                              saveLogLikelihood=saveLogLikelihood,
                              priorDistribution=prior, startWithMLE=startWithMLE,
                              startingPointValueRange=valueRange, nProcesses=1,
                              displayProgress=False)
    meta = {"case": case, "nChains": nChains, "nIter": nIter, "nSamples": nSamples,
            "parameterName": list(names), "nGroups": nGroups,
            "nResponsesPerGroup": nResp, "pooling": pooling,
            "prior": priorSpec(prior), "startWithMLE": startWithMLE,
            "startingPointValueRange": valueRange,
            "saveLogLikelihood": saveLogLikelihood, "data": dataMeta,
            "versions": {"numpy": numpy.__version__,
                         "scipy": __import__("scipy").__version__,
                         "pandas": __import__("pandas").__version__},
            "sha256": {}}
    for c in range(nChains):
        f = os.path.join(out, "sample", "sample.%i.csv" % c)
        meta["sha256"]["sample.%i.csv" % c] = sha256(f)
        if keepCsv:
            shutil.copy(f, caseDir)
        if saveLogLikelihood:
            f = os.path.join(out, "sample", "logLikelihood.%i.csv" % c)
            meta["sha256"]["logLikelihood.%i.csv" % c] = sha256(f)
            if keepLL and c == 0:
                shutil.copy(f, caseDir)
    with open(os.path.join(caseDir, "meta.json"), "w") as h:
        json.dump(meta, h, indent=1, sort_keys=True)
    if doDiag:
        diagnose(out, caseDir)
    shutil.rmtree(tmp, ignore_errors=True)
    print("made", case)


# --- the reference's own example workloads ---------------------------------
def distributionData():
    """example/distribution.py:13,26-39 with its module-level seed(12345)."""
    ns = runpy.run_path(os.path.join(REF, "example", "distribution.py"), run_name="golden")
    names = ("a", "b", "c")
    # draw the same data the example draws, and keep it for the fixture
    state = numpy.random.get_state()
    mu, sd = [], []
    for _ in names:
        mu.append(numpy.random.normal(loc=0, scale=1, size=10).tolist())
        sd.append(float(numpy.random.gamma(1)))
    numpy.random.set_state(state)
    func, prior, _ = ns["getFunction"](names, 10, 10)
    return func, prior, {"objective": "gaussian_distribution", "mu": mu, "sd": sd}


def regressionData(nGroups=10, nResp=10):
    """example/regression.py:13,16-50 with its module-level seed(12345)."""
    ns = runpy.run_path(os.path.join(REF, "example", "regression.py"), run_name="golden")
    import functools
    n = nResp if isinstance(nResp, int) else None
    if n is not None:
        data, _ = ns["generateData"](nGroups, nResp)
    else:
        # ragged variant: same generator, then drop rows to the wanted sizes
        full, _ = ns["generateData"](nGroups, max(nResp))
        keep = numpy.concatenate([numpy.arange(g * max(nResp), g * max(nResp) + r)
                                  for g, r in enumerate(nResp)])
        data = {"group": full["group"][keep], "X": full["X"][keep], "y": full["y"][keep]}
    objective = functools.partial(ns["computeLogLikelihood"], data=data)
    meta = {"objective": "linear_regression",
            "X": [[repr(float(v)) for v in row] for row in data["X"]],
            "y": [repr(float(v)) for v in data["y"]]}
    return objective, meta


REG_NAMES = ("b0", "b1", "sigma")
REG_RANGE = {"b0": [-100, 100], "b1": [0, 200], "sigma": [0.00, 100.]}


def regPrior():
    return [scipy.stats.norm(loc=0, scale=10), scipy.stats.norm(loc=100, scale=10),
            scipy.stats.gamma(10)]


def case_c1_distribution_partial():
    f, prior, meta = distributionData()
    runCase("c1_distribution_partial", f, meta, 2, 1000, 100, ("a", "b", "c"), 10, 10,
            "partial", prior, False, None, keepCsv=True, saveLogLikelihood=True, keepLL=True)


def case_dist_none():
    f, prior, meta = distributionData()
    runCase("dist_none", f, meta, 2, 400, 100, ("a", "b", "c"), 10, 10,
            "none", prior, False, None, keepCsv=True)


def case_dist_complete():
    f, prior, meta = distributionData()
    runCase("dist_complete", f, meta, 2, 400, 100, ("a", "b", "c"), 10, 10,
            "complete", prior, False, None, keepCsv=True)


def case_reg(pooling):
    f, meta = regressionData()
    runCase("reg_" + pooling, f, meta, 2, 600, 100, REG_NAMES, 10, 10, pooling,
            regPrior(), True, REG_RANGE, keepCsv=True,
            saveLogLikelihood=(pooling == "none"), keepLL=(pooling == "none"))


def case_reg_ragged():
    nResp = [3, 7, 10, 5, 1, 8, 10, 2, 6, 9]
    f, meta = regressionData(10, nResp)
    runCase("reg_ragged_partial", f, meta, 1, 240, 120, REG_NAMES, 10, nResp, "partial",
            None, False, REG_RANGE, keepCsv=True)


def case_c2(pooling):
    """BASELINE.json config 2: example.regression exactly (4 chains x 2000 it)."""
    f, meta = regressionData()
    runCase("c2_regression_" + pooling, f, meta, 4, 2000, 1000, REG_NAMES, 10, 10, pooling,
            regPrior(), True, REG_RANGE, keepCsv=False)


CASES = {
    "c1_distribution_partial": case_c1_distribution_partial,
    "dist_none": case_dist_none,
    "dist_complete": case_dist_complete,
    "reg_partial": lambda: case_reg("partial"),
    "reg_none": lambda: case_reg("none"),
    "reg_complete": lambda: case_reg("complete"),
    "reg_ragged_partial": case_reg_ragged,
    "c2_regression_partial": lambda: case_c2("partial"),
}

if __name__ == "__main__":
    todo = sys.argv[1:] or list(CASES)
    for name in todo:
        CASES[name]()
