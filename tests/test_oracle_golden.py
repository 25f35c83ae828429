"""Pin the CPU oracle against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only.

Sampler: the oracle's sample CSV must equal the reference's byte for byte
(sha256 for every case, full text where the fixture keeps the file).
Diagnostics: R-hat / ESS / median / HDI within 1e-12 relative, CSV texts equal.
"""

import hashlib
import json
import os

import numpy
import pytest
import scipy.special
import scipy.stats

from conftest import GOLDEN, goldenPath, loadGolden, oracleObjectiveFromMeta
from oracle import diagnosis_oracle as do
from oracle import posterior_oracle as po

SMALL_CASES = ["dist_none", "dist_complete", "reg_partial", "reg_none",
               "reg_complete", "reg_ragged_partial"]
FULL_CASES = ["c1_distribution_partial", "c2_regression_partial"]


def _have(case):
    return os.path.exists(os.path.join(GOLDEN, case, "meta.json"))


def _runOracle(meta, tmp_path):
    obj, prior = oracleObjectiveFromMeta(meta)
    po.samplePosteriorOracle(
        meta["nChains"], meta["nIter"], meta["nSamples"], tuple(meta["parameterName"]),
        meta["nGroups"], meta["nResponsesPerGroup"], meta["pooling"], obj, str(tmp_path),
        saveLogLikelihood=meta["saveLogLikelihood"], priorDistribution=prior,
        startWithMLE=meta["startWithMLE"],
        startingPointValueRange=meta["startingPointValueRange"])
    return os.path.join(str(tmp_path), "sample")


def _sha(path):
    with open(path, "rb") as h:
        return hashlib.sha256(h.read()).hexdigest()


@pytest.mark.parametrize("case", SMALL_CASES + FULL_CASES)
def test_sampler_oracle_matches_reference_bytes(case, tmp_path):
    if not _have(case):
        pytest.skip("fixture %s not generated" % case)
    meta = loadGolden(case)
    sampleDir = _runOracle(meta, tmp_path)
    for name, digest in meta["sha256"].items():
        got = os.path.join(sampleDir, name)
        ref = goldenPath(case, name)
        if os.path.exists(ref):
            with open(ref) as a, open(got) as b:
                assert a.read() == b.read(), name
        assert _sha(got) == digest, name

    # diagnostics oracle on the (byte-identical) sample files vs the reference's Diagnostic
    with open(goldenPath(case, "diag.json")) as h:
        ref = json.load(h)
    d = do.DiagnosticOracle(sampleDir)
    assert d._m == ref["m"] and d._n == ref["n"]
    assert d.partiallyPooled == ref["partiallyPooled"]
    assert d.completelyPooled == ref["completelyPooled"]
    d._compute()
    for k in ref["rhat"]:
        numpy.testing.assert_allclose(d.rhat[k], float(ref["rhat"][k]), rtol=1e-12)
        numpy.testing.assert_allclose(d.effectiveN[k], float(ref["effectiveN"][k]), rtol=1e-10)
        numpy.testing.assert_allclose(d.median[k], float(ref["median"][k]), rtol=0, atol=0)
        assert d.hdi[k][0] == float(ref["hdi"][k][0]) and d.hdi[k][1] == float(ref["hdi"][k][1])
    with open(goldenPath(case, "diagnosticAssessment.csv")) as h:
        assert d.assessmentString(False) == h.read()
    if ref["partiallyPooled"]:
        with open(goldenPath(case, "diagnosticAssessmentHyperOnly.csv")) as h:
            assert d.assessmentString(True) == h.read()
    if not ref["completelyPooled"]:
        with open(goldenPath(case, "diagnosticAssessmentIndividual.csv")) as h:
            assert d.summaryString() == h.read()
    with open(goldenPath(case, "summary.csv")) as h:
        assert do.summaryOracle(sampleDir) == h.read()


def test_c1_survey_known_answers():
    """C1 partial chain 0 start state and first retained row."""
    if not _have("c1_distribution_partial"):
        pytest.skip("fixture not generated")
    meta = loadGolden("c1_distribution_partial")
    assert meta["sha256"]["sample.0.csv"].startswith("44da431feeb436a0")
    assert meta["sha256"]["sample.1.csv"].startswith("ddd18b9a0c17de7a")
    obj, prior = oracleObjectiveFromMeta(meta)
    oc = po.OracleChain(0, 0, 1000, 100, ("a", "b", "c"), 10, 10, "partial", obj, prior)
    numpy.testing.assert_allclose(oc.startingPoint, [1.76405235, 0.40015721, 0.97873798], atol=5e-9)
    numpy.testing.assert_allclose(oc.sigma2, [0.42000623, 0.20003930, 0.31284788], atol=5e-9)


def test_norm_logpdf_restatement_is_scipy_bit_exact():
    rs = numpy.random.RandomState(1)
    x = rs.normal(0, 50, 2000)
    loc = rs.normal(0, 50, 2000)
    scale = numpy.abs(rs.normal(0, 3, 2000))
    scale[:5] = [0.0, -1.0, numpy.inf, 1e-300, numpy.nan]
    x[5:8] = [numpy.inf, -numpy.inf, numpy.nan]
    with numpy.errstate(all="ignore"):
        ref = scipy.stats.norm(loc=loc, scale=scale).logpdf(x)
    got = po.norm_logpdf(x, loc, scale)
    assert numpy.array_equal(ref, got, equal_nan=True)


def test_invgamma_draw_restatement():
    """HyperParameter._sampleInvChisq (posteriorSampling.py:497-498) vs the tape decomposition."""
    r1, r2 = numpy.random.RandomState(7), numpy.random.RandomState(7)
    for k in range(50):
        a, s2 = 4.5 + k, 0.3 + 0.1 * k
        v1 = scipy.stats.invgamma(a, scale=a * s2).rvs(random_state=r1)
        v2 = (1.0 / scipy.special.gammainccinv(a, r2.random_sample())) * (a * s2) + 0.0
        assert v1 == v2


def test_tape_replays_the_chain():
    """The tape alone (no RNG) must reproduce the oracle chain's decisions."""
    meta = loadGolden("reg_partial") if _have("reg_partial") else None
    if meta is None:
        pytest.skip("fixture not generated")
    obj, prior = oracleObjectiveFromMeta(meta)
    oc = po.OracleChain(1, 1, 200, 100, tuple(meta["parameterName"]), 10, 10, "partial",
                        obj, prior, False, meta["startingPointValueRange"], recordTape=True)
    value0 = oc.value.copy()
    oc.run()
    tp = oc.tape
    # proposals are value + scale*z and the accept bit is log(u) < diff when a uniform was drawn
    drew = ~numpy.isnan(tp.u_acc)
    with numpy.errstate(all="ignore"):
        assert numpy.array_equal(tp.accept[drew] == 1, numpy.log(tp.u_acc[drew]) < tp.diff[drew])
    assert drew.mean() > 0.8
    assert value0.shape == (3, 10)
