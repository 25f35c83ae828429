"""GPU tests of the drop-in modules (posteriorSampling.samplePosterior,
sampleDiagnosis.diagnoseSamples / Diagnostic / Summary / computeHpdInterval) through the
product package, checked against the reference's golden outputs and the CPU oracle."""

import json
import os
import shutil

import numpy
import pytest
import scipy.stats

from conftest import GOLDEN, goldenPath, loadGolden, oracleObjectiveFromMeta
import parity
from oracle import diagnosis_oracle as do
from oracle import posterior_oracle as po

pytestmark = pytest.mark.gpu

DIAG_CASES = ["reg_partial", "reg_none", "reg_complete", "dist_none", "dist_complete",
              "c1_distribution_partial", "reg_ragged_partial"]


def _stage(case, tmp_path):
    """Copy a golden case's reference-written sample files into an output directory."""
    out = tmp_path / "out"
    (out / "sample").mkdir(parents=True)
    for f in os.listdir(os.path.join(GOLDEN, case)):
        if f.startswith("sample.") and f.endswith(".csv"):
            shutil.copy(os.path.join(GOLDEN, case, f), out / "sample" / f)
    return str(out)


@pytest.mark.parametrize("case", DIAG_CASES)
def test_diagnostic_matches_reference_within_1e10(case, tmp_path):
    """north star: diagnoseSamples within 1e-10 on an identical sample array."""
    import sampleDiagnosis as sd
    out = _stage(case, tmp_path)
    with open(goldenPath(case, "diag.json")) as h:
        ref = json.load(h)
    d = sd.Diagnostic(out + "/sample/")
    assert d._m == ref["m"] and d._n == ref["n"]
    assert d.partiallyPooled == ref["partiallyPooled"] and d.completelyPooled == ref["completelyPooled"]
    for k in ref["rhat"]:
        numpy.testing.assert_allclose(d.rhat[k], float(ref["rhat"][k]), rtol=1e-10)
        numpy.testing.assert_allclose(d.effectiveN[k], float(ref["effectiveN"][k]), rtol=1e-10)
        numpy.testing.assert_allclose(d.median[k], float(ref["median"][k]), rtol=1e-10, atol=1e-300)
        numpy.testing.assert_allclose(d.hdi[k][0], float(ref["hdi"][k][0]), rtol=1e-10, atol=1e-300)
        numpy.testing.assert_allclose(d.hdi[k][1], float(ref["hdi"][k][1]), rtol=1e-10, atol=1e-300)


@pytest.mark.parametrize("case", DIAG_CASES)
def test_diagnoseSamples_writes_the_reference_files(case, tmp_path, capsys):
    import sampleDiagnosis as sd
    out = _stage(case, tmp_path)
    sd.diagnoseSamples(out, nFigures=0)
    printed = capsys.readouterr().out
    for name in ("diagnosticAssessment.csv", "diagnosticAssessmentHyperOnly.csv",
                 "diagnosticAssessmentIndividual.csv"):
        ref = goldenPath(case, name)
        got = os.path.join(out, "diagnostic", name)
        assert os.path.exists(ref) == os.path.exists(got), name
        if os.path.exists(ref):
            assert open(ref).read() == open(got).read(), name
    assert open(goldenPath(case, "summary.csv")).read() == open(os.path.join(out, "sample", "summary.csv")).read()
    assert printed == open(goldenPath(case, "diagnose.stdout.txt")).read()


def test_diagnoseSamples_writes_the_figures(tmp_path, capsys):
    """nFigures > 0 (:73-85): the tables as before, then figure/logLikelihood.png and one trace plot and one
    bivariate plot per key suffix, hyper-parameters first."""
    import figures
    import sampleDiagnosis as sd
    case = "c1_distribution_partial"
    out = _stage(case, tmp_path)
    shutil.copy(os.path.join(GOLDEN, case, "logLikelihood.0.csv"), os.path.join(out, "sample"))
    sd.diagnoseSamples(out, nFigures=2)
    printed = capsys.readouterr().out
    tables = open(goldenPath(case, "diagnose.stdout.txt")).read()
    assert printed.startswith(tables)
    assert "Creating loglikelihood plot: Done" in printed[len(tables):] and printed.rstrip().endswith("Creating bivariate plots: Done.")
    assert sorted(os.listdir(out + "/figure/traceplot")) == ["traceplot[000].png", "traceplot_.png"]
    assert sorted(os.listdir(out + "/figure/bivariate")) == ["bivariate[000].png", "bivariate_.png"]
    assert figures.readPng(out + "/figure/logLikelihood.png").shape == (300, 1200, 3)
    img = figures.readPng(out + "/figure/traceplot/traceplot_.png")
    assert img.shape == (1200, 1200, 3) and (img != 255).any()


def test_computeHpdInterval_matches_reference_formula():
    import sampleDiagnosis as sd
    rs = numpy.random.RandomState(3)
    for n in (2, 3, 10, 101, 4000):
        x = rs.normal(size=n)
        x[: n // 3] = numpy.round(x[: n // 3], 1)          # ties
        got = sd.computeHpdInterval(x, 95)
        ref = do.computeHpdInterval(x, 95)
        assert got[0] == ref[0] and got[1] == ref[1]


def _regCase():
    meta = loadGolden("reg_partial")
    obj, prior = oracleObjectiveFromMeta(meta)
    return meta, obj, prior


@pytest.mark.parametrize("pooling", ["partial", "none", "complete"])
def test_start_state_equals_reference_chain(pooling):
    """Engine.initialise draws from each chain's MT19937 stream in the reference's order."""
    from engine import Engine
    meta, obj, prior = _regCase()
    names = tuple(meta["parameterName"])
    nC = 5
    eng = Engine(parity.deviceObjective(obj, 10, "fp64"), 10, 10, pooling, nC, priorDistribution=prior, chainId0=2)
    eng.initialise(names, meta["startingPointValueRange"], False)
    st = eng.getState()
    for c in range(nC):
        oc = po.OracleChain(2 + c, 2 + c, 100, 50, names, 10, 10, pooling, obj, prior, False,
                            meta["startingPointValueRange"])
        numpy.testing.assert_array_equal(st["theta"][:, :, c], oc.value)
        if pooling == "partial":
            numpy.testing.assert_array_equal(st["mu"][:, c], oc.mu)
            numpy.testing.assert_array_equal(st["sigma2"][:, c], oc.sigma2)
            numpy.testing.assert_allclose(st["ll"][:, c], oc.LL, rtol=1e-12)
        else:
            assert numpy.isnan(st["ll"][:, c]).all()
            numpy.testing.assert_allclose(st["lprior"][:, :, c], oc.logPrior, rtol=1e-14)


def test_start_state_with_redrawn_groups_keeps_the_stale_log_priors():
    """Partial pooling, a start range that makes many group-level noise sds negative: the groups whose
    log-likelihood is not finite are redrawn (:746-758) from the chain's stream, and their stored log-priors
    stay those of the FIRST draw (:284-288) -- theta, log-likelihoods and the stale log-priors must
    equal the reference chain's."""
    from engine import Engine
    meta, obj, prior = _regCase()
    names = tuple(meta["parameterName"])
    ranges = dict(meta["startingPointValueRange"], sigma=[0.01, 0.2])
    nC = 6
    eng = Engine(parity.deviceObjective(obj, 10, "fp64"), 10, 10, "partial", nC, chainId0=0)
    eng.initialise(names, ranges, False)
    assert eng.lpriorStale
    st = eng.getState()
    redrawn = 0
    for c in range(nC):
        oc = po.OracleChain(c, c, 100, 50, names, 10, 10, "partial", obj, prior, False, ranges)
        numpy.testing.assert_array_equal(st["theta"][:, :, c], oc.value)
        numpy.testing.assert_allclose(st["ll"][:, c], oc.LL, rtol=1e-12)
        numpy.testing.assert_allclose(st["lprior"][:, :, c], oc.logPrior, rtol=1e-13, atol=1e-13)
        with numpy.errstate(all="ignore"):
            fresh = scipy.stats.norm(oc.mu[:, None], numpy.sqrt(oc.sigma2)[:, None]).logpdf(oc.value)
        redrawn += int((numpy.abs(fresh - oc.logPrior) > 1e-9).sum())
    assert redrawn > 0            # some stored log-priors really are stale


def test_mle_start_close_to_scipy_nelder_mead():
    from engine import Engine
    meta, obj, prior = _regCase()
    names = tuple(meta["parameterName"])
    eng = Engine(parity.deviceObjective(obj, 10, "fp64"), 10, 10, "partial", 3, chainId0=0)
    eng.initialise(names, meta["startingPointValueRange"], True)
    for c in range(3):
        oc = po.OracleChain(c, c, 100, 50, names, 10, 10, "partial", obj, prior, True,
                            meta["startingPointValueRange"])
        numpy.testing.assert_allclose(eng.startingPoint[:, c], oc.startingPoint, rtol=1e-6, atol=1e-6)


def test_samplePosterior_free_running_matches_oracle_within_mc_error(tmp_path):
    """Free-running (Philox) posterior means and SDs agree with the oracle within MC error."""
    import posteriorSampling as ps
    import sampleDiagnosis as sd
    from objectives import Objective
    meta, obj, prior = _regCase()
    names = tuple(meta["parameterName"])
    nIter, nSamples, nChains = 4000, 1000, 16
    out = str(tmp_path / "gpu")
    ps.samplePosterior(nChains, nIter, nSamples, names, 10, 10, "partial",
                       Objective.linear_regression(obj.X, obj.y), out, saveLogLikelihood=False,
                       priorDistribution=prior, startingPointValueRange=meta["startingPointValueRange"],
                       displayProgress=False)
    keys, gpu, chains = sd.loadSamples(out + "/sample/")
    assert gpu.shape == (nChains, 1000, 36) and chains == list(range(nChains))
    text = open(out + "/sample/sample.3.csv").read().splitlines()
    assert text[0].startswith("index,chain,b0_mu,b0_sigma2,b0[000],") and text[1].startswith("2000,3,")
    assert len(text) == 1001 and os.path.exists(out + "/log/samplePosterior.log")
    # oracle: same model, its own RNG
    rows = []
    for c in range(6):
        oc = po.OracleChain(100 + c, 100 + c, nIter, nSamples, names, 10, 10, "partial", obj, prior, False,
                            meta["startingPointValueRange"])
        rows.append(numpy.stack([r[1] for r in oc.run()]))
    ora = numpy.stack(rows)                                    # [chains][rows][keys]
    gm, om = gpu.mean(axis=(0, 1)), ora.mean(axis=(0, 1))
    gs, os_ = gpu.std(axis=(0, 1)), ora.std(axis=(0, 1))
    # Monte-Carlo standard error from between-chain spread of the chain means (both arms)
    se = numpy.sqrt(gpu.mean(axis=1).var(axis=0, ddof=1) / nChains + ora.mean(axis=1).var(axis=0, ddof=1) / 6)
    z = numpy.abs(gm - om) / numpy.maximum(se, 1e-12)
    assert numpy.median(z) < 1.5 and z.max() < 6.0, (z.max(), keys[int(z.argmax())])
    numpy.testing.assert_allclose(gs, os_, rtol=0.35)


def test_samplePosterior_loglikelihood_file_and_binary_store(tmp_path, monkeypatch):
    import posteriorSampling as ps
    import sampleDiagnosis as sd
    from objectives import Objective
    meta, obj, prior = _regCase()
    names = tuple(meta["parameterName"])
    out = str(tmp_path / "csv")
    ps.samplePosterior(3, 200, 50, names, 10, 10, "none", Objective.linear_regression(obj.X, obj.y, "fp64"),
                       out, saveLogLikelihood=True, priorDistribution=prior,
                       startingPointValueRange=meta["startingPointValueRange"], displayProgress=False)
    keys, smp, _ = sd.loadSamples(out + "/sample/")
    ll = numpy.loadtxt(out + "/sample/logLikelihood.1.csv", delimiter=",")
    assert ll.shape == (50, 100)
    # pointwise log-likelihood of the retained state, recomputed by the oracle objective
    theta = smp[1, -1].reshape(3, 10)
    ref = obj([numpy.repeat(theta[p], 10) for p in range(3)])
    numpy.testing.assert_allclose(ll[-1], ref, atol=2e-6)
    # the same run forced into the binary store
    monkeypatch.setattr(ps, "CSV_VALUE_LIMIT", 0)
    out2 = str(tmp_path / "bin")
    ps.samplePosterior(3, 200, 50, names, 10, 10, "none", Objective.linear_regression(obj.X, obj.y, "fp64"),
                       out2, saveLogLikelihood=False, priorDistribution=prior,
                       startingPointValueRange=meta["startingPointValueRange"], displayProgress=False)
    assert os.path.exists(out2 + "/sample/manifest.json")
    keys2, smp2, _ = sd.loadSamples(out2 + "/sample/")
    assert keys2 == keys
    numpy.testing.assert_allclose(smp2, smp, rtol=1e-6, atol=1e-6)     # float32 store, same Philox streams
    d = sd.Diagnostic(out2 + "/sample/")
    assert set(d.rhat) == set(keys)


def test_tune_interval_beyond_the_16_bit_counters_is_refused():
    from engine import Engine
    meta, obj, prior = _regCase()
    eng = Engine(parity.deviceObjective(obj, 10, "fp32"), 10, 10, "partial", 2)
    eng.initialise(tuple(meta["parameterName"]), meta["startingPointValueRange"])
    with pytest.raises(RuntimeError, match="16 bits"):
        eng.run(0, 5, 3, 1, tuneInterval=70000)
    eng.run(0, 5, 3, 1, tuneInterval=65535)


def test_samplePosterior_argument_errors(tmp_path):
    import posteriorSampling as ps
    from objectives import Objective
    meta, obj, prior = _regCase()
    names = tuple(meta["parameterName"])
    handle = Objective.linear_regression(obj.X, obj.y)
    with pytest.raises(TypeError):
        ps.samplePosterior(1, 10, 5, names, 10, 10, "partial", lambda p: p, str(tmp_path / "a"), displayProgress=False)
    with pytest.raises(Exception):
        ps.samplePosterior(1, 10, 5, names, 10, 10, "hierarchical", handle, str(tmp_path / "b"), displayProgress=False)
    with pytest.raises(Exception):
        ps.samplePosterior(1, 5, 10, names, 10, 10, "partial", handle, str(tmp_path / "c"), displayProgress=False)
    with pytest.raises(ValueError, match="Invalid prior"):
        ps.samplePosterior(1, 10, 5, names, 10, 10, "none", handle, str(tmp_path / "d"),
                           startingPointValueRange=meta["startingPointValueRange"], displayProgress=False)


POISSON_SRC = """
// Poisson regression: y ~ Poisson(exp(a + b*x)); record = (x, y, lgamma(y+1))
__device__ mcmc_real mcmc_obj_loglik(const mcmc_real* theta, const mcmc_real* obs, const mcmc_real* hdr,
                                     int obs_index, int group) {
    const mcmc_real eta = theta[0] + theta[1] * obs[0] + hdr[0];
    return obs[1] * eta - exp(eta) - obs[2];
}
"""


class _PoissonOracle(object):
    """numpy twin of POISSON_SRC in the reference's objective style."""

    def __init__(self, x, y, offset, nResp):
        import scipy.special
        self.x, self.y = x, y
        self.lg = scipy.special.gammaln(y + 1)
        self.off = numpy.repeat(offset, nResp)

    def __call__(self, parameter):
        eta = numpy.asarray(parameter[0]) + numpy.asarray(parameter[1]) * self.x + self.off
        with numpy.errstate(all="ignore"):
            return self.y * eta - numpy.exp(eta) - self.lg


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-11), ("fp32", 1e-5)])
def test_user_objective_compiled_by_nvrtc_replays_like_the_oracle(precision, tol):
    """north star (1): a user objective given as CUDA source, with a per-group header."""
    import torch
    from engine import Engine, SampleStore
    from objectives import Objective
    rs = numpy.random.RandomState(5)
    G, R = 7, 12
    nResp = [R] * G
    x = rs.normal(size=G * R)
    offset = rs.normal(0, 0.2, size=G)
    y = rs.poisson(numpy.exp(0.3 + 0.5 * x + numpy.repeat(offset, R))).astype(float)
    ora = _PoissonOracle(x, y, offset, nResp)
    import scipy.special
    handle = Objective.from_source(POISSON_SRC, 2, numpy.stack([x, y, scipy.special.gammaln(y + 1)], axis=1),
                                   header=offset[:, None], precision=precision)
    names, ranges = ("a", "b"), {"a": [-1, 1], "b": [-1, 1]}
    nC, nIter = 3, 80
    chains, start = [], []
    for c in range(nC):
        oc = po.OracleChain(c, c, nIter, 40, names, G, nResp, "partial", ora, None, False, ranges, recordTape=True)
        start.append(dict(value=oc.value.copy(), logPrior=oc.logPrior.copy(), LL=oc.LL.copy(),
                          mu=oc.mu.copy(), sigma2=oc.sigma2.copy()))
        oc.rows = oc.run()
        chains.append(oc)
    eng = Engine(handle, G, nResp, "partial", nC)
    stack = lambda k: numpy.stack([s[k] for s in start], axis=-1)
    eng.setState(stack("value"), stack("LL"), stack("logPrior"), stack("mu"), stack("sigma2"))

    def tens(field, shape, dtype=torch.float64):
        t = torch.zeros(shape + (eng.S,), dtype=dtype, device=eng.device)
        t[..., :nC] = torch.from_numpy(numpy.stack([getattr(c.tape, field) for c in chains], axis=-1)).to(eng.device).to(dtype)
        return t

    tape = {"z": tens("z_prop", (nIter, 2, G)), "u": torch.nan_to_num(tens("u_acc", (nIter, 2, G)), nan=0.5),
            "accept": tens("accept", (nIter, 2, G), torch.uint8),
            "zmu": tens("z_mu", (nIter, 2)), "qsig": tens("q_sig", (nIter, 2))}
    tr = eng.run(0, nIter, chains[0].burn, chains[0].thin, tape=tape, trace=True, useLpriorOverride=True)
    torch.cuda.synchronize()
    got = tr["ll"][..., :nC].cpu().numpy()
    want = numpy.stack([c.tape.ll_prop for c in chains], axis=-1)
    assert parity.relErr(got, want).max() <= tol
    assert (tr["accept"][..., :nC].cpu().numpy() != numpy.stack([c.tape.accept for c in chains], axis=-1)).sum() <= 2
    # free-running production kernels on the same compiled objective
    eng.run(nIter, 50, 0, 1)
    torch.cuda.synchronize()
    assert numpy.isfinite(eng.getState()["theta"]).all()


def test_user_objective_compile_error_is_reported():
    from objectives import Objective
    with pytest.raises(RuntimeError, match="failed to compile"):
        Objective.from_source("__device__ mcmc_real mcmc_obj_loglik(const mcmc_real* t) { return nope; }",
                              2, numpy.zeros((4, 1)))


def test_convergence_from_device_store_matches_diagnostic():
    """convergenceFromStore (R-hat / ESS straight from the device-resident sample store, used by
    bench.py's min-ESS figure) against Diagnostic on the same draws (north star: 1e-10)."""
    import torch
    import sampleDiagnosis as sd
    rs = numpy.random.RandomState(3)
    nChains, rows, ncol, S = 5, 40, 7, 32
    draws = numpy.cumsum(rs.normal(size=(nChains, rows, ncol)), axis=1) * 0.1 + rs.normal(size=(nChains, 1, ncol))
    store = torch.zeros((rows + 3, ncol, S), dtype=torch.float64, device="cuda")
    store[:rows, :, :nChains] = torch.from_numpy(numpy.transpose(draws, (1, 2, 0))).cuda()
    rhat, ess = sd.convergenceFromStore(store, rows, nChains)
    keys = ["k[%03d]" % i for i in range(ncol)]
    d = sd.Diagnostic(samples=draws, keys=keys)
    numpy.testing.assert_allclose(rhat.cpu().numpy(), [d.rhat[k] for k in keys], rtol=1e-10)
    numpy.testing.assert_allclose(ess.cpu().numpy(), [d.effectiveN[k] for k in keys], rtol=1e-10)


def test_one_pass_hyper_update_equals_two_pass_at_c3_shape(monkeypatch):
    """The Gibbs hyper update reads theta once when there are many groups (shifted sums); it must
    agree with the two-pass kernel (the reference's formulas, :485-:495) on the same state and tape."""
    import torch
    from engine import Engine
    obj, names, nResp, ranges = parity.syntheticRegression(G=1024, R=8, K=2)
    out = {}
    for label in ("one", "two"):
        if label == "two":
            monkeypatch.setenv("MCMCN_HYPER_TWO_PASS", "1")
        eng = Engine(parity.deviceObjective(obj, nResp, "fp32"), 1024, nResp, "partial", 64, seed=3)
        eng.initialise(names, ranges)
        gen = torch.Generator(device="cpu").manual_seed(5)
        P, G = 3, 1024
        tape = {"z": torch.randn((1, P, G, eng.S), generator=gen, dtype=torch.float64).to(eng.device),
                "u": torch.rand((1, P, G, eng.S), generator=gen, dtype=torch.float64).to(eng.device),
                "zmu": torch.randn((1, P, eng.S), generator=gen, dtype=torch.float64).to(eng.device),
                "qsig": torch.rand((1, P, eng.S), generator=gen, dtype=torch.float64).add_(0.5).to(eng.device)}
        eng.run(0, 1, 0, 1, tape=tape)
        torch.cuda.synchronize()
        st = eng.getState()
        out[label] = (st["mu"], st["sigma2"])
    numpy.testing.assert_allclose(out["one"][0], out["two"][0], rtol=1e-13, atol=1e-13)
    numpy.testing.assert_allclose(out["one"][1], out["two"][1], rtol=1e-12)
