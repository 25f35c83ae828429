"""GPU parity tests proper (north star "Equivalence"): tape-driven replay of the CUDA step
kernels against the CPU oracle, through the C ABI.

  * per-step proposal log-likelihood within 1e-5 relative (FP32 objective math) /
    1e-11 (FP64), proposal log-prior within 1e-12;
  * accept/reject trajectory identical; in FP32 mode the oracle's decisions are
    teacher-forced and any decision the engine would have taken differently must be a
    documented near-threshold tie (|log u - diff| <= 1e-5 * max(|ll|, 1));
  * retained rows equal the oracle's (FP64: to 1e-9 and as "%f" text).
"""

import numpy
import pytest
import scipy.stats

from conftest import loadGolden, oracleObjectiveFromMeta
import parity

pytestmark = pytest.mark.gpu


def _goldenCase(case):
    meta = loadGolden(case)
    obj, prior = oracleObjectiveFromMeta(meta)
    return meta, obj, prior


def _fmt(rows):
    return [",".join("%f" % v for v in r) for r in rows]


@pytest.mark.parametrize("case", ["reg_partial", "reg_none", "reg_complete", "dist_none",
                                  "dist_complete", "c1_distribution_partial", "reg_ragged_partial"])
def test_replay_fp64_is_trajectory_exact(case):
    meta, obj, prior = _goldenCase(case)
    nIter, nSamples = 300, 100
    res = parity.replay(obj, tuple(meta["parameterName"]), meta["nGroups"], meta["nResponsesPerGroup"],
                        meta["pooling"], prior, meta["startingPointValueRange"], nChains=3,
                        nIter=nIter, nSamples=nSamples, precision="fp64", force=False)
    err, ties = parity.checkReplay(res, 1e-11, 0.0)
    assert ties == 0
    numpy.testing.assert_allclose(res.final["theta"], res.oracleFinal, rtol=1e-12, atol=1e-12)
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)
    for c in range(3):   # the sample file text the CSV writer would emit
        assert _fmt(res.rows[:, :, c]) == _fmt(res.oracleRows[:, :, c])


@pytest.mark.parametrize("case", ["reg_partial", "reg_none", "reg_complete", "dist_none",
                                  "c1_distribution_partial", "reg_ragged_partial"])
def test_replay_fp32_log_density_and_ties(case):
    meta, obj, prior = _goldenCase(case)
    res = parity.replay(obj, tuple(meta["parameterName"]), meta["nGroups"], meta["nResponsesPerGroup"],
                        meta["pooling"], prior, meta["startingPointValueRange"], nChains=4,
                        nIter=300, nSamples=100, precision="fp32")
    err, ties = parity.checkReplay(res, 1e-5, 1e-5)
    decisions = res.oracle["accept"].size
    assert ties <= max(3, decisions // 2000), "too many ties: %d of %d" % (ties, decisions)
    # teacher-forced, so the state follows the oracle exactly (proposals are FP64)
    numpy.testing.assert_allclose(res.final["theta"], res.oracleFinal, rtol=1e-12, atol=1e-12)
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)


def test_replay_mle_start_regression_example():
    """C2: example.regression with startWithMLE (oracle start state fed to the engine)."""
    meta, obj, prior = _goldenCase("reg_partial")
    res = parity.replay(obj, tuple(meta["parameterName"]), 10, 10, "partial", prior,
                        meta["startingPointValueRange"], nChains=2, nIter=400, nSamples=100,
                        precision="fp64", force=False, startWithMLE=True)
    err, ties = parity.checkReplay(res, 1e-11, 0.0)
    assert ties == 0
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_replay_c2_full_config_reproduces_the_reference_files(precision, tmp_path):
    """BASELINE config 2 as the reference runs it (example.regression partial: 4 chains x 2,000 iterations,
    nSamples 1,000, startWithMLE) in replay mode on the GPU: per-step log-densities and the accept
    trajectory against the oracle, and the sample files written from the engine's retained rows must
    be the reference's own files (sha256 of sample.<chain>.csv recorded by tests/golden/make_golden.py
    from the unmodified reference)."""
    import ast
    import hashlib
    import posteriorSampling as ps
    meta, obj, prior = _goldenCase("c2_regression_partial")
    names = tuple(meta["parameterName"])
    res = parity.replay(obj, names, meta["nGroups"], meta["nResponsesPerGroup"], meta["pooling"], prior,
                        meta["startingPointValueRange"], nChains=meta["nChains"], nIter=meta["nIter"],
                        nSamples=meta["nSamples"], precision=precision, startWithMLE=meta["startWithMLE"])
    if precision == "fp64":
        err, ties = parity.checkReplay(res, 1e-11, 0.0)
        assert ties == 0
    else:
        err, ties = parity.checkReplay(res, 1e-5, 1e-5)
        assert ties <= 30, ties                      # of 240,000 decisions; teacher-forced
    sha = meta["sha256"] if isinstance(meta["sha256"], dict) else ast.literal_eval(meta["sha256"])
    header = ps.sampleHeader(names, meta["nGroups"], meta["pooling"])
    assert len(res.store.iterations) == 1000
    for c in range(meta["nChains"]):
        path = str(tmp_path / ("sample.%d.csv" % c))
        ps.writeSampleCsv(path, c, header, res.store.iterations, res.rows[:, :, c])
        assert hashlib.sha256(open(path, "rb").read()).hexdigest() == sha["sample.%d.csv" % c], c


@pytest.mark.parametrize("precision,tol,tensorCore", [("fp32", 1e-5, True), ("fp32", 1e-5, False), ("fp64", 1e-11, False)])
def test_replay_c3_shape_wide_path(precision, tol, tensorCore, monkeypatch):
    """C3's shape (K = 8 coefficients + sigma, R = 200) on the tcgen05 step kernel (two
    accumulator chunks of 112 + 96 observations) and on the 4-chains-per-lane FP32-pipe kernel,
    with a chain count that leaves partial warps / lanes."""
    if not tensorCore:
        monkeypatch.setenv("MCMCN_NO_TC", "1")
    obj, names, nResp, ranges = parity.syntheticRegression(G=6, R=200, K=8)
    res = parity.replay(obj, names, 6, nResp, "partial", None, ranges, nChains=133, nIter=25,
                        nSamples=10, precision=precision)
    err, ties = parity.checkReplay(res, tol, 1e-5)
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("G,R,K,ragged,pooling,nChains", [
    (5, 250, 8, False, "partial", 130),      # three accumulator chunks (112 + 112 + 32), two chain blocks
    (9, 40, 3, True, "partial", 5),          # ragged groups of 1..79 observations, K < 8
    (4, 112, 8, False, "none", 1),           # exactly one full chunk, fixed priors, a single chain
    (7, 200, 5, False, "partial", 257),      # three chain blocks, the last with one chain
    (3, 500, 8, False, "partial", 40),       # 48 KB blocks: one TMA stage, five accumulator chunks
    (4, 160, 8, False, "partial", 70),       # two chunks of 112 + 48: the straight-line read-back with a looped second chunk
    (5, 205, 3, True, "partial", 33),        # ragged groups of 1..409 observations: one, two and more chunks side by side
    (4, 200, 12, False, "partial", 70),      # K = 9..16: two K blocks, accumulator chunks of 96 (96 + 96 + 16), 7 MMAs per chunk
    (5, 100, 16, True, "partial", 33),       # all 16 coefficient columns, ragged groups of 1..199 observations
    (3, 250, 9, False, "none", 130),         # one coefficient in the second K block, fixed priors, 33 KB blocks: one TMA stage
])
def test_replay_tensor_core_kernel_shapes(G, R, K, ragged, pooling, nChains):
    """The tcgen05 step kernel over the shapes that change its control flow: number of
    accumulator chunks, padding rows, unused coefficient slots, lanes past the last chain."""
    obj, names, nResp, ranges = parity.syntheticRegression(G=G, R=R, K=K, ragged=ragged)
    prior = None
    if pooling == "none":
        prior = [scipy.stats.norm(0, 10)] * K + [scipy.stats.gamma(2)]
    res = parity.replay(obj, names, G, nResp, pooling, prior, ranges, nChains=nChains, nIter=20,
                        nSamples=10, precision="fp32")
    err, ties = parity.checkReplay(res, 1e-5, 1e-5)
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)


def test_tensor_core_and_fp32_pipe_kernels_agree_at_c3_size(monkeypatch):
    """BASELINE config 3 at full width (1,024 groups x 200 observations x 8 coefficients,
    1,024 chains): one traced iteration on the tcgen05 kernel and on the FP32-pipe kernel from
    the same state and the same replay tape; every proposal log-likelihood must agree
    within 2e-5 relative (each is within 1e-5 of the FP64 oracle on the shapes the oracle can do)."""
    import torch
    from engine import Engine
    obj, names, nResp, ranges = parity.syntheticRegression(G=1024, R=200, K=8)
    nC, G, P = 1024, 1024, 9
    out = {}
    for label, env in (("tc", None), ("pipe", "1")):
        if env:
            monkeypatch.setenv("MCMCN_NO_TC", env)
        eng = Engine(parity.deviceObjective(obj, nResp, "fp32"), G, nResp, "partial", nC, seed=7)
        eng.initialise(names, ranges)
        gen = torch.Generator(device="cpu").manual_seed(11)
        tape = {"z": torch.randn((1, P, G, eng.S), generator=gen, dtype=torch.float64).mul_(0.05).to(eng.device),
                "u": torch.rand((1, P, G, eng.S), generator=gen, dtype=torch.float64).to(eng.device),
                "accept": (torch.rand((1, P, G, eng.S), generator=gen) < 0.4).to(torch.uint8).to(eng.device),
                "zmu": torch.randn((1, P, eng.S), generator=gen, dtype=torch.float64).to(eng.device),
                "qsig": torch.rand((1, P, eng.S), generator=gen, dtype=torch.float64).add_(0.5).to(eng.device)}
        tr = eng.run(0, 1, 0, 1, tape=tape, trace=True)
        torch.cuda.synchronize()
        out[label] = (tr["ll"][..., :nC].cpu().numpy(), eng.getState()["theta"])
    # a proposed sigma <= 0 gives NaN on both kernels (scipy's scale check, :354-356); relErr wants the same NaN
    assert numpy.isfinite(out["tc"][0]).mean() > 0.99
    assert parity.relErr(out["tc"][0], out["pipe"][0]).max() <= 2e-5
    numpy.testing.assert_array_equal(out["tc"][1], out["pipe"][1])      # forced decisions: identical states


def test_replay_more_than_8_coefficients_fp32_pipe_and_fp64(monkeypatch):
    """K = 9..16 off the tensor core: the FP32-pipe kernel (two chains per lane) and the FP64 kernel
    (trajectory-exact), LinReg<12>."""
    obj, names, nResp, ranges = parity.syntheticRegression(G=5, R=60, K=12)
    res = parity.replay(obj, names, 5, nResp, "partial", None, ranges, nChains=3, nIter=60, nSamples=20,
                        precision="fp64", force=False)
    err, ties = parity.checkReplay(res, 1e-11, 0.0)
    assert ties == 0 and not res.engine.usesTensorCore
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)
    monkeypatch.setenv("MCMCN_NO_TC", "1")
    res = parity.replay(obj, names, 5, nResp, "partial", None, ranges, nChains=70, nIter=30, nSamples=10,
                        precision="fp32")
    assert not res.engine.usesTensorCore
    parity.checkReplay(res, 1e-5, 1e-5)
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("workload", ["c3-tcgen05", "c3-fp32-pipe", "c5"])
def test_full_size_iteration_matches_fp64_numpy(workload, monkeypatch):
    """BASELINE configs 3 and 5 at their full data size (1,024 groups x 200 observations x 8
    coefficients; 10,000 groups x 50 trials), 64 chains: one traced iteration with forced decisions.
    Every proposal log-likelihood (P x G x chains of them) must be within 1e-5 relative of the
    objective's formula (example/regression.py:53-67) evaluated in FP64 numpy on
    the same proposals, and the state afterwards must be the forced one."""
    import torch
    from engine import Engine
    nC = 64
    if workload == "c5":
        obj, names, nResp, ranges = parity.syntheticLogit(G=10000, R=50)
        G, R, P = 10000, 50, 2
        x, y = obj.x.reshape(G, R, 1), obj.y.reshape(G, R, 1)

        def loglik(th):                                       # th [P][G][nC] -> [G][nC]
            eta = th[0][:, None, :] + th[1][:, None, :] * x
            return (y * eta - numpy.logaddexp(0.0, eta)).sum(axis=1)
    else:
        if workload == "c3-fp32-pipe":
            monkeypatch.setenv("MCMCN_NO_TC", "1")
        obj, names, nResp, ranges = parity.syntheticRegression(G=1024, R=200, K=8)
        G, R, P = 1024, 200, 9
        X, y = obj.X.reshape(G, R, 8), obj.y.reshape(G, R, 1)

        def loglik(th):
            resid = numpy.einsum("grk,kgn->grn", X, th[:8]) - y
            sg = numpy.where(th[8] > 0, th[8], numpy.nan)     # scipy: scale <= 0 -> nan
            return -0.5 * (resid ** 2).sum(axis=1) / sg ** 2 - R * (numpy.log(sg) + 0.5 * numpy.log(2 * numpy.pi))
    eng = Engine(parity.deviceObjective(obj, nResp, "fp32"), G, nResp, "partial", nC, seed=5)
    eng.initialise(names, ranges)
    assert bool(eng.usesTensorCore) == (workload == "c3-tcgen05")
    st0 = eng.getState()
    gen = torch.Generator(device="cpu").manual_seed(13)
    z = torch.randn((1, P, G, eng.S), generator=gen, dtype=torch.float64).mul_(0.05)
    acc = (torch.rand((1, P, G, eng.S), generator=gen) < 0.4).to(torch.uint8)
    tape = {"z": z.to(eng.device), "u": torch.rand((1, P, G, eng.S), generator=gen, dtype=torch.float64).to(eng.device),
            "accept": acc.to(eng.device),
            "zmu": torch.randn((1, P, eng.S), generator=gen, dtype=torch.float64).to(eng.device),
            "qsig": torch.rand((1, P, eng.S), generator=gen, dtype=torch.float64).add_(0.5).to(eng.device)}
    tr = eng.run(0, 1, 0, 1, tape=tape, trace=True)
    torch.cuda.synchronize()
    got = tr["ll"][0, ..., :nC].cpu().numpy()
    theta = st0["theta"].copy()
    zz, aa = z[0, ..., :nC].numpy(), acc[0, ..., :nC].numpy() != 0
    worst = 0.0
    for p in range(P):
        prop = theta[p] + st0["scale"][p] * zz[p]
        cand = theta.copy()
        cand[p] = prop
        worst = max(worst, parity.relErr(got[p], loglik(cand)).max())
        theta[p] = numpy.where(aa[p], prop, theta[p])
    assert worst <= 1e-5, worst
    numpy.testing.assert_array_equal(eng.getState()["theta"], theta)


def test_tensor_core_and_fp32_pipe_kernels_draw_the_same_random_numbers(monkeypatch):
    """Free-running, same seed: both step kernels take their proposals and uniforms from the same
    Philox counters (one call per two sweeps), so after a few iterations the two chains' states
    are bit-identical except where a rounding-level difference of the log-likelihood flipped a
    near-threshold decision."""
    from engine import Engine
    obj, names, nResp, ranges = parity.syntheticRegression(G=8, R=200, K=8)
    out = {}
    for label, env in (("tc", None), ("pipe", "1")):
        if env:
            monkeypatch.setenv("MCMCN_NO_TC", env)
        eng = Engine(parity.deviceObjective(obj, nResp, "fp32"), 8, nResp, "partial", 160, seed=3)
        eng.initialise(names, ranges)
        eng.run(0, 3, 0, 1)
        out[label] = eng.getState()
    assert numpy.isfinite(out["tc"]["theta"]).all()
    assert (out["tc"]["theta"] == out["pipe"]["theta"]).mean() > 0.995


@pytest.mark.parametrize("pooling", ["partial", "none"])
def test_replay_bernoulli_logit(pooling):
    obj, names, nResp, ranges = parity.syntheticLogit(G=30, R=50)
    prior = [scipy.stats.norm(0, 5), scipy.stats.norm(0, 5)] if pooling == "none" else None
    res = parity.replay(obj, names, 30, nResp, pooling, prior, ranges, nChains=5, nIter=120,
                        nSamples=40, precision="fp32")
    err, ties = parity.checkReplay(res, 1e-5, 1e-5)
    res64 = parity.replay(obj, names, 30, nResp, pooling, prior, ranges, nChains=2, nIter=120,
                          nSamples=40, precision="fp64", force=False)
    err64, ties64 = parity.checkReplay(res64, 1e-11, 0.0)
    assert ties64 == 0


@pytest.mark.parametrize("P,pooling", [(5, "partial"), (8, "partial"), (6, "none"), (7, "complete")])
def test_replay_gaussian_distribution_with_5_to_8_parameters(P, pooling):
    """The Gaussian-distribution objective (example/distribution.py:18-24) beyond the example's own sizes: 5..8
    parameters (GaussDist<5..8>, csrc/mcmcn_sets_gauss_b.cu), ragged groups, all three pooling modes."""
    rs = numpy.random.RandomState(100 + P)
    G = 9
    nResp = [int(v) for v in rs.randint(3, 12, size=G)]
    mu = rs.normal(0, 2, size=(P, G))
    sd = rs.uniform(0.5, 2.0, size=P)
    obj = parity.po.GaussianDistributionObjective(mu, sd, nResp)
    names = tuple("m%d" % j for j in range(P))
    ranges = dict((n, [-3, 3]) for n in names)
    prior = None if pooling == "partial" else [scipy.stats.norm(0, 10)] * P
    res64 = parity.replay(obj, names, G, nResp, pooling, prior, ranges, nChains=3, nIter=90, nSamples=30,
                          precision="fp64", force=False)
    err64, ties64 = parity.checkReplay(res64, 1e-11, 0.0)
    assert ties64 == 0
    numpy.testing.assert_allclose(res64.rows, res64.oracleRows, rtol=1e-9, atol=1e-9)
    res = parity.replay(obj, names, G, nResp, pooling, prior, ranges, nChains=3, nIter=90, nSamples=30, precision="fp32")
    parity.checkReplay(res, 1e-5, 1e-5)


@pytest.mark.parametrize("objective", ["regression", "regression-8", "regression-12", "regression-fp32-pipe", "logit"])
def test_replay_complete_pooling_split_over_observations(objective, monkeypatch):
    """Complete pooling at scale: the engine evaluates the single group of all N observations as
    small groups in parallel (mcmcn_model.split: propose / eval / decide kernels, the next proposal
    formed by the decide launch) -- same tape replay criteria as the step kernels.  Linear regression
    with FP32 observation math evaluates on the tensor core (eval_tc_kernel: groups of 112 observations,
    96 for K > 8, all centred on the pooled least-squares fit), everything else with the FP32-pipe
    eval_kernel over groups of 128; N = 5,000 leaves a short last group."""
    # (coefficients, groups x observations of the data, chains, iterations, retained): the oracle runs chain by
    # chain, so the wide case (two 128-chain blocks) is short
    K, G, R, nC, nIter, nSamples = {"regression": (2, 50, 100, 37, 60, 20), "regression-8": (8, 20, 110, 130, 12, 6),
                                    "regression-12": (12, 20, 110, 9, 30, 10), "regression-fp32-pipe": (2, 50, 100, 5, 40, 20),
                                    "logit": (0, 50, 100, 37, 60, 20)}[objective]
    if K:
        if objective == "regression-fp32-pipe":
            monkeypatch.setenv("MCMCN_NO_TC", "1")
        obj, names, nResp, ranges = parity.syntheticRegression(G=G, R=R, K=K)
        prior = [scipy.stats.norm(0, 10)] * K + [scipy.stats.gamma(2)]
    else:
        obj, names, nResp, ranges = parity.syntheticLogit(G=G, R=R)
        prior = [scipy.stats.norm(0, 5), scipy.stats.cauchy(0, 5)]
    res = parity.replay(obj, names, G, nResp, "complete", prior, ranges, nChains=nC, nIter=nIter,
                        nSamples=nSamples, precision="fp32")
    assert res.engine.model.split
    err, ties = parity.checkReplay(res, 1e-5, 1e-5)
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)
    res = parity.replay(obj, names, G, nResp, "complete", prior, ranges, nChains=3, nIter=nIter,
                        nSamples=nSamples, precision="fp64", force=False)
    err, ties = parity.checkReplay(res, 1e-11, 0.0)
    assert ties == 0
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-12), ("fp32", 1e-5)])
def test_complete_pooling_loglik_split_equals_single_group(precision, tol, monkeypatch):
    """The log-likelihood the start-state search and the MLE start evaluate under complete pooling:
    summed over groups of 128 (split) it must equal the single group of all N observations, for the
    current state and for an explicit pooled vector."""
    from engine import Engine
    obj, names, nResp, ranges = parity.syntheticRegression(G=40, R=77, K=3)
    prior = [scipy.stats.norm(0, 10)] * 3 + [scipy.stats.gamma(2)]
    eng = Engine(parity.deviceObjective(obj, nResp, precision), 40, nResp, "complete", 9, priorDistribution=prior, seed=2)
    assert eng.model.split
    rs = numpy.random.RandomState(4)
    theta = rs.normal(size=(4, 1, 9))
    theta[3] = 0.5 + rs.random_sample((1, 9))
    eng.setState(theta, numpy.zeros((1, 9)), numpy.zeros((4, 1, 9)))
    pooled = rs.normal(size=(4, 9))
    pooled[3] = 1.0 + rs.random_sample(9)
    got = [eng.groupLogLikelihood()[0, :9].cpu().numpy(), eng.groupLogLikelihood(pooled)[0, :9].cpu().numpy(), eng.pooledNll(pooled)]
    monkeypatch.setenv("MCMCN_NO_SPLIT", "1")
    want = [eng.groupLogLikelihood()[0, :9].cpu().numpy(), eng.groupLogLikelihood(pooled)[0, :9].cpu().numpy(), eng.pooledNll(pooled)]
    for g, w in zip(got, want):
        assert numpy.isfinite(w).all()
        assert parity.relErr(g, w).max() <= tol


def test_replay_streaming_group_larger_than_tile(monkeypatch):
    """Complete pooling makes one group of all N observations; with N beyond the shared-memory
    tile the block is streamed through it by repeated TMA copies (the path below the split
    threshold, forced here with MCMCN_NO_SPLIT)."""
    monkeypatch.setenv("MCMCN_NO_SPLIT", "1")
    obj, names, nResp, ranges = parity.syntheticRegression(G=50, R=100, K=2)
    prior = [scipy.stats.norm(0, 10), scipy.stats.norm(0, 10), scipy.stats.gamma(2)]
    res = parity.replay(obj, names, 50, nResp, "complete", prior, ranges, nChains=3, nIter=40,
                        nSamples=20, precision="fp32")
    err, ties = parity.checkReplay(res, 1e-5, 1e-5)
    res = parity.replay(obj, names, 50, nResp, "complete", prior, ranges, nChains=2, nIter=40,
                        nSamples=20, precision="fp64", force=False)
    parity.checkReplay(res, 1e-11, 0.0)


@pytest.mark.parametrize("families", ["uniform-expon-halfnorm", "cauchy-t-lognorm", "laplace-logistic-invgamma",
                                      "beta-norm-chi2"])
def test_ragged_groups_and_prior_families(families):
    """Every device prior family against frozen scipy distributions (what the reference takes in
    priorDistribution, posteriorSampling.py:293-294): proposal log-prior within 1e-12, FP64
    trajectory identical."""
    obj, names, nResp, ranges = parity.syntheticRegression(G=9, R=7, K=2, ragged=True)
    prior = {"uniform-expon-halfnorm": [scipy.stats.uniform(-20, 40), scipy.stats.expon(-6, 4), scipy.stats.halfnorm(0, 3)],
             "cauchy-t-lognorm": [scipy.stats.cauchy(0, 5), scipy.stats.t(4, 1, 3), scipy.stats.lognorm(0.8, scale=2)],
             "laplace-logistic-invgamma": [scipy.stats.laplace(0, 4), scipy.stats.logistic(1, 3), scipy.stats.invgamma(3, scale=2)],
             "beta-norm-chi2": [scipy.stats.beta(2, 3, loc=-30, scale=60), scipy.stats.norm(0, 10), scipy.stats.chi2(4)]}[families]
    res = parity.replay(obj, names, 9, nResp, "none", prior, ranges, nChains=3, nIter=150,
                        nSamples=50, precision="fp64", force=False)
    err, ties = parity.checkReplay(res, 1e-11, 0.0)
    assert ties == 0
