"""Shared helpers for the GPU parity tests and __graft_entry__.smoke().

The oracle (oracle/) is the checker here, never the thing measured: every engine call
below goes through the product package and the C ABI.
"""

import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mcmc-for-nested-data_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from oracle import posterior_oracle as po  # noqa: E402


def deviceObjective(oracleObjective, nResponsesPerGroup, precision):
    """The product-side Objective handle for the same data as an oracle objective."""
    from objectives import Objective
    if isinstance(oracleObjective, po.LinearRegressionObjective):
        return Objective.linear_regression(oracleObjective.X, oracleObjective.y, precision)
    if isinstance(oracleObjective, po.BernoulliLogitObjective):
        return Objective.bernoulli_logit(oracleObjective.x, oracleObjective.y, precision)
    if isinstance(oracleObjective, po.GaussianDistributionObjective):
        return Objective.gaussian_distribution(oracleObjective.mu, oracleObjective.sd,
                                               nResponsesPerGroup, precision)
    raise TypeError(type(oracleObjective))


def syntheticRegression(G, R, K, seed=20261018, ragged=False):
    """SURVEY.md section 8d config C3's generator at arbitrary size."""
    rs = numpy.random.RandomState(seed)
    nResp = [R] * G
    if ragged:
        nResp = [int(v) for v in rs.randint(1, 2 * R, size=G)]
    N = sum(nResp)
    X = numpy.ones((N, K))
    X[:, 1:] = rs.normal(size=(N, K - 1)).astype(numpy.float32)
    mu = numpy.arange(K) - 3.5
    beta = rs.normal(mu, 1.0, size=(G, K))
    gi = numpy.repeat(numpy.arange(G), nResp)
    y = numpy.sum(X * beta[gi], axis=1) + rs.normal(size=N)
    names = tuple("b%d" % k for k in range(K)) + ("sigma",)
    ranges = dict((n, [-5, 5]) for n in names[:-1])
    ranges["sigma"] = [0.5, 2]
    return po.LinearRegressionObjective(X, y), names, nResp, ranges


def syntheticLogit(G, R, seed=20261019):
    """SURVEY.md section 8d config C5's generator at arbitrary size."""
    rs = numpy.random.RandomState(seed)
    N = G * R
    x = rs.normal(size=N)
    a = rs.normal(0, 1, size=G)
    b = rs.normal(1, 0.5, size=G)
    gi = numpy.repeat(numpy.arange(G), R)
    eta = a[gi] + b[gi] * x
    y = (rs.random_sample(N) < 1 / (1 + numpy.exp(-eta))).astype(float)
    return po.BernoulliLogitObjective(x, y), ("a", "b"), [R] * G, {"a": [-2, 2], "b": [-1, 3]}


class ReplayResult(object):
    pass


def replay(objective, names, nGroups, nResp, pooling, prior, ranges, nChains, nIter, nSamples,
           precision="fp32", force=None, startWithMLE=False, chainId0=0):
    """Run the oracle for nChains chains recording tapes, then replay the tapes through the
    CUDA engine from the oracle's start state.  force=None -> teacher-force the oracle's
    decisions only for fp32 (documented near-threshold ties); the engine's own decisions
    are recorded either way."""
    import torch
    from engine import Engine, SampleStore
    if force is None:
        force = precision == "fp32"
    chains = []
    start = []
    for c in range(nChains):
        oc = po.OracleChain(chainId0 + c, chainId0 + c, nIter, nSamples, names, nGroups, nResp, pooling,
                            objective, prior, startWithMLE, ranges, recordTape=True)
        start.append(dict(value=oc.value.copy(), logPrior=oc.logPrior.copy(), LL=oc.LL.copy(),
                          mu=oc.mu.copy(), sigma2=oc.sigma2.copy()))
        oc.rows = oc.run()
        chains.append(oc)
    oc0 = chains[0]
    P, G = oc0.P, oc0.G
    burn, thin = oc0.burn, oc0.thin

    eng = Engine(deviceObjective(objective, nResp, precision), nGroups, nResp, pooling, nChains,
                 priorDistribution=prior, chainId0=chainId0)
    stack = lambda key: numpy.stack([s[key] for s in start], axis=-1)
    eng.setState(stack("value"), stack("LL"), stack("logPrior"), stack("mu"), stack("sigma2"))

    S, dev = eng.S, eng.device

    def tens(field, shape, dtype=torch.float64):
        t = torch.zeros(shape + (S,), dtype=dtype, device=dev)
        arr = numpy.stack([getattr(c.tape, field) for c in chains], axis=-1)
        t[..., :nChains] = torch.from_numpy(arr).to(dev).to(dtype)
        return t

    tape = {"z": tens("z_prop", (nIter, P, G)), "u": tens("u_acc", (nIter, P, G))}
    # a uniform that the oracle never drew is never needed by an identical trajectory; feed 0.5
    tape["u"] = torch.nan_to_num(tape["u"], nan=0.5)
    if force:
        tape["accept"] = tens("accept", (nIter, P, G), torch.uint8)
    if pooling == "partial":
        tape["zmu"] = tens("z_mu", (nIter, P))
        tape["qsig"] = tens("q_sig", (nIter, P))
    store = SampleStore(eng, len(oc0.rows), torch.float64)
    tr = eng.run(0, nIter, burn, thin, store=store, tape=tape, trace=True, useLpriorOverride=(pooling == "partial"))
    torch.cuda.synchronize()

    res = ReplayResult()
    res.engine, res.chains, res.store = eng, chains, store
    res.trace = dict((k, v[..., :nChains].cpu().numpy()) for k, v in tr.items())
    res.oracle = dict((f, numpy.stack([getattr(c.tape, f) for c in chains], axis=-1))
                      for f in ("ll_prop", "lp_prop", "diff", "accept", "u_acc"))
    res.rows = store.hostArray()                                       # [rows][ncol][nC]
    res.oracleRows = numpy.stack([numpy.stack([r[1] for r in c.rows]) for c in chains], axis=-1)
    res.final = eng.getState()
    res.oracleFinal = numpy.stack([c.value for c in chains], axis=-1)
    return res


def relErr(a, b):
    """Element-wise relative error with exact agreement required on non-finite values."""
    a, b = numpy.asarray(a, dtype=float), numpy.asarray(b, dtype=float)
    fin = numpy.isfinite(a) & numpy.isfinite(b)
    same = (numpy.isnan(a) & numpy.isnan(b)) | (a == b)
    err = numpy.zeros(a.shape)
    err[fin] = numpy.abs(a[fin] - b[fin]) / numpy.maximum(numpy.abs(b[fin]), 1.0)
    err[~fin & ~same] = numpy.inf
    return err


def checkReplay(res, llTol, tieTol):
    """Assert the north-star replay criteria; returns (max log-density error, number of ties)."""
    e_ll = relErr(res.trace["ll"], res.oracle["ll_prop"])
    e_lp = relErr(res.trace["lp"], res.oracle["lp_prop"])
    assert e_ll.max() <= llTol, "proposal log-likelihood off by %g" % e_ll.max()
    assert e_lp.max() <= 1e-12, "proposal log-prior off by %g" % e_lp.max()
    mism = res.trace["accept"] != res.oracle["accept"]
    if mism.any():
        with numpy.errstate(all="ignore"):
            margin = numpy.abs(numpy.log(res.oracle["u_acc"]) - res.oracle["diff"])
        scale = numpy.maximum(numpy.abs(res.oracle["ll_prop"]), 1.0)
        assert numpy.all(margin[mism] <= tieTol * scale[mism]), \
            "accept/reject differs away from the threshold: margins %r" % (margin[mism][:5],)
    return float(e_ll.max()), int(mism.sum())


def smokeCheck(verbose=False):
    """One small replay (fp32, teacher-forced) + a few free-running iterations."""
    import torch
    obj, names, nResp, ranges = syntheticRegression(G=12, R=10, K=2)
    res = replay(obj, names, 12, nResp, "partial", None, ranges, nChains=3, nIter=60, nSamples=20)
    err, ties = checkReplay(res, 1e-5, 1e-5)
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-6, atol=1e-6)
    eng = res.engine
    eng.run(60, 20, 30, 2)          # free-running Philox iterations on the same engine
    torch.cuda.synchronize()
    st = eng.getState()
    assert numpy.isfinite(st["theta"]).all() and numpy.isfinite(st["ll"]).all()
    # the bench shape of a group (200 observations x 8 coefficients: two accumulator chunks of the
    # tcgen05 step kernel), a few groups and chains
    obj2, names2, nResp2, ranges2 = syntheticRegression(G=3, R=200, K=8)
    res2 = replay(obj2, names2, 3, nResp2, "partial", None, ranges2, nChains=5, nIter=12, nSamples=6)
    err2, ties2 = checkReplay(res2, 1e-5, 1e-5)
    assert res2.engine.usesTensorCore
    if verbose:
        print("smoke ok: replay max rel log-density error %.3g / %.3g (C3 group shape, tcgen05 kernel), "
              "%d near-threshold ties" % (err, err2, ties + ties2))
    return max(err, err2), ties + ties2
