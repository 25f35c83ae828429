"""Parity of the PRODUCTION kernel instantiations -- the ones bench.py times.

The replay tests (test_gpu_replay.py) need tapes and traces and therefore drive the general
instantiations (``sweep_tc_kernel<-1, *>``, ``sweep_kernel<.., -1>``); free-running chains run
``sweep_tc_kernel<0..3, *>`` / ``sweep_kernel<.., 0..3>`` and, from 512 groups on,
``hyper_onepass_kernel``.  Three links close the gap:

  1. same seed, same start state: production and general instantiations (``MCMCN_GENERAL=1``)
     must leave BIT-IDENTICAL theta / ll / scale / counts / hyper / retained rows, across a tune
     iteration and across the end of burn-in, for every kernel family;
  2. the one-pass Gibbs kernel meets the oracle directly in a tape replay with 512 groups;
  3. the random numbers themselves: Philox4x32-10 known answers (Random123), and the device
     normals, uniforms and inverse-gamma draws against their distributions
     (posteriorSampling.py:304-306, :362, :485-498).
"""

import ctypes

import numpy
import pytest
import scipy.stats

import parity

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ 1. production == general
def _runBoth(monkeypatch, makeEngine, names, ranges, nIter, burn, thin, env=()):
    """Run the same chains twice from the same start state and seed: production instantiations,
    then MCMCN_GENERAL=1.  Returns the two final states (+ retained rows)."""
    import torch
    from engine import SampleStore
    out = []
    for general in (False, True):
        for k, v in env:
            monkeypatch.setenv(k, v)
        if general:
            monkeypatch.setenv("MCMCN_GENERAL", "1")
        else:
            monkeypatch.delenv("MCMCN_GENERAL", raising=False)
        eng = makeEngine()
        eng.initialise(names, ranges)
        nRows = len([i for i in range(nIter) if i % thin == 0 and i >= burn])
        store = SampleStore(eng, nRows, torch.float64)
        # two calls: the second starts inside burn-in, so both the counting and the plain variants run
        eng.run(0, 60, burn, thin, store=store)
        eng.run(60, nIter - 60, burn, thin, store=store)
        torch.cuda.synchronize()
        st = eng.getState()
        st["counts"] = eng.counts[..., :eng.nChains].cpu().numpy()
        st["rows"] = store.hostArray()
        st["tc"] = eng.usesTensorCore
        out.append(st)
    monkeypatch.delenv("MCMCN_GENERAL", raising=False)
    return out


def _assertIdentical(prod, gen):
    for key in ("theta", "ll", "scale", "counts", "rows", "lprior", "mu", "sigma2"):
        if key in prod:
            a, b = prod[key], gen[key]
            assert a.shape == b.shape, key
            same = (a == b) | (numpy.isnan(a) & numpy.isnan(b)) if a.dtype.kind == "f" else (a == b)
            assert same.all(), "%s differs in %d of %d entries" % (key, int((~same).sum()), same.size)
    assert numpy.isfinite(prod["theta"]).all()
    # step sizes were tuned and rows differ from one another: the comparison is not vacuous
    assert (prod["scale"] != 1.0).any() and (prod["rows"][0] != prod["rows"][-1]).mean() > 0.5


@pytest.mark.parametrize("label,R,ragged,pooling,noTc", [
    ("tc-two-k-blocks", 200, False, "partial", False),   # K = 12: sweep_tc_kernel<F, false, 2>
    ("tc-uniform208", 200, False, "partial", False),     # C3's group shape: sweep_tc_kernel<F, true>
    ("tc-uniform208-fixed-priors", 200, False, "none", False),
    ("tc-any", 160, False, "partial", False),            # two chunks, second looped: sweep_tc_kernel<F, false>
    ("tc-ragged", 90, True, "partial", False),           # one, two and more chunks side by side
    ("fp32-pipe", 200, False, "partial", True),          # sweep_kernel<LinReg<8>, 4, float, 3, F, 128>
    ("fp32-pipe-fixed-priors", 200, False, "none", True),
])
def test_production_regression_kernels_equal_general_bit_for_bit(label, R, ragged, pooling, noTc, monkeypatch):
    from engine import Engine
    G, K, nC = 12, (12 if label == "tc-two-k-blocks" else 8), 161
    obj, names, nResp, ranges = parity.syntheticRegression(G=G, R=R, K=K, ragged=ragged)
    prior = [scipy.stats.norm(0, 10)] * K + [scipy.stats.gamma(2)] if pooling == "none" else None
    env = (("MCMCN_NO_TC", "1"),) if noTc else ()
    if not noTc:
        monkeypatch.delenv("MCMCN_NO_TC", raising=False)

    def make():
        return Engine(parity.deviceObjective(obj, nResp, "fp32"), G, nResp, pooling, nC,
                      priorDistribution=prior, chainId0=1000, seed=42)
    prod, gen = _runBoth(monkeypatch, make, names, ranges, nIter=230, burn=150, thin=4, env=env)
    assert prod["tc"] == (not noTc)
    _assertIdentical(prod, gen)


@pytest.mark.parametrize("pooling", ["partial", "none"])
def test_production_logit_kernel_equals_general_bit_for_bit(pooling, monkeypatch):
    from engine import Engine
    G, R, nC = 40, 50, 300
    obj, names, nResp, ranges = parity.syntheticLogit(G=G, R=R)
    prior = [scipy.stats.norm(0, 5), scipy.stats.norm(0, 5)] if pooling == "none" else None

    def make():
        return Engine(parity.deviceObjective(obj, nResp, "fp32"), G, nResp, pooling, nC,
                      priorDistribution=prior, chainId0=7, seed=9)
    prod, gen = _runBoth(monkeypatch, make, names, ranges, nIter=230, burn=150, thin=4)
    _assertIdentical(prod, gen)


def test_results_do_not_depend_on_the_task_size(monkeypatch):
    """MCMCN_TASK_OBS (observations per CTA task of the FP32-pipe step kernel) is a tuning knob: a task is a run
    of whole groups, a group's sum is formed by one thread either way."""
    import torch
    from engine import Engine
    G, R, nC = 40, 50, 130
    obj, names, nResp, ranges = parity.syntheticLogit(G=G, R=R)
    states = []
    for taskObs in (None, "50", "1000"):
        if taskObs is None:
            monkeypatch.delenv("MCMCN_TASK_OBS", raising=False)
        else:
            monkeypatch.setenv("MCMCN_TASK_OBS", taskObs)
        eng = Engine(parity.deviceObjective(obj, nResp, "fp32"), G, nResp, "partial", nC, chainId0=3, seed=11)
        eng.initialise(names, ranges)
        eng.run(0, 120, 100, 2)
        torch.cuda.synchronize()
        states.append((eng.model.n_tasks, eng.getState()))
    monkeypatch.delenv("MCMCN_TASK_OBS", raising=False)
    assert states[1][0] == G and states[2][0] < states[0][0] < G          # one group per task ... many
    for _, st in states[1:]:
        for key in ("theta", "ll", "scale", "mu", "sigma2"):
            assert (st[key] == states[0][1][key]).all(), key


def test_production_kernels_with_many_groups_equal_general_bit_for_bit(monkeypatch):
    """G >= 512: the production run also takes hyper_onepass_kernel; MCMCN_GENERAL only swaps the
    step kernel, so both arms use the same Gibbs kernel and must agree bit for bit."""
    from engine import Engine
    G, R, K, nC = 520, 24, 3, 64
    obj, names, nResp, ranges = parity.syntheticRegression(G=G, R=R, K=K)

    def make():
        return Engine(parity.deviceObjective(obj, nResp, "fp32"), G, nResp, "partial", nC, chainId0=0, seed=5)
    prod, gen = _runBoth(monkeypatch, make, names, ranges, nIter=130, burn=110, thin=2)
    _assertIdentical(prod, gen)


@pytest.mark.parametrize("noTc", [False, True])
def test_same_seed_twice_is_bit_identical(noTc, monkeypatch):
    """No atomics, no order-dependent reductions: two runs from the same seed and start give the same bits
    (compute-sanitizer's racecheck is not available on the GPU pool; a shared-memory, tensor-memory or
    TMA-stage hazard in the step kernels would show up here as a difference between runs).  C3's group
    shape, four chain blocks, group ranges of several groups, across a tune iteration."""
    import torch
    from engine import Engine, SampleStore
    if noTc:
        monkeypatch.setenv("MCMCN_NO_TC", "1")
    else:
        monkeypatch.delenv("MCMCN_NO_TC", raising=False)
    G, R, K, nC = 96, 200, 8, 500
    obj, names, nResp, ranges = parity.syntheticRegression(G=G, R=R, K=K)
    out = []
    for rep in range(3):
        eng = Engine(parity.deviceObjective(obj, nResp, "fp32"), G, nResp, "partial", nC, chainId0=5, seed=77)
        eng.initialise(names, ranges)
        store = SampleStore(eng, 10, torch.float64)
        eng.run(0, 140, 120, 2, store=store)
        torch.cuda.synchronize()
        st = eng.getState()
        out.append((st["theta"], st["ll"], st["scale"], st["mu"], st["sigma2"], store.hostArray()))
        del eng, store
    for other in out[1:]:
        for a, b in zip(out[0], other):
            numpy.testing.assert_array_equal(a, b)


def test_results_do_not_depend_on_how_many_chains_share_the_gpu():
    """Philox and the start-state streams are keyed by the global chain id, and every sum over groups is taken
    in a fixed order whatever the launch shape (the one-pass Gibbs kernel picks its block shape from the chain
    count: 32, 64 or 128 chains per block row), so chains 0..63 come out bit-identical whether they run alone
    or next to 9,408 others -- the property that makes a run independent of the number of GPUs."""
    import torch
    from engine import Engine
    G, R, K = 520, 12, 3
    obj, names, nResp, ranges = parity.syntheticRegression(G=G, R=R, K=K)
    out = []
    for nC in (64, 4736, 9472):                        # hyper_onepass_kernel<32,1>, <64,1>, <128,2>
        eng = Engine(parity.deviceObjective(obj, nResp, "fp32"), G, nResp, "partial", nC, chainId0=0, seed=11)
        eng.initialise(names, ranges)
        eng.run(0, 130, 110, 2)
        torch.cuda.synchronize()
        st = eng.getState()
        out.append(dict((k, v[..., :64].copy()) for k, v in st.items()))
        del eng
        torch.cuda.empty_cache()
    for other in out[1:]:
        for key in ("theta", "ll", "scale", "mu", "sigma2"):
            numpy.testing.assert_array_equal(out[0][key], other[key], err_msg=key)


# ------------------------------------------------------------------ 2. one-pass Gibbs kernel vs the oracle
@pytest.mark.parametrize("precision,tol", [("fp64", 1e-11), ("fp32", 1e-5)])
def test_replay_with_512_groups_meets_the_one_pass_hyper_kernel(precision, tol, monkeypatch):
    """From 512 groups on mcmcn_run updates the hyper-parameters with hyper_onepass_kernel (shifted
    single-pass sums).  Tape replay against the oracle: mu / sigma2 in the retained rows, the
    log-priors they imply and (FP64) the whole trajectory."""
    monkeypatch.delenv("MCMCN_HYPER_TWO_PASS", raising=False)
    G, R, K = 512, 6, 2
    obj, names, nResp, ranges = parity.syntheticRegression(G=G, R=R, K=K)
    res = parity.replay(obj, names, G, nResp, "partial", None, ranges, nChains=3, nIter=40, nSamples=20,
                        precision=precision, force=(precision == "fp32"))
    err, ties = parity.checkReplay(res, tol, 0.0 if precision == "fp64" else 1e-5)
    if precision == "fp64":
        assert ties == 0
    numpy.testing.assert_allclose(res.rows, res.oracleRows, rtol=1e-9, atol=1e-9)
    per = G + 2
    hyperCols = [p * per + j for p in range(K + 1) for j in (0, 1)]
    numpy.testing.assert_allclose(res.rows[:, hyperCols, :], res.oracleRows[:, hyperCols, :], rtol=1e-10, atol=1e-12)
    # ... and the two-pass kernel on the same tape gives the same rows to rounding
    monkeypatch.setenv("MCMCN_HYPER_TWO_PASS", "1")
    res2 = parity.replay(obj, names, G, nResp, "partial", None, ranges, nChains=3, nIter=40, nSamples=20,
                         precision=precision, force=(precision == "fp32"))
    numpy.testing.assert_allclose(res.rows[:, hyperCols, :], res2.rows[:, hyperCols, :], rtol=1e-10, atol=1e-12)


# ------------------------------------------------------------------ 3. random numbers
def _draws(kind, n, seed=12345, a=0.0, pair=False):
    import torch
    import mcmcn_native as nat
    out = torch.empty((2 * n if pair else n,), dtype=torch.float64, device="cuda")
    nat.call("mcmcn_debug_draws", kind, n, seed, float(a), ctypes.c_void_p(out.data_ptr()), None)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_philox4x32_10_known_answers():
    """Random123's known-answer vectors for philox4x32-10 (kat_vectors), through the device function
    the step kernels call."""
    import torch
    import mcmcn_native as nat
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c = torch.tensor(numpy.array(ctr, dtype=numpy.uint32).view(numpy.int32), device="cuda")
        k = torch.tensor(numpy.array(key, dtype=numpy.uint32).view(numpy.int32), device="cuda")
        o = torch.zeros(4, dtype=torch.int32, device="cuda")
        nat.call("mcmcn_debug_philox", ctypes.c_void_p(c.data_ptr()), ctypes.c_void_p(k.data_ptr()),
                 ctypes.c_void_p(o.data_ptr()))
        got = tuple(int(v) for v in o.cpu().numpy().view(numpy.uint32))
        assert got == want, (hex(got[0]), hex(want[0]))


def _momentsOk(x, mean, var, label, excessKurtosis=0.0):
    """Sample mean and variance within 5 / 6 standard errors (Var s^2 = (2 + excess kurtosis) var^2 / n)."""
    n = x.size
    assert abs(x.mean() - mean) < 5 * numpy.sqrt(var / n), (label, x.mean())
    assert abs(x.var() - var) < 6 * var * numpy.sqrt((2.0 + excessKurtosis) / n), (label, x.var())


def test_device_normals_are_standard_normal():
    import mcmcn_native as nat
    n = 1 << 20
    z = _draws(nat.DRAW_SWEEP_NORMALS, n, pair=True)
    zc, zs = z[0::2], z[1::2]
    for lab, v in (("cos branch", zc), ("sin branch", zs), ("hyper", _draws(nat.DRAW_HYPER_NORMAL, n))):
        assert numpy.isfinite(v).all()
        _momentsOk(v, 0.0, 1.0, lab)
        assert abs(scipy.stats.skew(v)) < 5 * numpy.sqrt(6.0 / n), lab
        assert abs(scipy.stats.kurtosis(v)) < 5 * numpy.sqrt(24.0 / n), lab
        ks = scipy.stats.kstest(v[:200000], "norm")
        assert ks.pvalue > 1e-4, (lab, ks)
        # tails: P(|z| > 3) = 2.6998e-3
        tail = (numpy.abs(v) > 3).mean()
        assert abs(tail - 2.6998e-3) < 5 * numpy.sqrt(2.6998e-3 / n), (lab, tail)
    # the two branches of one Box-Muller transform are independent
    assert abs(numpy.corrcoef(zc, zs)[0, 1]) < 5 / numpy.sqrt(n)
    assert abs(numpy.corrcoef(zc ** 2, zs ** 2)[0, 1]) < 5 / numpy.sqrt(n)


def test_device_uniforms_are_uniform_on_the_open_interval():
    import mcmcn_native as nat
    n = 1 << 20
    u = _draws(nat.DRAW_SWEEP_UNIFORMS, n, pair=True)
    u53 = _draws(nat.DRAW_UNIFORM53, n)
    assert (u > 0).all() and (u < 1).all()                 # log(u) is finite: branch 4/5 never sees log(0)
    assert (u53 >= 0).all() and (u53 < 1).all()
    for lab, v in (("word z", u[0::2]), ("word w", u[1::2]), ("53 bit", u53)):
        _momentsOk(v, 0.5, 1.0 / 12.0, lab, -1.2)
        ks = scipy.stats.kstest(v[:200000], "uniform")
        assert ks.pvalue > 1e-4, (lab, ks)
    assert abs(numpy.corrcoef(u[0::2], u[1::2])[0, 1]) < 5 / numpy.sqrt(n)
    # (w + 0.5) 2^-32 exactly: 32-bit resolution, symmetric about 1/2
    w = u * 4294967296.0 - 0.5
    assert (w == numpy.round(w)).all()


@pytest.mark.parametrize("G", [2, 3, 10, 50, 1024])
def test_device_inverse_gamma_draw_matches_scipy_invgamma(G):
    """The free-running sigma2 draw is q * (a * hat), q = 1 / Gamma(a, 1) by Marsaglia-Tsang on the
    chain's Philox stream, a = (G - 1) / 2; the reference draws scipy.stats.invgamma(a, scale=a * hat)
    by inverse CDF (posteriorSampling.py:489-498).  Same law: KS against scipy's CDF, and the
    moments of 1 / q (Gamma(a, 1): mean a, variance a)."""
    import mcmcn_native as nat
    a = (G - 1) / 2.0
    n = 1 << 19
    q = _draws(nat.DRAW_UNIT_INVGAMMA, n, seed=777 + G, a=a)
    assert numpy.isfinite(q).all() and (q > 0).all()
    ks = scipy.stats.kstest(q[:200000], scipy.stats.invgamma(a).cdf)
    assert ks.pvalue > 1e-4, ks
    g = 1.0 / q
    _momentsOk(g, a, a, "gamma(%g)" % a, 6.0 / a)
    # a second seed gives a different, equally valid stream
    q2 = _draws(nat.DRAW_UNIT_INVGAMMA, 1000, seed=778 + G, a=a)
    assert (q2 != q[:1000]).mean() > 0.99
