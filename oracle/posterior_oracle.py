"""CPU oracle of the sampler hot path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Array-based numpy restatement of the reference's one-parameter-at-a-time
random-walk Metropolis-within-Gibbs sampler.  It consumes the legacy MT19937
stream in exactly the reference's order, so for a given (chain id, inputs)
it reproduces the reference's ``sample.<chain>.csv`` byte for byte (pinned by
tests/test_oracle_golden.py against fixtures made by the unmodified
reference, tests/golden/make_golden.py).

Reference (``/root/reference/posteriorSampling.py``) lines followed:
  burn / thin arithmetic ............ :1018-1027
  starting point search ............. :1060-1095, :1102-1105
  Nelder-Mead MLE start ............. :1107-1141
  group offsets ..................... :552-574
  partial-pooling start state ....... :725-758 (incl. stale log-prior quirk :284-288)
  sweep order / proposal ............ :594-613, :304-306
  objective call + group sums ....... :615-635
  Metropolis decision tree .......... :334-383
  step-size tuning .................. :385-437
  Gibbs hyper update ................ :463-502, :763-769
  iteration loop, retention, CSV .... :862-909, :640-654, :771-787

On top of the reference behaviour it can record a replay *tape* (the raw
standard normals / uniforms in consumption order plus per-decision
log-densities and accept bits) that the CUDA engine's replay mode is fed
with (SURVEY.md section 8c "Replay-tape recipe").
"""

import math
import os

import numpy
import scipy.optimize
import scipy.special
import scipy.stats

_NORM_PDF_LOGC = numpy.log(numpy.sqrt(2 * numpy.pi))


# --------------------------------------------------------------------------
# scipy.stats.norm(loc, scale).logpdf restated (scipy _distn_infrastructure
# rv_continuous.logpdf + _continuous_distns._norm_logpdf); used for the
# Gaussian group-level prior (posteriorSampling.py:500-502) and by the
# example objectives.  Checked bit-for-bit against scipy in the CPU tests.
# --------------------------------------------------------------------------
def norm_logpdf(x, loc, scale):
    x, loc, scale = numpy.broadcast_arrays(
        numpy.asarray(x, dtype=float), numpy.asarray(loc, dtype=float),
        numpy.asarray(scale, dtype=float))
    with numpy.errstate(all="ignore"):
        y = (x - loc) / scale
        out = -y ** 2 / 2.0 - _NORM_PDF_LOGC - numpy.log(scale)
        out = numpy.where(numpy.isfinite(y) | numpy.isnan(y), out, -numpy.inf)
        bad = ~(scale > 0) | numpy.isnan(y)
        out = numpy.where(bad, numpy.nan, out)
    return out


# --------------------------------------------------------------------------
# Objectives, written the way the reference's examples write them.
# Contract (posteriorSampling.py:61-102): f(list[P][N]) -> array[N].
# --------------------------------------------------------------------------
class GaussianDistributionObjective(object):
    """example/distribution.py:18-24 -- ll[i] = sum_j norm(mu[j][g(i)], sd[j]).logpdf(theta_j[i])."""

    name = "gaussian_distribution"

    def __init__(self, mu, sd, nResponsesPerGroup):
        self.mu = numpy.asarray(mu, dtype=float)           # [P][G]
        self.sd = numpy.asarray(sd, dtype=float)           # [P]
        self.groupIndex = numpy.repeat(numpy.arange(self.mu.shape[1]),
                                       nResponsesPerGroup)

    def __call__(self, parameter):
        gi = self.groupIndex
        ll = 0
        for j in range(self.mu.shape[0]):
            theta = numpy.asarray(parameter[j], dtype=float)
            # builtin sum() over the per-name terms starts from int 0 and
            # adds left to right (example/distribution.py:21-22)
            ll = ll + norm_logpdf(theta, self.mu[j][gi], self.sd[j])
        return ll


class LinearRegressionObjective(object):
    """example/regression.py:53-67 generalised to K coefficients + noise sd."""

    name = "linear_regression"

    def __init__(self, X, y):
        self.X = numpy.asarray(X, dtype=float)             # [N][K]
        self.y = numpy.asarray(y, dtype=float)             # [N]

    def __call__(self, parameter):
        K = self.X.shape[1]
        betaHat = numpy.vstack([parameter[k] for k in range(K)]).T
        yHat = numpy.sum(self.X * betaHat, axis=1)
        noise = numpy.array(parameter[K], dtype=float)
        # scipy.stats.norm(loc=y, scale=noise).logpdf(yHat)
        return norm_logpdf(yHat, self.y, noise)


class BernoulliLogitObjective(object):
    """ll_i = y_i*eta_i - log(1+exp(eta_i)), eta = a + b*x."""

    name = "bernoulli_logit"

    def __init__(self, x, y):
        self.x = numpy.asarray(x, dtype=float)
        self.y = numpy.asarray(y, dtype=float)

    def __call__(self, parameter):
        a = numpy.asarray(parameter[0], dtype=float)
        b = numpy.asarray(parameter[1], dtype=float)
        eta = a + b * self.x
        with numpy.errstate(all="ignore"):
            softplus = numpy.maximum(eta, 0.0) + numpy.log1p(numpy.exp(-numpy.abs(eta)))
        return self.y * eta - softplus


# --------------------------------------------------------------------------
def burn_thin(nIter, nSamples):
    """posteriorSampling.py:1018-1027."""
    if nIter < nSamples:
        raise Exception()
    elif nIter // 2 > nSamples:
        burn = nIter // 2
    else:
        burn = nIter - nSamples
    thin = int(numpy.ceil((nIter - burn) / nSamples))
    return burn, thin


def sequential_group_sums(ll, switch):
    """Strict left-to-right fp64 sum per group (posteriorSampling.py:631-633:
    builtin ``sum`` over numpy scalars adds sequentially starting from int 0)."""
    ll = numpy.asarray(ll, dtype=float)
    sizes = numpy.diff(switch)
    G = len(sizes)
    with numpy.errstate(all="ignore"):
        if G and numpy.all(sizes == sizes[0]):
            cols = ll.reshape(G, int(sizes[0]))
            acc = numpy.zeros(G)
            for j in range(cols.shape[1]):
                acc = acc + cols[:, j]
            return acc
        out = numpy.zeros(G)
        for g in range(G):
            acc = 0
            for v in ll[switch[g]:switch[g + 1]]:
                acc = acc + v
            out[g] = acc
        return out


class Tape(object):
    """Replay tape of one chain, in consumption order.

    Per iteration t, name p, group g:
      z_prop[t,p,g]   standard normal behind the proposal (:304-306)
      u_acc[t,p,g]    uniform of the accept test, NaN when none was drawn (:362)
      ll_prop, lp_prop, diff, accept[t,p,g]   what the oracle computed (:335-367)
    Per iteration t, name p (partial pooling only):
      z_mu[t,p]       standard normal behind the mu draw (:487)
      q_sig[t,p]      1/gammainccinv(a, U): the unit inverse-gamma draw (:498)
      mu[t,p], sigma2[t,p]   the resulting hyper-parameters
    """

    def __init__(self, nIter, P, G):
        shape = (nIter, P, G)
        self.z_prop = numpy.zeros(shape)
        self.u_acc = numpy.full(shape, numpy.nan)
        self.ll_prop = numpy.zeros(shape)
        self.lp_prop = numpy.zeros(shape)
        self.diff = numpy.zeros(shape)
        self.accept = numpy.zeros(shape, dtype=numpy.uint8)
        self.scale = numpy.zeros(shape)
        self.z_mu = numpy.zeros((nIter, P))
        self.q_sig = numpy.zeros((nIter, P))
        self.mu = numpy.zeros((nIter, P))
        self.sigma2 = numpy.zeros((nIter, P))


class OracleChain(object):
    """One chain of the reference sampler (MCMC + StepMethod + Sampler, restated)."""

    def __init__(self, chain, seed, nIter, nSamples, parameterName, nGroups,
                 nResponsesPerGroup, pooling, logLikelihoodFunction,
                 priorDistribution=None, startWithMLE=False,
                 startingPointValueRange=None, recordTape=False,
                 randomState=None):
        self.chain = chain
        # numpy.random.seed(seed) on the global legacy RandomState (:1015);
        # a private RandomState(seed) yields the identical stream.
        self.rs = randomState if randomState is not None \
            else numpy.random.RandomState(seed)
        self.nIter = nIter
        self.nSamples = nSamples
        self.burn, self.thin = burn_thin(nIter, nSamples)
        self.names = tuple(parameterName)
        self.P = len(self.names)
        if type(nResponsesPerGroup) == int:
            nResponsesPerGroup = [nResponsesPerGroup] * nGroups
        self.nResponses = int(sum(nResponsesPerGroup))
        if pooling not in ("partial", "none", "complete"):
            raise Exception("Invalid pooling: ", pooling)
        self.pooling = pooling
        self.f = logLikelihoodFunction
        self.prior = priorDistribution
        self.recordTape = recordTape
        self.tape = None

        self._findStartingPoint(startWithMLE, startingPointValueRange)

        # StepMethod construction (:517-582, :662-719)
        if pooling == "complete":
            self.G = 1
            nResponsesPerGroup = [self.nResponses]
        else:
            self.G = nGroups
        if pooling in ("none", "complete"):
            if priorDistribution is None or len(self.names) != len(priorDistribution):
                raise ValueError("Invalid prior")
        self.switch = numpy.hstack([0, numpy.cumsum(nResponsesPerGroup)]).astype(int)
        self.groupIndex = numpy.repeat(numpy.arange(self.G), nResponsesPerGroup)

        P, G = self.P, self.G
        self.value = numpy.zeros((P, G))
        self.logPrior = numpy.zeros((P, G))
        self.LL = numpy.full(G, numpy.nan)                  # :265
        self.scaleFactor = numpy.ones((P, G))               # :269
        self.nAccepted = numpy.zeros((P, G))
        self.nRejected = numpy.zeros((P, G))
        self.mu = numpy.zeros(P)
        self.sigma2 = numpy.zeros(P)
        self._setStartingPoint()

    # ------------------------------------------------------------------ start
    def _pooledNll(self, x):
        """:1102-1105"""
        param = [numpy.full(self.nResponses, float(p)) for p in x]
        with numpy.errstate(all="ignore"):
            return -1 * numpy.sum(self.f(param))

    def _findStartingPoint(self, startWithMLE, valueRange):
        if valueRange is None:
            valueRange = {}
        ll = numpy.inf
        x = [0] * self.P
        counter = 0
        while not numpy.isfinite(ll):
            for i, name in enumerate(self.names):
                if name in valueRange:
                    x[i] = self.rs.uniform(low=valueRange[name][0],
                                           high=valueRange[name][1])
                elif self.prior is not None:
                    x[i] = self.prior[i].rvs(random_state=self.rs)
                else:
                    # numpy.random.norm does not exist (:1079)
                    raise AttributeError("module 'numpy.random' has no attribute 'norm'")
            ll = self._pooledNll(x)
            counter += 1
            if counter > 1000:
                raise RuntimeError("Failed to find a valid starting state: ll =", ll)
        self.startingPoint = x
        if startWithMLE:
            self._optimizeStartingPoint()

    def _optimizeStartingPoint(self):
        """:1107-1141 (the unreachable re-draw at :1131 is not replicated)."""
        optimised = False
        n = 0
        while not optimised:
            n += 1
            res = scipy.optimize.minimize(self._pooledNll, self.startingPoint,
                                          method="Nelder-Mead",
                                          options={"maxiter": None, "maxfev": None,
                                                   "xtol": 0.0001, "ftol": 0.0001})
            if numpy.isfinite(res.fun):
                self.startingPoint = res.x
                if res.success:
                    optimised = True
            else:
                raise RuntimeError("non-finite MLE objective")
            if n > 10:
                self.startingPoint = res.x
                optimised = True

    def _fixedLogPrior(self, p, x):
        with numpy.errstate(all="ignore"):
            return numpy.asarray(self.prior[p].logpdf(x), dtype=float)

    def _hyperLogPrior(self, p, x):
        return norm_logpdf(x, self.mu[p], numpy.sqrt(self.sigma2[p]))

    def _logPriorOf(self, p, x):
        if self.pooling == "partial":
            return self._hyperLogPrior(p, x)
        return self._fixedLogPrior(p, x)

    def _setStartingPoint(self):
        P, G = self.P, self.G
        if self.pooling != "partial":
            # :584-592 -- every group starts at the same point; LL stays NaN
            for p in range(P):
                self.value[p, :] = self.startingPoint[p]
                self.logPrior[p, :] = self._fixedLogPrior(p, self.value[p])
            return
        # :725-744
        for p in range(P):
            val = self.startingPoint[p]
            self.mu[p] = val
            self.sigma2[p] = numpy.sqrt(numpy.abs(val) / 10.)   # sic: sqrt used as a variance
            sd = numpy.sqrt(self.sigma2[p])
            z = self.rs.standard_normal(G)
            self.value[p] = z * sd + self.mu[p]                 # scipy rvs: vals*scale+loc
            self.logPrior[p] = self._hyperLogPrior(p, self.value[p])
        # :746-758
        ll = numpy.full(G, -numpy.inf)
        while not numpy.all(numpy.isfinite(ll)):
            ll = self._groupLogLikelihood(None, None)
            fin = numpy.isfinite(ll)
            self.LL[fin] = ll[fin]
            bad = numpy.nonzero(~fin)[0]
            for p in range(P):
                if len(bad):
                    sd = numpy.sqrt(self.sigma2[p])
                    z = self.rs.standard_normal(len(bad))
                    # log-prior deliberately left stale (:284-288)
                    self.value[p, bad] = z * sd + self.mu[p]

    # ------------------------------------------------------------- likelihood
    def _pointwise(self, p, proposal):
        """:615-627"""
        rows = []
        for q in range(self.P):
            src = proposal if q == p else self.value[q]
            rows.append(src[self.groupIndex])
        with numpy.errstate(all="ignore"):
            ll = numpy.asarray(self.f(rows), dtype=float)
        assert len(ll) == self.nResponses
        return ll

    def _groupLogLikelihood(self, p, proposal):
        """:629-635"""
        return sequential_group_sums(self._pointwise(p, proposal), self.switch)

    # ------------------------------------------------------------------ sweep
    def _stepOneParameter(self, p, tune, t):
        G = self.G
        sd = 1. * self.scaleFactor[p]
        z = self.rs.standard_normal(G)
        proposal = self.value[p] + sd * z                       # numpy.random.normal(value, sd)
        llProp = self._groupLogLikelihood(p, proposal)

        with numpy.errstate(all="ignore"):
            lpProp = self._logPriorOf(p, proposal)
            postProp = lpProp + llProp
            postCur = self.logPrior[p] + self.LL                # :331-332 invariant
            diff = postProp - postCur
        b1 = (~numpy.isfinite(postCur)) & numpy.isfinite(postProp)   # :347-352
        b2 = (~b1) & (~numpy.isfinite(llProp))                       # :354-356
        b3 = (~b1) & (~b2) & (~numpy.isfinite(diff))                 # :358-360
        needU = ~(b1 | b2 | b3)
        u = self.rs.random_sample(int(needU.sum()))                  # :362, in group order
        accept = b1.copy()
        with numpy.errstate(all="ignore"):
            accept[needU] = numpy.log(u) < diff[needU]

        if self.tape is not None:
            tp = self.tape
            tp.z_prop[t, p] = z
            tp.u_acc[t, p, needU] = u
            tp.ll_prop[t, p] = llProp
            tp.lp_prop[t, p] = lpProp
            tp.diff[t, p] = diff
            tp.accept[t, p] = accept
            tp.scale[t, p] = sd

        # :369-383, :608-610
        self.value[p, accept] = proposal[accept]
        self.LL[accept] = llProp[accept]
        self.logPrior[p, accept] = lpProp[accept]
        self.nAccepted[p, accept] += 1.
        self.nRejected[p, ~accept] += 1.

        if tune:
            self._tune(p)

    def _tune(self, p):
        """:385-437, vectorised over groups."""
        acc, rej = self.nAccepted[p], self.nRejected[p]
        tot = acc + rej
        active = tot != 0
        with numpy.errstate(all="ignore"):
            rate = acc / tot
        cur = self.scaleFactor[p].copy()
        f = numpy.ones_like(cur)
        f = numpy.where(rate < 0.001, 0.1,
            numpy.where(rate < 0.05, 0.5,
            numpy.where(rate < 0.2, 0.9,
            numpy.where(rate > 0.95, 10.0,
            numpy.where(rate > 0.75, 2.0,
            numpy.where(rate > 0.5, 1.1, 1.0))))))
        new = cur * f
        new = numpy.where(new == 0, cur, new)                   # :428-430
        self.scaleFactor[p] = numpy.where(active, new, cur)
        self.nAccepted[p] = numpy.where(active, 0., acc)
        self.nRejected[p] = numpy.where(active, 0., rej)

    def _stepHyperParameter(self, p, t):
        """:763-769 -> :463-498 -> setPrior :273-282"""
        if self.pooling != "partial":
            return
        x = numpy.array(list(self.value[p]))
        n = len(x)
        muHat = numpy.mean(x)
        sd = numpy.sqrt(self.sigma2[p] / n)
        zmu = self.rs.standard_normal()
        self.mu[p] = muHat + sd * zmu                           # numpy.random.normal(muHat, sd)
        with numpy.errstate(all="ignore"):
            hat = numpy.sum((x - self.mu[p]) ** 2) / (n - 1)
            a = (n - 1) / 2.
            U = self.rs.random_sample()
            q = 1.0 / scipy.special.gammainccinv(a, U)          # scipy invgamma._ppf
            self.sigma2[p] = q * (a * hat) + 0.0                # rvs: vals*scale+loc
        self.logPrior[p] = self._hyperLogPrior(p, self.value[p])
        if self.tape is not None:
            self.tape.z_mu[t, p] = zmu
            self.tape.q_sig[t, p] = q
            self.tape.mu[t, p] = self.mu[p]
            self.tape.sigma2[t, p] = self.sigma2[p]

    # ------------------------------------------------------------- output
    @property
    def header(self):
        """:640-646, :771-778, :504-507, :263"""
        cols = []
        for p, name in enumerate(self.names):
            if self.pooling == "partial":
                cols += ["%s_mu" % name, "%s_sigma2" % name]
            cols += ["%s[%.3i]" % (name, j) for j in range(self.G)]
        return cols

    @property
    def values(self):
        """:648-654, :780-787"""
        out = []
        for p in range(self.P):
            if self.pooling == "partial":
                out += [self.mu[p], self.sigma2[p]]
            out += list(self.value[p])
        return out

    def run(self, sampleFile=None, llFile=None, saveLogLikelihood=False,
            keepRows=True):
        """Sampler._loop :862-896.  Returns (iteration index, values) rows."""
        if self.recordTape:
            self.tape = Tape(self.nIter, self.P, self.G)
        rows = []
        sampleLines, llLines = [], []
        for i in range(self.nIter):
            tune = bool(i and (i < self.burn) and (i % 100 == 0))
            for p in range(self.P):                             # :594-597
                self._stepOneParameter(p, tune, i)
                self._stepHyperParameter(p, i)
            if i == self.burn:
                sampleLines.append("index,chain," + ",".join(self.header))
            if i % self.thin == 0 and i >= self.burn:
                vals = self.values
                if keepRows:
                    rows.append((i, numpy.array(vals, dtype=float)))
                sampleLines.append("%i,%i," % (i, self.chain) +
                                   ",".join(["%f" % v for v in vals]))
                if saveLogLikelihood:
                    llLines.append(",".join(
                        ["%f" % v for v in self._pointwise(None, None)]))
        if sampleFile is not None:
            with open(sampleFile, "w") as h:
                h.write("\n".join(sampleLines) + "\n")
        if llFile is not None and saveLogLikelihood:
            with open(llFile, "w") as h:
                h.write("\n".join(llLines) + "\n")
        return rows


def samplePosteriorOracle(nChains, nIter, nSamples, parameterName, nGroups,
                          nResponsesPerGroup, pooling, logLikelihoodFunction,
                          outputDirectory, saveLogLikelihood=True,
                          priorDistribution=None, startWithMLE=False,
                          startingPointValueRange=None, recordTape=False):
    """Serial restatement of samplePosterior (:28-216); seed = chain (:225)."""
    sampleDirectory = os.path.join(outputDirectory, "sample")
    os.makedirs(sampleDirectory, exist_ok=True)
    chains = []
    for chain in range(nChains):
        oc = OracleChain(chain, chain, nIter, nSamples, parameterName, nGroups,
                         nResponsesPerGroup, pooling, logLikelihoodFunction,
                         priorDistribution, startWithMLE,
                         startingPointValueRange, recordTape)
        oc.run(os.path.join(sampleDirectory, "sample.%i.csv" % chain),
               os.path.join(sampleDirectory, "logLikelihood.%i.csv" % chain),
               saveLogLikelihood)
        chains.append(oc)
    return chains
