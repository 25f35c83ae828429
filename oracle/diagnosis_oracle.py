"""CPU oracle of the diagnostics hot path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Vectorised numpy restatement of ``/root/reference/sampleDiagnosis.py``
(``Diagnostic`` :88-427, ``Summary`` :430-491, ``computeHpdInterval`` :766-776).
Same formulas, same quirks (ESS sums rho from lag 0, truncation tests only
even lags, banker's rounding in the HDI gap, mode detection by substring),
but without the pure-Python O(keys * m * n^2) loops.  Pinned against outputs
of the unmodified reference in tests/test_oracle_golden.py.
"""

import glob

import numpy
import pandas


def computeHpdInterval(samples, hdi_p=95):
    """sampleDiagnosis.py:766-776."""
    prob = hdi_p / 100.
    s = numpy.sort(numpy.asarray(samples, dtype=float), kind="stable")
    n = len(s)
    gap = max(1, min(n - 1, round(n * prob)))        # Python banker's rounding
    width = s[gap:] - s[:n - gap]
    i = int(numpy.argmin(width))                     # first minimum
    return (s[i], s[i + gap])


def loadSampleDirectory(sampleDirectory):
    """sampleDiagnosis.py:102, :118-156 -- returns (keys in column order,
    dict key -> (m, n) array of split half-chains, nChains)."""
    files = glob.glob(sampleDirectory + "/sample*.csv")
    samples = {}
    keys = None
    m = len(files) * 2
    for i, filename in enumerate(files):
        d = pandas.read_csv(filename, engine="python")
        if i == 0:
            N = d.shape[0]
            n = N // 2
            keys = [k for k in d.columns if k not in ("chain", "index")]
            for k in keys:
                samples[k] = numpy.zeros((m, n))
        for k in keys:
            col = d[k].to_numpy(dtype=float)
            samples[k][2 * i, :] = col[0:n]
            samples[k][2 * i + 1, :] = col[n:N]       # odd N raises, as upstream
    return keys, samples


class DiagnosticOracle(object):
    def __init__(self, sampleDirectory=None, samples=None, keys=None):
        if samples is None:
            keys, samples = loadSampleDirectory(sampleDirectory)
        self.keys = list(keys if keys is not None else samples.keys())
        self._samples = samples
        first = samples[self.keys[0]]
        self._m, self._n = first.shape
        self.partiallyPooled = any("_" in k for k in self.keys)          # :143-144
        self.completelyPooled = not any("01]" in k for k in self.keys)   # :146-147
        self._hdiP = 95
        self._done = False

    def _compute(self):
        if self._done:
            return
        m, n = self._m, self._n
        self.B, self.W, self.vhat, self.rho = {}, {}, {}, {}
        self.rhat, self.effectiveN, self.median, self.hdi = {}, {}, {}, {}
        for k in self.keys:
            x = self._samples[k]
            B = n * numpy.var(numpy.mean(x, axis=1), ddof=1)             # :164-166
            W = numpy.mean(numpy.var(x, axis=1, ddof=1))                 # :174-176
            vhat = W * (n - 1) / n + B / n                               # :186-187
            rho = numpy.zeros(n)
            for t in range(n):                                           # :189-208
                d = x[:, t:] - x[:, :n - t]
                V = numpy.sum(d * d) / (m * (n - t))
                rho[t] = 1. - V / (2. * vhat)
            T = None                                                     # :241-251
            for t in range(0, n - 2, 2):
                if (rho[t + 1] + rho[t + 2]) < 0:
                    T = t
                    break
            if T is None:
                T = n - 1
            self.B[k], self.W[k], self.vhat[k], self.rho[k] = B, W, vhat, rho
            self.rhat[k] = numpy.sqrt(vhat / W)                          # :224
            self.effectiveN[k] = (m * n) / (1 + 2 * numpy.sum(rho[0:T + 1]))   # :253-255
            flat = x.flatten()
            self.median[k] = numpy.median(flat)                          # :426
            self.hdi[k] = computeHpdInterval(flat, self._hdiP)           # :427
        self._done = True

    @property
    def assessment(self):
        """:263-289"""
        self._compute()
        a = numpy.array([(k.encode(), self.rhat[k], self.rhat[k] < 1.1,
                          self.effectiveN[k], self.effectiveN[k] > self._m * 10,
                          self.median[k], self.hdi[k][0], self.hdi[k][1])
                         for k in self.keys],
                        dtype=[("parameter", "S40"), ("rhat", float),
                               ("converged", bool), ("effective n", float),
                               ("enough n", bool), ("median", float),
                               ("HDI lower", float), ("HDI upper", float)])
        return numpy.sort(a, order="parameter")

    @property
    def summary(self):
        """:297-329"""
        a = self.assessment
        names = sorted(set(r[0].decode("ascii").split("[")[0] for r in a if b"[" in r[0]))
        rh = dict((nm, []) for nm in names)
        cv = dict((nm, []) for nm in names)
        for r in a:
            if b"[" not in r[0]:
                continue
            nm = r[0].decode("ascii").split("[")[0]
            rh[nm].append(r[1])
            cv[nm].append(r[2])
        return numpy.array([(nm, min(rh[nm]), numpy.median(rh[nm]), max(rh[nm]),
                             numpy.mean(cv[nm])) for nm in names],
                           dtype=[("parameter", "S40"), ("rhat min", float),
                                  ("rhat median", float), ("rhat max", float),
                                  ("proportion converged", float)])

    def assessmentString(self, hyperOnly):
        """:381-394"""
        a = self.assessment
        out = ",".join(a.dtype.names) + "\n"
        for r in a:
            if hyperOnly and (b"_" not in r[0]):
                continue
            out += "'%s',%.3f,%s,%.3f,%s,%.3f,%.3f,%.3f\n" % (
                r[0].decode("ascii"), r[1], r[2], r[3], r[4], r[5], r[6], r[7])
        return out

    def summaryString(self):
        """:396-405"""
        s = self.summary
        out = ",".join(s.dtype.names) + "\n"
        for r in s:
            out += "'%s',%.3f,%.3f,%.3f,%.3f\n" % (
                r[0].decode("ascii"), r[1], r[2], r[3], r[4])
        return out


def summaryOracle(sampleDirectory):
    """Summary (:430-491): per retained row the mean / median over groups of
    each name, then mean, median and 95% HDI of those.  Returns the CSV text."""
    files = glob.glob(sampleDirectory + "/sample*.csv")
    mean, median = {}, {}
    names = None
    for i, filename in enumerate(files):
        d = pandas.read_csv(filename, engine="python")
        if i == 0:
            names = numpy.unique([c.split("[")[0] for c in d.columns if "[" in c])
            for nm in names:
                mean[nm], median[nm] = [], []
        for nm in names:
            cols = [c for c in d.columns if (nm + "[") in c]            # substring match :466-467
            x = d[cols].to_numpy(dtype=float)
            mean[nm] += list(numpy.mean(x, axis=1))
            median[nm] += list(numpy.median(x, axis=1))
    out = "stats,parameter,mean,median,HDI lower,HDI upper\n"
    for s, dd in zip(("groupMean", "groupMedian"), (mean, median)):
        for nm in sorted(dd):
            hdi = computeHpdInterval(dd[nm], 95.)
            out += "%s,%s,%.4f,%.4f,%.4f,%.4f\n" % (
                s, nm, numpy.mean(dd[nm]), numpy.median(dd[nm]), hdi[0], hdi[1])
    return out
