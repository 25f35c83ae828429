"""Starting points for all chains at once.

Restates MCMC._findStartingPoint (posteriorSampling.py:1060-1095) and
MCMC._optimizeStartingPoint (:1107-1141) with the pooled negative log-likelihood
(:1102-1105) evaluated on the device for every chain in one launch.  The Nelder-Mead
search follows scipy.optimize's algorithm (the reference calls
``scipy.optimize.minimize(method="Nelder-Mead")`` per chain; its ``xtol``/``ftol``
options are unknown to current scipy and ignored, so the defaults xatol = fatol = 1e-4
apply) but advances all chains in lock-step, masking the branch each chain takes.
"""

import ctypes
import os

import numpy

import mcmcn_native as nat


def _hostThreads():
    """Host threads for the chains' start-state streams: the cores of the box shared among its ranks."""
    if os.environ.get("MCMCN_START_THREADS"):
        return max(1, int(os.environ["MCMCN_START_THREADS"]))
    return max(1, min(32, (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))


class ChainStreams(object):
    """The chains' own legacy numpy streams (the reference seeds each chain's process with the chain index,
    posteriorSampling.py:225, :1015): stream c equals ``numpy.random.RandomState(seed0 + c)`` bit for bit,
    held and advanced natively (``mcmcn_streams_*``, csrc/mcmcn_streams.cu) by host threads instead of one
    Python object per chain."""

    def __init__(self, nChains, seed0=0):
        self.n = int(nChains)
        self.threads = _hostThreads()
        h = ctypes.c_void_p()
        nat.call("mcmcn_streams_create", self.n, int(seed0), self.threads, ctypes.byref(h))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            nat.load().mcmcn_streams_free(h)

    @staticmethod
    def _which(chains):
        return numpy.ascontiguousarray(chains, dtype=numpy.int64)

    def uniform(self, chains, low, high):
        """[len(chains)][len(low)]: for each listed chain, in order i, ``uniform(low[i], high[i])``."""
        which = self._which(chains)
        low, high = numpy.ascontiguousarray(low, dtype=numpy.float64), numpy.ascontiguousarray(high, dtype=numpy.float64)
        out = numpy.empty((len(which), len(low)))
        nat.call("mcmcn_streams_uniform", self._h, which.ctypes.data, len(which), len(low), low.ctypes.data,
                 high.ctypes.data, out.ctypes.data, self.threads)
        return out

    def standardNormal(self, chains, counts):
        """Flat array: counts[j] (or the one count) ``standard_normal`` draws of each listed chain, chain after chain."""
        which = self._which(chains)
        counts = numpy.broadcast_to(numpy.asarray(counts, dtype=numpy.int64), (len(which),))
        off = numpy.zeros(len(which) + 1, dtype=numpy.int64)
        off[1:] = numpy.cumsum(counts)
        out = numpy.empty(int(off[-1]))
        nat.call("mcmcn_streams_normal", self._h, which.ctypes.data, len(which), off.ctypes.data, out.ctypes.data,
                 self.threads)
        return out

    def randomState(self, c):
        """Chain c's stream as a numpy RandomState (for ``scipy_prior.rvs(random_state=...)``); give it back
        with ``adopt`` after drawing."""
        key = numpy.empty(624, dtype=numpy.uint32)
        pos, has, g = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
        nat.call("mcmcn_streams_get_state", self._h, int(c), key.ctypes.data, ctypes.byref(pos), ctypes.byref(has),
                 ctypes.byref(g))
        rs = numpy.random.RandomState(0)
        rs.set_state(("MT19937", key, pos.value, has.value, g.value))
        return rs

    def adopt(self, c, rs):
        _, key, pos, has, g = rs.get_state()
        key = numpy.ascontiguousarray(key, dtype=numpy.uint32)
        nat.call("mcmcn_streams_set_state", self._h, int(c), key.ctypes.data, int(pos), int(has), float(g))


def findStartingPoints(engine, streams, parameterName, valueRange, startWithMLE, logger=None):
    """Returns x[P][nChains].  ``streams`` holds every chain's legacy numpy stream (ChainStreams)."""
    P, nC = engine.P, engine.nChains
    if valueRange is None:
        valueRange = {}
    for i, name in enumerate(parameterName):
        if name not in valueRange and engine.priorScipy is None:
            # the reference calls the non-existent numpy.random.norm here (:1079)
            raise ValueError("parameter %r needs a startingPointValueRange entry or a prior" % (name,))
    allRanged = all(name in valueRange for name in parameterName)
    if allRanged:
        low = [valueRange[name][0] for name in parameterName]
        high = [valueRange[name][1] for name in parameterName]
    x = numpy.zeros((P, nC))
    pending = list(range(nC))
    counter = 0
    while pending:
        if allRanged:                                   # one native call: P uniforms of every pending chain (:1069-1077)
            x[:, pending] = streams.uniform(pending, low, high).T
        else:                                           # a prior's own sampler in between (:1079-1081): chain by chain
            for c in pending:
                rs = streams.randomState(c)
                for i, name in enumerate(parameterName):
                    if name in valueRange:
                        x[i, c] = rs.uniform(low=valueRange[name][0], high=valueRange[name][1])
                    else:
                        x[i, c] = engine.priorScipy[i].rvs(random_state=rs)
                streams.adopt(c, rs)
        nll = engine.pooledNll(x)
        pending = [c for c in pending if not numpy.isfinite(nll[c])]
        counter += 1
        if counter > 1000 and pending:
            raise RuntimeError("Failed to find a valid starting state: ll =", nll[pending[0]])
    if startWithMLE:
        x = optimiseStartingPoints(engine, x, logger)
    return x


def optimiseStartingPoints(engine, x0, logger=None):
    """:1107-1141 -- repeat Nelder-Mead from the last point until it reports success (<= 11 runs)."""
    x = x0.copy()
    todo = numpy.ones(x.shape[1], dtype=bool)
    for n in range(1, 12):
        xn, fn, ok = nelderMead(engine, x, todo)
        fin = numpy.isfinite(fn)
        if (todo & ~fin).any():
            raise RuntimeError("non-finite log-likelihood while optimising the starting state")
        x[:, todo] = xn[:, todo]
        todo = todo & ~ok
        if not todo.any():
            break
    return x


def nelderMead(engine, x0, active, xatol=1e-4, fatol=1e-4):
    """scipy.optimize._minimize_neldermead, batched over chains (columns of x0 [P][nC]).
    Returns (x [P][nC], f [nC], success [nC])."""
    N, nC = x0.shape
    rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
    maxiter = maxfun = N * 200

    def func(pts, who):
        """pts [P][nC]; evaluate all, count a call for chains in `who`."""
        f = engine.pooledNll(pts)
        fcalls[who] += 1
        return f

    fcalls = numpy.zeros(nC, dtype=int)
    sim = numpy.empty((N + 1, N, nC))
    sim[0] = x0
    for k in range(N):
        y = x0.copy()
        y[k] = numpy.where(y[k] != 0, (1 + 0.05) * y[k], 0.00025)
        sim[k + 1] = y
    fsim = numpy.empty((N + 1, nC))
    everyone = numpy.ones(nC, dtype=bool)
    for k in range(N + 1):
        fsim[k] = func(sim[k], everyone)

    def sort():
        ind = numpy.argsort(fsim, axis=0, kind="stable")
        fs = numpy.take_along_axis(fsim, ind, axis=0)
        sm = numpy.take_along_axis(sim, ind[:, None, :], axis=0)
        return sm, fs

    sim, fsim = sort()
    iterations = numpy.ones(nC, dtype=int)
    run = active.copy()
    while True:
        with numpy.errstate(all="ignore"):
            conv = (numpy.max(numpy.abs(sim[1:] - sim[0]), axis=(0, 1)) <= xatol) & \
                   (numpy.max(numpy.abs(fsim[0] - fsim[1:]), axis=0) <= fatol)
        run = run & ~conv & (fcalls < maxfun) & (iterations < maxiter)
        if not run.any():
            break
        xbar = numpy.add.reduce(sim[:-1], 0) / N
        xr = (1 + rho) * xbar - rho * sim[-1]
        fxr = func(xr, run)
        newx = sim[-1].copy()
        newf = fsim[-1].copy()
        doshrink = numpy.zeros(nC, dtype=bool)

        bExp = run & (fxr < fsim[0])
        if bExp.any():
            xe = (1 + rho * chi) * xbar - rho * chi * sim[-1]
            fxe = func(xe, bExp)
            useE = bExp & (fxe < fxr)
            useR = bExp & ~useE
            newx[:, useE], newf[useE] = xe[:, useE], fxe[useE]
            newx[:, useR], newf[useR] = xr[:, useR], fxr[useR]
        rest = run & ~bExp
        bRefl = rest & (fxr < fsim[-2])
        newx[:, bRefl], newf[bRefl] = xr[:, bRefl], fxr[bRefl]
        contr = rest & ~bRefl
        bOut = contr & (fxr < fsim[-1])
        if bOut.any():
            xc = (1 + psi * rho) * xbar - psi * rho * sim[-1]
            fxc = func(xc, bOut)
            ok = bOut & (fxc <= fxr)
            newx[:, ok], newf[ok] = xc[:, ok], fxc[ok]
            doshrink |= bOut & ~ok
        bIn = contr & ~bOut
        if bIn.any():
            xcc = (1 - psi) * xbar + psi * sim[-1]
            fxcc = func(xcc, bIn)
            ok = bIn & (fxcc < fsim[-1])
            newx[:, ok], newf[ok] = xcc[:, ok], fxcc[ok]
            doshrink |= bIn & ~ok
        sim[-1], fsim[-1] = newx, newf
        if doshrink.any():
            for j in range(1, N + 1):
                sj = sim[0] + sigma * (sim[j] - sim[0])
                fj = func(sj, doshrink)
                sim[j][:, doshrink] = sj[:, doshrink]
                fsim[j][doshrink] = fj[doshrink]
        iterations[run] += 1
        keepSim, keepF = sim.copy(), fsim.copy()
        sim, fsim = sort()
        # chains that were not running keep their simplex untouched
        sim[:, :, ~run], fsim[:, ~run] = keepSim[:, :, ~run], keepF[:, ~run]
    success = (fcalls < maxfun) & (iterations < maxiter)
    return sim[0], numpy.min(fsim, axis=0), success
