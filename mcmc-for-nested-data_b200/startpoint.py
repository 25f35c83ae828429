"""Starting points for all chains at once.

Restates MCMC._findStartingPoint (posteriorSampling.py:1060-1095) and
MCMC._optimizeStartingPoint (:1107-1141) with the pooled negative log-likelihood
(:1102-1105) evaluated on the device for every chain in one launch.  The Nelder-Mead
search follows scipy.optimize's algorithm (the reference calls
``scipy.optimize.minimize(method="Nelder-Mead")`` per chain; its ``xtol``/``ftol``
options are unknown to current scipy and ignored, so the defaults xatol = fatol = 1e-4
apply) but advances all chains in lock-step, masking the branch each chain takes.
"""

import numpy


def findStartingPoints(engine, rss, parameterName, valueRange, startWithMLE, logger=None):
    """Returns x[P][nChains].  ``rss[c]`` is chain c's legacy RandomState."""
    P, nC = engine.P, engine.nChains
    if valueRange is None:
        valueRange = {}
    for i, name in enumerate(parameterName):
        if name not in valueRange and engine.priorScipy is None:
            # the reference calls the non-existent numpy.random.norm here (:1079)
            raise ValueError("parameter %r needs a startingPointValueRange entry or a prior" % (name,))
    x = numpy.zeros((P, nC))
    pending = list(range(nC))
    counter = 0
    while pending:
        for c in pending:
            for i, name in enumerate(parameterName):
                if name in valueRange:
                    x[i, c] = rss[c].uniform(low=valueRange[name][0], high=valueRange[name][1])
                else:
                    x[i, c] = engine.priorScipy[i].rvs(random_state=rss[c])
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
