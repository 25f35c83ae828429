#!/usr/bin/env python
"""Time the UNMODIFIED reference (baseline/_ref/posteriorSampling.py, staged by __graft_entry__.build())
through its own public API, samplePosterior, on a synthetic workload of bench.py's shape.

    python baseline/run_reference.py --groups 64 --obs 200 --coef 8 --chains 16 --processes 16 --iters 6

The objective is a numpy callable in the style of the reference's examples (example/regression.py:53-67:
one vectorised scipy.stats call over all observations); chains run as the reference runs them, one OS
process per chain in batches of --processes (posteriorSampling.py:173-201).  Prints one JSON line:
wall seconds of the whole call, and the seconds of the iteration loop alone, read from the "Sampling
started" / "100% complete" lines the reference itself logs per chain (posteriorSampling.py:862-896).
Runs in its own process with baseline/_ref first on sys.path: the product package has modules of the
same names on purpose, and must not be importable here."""

import argparse
import datetime
import functools
import json
import os
import re
import shutil
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
sys.path.insert(0, os.path.dirname(HERE))      # workloads.py
sys.path.insert(0, REF)

import numpy  # noqa: E402
import scipy.stats  # noqa: E402


def regressionLogLikelihood(parameter, X, y, K):
    yHat = numpy.sum(X * numpy.vstack(parameter[:K]).T, axis=1)
    return scipy.stats.norm(loc=y, scale=numpy.array(parameter[K])).logpdf(yHat)


def logitLogLikelihood(parameter, x, y):
    eta = numpy.array(parameter[0]) + numpy.array(parameter[1]) * x
    return y * eta - numpy.logaddexp(0.0, eta)


def loopSeconds(logDirectory, nChains):
    """Per chain: seconds between the reference's own 'Sampling started' and '100% complete' log lines."""
    stamp = re.compile(r"^(\d{4}-\d\d-\d\d \d\d:\d\d:\d\d),(\d{3}) - ")
    out = []
    for c in range(nChains):
        t0 = t1 = None
        last = None
        with open(os.path.join(logDirectory, "mcmc.chain%.2i.log" % c)) as h:
            for line in h:
                m = stamp.match(line)
                if m:
                    last = datetime.datetime.strptime(m.group(1), "%Y-%m-%d %H:%M:%S").timestamp() + int(m.group(2)) / 1e3
                elif "Sampling started" in line:
                    t0 = last
                elif "100% complete" in line:
                    t1 = last
        out.append(None if t0 is None or t1 is None else t1 - t0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=64)
    ap.add_argument("--obs", type=int, default=200)
    ap.add_argument("--coef", type=int, default=8, help="0 = Bernoulli-logit (config 5)")
    ap.add_argument("--pooling", default="partial")
    ap.add_argument("--chains", type=int, default=1)
    ap.add_argument("--processes", type=int, default=1)
    ap.add_argument("--iters", type=int, default=6, help=">= 6 (posteriorSampling.py:868 takes i % round(nIter / 10.): nIter <= 5 divides by zero)")
    ap.add_argument("--full-groups", type=int, default=0, help="groups of the full config the sample is a slice of")
    args = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "posteriorSampling.py")):
        print(json.dumps({"unavailable": "baseline/_ref is not staged (python __graft_entry__.py in the build container)"}))
        return
    import posteriorSampling as reference
    assert os.path.dirname(os.path.abspath(reference.__file__)) == REF, reference.__file__
    import workloads
    full = args.full_groups or args.groups
    prior = None
    if args.coef:
        X, y, names, ranges = workloads.makeWorkload(full, args.obs, args.coef)
        n = args.groups * args.obs
        objective = functools.partial(regressionLogLikelihood, X=X[:n], y=y[:n], K=args.coef)
        if args.pooling != "partial":
            prior = [scipy.stats.norm(0, 10)] * args.coef + [scipy.stats.gamma(2)]
    else:
        x, y, names, ranges = workloads.makeLogitWorkload(full, args.obs)
        n = args.groups * args.obs
        objective = functools.partial(logitLogLikelihood, x=x[:n], y=y[:n])
        if args.pooling != "partial":
            prior = [scipy.stats.norm(0, 5), scipy.stats.norm(0, 5)]
    out = tempfile.mkdtemp(prefix="mcmcn_ref_")
    try:
        t0 = time.perf_counter()
        reference.samplePosterior(args.chains, args.iters, 2, names, args.groups, args.obs, args.pooling, objective,
                                  out, saveLogLikelihood=False, priorDistribution=prior,
                                  startingPointValueRange=ranges, nProcesses=args.processes, displayProgress=False)
        wall = time.perf_counter() - t0
        loops = loopSeconds(os.path.join(out, "log"), args.chains)
    finally:
        shutil.rmtree(out, ignore_errors=True)
    print(json.dumps({"wall_s": wall, "loop_s": loops, "chains": args.chains, "processes": args.processes,
                      "iters": args.iters, "groups": args.groups, "obs": args.obs, "coef": args.coef,
                      "pooling": args.pooling, "numpy": numpy.__version__,
                      "scipy": __import__("scipy").__version__, "cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
