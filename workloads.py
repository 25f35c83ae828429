"""Synthetic workloads of BASELINE.json's configs 3-5, shared by bench.py (GPU arm
and oracle-port CPU arm) and baseline/run_reference.py (the unmodified reference).  numpy only."""

import numpy


def makeWorkload(G, R, K, seed=20261018):
    """Config C3: X[:,0]=1, X[:,1:]~N(0,1) fp32; beta_gk ~ N(k-3.5, 1); y = X.beta + N(0,1);
    parameters (b0..b{K-1}, sigma); ranges b_k in [-5,5], sigma in [0.5,2]."""
    rs = numpy.random.RandomState(seed)
    N = G * R
    X = numpy.ones((N, K))
    X[:, 1:] = rs.normal(size=(N, K - 1)).astype(numpy.float32)
    beta = rs.normal(numpy.arange(K) - 3.5, 1.0, size=(G, K))
    gi = numpy.repeat(numpy.arange(G), R)
    y = numpy.sum(X * beta[gi], axis=1) + rs.normal(size=N)
    names = tuple("b%d" % k for k in range(K)) + ("sigma",)
    ranges = dict((n, [-5, 5]) for n in names[:-1])
    ranges["sigma"] = [0.5, 2]
    return X, y, names, ranges


def makeLogitWorkload(G, R, seed=20261019):
    """Config C5: x ~ N(0,1); a_g ~ N(0,1), b_g ~ N(1,0.5); y ~ Bernoulli(sigmoid(a_g + b_g x));
    parameters (a, b)."""
    rs = numpy.random.RandomState(seed)
    N = G * R
    x = rs.normal(size=N)
    a = rs.normal(0, 1, size=G)
    b = rs.normal(1, 0.5, size=G)
    gi = numpy.repeat(numpy.arange(G), R)
    eta = a[gi] + b[gi] * x
    y = (rs.random_sample(N) < 1 / (1 + numpy.exp(-eta))).astype(float)
    return x, y, ("a", "b"), {"a": [-2, 2], "b": [-1, 3]}
