"""Time complete pooling at BASELINE config 3's data size (204,800 observations, 8 coefficients + sigma,
1,024 chains): split over observations (default) against the single-group step kernel
(MCMCN_NO_SPLIT=1).  usage: python tools/complete_pooling_timing.py [iterations]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mcmc-for-nested-data_b200")]
import scipy.stats  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from engine import Engine  # noqa: E402
from objectives import Objective  # noqa: E402

nIter = int(sys.argv[1]) if len(sys.argv) > 1 else 20
X, y, names, ranges = bench.makeWorkload(1024, 200, 8)
prior = [scipy.stats.norm(0, 10)] * 8 + [scipy.stats.gamma(2)]
for label, env, n in (("split over observations, tcgen05 evaluation", None, nIter),
                      ("split over observations, FP32-pipe evaluation", "MCMCN_NO_TC", nIter),
                      ("single group, one warp per 128 chains", "MCMCN_NO_SPLIT", 2)):
    os.environ.pop("MCMCN_NO_TC", None)
    if env:
        os.environ[env] = "1"
    eng = Engine(Objective.linear_regression(X, y, "fp32"), 1024, 200, "complete", 1024, priorDistribution=prior, seed=1)
    eng.initialise(names, ranges)
    eng.run(0, 2, 1000, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.run(2, n, 1000, 1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("%-40s %9.3f ms per iteration  (%.3g chain-observation evaluations/s)" % (label, dt * 1e3, 1024 * 9 * 204800 / dt))
