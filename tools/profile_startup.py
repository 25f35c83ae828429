#!/usr/bin/env python
"""Where samplePosterior's start-up goes at config-3 size (GPU box): cProfile of a short call
(1,024 chains x 200 iterations, binary store in /dev/shm) after two warm calls.
usage: python tools/profile_startup.py [chains [iterations [rows per chain]]]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mcmc-for-nested-data_b200")]
import torch  # noqa: E402

torch.zeros(1).cuda()
import posteriorSampling as ps  # noqa: E402
from objectives import Objective  # noqa: E402
from workloads import makeWorkload  # noqa: E402

CHAINS = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ROWS = int(sys.argv[3]) if len(sys.argv) > 3 else 20
X, y, names, ranges = makeWorkload(1024, 200, 8)
ps.CSV_VALUE_LIMIT = 0
ps.STORE_DTYPE = "float32"


def run():
    handle = Objective.linear_regression(X, y)
    ps.samplePosterior(CHAINS, ITERS, ROWS, names, 1024, 200, "partial", handle, "/dev/shm/mcmcn_profile_startup",
                       saveLogLikelihood=False, startingPointValueRange=ranges, displayProgress=False)


run()
t = time.time()
run()
print("second call wall", time.time() - t)
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
