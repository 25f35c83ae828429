#!/usr/bin/env python
"""Diagnostics at config-3 size, for timing and for ncu (GPU box):

    python tools/diag_profile.py kernels [keys]     R-hat / ESS / median / HDI of `keys` columns (default 128) of a
                                                    device-resident store of 1,024 chains x 1,000 rows (one slab):
                                                    per-kernel CUDA-event times; the command ncu wraps
    python tools/diag_profile.py files [keys]       the same columns through a .npy store in /dev/shm and
                                                    sampleDiagnosis.Diagnostic: where the wall time goes
"""
import os
import sys
import time

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mcmc-for-nested-data_b200")):
    sys.path.insert(0, p)
import sampleDiagnosis as sd  # noqa: E402


def ar1(rows, ncol, chains, phi=0.9, seed=1):
    """AR(1) draws [rows][ncol][chains] float32 on the device: autocorrelated like thinned MCMC output."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.empty((rows, ncol, chains), dtype=torch.float32, device="cuda")
    cur = torch.randn((ncol, chains), generator=g, device="cuda")
    for r in range(rows):
        cur = phi * cur + (1 - phi * phi) ** 0.5 * torch.randn((ncol, chains), generator=g, device="cuda")
        x[r] = cur
    return x


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "kernels"
    keys = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    rows, chains = 1000, 1024
    store = ar1(rows, keys, chains)
    torch.cuda.synchronize()
    if mode == "kernels":
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for rep in range(2):
            ev[0].record()
            rhat, ess = sd.convergenceFromStore(store, rows, chains)
            ev[1].record()
            stats = sd.orderStatisticsFromStore(store, rows, chains)
            ev[2].record()
            torch.cuda.synchronize()
        print("keys %d, half-chains %d x %d draws: R-hat / ESS %.1f ms, median / HDI %.1f ms; min ESS %.0f, max R-hat %.4f"
              % (keys, 2 * chains, rows // 2, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]),
                 float(ess.min()), float(rhat.max())))
        pairs = keys * 2 * chains * (rows // 2) * (rows // 2 - 1) / 2
        print("variogram work: %.3g squared differences (2 FP64 instructions each)" % pairs)
        return
    d = "/dev/shm/mcmcn_diag_profile"
    os.makedirs(d, exist_ok=True)
    numpy.save(d + "/samples.npy", store.cpu().numpy())
    import json
    json.dump({"header": ["k%04d" % i for i in range(keys)], "iterations": list(range(rows)), "nChains": chains,
               "shards": [{"file": "samples.npy", "chains": [0, chains]}]}, open(d + "/manifest.json", "w"))
    t0 = time.perf_counter()
    diag = sd.Diagnostic(d + "/")
    r = diag.rhat
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("Diagnostic over the memory-mapped store, %d keys: %.2f s (%.1f ms per key)" % (keys, t1 - t0, 1e3 * (t1 - t0) / keys))
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    diag2 = sd.Diagnostic(d + "/")
    r = diag2.rhat
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
    import shutil
    shutil.rmtree(d)


if __name__ == "__main__":
    main()
