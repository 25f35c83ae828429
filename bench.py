#!/usr/bin/env python
"""bench.py -- throughput of the MCMC step path on BASELINE.json's config 3
("synthetic hierarchical linear regression, partial pooling: 1,024 groups x 200 obs x 8
coefficients, 1,024 chains ... on 1 B200"), per GPU; weak scaling over --gpus.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (torchrun launches N > 1)
    python bench.py --impl reference ...                            CPU arm (oracle port, all cores)

One step = ITERS_PER_STEP sampler iterations (posteriorSampling.py:862-896: P sweeps, P Gibbs
hyper-updates, tuning while burning in, retained-sample write-back after) of every chain.
metric  chain-iterations/s summed over all GPUs, inputs resident in HBM (`value`) and
through the host-buffer API with H2D of the observation data and D2H of the step's
retained samples inside the timed region (`e2e`).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "mcmc-for-nested-data_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "chain-iterations/sec"
UNIT = "chain-iterations/s"
FLOP_PER_EVAL = 20.0      # 2K+4 at K=8 (DESIGN.md "algorithmic work")


def makeWorkload(G, R, K, seed=20261018):
    """X[:,0]=1, X[:,1:]~N(0,1) fp32; beta_gk ~ N(k-3.5, 1);
    y = X.beta + N(0,1); parameters (b0..b{K-1}, sigma); ranges b_k in [-5,5], sigma in [0.5,2]."""
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
    """x ~ N(0,1); a_g ~ N(0,1), b_g ~ N(1,0.5);
    y ~ Bernoulli(sigmoid(a_g + b_g x)); parameters (a, b)."""
    rs = numpy.random.RandomState(seed)
    N = G * R
    x = rs.normal(size=N)
    a = rs.normal(0, 1, size=G)
    b = rs.normal(1, 0.5, size=G)
    gi = numpy.repeat(numpy.arange(G), R)
    eta = a[gi] + b[gi] * x
    y = (rs.random_sample(N) < 1 / (1 + numpy.exp(-eta))).astype(float)
    return x, y, ("a", "b"), {"a": [-2, 2], "b": [-1, 3]}


def fixedPriors(args):
    """Priors of the no-pooling variant (C5 uses N(0, 5) x 2; the regression
    shape N(0, 10) on the coefficients and Gamma(2) on sigma); None for partial pooling."""
    if args.pooling == "partial":
        return None
    import scipy.stats
    if args.coef == 0:
        return [scipy.stats.norm(0, 5), scipy.stats.norm(0, 5)]
    return [scipy.stats.norm(0, 10)] * args.coef + [scipy.stats.gamma(2)]


def schedule(args):
    """A (warmup+steps)-step slice of C3's schedule: burn = half the iterations, thin 10."""
    total = (args.warmup + args.steps) * args.iters_per_step
    burn = total // 2
    return total, burn, args.thin


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(numpy.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm
def _cpuChainWorker(job):
    """One chain of the oracle port for `nIter` iterations at the bench shape; returns seconds
    spent in the iteration loop (start-up excluded, like the GPU arm)."""
    chain, nIter, G, R, K, pooling, prior = job
    from oracle import posterior_oracle as po
    if K == 0:                                   # C5: Bernoulli-logit
        x, y, names, ranges = makeLogitWorkload(G, R)
        obj = po.BernoulliLogitObjective(x, y)
    else:
        X, y, names, ranges = makeWorkload(G, R, K)
        obj = po.LinearRegressionObjective(X, y)
    oc = po.OracleChain(chain, chain, max(nIter, 10), max(nIter, 10) // 2, names, G, R, pooling,
                        obj, prior, False, ranges)
    oc.nIter = nIter
    t0 = time.perf_counter()
    oc.run(keepRows=False)
    return time.perf_counter() - t0


def cpuBaseline(args, cores, iters):
    """chain-iterations/s of the oracle port (the reference's algorithm restated in numpy) on
    `cores` host processes, `iters` iterations of one chain each."""
    import multiprocessing
    jobs = [(c, iters, args.groups, args.obs, args.coef, args.pooling, fixedPriors(args)) for c in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        dt = _cpuChainWorker(jobs[0])
    else:
        with multiprocessing.get_context("fork").Pool(cores) as pool:
            t0 = time.perf_counter()
            pool.map(_cpuChainWorker, jobs)
            dt = time.perf_counter() - t0
    return cores * iters / dt, dt


def runReference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    iters = args.ref_iters
    for _ in range(args.warmup):
        cpuBaseline(args, cores, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpuBaseline(args, cores, iters)     # includes process start-up and data generation per step
    dt = time.perf_counter() - t0
    value = cores * iters * args.steps / dt
    P, N = (args.coef + 1 if args.coef else 2), args.groups * args.obs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workloadConfig(args, cores),
            "evals_per_sec": value * P * N,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d chains (one per host core) x %d iterations per step of the same "
                                       "workload, numpy restatement of the reference (oracle/)" % (cores, iters)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(line))


def workloadConfig(args, chains):
    pool = {"partial": "partial pooling", "none": "no pooling (fixed priors)"}[args.pooling]
    name = ("C3: hierarchical linear regression, %s, %d groups x %d obs x %d coefficients (+sigma)"
            % (pool, args.groups, args.obs, args.coef)) if args.coef else \
           ("C5: hierarchical Bernoulli-logit, %s, %d groups x %d trials, 2 parameters" % (pool, args.groups, args.obs))
    return {"workload": name,
            "chains_per_gpu": chains, "iters_per_step": args.iters_per_step,
            "schedule": "burn = first half of the run (tune every 100), thin %d after" % args.thin,
            "l2": "chain state + sample write-back exceed L2 (%.0f MB touched per iteration)"
                  % (chains * args.groups * ((args.coef + 1) if args.coef else 2) * 28 / 1e6)}


def roofline(args, tensorCore, sweepMs, chains, peakFp32, peakTf32, peakMufu):
    """Roofline of the step kernel.  Algorithmic work per chain-observation
    log-density evaluation: 2K+4 = 20 FP32 flop.
    FP32-pipe kernel: bound by the FP32 pipe, achieved = 20 flop/eval over the launch time.
    tcgen05 kernel: the contraction runs on the tensor pipe as 3xTF32 -- per 128-chain x group tile
    and sweep, 4 MMAs (A_hi.X_hi, A_lo.X_hi, A_hi.X_lo, 1.NE) of 2*128*Np*8 flop, Np = the group's
    observations rounded up to 16 -- and that pipe is the one that binds once latencies are hidden,
    so achieved = those TF32 flop over the launch time against the pipe's measured MMA rate."""
    G, R, K = args.groups, args.obs, args.coef
    if K == 0:
        # C5.  Algorithmic figure: 2 MUFU (ex2, lg2) + 6 FP32 flop per evaluation,
        # the MUFU pipe binds -> `achieved` / `frac`.  The kernel itself executes one ex2 per evaluation
        # and one lg2 per 16 (logarithm of the product of 16 factors), 1.0625 MUFU per evaluation:
        # `mufu_pipe_utilisation` is that executed rate over the same peak.
        evals = 2.0 * G * R * chains
        ops = 2.0 * evals / (sweepMs * 1e-3)
        executed = 1.0625 * evals / (sweepMs * 1e-3)
        return {"bound": "mufu", "kernel": "sweep_kernel<Logit,4,float>", "achieved": ops / 1e9, "peak": peakMufu / 1e9,
                "unit": "Gop/s", "frac": ops / peakMufu, "mufu_per_eval": 2.0, "mufu_executed_per_eval": 1.0625,
                "mufu_pipe_utilisation": executed / peakMufu, "traffic": 2.499e9,
                "fp32_pipe_peak_tflops": peakFp32 / 1e12,
                "peak_source": "MUFU pipe limit measured in this run by an ex2-only microbenchmark (mcmcn_peak_mufu); "
                               "nominal 148 SM x 16 lanes x 1.965 GHz = 4653; traffic from profiles/r1_c5_kernel_ncu_summary.txt"}
    P, N = K + 1, G * R
    algFlops = FLOP_PER_EVAL * P * N * chains
    algTflops = algFlops / (sweepMs * 1e-3) / 1e12
    common = {"flop_per_eval": FLOP_PER_EVAL, "algorithmic_fp32_tflops": algTflops,
              "fp32_pipe_peak_tflops": peakFp32 / 1e12, "mufu_peak_gops": peakMufu / 1e9,
              "traffic": args.traffic if tensorCore else args.traffic_pipe}
    if not tensorCore:
        common.update({"bound": "fp32", "kernel": "sweep_kernel<LinReg<8>,4,float>", "achieved": algTflops,
                       "peak": peakFp32 / 1e12, "unit": "TFLOP/s", "frac": algFlops / (sweepMs * 1e-3) / peakFp32,
                       "peak_source": "FP32 pipe limit measured in this run by an FFMA-only microbenchmark "
                                      "(mcmcn_peak_fp32); MEASURED_PEAKS.json has no FP32 figure; nominal 74.4"})
        return common
    npad = max(16, (R + 15) // 16 * 16)
    tiles = ((chains + 127) // 128) * G * P
    tf32Flops = tiles * 4 * 2.0 * 128 * npad * 8
    common.update({"bound": "tensor", "kernel": "sweep_tc_kernel (tcgen05.mma kind::tf32, 3xTF32 + ne)",
                   "achieved": tf32Flops / (sweepMs * 1e-3) / 1e12, "peak": peakTf32 / 1e12, "unit": "TFLOP/s",
                   "frac": tf32Flops / (sweepMs * 1e-3) / peakTf32,
                   "tf32_flop_per_eval": tf32Flops / (P * N * chains),
                   "peak_source": "tensor pipe limit for this MMA shape (M128 N208 K8 kind::tf32, A in TMEM) measured "
                                  "in this run by an MMA-only microbenchmark (mcmcn_peak_tf32); MEASURED_PEAKS.json "
                                  "has bf16 only (1658 TFLOP/s burst; TF32 runs at half the bf16 rate = 829)"})
    return common


# ----------------------------------------------------------------------------- GPU arm
def runGpu(args):
    import torch
    import torch.distributed as dist
    from engine import Engine, SampleStore
    from objectives import Objective
    import mcmcn_native as nat
    import ctypes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    G, R, K = args.groups, args.obs, args.coef
    N = G * R
    chains = args.chains_per_gpu
    if K == 0:
        x, y, names, ranges = makeLogitWorkload(G, R)
        obj = Objective.bernoulli_logit(x, y, args.precision)
    else:
        X, y, names, ranges = makeWorkload(G, R, K)
        obj = Objective.linear_regression(X, y, args.precision)
    P = len(names)
    eng = Engine(obj, G, R, args.pooling, chains, priorDistribution=fixedPriors(args), chainId0=rank * chains, seed=args.seed)
    eng.initialise(names, ranges)
    total, burn, thin = schedule(args)
    ips = args.iters_per_step
    nRows = len([i for i in range(total) if i % thin == 0 and i >= burn])
    store = SampleStore(eng, max(nRows, 1), torch.float32)

    lib = nat.load()
    peak = ctypes.c_double(0.0)
    nat.check(lib.mcmcn_peak_fp32(ctypes.byref(peak), eng.stream))
    peakFlops = peak.value
    mufu = ctypes.c_double(0.0)
    nat.check(lib.mcmcn_peak_mufu(ctypes.byref(mufu), eng.stream))
    tensorCore = eng.usesTensorCore
    peakTf32 = ctypes.c_double(0.0)
    if tensorCore:
        nat.check(lib.mcmcn_peak_tf32(ctypes.byref(peakTf32), eng.stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pass 1: inputs resident in HBM
    for w in range(args.warmup):
        eng.run(w * ips, ips, burn, thin, store=store)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    timing = numpy.zeros(12)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        eng.run((args.warmup + k) * ips, ips, burn, thin, store=store, timing=timing)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    eng.collectTiming()            # durations of the launches sampled (every 8th iteration) inside the timed region

    # ---- pass 2: same steps through host buffers (H2D of the observation data, D2H of the
    # rows the step retained and of the hyper-parameters), fresh Philox seed
    stepInput = eng.stepInput           # the observation blocks the step kernel reads
    pinData = torch.from_numpy(stepInput.cpu().numpy()).pin_memory()
    rowBytes = eng.nCol * eng.S * 4
    maxRows = ips // thin + 1
    pinRows = torch.empty((maxRows, eng.nCol, eng.S), dtype=torch.float32).pin_memory()
    hyper = eng.hyper if eng.hyper is not None else torch.zeros((1, 1, eng.S), dtype=torch.float64, device=dev)
    pinHyper = torch.empty(tuple(hyper.shape), dtype=torch.float64).pin_memory()
    eng.seed = args.seed + 1
    store.iterations = store.iterations[:len([i for i in range(args.warmup * ips) if i % thin == 0 and i >= burn])]
    h2d = pinData.numel() * pinData.element_size()
    d2h = 0
    barrier()
    # the device->host reads run on a second stream, so the rows of step k travel while step k + 1
    # computes; every copy is finished before the closing event (the main stream waits for the copy stream)
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)
    hyperSnap = torch.empty_like(hyper)          # the step's hyper-parameters, frozen before the next step overwrites them
    snapFree = torch.cuda.Event()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        stepInput.copy_(pinData, non_blocking=True)
        r0 = len(store.iterations)
        eng.run((args.warmup + k) * ips, ips, burn, thin, store=store)
        r1 = len(store.iterations)
        if k:
            main.wait_event(snapFree)
        hyperSnap.copy_(hyper, non_blocking=True)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if r1 > r0:                              # retained rows are append-only: nothing overwrites them
                pinRows[:r1 - r0].copy_(store.tensor[r0:r1], non_blocking=True)
            pinHyper.copy_(hyperSnap, non_blocking=True)
            snapFree.record(side)
        d2h += (r1 - r0) * rowBytes + pinHyper.numel() * 8
    main.wait_stream(side)
    e1.record()
    barrier()
    msE2e = e0.elapsed_time(e1)

    # ---- between-chain diagnostics of the rows the e2e pass retained, on the device; with N > 1 the
    # half-chain moments and per-lag sums cross NVLink in one NCCL all-gather (the path's only exchange)
    from sampleDiagnosis import convergenceFromStore
    nKept = len(store.iterations)
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    diag = None
    if nKept >= 6:                                   # two half-chains of at least 3 rows each (mcmcn_diag_ess)
        barrier()
        d0.record()
        rhat, ess = convergenceFromStore(store.tensor, nKept, chains, group=dist.group.WORLD if world > 1 else None)
        d1.record()
        barrier()
        fin = torch.isfinite(ess)
        diag = {"rows": nKept, "half_chains": 2 * chains * world, "keys": int(ess.numel()),
                "min_ess": float(ess[fin].min()) if bool(fin.any()) else None,
                "max_rhat": float(rhat[torch.isfinite(rhat)].max()),
                "diag_ms": d0.elapsed_time(d1)}

    t = torch.tensor([ms, msE2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, msE2e = float(t[0]), float(t[1])

    if rank == 0:
        chainIters = world * chains * ips * args.steps
        value = chainIters / (ms * 1e-3)
        e2e = chainIters / (msE2e * 1e-3)
        sweepMs = timing[0] / max(timing[6], 1.0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
                "config": workloadConfig(args, chains),
                "evals_per_sec": value * P * N,
                "evals_per_sec_per_gpu": value * P * N / world,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h / max(args.steps, 1))},
                "gpu_launches": int(timing[3] + timing[4] + timing[5]),
                "kernel_ms": {"step_kernel_avg": sweepMs, "hyper_kernel_avg": timing[1] / max(timing[7], 1.0),
                              "writeback_avg": timing[2] / max(timing[8], 1.0),
                              "timed_launches": int(timing[6] + timing[7] + timing[8]),
                              "step_kernel_share": sweepMs * timing[3] / ms},
                "roofline": roofline(args, tensorCore, sweepMs, chains, peakFlops, peakTf32.value, mufu.value),
                "clocks": clocks}
        if diag is not None and diag["min_ess"] is not None:
            samplingS = (args.warmup + args.steps) * msE2e / args.steps * 1e-3     # every iteration it took to get the rows
            diag["min_ess_per_sec"] = diag["min_ess"] / samplingS
            diag["note"] = "min over all keys of the effective sample size (sampleDiagnosis.py:232-255) of the retained " \
                           "rows of all chains / wall time of burn-in + sampling at the e2e rate"
        line["min_ess"] = diag
        if world == 1 and not args.no_cpu_baseline:
            v, dt = cpuBaseline(args, 1, args.cpu_iters)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "1 chain x %d iterations of the same workload (%.1f s), numpy "
                                              "restatement of the reference (oracle/)" % (args.cpu_iters, dt)}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(text):
    """The one JSON line, on the process's real stdout."""
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def main():
    global _REAL_STDOUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"],
                    help="c3 (default, the config BASELINE.json's metric is quoted on) or c5 (Bernoulli-logit: "
                         "10,000 groups x 50 trials, 4,096 chains; sets --groups/--obs/--coef/--chains-per-gpu)")
    ap.add_argument("--chains-per-gpu", type=int, default=1024)
    ap.add_argument("--groups", type=int, default=1024)
    ap.add_argument("--obs", type=int, default=200)
    ap.add_argument("--coef", type=int, default=8)
    ap.add_argument("--iters-per-step", type=int, default=50)
    ap.add_argument("--thin", type=int, default=10)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--cpu-iters", type=int, default=40)
    ap.add_argument("--ref-iters", type=int, default=5)
    ap.add_argument("--pooling", default="partial", choices=["partial", "none"],
                    help="partial (default, the BASELINE metric's mode) or none (BASELINE config 5 names both)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traffic", type=float, default=199.9e6,
                    help="dram__bytes_read.sum + dram__bytes_write.sum per step-kernel launch, from the committed "
                         "ncu --set full capture (profiles/r1_final_tc_kernel_ncu_summary.txt); not measured live")
    ap.add_argument("--traffic-pipe", type=float, default=204.2e6,
                    help="same for the FP32-pipe kernel (profiles/r1_sweep_kernel_ncu_summary.txt)")
    args = ap.parse_args()
    # Everything libraries print on file descriptor 1 (NCCL's version banner under NCCL_DEBUG=VERSION,
    # for one) goes to stderr; stdout carries the JSON line and nothing else.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.workload == "c5":
        args.groups, args.obs, args.coef, args.chains_per_gpu = 10000, 50, 0, 4096
        args.iters_per_step, args.thin = min(args.iters_per_step, 20), max(args.thin, 20)
    if args.warmup < 3 and args.impl == "b200":
        print("warning: the timing rules ask for >= 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        runReference(args)
    else:
        runGpu(args)


if __name__ == "__main__":
    main()
