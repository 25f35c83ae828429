"""Drop-in for the reference's ``posteriorSampling`` module: same ``samplePosterior``
signature and outputs (/root/reference/posteriorSampling.py:28-216), but every chain is
advanced together on the GPU by the CUDA step kernels behind the C ABI (include/mcmcn.h).

What changes on purpose (BASELINE.json north star):
  * ``logLikelihoodFunction`` must be an ``objectives.Objective`` handle (a device function
    from the registry, or NVRTC-compiled source).  A Python callable raises TypeError; there is
    no CPU fallback.
  * ``nProcesses`` is accepted and ignored: chains are GPU threads, not OS processes.
  * proposals / accept uniforms come from per-chain Philox streams keyed by the global chain id
    (the reference uses numpy's global MT19937, seed = chain); start states still come from each
    chain's MT19937 stream in the reference's order, so they equal the reference's.
  * at scale the retained samples go to a binary store (``sample/samples.npy`` +
    ``sample/manifest.json``) that ``sampleDiagnosis`` reads; at example scale the reference's
    ``sample.<chain>.csv`` / ``logLikelihood.<chain>.csv`` files are written.
Under ``torch.distributed`` (one process per GPU) the chains are split contiguously over the
ranks; sampling needs no communication.
"""

import datetime
import json
import logging
import os
import shutil

import numpy
import torch

from engine import Engine, SampleStore, burnThin, retainedIterations
from objectives import Objective

# above this many retained values (rows x columns x chains) the samples go to the binary store
CSV_VALUE_LIMIT = int(os.environ.get("MCMCN_CSV_LIMIT", 20000000))
# the pointwise log-likelihood output is refused above this many values (rows x N x chains)
LOGLIK_VALUE_LIMIT = int(os.environ.get("MCMCN_LOGLIK_LIMIT", 200000000))


def samplePosterior(nChains, nIter, nSamples,
                    parameterName, nGroups, nResponsesPerGroup,
                    pooling, logLikelihoodFunction,
                    outputDirectory,
                    saveLogLikelihood=True,
                    priorDistribution=None,
                    startWithMLE=False, startingPointValueRange=None,
                    nProcesses=1, displayProgress=True, loggingLevel="info"):
    """Samples from the posterior distribution; samples are saved under ``outputDirectory``
    (see the reference's docstring, posteriorSampling.py:36-144, for the arguments)."""
    startTime = datetime.datetime.now()
    rank, world = _rankWorld()

    if rank == 0:
        if os.path.exists(outputDirectory):
            shutil.rmtree(outputDirectory)                         # :149-150
    _barrier(world)
    sampleDirectory = outputDirectory + "/sample/"
    os.makedirs(sampleDirectory, exist_ok=True)
    logDirectory = outputDirectory + "/log/"
    os.makedirs(logDirectory, exist_ok=True)

    logger = _getLogger(logDirectory + "samplePosterior.log", "samplePosterior", loggingLevel)
    msg = "MCMC sampling.\n"
    msg += "\tpooling: %s.\n" % pooling
    msg += "\tnChains: %i, nIterPerChain: %i, nSamplesPerChain: %i." % (nChains, nIter, nSamples)
    logger.info(msg)
    if displayProgress and rank == 0:
        print(msg)

    if not isinstance(logLikelihoodFunction, Objective):
        raise TypeError("logLikelihoodFunction must be an objectives.Objective handle (a device function); "
                        "Python callables cannot run on the GPU and there is no CPU fallback")
    if pooling not in ("partial", "none", "complete"):             # :1040-1043
        raise Exception("Invalid pooling: ", pooling)
    burn, thin = burnThin(nIter, nSamples)                          # :1018-1027

    # chains of this rank (contiguous global ids; Philox and the start-state RNG are keyed by them)
    lo, hi = (nChains * rank) // world, (nChains * (rank + 1)) // world
    myChains = hi - lo
    if myChains < 1:
        raise ValueError("fewer chains (%d) than ranks (%d)" % (nChains, world))

    eng = Engine(logLikelihoodFunction, nGroups, nResponsesPerGroup, pooling, myChains,
                 priorDistribution=priorDistribution, chainId0=lo,
                 seed=int(os.environ.get("MCMCN_SEED", "0")))
    chainLoggers = [_getLogger(logDirectory + "/mcmc.chain%.2i.log" % c, "mcmc.chain%.2i" % c, loggingLevel)
                    for c in range(lo, hi)] if myChains <= 64 else []
    if pooling == "partial" and priorDistribution is not None:
        logger.info("Partial pooling ignores prior distribution.")  # :713-714
    _progress(logger, displayProgress and rank == 0, "Started looking for a reasonable starting state.")
    eng.initialise(parameterName, startingPointValueRange, startWithMLE, logger)
    _progress(logger, displayProgress and rank == 0, "Found a reasonable starting state.")

    retained = retainedIterations(nIter, burn, thin)
    nValues = len(retained) * eng.nCol * nChains
    useCsv = nValues <= CSV_VALUE_LIMIT
    store = SampleStore(eng, max(len(retained), 1), torch.float64 if useCsv else torch.float32)
    pointwise = []
    if saveLogLikelihood:
        if len(retained) * eng.nObservations * nChains > LOGLIK_VALUE_LIMIT:
            raise ValueError("saveLogLikelihood=True would write %d x %d x %d log-likelihood values; pass "
                             "saveLogLikelihood=False (or raise MCMCN_LOGLIK_LIMIT)"
                             % (len(retained), eng.nObservations, nChains))

    # ---- Sampler._loop (:862-896): segments end at progress marks and, when the pointwise
    # log-likelihood is wanted, at every retained iteration
    stops = set([nIter])
    loggingInterval = int(numpy.round(nIter / 10.))
    if loggingInterval > 0:
        stops.update(range(loggingInterval, nIter, loggingInterval))
    if saveLogLikelihood:
        stops.update(i + 1 for i in retained)
    loopStart = datetime.datetime.now()
    _progress(logger, displayProgress and rank == 0, r"Sampling started. 0% complete.")
    cur = 0
    for stop in sorted(stops):
        eng.run(cur, stop - cur, burn, thin, store=store)
        cur = stop
        if saveLogLikelihood and (stop - 1) in retained:
            pointwise.append(eng.pointwiseLogLikelihood())          # [N][myChains]
        if loggingInterval > 0 and stop % loggingInterval == 0 and stop < nIter:
            torch.cuda.synchronize()
            now = datetime.datetime.now()
            percentage = float(stop) / nIter
            remain = (1 - percentage) * (now - loopStart) / percentage
            _progress(logger, displayProgress and rank == 0,
                      "%i%% complete. ETA: %s." % (percentage * 100, _getStrfTime(now + remain)))
    torch.cuda.synchronize()
    elapsed = datetime.datetime.now() - loopStart
    _progress(logger, displayProgress and rank == 0, "100%% complete. Elapsed Time: %s."
              % datetime.timedelta(seconds=int(elapsed.total_seconds())))

    # ---- outputs
    rows = store.hostArray()                                        # [rows][ncol][myChains]
    header = sampleHeader(parameterName, eng.G, pooling)
    if useCsv:
        for c in range(myChains):
            writeSampleCsv(sampleDirectory + "/sample.%i.csv" % (lo + c), lo + c, header,
                           retained, rows[:, :, c])
    else:
        _writeBinaryStore(sampleDirectory, rows, header, retained, lo, hi, rank, world, pooling)
    if saveLogLikelihood:
        for c in range(myChains):
            with open(sampleDirectory + "/logLikelihood.%i.csv" % (lo + c), "w") as h:
                for pw in pointwise:
                    h.write(",".join(["%f" % v for v in pw[:, c]]))
                    h.write("\n")
    for c, lg in enumerate(chainLoggers):
        lg.info("chain %i. 100%% complete. Elapsed Time: %s."
                % (lo + c, datetime.timedelta(seconds=int(elapsed.total_seconds()))))

    _barrier(world)
    endTime = datetime.datetime.now()
    msg = "Finished. The elapsed time in total is %s." \
        % datetime.timedelta(seconds=int((endTime - startTime).total_seconds()))
    logger.info(msg)
    if displayProgress and rank == 0:
        print("")
        _printProgress(msg)


# ------------------------------------------------------------------------------ output format
def sampleHeader(parameterName, nSteppedGroups, pooling):
    """Column names of a sample file (posteriorSampling.py:640-646, :771-778, :504-507, :263)."""
    cols = []
    for name in parameterName:
        if pooling == "partial":
            cols += ["%s_mu" % name, "%s_sigma2" % name]
        cols += ["%s[%.3i]" % (name, j) for j in range(nSteppedGroups)]
    return cols


def writeSampleCsv(path, chain, header, iterations, rows):
    """``sample.<chain>.csv`` exactly as Sampler._printHeader/_printSample write it (:898-905)."""
    with open(path, "w") as h:
        h.write("index,chain," + ",".join(header))
        h.write("\n")
        for i, row in zip(iterations, rows):
            h.write("%i,%i," % (i, chain))
            h.write(",".join(["%f" % v for v in row]))
            h.write("\n")


def _writeBinaryStore(sampleDirectory, rows, header, iterations, lo, hi, rank, world, pooling):
    """samples.npy [rows][ncol][chains] (+ manifest.json); one file per rank when sharded."""
    name = "samples.npy" if world == 1 else "samples.rank%d.npy" % rank
    numpy.save(os.path.join(sampleDirectory, name), rows)
    if world == 1:
        man = {"file": name, "header": header, "iterations": list(map(int, iterations)),
               "chains": list(range(lo, hi)), "pooling": pooling, "dtype": str(rows.dtype),
               "layout": "[rows][columns][chains]"}
        with open(os.path.join(sampleDirectory, "manifest.json"), "w") as h:
            json.dump(man, h)
    else:
        man = {"file": name, "header": header, "iterations": list(map(int, iterations)),
               "chains": list(range(lo, hi)), "pooling": pooling, "dtype": str(rows.dtype),
               "layout": "[rows][columns][chains]", "rank": rank, "world": world}
        with open(os.path.join(sampleDirectory, "manifest.rank%d.json" % rank), "w") as h:
            json.dump(man, h)


# ------------------------------------------------------------------------------ plumbing
def _rankWorld():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def _progress(logger, display, msg):
    logger.info(msg)
    if display:
        _printProgress(msg)


def _getLogger(logFile, logName, loggingLevel):
    """:1161-1182 (one handler per log file; the reference adds a handler on every call)."""
    level = {"debug": logging.DEBUG, "info": logging.INFO, "warning": logging.WARNING,
             "error": logging.ERROR}[loggingLevel]
    logger = logging.getLogger(logName)
    logger.setLevel(level)
    for hd in list(logger.handlers):
        logger.removeHandler(hd)
        hd.close()
    handler = logging.FileHandler(logFile)
    handler.setLevel(level)
    handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s\n%(message)s\n"))
    logger.addHandler(handler)
    return logger


def _printProgress(msg):
    print(_getStrfTime(datetime.datetime.now()) + "\t" + msg)


def _getStrfTime(time):
    return time.strftime("%Y/%m/%d %H:%M:%S")
