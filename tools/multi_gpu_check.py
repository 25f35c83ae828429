#!/usr/bin/env python
"""Multi-GPU check, run under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py OUTDIR

1. samplePosterior with the chains sharded over the ranks writes, chain for chain, the same
   sample files as a single-GPU run (Philox and the start-state streams are keyed by the global
   chain id, so results do not depend on the GPU count).
2. Diagnostic over the sharded chains (all-gather of per-half-chain summaries over NCCL; pooled
   median / HDI through a key-partitioned all-to-all) equals Diagnostic over all chains in one process.
3. The same run forced into the binary store writes one shard file per rank plus one manifest;
   diagnoseSamples under the process group (every rank opens its own shard; all-gather / all-to-all
   for the between-chain statistics, pooled order statistics and Summary; rank 0 writes) produces the
   same files and stdout as a single process reading all shards.
Prints MULTI_GPU_OK on rank 0."""
import contextlib
import io
import shutil
import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mcmc-for-nested-data_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import parity  # noqa: E402
import posteriorSampling as ps  # noqa: E402
import sampleDiagnosis as sd  # noqa: E402
from objectives import Objective  # noqa: E402


def main():
    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    obj, names, nResp, ranges = parity.syntheticRegression(G=24, R=30, K=4)
    handle = Objective.linear_regression(obj.X, obj.y)
    nChains, nIter, nSamples = 12, 600, 200
    ps.samplePosterior(nChains, nIter, nSamples, names, 24, nResp, "partial", handle, out + "/multi",
                       saveLogLikelihood=False, startingPointValueRange=ranges, displayProgress=False)
    # sharded diagnostics: every rank loads only its own chains
    lo, hi = sd.chainRange(nChains, rank, world)
    keys, allSamples, chains = sd.loadSamples(out + "/multi/sample/")
    mine = allSamples[lo:hi]
    dShard = sd.Diagnostic(samples=mine, keys=keys, group=dist.group.WORLD)
    rhat, ess, med, hdi = dShard.rhat, dShard.effectiveN, dShard.median, dShard.hdi
    # 3. sharded binary store
    csvLimit, ps.CSV_VALUE_LIMIT = ps.CSV_VALUE_LIMIT, 0
    ps.samplePosterior(nChains, nIter, nSamples, names, 24, nResp, "partial", handle, out + "/multibin",
                       saveLogLikelihood=False, startingPointValueRange=ranges, displayProgress=False)
    ps.CSV_VALUE_LIMIT = csvLimit
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        sd.diagnoseSamples(out + "/multibin", nFigures=0)
    shardedStdout = buf.getvalue()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        files = ("diagnostic/diagnosticAssessment.csv", "diagnostic/diagnosticAssessmentHyperOnly.csv",
                 "diagnostic/diagnosticAssessmentIndividual.csv", "sample/summary.csv")
        sharded = [open(out + "/multibin/" + f).read() for f in files]
        shutil.rmtree(out + "/multibin/diagnostic")
        os.remove(out + "/multibin/sample/summary.csv")
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            sd.diagnoseSamples(out + "/multibin", nFigures=0)          # one process, all shards
        assert buf.getvalue() == shardedStdout and len(shardedStdout) > 1000
        assert [open(out + "/multibin/" + f).read() for f in files] == sharded
        kb, binSamples, chainsB = sd.loadSamples(out + "/multibin/sample/")
        assert kb == keys and chainsB == list(range(nChains))
        numpy.testing.assert_allclose(binSamples, allSamples, rtol=0, atol=5.1e-7)   # the CSV files hold the same draws at "%f"
        # single-GPU reference run of the same global chains (no process group any more)
        ps.samplePosterior(nChains, nIter, nSamples, names, 24, nResp, "partial", handle, out + "/single",
                           saveLogLikelihood=False, startingPointValueRange=ranges, displayProgress=False)
        for c in range(nChains):
            a = open(out + "/multi/sample/sample.%d.csv" % c).read()
            b = open(out + "/single/sample/sample.%d.csv" % c).read()
            assert a == b, "chain %d differs between %d GPUs and 1 GPU" % (c, world)
        dAll = sd.Diagnostic(samples=allSamples, keys=keys)
        for k in keys:
            numpy.testing.assert_allclose(rhat[k], dAll.rhat[k], rtol=1e-12)
            numpy.testing.assert_allclose(ess[k], dAll.effectiveN[k], rtol=1e-11)
            assert med[k] == dAll.median[k] and hdi[k] == dAll.hdi[k]     # key-partitioned all-to-all: exact
        print("MULTI_GPU_OK world=%d chains=%d keys=%d" % (world, nChains, len(keys)))


if __name__ == "__main__":
    main()
