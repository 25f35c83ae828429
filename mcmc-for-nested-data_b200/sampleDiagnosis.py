"""Drop-in for the reference's ``sampleDiagnosis`` module, computed on the GPU.

Same public names, signatures, files written and text printed as
``/root/reference/sampleDiagnosis.py`` (``diagnoseSamples`` :11-85, ``Diagnostic`` :88-427,
``Summary`` :430-491, ``computeHpdInterval`` :766-776); figures (``Figure`` :494-759) are out
of scope and skipped.  The reductions run as batched FP64 kernels through the C ABI
(include/mcmcn.h, ``mcmcn_diag_*``); the reference's quirks are kept on purpose: ESS sums
rho from lag 0, the truncation scan tests only even lags, the HDI gap uses Python's
banker's rounding, pooling mode is detected by substring, names are cut to 40 bytes.

Input is either the reference-format ``sample/sample.<chain>.csv`` files or the binary
store (``sample/manifest.json`` + ``samples.npy``) that ``samplePosterior`` writes at scale.
Under ``torch.distributed`` (one process per GPU, chains sharded) the per-half-chain moments
and per-lag sums are exchanged with one all-gather each and merged in rank order.
"""

import ctypes
import glob
import json
import os

import numpy
import pandas
import torch

import mcmcn_native as nat


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("sampleDiagnosis needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _sortedMedianHdiDevice(x, hdi_p):
    """x: device double [n_keys][len] (destroyed: sorted in place).  Returns device [n_keys][3]
    = median, HDI lower, HDI upper (computeHpdInterval, :766-776)."""
    dev = x.device
    nKeys, length = x.shape
    prob = hdi_p / 100.
    gap = max(1, min(length - 1, round(length * prob)))          # Python banker's rounding
    out = torch.empty((nKeys, 3), dtype=torch.float64, device=dev)
    if nKeys:
        nat.call("mcmcn_diag_sort_keys", _ptr(x), nKeys, length, _stream(dev))
        nat.call("mcmcn_diag_median_hdi", _ptr(x), nKeys, length, gap, _ptr(out), _stream(dev))
    return out


def _sortedMedianHdi(x, hdi_p):
    return _sortedMedianHdiDevice(x, hdi_p).cpu().numpy()


def keyRange(nKeys, rank, world):
    """Contiguous slice [lo, hi) of the keys a rank owns in the key-partitioned exchange."""
    return (nKeys * rank) // world, (nKeys * (rank + 1)) // world


def exchangeByKey(local, group):
    """Key-partitioned exchange for the pooled order statistics:
    ``local`` is this rank's [n_keys][len] draws; rank r becomes the owner of keys keyRange(r)
    and receives every rank's draws of those keys, concatenated in rank (= chain) order:
    returns [keys_owned][world * len].  One all-to-all over NVLink with NCCL; backends without
    all-to-all (gloo, used by the CPU tests) run the same exchange as one gather per owner."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nKeys, length = local.shape
    send = [local[slice(*keyRange(nKeys, r, world))].contiguous() for r in range(world)]
    lo, hi = keyRange(nKeys, rank, world)
    recv = [torch.empty((hi - lo, length), dtype=local.dtype, device=local.device) for _ in range(world)]
    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv, send, group=group)
    else:
        for r in range(world):
            dist.gather(send[r], gather_list=recv if r == rank else None, dst=dist.get_global_rank(group, r) if group is not None else r,
                        group=group)
    return torch.cat(recv, dim=1).contiguous()


def pooledMedianHdi(local, hdi_p, group, keySlab=2048):
    """numpy.median and the HDI (:419-427, :766-776) of every key over ALL ranks' draws without
    any rank ever holding all draws of all keys: keys go through exchangeByKey in slabs, their
    owners sort them, and the [n_keys][3] results are all-gathered.  Returns device [n_keys][3]."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nKeys = local.shape[0]
    out = torch.empty((nKeys, 3), dtype=torch.float64, device=local.device)
    for k0 in range(0, nKeys, keySlab):
        k1 = min(nKeys, k0 + keySlab)
        mine = _sortedMedianHdiDevice(exchangeByKey(local[k0:k1], group), hdi_p)     # [keys owned in this slab][3]
        most = max(keyRange(k1 - k0, r, world)[1] - keyRange(k1 - k0, r, world)[0] for r in range(world))
        pad = torch.zeros((most, 3), dtype=torch.float64, device=local.device)
        pad[:mine.shape[0]] = mine
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        for r in range(world):
            lo, hi = keyRange(k1 - k0, r, world)
            out[k0 + lo:k0 + hi] = parts[r][:hi - lo]
    return out


def computeHpdInterval(samples, hdi_p=95):
    """sampleDiagnosis.py:766-776."""
    dev = _device()
    x = torch.as_tensor(numpy.ascontiguousarray(samples, dtype=numpy.float64), device=dev).reshape(1, -1).clone()
    res = _sortedMedianHdi(x, hdi_p)
    return (res[0, 1], res[0, 2])


def loadSamples(sampleDirectory):
    """Returns (keys in column order, array [nChains][rows][nKeys] float64, chain ids).
    Reads the binary store if its manifest is present, else every ``sample*.csv`` (:102, :132)."""
    manifest = os.path.join(sampleDirectory, "manifest.json")
    if os.path.exists(manifest):
        with open(manifest) as h:
            man = json.load(h)
        arr = numpy.load(os.path.join(sampleDirectory, man["file"]), mmap_mode="r")   # [rows][ncol][nChains]
        data = numpy.ascontiguousarray(numpy.transpose(arr, (2, 0, 1)), dtype=numpy.float64)
        return list(man["header"]), data, list(man["chains"])
    files = glob.glob(sampleDirectory + "/sample*.csv")
    if not files:
        raise FileNotFoundError("no sample*.csv under %s" % sampleDirectory)

    def chainOf(path):      # sample.<chain>.csv; the reference takes glob order, which is arbitrary
        parts = os.path.basename(path).split(".")
        return int(parts[1]) if len(parts) > 2 and parts[1].isdigit() else 1 << 30

    files.sort(key=lambda f: (chainOf(f), f))
    frames, keys = [], None
    for filename in files:
        d = pandas.read_csv(filename, float_precision="round_trip")
        if keys is None:
            keys = [k for k in d.columns if k not in ("chain", "index")]
        frames.append(d[keys].to_numpy(dtype=numpy.float64))
    return keys, numpy.stack(frames), [chainOf(f) for f in files]


def gatherShards(t, group):
    """all-gather a [n_keys][m_local][..] tensor along the half-chain axis (dim 1), rank order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous(), group=group)
    return torch.cat(parts, dim=1).contiguous()


def mergeShards(mean, var, vario, group):
    """The one exchange the between-chain diagnostics need (SURVEY.md sections 5 and 8e): every
    rank contributes its chains' half-chain means / variances [n_keys][m_local] and its per-lag
    sums [n_keys][n]; returns (mean, var) over all m half-chains and the per-lag sums added in
    fixed rank order, so every rank gets bit-identical results."""
    mean, var = gatherShards(mean, group), gatherShards(var, group)
    parts = gatherShards(vario.reshape(vario.shape[0], 1, vario.shape[1]), group)   # [n_keys][world][n]
    total = torch.zeros_like(vario)
    for r in range(parts.shape[1]):
        total += parts[:, r, :]
    return mean, var, total


def convergenceFromStore(storeTensor, nRows, nChains, group=None):
    """R-hat and effective sample size of every column of a device-resident sample store
    ([rows][ncol][S], chain fastest; engine.SampleStore) without leaving the GPU: the same
    kernels and formulas as Diagnostic (:158-255), the rows split into first / second half per
    chain (:118-156).  With ``group`` every rank holds a shard of the chains and the half-chain
    moments and per-lag sums are exchanged by one NCCL all-gather (mergeShards).
    Returns (rhat[ncol], ess[ncol]) as device tensors."""
    n = int(nRows) // 2
    if n < 2:
        raise ValueError("need at least 4 retained rows per chain")
    ncol = storeTensor.shape[1]
    dev = storeTensor.device
    x = storeTensor[:2 * n, :, :nChains].permute(1, 2, 0).to(torch.float64)        # [ncol][nC][2n]
    x = x.reshape(ncol, 2 * nChains, n).contiguous()                               # half-chains: chain c -> 2c, 2c+1
    st = _stream(dev)
    mL = 2 * nChains
    mean = torch.empty((ncol, mL), dtype=torch.float64, device=dev)
    var = torch.empty((ncol, mL), dtype=torch.float64, device=dev)
    nat.call("mcmcn_diag_moments", _ptr(x), ncol, mL, n, _ptr(mean), _ptr(var), st)
    vario = torch.empty((ncol, n), dtype=torch.float64, device=dev)
    nat.call("mcmcn_diag_variogram", _ptr(x), ncol, mL, n, _ptr(vario), st)
    if group is not None:
        mean, var, vario = mergeShards(mean, var, vario, group)
    m = mean.shape[1]
    rh = torch.empty((ncol, 4), dtype=torch.float64, device=dev)
    nat.call("mcmcn_diag_rhat", _ptr(mean), _ptr(var), ncol, m, n, _ptr(rh), st)
    ess = torch.empty((ncol,), dtype=torch.float64, device=dev)
    nat.call("mcmcn_diag_ess", _ptr(vario), _ptr(rh), ncol, m, n, None, _ptr(ess), st)
    return rh[:, 3].contiguous(), ess


def chainRange(nChains, rank, world):
    """Contiguous global chain ids [lo, hi) of a rank."""
    return (nChains * rank) // world, (nChains * (rank + 1)) // world


class Diagnostic(object):
    def __init__(self, sampleDirectory=None, samples=None, keys=None, group=None):
        """Diagnostic(sampleDirectory) as in the reference (:89-116).  Alternatively pass
        ``samples`` = array [nChains][rows][nKeys] (+ ``keys``) already in memory.
        ``group``: a torch.distributed process group whose ranks each hold a shard of the
        chains; results are then over all ranks' chains."""
        if samples is None:
            keys, samples, _ = loadSamples(sampleDirectory)
        samples = numpy.asarray(samples, dtype=numpy.float64)
        self._keys = list(keys)
        self._group = group
        self._organiseSamples(samples)
        self._done = False
        self._hdiP = 95
        self._assessment = None
        self._summary = None

    def _organiseSamples(self, samples):
        """:118-156 -- split every chain's rows into first / second half."""
        nChains, N, nKeys = samples.shape
        n = N // 2
        if N != 2 * n:
            # the reference fails with a broadcast error on an odd row count
            raise ValueError("could not broadcast input array from shape (%d,) into shape (%d,)" % (N - n, n))
        self._mLocal = 2 * nChains
        self._n = n
        self.partiallyPooled = any("_" in k for k in self._keys)          # :143-144
        self.completelyPooled = not any("01]" in k for k in self._keys)   # :146-147
        dev = _device()
        # [nKeys][m][n]: key-major, half-chains (chain c -> rows 2c, 2c+1), draws contiguous
        x = numpy.ascontiguousarray(numpy.transpose(samples.reshape(nChains, 2, n, nKeys), (3, 0, 1, 2)))
        self._x = torch.from_numpy(x.reshape(nKeys, self._mLocal, n)).to(dev)
        self._m = self._mLocal
        if self._group is not None:
            import torch.distributed as dist
            self._m = self._mLocal * dist.get_world_size(self._group)

    # ---------------------------------------------------------------- device pipeline
    def _compute(self):
        if self._done:
            return
        dev = self._x.device
        nKeys, mL, n = self._x.shape
        st = _stream(dev)
        f64 = torch.float64
        mean = torch.empty((nKeys, mL), dtype=f64, device=dev)
        var = torch.empty((nKeys, mL), dtype=f64, device=dev)
        nat.call("mcmcn_diag_moments", _ptr(self._x), nKeys, mL, n, _ptr(mean), _ptr(var), st)
        vario = torch.empty((nKeys, n), dtype=f64, device=dev)
        nat.call("mcmcn_diag_variogram", _ptr(self._x), nKeys, mL, n, _ptr(vario), st)
        if self._group is not None:
            mean, var, vario = mergeShards(mean, var, vario, self._group)
        m = mean.shape[1]
        rh = torch.empty((nKeys, 4), dtype=f64, device=dev)
        nat.call("mcmcn_diag_rhat", _ptr(mean), _ptr(var), nKeys, m, n, _ptr(rh), st)
        ess = torch.empty((nKeys,), dtype=f64, device=dev)
        rho = torch.empty((nKeys, n), dtype=f64, device=dev)
        nat.call("mcmcn_diag_ess", _ptr(vario), _ptr(rh), nKeys, m, n, _ptr(rho), _ptr(ess), st)
        pooled = self._x.reshape(nKeys, mL * n)
        if self._group is not None:       # key-partitioned all-to-all: no rank holds every key's pooled draws
            mh = pooledMedianHdi(pooled, self._hdiP, self._group).cpu().numpy()
        else:
            mh = _sortedMedianHdi(pooled.clone(), self._hdiP)
        rh_h, ess_h = rh.cpu().numpy(), ess.cpu().numpy()
        self._rhoArr = rho.cpu().numpy()
        k = self._keys
        self._B = dict(zip(k, rh_h[:, 0]))
        self._W = dict(zip(k, rh_h[:, 1]))
        self._vhat = dict(zip(k, rh_h[:, 2]))
        self._rhat = dict(zip(k, rh_h[:, 3]))
        self._rho = dict(zip(k, self._rhoArr))
        self._effectiveN = dict(zip(k, ess_h))
        self._median = dict(zip(k, mh[:, 0]))
        self._hdi = dict((key, (mh[i, 1], mh[i, 2])) for i, key in enumerate(k))
        self._done = True

    @property
    def rhat(self):
        self._compute()
        return self._rhat

    @property
    def effectiveN(self):
        self._compute()
        return self._effectiveN

    @property
    def median(self):
        self._compute()
        return self._median

    @property
    def hdi(self):
        self._compute()
        return self._hdi

    # ---------------------------------------------------------------- tables (:257-405)
    # One column specification per table drives the structured array (what `.assessment` /
    # `.summary` return, as in the reference) and the CSV text; both tables are derived lazily.
    _ASSESSMENT_SPEC = (("parameter", "S40", "'%s'"), ("rhat", float, "%.3f"), ("converged", bool, "%s"),
                        ("effective n", float, "%.3f"), ("enough n", bool, "%s"), ("median", float, "%.3f"),
                        ("HDI lower", float, "%.3f"), ("HDI upper", float, "%.3f"))
    _SUMMARY_SPEC = (("parameter", "S40", "'%s'"), ("rhat min", float, "%.3f"), ("rhat median", float, "%.3f"),
                     ("rhat max", float, "%.3f"), ("proportion converged", float, "%.3f"))

    @staticmethod
    def _structured(spec, records):
        return numpy.array(records, dtype=[(title, kind) for title, kind, _ in spec])

    @staticmethod
    def _csvText(spec, table, keep=None):
        """Header line + one formatted line per record whose (decoded) name passes `keep`."""
        pattern = ",".join(fmt for _, _, fmt in spec)
        lines = [",".join(title for title, _, _ in spec)]
        for record in table:
            label = record["parameter"].decode("ascii")
            if keep is None or keep(label):
                lines.append(pattern % ((label,) + tuple(record)[1:]))
        return "\n".join(lines) + "\n"

    @property
    def assessment(self):
        """Per key: R-hat, converged (< 1.1), ESS, enough (> 10 m), median, HDI; sorted by name bytes (:263-289)."""
        if self._assessment is None:
            self._compute()
            enough = self._m * 10
            records = []
            for key in self._keys:
                rhat, ess = self._rhat[key], self._effectiveN[key]
                lower, upper = self._hdi[key]
                records.append((key.encode(), rhat, rhat < 1.1, ess, ess > enough, self._median[key], lower, upper))
            self._assessment = numpy.sort(self._structured(self._ASSESSMENT_SPEC, records), order="parameter")
        return self._assessment

    def _assess(self):
        return self.assessment

    @property
    def summary(self):
        """Per parameter name, over its per-group keys (those with a '['): min / median / max R-hat and the
        share of converged keys (:297-329)."""
        if self._summary is None:
            perName = {}
            for record in self.assessment:
                label = record["parameter"].decode("ascii")
                if "[" in label:
                    bucket = perName.setdefault(label.split("[")[0], ([], []))
                    bucket[0].append(record["rhat"])
                    bucket[1].append(record["converged"])
            records = [(name, min(r), numpy.median(r), max(r), numpy.mean(c))
                       for name, (r, c) in sorted(perName.items())]
            self._summary = self._structured(self._SUMMARY_SPEC, records)
        return self._summary

    def _summarise(self):
        return self.summary

    def print(self, csvfile, individualSummary, hyperOnly):
        """Write one of the three tables to `csvfile`, or to stdout under its banner when None (:331-379)."""
        refusals = ((individualSummary and self.completelyPooled,
                     "MCMC was completely pooled. There is no individual summary."),
                    (hyperOnly and not self.partiallyPooled,
                     "MCMC was not partially pooled. There is no hyper-parameter."),
                    (individualSummary and hyperOnly, "Choose individualSummary or hyperOnly. Not both."))
        for refused, why in refusals:
            if refused:
                raise ValueError(why)
        text = self._getSummaryString() if individualSummary else self._getAssessmentString(hyperOnly)
        if csvfile is not None:
            with open(csvfile, "w") as h:
                h.write(text)
            return
        banner = "MCMC convergence diagnostic."
        if hyperOnly:
            banner = "MCMC convergence diagnostic for hyper-parameters."
        elif individualSummary:
            banner = "Summary of MCMC convergence diagnostic."
        print(banner)
        _stdout_csv(text)

    def _getAssessmentString(self, hyperOnly):
        keep = (lambda label: "_" in label) if hyperOnly else None       # hyper-parameters: <name>_mu, <name>_sigma2
        return self._csvText(self._ASSESSMENT_SPEC, self.assessment, keep)

    def _getSummaryString(self):
        return self._csvText(self._SUMMARY_SPEC, self.summary)


class Summary(object):
    """Summarise individual parameter values (:430-491): per retained row the mean and the
    median over groups of each parameter name, then mean / median / 95% HDI of those."""

    def __init__(self, sampleDirectory=None, samples=None, keys=None):
        if samples is None:
            keys, samples, _ = loadSamples(sampleDirectory)
        samples = numpy.asarray(samples, dtype=numpy.float64)      # [nChains][rows][nKeys]
        self._n = samples.shape[1]
        names = numpy.unique([name.split("[")[0] for name in keys if "[" in name])
        dev = _device()
        st = _stream(dev)
        self._rows = {}
        for s in ("groupMean", "groupMedian"):
            self._rows[s] = {}
        for name in names:
            cols = [i for i, key in enumerate(keys) if (name + "[") in key]      # substring match, :466-467
            x = torch.from_numpy(numpy.ascontiguousarray(samples[:, :, cols])).to(dev)   # [chains][rows][G]
            rows = x.shape[0] * x.shape[1]
            G = len(cols)
            flat = x.reshape(rows, G).contiguous()
            mean = torch.empty((rows,), dtype=torch.float64, device=dev)
            nat.call("mcmcn_diag_row_mean", _ptr(flat), rows, G, _ptr(mean), st)
            if G >= 2:
                srt = flat.clone()
                nat.call("mcmcn_diag_sort_keys", _ptr(srt), rows, G, st)
                mh3 = torch.empty((rows, 3), dtype=torch.float64, device=dev)
                nat.call("mcmcn_diag_median_hdi", _ptr(srt), rows, G, 1, _ptr(mh3), st)
                med = mh3[:, 0]
            else:
                med = flat[:, 0]
            for s, v in (("groupMean", mean), ("groupMedian", med)):
                v = v.contiguous()
                avg = torch.empty((1,), dtype=torch.float64, device=dev)
                nat.call("mcmcn_diag_row_mean", _ptr(v), 1, rows, _ptr(avg), st)
                mh = _sortedMedianHdi(v.reshape(1, rows).clone(), 95.)
                self._rows[s][str(name)] = (float(avg.cpu()[0]), mh[0, 0], mh[0, 1], mh[0, 2])
        self._summarise()

    def _summarise(self):
        self._summary = "stats,parameter,mean,median,HDI lower,HDI upper\n"
        for s in ("groupMean", "groupMedian"):
            for name in sorted(self._rows[s]):
                mean, median, lo, hi = self._rows[s][name]
                self._summary += "%s,%s,%.4f,%.4f,%.4f,%.4f\n" % (s, name, mean, median, lo, hi)

    def print(self, csvfile):
        if csvfile is None:
            print("Summary of individual parameters.")
            _stdout_csv(self._summary)
        else:
            with open(csvfile, "w") as h:
                h.write(self._summary)


def diagnoseSamples(outputDirectory, assessConvergence=True, printSummary=True, nFigures=10):
    """Diagnose samples (sampleDiagnosis.py:11-85).  Same files and stdout as the reference;
    ``nFigures`` is accepted for compatibility but figures are not produced."""
    sampleDirectory = outputDirectory + "/sample/"
    diagnosticDirectory = outputDirectory + "/diagnostic/"
    os.makedirs(diagnosticDirectory, exist_ok=True)

    if assessConvergence:
        print("- Convergence Diagnostic -")
        diagnostic = Diagnostic(sampleDirectory)
        pooled, hyper = diagnostic.completelyPooled, diagnostic.partiallyPooled
        # (wanted, file, individualSummary, hyperOnly, echoed to stdout as well)
        reports = ((True, "diagnosticAssessment.csv", False, False, pooled),
                   (hyper, "diagnosticAssessmentHyperOnly.csv", False, True, True),
                   (not pooled, "diagnosticAssessmentIndividual.csv", True, False, True))
        for wanted, fileName, individual, hyperOnly, echo in reports:
            if not wanted:
                continue
            diagnostic.print(diagnosticDirectory + "/" + fileName, individual, hyperOnly)
            if echo:
                diagnostic.print(None, individual, hyperOnly)

    if printSummary:
        summary = Summary(sampleDirectory)
        summary.print(sampleDirectory + "/summary.csv")
        summary.print(None)
    # figures (Figure, :494-759) are out of scope for the GPU engine


def _stdout_csv(content):
    """Tab-indented, comma-spaced echo of a CSV text (:762-763)."""
    print("\t" + "\n\t".join(line.replace(",", ", ") for line in content.split("\n")))
