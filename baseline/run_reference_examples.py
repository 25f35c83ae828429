#!/usr/bin/env python
"""BASELINE configs 1 and 2 on the UNMODIFIED reference (baseline/_ref, staged by __graft_entry__.build()):

    python baseline/run_reference_examples.py distribution partial
    python baseline/run_reference_examples.py regression partial|none|complete

Runs the reference example's own main() -- its data generator, objective, priors, chain counts and
samplePosterior / diagnoseSamples calls, unchanged -- with one substitution: diagnoseSamples is called
with nFigures=0 (matplotlib is not installed in this image, and the reference's figure code uses a
keyword current matplotlib removed).  Prints the wall seconds of sampling and of the diagnostics."""
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
sys.path.insert(0, REF)


def main():
    which, pooling = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "partial")
    if not os.path.exists(os.path.join(REF, "example", which + ".py")):
        print(json.dumps({"unavailable": "baseline/_ref is not staged"}))
        return
    os.chdir(tempfile.mkdtemp(prefix="mcmcn_ref_example_"))
    import importlib
    import posteriorSampling
    import sampleDiagnosis
    assert os.path.dirname(os.path.abspath(posteriorSampling.__file__)) == REF
    mod = importlib.import_module("example." + which)
    t = {}
    realSample, realDiagnose = mod.samplePosterior, mod.diagnoseSamples

    def timedSample(*a, **k):
        t0 = time.perf_counter()
        realSample(*a, **k)
        t["samplePosterior_s"] = time.perf_counter() - t0

    def timedDiagnose(outputDirectory, *a, **k):
        t0 = time.perf_counter()
        realDiagnose(outputDirectory, nFigures=0)
        t["diagnoseSamples_s"] = time.perf_counter() - t0
    mod.samplePosterior, mod.diagnoseSamples = timedSample, timedDiagnose
    t0 = time.perf_counter()
    mod.main(pooling)
    t["total_s"] = time.perf_counter() - t0
    t.update({"example": which, "pooling": pooling, "cores": os.cpu_count()})
    print("REFERENCE_EXAMPLE " + json.dumps(t))


if __name__ == "__main__":
    main()
