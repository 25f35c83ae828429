"""Shared driver of the two example workloads (BASELINE configs 1 and 2): a workload is a small record --
names, sizes, chain counts, how to draw its mock data -- and `run` feeds it to the drop-in calls."""

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mcmc-for-nested-data_b200"))

from posteriorSampling import samplePosterior  # noqa: E402
from sampleDiagnosis import diagnoseSamples  # noqa: E402


class Workload(object):
    def __init__(self, title, directory, names, groups, responses, chains, iterations, retained, **samplerOptions):
        self.title, self.directory, self.names = title, directory, tuple(names)
        self.groups, self.responses = groups, responses
        self.chains, self.iterations, self.retained = chains, iterations, retained
        self.samplerOptions = samplerOptions

    def run(self, pooling, objective, truth):
        samplePosterior(self.chains, self.iterations, self.retained, self.names, self.groups, self.responses,
                        pooling, objective, self.directory, **self.samplerOptions)
        print(truth)
        diagnoseSamples(self.directory)


def poolingFromCommandLine(title):
    parser = argparse.ArgumentParser(description=title)
    parser.add_argument("pooling", nargs="?", default="partial", choices=["partial", "complete", "none"],
                        help="pooling method (default: partial)")
    return parser.parse_args().pooling
