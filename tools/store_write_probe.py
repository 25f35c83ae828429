#!/usr/bin/env python
"""How fast host threads fill a memory-mapped .npy in /dev/shm (the sample store's file): populate rate
(MADV_POPULATE_WRITE) and copy rate into populated / unpopulated pages.  usage: store_write_probe.py [GB] [threads]"""
import ctypes, os, sys, time, threading
import numpy
from concurrent.futures import ThreadPoolExecutor

GB = float(sys.argv[1]) if len(sys.argv) > 1 else 16
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 8
path = "/dev/shm/mcmcn_store_probe.npy"
rowBytes = 37822464            # one C3 row (9,234 columns x 1,024 chains x 4 bytes)
rows = int(GB * 1e9 // rowBytes)
libc = ctypes.CDLL(None, use_errno=True)
page = os.sysconf("SC_PAGE_SIZE")


def populate(sink, nThreads):
    addr = sink.ctypes.data
    lo0 = addr - addr % page
    end = addr + sink.nbytes
    step = 64 << 20

    def work(k):
        lo = lo0 + k * step
        while lo < end:
            n = min(step, end - lo)
            if libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(n), 23) != 0:
                print("madvise failed", ctypes.get_errno()); return
            lo += nThreads * step
    ts = [threading.Thread(target=work, args=(k,)) for k in range(nThreads)]
    t = time.time(); [x.start() for x in ts]; [x.join() for x in ts]
    return time.time() - t


def copy(sink, src, nThreads):
    ex = ThreadPoolExecutor(nThreads)
    ncol = sink.shape[1]
    per = 3
    step = -(-ncol // per)
    t = time.time()
    for r0 in range(0, rows - 2, 3):
        jobs = [(r0 + r, k0, min(ncol, k0 + step)) for r in range(3) for k0 in range(0, ncol, step)]

        def put(j):
            r, a, b = j
            sink[r, a:b] = src[r - r0, a:b]
        list(ex.map(put, jobs))
    dt = time.time() - t
    ex.shutdown()
    return dt


src = numpy.random.rand(3, 9234, 1024).astype(numpy.float32)
for mode in ("populated", "unpopulated"):
    if os.path.exists(path):
        os.remove(path)
    sink = numpy.lib.format.open_memmap(path, mode="w+", dtype=numpy.float32, shape=(rows, 9234, 1024))
    if mode == "populated":
        dt = populate(sink, NT)
        print("populate %d threads: %.2f s, %.1f GB/s" % (NT, dt, sink.nbytes / dt / 1e9))
    dt = copy(sink, src, NT)
    print("copy into %s pages, %d threads: %.2f s, %.1f GB/s" % (mode, NT, dt, sink.nbytes / dt / 1e9))
    del sink
os.remove(path)
