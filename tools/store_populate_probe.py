#!/usr/bin/env python
"""Is the page-population rate of a tmpfs file per file or per box?  Populates 16 GB as 1, 2 and 4 files (8 threads in
all) with MADV_POPULATE_WRITE, and once with posix_fallocate from 8 threads on disjoint ranges."""
import ctypes, mmap, os, sys, time, threading

libc = ctypes.CDLL(None, use_errno=True)
TOTAL = 16 << 30
NT = 8
step = 64 << 20


def populate_files(nFiles):
    size = TOTAL // nFiles
    maps = []
    for f in range(nFiles):
        p = "/dev/shm/mcmcn_pop_probe_%d" % f
        fd = os.open(p, os.O_RDWR | os.O_CREAT | os.O_TRUNC)
        os.ftruncate(fd, size)
        m = mmap.mmap(fd, size)
        os.close(fd)
        maps.append((p, m, ctypes.addressof(ctypes.c_char.from_buffer(m))))
    per = NT // nFiles

    def work(f, k):
        lo = maps[f][2] + k * step
        end = maps[f][2] + size
        while lo < end:
            if libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(min(step, end - lo)), 23) != 0:
                print("madvise errno", ctypes.get_errno()); return
            lo += per * step
    ts = [threading.Thread(target=work, args=(f, k)) for f in range(nFiles) for k in range(per)]
    t = time.time(); [x.start() for x in ts]; [x.join() for x in ts]
    dt = time.time() - t
    print("%d file(s), %d threads: %.2f s, %.1f GB/s" % (nFiles, NT, dt, TOTAL / dt / 1e9))
    for p, m, a in maps:
        del a
    for p, m, a in maps:
        try:
            m.close()
        except BufferError:
            pass
        os.remove(p)


def fallocate():
    p = "/dev/shm/mcmcn_pop_probe_f"
    fd = os.open(p, os.O_RDWR | os.O_CREAT | os.O_TRUNC)
    os.ftruncate(fd, TOTAL)

    def work(k):
        lo = k * step
        while lo < TOTAL:
            os.posix_fallocate(fd, lo, min(step, TOTAL - lo))
            lo += NT * step
    ts = [threading.Thread(target=work, args=(k,)) for k in range(NT)]
    t = time.time(); [x.start() for x in ts]; [x.join() for x in ts]
    dt = time.time() - t
    print("posix_fallocate, %d threads: %.2f s, %.1f GB/s" % (NT, dt, TOTAL / dt / 1e9))
    m = mmap.mmap(fd, TOTAL)
    a = ctypes.addressof(ctypes.c_char.from_buffer(m))
    t = time.time()
    def work2(k):
        lo = a + k * step
        while lo < a + TOTAL:
            libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(min(step, a + TOTAL - lo)), 23)
            lo += NT * step
    ts = [threading.Thread(target=work2, args=(k,)) for k in range(NT)]
    [x.start() for x in ts]; [x.join() for x in ts]
    dt = time.time() - t
    print("   then mapping the allocated pages: %.2f s, %.1f GB/s" % (dt, TOTAL / dt / 1e9))
    os.close(fd)
    os.remove(p)


for n in (1, 2, 4):
    populate_files(n)
fallocate()
print(open("/sys/kernel/mm/transparent_hugepage/shmem_enabled").read().strip())
