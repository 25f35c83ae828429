"""Figures of a run (reference: ``Figure``, sampleDiagnosis.py:494-759): the log-likelihood trace, the
trace plots (histogram + trace of every parameter of a group, :595-664) and the bivariate scatter matrices
(:690-759), written as PNG files under the names the reference uses.

The reference draws them with matplotlib, which this image does not have (and whose ``normed`` keyword the
reference's ``_hist`` passes was removed upstream, so its trace plots no longer run on a current install).
Here the plots are rasterised directly: a small RGB canvas on numpy (alpha-blended spans, polylines sampled
per pixel, point clouds accumulated as counts and blended once per chain), a 5 x 7 bitmap typeface for
titles / labels / ticks, and a PNG encoder on zlib.  Same content per file: one colour per chain (the
reference's list of seven, cycled when there are more chains than colours -- the reference indexes past its
list there), histograms with ``len(d) // 10`` bins normalised to a density and filled at alpha 0.5, traces
against the retained-row index, scatter points at alpha 0.1, ``rhat`` from ``diagnosticAssessment.csv`` in
the titles.  Reads the samples through ``sampleDiagnosis.openSamples`` (CSV files or the binary store)."""

import csv
import glob
import os
import struct
import zlib

import numpy

COLOURS = {"blue": (31, 60, 230), "red": (220, 30, 30), "green": (20, 140, 40), "magenta": (200, 30, 200),
           "cyan": (20, 180, 190), "yellow": (200, 180, 20), "black": (0, 0, 0)}
CHAIN_COLOURS = ["blue", "red", "green", "magenta", "cyan", "yellow", "black"]      # :519-520
DPI = 100                                                                            # matplotlib's default: figsize inches -> pixels
MAX_CHAINS = int(os.environ.get("MCMCN_FIGURE_MAX_CHAINS", "56"))                    # chains drawn per figure (8 per colour)

# 5 x 7 typeface, ASCII 32..126: five column bytes per glyph, bit 0 = top row (bit 7: descenders)
_GLYPHS = bytes.fromhex(
    "0000000000" "00005f0000" "0007000700" "147f147f14" "242a7f2a12" "2313086462" "3649562050" "0008070300"
    "001c224100" "0041221c00" "2a1c7f1c2a" "08083e0808" "0080703000" "0808080808" "0000606000" "2010080402"
    "3e5149453e" "00427f4000" "7249494946" "2141494d33" "1814127f10" "2745454539" "3c4a494931" "4121110907"
    "3649494936" "464949291e" "0000140000" "0040340000" "0008142241" "1414141414" "0041221408" "0201590906"
    "3e415d594e" "7c1211127c" "7f49494936" "3e41414122" "7f4141413e" "7f49494941" "7f09090901" "3e41415173"
    "7f0808087f" "00417f4100" "2040413f01" "7f08142241" "7f40404040" "7f021c027f" "7f0408107f" "3e4141413e"
    "7f09090906" "3e4151215e" "7f09192946" "2649494932" "03017f0103" "3f4040403f" "1f2040201f" "3f4038403f"
    "6314081463" "0304780403" "6159494d43" "007f414141" "0204081020" "004141417f" "0402010204" "4040404040"
    "0003070800" "2054547840" "7f28444438" "3844444428" "384444287f" "3854545418" "00087e0902" "18a4a49c78"
    "7f08040478" "00447d4000" "2040403d00" "7f10284400" "00417f4000" "7c04780478" "7c08040478" "3844444438"
    "fc18242418" "18242418fc" "7c08040408" "4854545424" "04043f4424" "3c4040207c" "1c2040201c" "3c4030403c"
    "4428102844" "4c9090907c" "4464544c44" "0008364100" "0000770000" "0041360800" "0201020402")


def _glyphBits():
    cols = numpy.frombuffer(_GLYPHS, dtype=numpy.uint8).reshape(95, 5)
    bits = (cols[:, :, None] >> numpy.arange(8)[None, None, :]) & 1                  # [glyph][column][row]
    return numpy.transpose(bits, (0, 2, 1)).astype(bool)                             # [glyph][row][column]


_BITS = _glyphBits()


class Canvas(object):
    """RGB raster, origin top-left, white."""

    def __init__(self, width, height):
        self.w, self.h = int(width), int(height)
        self.px = numpy.full((self.h, self.w, 3), 255, dtype=numpy.float32)

    def _blend(self, ys, xs, colour, alpha):
        c = numpy.asarray(colour, dtype=numpy.float32)
        a = numpy.asarray(alpha, dtype=numpy.float32)
        if a.ndim:
            a = a[..., None]
        self.px[ys, xs] = self.px[ys, xs] * (1.0 - a) + c * a

    def fillRect(self, x0, y0, x1, y1, colour, alpha=1.0):
        """Pixels x0 <= x < x1, y0 <= y < y1 (clipped)."""
        x0, x1 = max(0, int(x0)), min(self.w, int(x1))
        y0, y1 = max(0, int(y0)), min(self.h, int(y1))
        if x1 > x0 and y1 > y0:
            self._blend(slice(y0, y1), slice(x0, x1), colour, alpha)

    def frame(self, x0, y0, x1, y1, colour=(0, 0, 0)):
        self.fillRect(x0, y0, x1 + 1, y0 + 1, colour)
        self.fillRect(x0, y1, x1 + 1, y1 + 1, colour)
        self.fillRect(x0, y0, x0 + 1, y1 + 1, colour)
        self.fillRect(x1, y0, x1 + 1, y1 + 1, colour)

    def polylines(self, x, y, colour, clip):
        """Polylines through pixel coordinates x, y ([lines][points] or [points]); every segment is sampled
        once per pixel of its longer extent.  clip = (x0, y0, x1, y1), inclusive."""
        x = numpy.atleast_2d(numpy.asarray(x, dtype=numpy.float64))
        y = numpy.atleast_2d(numpy.asarray(y, dtype=numpy.float64))
        if x.shape[1] == 1:
            x, y = numpy.repeat(x, 2, axis=1), numpy.repeat(y, 2, axis=1)
        ax, ay = x[:, :-1].ravel(), y[:, :-1].ravel()
        dx, dy = (x[:, 1:] - x[:, :-1]).ravel(), (y[:, 1:] - y[:, :-1]).ravel()
        ok = numpy.isfinite(ax + ay + dx + dy)
        ax, ay, dx, dy = ax[ok], ay[ok], dx[ok], dy[ok]
        if ax.size == 0:
            return
        steps = numpy.minimum(numpy.maximum(numpy.abs(dx), numpy.abs(dy)), 4.0 * (self.w + self.h)).astype(numpy.int64) + 1
        seg = numpy.repeat(numpy.arange(steps.size), steps)
        t = (numpy.arange(seg.size) - numpy.repeat(numpy.cumsum(steps) - steps, steps)) / numpy.repeat(steps, steps)
        ends = numpy.isfinite(x[:, -1] + y[:, -1])                                      # segments are [start, end): the last point
        self.points(numpy.concatenate([ax[seg] + dx[seg] * t, x[ends, -1]]),
                    numpy.concatenate([ay[seg] + dy[seg] * t, y[ends, -1]]), colour, 1.0, 1, clip)

    def points(self, x, y, colour, alpha, size, clip):
        """Squares of `size` pixels at (x, y); n overlapping points of one call cover 1 - (1 - alpha)^n."""
        xi = numpy.rint(numpy.asarray(x, dtype=numpy.float64)).astype(numpy.int64).ravel()
        yi = numpy.rint(numpy.asarray(y, dtype=numpy.float64)).astype(numpy.int64).ravel()
        cx0, cy0, cx1, cy1 = clip
        cx0, cy0, cx1, cy1 = max(cx0, 0), max(cy0, 0), min(cx1, self.w - 1), min(cy1, self.h - 1)
        if cx1 < cx0 or cy1 < cy0:
            return
        bw = cx1 - cx0 + 1
        count = numpy.zeros((cy1 - cy0 + 1) * bw, dtype=numpy.int64)                 # the clip box only
        lo = -(size // 2)
        for oy in range(lo, lo + size):
            for ox in range(lo, lo + size):
                xs, ys = xi + ox, yi + oy
                ok = (xs >= cx0) & (xs <= cx1) & (ys >= cy0) & (ys <= cy1)
                count += numpy.bincount((ys[ok] - cy0) * bw + (xs[ok] - cx0), minlength=count.size)
        hit = numpy.nonzero(count)[0]
        if hit.size:
            cover = 1.0 - (1.0 - alpha) ** count[hit] if alpha < 1.0 else 1.0
            self._blend(hit // bw + cy0, hit % bw + cx0, colour, cover)

    def text(self, x, y, s, colour=(0, 0, 0), scale=1, anchor="l", vertical=False):
        """Top-left of the string at (x, y); anchor "c" / "r" centres / right-aligns it on x.  vertical: rotated
        by 90 degrees (reading bottom to top), anchored the same way along y."""
        codes = [ord(c) - 32 if 32 <= ord(c) <= 126 else ord("?") - 32 for c in str(s)]
        if not codes:
            return
        strip = numpy.zeros((8, 6 * len(codes)), dtype=bool)
        for i, g in enumerate(codes):
            strip[:, 6 * i:6 * i + 5] = _BITS[g]
        if scale > 1:
            strip = numpy.repeat(numpy.repeat(strip, scale, axis=0), scale, axis=1)
        if vertical:
            strip = numpy.rot90(strip)
            extent = strip.shape[0]
            y = y - (extent // 2 if anchor == "c" else extent if anchor == "r" else 0)
        else:
            extent = strip.shape[1]
            x = x - (extent // 2 if anchor == "c" else extent if anchor == "r" else 0)
        ys, xs = numpy.nonzero(strip)
        xs, ys = xs + int(x), ys + int(y)
        ok = (xs >= 0) & (xs < self.w) & (ys >= 0) & (ys < self.h)
        self._blend(ys[ok], xs[ok], colour, 1.0)

    def png(self):
        """The canvas as PNG bytes (8-bit RGB, one zlib stream, filter 0 on every row)."""
        rgb = numpy.clip(numpy.rint(self.px), 0, 255).astype(numpy.uint8)
        raw = numpy.concatenate([numpy.zeros((self.h, 1), dtype=numpy.uint8), rgb.reshape(self.h, self.w * 3)], axis=1)

        def chunk(tag, body):
            return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)
        return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", self.w, self.h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)) + chunk(b"IEND", b""))

    def save(self, path):
        with open(path, "wb") as h:
            h.write(self.png())


def readPng(path):
    """[height][width][3] uint8 of a PNG written by Canvas.save (tests; not a general decoder)."""
    with open(path, "rb") as h:
        blob = h.read()
    if blob[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file: %s" % path)
    pos, data, w, hgt = 8, b"", 0, 0
    while pos < len(blob):
        n, tag = struct.unpack(">I4s", blob[pos:pos + 8])
        body = blob[pos + 8:pos + 8 + n]
        if zlib.crc32(tag + body) & 0xFFFFFFFF != struct.unpack(">I", blob[pos + 8 + n:pos + 12 + n])[0]:
            raise ValueError("bad chunk checksum in %s" % path)
        if tag == b"IHDR":
            w, hgt = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            data += body
        pos += 12 + n
    raw = numpy.frombuffer(zlib.decompress(data), dtype=numpy.uint8).reshape(hgt, 1 + 3 * w)
    return raw[:, 1:].reshape(hgt, w, 3)


def niceTicks(lo, hi, target=5):
    """Round tick values (1, 2, 2.5, 5 x 10^k apart) inside [lo, hi]."""
    if not (numpy.isfinite(lo) and numpy.isfinite(hi)) or hi <= lo:
        return [lo]
    raw = (hi - lo) / max(target, 1)
    mag = 10.0 ** numpy.floor(numpy.log10(raw))
    step = min((m for m in (1.0, 2.0, 2.5, 5.0, 10.0) if m * mag >= raw), default=10.0) * mag
    first = numpy.ceil(lo / step - 1e-9)
    ticks = [(first + i) * step for i in range(int(numpy.floor(hi / step + 1e-9) - first) + 1)]
    return [0.0 if abs(t) < step * 1e-9 else t for t in ticks]


class Axes(object):
    """A plot box on a canvas: data -> pixel mapping with 5 % margins, frame, ticks, title and labels."""

    def __init__(self, canvas, x0, y0, x1, y1):
        self.c = canvas
        self.box = (int(x0), int(y0), int(x1), int(y1))
        self.xr, self.yr = (0.0, 1.0), (0.0, 1.0)

    @staticmethod
    def _padded(lo, hi):
        lo, hi = float(lo), float(hi)
        if not (numpy.isfinite(lo) and numpy.isfinite(hi)):
            return 0.0, 1.0
        if hi <= lo:
            pad = 0.5 if lo == 0 else abs(lo) * 0.05
            return lo - pad, hi + pad
        pad = 0.05 * (hi - lo)
        return lo - pad, hi + pad

    def limits(self, xlo, xhi, ylo, yhi, padY=True):
        self.xr = self._padded(xlo, xhi)
        self.yr = self._padded(ylo, yhi) if padY else (float(ylo), float(yhi) * 1.05 if yhi > ylo else float(ylo) + 1.0)

    def X(self, v):
        x0, _, x1, _ = self.box
        return x0 + (numpy.asarray(v, dtype=numpy.float64) - self.xr[0]) / (self.xr[1] - self.xr[0]) * (x1 - x0)

    def Y(self, v):
        _, y0, _, y1 = self.box
        return y1 - (numpy.asarray(v, dtype=numpy.float64) - self.yr[0]) / (self.yr[1] - self.yr[0]) * (y1 - y0)

    def decorate(self, title=None, xlabel=None, ylabel=None, yticks=True):
        x0, y0, x1, y1 = self.box
        self.c.frame(x0, y0, x1, y1)
        for t in niceTicks(*self.xr):
            px = int(round(float(self.X(t))))
            self.c.fillRect(px, y1, px + 1, y1 + 4, (0, 0, 0))
            self.c.text(px, y1 + 6, "%g" % t, anchor="c")
        if yticks:
            for t in niceTicks(*self.yr):
                py = int(round(float(self.Y(t))))
                self.c.fillRect(x0 - 3, py, x0, py + 1, (0, 0, 0))
                self.c.text(x0 - 6, py - 3, "%g" % t, anchor="r")
        if title:
            self.c.text((x0 + x1) // 2, y0 - 14, title, anchor="c")
        if xlabel:
            self.c.text((x0 + x1) // 2, y1 + 18, xlabel, anchor="c")
        if ylabel:
            self.c.text(x0 - (56 if yticks else 14), (y0 + y1) // 2, ylabel, anchor="c", vertical=True)

    def lines(self, x, y, colour):
        self.c.polylines(self.X(x), self.Y(y), colour, self.box)

    def scatter(self, x, y, colour, alpha=0.1, size=2):
        self.c.points(self.X(x), self.Y(y), colour, alpha, size, self.box)

    def bars(self, edges, heights, colour, alpha=0.5):
        """Filled step histogram (histtype="stepfilled"): per pixel column of the box the height of the bin it falls
        in, one blend for the whole histogram (a bin narrower than a pixel still shows: the tallest bin of a column wins)."""
        bx0, by0, bx1, by1 = self.box
        xs = numpy.rint(self.X(edges)).astype(numpy.int64)
        tops = numpy.clip(numpy.rint(self.Y(heights)).astype(numpy.int64), by0, by1)
        base = min(int(round(float(self.Y(0.0)))), by1)
        top = numpy.full(bx1 - bx0 + 1, base + 1, dtype=numpy.int64)          # per column: first filled row (none: below the base)
        lo = numpy.clip(xs[:-1], bx0, bx1 + 1) - bx0
        hi = numpy.clip(numpy.maximum(xs[1:], xs[:-1] + 1), bx0, bx1 + 1) - bx0
        for i in numpy.nonzero(hi > lo)[0]:
            numpy.minimum(top[lo[i]:hi[i]], min(tops[i], base), out=top[lo[i]:hi[i]])
        rows = numpy.arange(by0, base + 1)[:, None]
        ys, xcol = numpy.nonzero(rows >= top[None, :])
        if ys.size:
            self.c._blend(ys + by0, xcol + bx0, colour, alpha)


def _colour(i):
    return COLOURS[CHAIN_COLOURS[i % len(CHAIN_COLOURS)]]


class Figure(object):
    """Figures to assess mixing and convergence (reference ``Figure``, sampleDiagnosis.py:494-759): same
    constructor argument, method names and output file names."""

    def __init__(self, sampleDirectory, source=None):
        import sampleDiagnosis
        self._src = source if source is not None else sampleDiagnosis.openSamples(sampleDirectory)
        self._keys = list(self._src.keys)
        # Chains drawn: all of them up to MAX_CHAINS (MCMCN_FIGURE_MAX_CHAINS), else that many evenly spaced ones -- a
        # panel of a thousand overlaid traces shows nothing a panel of 56 does not (the reference stops at 7 chains)
        nAll = self._src.nChains
        self._chains = numpy.arange(nAll) if nAll <= MAX_CHAINS else \
            numpy.unique(numpy.linspace(0, nAll - 1, MAX_CHAINS).round().astype(numpy.int64))
        self._m = len(self._chains)
        self._n = self._src.nRows
        self._column = {k: i for i, k in enumerate(self._keys)}
        self._rhat = self._loadSummary(sampleDirectory)
        self._loglikelihoods = self._loadLogLikelihoods(sampleDirectory)
        suffices = sorted(set("[" + name.split("[")[1] for name in self._keys if "[" in name))      # :512-513
        if any("_" in name for name in self._keys):                                                 # :515-517
            suffices = ["_"] + suffices
        self._keySuffices = suffices

    # ------------------------------------------------------------------ data
    def _samples(self, keys):
        """float64 [len(keys)][rows][chains] of the named columns (host): a few columns of every block."""
        cols = [self._column[k] for k in keys]
        parts = []
        for arr, ids in self._src.blocks:
            if hasattr(arr, "cpu"):                                   # device-resident store
                import torch
                blk = arr[:self._n].index_select(1, torch.as_tensor(cols, device=arr.device))[:, :, :len(ids)].cpu().numpy()
            else:
                blk = numpy.asarray(arr[:self._n][:, cols, :len(ids)])
            parts.append(numpy.asarray(blk, dtype=numpy.float64))
        data = numpy.concatenate(parts, axis=2)
        if self._m < data.shape[2]:
            data = data[:, :, self._chains]
        return numpy.transpose(data, (1, 0, 2))

    @staticmethod
    def _loadSummary(sampleDirectory):
        """rhat per parameter from diagnostic/diagnosticAssessment.csv, when diagnoseSamples wrote it (:545-550)."""
        path = os.path.join(sampleDirectory, "..", "diagnostic", "diagnosticAssessment.csv")
        if not os.path.exists(path):
            return None
        with open(path, newline="") as h:
            rows = list(csv.reader(h, quotechar="'"))
        head = [c.strip() for c in rows[0]]
        ip, ir = head.index("parameter"), head.index("rhat")
        return {r[ip].strip(): float(r[ir]) for r in rows[1:] if len(r) > max(ip, ir)}

    @staticmethod
    def _loadLogLikelihoods(sampleDirectory):
        """Per chain, the log-likelihood of every retained row: the row sums of logLikelihood.<chain>.csv (:552-570)."""
        files = glob.glob(os.path.join(sampleDirectory, "logLikelihood*.csv"))
        if not files or sum(os.path.getsize(f) for f in files) == 0:
            return None

        def chainOf(path):
            parts = os.path.basename(path).split(".")
            return int(parts[1]) if len(parts) > 2 and parts[1].isdigit() else 1 << 30
        files.sort(key=lambda f: (chainOf(f), f))
        return [numpy.atleast_2d(numpy.loadtxt(f, delimiter=",", dtype=numpy.float64)).sum(axis=1) for f in files]

    # ------------------------------------------------------------------ plots
    @staticmethod
    def _finish(canvas, dest):
        if dest is not None:
            canvas.save(dest)
        return canvas

    def loglikelihood(self, dest):
        """One line per chain: total log-likelihood against the retained-row index (:572-593)."""
        if self._loglikelihoods is None:
            return None
        print("Creating loglikelihood plot", end="")
        cv = Canvas(12 * DPI, 3 * DPI)
        ax = Axes(cv, 80, 20, cv.w - 110, cv.h - 45)
        series = self._loglikelihoods
        finite = numpy.concatenate([s[numpy.isfinite(s)] for s in series])
        ax.limits(0, max(len(s) for s in series) - 1, finite.min() if finite.size else 0.0, finite.max() if finite.size else 1.0)
        for i, s in enumerate(series):
            ax.lines(numpy.arange(len(s)), s, _colour(i))
        ax.decorate(xlabel="Iteration", ylabel="Log Likelihood")
        lx = cv.w - 100                                               # legend, title "Chain" (:583)
        cv.text(lx, 22, "Chain")
        for i in range(min(len(series), 12)):
            cv.fillRect(lx, 38 + 12 * i, lx + 14, 40 + 12 * i, _colour(i))
            cv.text(lx + 20, 35 + 12 * i, str(i))
        print(": Done")
        return self._finish(cv, dest)

    def _progress(self, what, i, n):
        msg = "Creating %s: %i out of %i" % (what, i + 1, n)
        if i:
            print("\r" * len(msg), end="")
        print(msg, end="")
        return msg

    @staticmethod
    def _progressDone(what, last):
        print("\r" * len(last) + " " * len(last) + "\r" * len(last), end="")
        print("Creating %s: Done." % what)

    def traceplots(self, dest, n=30):
        """traceplot<suffix>.png for the first n key suffices: the hyper-parameters ("_"), then one figure per
        group with every parameter of that group (:595-623)."""
        last = ""
        for i, suffix in enumerate(self._keySuffices[:n]):
            last = self._progress("traceplots", i, n)
            self.traceplot(sorted(k for k in self._keys if suffix in k), dest + "/traceplot%s.png" % suffix)
        self._progressDone("traceplots", last)

    def traceplot(self, keys, dest):
        """Per key a row of two panels: the chains' histograms (density, alpha 0.5) and their traces (:625-688)."""
        if keys is None:
            keys = sorted(k for k in self._keys if "_mean" in k)
        n = len(keys)
        if n == 0:
            return None
        data = self._samples(keys)                                    # [keys][rows][chains]
        cv = Canvas(12 * DPI, 2 * DPI * n)
        half = cv.w // 2
        for i, key in enumerate(keys):
            d = data[i]
            title = key.replace("_", " ")
            if self._rhat is not None and key in self._rhat:
                title += " (rhat=%.3f)" % self._rhat[key]
            top = 2 * DPI * i
            finite = d[numpy.isfinite(d)]
            lo, hi = (finite.min(), finite.max()) if finite.size else (0.0, 1.0)
            # histogram panel: every chain binned over its own range, len(d) // 10 bins (:678-684)
            hx = Axes(cv, 40, top + 30, half - 30, top + 2 * DPI - 45)
            hists = []
            nb = max(d.shape[0] // 10, 1)
            for c in range(d.shape[1]):
                col = d[:, c][numpy.isfinite(d[:, c])]
                if col.size == 0:
                    continue
                cl, ch = col.min(), col.max()
                if ch <= cl:
                    cl, ch = cl - 0.5, ch + 0.5
                dens, edges = numpy.histogram(col, bins=nb, range=(cl, ch), density=True)
                hists.append((c, edges, dens))
            hx.limits(lo, hi, 0.0, max([h[2].max() for h in hists] + [1e-300]), padY=False)
            for c, edges, dens in hists:
                hx.bars(edges, dens, _colour(c), 0.5)
            hx.decorate(title=title, xlabel="Sample Value", ylabel="Log Density", yticks=False)
            # trace panel
            tx = Axes(cv, half + 70, top + 30, cv.w - 20, top + 2 * DPI - 45)
            tx.limits(0, max(d.shape[0] - 1, 1), lo, hi)
            xs = numpy.arange(d.shape[0])
            for k in range(min(len(CHAIN_COLOURS), d.shape[1])):
                cols = d[:, k::len(CHAIN_COLOURS)].T                  # every chain of colour k, drawn in one call
                tx.lines(numpy.broadcast_to(xs, cols.shape), cols, _colour(k))
            tx.decorate(title=title, xlabel="Iteration", ylabel="Sample Value")
        return self._finish(cv, dest)

    def bivariates(self, dest, n=30):
        """bivariate<suffix>.png for the first n key suffices (:690-720)."""
        last = ""
        for i, suffix in enumerate(self._keySuffices[:n]):
            last = self._progress("bivariate plots", i, n)
            self.bivariate(sorted(k for k in self._keys if suffix in k), dest + "/bivariate%s.png" % suffix)
        self._progressDone("bivariate plots", last)

    def bivariate(self, keys, dest):
        """n x n panels: the name on the diagonal, elsewhere key x against key y, one colour per chain, alpha 0.1
        (:722-759).  Nothing for a single key."""
        if keys is None:
            keys = sorted(k for k in self._keys if "_mean" in k)
        n = len(keys)
        if n <= 1:
            return 0
        data = self._samples(keys)
        cell = 2 * DPI
        cv = Canvas(cell * n, cell * n)
        for i in range(n):
            for j in range(n):
                x0, y0 = cell * j, cell * i
                if i == j:
                    cv.text(x0 + cell // 2, y0 + cell // 2 - 4, keys[j].replace("_", " "), anchor="c")
                    continue
                ax = Axes(cv, x0 + 45, y0 + 12, x0 + cell - 10, y0 + cell - 28)
                dx, dy = data[j], data[i]
                fx, fy = dx[numpy.isfinite(dx)], dy[numpy.isfinite(dy)]
                ax.limits(fx.min() if fx.size else 0, fx.max() if fx.size else 1, fy.min() if fy.size else 0, fy.max() if fy.size else 1)
                for k in range(min(len(CHAIN_COLOURS), dx.shape[1])):
                    ax.scatter(dx[:, k::len(CHAIN_COLOURS)], dy[:, k::len(CHAIN_COLOURS)], _colour(k), 0.1, 3)
                ax.decorate()
        return self._finish(cv, dest)
