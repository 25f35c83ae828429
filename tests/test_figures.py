"""figures.py (reference: Figure, sampleDiagnosis.py:494-759): the rasteriser, the PNG encoder and the
reference's figure set on golden sample files.  Host only -- no device needed."""
import os
import shutil

import numpy
import pytest

from conftest import GOLDEN


def _stage(case, tmp_path, diagnostic=True):
    out = tmp_path / case
    (out / "sample").mkdir(parents=True)
    for f in os.listdir(os.path.join(GOLDEN, case)):
        if f.startswith("sample.") or f.startswith("logLikelihood."):
            shutil.copy(os.path.join(GOLDEN, case, f), out / "sample" / f)
    if diagnostic:
        (out / "diagnostic").mkdir()
        shutil.copy(os.path.join(GOLDEN, case, "diagnosticAssessment.csv"), out / "diagnostic")
    return str(out)


def _has(img, colour, tol=40):
    return bool((numpy.abs(img.astype(int) - numpy.array(colour)).sum(axis=2) < tol).any())


def test_png_round_trip_and_primitives(tmp_path):
    import figures
    cv = figures.Canvas(64, 48)
    cv.fillRect(4, 4, 20, 10, (220, 30, 30))
    cv.fillRect(10, 6, 30, 20, (31, 60, 230), 0.5)                   # half-transparent over red and over white
    cv.polylines([2, 60], [40, 40], (0, 0, 0), (0, 0, 63, 47))
    cv.text(2, 30, "I")
    path = str(tmp_path / "c.png")
    cv.save(path)
    img = figures.readPng(path)
    assert img.shape == (48, 64, 3)
    assert tuple(img[5, 5]) == (220, 30, 30)
    assert tuple(img[7, 12]) == (126, 45, 130)                      # (red + blue) / 2, rounded to even
    assert tuple(img[15, 25]) == (143, 158, 242)                    # (white + blue) / 2
    assert (img[40, 2:61] == 0).all() and (img[41, 2:61] == 255).all()
    # "I": top bar, stem in the middle column, bottom bar (5 x 7 glyph)
    glyph = (img[30:37, 2:7].sum(axis=2) == 0)
    assert glyph[0, 1:4].all() and glyph[6, 1:4].all() and glyph[1:6, 2].all() and not glyph[3, 0] and not glyph[3, 4]
    with open(path, "rb") as h:
        assert h.read(8) == b"\x89PNG\r\n\x1a\n"


def test_point_clouds_accumulate_alpha():
    import figures
    cv = figures.Canvas(10, 10)
    cv.points([3, 3, 3], [4, 4, 4], (0, 0, 0), 0.5, 1, (0, 0, 9, 9))       # three points on one pixel: 1 - 0.5^3
    assert abs(cv.px[4, 3, 0] - 255 * 0.125) < 1e-3 and cv.px[4, 4, 0] == 255
    cv.points([20, -3], [4, 4], (0, 0, 0), 1.0, 1, (0, 0, 9, 9))           # outside the clip box: nothing
    assert (cv.px[:, :, 0] < 255).sum() == 1


def test_nice_ticks():
    import figures
    assert figures.niceTicks(0.0, 100.0) == [0.0, 20.0, 40.0, 60.0, 80.0, 100.0]
    assert figures.niceTicks(-0.6, 1.5) == [-0.5, 0.0, 0.5, 1.0, 1.5]
    t = figures.niceTicks(3.2e-5, 9.7e-5)
    assert all(3.2e-5 <= v <= 9.7e-5 for v in t) and 3 <= len(t) <= 8
    assert figures.niceTicks(2.0, 2.0) == [2.0]


def test_figure_set_of_a_partial_pooling_run(tmp_path, capsys):
    """The reference's files (:73-85): logLikelihood.png, traceplot<suffix>.png, bivariate<suffix>.png for the
    hyper-parameters ("_") and the first groups; sizes follow its figsize (12 x 2n and 2n x 2n inches at 100 dpi)."""
    import figures
    out = _stage("c1_distribution_partial", tmp_path)
    fig = figures.Figure(out + "/sample/")
    assert fig._keySuffices[:3] == ["_", "[000]", "[001]"] and fig._m == 2 and fig._n == 100
    os.makedirs(out + "/figure/traceplot")
    os.makedirs(out + "/figure/bivariate")
    fig.loglikelihood(out + "/figure/logLikelihood.png")
    fig.traceplots(out + "/figure/traceplot", 3)
    fig.bivariates(out + "/figure/bivariate", 3)
    printed = capsys.readouterr().out
    assert "Creating loglikelihood plot: Done" in printed
    assert printed.rstrip().endswith("Creating bivariate plots: Done.") and "Creating traceplots: Done." in printed
    assert sorted(os.listdir(out + "/figure/traceplot")) == ["traceplot[000].png", "traceplot[001].png", "traceplot_.png"]
    assert sorted(os.listdir(out + "/figure/bivariate")) == ["bivariate[000].png", "bivariate[001].png", "bivariate_.png"]
    blue, red = figures.COLOURS["blue"], figures.COLOURS["red"]
    ll = figures.readPng(out + "/figure/logLikelihood.png")
    assert ll.shape == (300, 1200, 3) and _has(ll, blue) and not _has(ll, red)          # only chain 0 saved its log-likelihood
    tr = figures.readPng(out + "/figure/traceplot/traceplot_.png")                        # a, b, c x (mu, sigma2)
    assert tr.shape == (1200, 1200, 3) and _has(tr[:, 600:], blue) and _has(tr[:, 600:], red)
    assert _has(tr[:, :600], (143, 158, 242), 12)                                        # histogram bars: blue at alpha 0.5
    assert (tr[10:24, 100:500].sum(axis=2) == 0).any()                                   # title text above the first panel
    g0 = figures.readPng(out + "/figure/traceplot/traceplot[000].png")                   # a[000], b[000], c[000]
    assert g0.shape == (600, 1200, 3)
    bv = figures.readPng(out + "/figure/bivariate/bivariate[000].png")
    assert bv.shape == (600, 600, 3)
    off = bv[0:200, 200:400]                                                             # panel (a, b): tinted points
    assert ((off[:, :, 2] > off[:, :, 0] + 10).any() and (off[:, :, 0] > off[:, :, 2] + 10).any())
    diag = bv[0:200, 0:200]                                                              # diagonal: the name only, no frame
    assert (diag.sum(axis=2) == 0).any() and not (diag[12] == 0).all(axis=1).sum() > 100


def test_figure_traces_follow_the_draws(tmp_path):
    """A trace panel is the chain's draws: the drawn pixels of one chain lie on the polyline through its values."""
    import figures
    import sampleDiagnosis as sd
    n, m = 40, 1
    x = numpy.linspace(0.0, 1.0, n)
    samples = numpy.stack([x, 1.0 - x], axis=1)[None]                                    # [chains][rows][keys]: up, down
    src = sd.SampleSource.fromArray(samples, ["up[000]", "down[000]"])
    fig = figures.Figure(str(tmp_path), source=src)
    assert fig._keySuffices == ["[000]"]
    cv = fig.traceplot(["up[000]", "down[000]"], None)
    img = numpy.rint(cv.px).astype(int)
    blue = numpy.abs(img - numpy.array(figures.COLOURS["blue"])).sum(axis=2) < 10
    for row0, rising in ((0, True), (200, False)):
        ys, xs = numpy.nonzero(blue[row0:row0 + 200, 670:1181])                            # the trace panel of this key
        assert xs.size > 300
        slope = numpy.polyfit(xs, ys, 1)[0]
        assert (slope < -0.1) if rising else (slope > 0.1)                               # pixel y grows downwards
    assert fig.bivariate(["up[000]"], None) == 0                                         # a single key: nothing (:731-732)
    assert fig.loglikelihood(str(tmp_path / "ll.png")) is None and not os.path.exists(str(tmp_path / "ll.png"))


def test_figures_read_the_binary_store(tmp_path):
    """manifest.json + .npy shard files (the store samplePosterior writes at scale) instead of CSV files; more chains
    than colours cycle through the seven."""
    import json
    import figures
    rs = numpy.random.RandomState(5)
    rows, chains = 60, 9
    header = ["b_mu", "b_sigma2", "b[000]", "b[001]"]
    a = rs.normal(size=(rows, len(header), 5)).astype(numpy.float32)
    b = rs.normal(size=(rows, len(header), 4)).astype(numpy.float32)
    sdir = tmp_path / "sample"
    sdir.mkdir()
    numpy.save(str(sdir / "samples.rank0.npy"), a)
    numpy.save(str(sdir / "samples.rank1.npy"), b)
    with open(str(sdir / "manifest.json"), "w") as h:
        json.dump({"format": "mcmcn-samples-2", "header": header, "iterations": list(range(rows)), "nChains": chains,
                   "pooling": "partial", "dtype": "float32", "layout": "[rows][columns][chains]",
                   "shards": [{"file": "samples.rank0.npy", "chains": [0, 5]}, {"file": "samples.rank1.npy", "chains": [5, 9]}]}, h)
    fig = figures.Figure(str(sdir) + "/")
    assert fig._m == chains and fig._n == rows and fig._keySuffices == ["_", "[000]", "[001]"]
    d = fig._samples(["b[001]", "b_mu"])
    assert d.shape == (2, rows, chains)
    numpy.testing.assert_array_equal(d[0][:, :5], a[:, 3, :])
    numpy.testing.assert_array_equal(d[1][:, 5:], b[:, 0, :])
    cv = fig.traceplot(["b_mu", "b_sigma2"], str(tmp_path / "t.png"))
    img = figures.readPng(str(tmp_path / "t.png"))
    assert img.shape == (400, 1200, 3)
    for name in figures.CHAIN_COLOURS:
        assert _has(img[:, 600:], figures.COLOURS[name], 10), name


def test_many_chains_are_thinned_to_evenly_spaced_ones(monkeypatch):
    """More chains than MCMCN_FIGURE_MAX_CHAINS: that many evenly spaced chains are drawn, first and last included."""
    import figures
    import sampleDiagnosis as sd
    rs = numpy.random.RandomState(1)
    samples = rs.normal(size=(40, 30, 2))
    samples[:, :, 0] += numpy.arange(40)[:, None]                                        # chain c sits at level c
    src = sd.SampleSource.fromArray(samples, ["a[000]", "b[000]"])
    monkeypatch.setattr(figures, "MAX_CHAINS", 5)
    fig = figures.Figure("/nonexistent/", source=src)
    assert fig._m == 5 and list(fig._chains) == [0, 10, 20, 29, 39]
    d = fig._samples(["a[000]"])
    assert d.shape == (1, 30, 5)
    numpy.testing.assert_array_equal(d[0], samples[[0, 10, 20, 29, 39], :, 0].T)
    cv = fig.traceplot(["a[000]", "b[000]"], None)
    assert cv.px.shape == (400, 1200, 3)
    monkeypatch.setattr(figures, "MAX_CHAINS", 56)
    assert figures.Figure("/nonexistent/", source=src)._m == 40
