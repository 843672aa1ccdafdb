"""Host side of camera creation (zoic_b200/csrc/host_setup.cpp through the C ABI's host-only entry point)
against the oracle: every derived constant, the exit-pupil LUT boxes and the bokeh CDF tables, bit for bit; the
lens-table grammar; the error codes.  No GPU needed (LUT candidates are classified by host threads here)."""
import os

import numpy as np
import pytest

from zutil import (GOLDEN, bits_equal, draw_zoic_header, golden_case, golden_names, product_constants_flat,
                   setup_log_values)

from zoic_b200 import capi, host_setup
from zoic_b200.synth import hex_bokeh_image
from zoic_b200.workloads import LENSES, lens_path


def _same_constants(a, b, nan_is_nan=False):
    """nan_is_nan: a NaN equals any NaN (a lens table mangled by the delimiter-run quirk scales to NaNs, whose sign
    depends on the compiler's instruction selection and means nothing)."""
    a = product_constants_flat(a)

    def same(x, y):
        x, y = np.asarray(x, np.float32), np.asarray(y, np.float32)
        if nan_is_nan:
            n = np.isnan(x)
            return x.shape == y.shape and np.array_equal(n, np.isnan(y)) and bits_equal(x[~n], y[~n])
        return bits_equal(x, y)
    for k in ("userApertureRadius", "originShift", "apertureDistance", "focalLengthRatio", "tracedFocalLength0",
              "tracedFocalLength1", "principalPlane0", "principalPlane1", "focalPoint0", "focalPoint1"):
        assert same([a[k]], [b[k]]), k
    assert a["lensCount"] == b["lensCount"] and a["apertureElement"] == b["apertureElement"]
    assert same(a["lenses"], b["lenses"])
    assert same(a["lut"], b["lut"])


@pytest.mark.parametrize("lens", sorted(LENSES))
def test_kolb_setup_matches_oracle_bit_for_bit(port, lens):
    fnum, focal = LENSES[lens]
    kw = dict(lensModel=1, lensDataPath=lens_path(lens), focalLength=focal, fStop=fnum)
    c, _ = host_setup(**kw)
    p = port.PortCamera(**kw)
    _same_constants(c, p.constants())
    assert c["lutSize"] == 32
    p.close()


def test_kolb_setup_other_parameters(port):
    for kw in (dict(focalLength=3.5, fStop=1.2, focalDistance=55.0), dict(focalLength=8.0, fStop=11.0, focalDistance=1000.0),
               dict(focalLength=5.0, fStop=2.8, focalDistance=23.0, kolbSamplingLUT=0)):
        kw = dict(dict(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat")), **kw)
        c, _ = host_setup(**kw)
        p = port.PortCamera(**kw)
        _same_constants(c, p.constants())
        p.close()


def test_setup_reproduces_reference_known_answers():
    """reference src/draw.zoic:1-10 and the node_update log lines of the compiled reference (golden.json)."""
    want = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "draw_zoic_header.txt"))]
    c, _ = host_setup(lensModel=1, lensDataPath=lens_path("double_gauss_f2.0.dat"), focalLength=5.0, fStop=2.8, focalDistance=23.0)
    assert draw_zoic_header(c, 23.0) == want
    for name in golden_names():
        if not name.startswith("kolb"):
            continue
        kw, image, meta, *_ = golden_case(name)
        c, _ = host_setup(image=image, **kw)
        got = setup_log_values(product_constants_flat(c), kw["fStop"])
        for key, val in meta["setup_log"].items():
            assert got[key] == val, (name, key)


def test_thin_lens_constants_and_bokeh_tables(port):
    for size in (255, 33):
        img = hex_bokeh_image(size)
        kw = dict(lensModel=0, focalLength=3.5, fStop=2.8, opticalVignettingDistance=2.0, useImage=1)
        c, tabs = host_setup(image=img, **kw)
        p = port.PortCamera(image=img, **kw)
        pc = p.constants()
        for k in ("fov", "tan_fov", "apertureRadius"):
            assert np.float32(c[k]).tobytes() == np.float32(pc[k]).tobytes()
        for a, b in zip(tabs, p.bokeh_tables()):
            assert np.array_equal(a, b)
        assert (c["bokehWidth"], c["bokehHeight"]) == (size, size)
        p.close()


def test_bokeh_tables_ragged_images(port):
    """Non-square, even-sized, 4-channel, flat (all ties) and single-row images."""
    rng = np.random.default_rng(5)
    imgs = [hex_bokeh_image(33)[:, :20].copy(), hex_bokeh_image(32), np.ones((7, 9, 3), np.float32),
            rng.random((1, 16, 4)).astype(np.float32), rng.random((16, 1, 3)).astype(np.float32)]
    for img in imgs:
        kw = dict(lensModel=0, focalLength=3.5, fStop=2.8, useImage=1)
        _, tabs = host_setup(image=img, **kw)
        p = port.PortCamera(image=img, **kw)
        for a, b in zip(tabs, p.bokeh_tables()):
            assert np.array_equal(a, b)
        p.close()


LENS_TEXT = {
    "tabs": "# c\n58.950\t7.520\t1.67\t50.4\n169.660\t0.240\t1.0\t50.4\n0\t9.000\t0\t34.2\n-28.990\t2.360\t1.603\t34.0\n-79.46\t72.228\t1.0\t40.0",
    "mixed-delims-5col": "## V\n\n74.062,18.55;1.611:58.8 31.6\n-114.427\t0.766\t0.0\t0.0\t31.6\n0 5.3 0 0 20\n55.173\t15.9\t1.611 58.8\t23.1\n",
    "crlf": "58.950\t7.520\t1.67\t50.4\r\n0\t9.000\t0\t34.2\r\n-79.46\t72.228\t1.0\t40.0\r\n",
    "trailing-delim": "58.950 7.520 1.67 50.4 \n0 9.000 0 34.2 \n-79.46 72.228 1.0 40.0 \n",
}


@pytest.mark.parametrize("name", sorted(LENS_TEXT))
def test_lens_table_grammar(port, tmp_path, name):
    path = tmp_path / (name + ".dat")
    path.write_bytes(LENS_TEXT[name].encode())
    kw = dict(lensModel=1, lensDataPath=str(path), focalLength=5.0, fStop=2.8, kolbSamplingLUT=0)
    c, _ = host_setup(**kw)
    p = port.PortCamera(**kw)
    _same_constants(c, p.constants())
    p.close()


def test_shipped_lens_tables_parse_like_the_reference_originals(port):
    """zoic_b200/data/lenses/*.dat are re-emitted tables (tools/import_lenses.py); where the reference tree is
    present, check that both spellings give identical element tables."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from import_lenses import NAMES
    ref_dir = "/root/reference/lenses_tabular"
    if not os.path.isdir(ref_dir):
        pytest.skip("reference tree not present")
    for src, (dst, _) in NAMES.items():
        kw = dict(lensModel=1, focalLength=LENSES[dst][1], fStop=2.8, kolbSamplingLUT=0)
        a, _ = host_setup(lensDataPath=os.path.join(ref_dir, src), **kw)
        b, _ = host_setup(lensDataPath=lens_path(dst), **kw)
        assert bits_equal(a["lenses"], b["lenses"]) and a["originShift"] == b["originShift"]


def test_error_codes(tmp_path):
    with pytest.raises(capi.ZoicError) as e:
        host_setup(lensModel=1, lensDataPath=str(tmp_path / "missing.dat"))
    assert e.value.code == capi.ERR_LENS_FILE
    with pytest.raises(capi.ZoicError) as e:
        host_setup(lensModel=1, lensDataPath="")
    assert e.value.code == capi.ERR_LENS_FILE
    three = tmp_path / "three.dat"
    three.write_text("1 2 3\n4 5 6\n")
    with pytest.raises(capi.ZoicError) as e:
        host_setup(lensModel=1, lensDataPath=str(three))
    assert e.value.code == capi.ERR_LENS_FILE and "fewer than 4" in str(e.value)
    six = tmp_path / "six.dat"
    six.write_text("1 2 3 4 5 6\n1 2 3 4 5 6\n")
    with pytest.raises(capi.ZoicError) as e:
        host_setup(lensModel=1, lensDataPath=str(six))
    assert e.value.code == capi.ERR_LENS_FILE and "more than 5" in str(e.value)
    two_stops = tmp_path / "two.dat"
    two_stops.write_text("50 5 1.6 40\n0 2 0 30\n0 2 0 30\n-50 60 1 40\n")
    with pytest.raises(capi.ZoicError) as e:
        host_setup(lensModel=1, lensDataPath=str(two_stops))
    assert e.value.code == capi.ERR_LENS_DATA
    junk = tmp_path / "junk.dat"
    junk.write_text("50 5 abc 40\n-50 60 1 40\n")
    with pytest.raises(capi.ZoicError) as e:
        host_setup(lensModel=1, lensDataPath=str(junk))
    assert e.value.code == capi.ERR_LENS_FILE
    with pytest.raises(capi.ZoicError) as e:
        host_setup(lensModel=0, useImage=1)
    assert e.value.code == capi.ERR_BOKEH_IMAGE
    # 1 or 2 channels: accepted like the reference accepts them -- as an invalid image without tables (src/zoic.cpp:135-137)
    c, tabs = host_setup(lensModel=0, useImage=1, image=np.ones((4, 5, 2), np.float32))
    assert c["bokehWidth"] == 5 and c["bokehHeight"] == 4 and all(not t.any() for t in tabs)
    with pytest.raises(capi.ZoicError) as e:
        host_setup(lensModel=7)
    assert e.value.code == capi.ERR_INVALID_ARGUMENT


def test_kolb_setup_random_parameters_all_lenses(port):
    """Seeded sweep over the node parameters' documented ranges (src/zoic.mtd) for every lens table: the derived
    constants, the rescaled element stack and all 32 exit-pupil boxes equal the oracle's bit for bit."""
    rng = np.random.default_rng(20261017)
    for lens in sorted(LENSES):
        fnum, focal = LENSES[lens]
        for _ in range(2):
            kw = dict(lensModel=1, lensDataPath=lens_path(lens),
                      focalLength=float(np.float32(focal * rng.uniform(0.6, 1.8))),
                      fStop=float(np.float32(rng.choice([1.0, 1.4, 2.0, 2.8, 4.0, 5.6, 8.0, 16.0]))),
                      focalDistance=float(np.float32(rng.uniform(15.0, 2000.0))),
                      sensorWidth=float(np.float32(rng.uniform(1.0, 4.5))),
                      kolbSamplingLUT=int(rng.integers(0, 2)))
            c, _ = host_setup(**kw)
            p = port.PortCamera(**kw)
            _same_constants(c, p.constants())
            p.close()


def test_lens_table_grammar_fuzz(port, tmp_path):
    """Seeded re-writings of every shipped lens table -- delimiters drawn from the reference's set (tab , ; : space,
    src/zoic.cpp:728,771), comment and blank lines, trailing blanks, CRLF, number formats -- parse to the same element
    stack and constants as the oracle's parser; with single-character delimiters also to the same as the original file.
    (Delimiter RUNS are a quirk of the reference that both reproduce: its counting pass collapses them, its reading pass
    advances the column counter for every delimiter, :771-790, so "a,  b" leaves a column unassigned.)"""
    rng = np.random.default_rng(42)
    delims = ["\t", ",", ";", ":", " ", "  ", "\t ", ", "]
    for lens in sorted(LENSES):
        rows = [l.split() for l in open(lens_path(lens)).read().splitlines() if l.strip() and not l.lstrip().startswith("#")]
        rows = [[t for t in " ".join(r).replace(",", " ").replace(";", " ").replace(":", " ").split()] for r in rows]
        fnum, focal = LENSES[lens]
        base_kw = dict(lensModel=1, lensDataPath=lens_path(lens), focalLength=focal, fStop=fnum, kolbSamplingLUT=0)
        base, _ = host_setup(**base_kw)
        for variant in range(3):
            runs = variant == 2   # the third variant uses delimiter runs
            lines = []
            # CRLF files: the reference keeps the '\r' on the last token of a line (std::stof ignores it), so such files
            # cannot have blank lines or trailing blanks (a lone "\r" would count as a data line / a token)
            crlf = rng.random() < 0.3
            if rng.random() < 0.7:
                lines.append("# rewritten %d" % variant)
            for r in rows:
                toks = []
                for t in r:
                    v = float(t)
                    style = rng.integers(0, 3)
                    toks.append(t if style == 0 else (repr(v) if style == 1 else ("%.6f" % v if float("%.6f" % v) == v else t)))
                line = toks[0]
                for t in toks[1:]:
                    line += str(rng.choice(delims if runs else delims[:5])) + t
                if not crlf and rng.random() < 0.3:
                    line += "  " if runs else " "   # a run at the end of a line shifts the columns of the NEXT line (same quirk)
                lines.append(line)
                if not crlf and rng.random() < 0.2:
                    lines.append("")
                if rng.random() < 0.15:
                    lines.append("#" + line)
            eol = "\r\n" if crlf else "\n"
            text = eol.join(lines) + (eol if rng.random() < 0.6 else "")
            path = tmp_path / ("%s_%d.dat" % (lens, variant))
            path.write_bytes(text.encode())
            kw = dict(base_kw, lensDataPath=str(path))
            c, _ = host_setup(**kw)
            p = port.PortCamera(**kw)
            _same_constants(c, p.constants(), nan_is_nan=runs)
            p.close()
            if not runs:
                assert bits_equal(product_constants_flat(c)["lenses"], product_constants_flat(base)["lenses"]), (lens, variant)


def test_bokeh_tables_random_images_with_ties(port):
    """Seeded images of random shape with few grey levels (ties in every sort), zero borders and repeated rows: the host
    statement of the table build equals the oracle's tables entry for entry (float bits, tie order included)."""
    rng = np.random.default_rng(99)
    for k in range(12):
        h, w = int(rng.integers(2, 70)), int(rng.integers(2, 90))
        levels = int(rng.choice([2, 3, 8, 64]))
        img = (rng.integers(0, levels, (h, w, 1)) / max(1, levels - 1)).astype(np.float32).repeat(3, axis=2)
        img[rng.random((h, w)) < 0.3] = 0.0
        if h > 4:
            img[1] = img[3]
        img[0, 0] = 1.0   # never entirely black
        kw = dict(lensModel=0, focalLength=3.5, fStop=2.8, useImage=1)
        _, tabs = host_setup(image=img, **kw)
        p = port.PortCamera(image=img, **kw)
        for a, b in zip(tabs, p.bokeh_tables()):
            assert a.dtype == b.dtype and np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                                                         b.view(np.uint32) if b.dtype == np.float32 else b), (k, h, w, levels)
        p.close()


def test_sqrt_threshold_decides_like_the_rounded_root():
    """The thin-lens kernels test qx^2 + qy^2 < ov_s_threshold instead of sqrt(qx^2 + qy^2) < radius (reference
    src/zoic.cpp:1302-1304).  The two agree for EVERY s iff the threshold is the smallest float whose correctly rounded
    root reaches the radius: checked here at the threshold and its predecessor, and on samples of s around it, with
    numpy's float32 sqrt (IEEE, like the device's __fsqrt_rn and the reference's sqrtf)."""
    lib = capi.load()
    rng = np.random.default_rng(5)
    radii = np.concatenate([
        np.float32([1e-30, 1.1754944e-38, 1e-41, 1e-20, 0.5, 1.0, 2.0, 3.0, 1.25, 0.1, 1e10, 1.8446743e19, 1.8446744e19, 3e38]),
        (rng.random(300, dtype=np.float32) * np.float32(10.0)).astype(np.float32),
        np.exp(rng.uniform(-80, 80, 300)).astype(np.float32)])
    with np.errstate(over="ignore", invalid="ignore"):
        for r in radii:
            t = np.float32(lib.zoicb_debug_sqrt_threshold(float(r)))
            assert np.sqrt(t) >= r, (r, t)
            below = np.nextafter(t, np.float32(0.0))
            assert t == 0 or np.sqrt(below) < r, (r, t)
            # a cloud of s values around the threshold: the two forms of the test agree on each
            bits = t.view(np.uint32).astype(np.int64) + np.arange(-2000, 2001)
            bits = bits[(bits >= 0) & (bits <= 0x7F800000)]
            s = bits.astype(np.uint32).view(np.float32)
            assert np.array_equal(np.sqrt(s) < r, s < t), r
    # nothing passes a radius that is zero, negative or NaN; every finite s passes an infinite one
    for r in (0.0, -0.0, -1.0, float("nan")):
        assert lib.zoicb_debug_sqrt_threshold(r) == 0.0
    assert lib.zoicb_debug_sqrt_threshold(float("inf")) == float("inf")
