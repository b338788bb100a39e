"""CPU: what pins the oracle.

The reference's own tests hold no golden pixels and OpenImageIO is not available; what pins the oracle to the reference's
output is tests/test_reference_filmstrips.py (the reference's own rendered film strips, thumbnail scale).  Pinned here:
  * regression: the oracle renders every frame of every fixture timeline to the committed
    sha256 (tests/golden/frames.json, made by tests/golden/make_golden.py);
  * closed-form known answers of SURVEY 8c (identities, constant images, wipe/dissolve endpoints);
  * the derived integer facts of SURVEY 9.1;
  * constants that can be cross-checked without OIIO (fast_sinpi error bound, colour-map knots
    against OpenCV's copies of the matplotlib tables);
  * the time-math worked examples of SURVEY 9.8.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import graph as G
from oracle import oracle as O
from tests import media
from tests import timelines as TL

GOLD = TL.golden()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_renders_golden_frames(name):
    tl = G.Timeline(os.path.join(TL.DATA, name))
    g = G.ImageGraph(tl, media.decode)
    frames = tl.frames()
    assert len(frames) == len(GOLD[name]["frames"])
    for t, rec in zip(frames, GOLD[name]["frames"]):
        assert t.value == rec["time"]
        node = g.exec(t)
        assert G.describe_tree(node) == rec["tree"]
        img = g.eval(node)
        if rec["sha256"]:
            assert list(img.shape) == rec["shape"] and str(img.dtype) == rec["dtype"]
            assert hashlib.sha256(img.tobytes()).hexdigest() == rec["sha256"], f"{name} @ {rec['time']}"
        else:
            assert img is None or img.size == 0


def test_time_math_worked_examples():
    # Transition.otio -- frames 0-1 plain, 2-7 inside the dissolve with value (f-2)/6, 8-9 the Fill
    trees = [r["tree"] for r in GOLD["Transition.otio"]["frames"]]
    assert all("toucan:Dissolve" not in "".join(t) for t in trees[:2] + trees[8:])
    for f in range(2, 8):
        line = [ln for ln in trees[f] if "toucan:Dissolve" in ln][0]
        assert "value=" + G._fmt_any((f - 2) / 6.0) in line
        assert any(f"Read: Counter.{f}.png" in ln for ln in trees[f])
    # LinearTimeWarp.otio: clip 1 (x2.0) reads Counter.0,2,4; clip 2 (x0.5) reads 0,1,1,2,2,3,3,4
    reads = ["".join(r["tree"]).split("Read: Counter.")[1].split(".png")[0] for r in GOLD["LinearTimeWarp.otio"]["frames"]]
    assert reads == ["0", "2", "4", "0", "1", "1", "2", "2", "3", "3", "4"]
    # Markers.otio starts at global_start_time 100
    assert GOLD["Markers.otio"]["start"] == 100.0


def test_u8_derived_facts():
    # u8 load -> store is the identity; invert == 255 - v; the u8 x u8 product through the
    # float path equals integer round-half-up
    v = np.arange(256, dtype=np.uint8).reshape(1, 64, 4)
    assert np.array_equal(O.convert(v, O.U8), v)
    inv = O.invert(v)
    assert np.array_equal(inv[..., :3], 255 - v[..., :3]) and np.array_equal(inv[..., 3], v[..., 3])
    a = np.zeros((256, 256, 4), np.uint8)
    a[..., 0] = np.arange(256)[:, None]
    a[..., 3] = np.arange(256)[None, :]
    pm = O.premult(a)
    aa, cc = a[..., 0].astype(np.int64), a[..., 3].astype(np.int64)
    want = np.where(cc == 255, aa, (2 * aa * cc + 255) // 510)
    assert np.array_equal(pm[..., 0], want)
    v16 = np.arange(65536, dtype=np.uint16).reshape(64, 256, 4)
    assert np.array_equal(O.convert(v16, O.U16), v16)


def test_known_answers():
    rng = np.random.default_rng(5)
    f = rng.random((37, 53, 4), dtype=np.float32)
    u = (f * 255).astype(np.uint8)
    assert np.array_equal(O.flip(O.flip(u)), u) and np.array_equal(O.flop(O.flop(f)), f)
    assert np.array_equal(O.invert(O.invert(u)), u)
    assert np.array_equal(O.dissolve(f, f[::-1].copy(), 0.0), f)
    assert np.array_equal(O.dissolve(f, f[::-1].copy(), 1.0), f[::-1])
    const = np.full((40, 60, 4), 0.375, np.float32)
    assert np.abs(O.blur(const, 10.0) - 0.375).max() < 1e-6
    assert np.abs(O.resize(f, (53, 37)) - f).max() < 1e-5       # same size: identity (fast_sinpi ripple)
    assert np.abs(O.rotate(f, 0.0) - f).max() < 1e-5            # no rotation: identity
    opaque = f.copy()
    opaque[..., 3] = 1.0
    assert np.array_equal(O.over(opaque, f[::-1].copy()), opaque)  # alpha = 1 fg covers
    clear = np.zeros_like(f)
    assert np.array_equal(O.over(clear, f), f)                    # alpha = 0 fg shows bg
    black = np.zeros((4, 4, 4), np.float32)
    white = np.ones((4, 4, 4), np.float32)
    # first / last knot: matplotlib plasma's ends (0.050383, 0.029803, 0.527975) / (0.940015, 0.975158, 0.131326), linearised
    assert np.allclose(O.color_map(black, "plasma")[0, 0, :3], _srgb_to_linear([0.050383, 0.029803, 0.527975]), atol=1e-6)
    assert np.allclose(O.color_map(white, "plasma")[0, 0, :3], _srgb_to_linear([0.940015, 0.975158, 0.131326]), atol=1e-6)
    assert O.fit((1280, 720), (640, 360)) == (0, 0, 1280, 720)
    assert O.fit((1280, 720), (100, 100)) == (280, 0, 720, 720)
    # wipes at the ends: value 0 -> the 200 px ramp starts at 0; far right is all `from`
    a = rng.random((8, 1280, 4), dtype=np.float32)
    b = rng.random((8, 1280, 4), dtype=np.float32)
    w0 = O.wipe(a, b, 0.0, False)
    assert np.array_equal(w0[:, 200:], a[:, 200:]) and np.array_equal(w0[:, 0], b[:, 0])


def test_fast_sinpi_error_bound():
    # OIIO documents max abs error 0.000918954611 for fast_sinpi on [-1, 1]
    x = np.linspace(-1, 1, 20001, dtype=np.float32)
    got = np.array([O.lib().o_fast_sinpi(float(v)) for v in x[::10]])
    err = np.abs(got - np.sin(np.pi * x[::10].astype(np.float64))).max()
    assert 0.0008 < err < 0.00093, err


def _srgb_to_linear(v):
    v = np.asarray(v, np.float64)
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)


def test_colour_map_knots_are_linearised_samples_at_the_17_knot_indices():
    # OIIO's color_map tables [V], pinned by the reference's render of data/Filter.otio (test_reference_filmstrips.py): 17 knots per
    # map = the published (sRGB-encoded) map at indices 0,16,...,240,255 converted to LINEAR light.  OpenCV ships the same 256-entry
    # tables at 8-bit precision, so every knot, re-encoded, must sit within half an 8-bit step (+ the de-quantisation's slack) of
    # that entry -- which fixes the colour space, the knot count and the spacing (round 1's sRGB-encoded knots at 0,15,...,225,255
    # fail all three).
    cv2 = pytest.importorskip("cv2")
    import ctypes

    index = list(range(0, 241, 16)) + [255]
    for name, code in [("plasma", cv2.COLORMAP_PLASMA), ("inferno", cv2.COLORMAP_INFERNO), ("magma", cv2.COLORMAP_MAGMA),
                       ("viridis", cv2.COLORMAP_VIRIDIS), ("turbo", cv2.COLORMAP_TURBO)]:
        lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), code)[0, :, ::-1].astype(np.float64) / 255.0
        p = ctypes.POINTER(ctypes.c_float)()
        n = O.lib().o_color_map_knots(name.encode(), ctypes.byref(p))
        assert n == 17, (name, n)
        knots = np.array([p[i] for i in range(3 * n)], np.float64).reshape(n, 3)
        want = _srgb_to_linear(lut[index])
        # half an 8-bit step in the encoded value, through the slope of the decode curve (<= 2.3), plus the table's 6 decimals
        slope = np.maximum(1.0 / 12.92, 2.4 / 1.055 * ((lut[index] + 0.055) / 1.055) ** 1.4)
        assert np.all(np.abs(knots - want) <= (1.0 / 255) * slope + 2e-6), (name, np.abs(knots - want).max())
    # magma's knots are the six-decimal matplotlib values: OIIO's published table starts 0.000113, 0.000036, 0.001073, 0.003066, ...
    p = ctypes.POINTER(ctypes.c_float)()
    O.lib().o_color_map_knots(b"magma", ctypes.byref(p))
    assert [round(p[i], 6) for i in range(6)] == [0.000113, 0.000036, 0.001073, 0.003066, 0.002406, 0.016033]


def test_kernel_and_oracle_colour_map_tables_are_the_same_numbers():
    # both generated by tools/make_colormaps.py; the product never reads oracle/, so the numbers exist twice
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = open(os.path.join(root, "oracle", "colormap_tables.inc")).read().replace("k_", "X")
    b = open(os.path.join(root, "toucan_b200", "csrc", "tc_colormap_tables.inc")).read().replace("kMap_", "X")
    assert a == b and a.count("static const float") == 8


def test_builtin_png_reader_matches_harness_decode(built):
    # the C++ host's built-in reader and the test harness' Pillow decode apply the same two
    # read-side conventions (A = 1 for RGB, alpha association with OIIO's u8 rounding)
    from toucan_b200 import render

    for f in ["Counter.3.png", "Letter_A.png", "Gradient.png"]:
        p = os.path.join(TL.DATA, f)
        assert np.array_equal(render.decode_png(p), media.decode(p)), f
    # ... and the "oiio:UnassociatedAlpha" switch leaves the file's values alone in both
    render.set_unassociated_alpha(True)
    try:
        for f in ["Counter.3.png", "Letter_A.png"]:
            p = os.path.join(TL.DATA, f)
            got = render.decode_png(p)
            assert np.array_equal(got, media.decode_unassociated(p)), f
            assert not np.array_equal(got, media.decode(p)), f
        assert tuple(render.decode_png(os.path.join(TL.DATA, "Counter.3.png"))[0, 0]) == (255, 255, 255, 0)  # transparent white
    finally:
        render.set_unassociated_alpha(False)


def test_colour_matrices_against_published_aces_values():
    # Published ACES conversion matrices (Bradford-adapted, D65 -> ACES white): the oracle derives its
    # matrices from primaries + Bradford, these are the values tabulated in the ACES documentation
    # and reproduced by OCIO's built-in config to 6 decimals.
    srgb_to_ap1 = np.array([[0.613097, 0.339523, 0.047379], [0.070194, 0.916354, 0.013452], [0.020616, 0.109570, 0.869815]])
    srgb_to_ap0 = np.array([[0.439633, 0.382989, 0.177378], [0.089776, 0.813439, 0.096784], [0.017541, 0.111547, 0.870912]])
    assert np.abs(O.color_matrix("rec709", "ap1") - srgb_to_ap1).max() < 2e-5
    assert np.abs(O.color_matrix("rec709", "ap0") - srgb_to_ap0).max() < 2e-4
    ap0_to_ap1 = np.array([[1.4514393161, -0.2365107469, -0.2149285693], [-0.0765537734, 1.1762296998, -0.0996759264],
                           [0.0083161484, -0.0060324498, 0.9977163014]])
    assert np.abs(O.color_matrix("ap0", "ap1") - ap0_to_ap1).max() < 1e-5
    # round trips are the identity
    for a, b in [("rec709", "p3d65"), ("rec2020", "ap1")]:
        assert np.abs(O.color_matrix(a, b).astype(np.float64) @ O.color_matrix(b, a).astype(np.float64) - np.eye(3)).max() < 1e-5
    # ACEScct (S-2016-001): continuous at the break point, 0.18 grey -> 0.4135884
    img = np.zeros((1, 2, 4), np.float32)
    img[0, 0, :3] = 0.18
    img[0, 1, :3] = 0.0078125
    out = O.colorconvert(img, "ACEScg", "ACEScct")
    assert abs(float(out[0, 0, 0]) - 0.4135884) < 1e-5
    assert abs(float(out[0, 1, 0]) - (10.5402377416545 * 0.0078125 + 0.0729055341958355)) < 1e-6


def test_noise_fast_path_stays_within_its_decision_band(tmp_path):
    """tests/cpp/noise_fastpath_check.c: the seeded Noise kernel evaluates sqrt(-2 ln r / r) with Newton-refined SFU seeds and
    fused multiply-adds and hands over to the exact IEEE chain within 1024 double ulps of a float rounding boundary; the C
    restatement of that fast path must stay within a few ulps of the oracle's chain (full sweep of all 3.2e8 float r: 2 ulps;
    here every 61st value)."""
    import subprocess

    O.build()
    here = os.path.dirname(os.path.abspath(__file__))
    build = os.path.join(os.path.dirname(here), "oracle", "_build")
    exe = str(tmp_path / "noise_fastpath_check")
    subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", os.path.join(here, "cpp", "noise_fastpath_check.c"), "-o", exe,
                    "-L", build, "-loracle", "-lm", f"-Wl,-rpath,{build}"], check=True, capture_output=True)
    out = subprocess.run([exe, "61"], check=True, capture_output=True, text=True).stdout
    assert int(out.split("max_ulps")[1]) <= 8, out
