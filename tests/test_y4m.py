"""y4m output conversion (bin/toucan-render/App.cpp:302-485).

CPU: oracle/y4m.py against vectors the REAL libswscale 9.1.100 (FFmpeg 8.0, the reference's pin) produced
(tests/golden/make_y4m_golden.py) -- every byte of the small cases, sha256 of the large ones.
GPU: tc_rgb_to_y4m against the same vectors and against the oracle on float / half frames.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import y4m as Y

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = np.load(os.path.join(HERE, "golden", "y4m_small.npz"))
LARGE = json.load(open(os.path.join(HERE, "golden", "y4m_large.json")))


def small_cases():
    return sorted({k.rsplit("/", 1)[0] for k in SMALL.files})


def golden_planes(case):
    n = 4 if case.startswith("444alpha") else 3
    return [SMALL[f"{case}/plane{k}"] for k in range(n)]


def seeded(fmt, w, h, seed):
    rng = np.random.default_rng(seed)
    if fmt == "444p16":
        return rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)
    return rng.integers(0, 256, (h, w, 4 if fmt == "444alpha" else 3), dtype=np.uint8)


def rgba(img):
    """Golden inputs are packed RGB(A); the render path's frames are RGBA."""
    if img.shape[2] == 4:
        return np.ascontiguousarray(img)
    full = np.full(img.shape[:2] + (4,), 65535 if img.dtype == np.uint16 else 255, dtype=img.dtype)
    full[..., :3] = img
    return full


def test_golden_is_from_the_pinned_swscale():
    assert LARGE["libswscale"] == "9.1.100"  # FFmpeg 8.0, cmake/SuperBuild/BuildFFmpeg.cmake:390


@pytest.mark.parametrize("case", small_cases())
def test_oracle_matches_libswscale_small(case):
    fmt = case.split("/")[0]
    got = Y.sws_planes(SMALL[f"{case}/in"], fmt)
    for g, w in zip(got, golden_planes(case)):
        assert g.dtype == w.dtype and np.array_equal(g, w), case


@pytest.mark.parametrize("case", sorted(LARGE["cases"]))
def test_oracle_matches_libswscale_large(case):
    fmt, size = case.split("/")
    w, h = [int(v) for v in size.split("x")]
    info = LARGE["cases"][case]
    got = Y.sws_planes(seeded(fmt, w, h, info["seed"]), fmt)
    assert [hashlib.sha256(p.tobytes()).hexdigest() for p in got] == info["sha256"]


def test_filter_shape_facts():
    # the fast-bilinear increment: same-size luma is a stretch whose left tap grows by 5/16384 per column
    pos, filt = Y.init_filter(65516, 64, 64)
    assert filt.shape[1] == 4 and (filt.sum(1) == 16384).all()
    assert pos[:7].tolist() == list(range(7)) and (filt[:7, 0] == 16384).all()  # the tiny left tap is dropped
    assert filt[7].tolist()[:2] == [38, 16346] and pos[7] == 6
    # 16-bit output keeps the plain increments: luma untouched, chroma a 1:2 bilinear upsample sited at 1/4, 3/4
    pos, filt = Y.init_filter(32768, 32, 64)
    assert filt[1].tolist() == [12288, 4096, 0, 0] and filt[2].tolist() == [4096, 12288, 0, 0] and pos[:4].tolist() == [0, 0, 0, 1]
    assert Y.x_increments("444p16", 64)[:2] == (65536, 32768) and Y.x_increments("422", 64)[:2] == (65516, 65516)


def test_header_and_frame_bytes():
    assert Y.header(1280, 720, 24.0, "422") == b"YUV4MPEG2 W1280 H720 F24:1 C422\n"
    assert Y.header(16, 8, 23.976, "444alpha") == b"YUV4MPEG2 W16 H8 F24000:1001 C444alpha\n"
    assert Y.to_rational(25.0) == (25, 1)
    f = np.zeros((4, 16, 4), np.float32)
    f[..., 3] = 1.0
    for fmt, n in (("422", 16 * 4 * 2), ("444", 16 * 4 * 3), ("444alpha", 16 * 4 * 4), ("444p16", 16 * 4 * 6)):
        b = Y.frame_bytes(f, fmt)
        assert len(b) == n
    assert Y.frame_bytes(f, "422")[:1] == b"\x10" and Y.frame_bytes(f, "444alpha")[-1:] == b"\xff"  # black = Y 16, opaque


def test_oracle_rejects_odd_and_tiny_widths():
    for w in (6, 9):
        with pytest.raises(ValueError):
            Y.sws_planes(np.zeros((2, w, 3), np.uint8), "422")


# ---- GPU ------------------------------------------------------------------------------
def _gpu_planes(arr, fmt):
    import torch

    from toucan_b200 import ops

    t = torch.from_numpy(arr.view(np.int16) if arr.dtype == np.uint16 else arr).cuda()
    if arr.dtype == np.uint16:
        t = t.view(torch.uint16)
    h, w = arr.shape[:2]
    out = ops.rgb_to_y4m(t, fmt).cpu().numpy()
    return out, [np.ascontiguousarray(p) for p in ops.y4m_planes(out, fmt, w, h)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", small_cases())
def test_kernel_matches_libswscale_small(case):
    fmt = case.split("/")[0]
    _, got = _gpu_planes(rgba(SMALL[f"{case}/in"]), fmt)
    for g, w in zip(got, golden_planes(case)):
        assert g.dtype == w.dtype and np.array_equal(g, w), case


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(LARGE["cases"]))
def test_kernel_matches_libswscale_large(case):
    fmt, size = case.split("/")
    w, h = [int(v) for v in size.split("x")]
    info = LARGE["cases"][case]
    _, got = _gpu_planes(rgba(seeded(fmt, w, h, info["seed"])), fmt)
    assert [hashlib.sha256(p.tobytes()).hexdigest() for p in got] == info["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", Y.FORMATS)
@pytest.mark.parametrize("dtype", [np.float32, np.float16, np.uint8, np.uint16])
def test_kernel_quantises_like_paste(fmt, dtype):
    """Frames in any working type: paste's storage conversion happens inside the kernel (App.cpp:357-371)."""
    rng = np.random.default_rng(11)
    w, h = 322, 9
    f = rng.random((h, w, 4), dtype=np.float32) * 1.2 - 0.1  # out-of-range values clamp in the conversion
    src = O.convert(f, {np.float32: O.F32, np.float16: O.F16, np.uint8: O.U8, np.uint16: O.U16}[dtype])
    raw, _ = _gpu_planes(src, fmt)
    assert raw.tobytes() == Y.frame_bytes(src, fmt)


@pytest.mark.gpu
def test_kernel_full_8k_frame_against_strips():
    """BASELINE's 7680x4320 f32 frame: rows are independent, so every row band of the full-size conversion must
    equal the conversion of that band alone (checked against the oracle on three bands)."""
    import torch

    from toucan_b200 import ops

    w, h = 7680, 4320
    g = torch.Generator(device="cuda").manual_seed(5)
    frame = torch.rand((h, w, 4), device="cuda", generator=g)
    for fmt in ("422", "444p16"):
        full = ops.y4m_planes(ops.rgb_to_y4m(frame, fmt), fmt, w, h)
        for y0 in (0, 2161, h - 2):
            band = frame[y0:y0 + 2].contiguous()
            want = Y.frame_planes(band.cpu().numpy(), fmt)
            for p, q in zip(full, want):
                assert np.array_equal(p[y0:y0 + 2].cpu().numpy(), q), (fmt, y0)


@pytest.mark.gpu
def test_kernel_rejects():
    import torch

    from toucan_b200 import capi, ops

    with pytest.raises(capi.TcError):
        ops.rgb_to_y4m(torch.zeros((4, 9, 4), dtype=torch.uint8, device="cuda"), "422")
    with pytest.raises(capi.TcError):
        ops.rgb_to_y4m(torch.zeros((4, 6, 4), dtype=torch.uint8, device="cuda"), "444")


# ---- the output stage at timeline level ------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fmt", Y.FORMATS)
def test_timeline_y4m_output_on_device(cuda, fmt):
    """`toucan-render - -y4m <fmt>` (App.cpp:340-485) through toucan::FrameScheduler: the bytes handed to the writer
    are the oracle's paste + swscale of the oracle's frame (Transition.otio is in the bit-exact class)."""
    from oracle import graph as G
    from tests import media
    from tests import timelines as TL

    name = "Transition.otio"
    tl = TL.open_timeline(name)
    otl = G.Timeline(os.path.join(TL.DATA, name))
    og = G.ImageGraph(otl, media.decode)
    got = {}
    tl.render_range(2, 5, lambda f, p: got.__setitem__(f, p.tobytes()), y4m_format=fmt)
    for i, t in enumerate(otl.frames()[2:5]):
        assert got[2 + i] == Y.frame_bytes(og.render(t), fmt), (fmt, i)
    tl.close()


@pytest.mark.gpu
def test_toucan_render_cli_y4m(cuda, tmp_path):
    """The command line renderer's y4m stream: header, then "FRAME\\n" + planes per frame."""
    import subprocess

    from oracle import graph as G
    from tests import media
    from tests import timelines as TL
    from toucan_b200 import capi, render

    exe = os.path.join(capi.LIB_DIR, "toucan-render")
    out = tmp_path / "transition2.y4m"
    name = "Transition2.otio"  # PNG media: the command line tool decodes with its built-in reader
    r = subprocess.run([exe, os.path.join(TL.DATA, name), str(out), "-y4m", "422"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    data = open(out, "rb").read()
    otl = G.Timeline(os.path.join(TL.DATA, name))
    og = G.ImageGraph(otl, media.decode)
    header = Y.header(1280, 720, 24.0, "422")
    assert render.y4m_header(1280, 720, 24.0, "422") == header and render.y4m_header(8, 8, 29.97, "444p16") == Y.header(8, 8, 29.97, "444p16")
    assert data.startswith(header)
    frames = otl.frames()
    per = 6 + 1280 * 720 * 2
    assert len(data) == len(header) + per * len(frames)
    for i in (0, len(frames) // 2, len(frames) - 1):
        chunk = data[len(header) + i * per:len(header) + (i + 1) * per]
        assert chunk[:6] == b"FRAME\n" and chunk[6:] == Y.frame_bytes(og.render(frames[i]), "422"), i
    bad = subprocess.run([exe, os.path.join(TL.DATA, name), "-", "-y4m", "420"], capture_output=True, text=True, timeout=120)
    assert bad.returncode == 1 and "Cannot find the given y4m format" in bad.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_kernel_matches_the_oracle_on_random_frames(seed):
    """Random frame sizes (odd heights, widths that are not a multiple of the kernel's groups) and working types against the
    restatement of libswscale that the golden vectors pin (every plane, every byte)."""
    rng = np.random.default_rng(8000 + seed)
    for _ in range(6):
        fmt = Y.FORMATS[int(rng.integers(0, len(Y.FORMATS)))]
        w = int(rng.integers(5, 200)) * 2 if fmt == "422" else int(rng.integers(8, 400))
        h = int(rng.integers(1, 200))
        dt = [np.float32, np.float16, np.uint8, np.uint16][int(rng.integers(0, 4))]
        f = rng.random((h, w, 4)).astype(np.float32) * 1.1 - 0.05  # a little outside [0, 1]: the conversion clamps
        arr = (np.clip(f, 0, 1) * 255).astype(np.uint8) if dt == np.uint8 else (np.clip(f, 0, 1) * 65535).astype(np.uint16) if dt == np.uint16 else f.astype(dt)
        try:
            want = Y.frame_bytes(arr, fmt)
        except ValueError:
            continue  # (a width libswscale's filter set-up refuses)
        out, _ = _gpu_planes(np.ascontiguousarray(arr), fmt)
        assert out.tobytes() == want, f"{fmt} {w}x{h} {np.dtype(dt).name}"
