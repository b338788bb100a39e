"""Ingest (lib/toucanRender/Read.cpp:34-216): the C++ host's built-in readers.

The reference reads stills through OpenImageIO, i.e. libjpeg(-turbo) / libpng with their defaults.  Pillow
links the same libjpeg-turbo, so comparing the built-in JPEG reader with Pillow's decode pins it to the real
library: every pixel must match (islow IDCT, fancy upsampling, integer colour tables are all exact integer
algorithms).  These tests run on the CPU -- decode is host code -- except the last, which renders a JPEG-backed
fixture through the command line tool.
"""
import hashlib
import os

import numpy as np
import pytest

from tests import media
from tests import timelines as TL

Image = pytest.importorskip("PIL.Image")


def pil_rgb(path):
    return np.asarray(Image.open(path).convert("RGB"))


def synth(w, h, kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    a = np.stack([xx * 255 // max(w - 1, 1), yy * 255 // max(h - 1, 1), ((xx + yy) * 7) % 256], -1).astype(np.uint8)
    a[h // 3:h // 2, w // 4:w // 2] = (255, 0, 0)  # a saturated edge: exercises the range limits
    return a


def test_reference_fixture_jpeg(built):
    """data/Charlie.jpg is a progressive 4:4:4 JPEG (spectral selection + successive approximation)."""
    from toucan_b200 import render

    path = os.path.join(TL.DATA, "Charlie.jpg")
    got = render.decode_jpeg(path)
    assert got.shape == (720, 1280, 4) and got.dtype == np.uint8
    assert np.array_equal(got[..., :3], pil_rgb(path)) and (got[..., 3] == 255).all()  # Read.cpp:86-94: A = 1
    assert np.array_equal(got, media.decode(path))  # what the parity tests feed the oracle


@pytest.mark.parametrize("size", [(1, 1), (2, 2), (3, 5), (8, 8), (17, 9), (33, 31), (100, 75), (257, 129)])
@pytest.mark.parametrize("subsampling", [0, 1, 2])  # 4:4:4, 4:2:2 (h2v1), 4:2:0 (h2v2)
def test_jpeg_matches_libjpeg_turbo(built, tmp_path, size, subsampling):
    from toucan_b200 import render

    w, h = size
    n = 0
    for kind in ("noise", "ramp"):
        for progressive in (False, True):
            for extra in ({"quality": 30}, {"quality": 92, "optimize": True}, {"quality": 75, "restart_marker_blocks": 3}):
                path = str(tmp_path / f"{kind}_{int(progressive)}_{n}.jpg")
                Image.fromarray(synth(w, h, kind, n)).save(path, subsampling=subsampling, progressive=progressive, **extra)
                assert np.array_equal(render.decode_jpeg(path)[..., :3], pil_rgb(path)), (size, subsampling, kind, progressive, extra)
                n += 1


def test_jpeg_rgb_colourspace_and_errors(built, tmp_path):
    from toucan_b200 import capi, render

    path = str(tmp_path / "rgb.jpg")
    Image.fromarray(synth(37, 21, "noise")).save(path, quality=85, keep_rgb=True)  # Adobe-less RGB components 'R','G','B'
    assert np.array_equal(render.decode_jpeg(path)[..., :3], pil_rgb(path))
    Image.fromarray(synth(9, 9, "noise")[..., 0]).save(path)
    with pytest.raises(capi.TcError, match="grayscale"):
        render.decode_jpeg(path)
    for junk in (b"\xff\xd8\xff\xd9", b"not a jpeg", b""):
        with open(path, "wb") as f:
            f.write(junk)
        with pytest.raises(capi.TcError):
            render.decode_jpeg(path)
    # a truncated stream keeps what was decoded (zero coefficients for the rest), like libjpeg's premature-EOF path
    Image.fromarray(synth(64, 64, "ramp")).save(path, quality=80)
    data = open(path, "rb").read()
    with open(path, "wb") as f:
        f.write(data[:len(data) * 2 // 3])
    assert render.decode_jpeg(path).shape == (64, 64, 4)


@pytest.mark.gpu
def test_cli_reads_jpeg_media_itself(cuda, tmp_path):
    """toucan-render on a fixture backed by Charlie.jpg with NO pre-decoded media: the read node decodes the file
    with the built-in reader and the frames equal the golden ones (which were rendered from Pillow's decode)."""
    import subprocess

    from toucan_b200 import capi

    gold = TL.golden()["TransitionWipe.otio"]["frames"]
    exe = os.path.join(capi.LIB_DIR, "toucan-render")
    out = tmp_path / "wipe.raw"
    r = subprocess.run([exe, os.path.join(TL.DATA, "TransitionWipe.otio"), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out, np.uint8).reshape(len(gold), 720, 1280, 4)
    for i, rec in enumerate(gold):
        assert hashlib.sha256(raw[i].tobytes()).hexdigest() == rec["sha256"], f"frame {i}"


# ---- bundles and "open a file as a timeline" (TimelineWrapper.cpp:68-306) ------------------------
def _bundle_doc(name):
    """The fixture with its media moved under media/, the layout OTIO's otioz / otiod adapters write."""
    import json

    doc = json.load(open(os.path.join(TL.DATA, name)))
    files = set()

    def walk(o):
        if isinstance(o, dict):
            schema = o.get("OTIO_SCHEMA", "")
            if schema.startswith("ExternalReference"):
                files.add(o["target_url"])
                o["target_url"] = "media/" + o["target_url"]
            elif schema.startswith("ImageSequenceReference"):
                n0 = int(o["available_range"]["start_time"]["value"])
                for i in range(n0, n0 + int(o["available_range"]["duration"]["value"])):
                    files.add(f"{o['name_prefix']}{i:0{o['frame_zero_padding']}d}{o['name_suffix']}")
                o["target_url_base"] = "media/"
            for v in o.values():
                walk(v)
        elif isinstance(o, list):
            for v in o:
                walk(v)

    walk(doc)
    return json.dumps(doc), sorted(files)


def make_otioz(path, name, media_compression=None):
    import zipfile

    text, files = _bundle_doc(name)
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("content.otio", text, compress_type=zipfile.ZIP_DEFLATED)  # the adapter deflates the timeline ...
        for f in files:
            z.write(os.path.join(TL.DATA, f), "media/" + f, compress_type=media_compression or zipfile.ZIP_STORED)  # ... and stores media


def make_otiod(path, name):
    import shutil

    text, files = _bundle_doc(name)
    os.makedirs(os.path.join(path, "media"))
    with open(os.path.join(path, "content.otio"), "w") as f:
        f.write(text)
    for f in files:
        shutil.copyfile(os.path.join(TL.DATA, f), os.path.join(path, "media", f))


def test_bundle_errors(built, tmp_path):
    """Open-time errors of the archive front-end, with the reference's messages (TimelineWrapper.cpp:150-211)."""
    import zipfile

    from toucan_b200 import capi, render

    bad = str(tmp_path / "compressed.otioz")
    make_otioz(bad, "Transition.otio", media_compression=zipfile.ZIP_DEFLATED)
    with pytest.raises(capi.TcError, match="Media is not uncompressed: media/Counter.0.png"):
        render.Timeline(bad)
    missing = str(tmp_path / "missing.otioz")
    text, _ = _bundle_doc("Transition.otio")
    with zipfile.ZipFile(missing, "w") as z:
        z.writestr("content.otio", text)
    with pytest.raises(capi.TcError, match="Cannot locate: media/"):
        render.Timeline(missing)
    empty = str(tmp_path / "empty.otioz")
    with zipfile.ZipFile(empty, "w") as z:
        z.writestr("readme.txt", "no timeline here")
    with pytest.raises(capi.TcError, match="Cannot locate: content.otio"):
        render.Timeline(empty)
    with open(str(tmp_path / "junk.otioz"), "wb") as f:
        f.write(b"this is not a zip archive at all")
    with pytest.raises(capi.TcError, match="Cannot open zip file"):
        render.Timeline(str(tmp_path / "junk.otioz"))
    with pytest.raises(capi.TcError, match="Unrecognized file"):
        render.Timeline(str(tmp_path / "movie.mov"))


def _render_all(tl, n):
    got = {}
    tl.render_range(0, n, lambda f, p: got.__setitem__(f, hashlib.sha256(p.tobytes()).hexdigest()))
    return [got[i] for i in range(n)]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Transition.otio", "TransitionWipe.otio", "CompositeTracks.otio"])
def test_otioz_and_otiod_render_like_the_plain_timeline(cuda, tmp_path, name):
    """PNG sequences and a JPEG still read in place from the memory-mapped archive (and from the unpacked
    bundle): every frame equals the golden frame of the plain .otio fixture."""
    from toucan_b200 import render

    gold = [r["sha256"] for r in TL.golden()[name]["frames"]]
    z = str(tmp_path / "bundle.otioz")
    make_otioz(z, name)
    tl = render.Timeline(z)
    tl.prepare()
    assert _render_all(tl, len(gold)) == gold
    tl.close()
    d = str(tmp_path / "bundle.otiod")
    make_otiod(d, name)
    tl = render.Timeline(d)
    tl.prepare()
    assert _render_all(tl, len(gold)) == gold
    tl.close()


@pytest.mark.gpu
def test_image_and_sequence_open_as_timelines(cuda, tmp_path):
    """TimelineWrapper.cpp:262-306: a still becomes a one-frame timeline at 24 fps, a numbered file the sequence it
    belongs to (ordered by number, same prefix / padding / extension)."""
    import shutil

    from toucan_b200 import render

    still = os.path.join(TL.DATA, "Charlie.jpg")
    tl = render.Timeline(still)
    tl.prepare()
    info = tl.info()
    assert (info["frames"], info["rate"], info["width"], info["height"]) == (1, 24.0, 1280, 720)
    out = []
    tl.render_range(0, 1, lambda f, p: out.append(p.copy()))
    assert np.array_equal(out[0], media.decode(still))
    tl.close()
    for i in (3, 4, 5, 7):  # a sequence with a hole and a decoy of another padding
        shutil.copyfile(os.path.join(TL.DATA, f"Counter.{i}.png"), str(tmp_path / f"shot.{i:04d}.png"))
    shutil.copyfile(os.path.join(TL.DATA, "Counter.9.png"), str(tmp_path / "shot.9.png"))
    tl = render.Timeline(str(tmp_path / "shot.0004.png"))
    tl.prepare()
    info = tl.info()
    assert (info["frames"], info["start"]) == (4, 0.0)  # 4 files found -> duration 4 from frame 3 (files 3, 4, 5, 6)
    frames = {}
    tl.render_range(0, 4, lambda f, p: frames.__setitem__(f, p.copy()))
    tl.close()
    # the same one-clip timeline spelled out for the oracle
    from oracle import graph as G

    def rt(v):
        return {"OTIO_SCHEMA": "RationalTime.1", "rate": 24.0, "value": float(v)}

    doc = {"OTIO_SCHEMA": "Timeline.1", "name": "", "tracks": {"OTIO_SCHEMA": "Stack.1", "name": "tracks", "children": [
        {"OTIO_SCHEMA": "Track.1", "name": "Video", "kind": "Video", "children": [
            {"OTIO_SCHEMA": "Clip.1", "name": "", "source_range": {"OTIO_SCHEMA": "TimeRange.1", "start_time": rt(3), "duration": rt(4)},
             "media_reference": {"OTIO_SCHEMA": "ImageSequenceReference.1", "target_url_base": str(tmp_path), "name_prefix": "shot.",
                                 "name_suffix": ".png", "start_frame": 3, "frame_step": 1, "rate": 24.0, "frame_zero_padding": 4,
                                 "available_range": None}}]}]}}
    otl = G.Timeline(doc)
    otl.dir = str(tmp_path)
    og = G.ImageGraph(otl, media.decode)
    for f, t in enumerate(otl.frames()):
        assert np.array_equal(frames[f], og.render(t)), f  # frame 3 names file 0006, which does not exist: background only


# ---- OpenEXR stills (half / float media) -------------------------------------------------------------------------
EXR_DIR = os.path.join(TL.GOLDEN, "exr")


def _exr_cases():
    return sorted(f for f in os.listdir(EXR_DIR) if f.endswith(".exr") and not f.startswith("refused"))


@pytest.mark.parametrize("name", _exr_cases())
def test_exr_reader_matches_openexr(built, name):
    """Files written by the real OpenEXR library (tests/golden/make_exr_golden.py): every sample equals what OpenEXR's own
    reader returns; the frame's type is the file's sample type; RGB files get A = 1 (Read.cpp:86-94)."""
    from toucan_b200 import render

    want = np.load(os.path.join(EXR_DIR, "expected.npz"))[name]
    got = render.decode_exr(os.path.join(EXR_DIR, name))
    assert got.dtype == (np.float16 if name.startswith("half") else np.float32) and got.shape == want.shape
    assert np.array_equal(got.astype(np.float32), want)
    if name.endswith("_rgb.exr"):
        assert (got[..., 3] == 1).all()


def test_exr_reader_refuses_what_it_does_not_decode(built, tmp_path):
    from toucan_b200 import capi, render

    with pytest.raises(capi.TcError, match="DWA"):
        render.decode_exr(os.path.join(EXR_DIR, "refused_dwaa_rgb.exr"))
    data = open(os.path.join(EXR_DIR, "float_zip_rgba.exr"), "rb").read()
    bad = str(tmp_path / "cut.exr")
    with open(bad, "wb") as f:
        f.write(data[:len(data) // 2])
    with pytest.raises(capi.TcError):
        render.decode_exr(bad)
    with open(bad, "wb") as f:
        f.write(b"\x00" * 64)
    with pytest.raises(capi.TcError, match="not an OpenEXR"):
        render.decode_exr(bad)


@pytest.mark.gpu
def test_exr_still_renders_as_a_timeline(cuda):
    """A half-float EXR opened as a one-frame timeline: decoded by the read node, composited over the black background
    in half precision -- equal to the oracle's evaluation of the same decoded frame."""
    from oracle import oracle as O
    from toucan_b200 import render

    path = os.path.join(EXR_DIR, "half_zip_rgba.exr")
    tl = render.Timeline(path)
    tl.prepare()
    info = tl.info()
    assert (info["frames"], info["width"], info["height"]) == (1, 37, 37)
    out = []
    tl.render_range(0, 1, lambda f, p: out.append(p.copy()))
    src = render.decode_exr(path)
    want = O.comp(src, O.fill(37, 37, [0.0, 0.0, 0.0, 1.0], O.U8), True)
    assert out[0].dtype == np.float16 and np.array_equal(out[0].view(np.uint16), want.view(np.uint16))
    tl.close()


# ---- image-sequence outputs of the command line tool (App.cpp:243-252) -----------------------------------------------
@pytest.mark.gpu
def test_cli_writes_png_and_exr_sequences(cuda, tmp_path):
    import json
    import subprocess

    from oracle import graph as G
    from toucan_b200 import capi, render, workloads

    exe = os.path.join(capi.LIB_DIR, "toucan-render")
    # u8 timeline -> 8-bit PNG sequence numbered from the output name's number; opaque frames read back unchanged
    gold = TL.golden()["Transition.otio"]["frames"]
    r = subprocess.run([exe, os.path.join(TL.DATA, "Transition.otio"), str(tmp_path / "out.0100.png")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    names = sorted(f for f in os.listdir(tmp_path) if f.endswith(".png"))
    assert names == [f"out.{100 + i:04d}.png" for i in range(len(gold))]
    for i in (0, 4, len(gold) - 1):
        im = np.asarray(Image.open(str(tmp_path / names[i])))
        assert im.shape == (720, 1280, 4) and hashlib.sha256(im.tobytes()).hexdigest() == gold[i]["sha256"], i
    # float EXR still + Invert -> float EXR sequence, every sample as rendered
    doc = workloads.filter_chain_timeline(frames=3, effects=[("toucan:Invert", {})], url="exr/float_zip_rgba.exr")
    otio = tmp_path / "invert.otio"
    otio.write_text(json.dumps(doc))
    os.symlink(EXR_DIR, str(tmp_path / "exr"))
    r = subprocess.run([exe, str(otio), str(tmp_path / "inv.exr")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    src = render.decode_exr(os.path.join(EXR_DIR, "float_zip_rgba.exr"))
    otl = G.Timeline(doc)
    otl.dir = str(tmp_path)
    want = G.ImageGraph(otl, lambda p: src if p.endswith("float_zip_rgba.exr") else None).render(G.RT(1.0, 24.0))
    for i in range(3):
        got = render.decode_exr(str(tmp_path / f"inv{i}.exr"))
        assert got.dtype == np.float32 and np.array_equal(got, want), i
    # the half variant is written as half; an 8-bit timeline asked for .exr is converted to half on the device
    r = subprocess.run([exe, os.path.join(TL.DATA, "Transition.otio"), str(tmp_path / "t.00.exr")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    got = render.decode_exr(str(tmp_path / "t.03.exr"))
    u8 = np.asarray(Image.open(str(tmp_path / names[3])))
    assert got.dtype == np.float16 and np.array_equal(got, (u8.astype(np.float32) * np.float32(1.0 / 255.0)).astype(np.float16))
    # encoders that are not built in are refused, not silently written as raw
    r = subprocess.run([exe, os.path.join(TL.DATA, "Transition.otio"), str(tmp_path / "out.mov")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Unsupported output" in r.stdout


def test_exr_writer_round_trips_and_is_valid_openexr(built, tmp_path):
    """The sequence writer's files decode to the same samples with the built-in reader and -- when OpenCV's OpenEXR codec
    is available -- with the OpenEXR library itself."""
    from toucan_b200 import render

    rng = np.random.default_rng(9)
    for dtype in (np.float16, np.float32):
        a = (rng.random((45, 83, 4), dtype=np.float32) * 4.0 - 1.5).astype(dtype)
        a[:20, :30] = dtype(0.5)
        path = str(tmp_path / f"w_{np.dtype(dtype).name}.exr")
        render.write_exr(str(tmp_path / "raw.exr"), a, zip=False)
        assert np.array_equal(render.decode_exr(str(tmp_path / "raw.exr")).view(np.uint8), a.view(np.uint8))
        render.write_exr(path, a)
        back = render.decode_exr(path)
        assert back.dtype == dtype and np.array_equal(back.view(np.uint16 if dtype == np.float16 else np.uint32),
                                                      a.view(np.uint16 if dtype == np.float16 else np.uint32))
        os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
        try:
            import cv2

            b = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        except Exception:
            b = None
        if b is not None:
            assert np.array_equal(b[..., [2, 1, 0, 3]], a.astype(np.float32))


@pytest.mark.gpu
def test_cli_renders_the_stack_workload_from_exr_stills(cuda, tmp_path):
    """The headline structure (tracks of blurred stills joined by dissolves) end to end from FILES:
    float EXR stills on disk -> read nodes -> device source cache -> fused stack kernel -> raw float frames, against
    the oracle's node-by-node evaluation."""
    import json
    import subprocess

    from oracle import graph as G
    from toucan_b200 import capi, render, workloads

    w, h, tracks = 192, 108, 3
    doc = workloads.stack_timeline(tracks=tracks, frames=80)
    rng = np.random.default_rng(21)
    stills = {}
    os.makedirs(str(tmp_path / "stills"))
    text = json.dumps(doc)
    for t in range(tracks):
        for k in range(2):
            a = rng.random((h, w, 4), dtype=np.float32)
            a[..., 3] = np.clip(rng.random((h, w), dtype=np.float32) * 1.4 - 0.2, 0, 1)
            rel = f"stills/t{t}_k{k}.exr"
            render.write_exr(str(tmp_path / rel), a)
            stills[str(tmp_path / rel)] = a
            text = text.replace(workloads.still_url(t, k), rel)
    otio = tmp_path / "stack.otio"
    otio.write_text(text)
    out = tmp_path / "stack.raw"
    r = subprocess.run([os.path.join(capi.LIB_DIR, "toucan-render"), str(otio), str(out), "-raw", "rgbaf32"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    frames = np.fromfile(out, np.float32).reshape(80, h, w, 4)
    otl = G.Timeline(json.loads(text))
    otl.dir = str(tmp_path)
    og = G.ImageGraph(otl, lambda p: stills.get(p))
    times = otl.frames()
    for i in (0, 24, 40, 47, 60, 79):  # plain frames and frames inside dissolves
        want = og.render(times[i])
        assert float(np.abs(frames[i] - want).max()) <= 1e-5, i


# ---- TIFF stills -------------------------------------------------------------------------------------------------
def _associate_u8(rgba):
    """OIIO associates unassociated alpha on read (what Pillow writes for RGBA): v' = store_u8(alpha * v) in float."""
    f = rgba.astype(np.float32) * np.float32(1.0 / 255.0)
    al = f[..., 3:4]
    rgb = np.where(al != 1.0, al * f[..., :3], f[..., :3]).astype(np.float32)
    q = np.clip(rgb * np.float32(255.0) + np.float32(0.5), 0, 255).astype(np.uint8)
    return np.concatenate([q, rgba[..., 3:4]], -1)


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (64, 33), (301, 77)])
@pytest.mark.parametrize("channels", [3, 4])
def test_tiff_reader_matches_libtiff(built, tmp_path, size, channels):
    """8-bit files written by Pillow (libtiff): uncompressed, LZW, deflate and PackBits strips, with and without the
    horizontal-differencing predictor -- every sample as Pillow decodes it."""
    from toucan_b200 import render

    w, h = size
    rng = np.random.default_rng(3)
    for kind in ("noise", "ramp"):
        a = rng.integers(0, 256, (h, w, channels), dtype=np.uint8) if kind == "noise" else \
            np.ascontiguousarray(np.broadcast_to((np.arange(w)[None, :, None] * 5 + np.arange(h)[:, None, None] * 3) % 256, (h, w, channels)).astype(np.uint8))
        for compression in (None, "tiff_lzw", "tiff_adobe_deflate", "packbits"):
            for predictor in (None, 2):
                if predictor and compression not in ("tiff_lzw", "tiff_adobe_deflate"):
                    continue
                path = str(tmp_path / "t.tif")
                kw = {}
                if compression:
                    kw["compression"] = compression
                if predictor:
                    kw["tiffinfo"] = {317: 2}
                Image.fromarray(a).save(path, **kw)
                want = np.asarray(Image.open(path))
                want = np.concatenate([want, np.full((h, w, 1), 255, np.uint8)], -1) if channels == 3 else _associate_u8(want)
                got = render.decode_tiff(path)
                assert got.dtype == np.uint8 and np.array_equal(got, want), (size, channels, kind, compression, predictor)


def test_tiff_reader_16_bit_and_float(built, tmp_path):
    """16-bit (LZW + predictor, what OpenCV's libtiff writes) and 32-bit float files against OpenCV's decode."""
    cv2 = pytest.importorskip("cv2")
    from toucan_b200 import capi, render

    rng = np.random.default_rng(4)
    for dtype, scale in ((np.uint16, 65535.0), (np.float32, 2.0)):
        for ch in (3, 4):
            a = (rng.random((45, 83, ch)) * scale).astype(dtype)
            path = str(tmp_path / "c.tif")
            assert cv2.imwrite(path, a)
            b = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            want = b[..., [2, 1, 0, 3]] if ch == 4 else np.concatenate([b[..., ::-1], np.full((45, 83, 1), dtype(scale if dtype == np.uint16 else 1.0), dtype)], -1)
            got = render.decode_tiff(path)
            assert got.dtype == dtype and np.array_equal(got, want), (dtype, ch)
    with open(str(tmp_path / "junk.tif"), "wb") as f:
        f.write(b"II*\x00" + b"\x00" * 20)
    with pytest.raises(capi.TcError):
        render.decode_tiff(str(tmp_path / "junk.tif"))
    Image.fromarray(rng.integers(0, 256, (9, 9), dtype=np.uint8)).save(str(tmp_path / "gray.tif"))
    with pytest.raises(capi.TcError, match="RGB"):
        render.decode_tiff(str(tmp_path / "gray.tif"))


def _write_tiled_tiff(path, a, tile, deflate=False, predictor=False):
    """A minimal little-endian TILED TIFF (Pillow and OpenCV only write strips): RGB / RGBA, 8- or 16-bit unsigned,
    tiles padded to full size by repeating the edge, optional deflate and horizontal differencing."""
    import struct
    import zlib

    h, w, ch = a.shape
    tw, th = tile
    bps = a.dtype.itemsize
    across, down = (w + tw - 1) // tw, (h + th - 1) // th
    blocks = []
    for ty in range(down):
        for tx in range(across):
            t = a[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            t = np.pad(t, ((0, th - t.shape[0]), (0, tw - t.shape[1]), (0, 0)), mode="edge")
            if predictor:
                d = t.astype(np.int64)
                d[:, 1:] -= t[:, :-1].astype(np.int64)
                t = (d % (1 << (8 * bps))).astype(a.dtype)
            raw = t.astype("<u%d" % bps).tobytes()
            blocks.append(zlib.compress(raw) if deflate else raw)
    entries = [(256, 4, [w]), (257, 4, [h]), (258, 3, [8 * bps] * ch), (259, 3, [8 if deflate else 1]), (262, 3, [2]), (277, 3, [ch]),
               (284, 3, [1]), (317, 3, [2 if predictor else 1]), (322, 4, [tw]), (323, 4, [th]), (324, 4, None), (325, 4, [len(b) for b in blocks]),
               (339, 3, [1] * ch)]
    if ch == 4:
        entries.append((338, 3, [2]))  # unassociated alpha, what Pillow writes for RGBA
    entries.sort()
    size = {3: 2, 4: 4}
    ifd_at = 8
    extra_at = ifd_at + 2 + 12 * len(entries) + 4
    extra = b""
    slots = {}
    for tag, typ, vals in entries:
        n = len(blocks) if vals is None else len(vals)
        if n * size[typ] > 4:
            slots[tag] = extra_at + len(extra)
            extra += b"\0" * (n * size[typ] + (n * size[typ]) % 2)
    data_at = extra_at + len(extra)
    offsets, at = [], data_at
    for b in blocks:
        offsets.append(at)
        at += len(b) + len(b) % 2
    out = bytearray(b"II*\0" + struct.pack("<I", ifd_at) + struct.pack("<H", len(entries)))
    extra = bytearray(extra)
    for tag, typ, vals in entries:
        vals = offsets if vals is None else vals
        packed = b"".join(struct.pack("<H" if typ == 3 else "<I", v) for v in vals)
        if tag in slots:
            o = slots[tag] - extra_at
            extra[o:o + len(packed)] = packed
            field = struct.pack("<I", slots[tag])
        else:
            field = packed.ljust(4, b"\0")
        out += struct.pack("<HHI", tag, typ, len(vals)) + field
    out += struct.pack("<I", 0) + extra
    for b in blocks:
        out += b + b"\0" * (len(b) % 2)
    with open(path, "wb") as f:
        f.write(bytes(out))


@pytest.mark.parametrize("size,tile", [((101, 77), (64, 64)), ((64, 64), (64, 64)), ((33, 130), (16, 32)), ((5, 3), (16, 16)), ((300, 40), (128, 16))])
def test_tiff_reader_tiled(built, tmp_path, size, tile):
    """Tiled files (TileWidth / TileLength / TileOffsets): uncompressed and deflate tiles, with and without the predictor,
    8-bit RGB / RGBA and 16-bit RGB, edge tiles padded.  The fixture is checked against Pillow's (libtiff's) decode first."""
    from toucan_b200 import render

    w, h = size
    rng = np.random.default_rng(8)
    for dtype, ch in ((np.uint8, 3), (np.uint8, 4), (np.uint16, 3)):
        a = rng.integers(0, 256 if dtype == np.uint8 else 65536, (h, w, ch)).astype(dtype)
        for deflate in (False, True):
            for predictor in ((False, True) if deflate else (False,)):  # libtiff applies the predictor inside LZW / deflate only
                path = str(tmp_path / "tiled.tif")
                _write_tiled_tiff(path, a, tile, deflate, predictor)
                if dtype == np.uint8:  # Pillow has no 16-bit RGB mode: OpenCV's libtiff checks those
                    assert np.array_equal(np.asarray(Image.open(path)), a), ("fixture", size, tile, ch, deflate, predictor)
                else:
                    cv2 = pytest.importorskip("cv2")
                    assert np.array_equal(cv2.imread(path, cv2.IMREAD_UNCHANGED)[..., ::-1], a), ("fixture", size, tile, ch, deflate, predictor)
                if ch == 3:
                    want = np.concatenate([a, np.full((h, w, 1), np.iinfo(dtype).max, dtype)], -1)
                else:
                    want = _associate_u8(a)
                got = render.decode_tiff(path)
                assert got.dtype == dtype and np.array_equal(got, want), (size, tile, dtype, ch, deflate, predictor)


def _write_tiled_exr(path, rgba, tile, dtype, zip_tiles, mode=0):
    """A minimal single-part TILED OpenEXR file (OpenCV only writes scanline files): channels A, B, G, R of type HALF or
    FLOAT, level mode ONE_LEVEL (0) or MIPMAP_LEVELS (1: only the full-resolution level is stored first, which is all a
    level-0 reader needs -- OpenEXR itself wants the other levels, so that variant is not shown to it), tiles clipped at
    the right / bottom edge, stored raw or ZIP-compressed (byte split + predictor + zlib, kept only when smaller)."""
    import struct
    import zlib

    h, w, nch = rgba.shape
    tw, th = tile
    names = ["A", "B", "G", "R"] if nch == 4 else ["B", "G", "R"]
    plane = {"R": 0, "G": 1, "B": 2, "A": 3}
    px = 1 if dtype == np.float16 else 2

    def attr(name, typ, payload):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<I", len(payload)) + payload

    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", px, 0, 0, 0, 0, 1, 1) for n in names) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = struct.pack("<II", 20000630, 2 | 0x200)
    header += attr("channels", "chlist", chlist) + attr("compression", "compression", bytes([3 if zip_tiles else 0]))
    header += attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0")
    header += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
    header += attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + attr("tiles", "tiledesc", struct.pack("<IIB", tw, th, mode)) + b"\0"
    nx, ny = (w + tw - 1) // tw, (h + th - 1) // th
    chunks = []
    for ty in range(ny):
        for tx in range(nx):
            t = rgba[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            raw = b"".join(np.ascontiguousarray(t[l, :, plane[n]]).astype(dtype).tobytes() for l in range(t.shape[0]) for n in names)
            data = raw
            if zip_tiles:
                b = np.frombuffer(raw, np.uint8)
                split = np.concatenate([b[0::2], b[1::2]]).astype(np.int32)
                pred = split.copy()
                pred[1:] = (split[1:] - split[:-1] + 128 + 256) % 256
                z = zlib.compress(pred.astype(np.uint8).tobytes())
                data = z if len(z) < len(raw) else raw
            chunks.append(struct.pack("<iiiii", tx, ty, 0, 0, len(data)) + data)
    table_at = len(header)
    at = table_at + 8 * len(chunks)
    table = b""
    for c in chunks:
        table += struct.pack("<Q", at)
        at += len(c)
    with open(path, "wb") as f:
        f.write(header + table + b"".join(chunks))


@pytest.mark.parametrize("size,tile", [((37, 37), (16, 16)), ((64, 32), (64, 32)), ((70, 9), (32, 32)), ((5, 50), (8, 16))])
def test_exr_reader_tiled(built, tmp_path, size, tile):
    """Tiled single-part files: half / float, RGB / RGBA, raw and ZIP tiles, edge tiles clipped; the ONE_LEVEL fixtures are
    shown to the OpenEXR library (through OpenCV) first, and a MIPMAP_LEVELS file is read at its full-resolution level."""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    cv2 = pytest.importorskip("cv2")
    from toucan_b200 import render

    w, h = size
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for dtype in (np.float16, np.float32):
        for nch in (3, 4):
            # smooth ramps with a flat patch (so that ZIP tiles really shrink) plus negative and > 1 values
            a = np.stack([xx / w * 2.5 - 0.5, yy / h, (xx + yy) / (w + h) * 1.5, np.clip(xx / w, 0, 1)][:nch], -1).astype(np.float32)
            a[: h // 2, : w // 2] = 0.25
            a = a.astype(dtype).astype(np.float32)  # what the file will hold
            for zip_tiles in (False, True):
                path = str(tmp_path / "tiled.exr")
                _write_tiled_exr(path, a, tile, dtype, zip_tiles)
                ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)
                if ref is not None:  # OpenCV built with its OpenEXR codec
                    ref = ref[..., [2, 1, 0, 3]] if nch == 4 else ref[..., ::-1]
                    assert np.array_equal(ref.astype(np.float32), a), ("fixture", size, tile, dtype, nch, zip_tiles)
                want = a if nch == 4 else np.concatenate([a, np.ones((h, w, 1), np.float32)], -1)
                got = render.decode_exr(path)
                assert got.dtype == dtype and np.array_equal(got.astype(np.float32), want), (size, tile, dtype, nch, zip_tiles)
                _write_tiled_exr(path, a, tile, dtype, zip_tiles, mode=1)
                assert np.array_equal(render.decode_exr(path).astype(np.float32), want), ("mipmap level 0", size, tile, dtype, nch, zip_tiles)


def _write_png(path, a, interlaced, filter_type=0):
    """A PNG written by hand (Pillow and OpenCV do not write Adam7): 8- or 16-bit RGB / RGBA, every row with the given
    filter type (0 None, 1 Sub, 2 Up), optionally Adam7-interlaced (seven separately filtered sub-images)."""
    import struct
    import zlib

    h, w, ch = a.shape
    depth = 8 * a.dtype.itemsize
    big = a.astype(">u%d" % a.dtype.itemsize)

    def rows(img):
        out = b""
        prev = np.zeros(img.shape[1] * ch * a.dtype.itemsize, np.uint8)
        bpp = ch * a.dtype.itemsize
        for r in img:
            cur = np.frombuffer(r.tobytes(), np.uint8)
            if filter_type == 1:
                left = np.concatenate([np.zeros(bpp, np.uint8), cur[:-bpp]])
                enc = (cur.astype(np.int32) - left) % 256
            elif filter_type == 2:
                enc = (cur.astype(np.int32) - prev) % 256
            else:
                enc = cur
            out += bytes([filter_type]) + enc.astype(np.uint8).tobytes()
            prev = cur
        return out

    if interlaced:
        passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
        raw = b"".join(rows(big[y0::dy, x0::dx]) for x0, y0, dx, dy in passes if big[y0::dy, x0::dx].size)
    else:
        raw = rows(big)

    def chunk(kind, payload):
        return struct.pack(">I", len(payload)) + kind + payload + struct.pack(">I", zlib.crc32(kind + payload))

    ihdr = struct.pack(">IIBBBBB", w, h, depth, 6 if ch == 4 else 2, 0, 0, 1 if interlaced else 0)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


@pytest.mark.parametrize("size", [(1, 1), (3, 2), (8, 8), (9, 17), (67, 41)])
def test_png_reader_interlaced_and_filters(built, tmp_path, size):
    """Adam7-interlaced and plain files, filter types None / Sub / Up, 8-bit RGB / RGBA (checked against Pillow's decode
    of the same file, alpha associated like OIIO's reader) and 16-bit RGB."""
    from tests import media
    from toucan_b200 import render

    w, h = size
    rng = np.random.default_rng(12)
    for dtype, ch in ((np.uint8, 3), (np.uint8, 4), (np.uint16, 3)):
        a = rng.integers(0, 256 if dtype == np.uint8 else 65536, (h, w, ch)).astype(dtype)
        for interlaced in (False, True):
            for ft in (0, 1, 2):
                path = str(tmp_path / f"p_{np.dtype(dtype).name}_{ch}_{int(interlaced)}_{ft}.png")
                _write_png(path, a, interlaced, ft)
                if dtype == np.uint8:
                    assert np.array_equal(np.asarray(Image.open(path)), a), ("fixture", size, ch, interlaced, ft)
                    want = media.decode(path)
                else:
                    want = np.concatenate([a, np.full((h, w, 1), 65535, np.uint16)], -1)
                got = render.decode_png(path)
                assert got.dtype == dtype and np.array_equal(got, want), (size, dtype, ch, interlaced, ft)


@pytest.mark.parametrize("colors,bits", [(2, 1), (4, 2), (16, 4), (200, 8)])
def test_png_reader_palette_depths(built, tmp_path, colors, bits):
    """Palette files with 1, 2, 4 and 8 bits per index (Pillow writes them), with and without a tRNS chunk, at widths that
    do not fill the last byte of a row -- against Pillow's decode with OIIO's alpha association."""
    from tests import media
    from toucan_b200 import render

    rng = np.random.default_rng(13)
    for w, h in ((1, 1), (7, 5), (33, 9)):
        idx = rng.integers(0, colors, (h, w)).astype(np.uint8)
        pal = rng.integers(0, 256, (colors, 3)).astype(np.uint8)
        for transparent in (False, True):
            im = Image.fromarray(idx, mode="P")
            im.putpalette(pal.reshape(-1).tolist())
            path = str(tmp_path / f"pal_{w}_{int(transparent)}.png")
            kw = {"bits": bits}
            if transparent:
                kw["transparency"] = bytes(rng.integers(0, 256, colors).astype(np.uint8))
            im.save(path, **kw)
            chk = Image.open(path)
            assert chk.mode == "P"
            got = render.decode_png(path)
            assert got.dtype == np.uint8 and np.array_equal(got, media.decode(path)), (colors, bits, w, h, transparent)


def test_readers_survive_damaged_files(built, tmp_path):
    """tests/cpp/reader_fuzz.cpp under AddressSanitizer + UBSan: every EXR / PNG / JPEG fixture plus tiled, interlaced and
    palette files written here, each decoded whole, truncated and with random bytes flipped.  A reader may refuse a damaged
    file (false / an exception); it may not read or write out of bounds, overflow or allocate what a header merely claims."""
    import glob
    import subprocess

    from toucan_b200 import build as B

    rng = np.random.default_rng(0)
    u8 = rng.integers(0, 256, (37, 41, 3)).astype(np.uint8)
    f32 = rng.random((37, 41, 4)).astype(np.float32)
    files = sorted(glob.glob(os.path.join(EXR_DIR, "*.exr"))) + sorted(glob.glob(os.path.join(TL.DATA, "*.png"))) + sorted(glob.glob(os.path.join(TL.DATA, "*.jpg")))
    for dt, z in ((np.float16, False), (np.float32, True)):
        files.append(str(tmp_path / f"tiled_{np.dtype(dt).name}.exr"))
        _write_tiled_exr(files[-1], f32.astype(dt).astype(np.float32), (16, 16), dt, z)
    for deflate, predictor in ((False, False), (True, True)):
        files.append(str(tmp_path / f"tiled_{int(deflate)}.tif"))
        _write_tiled_tiff(files[-1], u8, (16, 16), deflate, predictor)
    files.append(str(tmp_path / "lzw.tif"))
    Image.fromarray(u8).save(files[-1], compression="tiff_lzw", tiffinfo={317: 2})
    files.append(str(tmp_path / "adam7.png"))
    _write_png(files[-1], u8, True, 1)
    exe = str(tmp_path / "reader_fuzz")
    src = [os.path.join(B.ROOT, "tests", "cpp", "reader_fuzz.cpp")] + [os.path.join(B.HOST, "toucanRender", f) for f in
                                                                        ("ExrRead.cpp", "TiffRead.cpp", "PngRead.cpp", "JpegRead.cpp")]
    r = subprocess.run([B.CXX, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I", os.path.join(B.ROOT, "include"),
                        "-I", B.HOST] + src + ["-o", exe, "-lz", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and ("asan" in r.stderr.lower() or "ubsan" in r.stderr.lower() or "fsanitize" in r.stderr):
        pytest.skip("no sanitizer runtime for " + B.CXX)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "60"] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "decoded" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])

