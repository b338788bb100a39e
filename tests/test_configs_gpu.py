"""GPU parity at the sizes and storage types of BASELINE.json configs[1..3]: the
structure of the reference's own fixture timelines (data/Filter.otio, MultipleEffects.otio,
CompositeTracks.otio, Transform.otio, ColorSpace.otio) with synthetic media of the config's size and type
registered under the fixture's media paths.  The C++ host renders through ImageGraph + the OpenFX
modules; the oracle evaluates the same node tree.  (configs[0] and the shipped media are covered by
tests/test_timelines_gpu.py; configs[4] by the stack tests and bench.py.)
"""
import os

import numpy as np
import pytest

from oracle import graph as G
from tests import timelines as TL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host(cuda):
    from toucan_b200 import render

    render.bind_device(0)
    return render.Host()


def synth(w, h, dtype, seed, alpha="one"):
    rng = np.random.default_rng(seed)
    a = rng.random((h, w, 4), dtype=np.float32)
    if alpha == "one":
        a[..., 3] = 1.0
    else:  # radial soft disc
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        a[..., 3] = np.clip(1.3 - 2.2 * np.hypot((xx - w / 2) / w, (yy - h / 2) / h), 0.0, 1.0)
    return np.ascontiguousarray(a.astype(dtype))


def _resized(doc, size):
    """The fixture with every generator's `size` parameter replaced (the shipped ones say 1280x720)."""
    if isinstance(doc, dict):
        return {k: (list(size) if k == "size" and isinstance(v, list) and len(v) == 2 else _resized(v, size)) for k, v in doc.items()}
    if isinstance(doc, list):
        return [_resized(v, size) for v in doc]
    return doc


def run(host, name, media, frames, check, generator_size=None):
    import json

    from toucan_b200 import render

    doc = json.load(open(os.path.join(TL.DATA, name)))
    if generator_size:
        doc = _resized(doc, generator_size)
    tl = render.Timeline(json_doc=doc, directory=TL.DATA)
    frames_by_path = {}
    for fname, frame in media.items():
        p = os.path.join(TL.DATA, fname)
        frames_by_path[p] = frame
        tl.set_media(p, frame)
    tl.prepare()
    otl = G.Timeline(doc)
    otl.dir = TL.DATA
    og = G.ImageGraph(otl, lambda p: frames_by_path.get(p))
    first = next(iter(media.values()))
    assert og.size == (first.shape[1], first.shape[0])
    times = otl.frames()
    for f in frames:
        t = times[f]
        got = tl.render(host, t.value, max_bytes=first.shape[0] * first.shape[1] * 16)
        want = og.render(t)
        assert got.shape == want.shape and got.dtype == want.dtype, (name, f, got.shape, want.shape, got.dtype, want.dtype)
        check(name, f, got, want)
    tl.close()


def close_f32(name, f, got, want):
    d = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max())
    assert d <= 1e-5, f"{name} frame {f}: max abs diff {d:.3e}"


def test_c2_filter_set_1080p_float(host):
    # ColorMap, Invert, Pow 2, Saturate 0, Blur 10, UnsharpMask(gaussian, 10, 1, 0) at 1920x1080 RGBA float32
    media = {"Charlie.jpg": synth(1920, 1080, np.float32, 1001)}
    run(host, "Filter.otio", media, range(6), close_f32)


def test_c3_stack_and_effects_4k_half(host):
    import json

    from tests.envelope import assert_within_envelope, envelope, ordered

    media = {"Charlie.jpg": synth(3840, 2160, np.float16, 1002)}
    doc = json.load(open(os.path.join(TL.DATA, "MultipleEffects.otio")))
    otl = G.Timeline(doc)
    otl.dir = TL.DATA
    by_path = {os.path.join(TL.DATA, k): v for k, v in media.items()}

    def close_half(name, f, got, want):
        # half storage is re-quantised at every node: Blur / ColorMap stores may flip by one code (+-1 LSB, north_star's bar),
        # Invert / over / Pow 3 downstream are exact given the flip -- the frame must lie inside the flip envelope
        # (tests/envelope.py), +-1 LSB for Pow itself; and flips are rare
        lo, hi, base = envelope(otl, lambda p: by_path.get(p), otl.frames()[f], alpha_varies=False)
        assert np.array_equal(base, want)
        assert_within_envelope(got, lo, hi, f"{name} frame {f}")
        assert float((np.abs(ordered(got) - ordered(want)) > 1).mean()) < 1e-3

    run(host, "MultipleEffects.otio", media, [0, 5], close_half)
    seq = {f"Counter.{i}.png": synth(3840, 2160, np.float16, 1100 + i, alpha="disc") for i in (0, 5)}

    def exactish(name, f, got, want):
        # premult + over of half frames only: every op on this path is held to +-1 LSB of half storage
        assert int(np.abs(ordered(got) - ordered(want)).max()) <= 1, f"{name} frame {f}"

    run(host, "CompositeTracks.otio", seq, [0, 5], exactish, generator_size=(3840, 2160))


def test_c4_transform_and_colorspace_4k_float(host):
    media = {"Charlie.jpg": synth(3840, 2160, np.float32, 1003)}
    # Crop, Flip, Flop (bit-exact class), Resize -> 640x360 + CompNode fit-resize back up, Rotate 45
    def check(name, f, got, want):
        if f in (2, 4):
            assert np.array_equal(got, want), f"{name} frame {f}: Flip / Flop must be bit-exact"
        elif got.dtype == np.uint8:
            # Crop / Resize outputs are smaller than the background: CompNode fit-resizes them and pastes
            # into a buffer of the background's (UINT8 generator) spec (Comp.cpp:44-68) -> +-1 LSB
            assert int(np.abs(got.astype(np.int64) - want.astype(np.int64)).max()) <= 1, f"{name} frame {f}"
        else:
            close_f32(name, f, got, want)

    run(host, "Transform.otio", media, [0, 2, 4, 6, 8], check)
    run(host, "ColorSpace.otio", media, [0, 3, 6], close_f32)


def test_c2_chained_filters_fuse_into_fewer_launches(cuda):
    """SURVEY 8d config 2, chained variant: one clip carrying ColorMap -> Invert -> Pow -> Saturate -> Blur -> UnsharpMask
    at 1920x1080 float32.  The four leading pointwise nodes run as ONE kernel (the plug-ins describe their operation
    to the host, which launches tc_pointwise_chain); frames are bit-identical to the node-by-node path and within
    1e-5 of the oracle."""
    import torch

    from oracle import graph as G
    from toucan_b200 import capi, render, workloads

    w, h = 1920, 1080
    rng = np.random.default_rng(1001)
    src = rng.random((h, w, 4), dtype=np.float32)
    src[..., 3] = 1.0
    doc = workloads.filter_chain_timeline(frames=6)
    render.bind_device(0)
    host = render.Host()
    tl = render.Timeline(json_doc=doc, directory="/synthetic")
    path = tl.media_path("synthetic://chain/source.exr")
    tl.set_media(path, src)
    tl.prepare()
    n0 = capi.launch_count()
    fused = tl.render(host, 2.0)
    fused_launches = capi.launch_count() - n0
    render.set_fusion(False)
    try:
        n0 = capi.launch_count()
        plain = tl.render(host, 2.0)
        plain_launches = capi.launch_count() - n0
    finally:
        render.set_fusion(True)
    assert fused.dtype == np.float32 and np.array_equal(fused, plain)
    # 4 pointwise launches became 1, the Fill background is composited as a colour instead of a rendered frame, and that
    # composite is the UnsharpMask kernel's epilogue: 8 launches -> 3 (chain, Blur, UnsharpMask + over)
    assert (plain_launches, fused_launches) == (8, 3), (plain_launches, fused_launches)
    otl = G.Timeline(doc)
    otl.dir = "/synthetic"
    want = G.ImageGraph(otl, lambda p: src if p == path else None).render(G.RT(2.0, 24.0))
    assert float(np.abs(fused - want).max()) <= 1e-5
    tl.close()
    del torch
