"""Decode stage of the test/bench harness (outside the hot path).

Media are decoded with Pillow and handed to BOTH the oracle and the CUDA path as raw
RGBA arrays, so decoder differences never enter a parity comparison.  The two read-side
conventions of the reference are replicated here:
  * RGB media get A = 1 appended (lib/toucanRender/Read.cpp:86-94);
  * OIIO's PNG reader associates alpha on read [V: pnginput, default png:linear_premult=0]:
    v' = store_u8(alpha * v) evaluated in float -- unless the "oiio:UnassociatedAlpha" open hint is set.  The
    reference opens its media without hints (Read.cpp:39), so association is the default here; the reference's own
    film strips of the transition timelines show UNassociated frames (tests/test_reference_filmstrips.py), so the
    convention is a switch: decode(path, associate=False), the product's tr_set_unassociated_alpha(1).
"""
from __future__ import annotations

import os

import numpy as np

_CACHE: dict[tuple, np.ndarray | None] = {}


def decode_unassociated(path: str) -> np.ndarray | None:
    return decode(path, associate=False)


def decode(path: str, associate: bool = True) -> np.ndarray | None:
    """Return the decoded frame as (h, w, 4) uint8/uint16, or None if unreadable."""
    key = (path, associate)
    if key in _CACHE:
        return _CACHE[key]
    out = None
    ext = os.path.splitext(path)[1].lower()
    if ext == ".svg":
        # No SVG rasteriser here (the reference uses lunasvg, Read.cpp:223-287, whose bitmap is premultiplied).  The shipped
        # <name>.png next to each data/<name>.svg is an Inkscape export of the same drawing at the same size: it stands in,
        # always associated, so that data/Gap.otio renders its letters and gaps instead of nothing.
        png = os.path.splitext(path)[0] + ".png"
        out = decode(png, True) if os.path.exists(path) or os.path.exists(png) else None
        _CACHE[key] = out
        return out
    if os.path.exists(path) and ext in (".png", ".jpg", ".jpeg", ".tif", ".tiff"):
        from PIL import Image

        im = Image.open(path)
        if im.mode in ("RGBA", "LA") or (im.mode == "P" and "transparency" in im.info):
            a = np.asarray(im.convert("RGBA"), dtype=np.uint8).copy()
            if ext == ".png" and associate:
                f = a.astype(np.float32) * np.float32(1.0 / 255.0)
                al = f[..., 3:4]
                rgb = np.where(al != 1.0, al * f[..., :3], f[..., :3]).astype(np.float32)
                q = np.clip(rgb * np.float32(255.0) + np.float32(0.5), 0, 255).astype(np.uint8)
                a[..., :3] = q
            out = a
        else:
            rgb = np.asarray(im.convert("RGB"), dtype=np.uint8)
            out = np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2)
        out = np.ascontiguousarray(out)
    _CACHE[key] = out
    return out
