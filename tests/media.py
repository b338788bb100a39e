"""Decode stage of the test/bench harness (outside the hot path).

Media are decoded with Pillow and handed to BOTH the oracle and the CUDA path as raw
RGBA arrays, so decoder differences never enter a parity comparison.  The two read-side
conventions of the reference are replicated here:
  * RGB media get A = 1 appended (lib/toucanRender/Read.cpp:86-94);
  * OIIO's PNG reader associates alpha on read [V: pnginput, default png:linear_premult=0]:
    v' = store_u8(alpha * v) evaluated in float.
"""
from __future__ import annotations

import os

import numpy as np

_CACHE: dict[str, np.ndarray | None] = {}


def decode(path: str) -> np.ndarray | None:
    """Return the decoded frame as (h, w, 4) uint8/uint16, or None if unreadable."""
    if path in _CACHE:
        return _CACHE[path]
    out = None
    ext = os.path.splitext(path)[1].lower()
    if os.path.exists(path) and ext in (".png", ".jpg", ".jpeg", ".tif", ".tiff"):
        from PIL import Image

        im = Image.open(path)
        if im.mode in ("RGBA", "LA") or (im.mode == "P" and "transparency" in im.info):
            a = np.asarray(im.convert("RGBA"), dtype=np.uint8).copy()
            if ext == ".png":
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
    _CACHE[path] = out
    return out
