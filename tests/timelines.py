"""Shared helpers for the timeline-level tests: the golden fixtures and the C++ host set-up."""
from __future__ import annotations

import json
import os

from tests import media

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATA = os.path.join(GOLDEN, "data")


SVG_STAND_INS = ("Letter_A.svg", "Letter_B.svg", "Letter_C.svg")


def golden():
    with open(os.path.join(GOLDEN, "frames.json")) as f:
        return json.load(f)


def open_timeline(name):
    """toucan::TimelineWrapper on a fixture with every raster medium of the data directory
    registered as a decoded frame (decode stays outside the hot path and identical for the
    oracle and the CUDA path)."""
    from toucan_b200 import render

    tl = render.Timeline(os.path.join(DATA, name))
    for f in sorted(os.listdir(DATA)):
        if f.lower().endswith((".png", ".jpg")):
            frame = media.decode(os.path.join(DATA, f))
            if frame is not None:
                tl.set_media(os.path.join(DATA, f), frame)
    for name in SVG_STAND_INS:  # the .svg media of Gap.otio: the PNG export stands in (tests/media.py)
        frame = media.decode(os.path.join(DATA, name))
        if frame is not None:
            tl.set_media(os.path.join(DATA, name), frame)
    tl.prepare()
    return tl
