"""Parity bar for chains that re-quantise at node boundaries (u8 / u16 / half storage).

north_star's bar is +-1 LSB at the output of a filtering / resampling / colour op.  Such an op cannot agree with the
reference's float result bit for bit (a separable FMA Gaussian and OIIO's 2-D sum differ by ~1e-7), so about one store in
10^4 rounds to the neighbouring code -- and every op DOWNSTREAM of that store sees a different input.  `1 - x` turns one
half LSB at x = 0.999 into hundreds of LSBs of the small result, `x^3` triples it: a bound on the final frame in its own
LSBs would either be wrong or meaningless.  The bar that IS exact: the kernel's frame must lie, value by value, between
the frames the ORACLE produces when the tolerant ops' stores are moved by -1 / 0 / +1 code (colour and alpha moved
independently), widened by the +-1 LSB the last tolerant op is itself allowed.  That statement says precisely "the only
disagreement is a +-1 LSB flip at the stores north_star allows; everything downstream of a flip is computed exactly".
It holds pixel-wise for pointwise downstream ops and for spatial ops with non-negative weights (Blur, over, Dissolve).
"""
from __future__ import annotations

import itertools

import numpy as np

from oracle import graph as G

TOLERANT_OPS = {"toucan:Blur", "toucan:UnsharpMask", "toucan:Resize", "toucan:Rotate", "toucan:ColorConvert", "toucan:ColorMap",
                "toucan:Pow", "toucan:Saturate"}


def ordered(a: np.ndarray) -> np.ndarray:
    """Storage codes as integers whose order is the values' order (so a difference is a count of LSBs)."""
    if a.dtype == np.float16:
        i = a.view(np.int16).astype(np.int64)
        return np.where(i < 0, -(i & 0x7FFF), i)
    if a.dtype == np.float32:
        i = a.view(np.int32).astype(np.int64)
        return np.where(i < 0, -(i & 0x7FFFFFFF), i)
    return a.astype(np.int64)


def shifted(a: np.ndarray, k_colour: int, k_alpha: int) -> np.ndarray:
    """The frame with every colour code moved by k_colour LSBs and every alpha code by k_alpha (clamped to the type's range)."""
    if a.dtype == np.float32 or (k_colour == 0 and k_alpha == 0):
        return a
    k = np.array([k_colour, k_colour, k_colour, k_alpha], np.int64)
    o = ordered(a) + k
    if a.dtype == np.float16:
        o = np.clip(o, -0x7BFF, 0x7BFF)  # stay finite
        return np.where(o < 0, (-o) | 0x8000, o).astype(np.uint16).view(np.float16).reshape(a.shape)
    info = np.iinfo(a.dtype)
    return np.clip(o, info.min, info.max).astype(a.dtype)


class FlipGraph(G.ImageGraph):
    """The oracle's graph with the stores of the tolerant ops moved by (k_colour, k_alpha) codes."""

    def __init__(self, timeline, media, k_colour=0, k_alpha=0, ops=TOLERANT_OPS, skip_last=True):
        super().__init__(timeline, media)
        self.k = (k_colour, k_alpha)
        self.ops = ops

    def eval_effect(self, name, meta, ins):
        out = super().eval_effect(name, meta, ins)
        if name in self.ops and out is not None and out.size:
            out = np.ascontiguousarray(shifted(out, *self.k))
        return out


def envelope(timeline, media, time, alpha_varies=True):
    """(lo, hi) of ordered codes over the flip variants of one frame, and the unshifted frame."""
    ks = list(itertools.product((-1, 0, 1), (-1, 0, 1) if alpha_varies else (0,)))
    lo = hi = base = None
    for kc, ka in ks:
        img = FlipGraph(timeline, media, kc, ka).render(time)
        o = ordered(img)
        lo = o if lo is None else np.minimum(lo, o)
        hi = o if hi is None else np.maximum(hi, o)
        if kc == 0 and ka == 0:
            base = img
    return lo, hi, base


def assert_within_envelope(got, lo, hi, what="", slack=1):
    g = ordered(got)
    bad = (g < lo - slack) | (g > hi + slack)
    if bad.any():
        i = tuple(np.argwhere(bad)[0])
        raise AssertionError(f"{what}: {int(bad.sum())} values outside the +-1 LSB flip envelope; first at {i}: got code {int(g[i])}, "
                             f"envelope [{int(lo[i])}, {int(hi[i])}]")
