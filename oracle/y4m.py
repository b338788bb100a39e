"""CPU restatement of toucan-render's y4m output path.  TEST INFRASTRUCTURE ONLY:
nothing on the product path may import this module.

Reference: bin/toucan-render/App.cpp:37-43 (y4mSpecs), :302-338 (_writeY4mHeader), :340-485
(_writeY4mFrame): the rendered frame is pasted into a u8 (u16 for 444p16) RGB (RGBA for 444alpha)
buffer with ImageBufAlgo::paste, converted with libswscale's sws_scale_frame(..., SWS_FAST_BILINEAR)
to YUV422P / YUV444P / YUVA444P / YUV444P16 and the planes are written back to back after "FRAME\\n".

libswscale is NOT in the reference tree (FFmpeg 8.0, cmake/SuperBuild/BuildFFmpeg.cmake:390), so its
algorithm for these four conversions is restated here from the published sources (libswscale/input.c
rgb24ToY_c, rgb24ToUV_half_c, rgbaToA_c, rgb48ToY/UV_half templates; utils.c initFilter, the
SWS_FAST_BILINEAR xInc adjustment in sws_init_context; swscale.c hscale16to15_c / hscale16to19_c;
output.c yuv2plane1_8_c / yuv2plane1_16_c).  PARITY IS PINNED: tests/golden/make_y4m_golden.py ran
the real libswscale 9.1.100 (FFmpeg 8.0, the build bundled with opencv-python-headless in this image)
over seeded images; this restatement reproduces every byte (tests/test_y4m.py).
"""
from __future__ import annotations

import numpy as np

from . import oracle as O

FORMATS = ("422", "444", "444alpha", "444p16")

# ff_rgb2yuv coefficients for SWS_CS_DEFAULT (BT.601, limited range), RGB2YUV_SHIFT = 15
RY, GY, BY = 8414, 16519, 3208
RU, GU, BU = -4865, -9528, 14392
RV, GV, BV = 14392, -12061, -2332


def _rounded_div(a: int, b: int) -> int:
    return (a + (b >> 1)) // b if a >= 0 else -((-a + (b >> 1)) // b)


def init_filter(x_inc: int, src_w: int, dst_w: int, src_pos: int = 128, dst_pos: int = 128, align: int = 4):
    """utils.c initFilter(), the `flags & SWS_FAST_BILINEAR` branch, with one = 1 << 14 and the x86
    filter alignment of 4.  Returns (filterPos[dst_w], filter[dst_w, size])."""
    one = 1 << 14
    fone = 1 << 54  # srcW / dstW is 0 or 1 for every case here
    if abs(x_inc - 0x10000) < 10 and src_pos == dst_pos:  # "unscaled"
        return np.arange(dst_w), np.full((dst_w, 1), one, dtype=np.int64)
    size = 2
    coef = [[0, 0] for _ in range(dst_w)]
    pos = [0] * dst_w
    x = ((dst_pos * x_inc) >> 8) - ((src_pos * 0x8000) >> 7)
    for i in range(dst_w):
        xx = (x - ((size - 1) << 15) + (1 << 15)) >> 16
        pos[i] = xx
        for j in range(size):
            coef[i][j] = max(fone - abs(xx * 65536 - x) * (fone >> 16), 0)
            xx += 1
        x += x_inc
    # "try to reduce the filter-size": leading coefficients below SWS_MAX_REDUCE_CUTOFF are dropped
    cutoff = 0.002 * fone
    min_size = 0
    for i in range(dst_w - 1, -1, -1):
        acc = 0
        for _ in range(size):
            acc += abs(coef[i][0])
            if acc > cutoff:
                break
            if i < dst_w - 1 and pos[i] >= pos[i + 1]:
                break
            coef[i] = coef[i][1:] + [0]
            pos[i] += 1
        acc, mn = 0, size
        for j in range(size - 1, 0, -1):
            acc += abs(coef[i][j])
            if acc > cutoff:
                break
            mn -= 1
        min_size = max(min_size, mn)
    fsz = (min_size + align - 1) & ~(align - 1)
    filt = [[coef[i][j] if j < size else 0 for j in range(fsz)] for i in range(dst_w)]
    # "fix borders"
    for i in range(dst_w):
        f = filt[i]
        if pos[i] < 0:
            for j in range(1, fsz):
                left = max(j + pos[i], 0)
                f[left] += f[j]
                f[j] = 0
            pos[i] = 0
        if pos[i] + fsz > src_w:
            shift = pos[i] + min(fsz - src_w, 0)
            acc = 0
            for j in range(fsz - 1, -1, -1):
                if pos[i] + j >= src_w:
                    acc += f[j]
                    f[j] = 0
            for j in range(fsz - 1, -1, -1):
                f[j] = 0 if j < shift else f[j - shift]
            pos[i] -= shift
            f[src_w - 1 - pos[i]] += acc
    # normalise to `one` with error feedback
    out = np.zeros((dst_w, fsz), dtype=np.int64)
    for i in range(dst_w):
        s = (sum(filt[i]) + one // 2) // one or 1
        err = 0
        for j in range(fsz):
            v = filt[i][j] + err
            iv = _rounded_div(v, s)
            out[i, j] = iv
            err = v - iv * s
    return np.array(pos), out


def _hscale(src: np.ndarray, pos: np.ndarray, filt: np.ndarray, shift: int, cap: int) -> np.ndarray:
    """hscale16to15_c (shift 13 for RGB input) / hscale16to19_c (shift 11 for 16-bit RGB input)."""
    w = src.shape[1]
    acc = np.zeros((src.shape[0], len(pos)), dtype=np.int64)
    for j in range(filt.shape[1]):
        acc += src[:, np.minimum(pos + j, w - 1)] * filt[:, j][None, :]
    return np.minimum(acc >> shift, cap)


def x_increments(fmt: str, w: int):
    """(lumXInc, chrXInc, chroma source width, chroma plane width) as sws_init_context leaves them on x86-64
    for RGB input with SWS_FAST_BILINEAR: the MMXEXT scaler is never used for RGB input, so 8-bit outputs
    take the `((srcW - 2) << 16) / (dstW - 2) - 20` branch; 16-bit output keeps the plain increments."""
    cw_src = (w + 1) // 2  # chrSrcHSubSample = 1 for RGB input under SWS_FAST_BILINEAR
    cw_dst = cw_src if fmt == "422" else w
    if fmt == "444p16":
        return ((w << 16) + (w >> 1)) // w, ((cw_src << 16) + (cw_dst >> 1)) // cw_dst, cw_src, cw_dst
    return ((w - 2) << 16) // (w - 2) - 20, ((cw_src - 2) << 16) // (cw_dst - 2) - 20, cw_src, cw_dst


def sws_planes(rgb: np.ndarray, fmt: str):
    """The planes sws_scale_frame produces for a packed u8 RGB / RGBA (u16 RGB for 444p16) image."""
    h, w, _ = rgb.shape
    if w < 8 or w & 1:
        raise ValueError("even widths >= 8 (the reference reads past the row for odd widths)")
    lx, cx, cw_src, cw_dst = x_increments(fmt, w)
    lpos, lfilt = init_filter(lx, w, w)
    cpos, cfilt = init_filter(cx, cw_src, cw_dst)
    c = rgb[..., :3].astype(np.int64)
    r, g, b = c[..., 0], c[..., 1], c[..., 2]
    if fmt == "444p16":
        y = (RY * r + GY * g + BY * b + (0x2001 << 14)) >> 15
        s = (c[:, 0::2] + c[:, 1::2] + 1) >> 1
        u = (RU * s[..., 0] + GU * s[..., 1] + BU * s[..., 2] + (0x10001 << 14)) >> 15
        v = (RV * s[..., 0] + GV * s[..., 1] + BV * s[..., 2] + (0x10001 << 14)) >> 15

        def fin(p, pos, filt):
            return np.clip((_hscale(p, pos, filt, 11, (1 << 19) - 1) + 4) >> 3, 0, 65535).astype(np.uint16)
    else:
        y = (RY * r + GY * g + BY * b + (32 << 14) + (1 << 8)) >> 9
        s = c[:, 0::2] + c[:, 1::2]
        u = (RU * s[..., 0] + GU * s[..., 1] + BU * s[..., 2] + (256 << 15) + (1 << 9)) >> 10
        v = (RV * s[..., 0] + GV * s[..., 1] + BV * s[..., 2] + (256 << 15) + (1 << 9)) >> 10

        def fin(p, pos, filt):
            return np.clip((_hscale(p, pos, filt, 13, (1 << 15) - 1) + 64) >> 7, 0, 255).astype(np.uint8)
    planes = [fin(y, lpos, lfilt), fin(u, cpos, cfilt), fin(v, cpos, cfilt)]
    if fmt == "444alpha":
        a = rgb[..., 3].astype(np.int64)
        planes.append(fin((a << 6) | (a >> 2), lpos, lfilt))
    return planes


def frame_planes(frame: np.ndarray, fmt: str):
    """App.cpp:346-371: paste into the y4m spec's storage (when type or channel count differ), then swscale."""
    t = O.U16 if fmt == "444p16" else O.U8
    q = frame if O.tcode(frame) == t else O.convert(frame, t)
    return sws_planes(q if fmt == "444alpha" else q[..., :3], fmt)


def frame_bytes(frame: np.ndarray, fmt: str) -> bytes:
    """What _writeY4mFrame writes after "FRAME\\n" (little-endian samples for 444p16)."""
    return b"".join(np.ascontiguousarray(p).astype(p.dtype.newbyteorder("<")).tobytes() for p in frame_planes(frame, fmt))


def to_rational(rate: float):
    """ftk::toRational (feather-tk, not in the reference tree; restated from its published source): the common
    video rates map to their exact fractions, anything else to (int(rate), 1)."""
    for num, den in ((24, 1), (30, 1), (60, 1), (24000, 1001), (30000, 1001), (60000, 1001)):
        if abs(rate - num / den) < 0.01:
            return num, den
    return int(rate), 1


def header(w: int, h: int, rate: float, fmt: str) -> bytes:
    """App.cpp:302-338."""
    num, den = to_rational(rate)
    return f"YUV4MPEG2 W{w} H{h} F{num}:{den} C{fmt}\n".encode()
