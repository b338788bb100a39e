"""ctypes front-end of the CPU oracle (oracle/toucan_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of toucan_oracle.c.  Nothing under
toucan_b200/ may import this module.  Parity: coarsely pinned by the reference's own film strips
(tests/test_reference_filmstrips.py); [V] below thumbnail resolution.

Images are numpy arrays of shape (h, w, 4) and dtype uint8 / uint16 / float16 / float32.
Every function returns a new array and cites the reference call site it restates.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

U8, U16, F16, F32 = 0, 1, 2, 3
_DT = {np.dtype(np.uint8): U8, np.dtype(np.uint16): U16, np.dtype(np.float16): F16, np.dtype(np.float32): F32}
NP_OF = {U8: np.uint8, U16: np.uint16, F16: np.float16, F32: np.float32}
NAME_OF = {U8: "u8", U16: "u16", F16: "half", F32: "float"}


class _OImg(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("w", ctypes.c_int), ("h", ctypes.c_int), ("type", ctypes.c_int)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "toucan_oracle.c")
    inc = os.path.join(_HERE, "colormap_tables.inc")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(inc)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.o_det_log.restype = ctypes.c_double
        _lib.o_det_log.argtypes = [ctypes.c_double]
        _lib.o_fast_exp.restype = ctypes.c_float
        _lib.o_fast_exp.argtypes = [ctypes.c_float]
        _lib.o_fast_sinpi.restype = ctypes.c_float
        _lib.o_fast_sinpi.argtypes = [ctypes.c_float]
    return _lib


def _img(a: np.ndarray) -> _OImg:
    assert a.ndim == 3 and a.shape[2] == 4 and a.flags.c_contiguous, (a.shape, a.flags)
    return _OImg(a.ctypes.data, a.shape[1], a.shape[0], _DT[a.dtype])


def new(w: int, h: int, t) -> np.ndarray:
    """ImageBuf(ImageSpec(w,h,4,type)) -- zero filled."""
    dt = NP_OF[t] if isinstance(t, int) else t
    return np.zeros((h, w, 4), dtype=dt)


def tcode(a: np.ndarray) -> int:
    return _DT[a.dtype]


def _f4(c):
    return (ctypes.c_float * 4)(*[float(v) for v in c])


def type_merge(a: int, b: int) -> int:
    return lib().o_type_merge(a, b)


# --- generators (GeneratorPlugin.cpp) -------------------------------------------------
def fill(w, h, color, t=U8):
    out = new(w, h, t)
    lib().o_fill(ctypes.byref(_img(out)), _f4(color))
    return out


def gradient(w, h, c1, c2, vertical, t=U8):
    out = new(w, h, t)
    if vertical:
        lib().o_fill_tb(ctypes.byref(_img(out)), _f4(c1), _f4(c2))
    else:
        lib().o_gradient_h(ctypes.byref(_img(out)), _f4(c1), _f4(c2))
    return out


def checkers(w, h, cw, ch, c1, c2, t=U8):
    out = new(w, h, t)
    lib().o_checker(ctypes.byref(_img(out)), int(cw), int(ch), _f4(c1), _f4(c2))
    return out


NOISE_TYPES = {"gaussian": 0, "normal": 0, "white": 1, "uniform": 1, "salt": 2}


def noise(w, h, ntype, a, b, mono, seed, t=U8):
    out = new(w, h, t)
    if ntype not in NOISE_TYPES:
        return out  # OIIO noise() fails on an unknown type: buffer stays zero
    lib().o_noise(ctypes.byref(_img(out)), NOISE_TYPES[ntype], ctypes.c_float(a), ctypes.c_float(b), int(bool(mono)), int(seed))
    return out


# --- filters (FilterPlugin.cpp) --------------------------------------------------------
def _unary(fn, src, *args, size=None):
    h, w = src.shape[:2]
    if size is not None:
        w, h = size
    out = new(w, h, src.dtype)
    fn(ctypes.byref(_img(out)), ctypes.byref(_img(src)), *args)
    return out


def invert(src):
    return _unary(lib().o_invert, src)


def pow_(src, value):
    return _unary(lib().o_pow, src, ctypes.c_float(value))


def saturate(src, value):
    return _unary(lib().o_saturate, src, ctypes.c_float(value))


def color_map(src, name):
    return _unary(lib().o_color_map, src, name.encode())


def premult(src):
    return _unary(lib().o_premult, src)


def unpremult(src):
    return _unary(lib().o_unpremult, src)


def blur(src, radius):
    return _unary(lib().o_blur, src, ctypes.c_float(radius))


def unsharp_mask(src, kernel, width, contrast, threshold):
    if kernel not in ("gaussian",):
        raise NotImplementedError("only the gaussian kernel is in scope")
    return _unary(lib().o_unsharp, src, ctypes.c_float(width), ctypes.c_float(contrast), ctypes.c_float(threshold))


def gauss_kernel(width):
    n = lib().o_make_gauss_kernel(ctypes.c_float(width), None, 0)
    k = np.zeros((n, n), np.float32)
    lib().o_make_gauss_kernel(ctypes.c_float(width), k.ctypes.data_as(ctypes.c_void_p), n * n)
    return k


# --- draw (DrawPlugin.cpp) --------------------------------------------------------------
def draw_box(src, pos1, pos2, color, fill=True):
    c = (ctypes.c_float * 4)(*[float(v) for v in color])
    return _unary(lib().o_draw_box, src, int(pos1[0]), int(pos1[1]), int(pos2[0]), int(pos2[1]), c, 1 if fill else 0)


def draw_line(src, pos1, pos2, color, skip_first_point=False):
    c = (ctypes.c_float * 4)(*[float(v) for v in color])
    return _unary(lib().o_draw_line, src, int(pos1[0]), int(pos1[1]), int(pos2[0]), int(pos2[1]), c, 1 if skip_first_point else 0)


# --- transforms (TransformPlugin.cpp) --------------------------------------------------
def crop(src, pos, size):
    return _unary(lib().o_crop, src, int(pos[0]), int(pos[1]), int(size[0]), int(size[1]), size=size)


def flip(src):
    return _unary(lib().o_flip, src)


def flop(src):
    return _unary(lib().o_flop, src)


FILTERS = {"": -1, "lanczos3": 0, "blackman-harris": 1}


def resize(src, size, filter_name="", filter_width=0.0, out_type=None):
    out = new(size[0], size[1], src.dtype if out_type is None else out_type)
    lib().o_resize(ctypes.byref(_img(out)), ctypes.byref(_img(src)), FILTERS[filter_name], ctypes.c_float(filter_width))
    return out


def rotate(src, angle_degrees, filter_name="", filter_width=0.0):
    assert filter_name in ("", "lanczos3") and filter_width in (0.0, 6.0)
    # TransformPlugin.cpp:465: angle / 360.F * 2.F * M_PI in double, narrowed to float by
    # the ImageBufAlgo::rotate(float angle) signature.
    ang = np.float32(float(angle_degrees) / 360.0 * 2.0 * math.pi)
    return _unary(lib().o_rotate, src, ctypes.c_float(ang))


# --- compositing / transitions ----------------------------------------------------------
def fit(a, b):
    out = (ctypes.c_int * 4)()
    lib().o_fit(int(a[0]), int(a[1]), int(b[0]), int(b[1]), out)
    return tuple(out)  # minx, miny, w, h


def paste(dst, ox, oy, src):
    lib().o_paste(ctypes.byref(_img(dst)), int(ox), int(oy), ctypes.byref(_img(src)))
    return dst


def over(fg, bg):
    h, w = bg.shape[:2]
    out = new(w, h, type_merge(tcode(fg), tcode(bg)))
    lib().o_over(ctypes.byref(_img(out)), ctypes.byref(_img(fg)), ctypes.byref(_img(bg)))
    return out


def comp(fg, bg, premult_=True):
    """CompNode::exec -- Comp.cpp:29-84."""
    fg_ok = fg is not None and fg.shape[0] > 0 and fg.shape[1] > 0
    bg_ok = bg is not None and bg.shape[0] > 0 and bg.shape[1] > 0
    if fg is not None and bg is not None:
        if premult_ and fg_ok:
            fg = premult(fg)
        if fg_ok and bg_ok and fg.shape[:2] != bg.shape[:2]:
            bw, bh = bg.shape[1], bg.shape[0]
            box = fit((bw, bh), (fg.shape[1], fg.shape[0]))
            resized = resize(fg, (box[2], box[3]))
            fg = new(bw, bh, bg.dtype)
            paste(fg, box[0], box[1], resized)
        if fg_ok:
            return over(fg, bg)
        return bg
    only = fg if fg is not None else bg
    if only is not None and premult_ and only.size:
        return premult(only)
    return only


def dissolve(frm, to, value, size=None):
    """DissolvePlugin::_render -- TransitionPlugin.cpp:192-267 (incl. the `&& ||` precedence
    of :203-206: the resize branch is taken whenever the heights differ, or both images
    are non-empty and the widths differ)."""
    fw, fh = frm.shape[1], frm.shape[0]
    tw, th = to.shape[1], to.shape[0]
    if (fw > 0 and fh > 0 and tw > 0 and th > 0 and tw != fw) or th != fh:
        fg_aspect = tw / float(th)
        bg_aspect = fw / float(fh)
        if fg_aspect > bg_aspect:
            width = fw
            height = int(width / fg_aspect)
        else:
            height = fh
            width = int(height * fg_aspect)
        resized = resize(to, (width, height))
        tmp = new(fw, fh, frm.dtype)
        paste(tmp, fw // 2 - width // 2, fh // 2 - height // 2, resized)
        to = tmp
    ow, oh = (fw, fh) if size is None else size
    out = new(ow, oh, frm.dtype)
    lib().o_dissolve(ctypes.byref(_img(out)), ctypes.byref(_img(frm)), ctypes.byref(_img(to)), ctypes.c_double(value))
    return out


def wipe(frm, to, value, vertical, size=None):
    ow, oh = (frm.shape[1], frm.shape[0]) if size is None else size
    out = new(ow, oh, frm.dtype)
    lib().o_wipe(ctypes.byref(_img(out)), ctypes.byref(_img(frm)), ctypes.byref(_img(to)), ctypes.c_double(value), int(vertical))
    return out


def convert(src, t, size=None):
    h, w = src.shape[:2]
    if size is not None:
        w, h = size
    out = new(w, h, t)
    lib().o_convert(ctypes.byref(_img(out)), ctypes.byref(_img(src)))
    return out


# --- ColorConvert: colour-space table of OCIO's built-in CG config [V] ------------------
_BRADFORD = np.array([[0.8951, 0.2664, -0.1614], [-0.7502, 1.7135, 0.0367], [0.0389, -0.0685, 1.0296]])
_W_ACES = (0.32168, 0.33767)
_W_D65 = (0.3127, 0.3290)
_PRIMS = {
    "ap0": ((0.7347, 0.2653), (0.0, 1.0), (0.0001, -0.077), _W_ACES),
    "ap1": ((0.713, 0.293), (0.165, 0.830), (0.128, 0.044), _W_ACES),
    "rec709": ((0.64, 0.33), (0.30, 0.60), (0.15, 0.06), _W_D65),
    "p3d65": ((0.680, 0.320), (0.265, 0.690), (0.150, 0.060), _W_D65),
    "rec2020": ((0.708, 0.292), (0.170, 0.797), (0.131, 0.046), _W_D65),
}
CURVE_LINEAR, CURVE_GAMMA, CURVE_SRGB, CURVE_ACESCCT, CURVE_ACESCC = 0, 1, 2, 3, 4
# name -> (curve id, curve param, primaries)
COLOR_SPACES = {
    "ACES2065-1": (CURVE_LINEAR, 1.0, "ap0"),
    "ACEScg": (CURVE_LINEAR, 1.0, "ap1"),
    "ACEScct": (CURVE_ACESCCT, 1.0, "ap1"),
    "ACEScc": (CURVE_ACESCC, 1.0, "ap1"),
    "Linear Rec.709 (sRGB)": (CURVE_LINEAR, 1.0, "rec709"),
    "Linear P3-D65": (CURVE_LINEAR, 1.0, "p3d65"),
    "Linear Rec.2020": (CURVE_LINEAR, 1.0, "rec2020"),
    "Gamma 1.8 Rec.709 - Texture": (CURVE_GAMMA, 1.8, "rec709"),
    "Gamma 2.2 Rec.709 - Texture": (CURVE_GAMMA, 2.2, "rec709"),
    "Gamma 2.4 Rec.709 - Texture": (CURVE_GAMMA, 2.4, "rec709"),
    "Gamma 2.2 AP1 - Texture": (CURVE_GAMMA, 2.2, "ap1"),
    "sRGB - Texture": (CURVE_SRGB, 1.0, "rec709"),
    "sRGB Encoded Rec.709 (sRGB)": (CURVE_SRGB, 1.0, "rec709"),
    "sRGB Encoded AP1 - Texture": (CURVE_SRGB, 1.0, "ap1"),
    "sRGB Encoded P3-D65 - Texture": (CURVE_SRGB, 1.0, "p3d65"),
}


def _xyz(xy):
    return np.array([xy[0] / xy[1], 1.0, (1.0 - xy[0] - xy[1]) / xy[1]])


def _npm(prims):
    r, g, b, w = prims
    P = np.stack([_xyz(r), _xyz(g), _xyz(b)], axis=1)
    S = np.linalg.solve(P, _xyz(w))
    return P * S[None, :]


def _cat(src_w, dst_w):
    if src_w == dst_w:
        return np.eye(3)
    s = _BRADFORD @ _xyz(src_w)
    d = _BRADFORD @ _xyz(dst_w)
    return np.linalg.inv(_BRADFORD) @ np.diag(d / s) @ _BRADFORD


def _to_ap0(prim):
    p = _PRIMS[prim]
    return np.linalg.inv(_npm(_PRIMS["ap0"])) @ _cat(p[3], _W_ACES) @ _npm(p)


def color_matrix(from_prim, to_prim):
    """3x3 (column-vector convention, row-major) taking from_prim RGB to to_prim RGB via ACES2065-1."""
    if from_prim == to_prim:
        return np.eye(3, dtype=np.float32)
    return (np.linalg.inv(_to_ap0(to_prim)) @ _to_ap0(from_prim)).astype(np.float32)


def colorconvert(src, fromspace, tospace, unpremult=False):
    """ColorConvertPlugin::_render -- ColorPlugin.cpp:200-250 with the effective premult flag 0
    (SURVEY 8b).  Unknown colour-space names make colorconvert() fail: the output buffer
    keeps its zero fill."""
    h, w = src.shape[:2]
    out = new(w, h, src.dtype)
    if fromspace not in COLOR_SPACES or tospace not in COLOR_SPACES:
        return out
    ci, pi, fp = COLOR_SPACES[fromspace]
    co, po, tp = COLOR_SPACES[tospace]
    m = np.ascontiguousarray(color_matrix(fp, tp), dtype=np.float32)
    lib().o_colorconvert2(ctypes.byref(_img(out)), ctypes.byref(_img(src)), ci, ctypes.c_float(pi),
                          m.ctypes.data_as(ctypes.c_void_p), co, ctypes.c_float(po), int(bool(unpremult)))
    return out
