"""Golden vectors for the y4m output conversion, produced by the REAL libswscale.

Run in the authoring container:
    python tests/golden/make_y4m_golden.py

The reference converts frames for `toucan-render -y4m` with libswscale from FFmpeg 8.0
(cmake/SuperBuild/BuildFFmpeg.cmake:390; call site bin/toucan-render/App.cpp:373-412:
sws_getContext(w, h, RGB24|RGBA|RGB48, w, h, YUV422P|YUV444P|YUVA444P|YUV444P16, SWS_FAST_BILINEAR)).
opencv-python-headless in this image bundles exactly that libswscale (9.1.100 = FFmpeg 8.0); this script
drives it through ctypes on seeded images and writes

  tests/golden/y4m_small.npz   inputs and output planes for small images (every byte), all four formats
  tests/golden/y4m_large.json  sha256 of the output planes for seeded 1280x720 / 1920x8 / 3840x4 / 7680x2 inputs
                               (the widths where the filter phase drifts furthest)

so the vectors pin oracle/y4m.py and the CUDA kernel to the real library, not to a restatement.
"""
import ctypes
import glob
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SWS_FAST_BILINEAR = 1
FORMATS = {"422": ("rgb24", "yuv422p"), "444": ("rgb24", "yuv444p"), "444alpha": ("rgba", "yuva444p"), "444p16": ("rgb48le", "yuv444p16le")}
SMALL = [(8, 3), (10, 2), (34, 5), (64, 4), (96, 7), (130, 3), (258, 2)]
LARGE = [(1280, 720), (1920, 8), (3840, 4), (7680, 2), (1022, 6), (2050, 3)]


def load():
    import cv2  # noqa: F401  resolves the bundled FFmpeg libraries' own dependencies

    d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    av = ctypes.CDLL(glob.glob(os.path.join(d, "libavutil-*"))[0])
    sws = ctypes.CDLL(glob.glob(os.path.join(d, "libswscale-*"))[0])
    sws.swscale_version.restype = ctypes.c_uint
    av.av_get_pix_fmt.argtypes = [ctypes.c_char_p]
    sws.sws_getContext.restype = ctypes.c_void_p
    sws.sws_getContext.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p] * 3
    sws.sws_scale.argtypes = [ctypes.c_void_p, ctypes.c_void_p * 4, ctypes.c_int * 4, ctypes.c_int, ctypes.c_int, ctypes.c_void_p * 4, ctypes.c_int * 4]
    sws.sws_freeContext.argtypes = [ctypes.c_void_p]
    return av, sws


def swscale(av, sws, img, fmt):
    src_name, dst_name = FORMATS[fmt]
    h, w, _ = img.shape
    ctx = sws.sws_getContext(w, h, av.av_get_pix_fmt(src_name.encode()), w, h, av.av_get_pix_fmt(dst_name.encode()), SWS_FAST_BILINEAR, None, None, None)
    assert ctx
    img = np.ascontiguousarray(img)
    cw = (w + 1) // 2 if fmt == "422" else w
    dt = np.uint16 if fmt == "444p16" else np.uint8
    planes = [np.zeros((h, w), dt), np.zeros((h, cw), dt), np.zeros((h, cw), dt)] + ([np.zeros((h, w), dt)] if fmt == "444alpha" else [])
    P4, I4 = ctypes.c_void_p * 4, ctypes.c_int * 4
    dp = [p.ctypes.data for p in planes] + [None] * (4 - len(planes))
    ds = [p.strides[0] for p in planes] + [0] * (4 - len(planes))
    assert sws.sws_scale(ctx, P4(img.ctypes.data, None, None, None), I4(img.strides[0], 0, 0, 0), 0, h, P4(*dp), I4(*ds)) == h
    sws.sws_freeContext(ctx)
    return planes


def image(fmt, w, h, seed):
    rng = np.random.default_rng(seed)
    if fmt == "444p16":
        return rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)
    return rng.integers(0, 256, (h, w, 4 if fmt == "444alpha" else 3), dtype=np.uint8)


def main():
    av, sws = load()
    v = sws.swscale_version()
    version = f"{v >> 16}.{(v >> 8) & 255}.{v & 255}"
    small = {}
    for fmt in FORMATS:
        for i, (w, h) in enumerate(SMALL):
            img = image(fmt, w, h, 100 + i)
            # extremes in the first row: black, white, saturated primaries
            img[0, :6, :3] = np.array([[0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0]]) * (65535 if fmt == "444p16" else 255)
            small[f"{fmt}/{w}x{h}/in"] = img
            for k, p in enumerate(swscale(av, sws, img, fmt)):
                small[f"{fmt}/{w}x{h}/plane{k}"] = p
    np.savez_compressed(os.path.join(HERE, "y4m_small.npz"), **small)
    large = {"libswscale": version, "seed_base": 200, "cases": {}}
    for fmt in FORMATS:
        for i, (w, h) in enumerate(LARGE):
            planes = swscale(av, sws, image(fmt, w, h, 200 + i), fmt)
            large["cases"][f"{fmt}/{w}x{h}"] = {"seed": 200 + i, "sha256": [hashlib.sha256(p.tobytes()).hexdigest() for p in planes]}
    with open(os.path.join(HERE, "y4m_large.json"), "w") as f:
        json.dump(large, f, indent=1)
    print("libswscale", version, len(small), "arrays,", len(large["cases"]), "hashed cases")


if __name__ == "__main__":
    main()
