"""OpenEXR fixtures for the built-in EXR reader, written by the REAL OpenEXR library (through OpenCV).

Run in the authoring container:
    python tests/golden/make_exr_golden.py

Writes tests/golden/exr/*.exr (seeded 37x37 frames: half / float samples, RGB / RGBA, compression NONE / RLE / ZIPS / ZIP / PIZ / PXR24 / B44 / B44A)
and tests/golden/exr/expected.npz with the pixels OpenCV's OpenEXR reader returns for each file (RGBA order, float32).
"""
import os

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"

import cv2  # noqa: E402
import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "exr")
TYPES = {"half": cv2.IMWRITE_EXR_TYPE_HALF, "float": cv2.IMWRITE_EXR_TYPE_FLOAT}
COMPRESSIONS = {"none": cv2.IMWRITE_EXR_COMPRESSION_NO, "rle": cv2.IMWRITE_EXR_COMPRESSION_RLE, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS,
                "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP, "piz": cv2.IMWRITE_EXR_COMPRESSION_PIZ}


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(77)
    w, h = 37, 37  # more than one 16-line ZIP chunk and more than one 32-line PIZ chunk, odd sizes
    expected = {}
    for tname, t in TYPES.items():
        for cname, c in COMPRESSIONS.items():
            for nch in (3, 4):
                a = (rng.random((h, w, nch), dtype=np.float32) * 3.0 - 1.0).astype(np.float32)  # negative and > 1 values
                a[:4, :6] = 0.25  # a flat patch: runs for the RLE coder
                name = f"{tname}_{cname}_{'rgba' if nch == 4 else 'rgb'}.exr"
                bgr = a[..., [2, 1, 0, 3]] if nch == 4 else a[..., ::-1]
                assert cv2.imwrite(os.path.join(OUT, name), np.ascontiguousarray(bgr), [cv2.IMWRITE_EXR_TYPE, t, cv2.IMWRITE_EXR_COMPRESSION, c])
                b = cv2.imread(os.path.join(OUT, name), cv2.IMREAD_UNCHANGED)
                rgba = b[..., [2, 1, 0, 3]] if nch == 4 else np.concatenate([b[..., ::-1], np.ones((h, w, 1), np.float32)], -1)
                expected[name] = np.ascontiguousarray(rgba.astype(np.float32))
    # PXR24 (zlib over byte planes of per-channel differences; float samples keep 24 bits, half samples are exact): its own
    # generator so the files above stay what they were
    rng24 = np.random.default_rng(78)
    for tname, t in TYPES.items():
        for nch in (3, 4):
            a = (rng24.random((h, w, nch), dtype=np.float32) * 3.0 - 1.0).astype(np.float32)
            a[:4, :6] = 0.25
            name = f"{tname}_pxr24_{'rgba' if nch == 4 else 'rgb'}.exr"
            bgr = a[..., [2, 1, 0, 3]] if nch == 4 else a[..., ::-1]
            assert cv2.imwrite(os.path.join(OUT, name), np.ascontiguousarray(bgr), [cv2.IMWRITE_EXR_TYPE, t, cv2.IMWRITE_EXR_COMPRESSION,
                                                                                     cv2.IMWRITE_EXR_COMPRESSION_PXR24])
            b = cv2.imread(os.path.join(OUT, name), cv2.IMREAD_UNCHANGED)
            rgba = b[..., [2, 1, 0, 3]] if nch == 4 else np.concatenate([b[..., ::-1], np.ones((h, w, 1), np.float32)], -1)
            expected[name] = np.ascontiguousarray(rgba.astype(np.float32))
    # B44 / B44A: lossy at the writer (half samples in 4 x 4 blocks of 14 or 3 bytes), exactly specified at the reader --
    # expected = what the OpenEXR library decodes.  Smooth content plus a flat patch (B44A's 3-byte blocks) plus noise.
    rng44 = np.random.default_rng(79)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for cname, c in (("b44", cv2.IMWRITE_EXR_COMPRESSION_B44), ("b44a", cv2.IMWRITE_EXR_COMPRESSION_B44A)):
        for tname, t in TYPES.items():
            for nch in (3, 4):
                a = np.stack([xx / w * 2.0 - 0.5, yy / h, (xx * yy) / (w * h), np.clip(1.2 - xx / w, 0, 1)][:nch], -1).astype(np.float32)
                a += (rng44.random((h, w, nch), dtype=np.float32) - 0.5) * 0.05
                a[5:20, 8:30] = 0.25
                name = f"{tname}_{cname}_{'rgba' if nch == 4 else 'rgb'}.exr"
                bgr = a[..., [2, 1, 0, 3]] if nch == 4 else a[..., ::-1]
                assert cv2.imwrite(os.path.join(OUT, name), np.ascontiguousarray(bgr), [cv2.IMWRITE_EXR_TYPE, t, cv2.IMWRITE_EXR_COMPRESSION, c])
                b = cv2.imread(os.path.join(OUT, name), cv2.IMREAD_UNCHANGED)
                rgba = b[..., [2, 1, 0, 3]] if nch == 4 else np.concatenate([b[..., ::-1], np.ones((h, w, 1), np.float32)], -1)
                expected[name] = np.ascontiguousarray(rgba.astype(np.float32))
    # a lossy DWAA file: a compression the reader must refuse
    a = rng.random((h, w, 3), dtype=np.float32)
    assert cv2.imwrite(os.path.join(OUT, "refused_dwaa_rgb.exr"), a, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION,
                                                                      cv2.IMWRITE_EXR_COMPRESSION_DWAA])
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
    print(len(expected), "files,", sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")


if __name__ == "__main__":
    main()
