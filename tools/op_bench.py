#!/usr/bin/env python
"""Per-kernel HBM throughput of the C-ABI ops (SURVEY 8d per-op figures): algorithmic bytes
(inputs once + output once) / CUDA-event time, against MEASURED_PEAKS.json.

  python tools/op_bench.py [--size 3840x2160] [--types f32,f16,u8] [--out profiles/xxx.json]

Inputs rotate over enough buffers to exceed the 126 MB L2 between iterations.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from toucan_b200 import capi, ops  # noqa: E402

TYPES = {"u8": torch.uint8, "u16": torch.uint16, "f16": torch.float16, "f32": torch.float32}


def make(w, h, dt, n):
    out = []
    for i in range(n):
        f = torch.rand((h, w, 4), device="cuda", dtype=torch.float32)
        if dt == torch.uint8:
            out.append((f * 255).to(torch.uint8))
        elif dt == torch.uint16:
            out.append((f * 65535).to(torch.int32).to(torch.uint16))
        else:
            out.append(f.to(dt))
    return out


def timeit(fn, iters=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="3840x2160")
    ap.add_argument("--types", default="f32,f16,u8")
    ap.add_argument("--out", default="")
    ap.add_argument("--ops", default="", help="only ops whose name contains this substring")
    args = ap.parse_args()
    w, h = [int(v) for v in args.size.split("x")]
    capi.check(capi.lib().tc_set_device(0), "tc_set_device")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rows = []
    for tname in args.types.split(","):
        dt = TYPES[tname]
        px = w * h * 4 * torch.empty((), dtype=dt).element_size()
        n = max(3, int(300e6 // px) + 2)
        a = make(w, h, dt, n)
        b = make(w, h, dt, n)
        outs = [torch.empty_like(a[0]) for _ in range(n)]
        tcode = ops.tcode(a[0])

        def rec(name, frames, ms, out_bytes=None):
            byts = frames * px if out_bytes is None else out_bytes
            gbs = byts / (ms * 1e-3) / 1e9
            rows.append({"op": name, "type": tname, "size": args.size, "ms": round(ms, 4), "bytes": byts, "GB/s": round(gbs, 1),
                         "frac_of_measured_peak": round(gbs / peak, 3)})
            print(f"{name:28s} {tname:4s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:6.1%}", flush=True)

        def run(name, frames, fn, iters=20, out_bytes=None):
            if args.ops and args.ops not in name:
                return
            rec(name, frames, timeit(fn, iters=iters), out_bytes)

        lib = capi.lib()
        B = ops._B
        S = ops._stream

        def unary(fn, *extra):
            return lambda i: capi.check(fn(B(ops.img(outs[i % n])), B(ops.img(a[i % n])), *extra, S()), "op")

        run("fill", 1, lambda i: capi.check(lib.tc_fill(B(ops.img(outs[i % n])), capi.f4([0.1, 0.2, 0.3, 1.0]), S()), "fill"))
        run("checkers", 1, lambda i: capi.check(lib.tc_checkers(B(ops.img(outs[i % n])), 100, 100, capi.f4([0, 0, 0, 1]), capi.f4([1, 1, 1, 1]), S()), "ck"))
        run("noise gaussian", 1, lambda i: capi.check(lib.tc_noise(B(ops.img(outs[i % n])), b"gaussian", 0.5, 0.1, 0, 1, S()), "noise"))
        run("invert", 2, unary(lib.tc_invert))
        run("flip", 2, unary(lib.tc_flip))
        run("flop", 2, unary(lib.tc_flop))
        run("pow 2.2", 2, unary(lib.tc_pow, 2.2))
        run("saturate", 2, unary(lib.tc_saturate, 0.5))
        run("color_map plasma", 2, unary(lib.tc_color_map, b"plasma"))
        run("premult", 2, unary(lib.tc_premult))
        run("colorconvert lin709->ACEScct", 2, unary(lib.tc_color_convert, b"Linear Rec.709 (sRGB)", b"ACEScct"))
        run("blur r10", 2, unary(lib.tc_blur, 10.0))
        run("blur r20", 2, unary(lib.tc_blur, 20.0))
        run("unsharp r10", 2, unary(lib.tc_unsharp_mask, b"gaussian", 10.0, 1.0, 0.0))
        run("rotate 45", 2, unary(lib.tc_rotate, 0.7853982, b"", 0.0), iters=5)
        half = [torch.empty((h // 2, w // 2, 4), device="cuda", dtype=dt) for _ in range(n)]
        run("resize 1/2 (lanczos3)", 1.25, lambda i: capi.check(lib.tc_resize(B(ops.img(half[i % n])), B(ops.img(a[i % n])), b"", 0.0, S()), "rs"), iters=5)
        run("resize x2 (blackman-harris)", 1.25, lambda i: capi.check(lib.tc_resize(B(ops.img(outs[i % n])), B(ops.img(half[i % n])), b"", 0.0, S()), "rs"), iters=5)
        run("dissolve", 3, lambda i: capi.check(lib.tc_dissolve(B(ops.img(outs[i % n])), B(ops.img(a[i % n])), B(ops.img(b[i % n])), 0.3, S()), "ds"))
        run("wipe h", 3, lambda i: capi.check(lib.tc_wipe(B(ops.img(outs[i % n])), B(ops.img(a[i % n])), B(ops.img(b[i % n])), 0.3, 0, S()), "wp"))
        run("premult+over", 3, lambda i: capi.check(lib.tc_over(B(ops.img(outs[i % n])), B(ops.img(a[i % n])), B(ops.img(b[i % n])), 1, S()), "ov"))
        for fmt in ("422", "444", "444alpha", "444p16"):
            ybytes = int(lib.tc_y4m_frame_bytes(capi.Y4M_FORMATS[fmt], w, h))
            ybuf = [torch.empty((ybytes,), dtype=torch.uint8, device="cuda") for _ in range(3)]
            run(f"y4m {fmt}", 1, lambda i, fmt=fmt, ybuf=ybuf: capi.check(lib.tc_rgb_to_y4m(
                ctypes.c_void_p(ybuf[i % 3].data_ptr()), B(ops.img(a[i % n])), capi.Y4M_FORMATS[fmt], S()), "y4m"), out_bytes=px + ybytes)
        del a, b, outs, half
        torch.cuda.empty_cache()
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump({"peak_gbs": peak, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
