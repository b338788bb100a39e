#!/usr/bin/env python
"""Throughput of the render path on the reference's own timelines (data/*.otio, shipped 1280x720
u8 media) and on BASELINE.json configs C2-C4 (the same timeline structures with synthetic media at
the named resolution / storage type), next to the CPU oracle on the box's host cores.

  python tools/config_bench.py [--out profiles/xxx.json]

Per timeline: GPU ms/frame with the decoded media resident in HBM (graph build + kernels, CUDA events
around the whole pass over the timeline), end-to-end frames/s through toucan::FrameScheduler (host
media uploaded every frame, frames downloaded), and the oracle's ms/frame (a sample of frames).
B_alg per frame = leaf source frames read once + final frame written once.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import graph as G  # noqa: E402
from tests import media, timelines as TL  # noqa: E402
from toucan_b200 import capi, render  # noqa: E402


def synth(w, h, dtype, seed, disc=False):
    rng = np.random.default_rng(seed)
    a = rng.random((h, w, 4), dtype=np.float32)
    if disc:
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        a[..., 3] = np.clip(1.3 - 2.2 * np.hypot((xx - w / 2) / w, (yy - h / 2) / h), 0.0, 1.0)
    else:
        a[..., 3] = 1.0
    return np.ascontiguousarray(a.astype(dtype))


def resized(doc, size):
    if isinstance(doc, dict):
        return {k: (list(size) if k == "size" and isinstance(v, list) and len(v) == 2 and v == [1280, 720] else resized(v, size))
                for k, v in doc.items()}
    if isinstance(doc, list):
        return [resized(v, size) for v in doc]
    return doc


ONLY = ""


def bench_timeline(host, label, name, frames_by_file, size=None, cpu_frames=2, doc=None):
    if ONLY and not any(o in f"{label} {name}" for o in ONLY.split(",")):
        return None
    # every row starts from an empty stream-ordered pool, like a fresh toucan-render process: blocks cached for the
    # previous timeline's frame sizes otherwise make the allocator split and re-map on every frame
    torch.cuda.synchronize()
    capi.check(capi.lib().tc_pool_trim(), "tc_pool_trim")
    torch.cuda.empty_cache()
    if doc is None:
        doc = json.load(open(os.path.join(TL.DATA, name)))
    if size:
        doc = resized(doc, size)
    by_path = {(f if os.path.isabs(f) else os.path.join(TL.DATA, f)): a for f, a in frames_by_file.items()}
    # --- GPU, media resident ---
    tl = render.Timeline(json_doc=doc, directory=TL.DATA)
    dev = {}
    for p, a in by_path.items():
        t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).cuda()
        if a.dtype == np.uint16:
            t = t.view(torch.uint16)
        dev[p] = t
        tl.set_media(p, t)
    tl.prepare()
    info = tl.info()
    n = info["frames"]
    times = [info["start"] + i for i in range(n)]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    render.bind_stream(0, stream.cuda_stream)
    reps = max(1, 60 // n)
    for t in times:
        tl.render_resident(host, t)
    torch.cuda.synchronize()
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0 = capi.alloc_stats()
    h0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        for t in times:
            tl.render_resident(host, t)
    e1.record()
    host_ms = (time.perf_counter() - h0) * 1e3 / (reps * n)  # time the host needed to ENQUEUE a frame
    a1 = capi.alloc_stats()
    alloc_ms = (a1[1] - a0[1]) * 1e3 / (reps * n)
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / (reps * n)
    launches = (capi.launch_count() - l0) / (reps * n)
    tl.close()
    # --- end to end: host media, scheduler ---
    tl2 = render.Timeline(json_doc=doc, directory=TL.DATA)
    pinned = {}
    for p, a in by_path.items():
        t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).pin_memory()
        pinned[p] = t
        tl2.set_media(p, t.view(torch.uint16) if a.dtype == np.uint16 else t)
    tl2.prepare()
    sink = [0]

    def writer(f, px):
        sink[0] += int(px.size > 0)

    tl2.render_range(0, min(n, 2), writer)
    t0 = time.perf_counter()
    for _ in range(reps):
        tl2.render_range(0, n, writer)
    e2e_fps = reps * n / (time.perf_counter() - t0)
    tl2.close()
    # --- CPU oracle ---
    otl = G.Timeline(doc)
    otl.dir = TL.DATA
    og = G.ImageGraph(otl, lambda p: by_path.get(p))
    fr = otl.frames()
    pick = [fr[i] for i in sorted(set(np.linspace(0, len(fr) - 1, cpu_frames).astype(int).tolist()))]
    og.render(pick[0])
    t0 = time.perf_counter()
    out = None
    for t in pick:
        out = og.render(t)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / len(pick)
    first = next(iter(by_path.values()))
    fbytes = first.nbytes
    obytes = out.nbytes if out is not None else 0
    row = {"config": label, "timeline": name, "frames": n, "size": f"{info['width']}x{info['height']}", "type": str(first.dtype),
           "gpu_ms_per_frame": round(gpu_ms, 4), "host_enqueue_ms_per_frame": round(host_ms, 4), "allocator_ms_per_frame": round(alloc_ms, 4), "gpu_frames_per_s": round(1e3 / gpu_ms, 1), "kernel_launches_per_frame": round(launches, 1),
           "alg_bytes_per_frame": fbytes + obytes, "alg_GBs": round((fbytes + obytes) / (gpu_ms * 1e-3) / 1e9, 1),
           "e2e_frames_per_s": round(e2e_fps, 1), "cpu_oracle_ms_per_frame": round(cpu_ms, 2), "cpu_cores": os.cpu_count(),
           "gpu_over_cpu": round(cpu_ms / gpu_ms, 1)}
    print(json.dumps(row), flush=True)
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", help="only rows whose 'config timeline' label contains this substring")
    args = ap.parse_args()
    global ONLY
    ONLY = args.only
    capi.check(capi.lib().tc_set_device(0), "tc_set_device")
    render.bind_device(0)
    host = render.Host()
    rows = []
    shipped = {f: media.decode(os.path.join(TL.DATA, f)) for f in sorted(os.listdir(TL.DATA)) if f.lower().endswith((".png", ".jpg"))}
    for name in ["Transition.otio", "Transition2.otio", "TransitionWipe.otio", "CompositeTracks.otio", "Generator.otio", "Filter.otio",
                 "MultipleEffects.otio", "Transform.otio", "ColorSpace.otio", "LinearTimeWarp.otio"]:
        rows.append(bench_timeline(host, "C1/fixtures 1280x720 u8", name, shipped))
    charlie = {"Charlie.jpg": synth(1920, 1080, np.float32, 1001)}
    rows.append(bench_timeline(host, "C2 1920x1080 f32", "Filter.otio", charlie))
    from toucan_b200 import workloads  # noqa: E402

    rows.append(bench_timeline(host, "C2 1920x1080 f32", "six filters chained on one clip (240 frames)",
                               {os.path.join(TL.DATA, "source.exr"): charlie["Charlie.jpg"]},
                               doc=workloads.filter_chain_timeline(frames=240, url="source.exr")))
    c3 = {"Charlie.jpg": synth(3840, 2160, np.float16, 1002)}
    rows.append(bench_timeline(host, "C3 3840x2160 half", "MultipleEffects.otio", c3))
    seq = {f"Counter.{i}.png": synth(3840, 2160, np.float16, 1100 + i, disc=True) for i in range(10)}
    rows.append(bench_timeline(host, "C3 3840x2160 half", "CompositeTracks.otio", seq, size=(3840, 2160)))
    c4 = {"Charlie.jpg": synth(3840, 2160, np.float32, 1003)}
    rows.append(bench_timeline(host, "C4 3840x2160 f32", "Transform.otio", c4))
    rows.append(bench_timeline(host, "C4 3840x2160 f32", "ColorSpace.otio", c4))
    rows = [r for r in rows if r]
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
