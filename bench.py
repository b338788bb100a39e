#!/usr/bin/env python
"""bench.py -- frames/s and HBM GB/s of the toucan render hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|s4k|c2|c3|c4]

Workloads (BASELINE.json `configs`; a "step" is one timeline frame):
  c5 (default, configs[4], the configuration the 1/2/4/8-GPU metric is quoted on): synthetic 1000-frame, 10-track
      7680x4320 RGBA float32 timeline, toucan:Blur radius 10 on every clip, toucan:Dissolve between clips, tracks
      composited premult+over on black.  Timed steps alternate plain (10 sources) and mid-transition (20 sources) frames.
  s4k the c5 timeline at 3840x2160: north_star's "synthetic 4K RGBA-float multi-effect timeline" (target >= 60 % of the HBM roofline)
  c2  configs[1]: the six filters of data/Filter.otio chained on one clip, 1920x1080 RGBA float32
  c3  configs[2]: data/MultipleEffects.otio's structure at 3840x2160 RGBA half
  c4  configs[3]: data/Transform.otio's structure at 3840x2160 RGBA float32

  value     frames/s with the decoded source stills resident in HBM (whole job, all ranks), CUDA events
  sustained the same over a >= 2 s loop (the burst above is tens of milliseconds at the top clock)
  e2e       frames/s through toucan::FrameScheduler from HOST buffers: the source stills start in pinned host memory, the
            device source cache starts empty, every still a frame reads is uploaded inside the timed region (once per still
            and GPU: uploads are de-duplicated by media path), every finished frame is downloaded to the ordered writer
  roofline  algorithmic bytes (sources read once + frame written once) / time, against the measured HBM copy peak
  cpu_baseline / --impl reference
            the CPU restatement of the reference's OIIO path (oracle/, "port": the real toucan-render cannot be built here)
            with an explicitly set OpenMP team of all host cores.  --impl reference renders FULL frames, nothing extrapolated
            (when K + W <= 32; else a fixed 1/16 band per step); the GPU arm's cpu_baseline times a fixed 1/16 band per frame
            plus one full frame to validate the x16

Frames are sharded in contiguous blocks across ranks (no collective on the data path; torch.distributed is used for
the barrier and the max-over-ranks time only).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from toucan_b200 import workloads  # noqa: E402

BAND_DIV = 16  # a CPU band step renders rows 0 .. H/16-1 of the frame; fixed, never adapted
FULL_FRAME_STEPS = 32  # --impl reference renders full frames when steps + warmup <= this


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# clocks sampler
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v == "Active":
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# c5 on the CPU: the oracle, the unfused node chain the reference executes per frame (bench.py is one of the places
# allowed to run oracle/)
def c5_sources_numpy(cfg):
    import numpy as np

    w, h = cfg["width"], cfg["height"]
    src = {}
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for t in range(cfg["tracks"]):
        for k in range(2):
            rng = np.random.default_rng(2000 + 10 * t + k)
            a = rng.random((h, w, 4), dtype=np.float32)
            r = np.hypot((xx - w * (0.3 + 0.04 * t)) / w, (yy - h * 0.5) / h)
            a[..., 3] = np.clip(1.2 - 2.0 * r, 0.0, 1.0)
            src[(t, k)] = a
    return src


def c5_cpu_frame(O, src, layers, rows=None):
    """ImageGraph.cpp:165-198 for this timeline: Fill, then per track Blur [Dissolve Blur] premult over.  rows = the
    number of top rows to render (a band: contiguous views of the same stills), None = the full frame."""
    first = next(iter(src.values()))
    h = first.shape[0] if rows is None else rows
    w = first.shape[1]
    acc = O.fill(w, h, [0.0, 0.0, 0.0, 1.0], O.F32)
    for a, b, ra, rb, v in layers:
        x = O.blur(src[a][:h], ra)
        if b is not None:
            x = O.dissolve(x, O.blur(src[b][:h], rb), v)
        acc = O.comp(x, acc, True)
    return acc


def cpu_team(O):
    """Set the OpenMP team explicitly: torchrun exports OMP_NUM_THREADS=1 to its ranks.  Returns the threads in use."""
    O.lib().o_set_threads(host_threads())
    return int(O.lib().o_get_threads())


def c5_cpu_arm(cfg, src, frames, warmup, full):
    """Times the oracle over frames[warmup:].  full: whole frames; else the fixed 1/16 band, scaled x16, plus ONE
    full frame whose measured time is reported next to 16 x the band time of the same frame."""
    from oracle import oracle as O

    O.lib()
    threads = cpu_team(O)
    rows = None if full else cfg["height"] // BAND_DIV
    for f in frames[:warmup]:
        c5_cpu_frame(O, src, workloads.stack_layers(f, **cfg), rows)
    t0 = time.perf_counter()
    for f in frames[warmup:]:
        c5_cpu_frame(O, src, workloads.stack_layers(f, **cfg), rows)
    dt = time.perf_counter() - t0
    n = len(frames) - warmup
    info = {"threads": threads, "seconds": dt, "steps": n}
    if full:
        info["fps"] = n / dt
        info["sample"] = f"{n} full {cfg['width']}x{cfg['height']} frames (nothing extrapolated)"
    else:
        info["fps"] = n / (dt * BAND_DIV)
        info["sample"] = (f"{n} frames x rows 0..{rows - 1} of {cfg['height']} (a fixed 1/{BAND_DIV} band of each "
                          f"{cfg['width']}x{cfg['height']} frame), time scaled x{BAND_DIV}")
        f = frames[-1]
        layers = workloads.stack_layers(f, **cfg)
        t0 = time.perf_counter()
        c5_cpu_frame(O, src, layers, rows)
        band = time.perf_counter() - t0
        t0 = time.perf_counter()
        c5_cpu_frame(O, src, layers, None)
        whole = time.perf_counter() - t0
        info["full_frame_check"] = {"frame": f, "full_frame_s": whole, "band_s_x16": band * BAND_DIV,
                                    "measured_over_extrapolated": whole / (band * BAND_DIV)}
    return info


def c5_frames(rank, world, cfg, count):
    """Frames a rank renders, inside its own contiguous block: plain frames (every track shows one clip: 10 sources)
    and mid-transition frames (20 sources) ALTERNATE, so any even number of timed steps is an exact 50:50 mix --
    the share the 1000-frame timeline has (offset 12 either side of a cut every 48 frames)."""
    lo, hi = workloads.frame_block(rank, world, cfg["frames"])
    plain, trans = [], []
    for f in range(lo, hi):
        (trans if workloads.stack_layers(f, **cfg)[0][1] is not None else plain).append(f)
    if not plain or not trans:
        return [lo + s % max(1, hi - lo) for s in range(count)]
    return [(plain[(s // 2) % len(plain)] if s % 2 == 0 else trans[(s // 2) % len(trans)]) for s in range(count)]


def c5_e2e_window(rank, world, cfg, steps):
    """`steps` CONSECUTIVE frames for the scheduler, half before and half inside a transition of the rank's block."""
    lo, hi = workloads.frame_block(rank, world, cfg["frames"])
    period, off = cfg["clip_frames"], cfg["offset"]
    first_cut = ((lo + steps) // period + 1) * period
    start = first_cut - off - steps // 2
    if start < lo or start + steps > hi:
        start = lo
    return start


def c5_config_json(cfg, world):
    if (cfg["width"], cfg["height"]) == (3840, 2160):
        label = ("s4k: the c5 timeline at 3840x2160 -- north_star's \"synthetic 4K RGBA-float multi-effect timeline\": 1000 frames, 10 tracks, "
                 "blur(radius 10)+dissolve per track, premult+over stack, RGBA float32")
    else:
        label = ("c5: synthetic 1000-frame 10-track 7680x4320 RGBA float32 timeline, blur(radius 10)+dissolve per track, "
                 "premult+over stack (BASELINE.json configs[4])")
    mb = cfg["width"] * cfg["height"] * 16 / 1e6
    return {"workload": label,
            "width": cfg["width"], "height": cfg["height"], "tracks": cfg["tracks"], "frames": cfg["frames"],
            "frame_mix": "timed steps alternate plain (10 sources) and mid-transition (20 sources) frames: 50:50",
            "sharding": f"contiguous frame blocks over {world} GPU(s), no collective",
            "l2": f"inputs larger than L2 (each source frame {mb:.0f} MB, 10-20 per step)"}


# ---------------------------------------------------------------------------------------------
# c2 - c4: the reference's own timeline structures at the named resolution / storage type
def small_workload(name):
    """-> dict(doc, media {relative url: ndarray}, label, dtype, size, directory)."""
    import numpy as np

    data = os.path.join(ROOT, "tests", "golden", "data")

    def synth(w, h, dtype, seed):
        rng = np.random.default_rng(seed)
        a = rng.random((h, w, 4), dtype=np.float32)
        a[..., 3] = 1.0
        return np.ascontiguousarray(a.astype(dtype))

    if name == "c2":
        return dict(doc=workloads.filter_chain_timeline(frames=240, url="source.exr"), media={"source.exr": synth(1920, 1080, np.float32, 1001)},
                    label="c2: ColorMap > Invert > Pow > Saturate > Blur(10) > UnsharpMask(10) chained on one clip, 1920x1080 RGBA float32 "
                          "(BASELINE.json configs[1], chained variant of data/Filter.otio)", dtype="f32", size=(1920, 1080), directory=data)
    if name == "c3":
        doc = json.load(open(os.path.join(data, "MultipleEffects.otio")))
        return dict(doc=doc, media={"Charlie.jpg": synth(3840, 2160, np.float16, 1002)},
                    label="c3: data/MultipleEffects.otio (Flip+Blur 20 / Flop+ColorMap clips, track Invert, stack Pow 3) at 3840x2160 RGBA half "
                          "(BASELINE.json configs[2])", dtype="f16", size=(3840, 2160), directory=data)
    if name == "c4":
        doc = json.load(open(os.path.join(data, "Transform.otio")))
        return dict(doc=doc, media={"Charlie.jpg": synth(3840, 2160, np.float32, 1003)},
                    label="c4: data/Transform.otio (Crop, Flip, Flop, Resize, Rotate 45) at 3840x2160 RGBA float32 (BASELINE.json configs[3])",
                    dtype="f32", size=(3840, 2160), directory=data)
    raise SystemExit("unknown workload " + name)


def small_config_json(wl, world):
    return {"workload": wl["label"], "width": wl["size"][0], "height": wl["size"][1],
            "sharding": f"contiguous frame blocks over {world} GPU(s), no collective",
            "l2": "source rotated over copies larger than L2 between timed frames" if wl.get("rotated") else "per-frame traffic is below L2 size; "
                  "sources are rotated over 3 copies so consecutive frames read different memory"}


def small_cpu_arm(wl, steps, warmup):
    import numpy as np

    from oracle import graph as G
    from oracle import oracle as O

    O.lib()
    threads = cpu_team(O)
    otl = G.Timeline(wl["doc"])
    otl.dir = wl["directory"]
    by_path = {os.path.join(wl["directory"], k): v for k, v in wl["media"].items()}
    og = G.ImageGraph(otl, lambda p: by_path.get(p))
    fr = otl.frames()
    seq = [fr[i % len(fr)] for i in range(steps + warmup)]
    for t in seq[:warmup]:
        og.render(t)
    t0 = time.perf_counter()
    for t in seq[warmup:]:
        og.render(t)
    dt = time.perf_counter() - t0
    del np
    return {"threads": threads, "seconds": dt, "steps": steps, "fps": steps / dt,
            "sample": f"{steps} full {wl['size'][0]}x{wl['size'][1]} frames of the timeline (nothing extrapolated)"}


# ---------------------------------------------------------------------------------------------
def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    if args.workload in ("c5", "s4k"):
        full = args.steps + args.warmup <= FULL_FRAME_STEPS
        frames = c5_frames(0, world, cfg, args.steps + args.warmup)
        info = c5_cpu_arm(cfg, c5_sources_numpy(cfg), frames, args.warmup, full)
        config, dtype = c5_config_json(cfg, world), "f32"
    else:
        wl = small_workload(args.workload)
        info = small_cpu_arm(wl, args.steps, args.warmup)
        config, dtype = small_config_json(wl, world), wl["dtype"]
    fps = info["fps"]
    line = {"impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / fps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": info["threads"], "kind": "port", "sample": info["sample"],
                             "seconds": info["seconds"]},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the toucan/OIIO path (oracle/) on rank 0's host cores, OpenMP team set explicitly; the real "
                    "toucan-render needs OIIO/OCIO/OTIO/OpenFX which are not installed and cannot be fetched"}
    if "full_frame_check" in info:
        line["cpu_baseline"]["full_frame_check"] = info["full_frame_check"]
    if args.scale != 1.0:
        line["INVALID"] = "debug run at reduced frame size (--scale)"
    print(json.dumps(line), flush=True)


def bind_to_gpu_cpus(torch, local):
    """Run this rank (and the threads it starts) on the CPUs NVML reports as local to its GPU, so pinned
    staging memory is first-touched on the GPU's NUMA node.  Best effort: silently skipped if NVML is absent."""
    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


class Dist:
    """One process per GPU: device, NCCL group for the barrier and the max-over-ranks time."""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device visible; the render path has no CPU fallback")
        torch.cuda.set_device(self.local)
        from toucan_b200 import capi

        capi.check(capi.lib().tc_set_device(self.local), "tc_set_device")
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            bind_to_gpu_cpus(torch, self.local)
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def timed_loop(D, step, items, warm):
    """W untimed steps, barrier + sync, K timed steps between CUDA events on the render stream, barrier + sync."""
    torch = D.torch
    from toucan_b200 import capi

    for it in warm:
        step(it)
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = capi.launch_count()
    e0.record()
    last = None
    for it in items:
        last = step(it)
    e1.record()
    D.barrier()
    D.launches = capi.launch_count() - l0  # kernels of libtoucan_cuda.so launched inside the timed region
    return D.max(e0.elapsed_time(e1)), last


def e2e_run(D, tl, first, steps, warm_first, y4m_format=None):
    """toucan::FrameScheduler through the C API from host media with a COLD device source cache: every still is uploaded
    inside the timed region the first time a frame reads it.  Wall clock around the whole call, max over ranks."""
    sink = {"n": 0, "acc": 0.0, "t": []}

    def writer(frame, pixels):
        sink["n"] += 1
        sink["t"].append(time.perf_counter())
        if pixels.size:
            sink["acc"] += float(pixels.reshape(-1)[0])  # the writer's read of the finished frame

    tl.render_range(warm_first, warm_first + 2, writer, devices=[D.local], y4m_format=y4m_format)  # pool growth, pinned ring, plug-in load
    tl.drop_source_cache()  # ... and forget the uploads: the timed region starts with nothing resident
    D.torch.cuda.synchronize()
    n0, b0 = tl.upload_stats()
    D.barrier()
    t0 = time.perf_counter()
    tl.render_range(first, first + steps, writer, devices=[D.local], y4m_format=y4m_format)
    dt = time.perf_counter() - t0
    D.barrier()
    n1, b1 = tl.upload_stats()
    assert sink["n"] == steps + 2
    if os.environ.get("TOUCAN_BENCH_TRACE"):
        sys.stderr.write("e2e frame arrival ms: " + " ".join(f"{(t - t0) * 1e3:.0f}" for t in sink["t"][2:]) + "\n")
    return D.max(dt), dt, b1 - b0, n1 - n0


def run_c5(args, cfg):
    D = Dist()
    torch = D.torch
    from toucan_b200 import capi, render

    rank, world, local, dev = D.rank, D.world, D.local, D.dev
    w, h = cfg["width"], cfg["height"]
    F = w * h * 16
    # synthetic decoded stills
    ys = torch.arange(h, device=dev, dtype=torch.float32).view(h, 1)
    xs = torch.arange(w, device=dev, dtype=torch.float32).view(1, w)
    src = {}
    for t in range(cfg["tracks"]):
        for k in range(2):
            g = torch.Generator(device=dev)
            g.manual_seed(2000 + 10 * t + k)
            a = torch.rand((h, w, 4), generator=g, device=dev, dtype=torch.float32)
            r = torch.hypot((xs - w * (0.3 + 0.04 * t)) / w, (ys - h * 0.5) / h)
            a[..., 3] = torch.clamp(1.2 - 2.0 * r, 0.0, 1.0)
            src[(t, k)] = a
    del ys, xs

    # the reference-facing path: OTIO timeline -> toucan::ImageGraph -> OpenFX host/plug-ins -> kernels
    doc = workloads.stack_timeline(**cfg)
    ofx = render.Host()
    tl = render.Timeline(json_doc=doc, directory="/synthetic")
    for (t, k), a in src.items():
        tl.set_media(tl.media_path(workloads.still_url(t, k)), a)  # resident in HBM
    tl.prepare()
    info = tl.info()
    assert (info["width"], info["height"], info["type"]) == (w, h, capi.F32), info
    stream = torch.cuda.Stream(dev)  # the render stream: CUDA events below are recorded on it
    torch.cuda.set_stream(stream)
    render.bind_stream(local, stream.cuda_stream)

    frames = c5_frames(rank, world, cfg, args.steps + args.warmup)
    nsrc = {f: sum(1 + (la[1] is not None) for la in workloads.stack_layers(f, **cfg)) for f in set(frames)}
    start = info["start"]

    def step(f):
        return tl.render_resident(ofx, start + f)

    clocks = Clocks(local) if rank == 0 else None
    ms, last = timed_loop(D, step, frames[args.warmup:], frames[:args.warmup])
    launches = D.launches
    ptr = last[0]
    fps = world * args.steps / (ms * 1e-3)
    alg_bytes = sum((nsrc[f] + 1) * F for f in frames[args.warmup:])
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    # per frame kind (CUDA events around each kind alone, same stream): where the two halves of the mix stand
    kinds = {}
    for kind, n in (("plain", cfg["tracks"]), ("transition", 2 * cfg["tracks"])):
        fr = [f for f in frames if nsrc[f] == n][:8]
        if fr:
            kms, _ = timed_loop(D, step, fr, fr[:1])
            kinds[kind] = {"ms": kms / len(fr), "algorithmic_bytes": (n + 1) * F,
                           "frac_of_peak": (n + 1) * F / (kms / len(fr) * 1e-3) / 1e9 / peak}
    # sustained: >= 2 s of the same alternating mix (the power cap, not the burst clock, sets this one)
    sus = None
    if os.environ.get("TOUCAN_BENCH_SUSTAINED", "1") != "0":
        n_sus = max(args.steps, int(2000.0 / max(1e-3, ms / args.steps)) // 2 * 2)
        sfr = [frames[args.warmup + (i % max(2, args.steps // 2 * 2))] for i in range(n_sus)]
        sms, _ = timed_loop(D, step, sfr, [])
        sbytes = sum((nsrc[f] + 1) * F for f in sfr)
        sus = {"value": world * n_sus / (sms * 1e-3), "unit": "frames/s", "steps": n_sus, "ms_per_step": sms / n_sus,
               "roofline_frac": sbytes / (sms * 1e-3) / 1e9 / peak}
    clk = clocks.stop() if clocks else None

    out_host = torch.empty((h, w, 4), dtype=torch.float32).pin_memory()
    capi.check(capi.lib().tc_download(ctypes.c_void_p(out_host.data_ptr()), ctypes.c_void_p(ptr), F,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "tc_download")
    torch.cuda.synchronize()
    checksum = float(out_host[::97, ::89].double().sum().item())
    last_frame = frames[-1]

    # ---- end to end: host buffers in, host buffers out, every copy inside the timed region ----
    e2e = None
    e2e_y4m = None
    e2e_reupload = None
    host_src = None
    e2e_steps = int(os.environ.get("TOUCAN_BENCH_E2E_STEPS", str(min(args.steps, 24))))
    if e2e_steps > 0:
        tl.close()
        host_src = {k: v.cpu().pin_memory() for k, v in src.items()}
        src.clear()
        torch.cuda.empty_cache()
        capi.lib().tc_pool_trim()
        tl2 = render.Timeline(json_doc=doc, directory="/synthetic")
        for (t, k), a in host_src.items():
            tl2.set_media(tl2.media_path(workloads.still_url(t, k)), a)  # pinned host memory
        tl2.set_source_cache(64 << 30)  # what `toucan-render` runs with by default (-cache 64)
        tl2.prepare()
        first = c5_e2e_window(rank, world, cfg, e2e_steps)
        dt, my_dt, up_bytes, up_n = e2e_run(D, tl2, first, e2e_steps, first + e2e_steps // 2)
        h2d_total = D.sum(float(up_bytes))
        e2e = {"value": world * e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_total / (world * e2e_steps)),
               "d2h_bytes_per_step": F, "steps": e2e_steps, "source_uploads_per_rank": int(up_n),
               "h2d_GBs_aggregate": h2d_total / dt / 1e9, "d2h_GBs_aggregate": world * e2e_steps * F / dt / 1e9,
               "note": "toucan::FrameScheduler through the C API: source stills in pinned host memory, device source cache EMPTY at the "
                       "start of the timed region, each still uploaded the first time a frame reads it (de-duplicated by media path: "
                       f"{int(up_n)} uploads of {F / 1e6:.0f} MB on this rank), {e2e_steps} consecutive frames (half before, half inside a "
                       "transition) rendered and every finished float32 frame downloaded into a pinned ring for the ordered writer; "
                       "wall clock, max over ranks"}
        if os.environ.get("TOUCAN_BENCH_Y4M", "1") != "0":
            # (informational) the same run with the reference README's pipe-to-ffmpeg output, `toucan-render timeline.otio - -y4m 444`:
            # the RGB -> planar YUV conversion of App.cpp:340-485 runs on the device, so 3 bytes per pixel cross PCIe instead of 16
            dty, _, upy, _ = e2e_run(D, tl2, first, e2e_steps, first + e2e_steps // 2, y4m_format="444")
            e2e_y4m = {"value": world * e2e_steps / dty, "unit": "frames/s", "h2d_bytes_per_step": int(D.sum(float(upy)) / (world * e2e_steps)),
                       "d2h_bytes_per_step": 3 * w * h, "steps": e2e_steps,
                       "note": "informational: as `e2e` (cold source cache, every still uploaded inside the timed region) but the finished frame is "
                               "converted to y4m 4:4:4 on the device and those bytes are downloaded"}
        if world > 1 and os.environ.get("TOUCAN_BENCH_FANOUT", "1") != "0":
            # N GPUs of one box read the SAME stills: instead of every rank pulling all 20 through the one PCIe root they share
            # (N x 10.6 GB), each still crosses PCIe ONCE -- uploaded by the rank that owns it (index % N) -- and is fanned out to the
            # other GPUs over NVLink (NCCL broadcast: the one real exchange step of the sharded job).  All of it inside the timed region.
            tl2.close()
            keys = sorted(host_src)
            dev_buf = {k: torch.zeros((h, w, 4), dtype=torch.float32, device=dev) for k in keys}
            tl3 = render.Timeline(json_doc=doc, directory="/synthetic")
            for k in keys:
                tl3.set_media(tl3.media_path(workloads.still_url(*k)), dev_buf[k])  # device media: the scheduler uploads nothing
            tl3.prepare()
            sink = {"n": 0}

            def writer(frame, pixels):
                sink["n"] += 1

            tl3.render_range(first + e2e_steps // 2, first + e2e_steps // 2 + 2, writer, devices=[local])  # warm-up
            for i, k in enumerate(keys):  # NCCL warm-up (communicator channels) on a small slice
                D.dist.broadcast(dev_buf[k][:8], src=i % world)
            D.barrier()
            t0 = time.perf_counter()
            own = 0
            for i, k in enumerate(keys):  # every rank uploads its own share at the same time ...
                if i % world == rank:
                    dev_buf[k].copy_(host_src[k], non_blocking=True)
                    own += 1
            for i, k in enumerate(keys):  # ... then the shares are exchanged over NVLink
                D.dist.broadcast(dev_buf[k], src=i % world)
            torch.cuda.synchronize()
            t_dist = time.perf_counter() - t0
            tl3.render_range(first, first + e2e_steps, writer, devices=[local])
            dt3 = time.perf_counter() - t0
            D.barrier()
            dt3 = D.max(dt3)
            assert sink["n"] == e2e_steps + 2
            e2e_independent = dict(e2e, note="informational: every rank uploads all the stills it reads itself (no exchange between GPUs); " + e2e["note"])
            h2d3 = D.sum(float(own * F))
            e2e = {"value": world * e2e_steps / dt3, "unit": "frames/s", "h2d_bytes_per_step": int(h2d3 / (world * e2e_steps)),
                   "d2h_bytes_per_step": F, "steps": e2e_steps, "source_uploads_per_rank": own, "source_distribution_s": D.max(t_dist),
                   "h2d_GBs_aggregate": h2d3 / dt3 / 1e9, "d2h_GBs_aggregate": world * e2e_steps * F / dt3 / 1e9,
                   "independent_uploads": e2e_independent,
                   "note": f"source stills in pinned host memory, each uploaded ONCE per job by the rank that owns it ({own} of {len(keys)} on this "
                           "rank) and broadcast to the other GPUs over NVLink (NCCL), then toucan::FrameScheduler renders "
                           f"{e2e_steps} consecutive frames per rank and downloads every finished float32 frame to the ordered writer; upload, "
                           "broadcast, render and download all inside the timed region; wall clock, max over ranks"}
            tl3.close()
            dev_buf.clear()
            tl2 = None
        if tl2 is not None and os.environ.get("TOUCAN_BENCH_REUPLOAD", "1") != "0":
            # (informational) round 1's arrangement: no source cache, every frame re-uploads every still it reads
            tl2.set_source_cache(0)
            n_re = min(e2e_steps, 8)
            dt2, _, up2, _ = e2e_run(D, tl2, first + e2e_steps // 2 - n_re // 2, n_re, first)
            e2e_reupload = {"value": world * n_re / dt2, "unit": "frames/s", "h2d_bytes_per_step": int(D.sum(float(up2)) / (world * n_re)),
                            "d2h_bytes_per_step": F, "steps": n_re,
                            "note": "informational: source cache off, every frame uploads all 10-20 stills it reads (the reference re-reads "
                                    "its media every frame); PCIe-bound by construction"}
        if tl2 is not None:
            tl2.close()

    cpu = None
    if rank == 0 and world == 1 and os.environ.get("TOUCAN_BENCH_CPU", "1") != "0":
        # the oracle on the SAME stills (their pinned host copies), a fixed 1/16 band of 12 plain + 12 transition frames
        # (about 10 s on 16 cores) + one full frame to validate the x16
        np_src = {k: v.numpy() for k, v in host_src.items()} if host_src else c5_sources_numpy(cfg)
        cf = c5_frames(0, 1, cfg, 24)
        info_cpu = c5_cpu_arm(cfg, np_src, [cf[0]] + cf[:23] + [last_frame], 1, full=False)
        cpu = {"value": info_cpu["fps"], "unit": "frames/s", "cores": info_cpu["threads"], "kind": "port", "sample": info_cpu["sample"],
               "seconds": info_cpu["seconds"], "full_frame_check": info_cpu["full_frame_check"]}
        if host_src:
            # free parity check at the full size: the GPU frame the resident loop rendered last vs the oracle's rows of it
            from oracle import oracle as O

            rows = cfg["height"] // BAND_DIV
            want = c5_cpu_frame(O, np_src, workloads.stack_layers(last_frame, **cfg), rows)
            got = out_host.numpy()[:rows - 8]  # the band's last rows see a clamped edge the full frame does not have
            cpu["parity_max_abs_vs_gpu_frame"] = float(abs(got - want[:rows - 8]).max())
    del out_host

    if rank == 0:
        line = {"metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": c5_config_json(cfg, world),
                "api": "toucan::ImageGraph::exec + IImageNode::exec through libtoucan_render.so (OpenFX host + 6 .ofx modules)",
                "hbm_gbs_algorithmic": achieved * world, "sustained": sus,
                "roofline": {"bound": "hbm", "kernel": "k_stack_f32", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                             "frac_sustained": sus["roofline_frac"] if sus else None, "per_frame_kind": kinds,
                             "traffic": measured_traffic([nsrc[f] for f in frames[args.warmup:]], cfg["tracks"]) if args.scale == 1.0 else None,
                             "traffic_source": "profiles/ stack traffic json (ncu dram__bytes_read+write per launch, plain / "
                                               "transition frames, weighted by this run's frame mix)",
                             "algorithmic_bytes_per_launch": alg_bytes / args.steps,
                             "note": f"bytes = (sources referenced + 1 output) x {F:,} B per frame; one launch per frame; "
                                     "rank-0-equivalent (max-over-ranks time)"},
                "cpu_baseline": cpu, "e2e": e2e, "e2e_y4m444": e2e_y4m, "e2e_reupload_every_frame": e2e_reupload, "gpu_launches": int(launches), "clocks": clk,
                "checksum": checksum}
        if args.scale != 1.0:
            line["INVALID"] = "debug run at reduced frame size (--scale)"
        print(json.dumps(line), flush=True)
    D.close()


def measured_traffic(nsrc_per_launch, tracks):
    """DRAM bytes per launch of the fused stack kernel from the committed ncu capture (newest round first), averaged over
    this run's mix of plain (n sources = tracks) and transition (2 x tracks) frames; None if absent."""
    for name in ("r02_stack_traffic.json", "r01_stack_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            plain = sum(1 for n in nsrc_per_launch if n == tracks)
            trans = sum(1 for n in nsrc_per_launch if n == 2 * tracks)
            if plain + trans != len(nsrc_per_launch) or tracks != 10:
                return None
            return (plain * t["plain_frame_bytes"] + trans * t["transition_frame_bytes"]) / len(nsrc_per_launch)
        except Exception:
            continue
    return None


def run_small(args):
    """c2 - c4 through the same API: media resident in HBM for `value`, pinned host media + cold source cache for e2e."""
    D = Dist()
    torch = D.torch
    import numpy as np

    from toucan_b200 import capi, render

    wl = small_workload(args.workload)
    rank, world, local = D.rank, D.world, D.local
    ofx = render.Host()
    # three timelines over three copies of the media: consecutive timed frames read different memory (c2's 33 MB source
    # would otherwise sit in the 126 MB L2)
    copies = 3
    tls, keep = [], []
    for c in range(copies):
        tl = render.Timeline(json_doc=wl["doc"], directory=wl["directory"])
        for url, a in wl["media"].items():
            t = torch.from_numpy(a).to(D.dev).clone()
            keep.append(t)
            tl.set_media(os.path.join(wl["directory"], url), t)
        tl.prepare()
        tls.append(tl)
    info = tls[0].info()
    n = info["frames"]
    stream = torch.cuda.Stream(D.dev)
    torch.cuda.set_stream(stream)
    render.bind_stream(local, stream.cuda_stream)
    lo, hi = workloads.frame_block(rank, world, n)
    if hi <= lo:
        lo, hi = 0, n
    seq = [(i % copies, info["start"] + lo + i % (hi - lo)) for i in range(args.steps + args.warmup)]

    def step(it):
        return tls[it[0]].render_resident(ofx, it[1])

    # set-up (not a step): every distinct frame of the block once per media copy, so one-time work -- per-geometry resize tap tables,
    # colour-map knots, pool growth for each frame size -- is not inside the W + K steps
    for c in range(copies):
        for f in range(lo, hi):
            tls[c].render_resident(ofx, info["start"] + f)
    torch.cuda.synchronize()
    clocks = Clocks(local) if rank == 0 else None
    ms, last = timed_loop(D, step, seq[args.warmup:], seq[:args.warmup])
    launches = D.launches
    _, ow, oh, ot = last
    src_bytes = sum(a.nbytes for a in wl["media"].values())
    out_bytes = ow * oh * 4 * {capi.U8: 1, capi.U16: 2, capi.F16: 2, capi.F32: 4}[ot]
    alg = src_bytes + out_bytes  # leaf frames read once + final frame written once
    peak, peak_src = measured_peak()
    fps = world * args.steps / (ms * 1e-3)
    achieved = alg * args.steps / (ms * 1e-3) / 1e9
    sus = None
    if os.environ.get("TOUCAN_BENCH_SUSTAINED", "1") != "0":
        n_sus = max(args.steps, int(1000.0 / max(1e-4, ms / args.steps)))
        sseq = [seq[args.warmup + i % args.steps] for i in range(n_sus)]
        sms, _ = timed_loop(D, step, sseq, [])
        sus = {"value": world * n_sus / (sms * 1e-3), "unit": "frames/s", "steps": n_sus, "ms_per_step": sms / n_sus,
               "roofline_frac": alg * n_sus / (sms * 1e-3) / 1e9 / peak}
    clk = clocks.stop() if clocks else None
    for tl in tls:
        tl.close()
    keep.clear()

    # e2e: pinned host media, cold source cache, frames downloaded
    e2e_steps = min(args.steps, max(2, hi - lo))
    tl2 = render.Timeline(json_doc=wl["doc"], directory=wl["directory"])
    pinned = {}
    for url, a in wl["media"].items():
        pinned[url] = torch.from_numpy(a).pin_memory()
        tl2.set_media(os.path.join(wl["directory"], url), pinned[url])
    tl2.set_source_cache(64 << 30)
    tl2.prepare()
    dt, _, up_bytes, up_n = e2e_run(D, tl2, lo, min(e2e_steps, hi - lo), lo)
    e2e_n = min(e2e_steps, hi - lo)
    e2e = {"value": world * e2e_n / dt, "unit": "frames/s", "h2d_bytes_per_step": int(D.sum(float(up_bytes)) / (world * e2e_n)),
           "d2h_bytes_per_step": out_bytes, "steps": e2e_n,
           "note": "toucan::FrameScheduler from pinned host media with a cold device source cache (each still uploaded once inside the timed "
                   "region), every finished frame downloaded to the ordered writer; wall clock, max over ranks"}
    tl2.close()

    cpu = None
    if rank == 0 and world == 1 and os.environ.get("TOUCAN_BENCH_CPU", "1") != "0":
        ic = small_cpu_arm(wl, 8, 1)
        cpu = {"value": ic["fps"], "unit": "frames/s", "cores": ic["threads"], "kind": "port", "sample": ic["sample"], "seconds": ic["seconds"]}
    if rank == 0:
        line = {"metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"],
                "data": "synthetic", "config": small_config_json(wl, world),
                "api": "toucan::ImageGraph::exec + IImageNode::exec through libtoucan_render.so (OpenFX host + 6 .ofx modules)",
                "sustained": sus,
                "roofline": {"bound": "hbm", "kernel": "whole frame (all launches of one ImageGraph evaluation)", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                             "frac_sustained": sus["roofline_frac"] if sus else None, "traffic": None,
                             "algorithmic_bytes_per_launch": alg,
                             "launches_per_frame": launches / args.steps,
                             "note": "SURVEY 8d graph-minimum bytes per frame = leaf source frames read once + final frame written once; "
                                     "intermediate node outputs count zero"},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line), flush=True)
    del np
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "s4k", "c2", "c3", "c4"])
    ap.add_argument("--scale", type=float, default=1.0, help="debug: shrink the c5 frame (NOT a valid bench number)")
    args = ap.parse_args()
    cfg = dict(workloads.C5)
    if args.scale != 1.0:
        cfg["width"] = max(64, int(cfg["width"] * args.scale))
        cfg["height"] = max(64, int(cfg["height"] * args.scale) // BAND_DIV * BAND_DIV)
    if args.workload == "s4k":
        cfg.update(width=3840, height=2160)
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.workload in ("c5", "s4k"):
        run_c5(args, cfg)
    else:
        run_small(args)


if __name__ == "__main__":
    main()
