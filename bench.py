#!/usr/bin/env python
"""bench.py -- frames/s and HBM GB/s of the toucan render hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5]

Workload (BASELINE.json configs[4], the configuration the 1/2/4/8-GPU metric is quoted on):
synthetic 1000-frame, 10-track 7680x4320 RGBA float32 timeline, toucan:Blur radius 10 on
every clip, toucan:Dissolve between clips (half of the frames are inside a transition),
tracks composited premult+over on black.  A "step" is one timeline frame.

  value     frames/s with the decoded source stills resident in HBM (whole job, all ranks)
  e2e       frames/s through the host-buffer API: every step uploads the frame's sources
            from pinned host memory, renders, and downloads the finished frame
  roofline  the fused stack kernel against the measured HBM copy peak
  cpu_baseline / --impl reference
            the CPU restatement of the reference's OIIO path (oracle/, "port": the real
            toucan-render cannot be built here) on all host cores, on a 1/16-frame band

Frames are sharded in contiguous blocks across ranks (no collective on the data path;
torch.distributed is used for the barrier and the max-over-ranks time only).
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

BAND_DIV = 16  # CPU arms render a 1/16-frame band per step


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
# CPU arm: the oracle on a band of the frame (bench.py is one of the places allowed to run oracle/)
def cpu_band_setup(cfg, div=BAND_DIV):
    import numpy as np

    w, h = cfg["width"], cfg["height"] // div
    src = {}
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for t in range(cfg["tracks"]):
        for k in range(2):
            rng = np.random.default_rng(2000 + 10 * t + k)
            a = rng.random((h, w, 4), dtype=np.float32)
            r = np.hypot((xx - w * (0.3 + 0.04 * t)) / w, (yy - h * 0.5) / (cfg["height"]))
            a[..., 3] = np.clip(1.2 - 2.0 * r, 0.0, 1.0)
            src[(t, k)] = a
    return w, h, src


def cpu_band_frame(O, w, h, src, layers):
    """The unfused node chain the reference executes per frame (ImageGraph.cpp:165-198)."""
    acc = O.fill(w, h, [0.0, 0.0, 0.0, 1.0], O.F32)
    for a, b, ra, rb, v in layers:
        x = O.blur(src[a], ra)
        if b is not None:
            x = O.dissolve(x, O.blur(src[b], rb), v)
        acc = O.comp(x, acc, True)
    return acc


def cpu_arm(cfg, frames, warmup, target_step_s=None):
    from oracle import oracle as O

    O.lib()
    div = BAND_DIV
    w, h, src = cpu_band_setup(cfg, div)
    cores = os.cpu_count() or 1
    if target_step_s:
        # bounded sample: shrink the band until one step fits the time budget (per-pixel cost is unchanged)
        t0 = time.perf_counter()
        cpu_band_frame(O, w, h, src, workloads.stack_layers(frames[-1], **cfg))
        dt = time.perf_counter() - t0
        while dt > target_step_s and div < 256:
            div *= 2
            dt /= 2
        if div != BAND_DIV:
            w, h, src = cpu_band_setup(cfg, div)
    for f in frames[:warmup]:
        cpu_band_frame(O, w, h, src, workloads.stack_layers(f, **cfg))
    t0 = time.perf_counter()
    for f in frames[warmup:]:
        cpu_band_frame(O, w, h, src, workloads.stack_layers(f, **cfg))
    dt = time.perf_counter() - t0
    n = len(frames) - warmup
    fps = n / (dt * div)
    return fps, dt, cores, f"{n} frames x rows 0..{h - 1} of {cfg['height']} (1/{div} of each {cfg['width']}x{cfg['height']} frame), time scaled x{div}"


def measured_traffic(nsrc_per_launch, tracks):
    """DRAM bytes per launch of the fused stack kernel from the committed ncu capture, averaged over this
    run's mix of plain (n sources = tracks) and transition (2 x tracks) frames; None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_stack_traffic.json")) as f:
            t = json.load(f)
        plain = sum(1 for n in nsrc_per_launch if n == tracks)
        trans = sum(1 for n in nsrc_per_launch if n == 2 * tracks)
        if plain + trans != len(nsrc_per_launch) or tracks != 10:
            return None
        return (plain * t["plain_frame_bytes"] + trans * t["transition_frame_bytes"]) / len(nsrc_per_launch)
    except Exception:
        return None


def frames_for(rank, world, cfg, count):
    """Frames a rank renders, inside its own contiguous block.  The window starts in the middle of a clip
    and wraps at a multiple of half a clip, so every rank (and every multiple of 24 steps) sees the same
    half-plain / half-transition mix."""
    lo, hi = workloads.frame_block(rank, world, cfg["frames"])
    period, half = cfg["clip_frames"], cfg["clip_frames"] // 2
    start = lo + (half - lo) % period
    avail = ((hi - start) // half) * half
    if avail < half:
        start, avail = lo, max(1, hi - lo)
    return [start + s % avail for s in range(count)]


def config_json(cfg, world):
    return {"workload": "c5: synthetic 1000-frame 10-track 7680x4320 RGBA float32 timeline, blur(radius 10)+dissolve per track, "
                        "premult+over stack (BASELINE.json configs[4])",
            "width": cfg["width"], "height": cfg["height"], "tracks": cfg["tracks"], "frames": cfg["frames"],
            "sharding": f"contiguous frame blocks over {world} GPU(s), no collective",
            "l2": "inputs larger than L2 (each source frame 531 MB, 10-20 per step)"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = frames_for(0, 1, cfg, args.steps + args.warmup)
    fps, dt, cores, sample = cpu_arm(cfg, frames, args.warmup, target_step_s=90.0 / max(1, args.steps + args.warmup))
    line = {"impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / fps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_json(cfg, 1),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the toucan/OIIO path (oracle/); the real toucan-render needs OIIO/OCIO/OTIO/OpenFX "
                    "which are not installed and cannot be fetched"}
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


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    from toucan_b200 import capi, render

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the render path has no CPU fallback")
    torch.cuda.set_device(local)
    capi.check(capi.lib().tc_set_device(local), "tc_set_device")
    if world > 1:
        bind_to_gpu_cpus(torch, local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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

    frames = frames_for(rank, world, cfg, args.steps + args.warmup)
    nsrc = [sum(1 + (la[1] is not None) for la in workloads.stack_layers(f, **cfg)) for f in frames]
    start = info["start"]

    for f in frames[:args.warmup]:
        tl.render_resident(ofx, start + f)
    barrier()
    clocks = Clocks(local) if rank == 0 else None
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in frames[args.warmup:]:
        ptr, ow, oh, ot = tl.render_resident(ofx, start + f)
    e1.record()
    barrier()
    launches = capi.launch_count() - l0
    clk = clocks.stop() if clocks else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    fps = world * args.steps / (ms * 1e-3)
    alg_bytes = sum((n + 1) * F for n in nsrc[args.warmup:])
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    out_host = torch.empty((h, w, 4), dtype=torch.float32).pin_memory()
    capi.check(capi.lib().tc_download(ctypes.c_void_p(out_host.data_ptr()), ctypes.c_void_p(ptr), F,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "tc_download")
    torch.cuda.synchronize()
    checksum = float(out_host[::97, ::89].double().sum().item())
    del out_host

    # ---- end to end: host buffers in, host buffers out, every copy inside the timed region ----
    e2e = None
    e2e_cached = None
    e2e_steps = int(os.environ.get("TOUCAN_BENCH_E2E_STEPS", str(min(args.steps, 24))))
    sink = {"n": 0, "acc": 0.0}

    def writer(frame, pixels):
        sink["n"] += 1
        sink["acc"] += float(pixels[0, 0, 0])  # the writer's read of the finished frame

    if e2e_steps > 0:
        # (informational) the deployment shape for stills: sources cached in HBM, every frame downloaded
        fr = frames_for(rank, world, cfg, e2e_steps)
        tl.render_range(fr[12 % len(fr)], fr[12 % len(fr)] + 2, writer, devices=[local])
        barrier()
        t0 = time.perf_counter()
        tl.render_range(fr[0], fr[0] + e2e_steps, writer, devices=[local])
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e_cached = {"value": world * e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": F,
                      "note": "informational: source stills cached in HBM (the scheduler's source cache), finished frames downloaded"}
        sink["n"] = 0
        tl.close()
        host_src = {k: v.cpu().pin_memory() for k, v in src.items()}
        src.clear()
        torch.cuda.empty_cache()
        capi.lib().tc_pool_trim()
        tl2 = render.Timeline(json_doc=doc, directory="/synthetic")
        for (t, k), a in host_src.items():
            tl2.set_media(tl2.media_path(workloads.still_url(t, k)), a)  # pinned host: uploaded on every frame that reads it
        tl2.prepare()
        fr = frames_for(rank, world, cfg, e2e_steps)
        # warm-up on two mid-transition frames (20 sources): pool growth, pinned ring, plug-in load
        tl2.render_range(fr[12 % len(fr)], fr[12 % len(fr)] + 2, writer, devices=[local])
        h2d = sum(sum(1 + (la[1] is not None) for la in workloads.stack_layers(f, **cfg)) for f in range(fr[0], fr[0] + e2e_steps)) * F
        barrier()
        t0 = time.perf_counter()
        tl2.render_range(fr[0], fr[0] + e2e_steps, writer, devices=[local])
        dt = time.perf_counter() - t0
        barrier()
        dt = max_over_ranks(dt)
        assert sink["n"] == e2e_steps + 2
        e2e = {"value": world * e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": h2d // e2e_steps,
               "d2h_bytes_per_step": F, "steps": e2e_steps,
               "note": "toucan::FrameScheduler through the C API: every step uploads all source frames the frame references "
                       "(pinned host -> HBM, as the reference re-reads its media every frame), renders, and downloads the finished "
                       "float32 frame into a pinned ring for the ordered writer; wall clock, max over ranks"}

    cpu = None
    if rank == 0 and world == 1 and os.environ.get("TOUCAN_BENCH_CPU", "1") != "0":
        cf = frames_for(0, 1, cfg, 24)
        # one whole 24-frame window (12 plain + 12 transition frames, the mix the timed region renders), 1 warm-up:
        # about 10 s of CPU work on 16 cores
        cfps, cdt, cores, sample = cpu_arm(cfg, [cf[0]] + cf, 1)
        cpu = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample, "seconds": cdt}

    if rank == 0:
        line = {"metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_json(cfg, world),
                "api": "toucan::ImageGraph::exec + IImageNode::exec through libtoucan_render.so (OpenFX host + 6 .ofx modules)",
                "hbm_gbs_algorithmic": achieved * world,
                "roofline": {"bound": "hbm", "kernel": "k_stack_f32", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                             "traffic": measured_traffic(nsrc[args.warmup:], cfg["tracks"]) if args.scale == 1.0 else None,
                             "traffic_source": "profiles/r01_stack_traffic.json (ncu dram__bytes_read+write per launch, plain / "
                                               "transition frames, weighted by this run's frame mix)",
                             "algorithmic_bytes_per_launch": alg_bytes / args.steps,
                             "note": "bytes = (sources referenced + 1 output) x 530,841,600 B per frame; one launch per frame; "
                                     "rank-0-equivalent (max-over-ranks time)"},
                "cpu_baseline": cpu, "e2e": e2e, "e2e_sources_cached": e2e_cached, "gpu_launches": int(launches), "clocks": clk, "checksum": checksum}
        if args.scale != 1.0:
            line["INVALID"] = "debug run at reduced frame size (--scale)"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="debug: shrink the frame (NOT a valid bench number)")
    args = ap.parse_args()
    cfg = dict(workloads.C5)
    if args.scale != 1.0:
        cfg["width"] = max(64, int(cfg["width"] * args.scale))
        cfg["height"] = max(64, int(cfg["height"] * args.scale) // BAND_DIV * BAND_DIV)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
