"""Per-GPU frame scheduler for stack timelines: host frames in, host frames out, in order.

Replaces the serial, one-frame-in-flight loop of bin/toucan-render/App.cpp:222-264 for one
GPU worker: the decoded source frames of frame i+1 are uploaded (H2D from pinned memory)
while frame i renders and frame i-1 is downloaded (D2H into pinned memory) for the
writer.  Three streams, two buffer sets, no allocation inside the loop.  torch supplies
device memory, streams and events only; the render itself is the fused sm_100a kernel
behind tc_stack_f32.
"""
from __future__ import annotations

import torch

from . import ops


class StackFrameScheduler:
    def __init__(self, width: int, height: int, max_sources: int, device="cuda", background=(0.0, 0.0, 0.0, 1.0)):
        self.w, self.h = width, height
        self.background = list(background)
        self.device = torch.device(device)
        self.src = [[torch.empty((height, width, 4), dtype=torch.float32, device=self.device) for _ in range(max_sources)]
                    for _ in range(2)]
        self.out = [torch.empty((height, width, 4), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.out_host = [torch.empty((height, width, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def render(self, frames, host_sources):
        """frames: iterable of layer lists [(key_a, key_b|None, radius_a, radius_b, value), ...];
        host_sources: key -> pinned float32 tensor (h, w, 4).  Yields (index, pinned result)
        in order; a yielded tensor is valid until the next frame is requested."""
        ev_in = [None, None]
        ev_run = [None, None]
        ev_out = [None, None]
        pending = None
        for i, layers in enumerate(frames):
            b = i & 1
            keys = []
            for la in layers:
                for k in (la[0], la[1]):
                    if k is not None and k not in keys:
                        keys.append(k)
            slot = {k: self.src[b][j] for j, k in enumerate(keys)}
            with torch.cuda.stream(self.s_in):
                if ev_run[b] is not None:
                    self.s_in.wait_event(ev_run[b])  # set b was last read by frame i-2
                for k in keys:
                    slot[k].copy_(host_sources[k], non_blocking=True)
                    self.h2d_bytes += slot[k].numel() * 4
                ev_in[b] = torch.cuda.Event()
                ev_in[b].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_in[b])
                if ev_out[b] is not None:
                    self.s_run.wait_event(ev_out[b])  # out[b] was last downloaded for frame i-2
                dev_layers = [(slot[la[0]], None if la[1] is None else slot[la[1]], la[2], la[3], la[4]) for la in layers]
                ops.stack_f32(self.w, self.h, self.background, dev_layers, out=self.out[b])
                ev_run[b] = torch.cuda.Event()
                ev_run[b].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run[b])
                self.out_host[b].copy_(self.out[b], non_blocking=True)
                self.d2h_bytes += self.out[b].numel() * 4
                ev_out[b] = torch.cuda.Event()
                ev_out[b].record(self.s_out)
            if pending is not None:
                pi, pb, pev = pending
                pev.synchronize()
                yield pi, self.out_host[pb]
            pending = (i, b, ev_out[b])
        if pending is not None:
            pi, pb, pev = pending
            pev.synchronize()
            yield pi, self.out_host[pb]
