#!/usr/bin/env python
"""Per-frame GPU time (CUDA events) of the C3 CompositeTracks config, fusion on and off."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import timelines as TL  # noqa: E402
from toucan_b200 import capi, render  # noqa: E402
from tools.config_bench import resized, synth  # noqa: E402

capi.check(capi.lib().tc_set_device(0), "tc_set_device")
render.bind_device(0)
host = render.Host()
doc = resized(json.load(open(os.path.join(TL.DATA, "CompositeTracks.otio"))), (3840, 2160))
tl = render.Timeline(json_doc=doc, directory=TL.DATA)
keep = []
for i in range(10):
    t = torch.from_numpy(synth(3840, 2160, np.float16, 1100 + i, disc=True)).cuda()
    keep.append(t)
    tl.set_media(os.path.join(TL.DATA, f"Counter.{i}.png"), t)
tl.prepare()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
render.bind_stream(0, stream.cuda_stream)
for fuse in (True, False, True):
    render.set_fusion(fuse)
    for f in range(10):
        tl.render_resident(host, float(f))
    torch.cuda.synchronize()
    out = []
    for f in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(5):
            tl.render_resident(host, float(f))
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / 5
        torch.cuda.synchronize()
        out.append((round(e0.elapsed_time(e1) / 5, 3), round(host_ms, 3)))
    print("fusion", fuse, out)
