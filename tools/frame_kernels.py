#!/usr/bin/env python
"""Render ONE frame of a config timeline (after a warm-up frame) so an ncu launch list shows which kernels a frame is made of.

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/frame_kernels.py C3
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import json  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import timelines as TL  # noqa: E402
from toucan_b200 import capi, render, workloads  # noqa: E402
from tools.config_bench import resized, synth  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "C3"
    capi.check(capi.lib().tc_set_device(0), "tc_set_device")
    render.bind_device(0)
    host = render.Host()
    if which == "C3":
        doc = json.load(open(os.path.join(TL.DATA, "MultipleEffects.otio")))
        media = {os.path.join(TL.DATA, "Charlie.jpg"): synth(3840, 2160, np.float16, 1002)}
        frame = 3.0
    elif which == "C3c":
        from tools.config_bench import resized as _rs
        doc = _rs(json.load(open(os.path.join(TL.DATA, "CompositeTracks.otio"))), (3840, 2160))
        media = {os.path.join(TL.DATA, f"Counter.{i}.png"): synth(3840, 2160, np.float16, 1100 + i, disc=True) for i in range(10)}
        frame = 3.0
    elif which == "C4":
        doc = json.load(open(os.path.join(TL.DATA, "Transform.otio")))
        media = {os.path.join(TL.DATA, "Charlie.jpg"): synth(3840, 2160, np.float32, 1003)}
        frame = float(sys.argv[2]) if len(sys.argv) > 2 else 7.0
    else:
        doc = workloads.filter_chain_timeline(frames=8, url="source.exr")
        media = {os.path.join(TL.DATA, "source.exr"): synth(1920, 1080, np.float32, 1001)}
        frame = 2.0
    tl = render.Timeline(json_doc=doc, directory=TL.DATA)
    keep = []
    for p, a in media.items():
        t = torch.from_numpy(a).cuda()
        keep.append(t)
        tl.set_media(p, t)
    tl.prepare()
    print(tl.describe(host, frame))
    tl.render_resident(host, frame)
    torch.cuda.synchronize()
    n0 = capi.launch_count()
    tl.render_resident(host, frame)
    torch.cuda.synchronize()
    print("launches", capi.launch_count() - n0)


if __name__ == "__main__":
    main()
