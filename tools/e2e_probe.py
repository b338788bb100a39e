import sys, time, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from toucan_b200 import capi, render, workloads
torch.cuda.set_device(0)
capi.check(capi.lib().tc_set_device(0), "x")
cfg = dict(workloads.C5)
w, h = cfg["width"], cfg["height"]
doc = workloads.stack_timeline(**cfg)
host_src = {}
for t in range(10):
    for k in range(2):
        host_src[(t, k)] = torch.rand((h, w, 4), dtype=torch.float32).pin_memory() if (t, k) == (0, 0) else host_src[(0, 0)].clone().pin_memory()
tl2 = render.Timeline(json_doc=doc, directory="/synthetic")
for (t, k), a in host_src.items():
    tl2.set_media(tl2.media_path(workloads.still_url(t, k)), a)
tl2.prepare()
ts = []
def writer(frame, pixels):
    ts.append((frame, time.perf_counter()))
t0 = time.perf_counter()
tl2.render_range(24, 60, writer, devices=[0])
print("total", time.perf_counter() - t0)
prev = t0
for f, t in ts:
    print(f, round((t - prev) * 1e3, 1), "ms")
    prev = t
# raw H2D bandwidth check
d = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    d.copy_(host_src[(0, 0)], non_blocking=True)
torch.cuda.synchronize()
print("H2D GB/s", 10 * w * h * 16 / (time.perf_counter() - t0) / 1e9)
