import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from toucan_b200 import capi, render, workloads
w, h = 3840, 2160
frames = 240
d = "/tmp/cli_probe"; os.makedirs(d + "/stills", exist_ok=True)
text = json.dumps(workloads.stack_timeline(tracks=10, frames=frames))
rng = np.random.default_rng(5)
for t in range(10):
    for k in range(2):
        rel = f"stills/t{t}_k{k}.exr"
        render.write_exr(os.path.join(d, rel), rng.random((h, w, 4), dtype=np.float32), zip=False)
        text = text.replace(workloads.still_url(t, k), rel)
open(d + "/stack.otio", "w").write(text)
capi.check(capi.lib().tc_set_device(0), "dev")
render.bind_device(0)
host = render.Host()
t0 = time.perf_counter()
tl = render.Timeline(d + "/stack.otio")
tl.set_source_cache(8 << 30)
tl.prepare()
print("open+prepare", round(time.perf_counter() - t0, 3))
for name, kw in (("y4m", dict(y4m_format="422")), ("raw", {})):
    stamps = []
    t0 = time.perf_counter()
    tl.render_range(0, frames, lambda f, p: stamps.append(time.perf_counter()), **kw)
    dt = time.perf_counter() - t0
    s = np.diff(np.array(stamps))
    print(name, "total", round(dt, 3), "first frame at", round(stamps[0] - t0, 3), "median gap ms", round(float(np.median(s)) * 1e3, 3),
          "p90", round(float(np.percentile(s, 90)) * 1e3, 3), "max", round(float(s.max()) * 1e3, 3), "sum of gaps", round(float(s.sum()), 3))
    big = np.argsort(s)[-8:]
    print("   largest gaps after frames", sorted(big.tolist()), [round(float(s[i]) * 1e3, 1) for i in sorted(big.tolist())])
