import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import graph as G
from tests.test_host_graph import _random_timeline
from toucan_b200 import render
render.bind_device(0)
host = render.Host()
def tame(effects):
    out = []
    for e in effects:
        n = e.get("effect_name", "")
        if n in ("toucan:Pow", "toucan:ColorMap", "toucan:UnsharpMask", "toucan:Rotate"):
            e = dict(e, effect_name="toucan:Invert", name="toucan:Invert", metadata={})
        out.append(e)
    return out
for seed, fusion in [(8, True), (8, False), (11, True), (9, False), (14, False)]:
    rng = np.random.default_rng(9000 + seed)
    doc = _random_timeline(rng)
    doc["tracks"]["effects"] = tame(doc["tracks"]["effects"])
    stills = {}
    for t in doc["tracks"]["children"]:
        t["effects"] = tame(t["effects"])
        for c in t["children"]:
            if c["OTIO_SCHEMA"] == "Clip.2":
                c["effects"] = tame(c["effects"])
                u = c["media_references"]["DEFAULT_MEDIA"]["target_url"]
                if u not in stills:
                    a = rng.random((30, 48, 4), dtype=np.float32)
                    a[..., 3] = np.where(rng.random((30, 48)) < 0.3, 1.0, a[..., 3])
                    stills[u] = np.ascontiguousarray(a)
    render.set_fusion(fusion)
    tl = render.Timeline(json_doc=doc, directory="/nonexistent")
    by_path = {}
    for u, a in stills.items():
        by_path[tl.media_path(u)] = a
        tl.set_media(tl.media_path(u), a)
    tl.prepare()
    otl = G.Timeline(doc); otl.dir = "/nonexistent"
    og = G.ImageGraph(otl, lambda p: by_path[p])
    shown = 0
    for t in otl.frames():
        want = og.render(t)
        try:
            got = tl.render(host, t.value, max_bytes=64 * 64 * 16)
        except Exception as e:
            print('seed', seed, 'fusion', fusion, 't', t.value, 'EXC', e)
            print(tl.describe(host, t.value))
            shown += 1
            if shown >= 2: break
            continue
        if want is None or want.size == 0:
            continue
        if got.shape != want.shape:
            print("seed", seed, fusion, "t", t.value, "SHAPE", got.shape, want.shape); shown += 1
        else:
            d = np.abs(got.astype(np.float64) - want.astype(np.float64))
            if d.max() > 2e-5:
                ys, xs = np.where(d.max(axis=2) > 2e-5)
                print("seed", seed, "fusion", fusion, "t", t.value, "max", d.max(), "bad px", len(ys), "of", d.shape[0] * d.shape[1], "bbox", ys.min(), ys.max(), xs.min(), xs.max())
                print(tl.describe(host, t.value))
                shown += 1
        if shown >= 2:
            break
    tl.close()
render.set_fusion(True)
