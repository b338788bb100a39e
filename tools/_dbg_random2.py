import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import graph as G
from oracle import oracle as O
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
for seed, times in [(9, [36.0, 37.0]), (14, [35.0])]:
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
    render.set_fusion(False)
    tl = render.Timeline(json_doc=doc, directory="/nonexistent")
    by_path = {}
    for u, a in stills.items():
        by_path[tl.media_path(u)] = a
        tl.set_media(tl.media_path(u), a)
    tl.prepare()
    otl = G.Timeline(doc); otl.dir = "/nonexistent"
    og = G.ImageGraph(otl, lambda p: by_path[p])
    for t in otl.frames():
        if t.value not in times: continue
        node = og.exec(t)
        want = og.eval(node)
        got = tl.render(host, t.value, max_bytes=64 * 64 * 16)
        d = np.abs(got.astype(np.float64) - want.astype(np.float64)).max(axis=2)
        ys, xs = np.where(d > 2e-5)
        print("seed", seed, "t", t.value, "bad", len(ys))
        for k in range(0, len(ys), max(1, len(ys) // 6)):
            y, x = ys[k], xs[k]
            print("  px", y, x, "got", got[y, x], "want", want[y, x])
        # evaluate oracle subtrees
        def walk(n, depth=0):
            if n is None:
                print("  " * depth + "(null)"); return
            img = og.eval(n)
            kind = n[0]
            label = n[1] if kind == "effect" else kind
            print("  " * depth + f"{label} -> {None if img is None else (img.shape, img.dtype, float(np.nanmin(img)) if img.size else None, float(np.nanmax(img)) if img.size else None)}")
            ins = n[3] if kind == "effect" else (n[1] if kind == "comp" else [])
            for i in ins if isinstance(ins, (list, tuple)) else []:
                walk(i, depth + 1)
        try:
            walk(node)
        except Exception as e:
            print("walk failed", e, node[:2] if node else None)
    tl.close()
render.set_fusion(True)
