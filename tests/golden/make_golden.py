"""Regenerates tests/golden/ from the reference's fixtures and the CPU oracle.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_golden.py

  tests/golden/data/       the reference's data/*.otio timelines and raster media, copied verbatim
                           (inputs that define the workloads; /root/reference does not exist on the GPU box)
  tests/golden/ref_images/ the reference's film strips images/*.png, copied verbatim: pixels the REFERENCE rendered
                           (tests/test_reference_filmstrips.py pins the oracle to them at thumbnail scale)
  tests/golden/frames.json per timeline and frame: the node tree ImageGraph::exec builds (text form) and the
                           sha256 + shape + dtype of the frame the CPU oracle renders from the decoded media
The reference's own tests hold no golden pixels (tests/toucanRenderTest/ImageGraphTest.cpp:62-74 only
writes PNGs), so these vectors pin the ORACLE (regression) and the time math / graph structure; they do
not pin parity with a real toucan-render build ("parity unpinned").
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/data"

from oracle import graph as G  # noqa: E402
from tests import media  # noqa: E402

SKIP = set()  # Draw.otio: Box and Line are drawn; Text copies its source (no FreeType / font: what render_text leaves when it finds no font)


def main():
    dst = os.path.join(HERE, "data")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(REF)):
        if f.endswith((".otio", ".png", ".jpg")):
            shutil.copyfile(os.path.join(REF, f), os.path.join(dst, f))
    # the reference's own renders: README film strips made by bin/toucan-filmstrip (120x67 tiles at a 140 px stride)
    strips = os.path.join(HERE, "ref_images")
    os.makedirs(strips, exist_ok=True)
    for f in sorted(os.listdir("/root/reference/images")):
        if f.endswith(".png") and not f.startswith("toucan-view"):
            shutil.copyfile(os.path.join("/root/reference/images", f), os.path.join(strips, f))
    out = {}
    for f in sorted(os.listdir(dst)):
        if not f.endswith(".otio") or f in SKIP:
            continue
        tl = G.Timeline(os.path.join(dst, f))
        g = G.ImageGraph(tl, media.decode)
        frames = []
        for t in tl.frames():
            node = g.exec(t)
            img = g.eval(node)
            rec = {"time": t.value, "tree": G.describe_tree(node)}
            if img is not None and img.size:
                rec.update(shape=list(img.shape), dtype=str(img.dtype), sha256=hashlib.sha256(img.tobytes()).hexdigest())
            else:
                rec.update(shape=[0, 0, 4], dtype="uint8", sha256="")
            frames.append(rec)
        out[f] = {"size": list(g.size), "dtype": g.dtype, "rate": tl.time_range.duration.rate, "start": tl.time_range.start.value,
                  "frames": frames}
        print(f, len(frames), "frames", g.size, g.dtype, flush=True)
    with open(os.path.join(HERE, "frames.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
