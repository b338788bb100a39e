"""CPU: the C++ host (ImageEffectHost + OpenFX modules + TimelineWrapper + ImageGraph) builds,
for every frame of every fixture timeline, the same node tree as the oracle's restatement of
lib/toucanRender/ImageGraph.cpp:144-466 (golden trees in tests/golden/frames.json).  No GPU:
only graph construction and the OpenFX describe actions run."""
import os

import pytest

from tests import timelines as TL

GOLD = TL.golden()

IDENTIFIERS = {
    "toucan:Checkers", "toucan:Fill", "toucan:Gradient", "toucan:Noise",
    "toucan:Blur", "toucan:ColorMap", "toucan:Invert", "toucan:Pow", "toucan:Saturate", "toucan:UnsharpMask",
    "toucan:Crop", "toucan:Flip", "toucan:Flop", "toucan:Resize", "toucan:Rotate",
    "toucan:Dissolve", "toucan:HorizontalWipe", "toucan:VerticalWipe",
    "toucan:ColorConvert", "toucan:Premult", "toucan:Unpremult",
    "toucan:Box", "toucan:Line", "toucan:Text",  # SURVEY 8 row f3 (Text copies its source: no FreeType / font here)
}


@pytest.fixture(scope="module")
def host(built):
    from toucan_b200 import render

    return render.Host()


def test_render_abi_exports(built):
    import re

    from toucan_b200 import render

    lib = render.lib()
    txt = open(os.path.join(os.path.dirname(TL.GOLDEN), "..", "include", "toucan_render.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"\b(tr_[a-z0-9_]+)\s*\(", txt)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in toucan_render.h but not exported"
    assert set(render.EXPORTS) == set(names)


def test_host_loads_all_effects(host):
    # the 21 in-scope plug-in identifiers of SURVEY 8b plus the three Draw effects, served by 6 *.ofx modules
    assert set(host.identifiers()) == IDENTIFIERS


def test_ofx_modules_export_entry_points(built):
    import ctypes

    from toucan_b200 import capi, render

    capi.lib()
    mods = sorted(f for f in os.listdir(render.PLUGIN_DIR) if f.endswith(".ofx"))
    assert len(mods) == 6
    total = 0
    for m in mods:
        so = ctypes.CDLL(os.path.join(render.PLUGIN_DIR, m))
        so.OfxGetNumberOfPlugins.restype = ctypes.c_int
        n = so.OfxGetNumberOfPlugins()
        assert n >= 3 and hasattr(so, "OfxGetPlugin")
        total += n
    assert total == 24


@pytest.mark.parametrize("name", sorted(GOLD))
def test_graph_matches_golden_tree(host, name):
    g = GOLD[name]
    tl = TL.open_timeline(name)
    info = tl.info()
    assert info["frames"] == len(g["frames"])
    assert info["rate"] == g["rate"] and info["start"] == g["start"]
    assert [info["width"], info["height"]] == g["size"]
    for rec in g["frames"]:
        got = tl.describe(host, rec["time"]).rstrip("\n").split("\n")
        assert got == rec["tree"], f"{name} @ {rec['time']}"
    tl.close()


def test_unknown_effect_is_skipped_and_unknown_transition_dissolves(host):
    from toucan_b200 import render, workloads

    doc = workloads.stack_timeline(tracks=1, frames=96)
    track = doc["tracks"]["children"][0]
    track["children"][1]["transition_type"] = "SMPTE_Dissolve"
    track["children"][0]["effects"].insert(0, {"OTIO_SCHEMA": "Effect.1", "effect_name": "nonesuch:Effect", "metadata": {}})
    tl = render.Timeline(json_doc=doc, directory="/nonexistent")
    import numpy as np

    for k in range(2):
        tl.set_media(tl.media_path(workloads.still_url(0, k)), np.zeros((8, 8, 4), np.float32))
    tl.prepare()
    tree = tl.describe(host, 40.0)
    assert "toucan:Dissolve {value=0.16666666666666666}" in tree
    assert "nonesuch" not in tree and tree.count("toucan:Blur") == 2


def test_host_fit_matches_oracle():
    """toucan::fit of the C++ host against the oracle's restatement of Util.cpp:123-147 (the box decides which pixels
    a fit-resize writes, so it must agree to the pixel): config sizes, odd sizes, degenerate boxes, 2000 random pairs."""
    import numpy as np

    from oracle import oracle as O
    from toucan_b200 import render

    cases = [((1280, 720), (1280, 720)), ((3840, 2160), (1920, 1080)), ((1280, 720), (720, 1280)), ((128, 72), (64, 36)),
             ((1920, 1080), (1000, 1000)), ((7, 5), (3, 11)), ((1, 1), (1, 1)), ((100, 100), (1, 0)), ((100, 0), (10, 10)),
             ((33, 77), (1280, 720)), ((7680, 4320), (1281, 719))]
    rng = np.random.default_rng(4)
    cases += [((int(a), int(b)), (int(c), int(d))) for a, b, c, d in rng.integers(1, 5000, size=(2000, 4))]
    for a, b in cases:
        assert render.fit(a, b) == tuple(O.fit(a, b)), (a, b)


def test_plugins_resolve_parameters_through_the_host_suites(tmp_path):
    """tests/cpp/chain_probe.cpp: effect nodes created like ImageGraph creates them, then the Render action in
    pointwise-chain mode -- the plug-ins read their parameters through the host's parameter suite exactly as in a real
    render (double, string and boolean values, described and effective defaults: SURVEY 8b), no kernel launched."""
    import subprocess

    from toucan_b200 import build as B

    B.build_all()
    exe = str(tmp_path / "chain_probe")
    subprocess.run([B.CXX, "-O1", "-std=c++17", "-I", os.path.join(B.ROOT, "include"), "-I", B.HOST, os.path.join(B.ROOT, "tests", "cpp", "chain_probe.cpp"),
                    "-o", exe, "-L", B.LIB, "-ltoucan_render", "-ltoucan_cuda", f"-Wl,-rpath,{B.LIB}"], check=True, capture_output=True)
    out = subprocess.run([exe, os.path.join(B.LIB, "plugins")], check=True, capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("toucan:")]
    assert lines == [
        "toucan:Invert 1 1 0 0 0 [] []",
        "toucan:Pow 1 1 1 2.20000005 0 [] []",
        "toucan:Pow 1 1 1 1 0 [] []",                                   # described default of "value"
        "toucan:Saturate 1 1 2 0.25 0 [] []",
        "toucan:ColorMap 1 1 3 0 0 [plasma] []",
        "toucan:ColorConvert 1 1 6 0 1 [sRGB - Texture] [ACEScg]",
        "toucan:ColorConvert 1 1 6 0 0 [] []",                          # effective default of "premult" is 0
        "toucan:Premult 1 1 4 0 0 [] []",
        "toucan:Unpremult 1 1 5 0 0 [] []",
        "toucan:Blur 1 1 11 10 0 [] []",                                # a convolution node: may START a chain (TC_CHAIN_BLUR, radius)
        "toucan:Pow 0 0 -1 0 0 [] []",                                  # output-size override
    ], out


def test_timeline_documents_survive_damage(tmp_path):
    """tests/cpp/otio_fuzz.cpp under AddressSanitizer + UBSan: every fixture timeline (and a .otioz bundle) parsed whole,
    truncated and with random bytes flipped, plus documents nested 300 000 levels deep -- refused, never a crash."""
    import glob
    import subprocess

    from tests.test_ingest import make_otioz
    from toucan_b200 import build as B

    files = sorted(glob.glob(os.path.join(TL.DATA, "*.otio")))
    bundle = str(tmp_path / "t.otioz")
    make_otioz(bundle, "Transition.otio")
    deep = [str(tmp_path / "deep_array.otio"), str(tmp_path / "deep_object.otio")]
    with open(deep[0], "w") as f:
        f.write("[" * 300000)
    with open(deep[1], "w") as f:
        f.write('{"a":' * 200000)
    exe = str(tmp_path / "otio_fuzz")
    src = [os.path.join(B.ROOT, "tests", "cpp", "otio_fuzz.cpp"), os.path.join(B.HOST, "toucanRender", "otio.cpp"), os.path.join(B.HOST, "toucanRender", "Archive.cpp")]
    r = subprocess.run([B.CXX, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I", os.path.join(B.ROOT, "include"),
                        "-I", B.HOST] + src + ["-o", exe, "-lz", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and ("asan" in r.stderr.lower() or "ubsan" in r.stderr.lower() or "fsanitize" in r.stderr):
        pytest.skip("no sanitizer runtime for " + B.CXX)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "300", str(tmp_path / "mutated.otioz")] + files + [bundle] + deep, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "parsed" in r.stdout, (r.stdout[-300:], r.stderr[-3000:])
    # and through the public API: a hopeless document is an error, not a crash
    from toucan_b200 import capi, render

    with pytest.raises(capi.TcError):
        render.Timeline(json_doc=None, path=deep[0])



def _random_timeline(rng):
    """A random but well-formed .otio document over synthetic stills: 1-3 video tracks (plus an audio track the graph skips) of
    clips, gaps and transitions, with effect chains and time warps on clips, tracks and the stack."""
    rate = 24.0

    def rt(v):
        return {"OTIO_SCHEMA": "RationalTime.1", "rate": rate, "value": float(v)}

    def tr(s, d):
        return {"OTIO_SCHEMA": "TimeRange.1", "start_time": rt(s), "duration": rt(d)}

    def effect():
        k = int(rng.integers(0, 12))
        meta = [("toucan:Invert", {}), ("toucan:Pow", {"value": float(rng.choice([0.5, 2.0, 2.2, 3.0]))}),
                ("toucan:Saturate", {"value": float(rng.uniform(0, 2))}), ("toucan:Blur", {"radius": float(rng.choice([2.0, 10.0, 20.0]))}),
                ("toucan:ColorMap", {"map_name": str(rng.choice(["plasma", "inferno", "turbo"]))}), ("toucan:Flip", {}), ("toucan:Flop", {}),
                ("toucan:Crop", {"pos": [int(rng.integers(0, 4)), int(rng.integers(0, 3))], "size": [int(rng.integers(1, 8)), int(rng.integers(1, 6))]}),
                ("toucan:Resize", {"size": [int(rng.integers(2, 20)), int(rng.integers(2, 16))]}),
                ("toucan:Rotate", {"angle": float(rng.uniform(-90, 90))}), ("toucan:UnsharpMask", {"width": 3.0, "contrast": 0.5}),
                ("nonesuch:Effect", {})][k]
        return {"OTIO_SCHEMA": "Effect.1", "name": meta[0], "effect_name": meta[0], "metadata": meta[1]}

    def effects(p_warp):
        out = [effect() for _ in range(int(rng.integers(0, 4)))]
        if rng.random() < p_warp:
            out.insert(int(rng.integers(0, len(out) + 1)),
                       {"OTIO_SCHEMA": "LinearTimeWarp.1", "name": "", "effect_name": "", "metadata": {}, "time_scalar": float(rng.choice([0.5, 1.5, 2.0, -1.0]))})
        return out

    tracks = []
    for t in range(int(rng.integers(1, 4))):
        children = []
        n = int(rng.integers(1, 6))
        for i in range(n):
            dur = int(rng.integers(6, 20))
            if i > 0 and rng.random() < 0.6:
                o = int(rng.integers(1, 3))
                children.append({"OTIO_SCHEMA": "Transition.1", "name": "x", "metadata": {},
                                 "transition_type": str(rng.choice(["toucan:Dissolve", "toucan:HorizontalWipe", "toucan:VerticalWipe", "SMPTE_Dissolve"])),
                                 "in_offset": rt(o), "out_offset": rt(int(rng.integers(1, 3)))})
            if rng.random() < 0.15 and children and children[-1]["OTIO_SCHEMA"] != "Transition.1":
                children.append({"OTIO_SCHEMA": "Gap.1", "name": "Gap", "metadata": {}, "effects": [], "source_range": tr(0, int(rng.integers(2, 8)))})
            if rng.random() < 0.2:  # a generator clip (Clip.1 with a single media_reference, as data/Generator.otio has them)
                kind = str(rng.choice(["toucan:Fill", "toucan:Checkers", "toucan:Gradient", "toucan:Noise"]))
                params = {"size": [48, 30]}
                if kind == "toucan:Fill":
                    params["color"] = [float(v) for v in rng.random(4)]
                elif kind == "toucan:Checkers":
                    params.update(checkerSize=[int(rng.integers(1, 20)), int(rng.integers(1, 20))], color1=[float(v) for v in rng.random(4)], color2=[float(v) for v in rng.random(4)])
                elif kind == "toucan:Gradient":
                    params.update(color1=[float(v) for v in rng.random(4)], color2=[float(v) for v in rng.random(4)], vertical=bool(rng.integers(0, 2)))
                else:
                    params.update(type=str(rng.choice(["gaussian", "uniform", "salt"])), a=float(rng.random()), b=float(rng.random()), mono=bool(rng.integers(0, 2)), seed=int(rng.integers(0, 100)))
                children.append({"OTIO_SCHEMA": "Clip.1", "name": f"t{t}g{i}", "metadata": {}, "source_range": tr(0, dur), "effects": effects(0.1),
                                 "media_reference": {"OTIO_SCHEMA": "GeneratorReference.1", "name": "", "metadata": {}, "available_range": tr(0, 40),
                                                     "generator_kind": kind, "parameters": params}})
                continue
            if rng.random() < 0.15:  # a numbered image sequence: the frame number read follows the clip's (warped) time
                children.append({
                    "OTIO_SCHEMA": "Clip.2", "name": f"t{t}s{i}", "metadata": {}, "source_range": tr(int(rng.integers(0, 6)), dur), "effects": effects(0.4),
                    "media_references": {"DEFAULT_MEDIA": {"OTIO_SCHEMA": "ImageSequenceReference.1", "name": "", "metadata": {},
                                                           "available_range": tr(int(rng.integers(0, 3)), 40), "target_url_base": f"synthetic://seq{t}/",
                                                           "name_prefix": "f.", "name_suffix": ".png", "start_frame": int(rng.integers(0, 3)), "frame_step": 1,
                                                           "rate": rate, "frame_zero_padding": 4, "missing_frame_policy": "error"}},
                    "active_media_reference_key": "DEFAULT_MEDIA"})
                continue
            children.append({
                "OTIO_SCHEMA": "Clip.2", "name": f"t{t}c{i}", "metadata": {}, "source_range": tr(int(rng.integers(0, 6)), dur),
                "effects": effects(0.3),
                "media_references": {"DEFAULT_MEDIA": {"OTIO_SCHEMA": "ExternalReference.1", "name": "", "metadata": {},
                                                       "available_range": tr(0, 40) if rng.random() < 0.8 else
                                                       {"OTIO_SCHEMA": "TimeRange.1", "start_time": {"OTIO_SCHEMA": "RationalTime.1", "rate": 30.0, "value": 0.0},
                                                        "duration": {"OTIO_SCHEMA": "RationalTime.1", "rate": 30.0, "value": 50.0}},
                                                       "target_url": f"synthetic://m{t}_{i % 3}.png"}},
                "active_media_reference_key": "DEFAULT_MEDIA"})
        if rng.random() < 0.15:  # a transition that opens or closes the track (nothing on its far side)
            x = {"OTIO_SCHEMA": "Transition.1", "name": "edge", "metadata": {}, "transition_type": "toucan:Dissolve", "in_offset": rt(2), "out_offset": rt(2)}
            children.insert(0, x) if rng.random() < 0.5 else children.append(x)
        track = {"OTIO_SCHEMA": "Track.1", "name": f"V{t}", "kind": "Video", "metadata": {}, "effects": effects(0.15), "children": children}
        if rng.random() < 0.2:  # a trimmed track
            track["source_range"] = tr(int(rng.integers(0, 8)), int(rng.integers(5, 30)))
        tracks.append(track)
    tracks.insert(int(rng.integers(0, len(tracks) + 1)),
                  {"OTIO_SCHEMA": "Track.1", "name": "A1", "kind": "Audio", "metadata": {}, "effects": [], "children": []})
    return {"OTIO_SCHEMA": "Timeline.1", "name": "random", "metadata": {},
            "tracks": {"OTIO_SCHEMA": "Stack.1", "name": "tracks", "metadata": {}, "effects": effects(0.15), "children": tracks}}


def _random_timeline_media(doc):
    """URLs of every still and of frames -80 .. 199 of every numbered sequence of a _random_timeline document."""
    urls = set()
    for t in doc["tracks"]["children"]:
        for c in t["children"]:
            mr = (c.get("media_references") or {}).get("DEFAULT_MEDIA")
            if not mr:
                continue
            if mr["OTIO_SCHEMA"].startswith("ExternalReference"):
                urls.add(mr["target_url"])
            elif mr["OTIO_SCHEMA"].startswith("ImageSequenceReference"):
                for n in range(-80, 200):  # (time warps reach before the first frame: "f.00-3.png", Util.cpp:46-49)
                    urls.add("%s%s%s%s" % (mr["target_url_base"], mr["name_prefix"], str(n).rjust(mr["frame_zero_padding"], "0"), mr["name_suffix"]))
    return sorted(urls)


@pytest.mark.parametrize("seed", range(160))
def test_random_timelines_build_the_oracles_node_trees(host, seed):
    """ImageGraph::exec of the C++ host against the oracle's restatement of lib/toucanRender/ImageGraph.cpp:144-466 on
    random timelines: for every frame the same node tree (which clip, which transition and its progress, every effect's
    resolved parameters, time warps applied at clip / track / stack level, gaps, skipped audio tracks and unknown effects)."""
    import numpy as np

    from oracle import graph as G
    from toucan_b200 import render

    rng = np.random.default_rng(7000 + seed)
    doc = _random_timeline(rng)
    still = np.zeros((6, 8, 4), np.uint8)
    tl = render.Timeline(json_doc=doc, directory="/nonexistent")
    for u in _random_timeline_media(doc):
        tl.set_media(tl.media_path(u), still)
    tl.prepare()
    otl = G.Timeline(doc)
    og = G.ImageGraph(otl, lambda p: still)
    info = tl.info()
    times = otl.frames()
    assert info["frames"] == len(times)
    for t in times:
        want = G.describe_tree(og.exec(t))
        got = tl.describe(host, t.value).rstrip("\n").split("\n")
        assert got == want, f"seed {seed} @ {t.value}:\n" + "\n".join(got) + "\n--- oracle ---\n" + "\n".join(want)
    tl.close()
