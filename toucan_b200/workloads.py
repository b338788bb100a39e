"""Synthetic timelines of BASELINE.json `configs` as OpenTimelineIO JSON (host-side only).

The timelines are plain .otio documents, so the same description feeds the C++ host
(ImageGraph), the CPU oracle's graph restatement (oracle/graph.py) and bench.py.

C5 (configs[4]): 1000-frame, 10-track, 7680x4320 RGBA float32 @24 fps.  Every track is
a run of 48-frame clips joined by `toucan:Dissolve` transitions with in/out offset 12,
every clip carries `toucan:Blur` radius 10, clip media alternate between two stills per
track; tracks are composited with premult + over onto the black background
(lib/toucanRender/ImageGraph.cpp:144-215).
"""
from __future__ import annotations

C5 = dict(width=7680, height=4320, tracks=10, frames=1000, rate=24.0, clip_frames=48, offset=12, blur_radius=10.0)


def _rt(v, rate):
    return {"OTIO_SCHEMA": "RationalTime.1", "rate": float(rate), "value": float(v)}


def _tr(start, dur, rate):
    return {"OTIO_SCHEMA": "TimeRange.1", "start_time": _rt(start, rate), "duration": _rt(dur, rate)}


def still_url(track: int, k: int) -> str:
    """Media key of still k (0/1) of a track; bench/tests register decoded frames under it."""
    return f"synthetic://track{track}/still{k}.exr"


def stack_timeline(tracks=10, frames=1000, rate=24.0, clip_frames=48, offset=12, blur_radius=10.0, **_):
    """The C5 timeline structure as an OTIO JSON dict (size-independent)."""
    n_clips = (frames + clip_frames - 1) // clip_frames
    track_nodes = []
    for t in range(tracks):
        children = []
        for i in range(n_clips):
            dur = min(clip_frames, frames - i * clip_frames)
            if i > 0:
                children.append({
                    "OTIO_SCHEMA": "Transition.1", "name": f"t{t}x{i}", "metadata": {},
                    "transition_type": "toucan:Dissolve",
                    "in_offset": _rt(offset, rate), "out_offset": _rt(offset, rate),
                })
            children.append({
                "OTIO_SCHEMA": "Clip.2", "name": f"t{t}c{i}", "metadata": {},
                # stills: the same frame for any clip-local time, with handles either side
                "source_range": _tr(offset, dur, rate),
                "effects": [{"OTIO_SCHEMA": "Effect.1", "name": "blur", "effect_name": "toucan:Blur",
                             "metadata": {"radius": float(blur_radius)}}],
                "media_references": {"DEFAULT_MEDIA": {
                    "OTIO_SCHEMA": "ExternalReference.1", "name": "", "metadata": {},
                    "available_range": _tr(0, clip_frames + 2 * offset, rate),
                    "target_url": still_url(t, i % 2)}},
                "active_media_reference_key": "DEFAULT_MEDIA",
            })
        track_nodes.append({"OTIO_SCHEMA": "Track.1", "name": f"V{t + 1}", "kind": "Video", "metadata": {},
                            "effects": [], "children": children})
    return {"OTIO_SCHEMA": "Timeline.1", "name": "synthetic-stack", "metadata": {},
            "tracks": {"OTIO_SCHEMA": "Stack.1", "name": "tracks", "metadata": {}, "effects": [], "children": track_nodes}}


# SURVEY 8d config 2, chained variant: ONE clip carrying the six filters of data/Filter.otio in this order
FILTER_CHAIN = [("toucan:ColorMap", {"map_name": "plasma"}), ("toucan:Invert", {}), ("toucan:Pow", {"value": 2.0}),
                ("toucan:Saturate", {"value": 0.0}), ("toucan:Blur", {"radius": 10.0}),
                ("toucan:UnsharpMask", {"kernel": "gaussian", "width": 10.0, "contrast": 1.0, "threshold": 0.0})]


def filter_chain_timeline(frames=240, rate=24.0, effects=None, url="synthetic://chain/source.exr"):
    """A single-track timeline whose only clip (a still held for `frames` frames) carries `effects` (default: FILTER_CHAIN)."""
    fx = [{"OTIO_SCHEMA": "Effect.1", "name": n.split(":")[1], "effect_name": n, "metadata": dict(m)} for n, m in (effects or FILTER_CHAIN)]
    clip = {"OTIO_SCHEMA": "Clip.2", "name": "chain", "metadata": {}, "source_range": _tr(0, frames, rate), "effects": fx,
            "media_references": {"DEFAULT_MEDIA": {"OTIO_SCHEMA": "ExternalReference.1", "name": "", "metadata": {},
                                                   "available_range": _tr(0, frames, rate), "target_url": url}},
            "active_media_reference_key": "DEFAULT_MEDIA"}
    track = {"OTIO_SCHEMA": "Track.1", "name": "V1", "kind": "Video", "metadata": {}, "effects": [], "children": [clip]}
    return {"OTIO_SCHEMA": "Timeline.1", "name": "filter-chain", "metadata": {},
            "tracks": {"OTIO_SCHEMA": "Stack.1", "name": "tracks", "metadata": {}, "effects": [], "children": [track]}}


def stack_layers(frame: int, tracks=10, frames=1000, clip_frames=48, offset=12, blur_radius=10.0, **_):
    """What ImageGraph::_track yields for `frame` of stack_timeline(), bottom track first:
    a list of (still_a, still_b | None, radius_a, radius_b, value) with still = (track, k).

    Restates ImageGraph.cpp:217-332 for this timeline shape: the clip containing the frame,
    plus the neighbouring transition whose range [cut-offset, cut+offset) contains it;
    value = (t - transition.start) / transition.duration; `from` is the earlier clip.
    """
    n_clips = (frames + clip_frames - 1) // clip_frames
    i = min(frame // clip_frames, n_clips - 1)
    out = []
    for t in range(tracks):
        a, b, v = (t, i % 2), None, 0.0
        cut_prev, cut_next = i * clip_frames, (i + 1) * clip_frames
        if i > 0 and cut_prev - offset <= frame < cut_prev + offset:
            a, b = (t, (i - 1) % 2), (t, i % 2)
            v = (frame - (cut_prev - offset)) / (2.0 * offset)
        elif i < n_clips - 1 and cut_next - offset <= frame < cut_next + offset:
            a, b = (t, i % 2), (t, (i + 1) % 2)
            v = (frame - (cut_next - offset)) / (2.0 * offset)
        out.append((a, b, blur_radius, blur_radius, v))
    return out


def frame_block(rank: int, world: int, frames: int):
    """Contiguous block of frames owned by a rank (ceil(N/G) frames per GPU)."""
    per = (frames + world - 1) // world
    lo = min(rank * per, frames)
    return lo, min(lo + per, frames)
