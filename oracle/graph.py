"""Python restatement of toucan's per-frame graph builder and its evaluation on the CPU oracle.

TEST INFRASTRUCTURE ONLY (see oracle/toucan_oracle.c).  Parity: coarsely pinned (tests/test_reference_filmstrips.py).

Follows, in this order of authority:
  * lib/toucanRender/ImageGraph.cpp:42-466 (constructor, exec, _track, _item, _timeWarps, _effects)
  * lib/toucanRender/ImageEffect.cpp:77-153 (output allocation rules per OFX context)
  * lib/toucanRender/Comp.cpp:29-84
  * lib/toucanRender/TimelineWrapper.cpp:311-316 (time range), bin/toucan-render/App.cpp:222-264 (frame loop)
  * OpenTimelineIO 0.18 time/range algebra as restated in SURVEY 9.8 [V]
  * the effective-parameter-default quirks of SURVEY 8b / 9.10

Media are supplied by a callback ``media(path: str) -> ndarray(h,w,4) | None`` (decoded
RGBA frames; decode is outside the hot path).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Optional

import numpy as np

from . import oracle as O


# ----------------------------------------------------------------------------------------
# opentime [V: OpenTimelineIO 0.18 rationalTime.h / timeRange.h]
@dataclass(frozen=True)
class RT:
    value: float = 0.0
    rate: float = 1.0

    def rescaled_value(self, rate):
        return self.value if rate == self.rate else (self.value * rate) / self.rate

    def __sub__(self, o):
        if self.rate < o.rate:
            return RT(self.rescaled_value(o.rate) - o.value, o.rate)
        return RT(self.value - o.rescaled_value(self.rate), self.rate)

    def __add__(self, o):
        if self.rate < o.rate:
            return RT(self.rescaled_value(o.rate) + o.value, o.rate)
        return RT(self.value + o.rescaled_value(self.rate), self.rate)

    def secs(self):
        return self.value / self.rate

    def eq(self, o):
        return self.rescaled_value(o.rate) == o.value

    def to_frames(self):
        return int(self.value)

    def round(self):
        # std::round: half away from zero
        v = self.value
        return RT(math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5), self.rate)


@dataclass(frozen=True)
class TR:
    start: RT = RT()
    duration: RT = RT()

    def end_exclusive(self):
        return self.start + self.duration

    def end_inclusive(self):
        et = self.end_exclusive()
        if (et - RT(self.start.rescaled_value(self.duration.rate), self.duration.rate)).value > 1:
            if self.duration.value != math.floor(self.duration.value):
                return RT(math.floor(et.value), et.rate)
            return et - RT(1, self.duration.rate)
        return self.start

    def contains(self, t: RT):
        return self.start.secs() <= t.secs() and t.secs() < self.end_exclusive().secs()


def _media_ref(j):
    """Clip.1 `media_reference`, or Clip.2 `media_references[active_media_reference_key]`."""
    mr = j.get("media_reference")
    if mr is None and isinstance(j.get("media_references"), dict):
        mr = j["media_references"].get(j.get("active_media_reference_key", "DEFAULT_MEDIA"))
    return mr or {}


def _rt(j):
    return RT(float(j["value"]), float(j["rate"])) if j else None


def _tr(j):
    return TR(_rt(j["start_time"]), _rt(j["duration"])) if j else None


# ----------------------------------------------------------------------------------------
# A light OTIO object model (Timeline/Stack/Track/Clip/Gap/Transition/Effect/LinearTimeWarp)
@dataclass(eq=False)  # nodes are compared by identity: two siblings may carry identical JSON
class Node:
    schema: str
    j: dict
    parent: Optional["Node"] = None
    children: list = field(default_factory=list)
    effects: list = field(default_factory=list)

    @property
    def kind(self):
        return self.schema.split(".")[0]

    @property
    def source_range(self):
        return _tr(self.j.get("source_range"))

    def is_transition(self):
        return self.kind == "Transition"

    # Item::duration / Transition::duration
    def duration(self) -> RT:
        if self.is_transition():
            return _rt(self.j["in_offset"]) + _rt(self.j["out_offset"])
        return self.trimmed_range().duration

    def trimmed_range(self) -> TR:
        sr = self.source_range
        return sr if sr is not None else self.available_range()

    def available_range(self) -> TR:
        k = self.kind
        if k == "Clip":
            mr = _media_ref(self.j)
            ar = _tr(mr.get("available_range"))
            return ar if ar is not None else TR()
        if k == "Track":
            dur = None
            for c in self.children:
                if not c.is_transition():
                    dur = c.duration() if dur is None else dur + c.duration()
            if dur is None:
                dur = RT(0, 1)
            if self.children:
                if self.children[0].is_transition():
                    dur = dur + _rt(self.children[0].j["in_offset"])
                if self.children[-1].is_transition():
                    dur = dur + _rt(self.children[-1].j["out_offset"])
            return TR(RT(0, dur.rate), dur)
        if k == "Stack":
            dur = None
            for c in self.children:
                d = c.duration()
                if dur is None or d.secs() > dur.secs():
                    dur = d
            if dur is None:
                dur = RT(0, 1)
            return TR(RT(0, dur.rate), dur)
        return TR()  # Gap: Item::available_range is not implemented -> TimeRange()

    # Track::range_of_child_at_index
    def range_of_child(self, child: "Node") -> TR:
        if self.kind == "Stack":
            d = child.duration()
            return TR(RT(0, d.rate), d)
        idx = next(i for i, c in enumerate(self.children) if c is child)
        d = child.duration()
        start = RT(0, d.rate)
        for c in self.children[:idx]:
            if not c.is_transition():
                start = start + c.duration()
        if child.is_transition():
            start = start - _rt(child.j["in_offset"])
        return TR(start, d)

    def trimmed_range_of_child(self, child: "Node") -> Optional[TR]:
        r = self.range_of_child(child)
        sr = self.source_range
        if sr is None:
            return r
        if sr.start.secs() >= r.end_exclusive().secs() or sr.end_exclusive().secs() <= r.start.secs():
            return None
        s, e = r.start, r.end_exclusive()
        if s.secs() < sr.start.secs():
            s = sr.start
        if e.secs() > sr.end_exclusive().secs():
            e = sr.end_exclusive()
        return TR(s, e - s)

    def trimmed_range_in_parent(self) -> Optional[TR]:
        return self.parent.trimmed_range_of_child(self) if self.parent else None

    # Item::transformed_time(time, to_item) for self = track, to_item = a child item
    def transformed_time(self, t: RT, item: "Node") -> RT:
        r = t
        r = r + item.trimmed_range().start
        r = r - self.range_of_child(item).start
        return r


def _build(j, parent=None) -> Node:
    n = Node(j["OTIO_SCHEMA"], j, parent)
    for e in j.get("effects", []) or []:
        n.effects.append(Node(e["OTIO_SCHEMA"], e, n))
    for c in j.get("children", []) or []:
        n.children.append(_build(c, n))
    return n


class Timeline:
    def __init__(self, path_or_json):
        if isinstance(path_or_json, (str, os.PathLike)):
            with open(path_or_json) as f:
                self.j = json.load(f)
            self.dir = os.path.dirname(os.path.abspath(path_or_json))
        else:
            self.j = path_or_json
            self.dir = ""
        self.stack = _build(self.j["tracks"])
        gst = _rt(self.j.get("global_start_time"))
        dur = self.stack.duration()
        # TimelineWrapper.cpp:311-316
        self.time_range = TR(gst if gst is not None else RT(0.0, dur.rate), dur)

    def frames(self):
        """App.cpp:222-224."""
        t = self.time_range.start
        end = self.time_range.end_inclusive()
        inc = RT(1.0, self.time_range.duration.rate)
        out = []
        while t.secs() <= end.secs():
            out.append(t)
            t = t + inc
        return out

    def video_clips(self):
        out = []

        def rec(n):
            for c in n.children:
                if c.kind == "Clip":
                    out.append(c)
                elif c.children:
                    rec(c)

        for tr in self.stack.children:
            if tr.kind == "Track" and tr.j.get("kind", "Video") == "Video":
                rec(tr)
        return out


# ----------------------------------------------------------------------------------------
# effect parameter handling: the *effective* defaults of SURVEY 9.10 (described int/bool
# defaults never reach the plug-in; RGBA defaults deliver only element 0).
def _is_int(v):
    return isinstance(v, int) and not isinstance(v, bool)


def _rgba(meta, key, described0):
    local = [0.0, 0.0, 0.0, 0.0]
    v = meta.get(key)
    if v is None:
        local[0] = described0  # double default index 0 is delivered
        return local
    if isinstance(v, list) and v and all(isinstance(x, float) for x in v):
        for i, x in enumerate(v[:4]):
            local[i] = x
    return local


def _int2(meta, key):
    v = meta.get(key)
    if isinstance(v, list) and v and all(_is_int(x) for x in v):
        out = [0, 0]
        for i, x in enumerate(v[:2]):
            out[i] = x
        return out
    return [0, 0]


def _bool_as_int(meta, key, local):
    """`int x = local; paramGetValue(h, &x)`: a bool is written into the low byte, an integer into the whole int."""
    v = meta.get(key)
    if isinstance(v, bool):
        return (local & ~0xff) | int(v)
    if _is_int(v):
        return v & 0xffffffff
    return local


def _dbl(meta, key, described, local=None):
    v = meta.get(key)
    if v is None:
        return described
    return v if isinstance(v, float) else (local if local is not None else described)


def _str(meta, key, described):
    v = meta.get(key)
    return v if isinstance(v, str) else described


GENERATORS = {"toucan:Fill", "toucan:Checkers", "toucan:Gradient", "toucan:Noise"}
FILTERS = {"toucan:Blur", "toucan:ColorMap", "toucan:Invert", "toucan:Pow", "toucan:Saturate",
           "toucan:UnsharpMask", "toucan:Crop", "toucan:Flip", "toucan:Flop", "toucan:Resize",
           "toucan:Rotate", "toucan:ColorConvert", "toucan:Premult", "toucan:Unpremult",
           "toucan:Box", "toucan:Line", "toucan:Text"}
TRANSITIONS = {"toucan:Dissolve", "toucan:HorizontalWipe", "toucan:VerticalWipe"}
KNOWN = GENERATORS | FILTERS | TRANSITIONS


def _size_of(meta):
    """ImageEffect.cpp:84-88: anyToVec wants an AnyVector of two int64."""
    v = meta.get("size")
    if isinstance(v, list) and len(v) == 2 and all(_is_int(x) for x in v):
        return (v[0], v[1])
    return (0, 0)


# A graph node is a tuple tree: ("read", path) | ("effect", name, meta, [inputs]) | ("comp", [inputs])
class ImageGraph:
    def __init__(self, timeline: Timeline, media: Callable[[str], Optional[np.ndarray]]):
        self.tl = timeline
        self.media = media
        self.time_range = timeline.time_range
        self.size = (0, 0)
        self.channels = 0
        self.dtype = ""
        # ImageGraph.cpp:53-123
        for clip in timeline.video_clips():
            mr = _media_ref(clip.j)
            s = mr.get("OTIO_SCHEMA", "")
            if s.startswith("ExternalReference") or s.startswith("ImageSequenceReference"):
                path = self._first_path(mr)
                img = self.media(path) if path else None
                if img is not None and img.shape[1] > 0:
                    self.size = (img.shape[1], img.shape[0])
                    self.channels = 4
                    self.dtype = O.NAME_OF[O.tcode(img)]
                    break
            elif s.startswith("GeneratorReference"):
                p = mr.get("parameters", {})
                if "size" in p and p["size"] is not None:
                    sz = _size_of(p)
                    self.size = sz
                    self.channels = 4
                    self.dtype = "u8"
                    break

    def _media_path(self, url):
        p = url.split("://", 1)[1] if "://" in url else url
        if not os.path.isabs(p):
            p = os.path.join(self.tl.dir, p)
        return p

    def _seq_path(self, mr, frame):
        base = self._media_path(mr.get("target_url_base", ""))
        pad = int(mr.get("frame_zero_padding", 0))
        # std::setw(padding) << std::setfill('0') << frame (Util.cpp:46-49): fill characters go LEFT of the sign ("00-1", not "-001")
        return os.path.join(base, "%s%s%s" % (mr.get("name_prefix", ""), str(frame).rjust(pad, "0"), mr.get("name_suffix", "")))

    def _first_path(self, mr):
        if mr["OTIO_SCHEMA"].startswith("ExternalReference"):
            return self._media_path(mr.get("target_url", ""))
        return self._seq_path(mr, int(mr.get("start_frame", 0)))

    # ------------------------------------------------------------------ graph construction
    def exec(self, time: RT):
        """ImageGraph::exec -- ImageGraph.cpp:144-215.  Returns the node tree."""
        node = ("effect", "toucan:Fill", {"size": [self.size[0], self.size[1]], "color": [0.0, 0.0, 0.0, 1.0]}, [])
        stack = self.tl.stack
        t = time - self.time_range.start
        t = self._time_warps(t, stack.available_range(), stack.effects)
        for track in stack.children:
            if track.kind != "Track" or track.j.get("kind", "Video") != "Video":
                continue
            if not any(c.kind == "Clip" for c in track.children):
                continue
            t2 = t
            if track.effects:
                t2 = self._time_warps(t2, track.available_range(), track.effects)
            tn = self._track(t2, track)
            tn = self._effects(track.effects, tn)
            nodes = [n for n in (tn, node) if n is not None]
            node = ("comp", nodes)
        node = self._effects(stack.effects, node)
        return node

    def _track(self, time: RT, track: Node):
        out = None
        item = None
        prev = prev2 = nxt = nxt2 = None
        ch = track.children
        for i, c in enumerate(ch):
            if c.is_transition():
                item = None
                continue
            item = c
            r = c.trimmed_range_in_parent()
            if r is not None and r.contains(time):
                out = self._item(track.transformed_time(time, c), c)
                if i > 0:
                    prev = ch[i - 1]
                if i > 1:
                    prev2 = ch[i - 2]
                if i < len(ch) - 1:
                    nxt = ch[i + 1]
                if len(ch) > 1 and i < len(ch) - 2:
                    nxt2 = ch[i + 2]
                break
        if item is not None and out is not None:
            if prev is not None and prev.is_transition():
                r = prev.trimmed_range_in_parent()
                if r is not None and r.contains(time) and prev2 is not None and not prev2.is_transition():
                    value = (time - r.start).value / r.duration.value
                    a = self._item(track.transformed_time(time, prev2), prev2)
                    out = self._transition(prev, value, [a, out])
            if nxt is not None and nxt.is_transition():
                r = nxt.trimmed_range_in_parent()
                if r is not None and r.contains(time) and nxt2 is not None and not nxt2.is_transition():
                    value = (time - r.start).value / r.duration.value
                    b = self._item(track.transformed_time(time, nxt2), nxt2)
                    out = self._transition(nxt, value, [out, b])
        return out

    def _transition(self, tr: Node, value: float, inputs):
        meta = dict(tr.j.get("metadata", {}) or {})
        meta["value"] = float(value)
        name = tr.j.get("transition_type", "")
        if name not in KNOWN:
            name = "toucan:Dissolve"
        return ("effect", name, meta, inputs)

    def _item(self, time: RT, item: Node):
        out = None
        t = self._time_warps(time, item.available_range(), item.effects)
        if item.kind == "Clip":
            mr = _media_ref(item.j)
            s = mr.get("OTIO_SCHEMA", "")
            if s.startswith("ExternalReference"):
                path = self._media_path(mr.get("target_url", ""))
                out = ("read", path) if self.media(path) is not None else None
            elif s.startswith("ImageSequenceReference"):
                # the reader exists if its first frame could be probed (Read.cpp:138-150 never throws)
                out = ("read", self._seq_path(mr, t.to_frames()))
            elif s.startswith("GeneratorReference"):
                kind = mr.get("generator_kind", "")
                if kind in KNOWN:
                    out = ("effect", kind, dict(mr.get("parameters", {}) or {}), [])
        elif item.kind == "Gap":
            out = ("effect", "toucan:Fill", {"size": [self.size[0], self.size[1]]}, [])
        return self._effects(item.effects, out)

    def _time_warps(self, time: RT, rng: TR, effects):
        out = time
        for e in effects:
            if e.kind == "LinearTimeWarp":
                s = float(e.j.get("time_scalar", 1.0))
                out = RT((out - rng.start).value * s, time.rate).round()
        return out

    def _effects(self, effects, inp):
        out = inp
        for e in effects:
            name = e.j.get("effect_name", "")
            if name in KNOWN:
                out = ("effect", name, dict(e.j.get("metadata", {}) or {}), [out])
        return out

    # ------------------------------------------------------------------ evaluation
    def render(self, time: RT) -> Optional[np.ndarray]:
        return self.eval(self.exec(time))

    def eval(self, node) -> Optional[np.ndarray]:
        if node is None:
            return None
        if node[0] == "read":
            img = self.media(node[1])
            return None if img is None else np.ascontiguousarray(img)
        if node[0] == "comp":
            ins = node[1]
            if len(ins) > 1:
                fg = self.eval(ins[0])
                bg = self.eval(ins[1])
                if fg is None:
                    fg = np.zeros((0, 0, 4), np.uint8)
                return O.comp(fg, bg, True)
            if len(ins) == 1:
                return O.comp(self.eval(ins[0]), None, True)
            return None
        _, name, meta, inputs = node
        return self.eval_effect(name, meta, [self.eval(i) for i in inputs])

    def eval_effect(self, name, meta, ins):
        """ImageEffectNode::exec (ImageEffect.cpp:77-153) + the plug-in _render bodies."""
        size = _size_of(meta)
        empty = np.zeros((0, 0, 4), np.uint8)
        if name in GENERATORS:
            w, h = size
            if w <= 0 or h <= 0:
                return empty
            if name == "toucan:Fill":
                return O.fill(w, h, _rgba(meta, "color", 0.0))
            if name == "toucan:Checkers":
                cs = _int2(meta, "checkerSize")
                return O.checkers(w, h, cs[0], cs[1], _rgba(meta, "color1", 0.0), _rgba(meta, "color2", 1.0))
            if name == "toucan:Gradient":
                v = meta.get("vertical")
                vertical = v if isinstance(v, bool) else False
                return O.gradient(w, h, _rgba(meta, "color1", 0.0), _rgba(meta, "color2", 1.0), vertical)
            if name == "toucan:Noise":
                mono = meta.get("mono")
                seed = meta.get("seed")
                # `int mono` receives a bool through a bool* write of its low byte (GeneratorPlugin.cpp:517-522)
                return O.noise(w, h, _str(meta, "type", "gaussian"), _dbl(meta, "a", 0.0), _dbl(meta, "b", 0.0),
                               bool(mono) if isinstance(mono, bool) else False, seed if _is_int(seed) else 0)
        if name in FILTERS:
            src = ins[0] if ins else None
            if src is None or src.size == 0:
                return empty
            sh, sw = src.shape[:2]
            ow, oh = (size if size[0] > 0 and size[1] > 0 else (sw, sh))
            if name == "toucan:Crop":
                pos = _int2(meta, "pos")
                csz = _int2(meta, "size")
                out = O.new(ow, oh, src.dtype)
                O.lib().o_crop(O.ctypes.byref(O._img(out)), O.ctypes.byref(O._img(src)), pos[0], pos[1], csz[0], csz[1])
                return out
            if name == "toucan:Resize":
                rs = _int2(meta, "size")
                fn = _str(meta, "filter_name", "")
                fw = _dbl(meta, "filter_width", 0.0)
                return O.resize(src, (ow, oh), fn, fw) if rs[0] > 0 else O.new(ow, oh, src.dtype)
            # all remaining filters render into an output of the (possibly overridden) size;
            # reads outside the source are black, so evaluate on a canvas of the output size.
            if (ow, oh) != (sw, sh):
                canvas = O.convert(src, src.dtype, (ow, oh))
            else:
                canvas = src
            if name == "toucan:Blur":
                return O.blur(canvas, _dbl(meta, "radius", 10.0)) if (ow, oh) == (sw, sh) else _blur_sized(src, ow, oh, _dbl(meta, "radius", 10.0))
            if name == "toucan:ColorMap":
                return O.color_map(canvas, _str(meta, "map_name", "plasma"))
            if name == "toucan:Invert":
                return O.invert(canvas)
            if name == "toucan:Pow":
                return O.pow_(canvas, _dbl(meta, "value", 1.0))
            if name == "toucan:Saturate":
                return O.saturate(canvas, _dbl(meta, "value", 1.0))
            if name == "toucan:UnsharpMask":
                return O.unsharp_mask(canvas, _str(meta, "kernel", "gaussian"), _dbl(meta, "width", 1.0, 3.0),
                                      _dbl(meta, "contrast", 1.0), _dbl(meta, "threshold", 1.0, 0.0))
            if name == "toucan:Flip":
                return O.flip(canvas)
            if name == "toucan:Flop":
                return O.flop(canvas)
            if name == "toucan:Rotate":
                return O.rotate(canvas, _dbl(meta, "angle", 0.0), _str(meta, "filter_name", ""), _dbl(meta, "filter_width", 0.0))
            if name == "toucan:ColorConvert":
                pm = meta.get("premult")
                return O.colorconvert(canvas, _str(meta, "fromspace", ""), _str(meta, "tospace", ""),
                                      unpremult=bool(pm) if isinstance(pm, bool) else False)
            if name == "toucan:Box":
                return O.draw_box(canvas, _int2(meta, "pos1"), _int2(meta, "pos2"), _rgba(meta, "color", 1.0), _bool_as_int(meta, "fill", 0))
            if name == "toucan:Line":
                return O.draw_line(canvas, _int2(meta, "pos1"), _int2(meta, "pos2"), _rgba(meta, "color", 1.0),
                                   _bool_as_int(meta, "skip_first_point", 0))
            if name == "toucan:Text":
                # render_text needs FreeType and a font file; without a font it draws nothing and the plug-in
                # ignores the failure (DrawPlugin.cpp:436-452): the output is the copy of the source
                return np.ascontiguousarray(canvas.copy())
            if name == "toucan:Premult":
                return O.premult(canvas)
            if name == "toucan:Unpremult":
                return O.unpremult(canvas)
        if name in TRANSITIONS:
            if len(ins) < 2 or ins[0] is None or ins[1] is None:
                return empty
            a, b = ins
            if a.size == 0:
                return empty
            ow, oh = (size if size[0] > 0 and size[1] > 0 else (a.shape[1], a.shape[0]))
            value = _dbl(meta, "value", 0.0)
            if name == "toucan:Dissolve":
                return O.dissolve(a, b, value, (ow, oh))
            return O.wipe(a, b, value, name == "toucan:VerticalWipe", (ow, oh))
        return None


def _blur_sized(src, ow, oh, radius):
    out = O.new(ow, oh, src.dtype)
    O.lib().o_blur(O.ctypes.byref(O._img(out)), O.ctypes.byref(O._img(src)), O.ctypes.c_float(radius))
    return out


# ----------------------------------------------------------------------------------------
# Text form of a node tree, identical to toucan::IImageNode::describeTree of the C++ host
# (one node per line, inputs indented by two spaces) -- lets the graph builders be compared
# without a GPU.
def _fmt_any(v):
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        s = "%.17g" % v
        if math.isfinite(v) and abs(v) < 9.2e18 and v == int(v):
            s += ".0"
        return s
    if isinstance(v, str):
        return '"' + v + '"'
    if isinstance(v, list):
        return "[" + ",".join(_fmt_any(x) for x in v) + "]"
    return "?"


def describe_tree(node, depth=0, out=None):
    out = [] if out is None else out
    pad = "  " * depth
    if node is None:
        out.append(pad + "(null)")
        return out
    if node[0] == "read":
        out.append(pad + "Read: " + os.path.basename(node[1]))
        return out
    if node[0] == "comp":
        out.append(pad + "Comp")
        for i in node[1]:
            describe_tree(i, depth + 1, out)
        return out
    _, name, meta, inputs = node
    out.append(pad + name + " {" + ", ".join(k + "=" + _fmt_any(meta[k]) for k in sorted(meta)) + "}")
    for i in inputs:
        describe_tree(i, depth + 1, out)
    return out
