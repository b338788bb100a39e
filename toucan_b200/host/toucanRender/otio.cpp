// otio.cpp -- JSON reader and time/range algebra for the minimal OTIO object model
// (opentimelineio/otio.h).  Behaviour follows OpenTimelineIO 0.18 as restated in
// Track::range_of_child_at_index, Composition::trim_child_range,
// Item::transformed_time, Track/Stack::available_range.
#include <opentimelineio/otio.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace opentimelineio { inline namespace v0_18 {

// ---- JSON ------------------------------------------------------------------------------
namespace {

struct Parser {
    const char* p;
    const char* end;
    std::string err;
    int depth = 0; // containers open at the cursor: bounded, because the parser recurses once per level
    static constexpr int maxDepth = 512;

    void ws()
    {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    }
    bool fail(const std::string& m)
    {
        if (err.empty()) err = m;
        return false;
    }
    bool value(std::any& out)
    {
        ws();
        if (p >= end) return fail("unexpected end of input");
        switch (*p) {
        case '{':
        case '[': {
            if (depth >= maxDepth) return fail("nesting deeper than 512 levels");
            ++depth;
            const bool ok = '{' == *p ? object(out) : array(out);
            --depth;
            return ok;
        }
        case '"': {
            std::string s;
            if (!string(s)) return false;
            out = s;
            return true;
        }
        case 't':
            if (end - p >= 4 && !strncmp(p, "true", 4)) { p += 4; out = true; return true; }
            return fail("bad literal");
        case 'f':
            if (end - p >= 5 && !strncmp(p, "false", 5)) { p += 5; out = false; return true; }
            return fail("bad literal");
        case 'n':
            if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; out = std::any(); return true; }
            return fail("bad literal");
        default: return number(out);
        }
    }
    bool number(std::any& out)
    {
        const char* s = p;
        bool is_int = true;
        if (p < end && (*p == '-' || *p == '+')) ++p;
        while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) {
            if (*p == '.' || *p == 'e' || *p == 'E') is_int = false;
            ++p;
        }
        if (p == s) {
            // NaN / Infinity spellings some OTIO writers emit
            if (end - p >= 3 && !strncmp(p, "NaN", 3)) { p += 3; out = std::nan(""); return true; }
            if (end - p >= 8 && !strncmp(p, "Infinity", 8)) { p += 8; out = HUGE_VAL; return true; }
            return fail(std::string("unexpected character '") + *p + "'");
        }
        const std::string tok(s, p);
        if (is_int) out = static_cast<int64_t>(std::strtoll(tok.c_str(), nullptr, 10));
        else out = std::strtod(tok.c_str(), nullptr);
        return true;
    }
    bool string(std::string& out)
    {
        ++p; // opening quote
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) {
                case 'n': out.push_back('\n'); break;
                case 't': out.push_back('\t'); break;
                case 'r': out.push_back('\r'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'u': {
                    if (end - p < 5) return fail("bad \\u escape");
                    unsigned cp = (unsigned)std::strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                    p += 4;
                    if (cp < 0x80) out.push_back((char)cp);
                    else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                    else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                    break;
                }
                default: out.push_back(*p); break;
                }
                ++p;
            } else
                out.push_back(*p++);
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    bool array(std::any& out)
    {
        ++p;
        AnyVector v;
        ws();
        if (p < end && *p == ']') { ++p; out = v; return true; }
        for (;;) {
            std::any e;
            if (!value(e)) return false;
            v.push_back(std::move(e));
            ws();
            if (p < end && *p == ',') { ++p; continue; }
            if (p < end && *p == ']') { ++p; break; }
            return fail("expected ',' or ']'");
        }
        out = std::move(v);
        return true;
    }
    bool object(std::any& out)
    {
        ++p;
        AnyDictionary d;
        ws();
        if (p < end && *p == '}') { ++p; out = d; return true; }
        for (;;) {
            ws();
            if (p >= end || *p != '"') return fail("expected a string key");
            std::string k;
            if (!string(k)) return false;
            ws();
            if (p >= end || *p != ':') return fail("expected ':'");
            ++p;
            std::any e;
            if (!value(e)) return false;
            d[k] = std::move(e);
            ws();
            if (p < end && *p == ',') { ++p; continue; }
            if (p < end && *p == '}') { ++p; break; }
            return fail("expected ',' or '}'");
        }
        out = std::move(d);
        return true;
    }
};

const AnyDictionary* as_dict(const std::any& a) { return std::any_cast<AnyDictionary>(&a); }
const AnyVector* as_vec(const std::any& a) { return std::any_cast<AnyVector>(&a); }
const std::any* find(const AnyDictionary& d, const char* k)
{
    auto i = d.find(k);
    return i == d.end() ? nullptr : &i->second;
}
std::string get_str(const AnyDictionary& d, const char* k, const std::string& def = std::string())
{
    const std::any* a = find(d, k);
    if (a)
        if (auto s = std::any_cast<std::string>(a)) return *s;
    return def;
}
double get_num(const AnyDictionary& d, const char* k, double def = 0.0)
{
    const std::any* a = find(d, k);
    if (!a) return def;
    if (auto v = std::any_cast<double>(a)) return *v;
    if (auto v = std::any_cast<int64_t>(a)) return (double)*v;
    return def;
}
std::string schema_of(const AnyDictionary& d)
{
    std::string s = get_str(d, "OTIO_SCHEMA");
    const size_t dot = s.find('.');
    return dot == std::string::npos ? s : s.substr(0, dot);
}
std::optional<RationalTime> read_rt(const std::any* a)
{
    if (!a) return std::nullopt;
    const AnyDictionary* d = as_dict(*a);
    if (!d) return std::nullopt;
    return RationalTime(get_num(*d, "value"), get_num(*d, "rate", 1.0));
}
std::optional<TimeRange> read_tr(const std::any* a)
{
    if (!a) return std::nullopt;
    const AnyDictionary* d = as_dict(*a);
    if (!d) return std::nullopt;
    auto s = read_rt(find(*d, "start_time"));
    auto du = read_rt(find(*d, "duration"));
    if (!s || !du) return std::nullopt;
    return TimeRange(*s, *du);
}
void read_common(SerializableObjectWithMetadata* o, const AnyDictionary& d)
{
    o->set_name(get_str(d, "name"));
    if (const std::any* m = find(d, "metadata"))
        if (auto md = as_dict(*m)) o->metadata() = *md;
}

std::shared_ptr<SerializableObject> build(const std::any& a, ErrorStatus* es);

void read_effects(Item* item, const AnyDictionary& d, ErrorStatus* es)
{
    if (const std::any* e = find(d, "effects"))
        if (auto v = as_vec(*e))
            for (const auto& x : *v)
                if (auto eff = std::dynamic_pointer_cast<Effect>(build(x, es))) item->effects().push_back(SerializableObject::Retainer<Effect>(eff));
}
void read_children(Composition* c, const AnyDictionary& d, ErrorStatus* es)
{
    if (const std::any* ch = find(d, "children"))
        if (auto v = as_vec(*ch))
            for (const auto& x : *v)
                if (auto child = std::dynamic_pointer_cast<Composable>(build(x, es))) c->append_child(SerializableObject::Retainer<Composable>(child));
}

std::shared_ptr<SerializableObject> build(const std::any& a, ErrorStatus* es)
{
    const AnyDictionary* dp = as_dict(a);
    if (!dp) return nullptr;
    const AnyDictionary& d = *dp;
    const std::string schema = schema_of(d);
    if (schema == "Timeline") {
        auto t = std::make_shared<Timeline>();
        read_common(t.get(), d);
        t->set_global_start_time(read_rt(find(d, "global_start_time")));
        if (const std::any* tr = find(d, "tracks"))
            if (auto st = std::dynamic_pointer_cast<Stack>(build(*tr, es))) t->set_tracks(SerializableObject::Retainer<Stack>(st));
        return t;
    }
    if (schema == "Stack" || schema == "Track") {
        std::shared_ptr<Composition> c;
        if (schema == "Stack") c = std::make_shared<Stack>();
        else {
            auto tr = std::make_shared<Track>();
            tr->set_kind(get_str(d, "kind", Track::Kind::video));
            c = tr;
        }
        read_common(c.get(), d);
        c->set_source_range(read_tr(find(d, "source_range")));
        read_effects(c.get(), d, es);
        read_children(c.get(), d, es);
        return c;
    }
    if (schema == "Clip") {
        auto c = std::make_shared<Clip>();
        read_common(c.get(), d);
        c->set_source_range(read_tr(find(d, "source_range")));
        read_effects(c.get(), d, es);
        const std::any* mr = find(d, "media_reference");
        if (!mr || !mr->has_value()) {
            // Clip.2: media_references map + active key
            if (const std::any* mrs = find(d, "media_references"))
                if (auto m = as_dict(*mrs)) {
                    auto i = m->find(get_str(d, "active_media_reference_key", "DEFAULT_MEDIA"));
                    if (i != m->end()) mr = &i->second;
                }
        }
        std::shared_ptr<MediaReference> ref;
        if (mr) ref = std::dynamic_pointer_cast<MediaReference>(build(*mr, es));
        if (!ref) ref = std::make_shared<MissingReference>();
        c->set_media_reference(SerializableObject::Retainer<MediaReference>(ref));
        return c;
    }
    if (schema == "Gap") {
        auto g = std::make_shared<Gap>();
        read_common(g.get(), d);
        auto sr = read_tr(find(d, "source_range"));
        g->set_source_range(sr ? sr : std::optional<TimeRange>(TimeRange()));
        read_effects(g.get(), d, es);
        return g;
    }
    if (schema == "Transition") {
        auto t = std::make_shared<Transition>();
        read_common(t.get(), d);
        t->set_transition_type(get_str(d, "transition_type"));
        if (auto v = read_rt(find(d, "in_offset"))) t->set_in_offset(*v);
        if (auto v = read_rt(find(d, "out_offset"))) t->set_out_offset(*v);
        return t;
    }
    if (schema == "LinearTimeWarp") {
        auto e = std::make_shared<LinearTimeWarp>();
        read_common(e.get(), d);
        e->set_effect_name(get_str(d, "effect_name"));
        e->set_time_scalar(get_num(d, "time_scalar", 1.0));
        return e;
    }
    if (schema == "Effect" || schema == "TimeEffect" || schema == "FreezeFrame") {
        auto e = std::make_shared<Effect>();
        read_common(e.get(), d);
        e->set_effect_name(get_str(d, "effect_name"));
        return e;
    }
    if (schema == "ExternalReference") {
        auto r = std::make_shared<ExternalReference>(get_str(d, "target_url"));
        read_common(r.get(), d);
        r->set_available_range(read_tr(find(d, "available_range")));
        return r;
    }
    if (schema == "ImageSequenceReference") {
        auto r = std::make_shared<ImageSequenceReference>(get_str(d, "target_url_base"), get_str(d, "name_prefix"),
                                                          get_str(d, "name_suffix"), (int)get_num(d, "start_frame", 1),
                                                          (int)get_num(d, "frame_step", 1), get_num(d, "rate", 1.0),
                                                          (int)get_num(d, "frame_zero_padding", 0));
        read_common(r.get(), d);
        r->set_available_range(read_tr(find(d, "available_range")));
        return r;
    }
    if (schema == "GeneratorReference") {
        auto r = std::make_shared<GeneratorReference>();
        read_common(r.get(), d);
        r->set_available_range(read_tr(find(d, "available_range")));
        r->set_generator_kind(get_str(d, "generator_kind"));
        if (const std::any* p = find(d, "parameters"))
            if (auto pd = as_dict(*p)) r->parameters() = *pd;
        return r;
    }
    if (schema == "MissingReference") {
        auto r = std::make_shared<MissingReference>();
        read_common(r.get(), d);
        r->set_available_range(read_tr(find(d, "available_range")));
        return r;
    }
    // markers and unknown schemas are not on the render path
    return nullptr;
}

} // namespace

std::any parse_json(const std::string& text, ErrorStatus* es)
{
    Parser ps { text.data(), text.data() + text.size(), {} };
    std::any out;
    if (!ps.value(out)) {
        if (es) {
            es->outcome = ErrorStatus::JSON_PARSE_ERROR;
            es->details = ps.err;
            es->full_description = "JSON parse error: " + ps.err;
        }
        return std::any();
    }
    return out;
}

Timeline::Timeline() : _tracks(std::make_shared<Stack>()) {}

SerializableObject* Timeline::from_json_string(const std::string& text, ErrorStatus* es)
{
    const std::any root = parse_json(text, es);
    if (!root.has_value()) return nullptr;
    auto obj = build(root, es);
    if (!obj) {
        if (es) {
            es->outcome = ErrorStatus::SCHEMA_ERROR;
            es->full_description = "the JSON root is not a known OTIO schema";
        }
        return nullptr;
    }
    // hand ownership to the caller's Retainer: park one reference in the object itself
    // until a Retainer adopts it (shared_from_this keeps it alive afterwards)
    auto* raw = obj.get();
    static thread_local std::vector<std::shared_ptr<SerializableObject>> parked;
    parked.clear();
    parked.push_back(obj);
    return raw;
}

SerializableObject* Timeline::from_json_file(const std::string& file_name, ErrorStatus* es)
{
    std::ifstream f(file_name, std::ios::binary);
    if (!f) {
        if (es) {
            es->outcome = ErrorStatus::FILE_OPEN_FAILED;
            es->details = file_name;
            es->full_description = "failed to open file for reading: " + file_name;
        }
        return nullptr;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    return from_json_string(ss.str(), es);
}

// ---- time / range algebra -----------------------------------------------------------------
TimeRange Item::available_range(ErrorStatus* es) const
{
    if (es) {
        es->outcome = ErrorStatus::NOT_IMPLEMENTED;
        es->full_description = "available_range is not implemented for this item";
    }
    return TimeRange();
}

TimeRange Clip::available_range(ErrorStatus* es) const
{
    MediaReference* ref = media_reference();
    if (!ref || !ref->available_range()) {
        if (es) {
            es->outcome = ErrorStatus::CANNOT_COMPUTE_AVAILABLE_RANGE;
            es->full_description = "No available_range set on media reference on clip";
        }
        return TimeRange();
    }
    return *ref->available_range();
}

std::optional<TimeRange> Item::trimmed_range_in_parent(ErrorStatus* es) const
{
    if (!parent()) return std::nullopt;
    return parent()->trimmed_range_of_child(this, es);
}
TimeRange Item::range_in_parent(ErrorStatus* es) const
{
    if (!parent()) return TimeRange();
    return parent()->range_of_child(this, es);
}
std::optional<TimeRange> Transition::trimmed_range_in_parent(ErrorStatus* es) const
{
    if (!parent()) return std::nullopt;
    return parent()->trimmed_range_of_child(this, es);
}

RationalTime Item::transformed_time(RationalTime time, const Item* to_item, ErrorStatus* es) const
{
    if (!to_item) return time;
    const Item* root = this;
    while (root->parent()) root = root->parent();
    const Item* item = this;
    RationalTime result = time;
    while (item != root && item != to_item) {
        const Composition* par = item->parent();
        result -= item->trimmed_range(es).start_time();
        result += par->range_of_child(item, es).start_time();
        item = par;
    }
    const Item* ancestor = item;
    item = to_item;
    while (item != root && item != ancestor) {
        const Composition* par = item->parent();
        result += item->trimmed_range(es).start_time();
        result -= par->range_of_child(item, es).start_time();
        item = par;
    }
    return result;
}

int Composition::index_of_child(const Composable* child) const
{
    for (size_t i = 0; i < _children.size(); ++i)
        if (_children[i].value == child) return (int)i;
    return -1;
}
TimeRange Composition::range_of_child(const Composable* child, ErrorStatus* es) const
{
    const int i = index_of_child(child);
    if (i < 0) return TimeRange();
    return range_of_child_at_index(i, es);
}
std::optional<TimeRange> Composition::trim_child_range(TimeRange child_range) const
{
    if (!source_range()) return child_range;
    const TimeRange sr = *source_range();
    const bool past_end = sr.start_time() >= child_range.end_time_exclusive();
    const bool before_start = sr.end_time_exclusive() <= child_range.start_time();
    if (past_end || before_start) return std::nullopt;
    if (child_range.start_time() < sr.start_time())
        child_range = TimeRange::range_from_start_end_time(sr.start_time(), child_range.end_time_exclusive());
    if (child_range.end_time_exclusive() > sr.end_time_exclusive())
        child_range = TimeRange::range_from_start_end_time(child_range.start_time(), sr.end_time_exclusive());
    return child_range;
}
std::optional<TimeRange> Composition::trimmed_range_of_child(const Composable* child, ErrorStatus* es) const
{
    return trim_child_range(range_of_child(child, es));
}
std::vector<SerializableObject::Retainer<Clip>> Composition::find_clips(ErrorStatus* es, std::optional<TimeRange>, bool shallow) const
{
    std::vector<Retainer<Clip>> out;
    for (const auto& c : _children) {
        if (auto clip = dynamic_cast<Clip*>(c.value)) out.push_back(Retainer<Clip>(clip));
        else if (!shallow)
            if (auto comp = dynamic_cast<Composition*>(c.value)) {
                auto sub = comp->find_clips(es);
                out.insert(out.end(), sub.begin(), sub.end());
            }
    }
    return out;
}

TimeRange Track::range_of_child_at_index(int index, ErrorStatus* es) const
{
    const Composable* child = _children[index].value;
    const RationalTime child_duration = child->duration(es);
    RationalTime start_time(0, child_duration.rate());
    for (int i = 0; i < index; ++i) {
        const Composable* c2 = _children[i].value;
        if (!c2->overlapping()) start_time += c2->duration(es);
    }
    if (child->overlapping()) start_time -= static_cast<const Transition*>(child)->in_offset();
    return TimeRange(start_time, child_duration);
}
TimeRange Track::available_range(ErrorStatus* es) const
{
    RationalTime duration;
    bool first = true;
    for (const auto& c : _children) {
        if (auto item = dynamic_cast<const Item*>(c.value)) {
            duration = first ? item->duration(es) : duration + item->duration(es);
            first = false;
        }
    }
    if (!_children.empty()) {
        if (auto t = dynamic_cast<const Transition*>(_children.front().value)) duration += t->in_offset();
        if (auto t = dynamic_cast<const Transition*>(_children.back().value)) duration += t->out_offset();
    }
    return TimeRange(RationalTime(0, duration.rate()), duration);
}
TimeRange Stack::range_of_child_at_index(int index, ErrorStatus* es) const
{
    const RationalTime d = _children[index].value->duration(es);
    return TimeRange(RationalTime(0, d.rate()), d);
}
TimeRange Stack::available_range(ErrorStatus* es) const
{
    if (_children.empty()) return TimeRange();
    RationalTime duration = _children[0].value->duration(es);
    for (size_t i = 1; i < _children.size(); ++i) {
        const RationalTime d = _children[i].value->duration(es);
        if (d > duration) duration = d;
    }
    return TimeRange(RationalTime(0, duration.rate()), duration);
}

}} // namespace opentimelineio::v0_18
