// Minimal OpenTimelineIO object model: the subset of OTIO 0.18 that toucan's per-frame
// graph builder touches (lib/toucanRender/ImageGraph.cpp, TimelineWrapper.cpp,
// TimelineAlgo.cpp).  OpenTimelineIO itself is not available in this build
// environment; class names, method names and the time algebra follow the OTIO 0.18 API
//  so the graph builder reads like the reference's.  One header instead of
// OTIO's per-class headers.
#pragma once

#include <any>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#define OTIO_NS opentimelineio::v0_18

namespace opentimelineio { inline namespace v0_18 {

// ---- opentime ------------------------------------------------------------------------
class RationalTime {
public:
    explicit constexpr RationalTime(double value = 0, double rate = 1) noexcept : _value(value), _rate(rate) {}
    constexpr double value() const noexcept { return _value; }
    constexpr double rate() const noexcept { return _rate; }
    constexpr double value_rescaled_to(double new_rate) const noexcept
    {
        return new_rate == _rate ? _value : (_value * new_rate) / _rate;
    }
    constexpr double value_rescaled_to(RationalTime rt) const noexcept { return value_rescaled_to(rt._rate); }
    constexpr RationalTime rescaled_to(double new_rate) const noexcept { return RationalTime(value_rescaled_to(new_rate), new_rate); }
    constexpr RationalTime rescaled_to(RationalTime rt) const noexcept { return rescaled_to(rt._rate); }
    constexpr double to_seconds() const noexcept { return _value / _rate; }
    int to_frames() const noexcept { return static_cast<int>(_value); }
    int to_frames(double rate) const noexcept { return static_cast<int>(value_rescaled_to(rate)); }
    RationalTime floor() const { return RationalTime(std::floor(_value), _rate); }
    RationalTime round() const { return RationalTime(std::round(_value), _rate); }

    friend constexpr RationalTime operator+(RationalTime lhs, RationalTime rhs) noexcept
    {
        return (lhs._rate < rhs._rate) ? RationalTime(lhs.value_rescaled_to(rhs._rate) + rhs._value, rhs._rate)
                                       : RationalTime(rhs.value_rescaled_to(lhs._rate) + lhs._value, lhs._rate);
    }
    friend constexpr RationalTime operator-(RationalTime lhs, RationalTime rhs) noexcept
    {
        return (lhs._rate < rhs._rate) ? RationalTime(lhs.value_rescaled_to(rhs._rate) - rhs._value, rhs._rate)
                                       : RationalTime(lhs._value - rhs.value_rescaled_to(lhs._rate), lhs._rate);
    }
    RationalTime& operator+=(RationalTime other) noexcept { return *this = *this + other; }
    RationalTime& operator-=(RationalTime other) noexcept { return *this = *this - other; }
    friend constexpr bool operator>(RationalTime lhs, RationalTime rhs) noexcept { return (lhs._value / lhs._rate) > (rhs._value / rhs._rate); }
    friend constexpr bool operator>=(RationalTime lhs, RationalTime rhs) noexcept { return (lhs._value / lhs._rate) >= (rhs._value / rhs._rate); }
    friend constexpr bool operator<(RationalTime lhs, RationalTime rhs) noexcept { return !(lhs >= rhs); }
    friend constexpr bool operator<=(RationalTime lhs, RationalTime rhs) noexcept { return !(lhs > rhs); }
    friend constexpr bool operator==(RationalTime lhs, RationalTime rhs) noexcept { return lhs.value_rescaled_to(rhs._rate) == rhs._value; }
    friend constexpr bool operator!=(RationalTime lhs, RationalTime rhs) noexcept { return !(lhs == rhs); }

private:
    double _value, _rate;
};

class TimeRange {
public:
    constexpr TimeRange() noexcept {}
    constexpr TimeRange(RationalTime start_time, RationalTime duration) noexcept : _start_time(start_time), _duration(duration) {}
    constexpr RationalTime start_time() const noexcept { return _start_time; }
    constexpr RationalTime duration() const noexcept { return _duration; }
    constexpr RationalTime end_time_exclusive() const noexcept { return _duration + _start_time.rescaled_to(_duration); }
    RationalTime end_time_inclusive() const
    {
        const RationalTime et = end_time_exclusive();
        if ((et - _start_time.rescaled_to(_duration)).value() > 1) {
            return _duration.value() != std::floor(_duration.value()) ? et.floor() : et - RationalTime(1, _duration.rate());
        }
        return _start_time;
    }
    constexpr bool contains(RationalTime other) const noexcept { return _start_time <= other && other < end_time_exclusive(); }
    static constexpr TimeRange range_from_start_end_time(RationalTime start, RationalTime end_exclusive) noexcept
    {
        return TimeRange(start, end_exclusive - start);
    }

private:
    RationalTime _start_time, _duration;
};

// ---- any containers --------------------------------------------------------------------
// JSON values map to: null -> empty any, bool, int64_t (no fraction/exponent), double,
// std::string, AnyVector, AnyDictionary -- the dynamic types the reference's parameter
// marshalling switches on (lib/toucanRender/ImageEffectHost.cpp:277-322).
using AnyDictionary = std::map<std::string, std::any>;
using AnyVector = std::vector<std::any>;

struct ErrorStatus {
    enum Outcome { OK = 0, NOT_IMPLEMENTED, FILE_OPEN_FAILED, JSON_PARSE_ERROR, SCHEMA_ERROR, CANNOT_COMPUTE_AVAILABLE_RANGE };
    Outcome outcome = OK;
    std::string details;
    std::string full_description;
};
inline bool is_error(const ErrorStatus& es) { return es.outcome != ErrorStatus::OK; }
inline bool is_error(const ErrorStatus* es) { return es && es->outcome != ErrorStatus::OK; }

std::any parse_json(const std::string& text, ErrorStatus* error_status);

// ---- object model ----------------------------------------------------------------------
class SerializableObject : public std::enable_shared_from_this<SerializableObject> {
public:
    virtual ~SerializableObject() = default;

    // OTIO's intrusive Retainer, here a shared_ptr with the same surface (.value, ->, bool).
    template <class T = SerializableObject>
    struct Retainer {
        Retainer() = default;
        Retainer(T* p) : value(p) { if (p) _hold = std::static_pointer_cast<SerializableObject>(p->shared_from_this_or_adopt()); }
        Retainer(const std::shared_ptr<T>& p) : value(p.get()), _hold(p) {}
        T* operator->() const { return value; }
        operator T*() const { return value; }
        explicit operator bool() const { return value != nullptr; }
        T* value = nullptr;
        std::shared_ptr<SerializableObject> _hold;
    };

    std::shared_ptr<SerializableObject> shared_from_this_or_adopt()
    {
        auto w = weak_from_this();
        if (auto s = w.lock()) return s;
        return std::shared_ptr<SerializableObject>(this);
    }
};

template <class T, class U>
SerializableObject::Retainer<T> dynamic_retainer_cast(const SerializableObject::Retainer<U>& r)
{
    return SerializableObject::Retainer<T>(dynamic_cast<T*>(r.value));
}

class SerializableObjectWithMetadata : public SerializableObject {
public:
    const std::string& name() const { return _name; }
    void set_name(const std::string& n) { _name = n; }
    AnyDictionary& metadata() { return _metadata; }
    const AnyDictionary& metadata() const { return _metadata; }

protected:
    std::string _name;
    AnyDictionary _metadata;
};

class Effect : public SerializableObjectWithMetadata {
public:
    const std::string& effect_name() const { return _effect_name; }
    void set_effect_name(const std::string& n) { _effect_name = n; }

private:
    std::string _effect_name;
};
class TimeEffect : public Effect {};
class LinearTimeWarp : public TimeEffect {
public:
    double time_scalar() const { return _time_scalar; }
    void set_time_scalar(double s) { _time_scalar = s; }

private:
    double _time_scalar = 1.0;
};

class MediaReference : public SerializableObjectWithMetadata {
public:
    const std::optional<TimeRange>& available_range() const { return _available_range; }
    void set_available_range(const std::optional<TimeRange>& r) { _available_range = r; }
    virtual bool is_missing_reference() const { return false; }

private:
    std::optional<TimeRange> _available_range;
};
class MissingReference : public MediaReference {
public:
    bool is_missing_reference() const override { return true; }
};
class ExternalReference : public MediaReference {
public:
    explicit ExternalReference(const std::string& target_url = std::string()) : _target_url(target_url) {}
    const std::string& target_url() const { return _target_url; }
    void set_target_url(const std::string& u) { _target_url = u; }

private:
    std::string _target_url;
};
class GeneratorReference : public MediaReference {
public:
    const std::string& generator_kind() const { return _generator_kind; }
    void set_generator_kind(const std::string& k) { _generator_kind = k; }
    AnyDictionary& parameters() { return _parameters; }
    const AnyDictionary& parameters() const { return _parameters; }

private:
    std::string _generator_kind;
    AnyDictionary _parameters;
};
class ImageSequenceReference : public MediaReference {
public:
    ImageSequenceReference(const std::string& target_url_base = std::string(), const std::string& name_prefix = std::string(),
                           const std::string& name_suffix = std::string(), int start_frame = 1, int frame_step = 1,
                           double rate = 1, int frame_zero_padding = 0)
        : _target_url_base(target_url_base), _name_prefix(name_prefix), _name_suffix(name_suffix), _start_frame(start_frame),
          _frame_step(frame_step), _rate(rate), _frame_zero_padding(frame_zero_padding)
    {
    }
    const std::string& target_url_base() const { return _target_url_base; }
    const std::string& name_prefix() const { return _name_prefix; }
    const std::string& name_suffix() const { return _name_suffix; }
    int start_frame() const { return _start_frame; }
    int frame_step() const { return _frame_step; }
    double rate() const { return _rate; }
    int frame_zero_padding() const { return _frame_zero_padding; }

    std::string _target_url_base, _name_prefix, _name_suffix;
    int _start_frame, _frame_step;
    double _rate;
    int _frame_zero_padding;
};

class Composition;
class Item;

class Composable : public SerializableObjectWithMetadata {
public:
    virtual bool visible() const { return false; }
    virtual bool overlapping() const { return false; }
    virtual RationalTime duration(ErrorStatus* error_status = nullptr) const = 0;
    Composition* parent() const { return _parent; }
    void _set_parent(Composition* p) { _parent = p; }

private:
    Composition* _parent = nullptr;
};

class Item : public Composable {
public:
    bool visible() const override { return true; }
    const std::optional<TimeRange>& source_range() const { return _source_range; }
    void set_source_range(const std::optional<TimeRange>& r) { _source_range = r; }
    std::vector<Retainer<Effect>>& effects() { return _effects; }
    const std::vector<Retainer<Effect>>& effects() const { return _effects; }
    virtual TimeRange available_range(ErrorStatus* error_status = nullptr) const;
    TimeRange trimmed_range(ErrorStatus* error_status = nullptr) const
    {
        return _source_range ? *_source_range : available_range(error_status);
    }
    RationalTime duration(ErrorStatus* error_status = nullptr) const override { return trimmed_range(error_status).duration(); }
    std::optional<TimeRange> trimmed_range_in_parent(ErrorStatus* error_status = nullptr) const;
    TimeRange range_in_parent(ErrorStatus* error_status = nullptr) const;
    RationalTime transformed_time(RationalTime time, const Item* to_item, ErrorStatus* error_status = nullptr) const;

private:
    std::optional<TimeRange> _source_range;
    std::vector<Retainer<Effect>> _effects;
};

class Gap : public Item {
public:
    bool visible() const override { return false; }
};

class Clip : public Item {
public:
    MediaReference* media_reference() const { return _media_reference.value; }
    void set_media_reference(MediaReference* r) { _media_reference = Retainer<MediaReference>(r); }
    void set_media_reference(const Retainer<MediaReference>& r) { _media_reference = r; }
    TimeRange available_range(ErrorStatus* error_status = nullptr) const override;

private:
    Retainer<MediaReference> _media_reference;
};

class Transition : public Composable {
public:
    bool overlapping() const override { return true; }
    const std::string& transition_type() const { return _transition_type; }
    void set_transition_type(const std::string& t) { _transition_type = t; }
    RationalTime in_offset() const { return _in_offset; }
    RationalTime out_offset() const { return _out_offset; }
    void set_in_offset(RationalTime t) { _in_offset = t; }
    void set_out_offset(RationalTime t) { _out_offset = t; }
    RationalTime duration(ErrorStatus* = nullptr) const override { return _in_offset + _out_offset; }
    std::optional<TimeRange> trimmed_range_in_parent(ErrorStatus* error_status = nullptr) const;

private:
    std::string _transition_type;
    RationalTime _in_offset, _out_offset;
};

class Composition : public Item {
public:
    const std::vector<Retainer<Composable>>& children() const { return _children; }
    void append_child(const Retainer<Composable>& c)
    {
        c.value->_set_parent(this);
        _children.push_back(c);
    }
    int index_of_child(const Composable* child) const;
    virtual TimeRange range_of_child_at_index(int index, ErrorStatus* error_status = nullptr) const = 0;
    TimeRange range_of_child(const Composable* child, ErrorStatus* error_status = nullptr) const;
    std::optional<TimeRange> trim_child_range(TimeRange child_range) const;
    std::optional<TimeRange> trimmed_range_of_child(const Composable* child, ErrorStatus* error_status = nullptr) const;
    std::vector<Retainer<Clip>> find_clips(ErrorStatus* error_status = nullptr, std::optional<TimeRange> search_range = std::nullopt,
                                           bool shallow_search = false) const;

protected:
    std::vector<Retainer<Composable>> _children;
};

class Track : public Composition {
public:
    struct Kind {
        static constexpr const char* video = "Video";
        static constexpr const char* audio = "Audio";
    };
    const std::string& kind() const { return _kind; }
    void set_kind(const std::string& k) { _kind = k; }
    TimeRange range_of_child_at_index(int index, ErrorStatus* error_status = nullptr) const override;
    TimeRange available_range(ErrorStatus* error_status = nullptr) const override;

private:
    std::string _kind = Kind::video;
};

class Stack : public Composition {
public:
    TimeRange range_of_child_at_index(int index, ErrorStatus* error_status = nullptr) const override;
    TimeRange available_range(ErrorStatus* error_status = nullptr) const override;
};

class Timeline : public SerializableObjectWithMetadata {
public:
    Timeline();
    Stack* tracks() const { return _tracks.value; }
    void set_tracks(const Retainer<Stack>& s) { _tracks = s; }
    const std::optional<RationalTime>& global_start_time() const { return _global_start_time; }
    void set_global_start_time(const std::optional<RationalTime>& t) { _global_start_time = t; }
    RationalTime duration(ErrorStatus* error_status = nullptr) const { return _tracks.value->duration(error_status); }
    std::vector<Retainer<Clip>> find_clips(ErrorStatus* error_status = nullptr) const { return _tracks.value->find_clips(error_status); }

    static SerializableObject* from_json_file(const std::string& file_name, ErrorStatus* error_status = nullptr);
    static SerializableObject* from_json_string(const std::string& text, ErrorStatus* error_status = nullptr);

private:
    Retainer<Stack> _tracks;
    std::optional<RationalTime> _global_start_time;
};

}} // namespace opentimelineio::v0_18
