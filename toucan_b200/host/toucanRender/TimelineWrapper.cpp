#include "TimelineWrapper.h"

#include "Util.h"

namespace toucan
{
    TimelineWrapper::TimelineWrapper(const std::filesystem::path& path) :
        _path(path),
        _store(MediaStore::global())
    {
        const std::string extension = toLower(_path.extension().string());
        if (".otio" != extension)
        {
            throw std::runtime_error("Unrecognized file");
        }
        OTIO_NS::ErrorStatus errorStatus;
        _timeline = OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline>(
            dynamic_cast<OTIO_NS::Timeline*>(OTIO_NS::Timeline::from_json_file(_path.string(), &errorStatus)));
        if (!_timeline)
        {
            throw std::runtime_error(errorStatus.full_description);
        }
        _init();
    }

    TimelineWrapper::TimelineWrapper(const std::string& json, const std::filesystem::path& directory) :
        _path(directory / "timeline.otio"),
        _store(MediaStore::global())
    {
        OTIO_NS::ErrorStatus errorStatus;
        _timeline = OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline>(
            dynamic_cast<OTIO_NS::Timeline*>(OTIO_NS::Timeline::from_json_string(json, &errorStatus)));
        if (!_timeline)
        {
            throw std::runtime_error(errorStatus.full_description);
        }
        _init();
    }

    void TimelineWrapper::_init()
    {
        // global start time if present, else zero at the timeline's rate (TimelineWrapper.cpp:311-316)
        const auto globalStartTime = _timeline->global_start_time();
        _timeRange = OTIO_NS::TimeRange(
            globalStartTime.has_value() ?
                globalStartTime.value() :
                OTIO_NS::RationalTime(0.0, _timeline->duration().rate()),
            _timeline->duration());
    }

    TimelineWrapper::~TimelineWrapper()
    {}

    const std::filesystem::path& TimelineWrapper::getPath() const { return _path; }

    const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline>& TimelineWrapper::getTimeline() const { return _timeline; }

    const OTIO_NS::TimeRange& TimelineWrapper::getTimeRange() const { return _timeRange; }

    std::string TimelineWrapper::getMediaPath(const std::string& url) const
    {
        std::filesystem::path path = splitURLProtocol(url).second;
        if (!path.is_absolute())
        {
            path = _path.parent_path() / path;
        }
        return path.string();
    }

    std::shared_ptr<IReadNode> TimelineWrapper::createReadNode(const OTIO_NS::MediaReference* ref)
    {
        std::shared_ptr<IReadNode> out;
        if (auto externalRef = dynamic_cast<const OTIO_NS::ExternalReference*>(ref))
        {
            out = toucan::createReadNode(getMediaPath(externalRef->target_url()), _store);
        }
        else if (auto sequenceRef = dynamic_cast<const OTIO_NS::ImageSequenceReference*>(ref))
        {
            out = toucan::createReadNode(
                getMediaPath(sequenceRef->target_url_base()),
                sequenceRef->name_prefix(),
                sequenceRef->name_suffix(),
                sequenceRef->start_frame(),
                sequenceRef->frame_step(),
                sequenceRef->rate(),
                sequenceRef->frame_zero_padding(),
                _store);
        }
        return out;
    }

    void TimelineWrapper::setMediaStore(const std::shared_ptr<MediaStore>& value) { _store = value; }

    const std::shared_ptr<MediaStore>& TimelineWrapper::getMediaStore() const { return _store; }

    std::vector<OTIO_NS::SerializableObject::Retainer<OTIO_NS::Clip> >
        getVideoClips(const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline>& timeline)
    {
        std::vector<OTIO_NS::SerializableObject::Retainer<OTIO_NS::Clip> > out;
        for (const auto& child : timeline->tracks()->children())
        {
            auto track = OTIO_NS::dynamic_retainer_cast<OTIO_NS::Track>(child);
            if (track && OTIO_NS::Track::Kind::video == track->kind())
            {
                const auto clips = track->find_clips(nullptr, std::nullopt, true);
                out.insert(out.end(), clips.begin(), clips.end());
            }
        }
        return out;
    }
}
