#include "TimelineWrapper.h"

#include "Util.h"

#include <algorithm>
#include <set>

namespace toucan
{
    namespace
    {
        OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline> timelineFromFile(const std::filesystem::path& path)
        {
            OTIO_NS::ErrorStatus errorStatus;
            OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline> out(
                dynamic_cast<OTIO_NS::Timeline*>(OTIO_NS::Timeline::from_json_file(path.string(), &errorStatus)));
            if (!out)
            {
                throw std::runtime_error(errorStatus.full_description);
            }
            return out;
        }
    }

    TimelineWrapper::TimelineWrapper(const std::filesystem::path& path) :
        _path(path),
        _store(MediaStore::global())
    {
        const std::string extension = toLower(_path.extension().string());
        if (".otio" == extension)
        {
            _timeline = timelineFromFile(_path);
        }
        else if (".otiod" == extension)
        {
            // an unpacked bundle: content.otio + media/ inside the directory (TimelineWrapper.cpp:84-95)
            _path = _path / "content.otio";
            _timeline = timelineFromFile(_path);
        }
        else if (".otioz" == extension)
        {
            // TimelineWrapper.cpp:146-246: map the archive, inflate content.otio, reference the stored media in place
            _archive = std::make_shared<ZipArchive>(path);
            const ZipArchive::Entry* content = _archive->find("content.otio");
            if (!content)
            {
                throw std::runtime_error("Cannot locate: content.otio");
            }
            const std::vector<unsigned char> json = _archive->read(*content);
            OTIO_NS::ErrorStatus errorStatus;
            _timeline = OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline>(dynamic_cast<OTIO_NS::Timeline*>(
                OTIO_NS::Timeline::from_json_string(std::string(json.begin(), json.end()), &errorStatus)));
            if (!_timeline)
            {
                throw std::runtime_error(errorStatus.full_description);
            }
            // URLs inside an archive are only meaningful for this timeline: a private store
            _store = std::make_shared<MediaStore>();
            auto reference = [this](const std::string& url)
            {
                const std::string name = splitURLProtocol(url).second;
                const ZipArchive::Entry* entry = _archive->find(name);
                if (!entry)
                {
                    throw std::runtime_error("Cannot locate: " + name);
                }
                if (entry->method != 0)
                {
                    throw std::runtime_error("Media is not uncompressed: " + name);
                }
                MemoryReference memory;
                memory.owner = _archive;
                memory.data = _archive->data() + entry->dataOffset;
                memory.size = entry->size;
                _memoryReferences[url] = memory;
            };
            for (const auto& clip : _timeline->find_clips())
            {
                if (auto externalRef = dynamic_cast<OTIO_NS::ExternalReference*>(clip->media_reference()))
                {
                    reference(externalRef->target_url());
                }
                else if (auto sequenceRef = dynamic_cast<OTIO_NS::ImageSequenceReference*>(clip->media_reference()))
                {
                    const OTIO_NS::TimeRange timeRange = clip->available_range();
                    for (int64_t frame = static_cast<int64_t>(timeRange.start_time().value());
                        frame <= static_cast<int64_t>(timeRange.end_time_inclusive().value());
                        ++frame)
                    {
                        reference(getSequenceFrame(
                            sequenceRef->target_url_base(),
                            sequenceRef->name_prefix(),
                            static_cast<int>(frame),
                            sequenceRef->frame_zero_padding(),
                            sequenceRef->name_suffix()));
                    }
                }
            }
        }
        else if (hasExtension(extension, ImageReadNode::getExtensions()))
        {
            // an image or a numbered image sequence opened as a one-clip timeline (TimelineWrapper.cpp:262-306)
            const auto sequence = getSequence(path);
            const auto split = splitFileNameNumber(sequence.front().stem().string());
            _timeline = OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline>(new OTIO_NS::Timeline);
            OTIO_NS::SerializableObject::Retainer<OTIO_NS::Track> track(new OTIO_NS::Track);
            _timeline->tracks()->append_child(OTIO_NS::SerializableObject::Retainer<OTIO_NS::Composable>(track.value));
            OTIO_NS::SerializableObject::Retainer<OTIO_NS::Clip> clip(new OTIO_NS::Clip);
            track->append_child(OTIO_NS::SerializableObject::Retainer<OTIO_NS::Composable>(clip.value));
            if (split.second.empty())
            {
                clip->set_media_reference(new OTIO_NS::ExternalReference(path.string()));
                clip->set_source_range(OTIO_NS::TimeRange(OTIO_NS::RationalTime(0.0, 24.0), OTIO_NS::RationalTime(1.0, 24.0)));
            }
            else
            {
                const int start = std::atoi(split.second.c_str());
                const double rate = 24.0;
                clip->set_media_reference(new OTIO_NS::ImageSequenceReference(
                    sequence.front().parent_path().string(),
                    split.first,
                    sequence.front().extension().string(),
                    start,
                    1,
                    rate,
                    static_cast<int>(getNumberPadding(split.second))));
                clip->set_source_range(OTIO_NS::TimeRange(
                    OTIO_NS::RationalTime(start, rate),
                    OTIO_NS::RationalTime(static_cast<double>(sequence.size()), rate)));
            }
        }
        else
        {
            // movie and SVG front-ends need FFmpeg / lunasvg decoders: outside this path
            throw std::runtime_error("Unrecognized file");
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
        if (_archive)
        {
            return url; // in-memory media are keyed by their URL (TimelineWrapper.cpp:339-342)
        }
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
            const auto memory = _memoryReferences.find(externalRef->target_url());
            out = toucan::createReadNode(
                getMediaPath(externalRef->target_url()),
                _store,
                memory != _memoryReferences.end() ? memory->second : MemoryReference());
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
                _store,
                _memoryReferences);
        }
        return out;
    }

    std::vector<std::pair<std::string, MemoryReference> > TimelineWrapper::getStillMedia() const
    {
        std::vector<std::pair<std::string, MemoryReference> > out;
        std::set<std::string> seen;
        for (const auto& clip : _timeline->find_clips())
        {
            if (auto externalRef = dynamic_cast<OTIO_NS::ExternalReference*>(clip->media_reference()))
            {
                const std::string path = getMediaPath(externalRef->target_url());
                if (!hasExtension(toLower(std::filesystem::path(path).extension().string()), ImageReadNode::getExtensions()) ||
                    !seen.insert(path).second)
                {
                    continue;
                }
                const auto memory = _memoryReferences.find(externalRef->target_url());
                out.emplace_back(path, memory != _memoryReferences.end() ? memory->second : MemoryReference());
            }
        }
        return out;
    }

    std::vector<std::pair<std::string, MemoryReference> > TimelineWrapper::getStillMediaByFirstUse(const OTIO_NS::RationalTime& from) const
    {
        // The order a render starting at `from` (a time in the tracks' coordinates) first touches the stills: the clips under
        // `from`, then the later ones by their start, then the ones already over (a transition's out-going clip is still read a
        // little past its end: it is under `from` or uploaded at first use, which the cache handles either way).
        struct Entry
        {
            double key0, key1;
            std::pair<std::string, MemoryReference> media;
        };
        std::vector<Entry> entries;
        const double t0 = from.to_seconds();
        for (const auto& child : _timeline->tracks()->children())
        {
            auto track = dynamic_cast<OTIO_NS::Track*>(child.value);
            if (!track) continue;
            for (const auto& clip : track->find_clips())
            {
                auto externalRef = dynamic_cast<OTIO_NS::ExternalReference*>(clip->media_reference());
                if (!externalRef) continue;
                const std::string path = getMediaPath(externalRef->target_url());
                if (!hasExtension(toLower(std::filesystem::path(path).extension().string()), ImageReadNode::getExtensions())) continue;
                double start = 0.0, end = 0.0;
                if (const auto range = clip->trimmed_range_in_parent())
                {
                    start = range->start_time().to_seconds();
                    end = range->end_time_exclusive().to_seconds();
                }
                const auto memory = _memoryReferences.find(externalRef->target_url());
                entries.push_back({ end > t0 ? 0.0 : 1.0, std::max(0.0, start - t0),
                    std::make_pair(path, memory != _memoryReferences.end() ? memory->second : MemoryReference()) });
            }
        }
        std::stable_sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& b)
        {
            return a.key0 != b.key0 ? a.key0 < b.key0 : a.key1 < b.key1;
        });
        std::vector<std::pair<std::string, MemoryReference> > out;
        std::set<std::string> seen;
        for (const auto& e : entries)
        {
            if (seen.insert(e.media.first).second) out.push_back(e.media);
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
