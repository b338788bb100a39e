// Opens a timeline and resolves its media.  Interface parity:
// lib/toucanRender/TimelineWrapper.h:21-52.  Opens .otio files, .otiod directories, .otioz
// archives (media referenced in place inside the memory-mapped ZIP), single images and
// numbered image sequences (as one-clip timelines) and, as an extension for tests and
// synthetic workloads, a timeline given as a JSON string.  Movie and SVG files need the
// FFmpeg / lunasvg decoders and throw "Unrecognized file".
#pragma once

#include <toucanRender/Archive.h>
#include <toucanRender/Read.h>

#include <opentimelineio/otio.h>

#include <filesystem>
#include <memory>

namespace toucan
{
    class TimelineWrapper : public std::enable_shared_from_this<TimelineWrapper>
    {
    public:
        TimelineWrapper(const std::filesystem::path&);

        //! Extension: a timeline from JSON text; relative media URLs resolve against `directory`.
        TimelineWrapper(const std::string& json, const std::filesystem::path& directory);

        ~TimelineWrapper();

        const std::filesystem::path& getPath() const;

        const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline>& getTimeline() const;

        const OTIO_NS::TimeRange& getTimeRange() const;

        std::string getMediaPath(const std::string& url) const;

        std::shared_ptr<IReadNode> createReadNode(const OTIO_NS::MediaReference*);

        //! The stills (single-file media) the timeline's clips reference, in clip order, without duplicates: what a
        //! background prefetch decodes ahead of the render threads.  Sequence frames are not listed (they stream).
        std::vector<std::pair<std::string, MemoryReference> > getStillMedia() const;
        //! The same stills in the order a render starting at `from` (a time in the tracks' coordinates) first touches them.
        std::vector<std::pair<std::string, MemoryReference> > getStillMediaByFirstUse(const OTIO_NS::RationalTime& from) const;

        //! Decoded frames the read nodes draw from (defaults to the process-wide store).
        void setMediaStore(const std::shared_ptr<MediaStore>&);
        const std::shared_ptr<MediaStore>& getMediaStore() const;

    private:
        void _init();

        std::filesystem::path _path;
        OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline> _timeline;
        OTIO_NS::TimeRange _timeRange;
        std::shared_ptr<MediaStore> _store;
        std::shared_ptr<ZipArchive> _archive;
        MemoryReferences _memoryReferences;
    };

    //! Get the video clips of a timeline, track by track (lib/toucanRender/TimelineAlgo.cpp:10-26).
    std::vector<OTIO_NS::SerializableObject::Retainer<OTIO_NS::Clip> >
        getVideoClips(const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Timeline>&);
}
