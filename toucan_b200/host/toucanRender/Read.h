// Read nodes: the leaves of the image graph.  In the reference these decode a file with
// OpenImageIO / FFmpeg / lunasvg on EVERY exec() (lib/toucanRender/Read.cpp:64-101,164-216);
// here a leaf is a decoded RGBA frame taken from a MediaStore -- host memory (uploaded on
// every exec, mirroring the per-frame decode) or device memory kept resident -- so the hot
// path starts at the decoded frame.  File decode itself is outside the
// kernel scope; PNG files are decoded by a small built-in reader when no frame was
// registered for the path.
#pragma once

#include <toucanRender/ImageNode.h>

#include <filesystem>
#include <condition_variable>
#include <list>
#include <map>
#include <mutex>
#include <set>

namespace toucan
{
    //! One decoded frame: tightly packed interleaved pixels on the host or the device.
    struct MediaFrame
    {
        const void* host = nullptr;
        void* device = nullptr;
        int width = 0;
        int height = 0;
        int nchannels = 4;
        int format = TC_U8;
        std::shared_ptr<std::vector<unsigned char> > owned; // storage for frames decoded by the built-in readers
        std::shared_ptr<unsigned char> storage;               // ... or a plain (uninitialised) array for the large ones
    };

    //! Reader option, the counterpart of OpenImageIO's "oiio:UnassociatedAlpha" open-time hint: PNG (and TIFF ExtraSamples = 2)
    //! files store UNassociated alpha; OIIO's readers multiply colour by alpha on read unless that hint is set, and the
    //! reference opens its media with no hints (lib/toucanRender/Read.cpp:39) -- so the default here is to associate.  The
    //! reference's own rendered film strips (images/Transition*.png) show UNassociated Counter frames entering the Dissolve, i.e.
    //! they were made by a build whose read path kept the file's values; `true` reproduces those renders
    //! (tests/test_reference_filmstrips.py runs both).  Process-wide; initial value from TOUCAN_UNASSOCIATED_ALPHA=1.
    void setKeepUnassociatedAlpha(bool);
    bool getKeepUnassociatedAlpha();

    //! A file held in memory (lib/toucanRender/MemoryMap.h: MemoryReference): an entry of a mapped .otioz.
    struct MemoryReference
    {
        std::shared_ptr<const void> owner; //!< keeps the mapping alive
        const unsigned char* data = nullptr;
        size_t size = 0;
        bool isValid() const { return data != nullptr && size > 0; }
    };

    //! In-memory files by URL (TimelineWrapper.cpp:196-246: the media of an .otioz archive).
    typedef std::map<std::string, MemoryReference> MemoryReferences;

    //! Registry path -> decoded frame, shared by all read nodes of a timeline.
    class MediaStore
    {
    public:
        void set(const std::string& path, const MediaFrame&);
        bool get(const std::string& path, MediaFrame&);
        void erase(const std::string& path);
        void clear();

        //! Process-wide default store.
        static const std::shared_ptr<MediaStore>& global();

        //! Device source cache: keep up to `bytes` of uploaded frames resident per
        //! GPU so a still is uploaded once instead of on every frame (the reference re-decodes
        //! its media on every exec(), lib/toucanRender/Read.cpp:64-101).  0 = off (the default:
        //! host frames are uploaded on every exec()).  Least recently used frames are evicted.
        void setDeviceCacheBytes(size_t bytes);
        size_t getDeviceCacheBytes() const;
        bool getCached(int device, const std::string& path, Image&);
        void putCached(int device, const std::string& path, const Image&);
        void dropCached();

        //! Decode coordination: the first caller for a path gets true and must call endLoad(); others block until it has.
        bool beginLoad(const std::string& path);
        void endLoad(const std::string& path);

        //! Upload coordination for the device cache: the first caller for (device, path) gets true, uploads, puts the frame in
        //! the cache and calls endUpload(); a concurrent caller blocks until then and gets false (the frame is cached by then).
        bool beginUpload(int device, const std::string& path);
        void endUpload(int device, const std::string& path);

        //! Host -> device uploads of source frames so far (diagnostics).
        void countUpload(size_t bytes);
        size_t getUploadCount() const;
        size_t getUploadBytes() const;

    private:
        struct Cached
        {
            int device = 0;
            std::string path;
            Image image;
        };
        mutable std::mutex _mutex;
        std::map<std::string, MediaFrame> _frames;
        size_t _cacheLimit = 0;
        size_t _uploads = 0, _uploadBytes = 0;
        std::set<std::string> _loading;
        std::set<std::pair<int, std::string> > _uploading;
        std::condition_variable _loaded;
        std::list<Cached> _cache; // most recently used first
    };

    //! Base class for read nodes.
    class IReadNode : public IImageNode
    {
    public:
        IReadNode(const std::string& name);

        virtual ~IReadNode() = 0;

        const ImageSpec& getSpec() const;
        bool sizeHint(int& width, int& height) const override;

        const OTIO_NS::TimeRange& getTimeRange() const;

    protected:
        Image _toDevice(const MediaFrame&, const std::shared_ptr<MediaStore>&, const std::string& path) const;

        ImageSpec _spec;
        OTIO_NS::TimeRange _timeRange;
    };

    //! Image read node.
    class ImageReadNode : public IReadNode
    {
    public:
        ImageReadNode(
            const std::filesystem::path&,
            const std::shared_ptr<MediaStore>& = MediaStore::global(),
            const MemoryReference& = MemoryReference());

        virtual ~ImageReadNode();

        std::string getLabel() const override;

        Image exec() override;

        static std::vector<std::string> getExtensions();

    private:
        std::filesystem::path _path;
        std::shared_ptr<MediaStore> _store;
        MemoryReference _memory;
    };

    //! Image sequence read node.
    class SequenceReadNode : public IReadNode
    {
    public:
        SequenceReadNode(
            const std::string& base,
            const std::string& namePrefix,
            const std::string& nameSuffix,
            int startFrame,
            int frameStep,
            double rate,
            int frameZeroPadding,
            const std::shared_ptr<MediaStore>& = MediaStore::global(),
            const MemoryReferences& = MemoryReferences());

        virtual ~SequenceReadNode();

        std::string getLabel() const override;

        Image exec() override;

        static std::vector<std::string> getExtensions();

    private:
        std::string _base;
        std::string _namePrefix;
        std::string _nameSuffix;
        int _startFrame = 1;
        int _frameStep = 1;
        double _rate = 1.0;
        int _frameZeroPadding = 0;
        std::shared_ptr<MediaStore> _store;
        MemoryReferences _memoryReferences;
    };

    //! Decode, in the background, the stills and sequence frames a timeline references (in timeline order) and park them
    //! in the store, so the render threads find them decoded.  Returns immediately; the returned handle joins on destruction.
    class MediaPrefetch
    {
    public:
        MediaPrefetch(const std::shared_ptr<MediaStore>&, const std::vector<std::pair<std::string, MemoryReference> >& media, int threads = 2);
        ~MediaPrefetch();

    private:
        struct State;
        std::shared_ptr<State> _state;
    };

    //! Upload a host frame into the store's device cache for `device` on the calling thread's stream (or find it there):
    //! one upload per (device, path) however many threads ask.
    Image uploadToCache(const MediaFrame&, const std::shared_ptr<MediaStore>&, int device, const std::string& path);

    //! Look a frame up in the store, falling back to the built-in file readers.
    //! `park`: keep the decoded frame in the store (stills, decoded once); sequence frames are used once and are not kept.
    bool readMedia(const std::shared_ptr<MediaStore>&, const std::string& path, MediaFrame&, const MemoryReference& = MemoryReference(),
        bool park = true);

    //! Create a read node for a single image.  Throws std::runtime_error if the media cannot
    //! be opened (the caller logs it and the clip yields no node, ImageGraph.cpp:355-367).
    std::shared_ptr<IReadNode> createReadNode(
        const std::filesystem::path&,
        const std::shared_ptr<MediaStore>& = MediaStore::global(),
        const MemoryReference& = MemoryReference());

    //! Create a read node for an image sequence.
    std::shared_ptr<IReadNode> createReadNode(
        const std::string& base,
        const std::string& namePrefix,
        const std::string& nameSuffix,
        int startFrame,
        int frameStep,
        double rate,
        int frameZeroPadding,
        const std::shared_ptr<MediaStore>& = MediaStore::global(),
        const MemoryReferences& = MemoryReferences());
}
