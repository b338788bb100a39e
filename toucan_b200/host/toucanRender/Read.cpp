#include "Read.h"

#include "ExrRead.h"
#include "JpegRead.h"
#include "PngRead.h"
#include "TiffRead.h"
#include "Util.h"

#include <atomic>
#include <cstdlib>
#include <sstream>
#include <thread>

namespace toucan
{
    void MediaStore::set(const std::string& path, const MediaFrame& frame)
    {
        std::lock_guard<std::mutex> lock(_mutex);
        _frames[path] = frame;
    }

    bool MediaStore::get(const std::string& path, MediaFrame& out)
    {
        std::lock_guard<std::mutex> lock(_mutex);
        const auto i = _frames.find(path);
        if (i == _frames.end())
        {
            return false;
        }
        out = i->second;
        return true;
    }

    void MediaStore::erase(const std::string& path)
    {
        std::lock_guard<std::mutex> lock(_mutex);
        _frames.erase(path);
    }

    void MediaStore::clear()
    {
        std::lock_guard<std::mutex> lock(_mutex);
        _frames.clear();
    }

    void MediaStore::setDeviceCacheBytes(size_t bytes)
    {
        std::lock_guard<std::mutex> lock(_mutex);
        _cacheLimit = bytes;
        if (0 == bytes) _cache.clear();
    }

    size_t MediaStore::getDeviceCacheBytes() const
    {
        std::lock_guard<std::mutex> lock(_mutex);
        return _cacheLimit;
    }

    bool MediaStore::getCached(int device, const std::string& path, Image& out)
    {
        std::lock_guard<std::mutex> lock(_mutex);
        for (auto i = _cache.begin(); i != _cache.end(); ++i)
        {
            if (i->device == device && i->path == path)
            {
                out = i->image;
                _cache.splice(_cache.begin(), _cache, i);
                return true;
            }
        }
        return false;
    }

    void MediaStore::putCached(int device, const std::string& path, const Image& image)
    {
        std::lock_guard<std::mutex> lock(_mutex);
        if (0 == _cacheLimit) return;
        _cache.push_front(Cached{ device, path, image });
        // evict the least recently used frames of this device beyond the budget
        size_t bytes = 0;
        for (auto i = _cache.begin(); i != _cache.end();)
        {
            if (i->device == device)
            {
                bytes += i->image.spec().image_bytes();
                if (bytes > _cacheLimit && i != _cache.begin())
                {
                    i = _cache.erase(i);
                    continue;
                }
            }
            ++i;
        }
    }

    bool MediaStore::beginLoad(const std::string& path)
    {
        std::unique_lock<std::mutex> lock(_mutex);
        if (_loading.insert(path).second)
        {
            return true;
        }
        _loaded.wait(lock, [this, &path] { return 0 == _loading.count(path); });
        return false;
    }

    void MediaStore::endLoad(const std::string& path)
    {
        {
            std::lock_guard<std::mutex> lock(_mutex);
            _loading.erase(path);
        }
        _loaded.notify_all();
    }

    bool MediaStore::beginUpload(int device, const std::string& path)
    {
        std::unique_lock<std::mutex> lock(_mutex);
        const auto key = std::make_pair(device, path);
        if (_uploading.insert(key).second)
        {
            return true;
        }
        _loaded.wait(lock, [this, &key] { return 0 == _uploading.count(key); });
        return false;
    }

    void MediaStore::endUpload(int device, const std::string& path)
    {
        {
            std::lock_guard<std::mutex> lock(_mutex);
            _uploading.erase(std::make_pair(device, path));
        }
        _loaded.notify_all();
    }

    void MediaStore::countUpload(size_t bytes)
    {
        std::lock_guard<std::mutex> lock(_mutex);
        ++_uploads;
        _uploadBytes += bytes;
    }

    size_t MediaStore::getUploadCount() const
    {
        std::lock_guard<std::mutex> lock(_mutex);
        return _uploads;
    }

    size_t MediaStore::getUploadBytes() const
    {
        std::lock_guard<std::mutex> lock(_mutex);
        return _uploadBytes;
    }

    void MediaStore::dropCached()
    {
        std::lock_guard<std::mutex> lock(_mutex);
        _cache.clear();
    }

    const std::shared_ptr<MediaStore>& MediaStore::global()
    {
        static const std::shared_ptr<MediaStore> store = std::make_shared<MediaStore>();
        return store;
    }

    namespace
    {
        bool decodeMedia(const std::string& path, MediaFrame& out, const MemoryReference& memory)
        {
            // Built-in readers: PNG (zlib), JPEG, scanline OpenEXR and stripped TIFF.
            const std::string extension = toLower(std::filesystem::path(path).extension().string());
            const bool inMemory = memory.isValid();
            try
            {
                if (".png" == extension)
                {
                    return inMemory ? readPngMemory(memory.data, memory.size, out) : readPng(path, out);
                }
                if (".jpg" == extension || ".jpeg" == extension)
                {
                    return inMemory ? readJpegMemory(memory.data, memory.size, out) : readJpeg(path, out);
                }
                if (".exr" == extension)
                {
                    return inMemory ? readExrMemory(memory.data, memory.size, out) : readExr(path, out);
                }
                if (".tif" == extension || ".tiff" == extension)
                {
                    return inMemory ? readTiffMemory(memory.data, memory.size, out) : readTiff(path, out);
                }
            }
            catch (const std::exception&)
            {
                // OIIO::ImageInput::open returns null for a file it cannot decode (Read.cpp:40-46)
            }
            return false;
        }
    }

    bool readMedia(const std::shared_ptr<MediaStore>& store, const std::string& path, MediaFrame& out, const MemoryReference& memory, bool park)
    {
        if (store && store->get(path, out))
        {
            return true;
        }
        if (!store || !park)
        {
            return decodeMedia(path, out, memory);
        }
        // The decoded frame is parked in the store so a still is decoded once (it is still uploaded on every exec()
        // unless the device source cache is on).  One thread decodes a given file; a concurrent caller -- a render
        // thread catching up with the prefetcher, or the other way round -- waits for it instead of decoding again.
        if (!store->beginLoad(path))
        {
            return store->get(path, out);
        }
        bool read = false;
        try
        {
            read = decodeMedia(path, out, memory);
            if (read)
            {
                // A parked still lives as long as the store: page-lock the large ones (when there is a CUDA device to
                // upload to), so every GPU's upload runs at PCIe speed.  Frames that are not parked stay pageable --
                // an asynchronous copy from locked memory could outlive them.
                const size_t bytes = tc_image_bytes(out.width, out.height, out.format);
                void* p = const_cast<void*>(out.host);
                if (bytes >= (size_t(32) << 20) && p && TC_OK == tc_host_register(p, bytes))
                {
                    const auto innerStorage = out.storage;
                    const auto innerOwned = out.owned;
                    out.storage = std::shared_ptr<unsigned char>(static_cast<unsigned char*>(p), [innerStorage, innerOwned](unsigned char* q)
                    {
                        tc_host_unregister(q); // before the captured owners free the memory
                    });
                }
                store->set(path, out);
            }
        }
        catch (...)
        {
            store->endLoad(path);
            throw;
        }
        store->endLoad(path);
        return read;
    }

    struct MediaPrefetch::State
    {
        std::shared_ptr<MediaStore> store;
        std::vector<std::pair<std::string, MemoryReference> > media;
        std::atomic<size_t> next{ 0 };
        std::atomic<bool> stop{ false };
        std::vector<std::thread> threads;
    };

    MediaPrefetch::MediaPrefetch(
        const std::shared_ptr<MediaStore>& store,
        const std::vector<std::pair<std::string, MemoryReference> >& media,
        int threads) :
        _state(std::make_shared<State>())
    {
        _state->store = store;
        _state->media = media;
        State* state = _state.get();
        for (int t = 0; t < std::max(1, threads); ++t)
        {
            _state->threads.emplace_back([state]
            {
                for (;;)
                {
                    const size_t i = state->next++;
                    if (i >= state->media.size() || state->stop) return;
                    MediaFrame frame;
                    readMedia(state->store, state->media[i].first, frame, state->media[i].second);
                }
            });
        }
    }

    MediaPrefetch::~MediaPrefetch()
    {
        _state->stop = true;
        for (auto& t : _state->threads)
        {
            if (t.joinable()) t.join();
        }
    }

    IReadNode::IReadNode(const std::string& name) :
        IImageNode(name)
    {}

    IReadNode::~IReadNode()
    {}

    const ImageSpec& IReadNode::getSpec() const { return _spec; }

    bool IReadNode::sizeHint(int& width, int& height) const
    {
        width = _spec.width;
        height = _spec.height;
        return width > 0 && height > 0;
    }

    const OTIO_NS::TimeRange& IReadNode::getTimeRange() const { return _timeRange; }

    Image uploadToCache(const MediaFrame& frame, const std::shared_ptr<MediaStore>& store, int device, const std::string& path)
    {
        const ImageSpec spec(frame.width, frame.height, 4, frame.format);
        Image cached;
        for (;;)
        {
            if (store->getCached(device, path, cached) && cached.spec().image_bytes() == spec.image_bytes())
            {
                // another thread of this GPU (a second worker, or the scheduler's upload-ahead thread) may have uploaded it
                // on its own stream: wait for that copy on ours
                cached.acquire();
                return cached;
            }
            if (store->beginUpload(device, path)) break; // ours to upload; otherwise someone just finished: look again
        }
        try
        {
            cached = Image::upload(frame.host, spec);
            cached.keepAlive(frame.storage ? std::shared_ptr<const void>(frame.storage) : std::shared_ptr<const void>(frame.owned));
            cached.publish(); // before any other thread can find it in the cache
            store->countUpload(spec.image_bytes());
            store->putCached(device, path, cached);
        }
        catch (...)
        {
            store->endUpload(device, path);
            throw;
        }
        store->endUpload(device, path);
        return cached;
    }

    Image IReadNode::_toDevice(const MediaFrame& frame, const std::shared_ptr<MediaStore>& store, const std::string& path) const
    {
        if (frame.nchannels != 4)
        {
            // RGB sources are padded with A = 1 at decode time (Read.cpp:86-94)
            throw std::runtime_error("read node: frames must be RGBA (pad RGB with A=1 at decode)");
        }
        const ImageSpec spec(frame.width, frame.height, 4, frame.format);
        if (frame.device)
        {
            return Image::wrap(frame.device, spec);
        }
        if (store && store->getDeviceCacheBytes() > 0)
        {
            return uploadToCache(frame, store, DeviceContext::device(), path);
        }
        if (store) store->countUpload(spec.image_bytes());
        Image uploaded = Image::upload(frame.host, spec);
        uploaded.keepAlive(frame.storage ? std::shared_ptr<const void>(frame.storage) : std::shared_ptr<const void>(frame.owned));
        return uploaded;
    }

    ImageReadNode::ImageReadNode(
        const std::filesystem::path& path,
        const std::shared_ptr<MediaStore>& store,
        const MemoryReference& memory) :
        IReadNode("ImageRead"),
        _path(path),
        _store(store),
        _memory(memory)
    {
        MediaFrame frame;
        if (!readMedia(_store, _path.string(), frame, _memory))
        {
            std::stringstream ss;
            ss << "Cannot open file: " << _path.string();
            throw std::runtime_error(ss.str());
        }
        _spec = ImageSpec(frame.width, frame.height, frame.nchannels, frame.format);
    }

    ImageReadNode::~ImageReadNode()
    {}

    std::string ImageReadNode::getLabel() const
    {
        return "Read: " + _path.filename().string();
    }

    Image ImageReadNode::exec()
    {
        MediaFrame frame;
        if (!readMedia(_store, _path.string(), frame, _memory))
        {
            return Image();
        }
        return _toDevice(frame, _store, _path.string());
    }

    std::vector<std::string> ImageReadNode::getExtensions()
    {
        return { ".exr", ".tif", ".tiff", ".jpg", ".jpeg", ".png" };
    }

    SequenceReadNode::SequenceReadNode(
        const std::string& base,
        const std::string& namePrefix,
        const std::string& nameSuffix,
        int startFrame,
        int frameStep,
        double rate,
        int frameZeroPadding,
        const std::shared_ptr<MediaStore>& store,
        const MemoryReferences& memoryReferences) :
        IReadNode("SequenceRead"),
        _base(base),
        _namePrefix(namePrefix),
        _nameSuffix(nameSuffix),
        _startFrame(startFrame),
        _frameStep(frameStep),
        _rate(rate),
        _frameZeroPadding(frameZeroPadding),
        _store(store),
        _memoryReferences(memoryReferences)
    {
        // Image information comes from the first frame; a missing first frame leaves an
        // empty spec instead of throwing (Read.cpp:138-150).
        MediaFrame frame;
        const std::string first = getSequenceFrame(_base, _namePrefix, _startFrame, _frameZeroPadding, _nameSuffix);
        const auto memory = _memoryReferences.find(first);
        if (readMedia(_store, first, frame, memory != _memoryReferences.end() ? memory->second : MemoryReference(), false))
        {
            _spec = ImageSpec(frame.width, frame.height, frame.nchannels, frame.format);
        }
        _timeRange = OTIO_NS::TimeRange(
            OTIO_NS::RationalTime(_startFrame, _rate),
            OTIO_NS::RationalTime(1.0, _rate));
    }

    SequenceReadNode::~SequenceReadNode()
    {}

    std::string SequenceReadNode::getLabel() const
    {
        // the frame the node is currently set to (the reference prints the first frame)
        return "Read: " + std::filesystem::path(
            getSequenceFrame(_base, _namePrefix, _time.to_frames(), _frameZeroPadding, _nameSuffix)).filename().string();
    }

    Image SequenceReadNode::exec()
    {
        // The clip-local time IS the file number (Read.cpp:169-174).
        const std::string url = getSequenceFrame(_base, _namePrefix, _time.to_frames(), _frameZeroPadding, _nameSuffix);
        MediaFrame frame;
        const auto memory = _memoryReferences.find(url);
        if (!readMedia(_store, url, frame, memory != _memoryReferences.end() ? memory->second : MemoryReference(), false))
        {
            return Image();
        }
        return _toDevice(frame, _store, url);
    }

    std::vector<std::string> SequenceReadNode::getExtensions()
    {
        return { ".exr", ".tif", ".tiff", ".jpg", ".jpeg", ".png" };
    }

    std::shared_ptr<IReadNode> createReadNode(
        const std::filesystem::path& path,
        const std::shared_ptr<MediaStore>& store,
        const MemoryReference& memory)
    {
        return std::make_shared<ImageReadNode>(path, store, memory);
    }

    std::shared_ptr<IReadNode> createReadNode(
        const std::string& base,
        const std::string& namePrefix,
        const std::string& nameSuffix,
        int startFrame,
        int frameStep,
        double rate,
        int frameZeroPadding,
        const std::shared_ptr<MediaStore>& store,
        const MemoryReferences& memoryReferences)
    {
        return std::make_shared<SequenceReadNode>(
            base, namePrefix, nameSuffix, startFrame, frameStep, rate, frameZeroPadding, store, memoryReferences);
    }
}
