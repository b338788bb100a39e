#include "Image.h"

#include <atomic>
#include <mutex>
#include <set>
#include <sstream>

namespace toucan
{
    namespace
    {
        struct Binding
        {
            bool bound = false;
            int device = 0;
            tc_stream stream = nullptr;
            bool ownStream = false;
        };
        thread_local Binding binding;

        // Streams created by bind() that are still alive.  A buffer can outlive the worker thread that allocated it (a cached
        // source frame): its release must not touch a destroyed stream handle.
        std::mutex ownedStreamsMutex;
        std::set<tc_stream> ownedStreams;

        bool streamIsGone(tc_stream stream, bool owned)
        {
            if (!owned) return false; // a caller-provided stream (tr_bind_stream): the caller keeps it alive
            std::lock_guard<std::mutex> lock(ownedStreamsMutex);
            return 0 == ownedStreams.count(stream);
        }

        void destroyOwnedStream(tc_stream stream)
        {
            {
                std::lock_guard<std::mutex> lock(ownedStreamsMutex);
                ownedStreams.erase(stream);
            }
            tc_stream_destroy(stream);
        }
    }

    void tcCheck(int status, const char* what)
    {
        if (status != TC_OK)
        {
            std::stringstream ss;
            ss << what << ": " << tc_last_error() << " (status " << status << ")";
            throw std::runtime_error(ss.str());
        }
    }

    void DeviceContext::bind(int device, tc_stream stream)
    {
        tcCheck(tc_set_device(device), "tc_set_device");
        if (binding.bound && binding.ownStream && binding.stream)
        {
            tc_stream_sync(binding.stream);
            destroyOwnedStream(binding.stream);
        }
        binding.device = device;
        binding.ownStream = false;
        binding.stream = stream;
        if (!stream)
        {
            tcCheck(tc_stream_create(&binding.stream), "tc_stream_create");
            binding.ownStream = true;
            std::lock_guard<std::mutex> lock(ownedStreamsMutex);
            ownedStreams.insert(binding.stream);
        }
        binding.bound = true;
    }

    void DeviceContext::unbind()
    {
        if (binding.bound && binding.stream)
        {
            tc_stream_sync(binding.stream);
            if (binding.ownStream) destroyOwnedStream(binding.stream);
        }
        binding = Binding();
    }

    bool DeviceContext::isBound() { return binding.bound; }

    int DeviceContext::device()
    {
        if (!binding.bound) bind(0);
        return binding.device;
    }

    tc_stream DeviceContext::stream()
    {
        if (!binding.bound) bind(0);
        return binding.stream;
    }

    bool DeviceContext::streamIsOwned()
    {
        return binding.bound && binding.ownStream;
    }

    void DeviceContext::synchronize()
    {
        tcCheck(tc_stream_sync(stream()), "tc_stream_sync");
    }

    std::string toImageDataType(int format)
    {
        switch (format)
        {
        case TC_U8: return "u8";
        case TC_U16: return "u16";
        case TC_F16: return "half";
        case TC_F32: return "float";
        default: break;
        }
        return "Unknown";
    }

    struct Image::Buffer
    {
        void* p = nullptr;
        tc_stream stream = nullptr;
        bool streamOwned = false;               // `stream` was created by DeviceContext::bind (it ends with its thread)
        std::shared_ptr<const void> hostSource; // page-locked host memory an asynchronous upload is still reading
        tc_event ready = nullptr;               // publish(): the contents are complete on `stream` at this point
        std::atomic<bool> shared{ false };      // acquire()d from another stream: readers are not ordered by `stream`
        ~Buffer()
        {
            // Stream-ordered release when only the allocating stream ever touched the buffer.  A buffer other streams
            // have read (cached sources), or whose stream / device binding is gone (a cached frame outliving its
            // worker), is released with cudaFree, which waits for all work on the device.
            if (p && (shared.load() || streamIsGone(stream, streamOwned) || tc_free(p, stream) != TC_OK)) tc_free_sync(p);
            if (ready) tc_event_destroy(ready);
        }
    };

    Image::Image(const ImageSpec& spec) : _spec(spec)
    {
        const size_t bytes = spec.image_bytes();
        if (bytes > 0)
        {
            if (spec.nchannels != 4)
            {
                throw std::runtime_error("toucan::Image: only 4-channel (RGBA) device images are supported");
            }
            _buffer = std::make_shared<Buffer>();
            _buffer->stream = DeviceContext::stream();
            _buffer->streamOwned = DeviceContext::streamIsOwned();
            tcCheck(tc_malloc(&_buffer->p, bytes, _buffer->stream), "tc_malloc");
            _data = _buffer->p;
        }
    }

    Image Image::wrap(void* devicePtr, const ImageSpec& spec)
    {
        Image out;
        out._spec = spec;
        out._data = devicePtr;
        return out;
    }

    Image Image::upload(const void* hostPixels, const ImageSpec& spec)
    {
        Image out(spec);
        if (out._data)
        {
            tcCheck(tc_upload(out._data, hostPixels, spec.image_bytes(), DeviceContext::stream()), "tc_upload");
        }
        return out;
    }

    void Image::keepAlive(const std::shared_ptr<const void>& hostSource)
    {
        if (_buffer) _buffer->hostSource = hostSource;
    }

    void Image::publish()
    {
        if (!_buffer) return;
        if (!_buffer->ready) tcCheck(tc_event_create(&_buffer->ready), "tc_event_create");
        tcCheck(tc_event_record(_buffer->ready, DeviceContext::stream()), "tc_event_record");
    }

    void Image::acquire() const
    {
        if (!_buffer || !_buffer->ready) return;
        const tc_stream mine = DeviceContext::stream();
        if (mine == _buffer->stream) return; // same stream: already ordered
        _buffer->shared = true;
        tcCheck(tc_stream_wait_event(mine, _buffer->ready), "tc_stream_wait_event");
    }

    void Image::download(void* hostPixels) const
    {
        if (_data)
        {
            tcCheck(tc_download(hostPixels, _data, _spec.image_bytes(), DeviceContext::stream()), "tc_download");
            DeviceContext::synchronize();
        }
    }

    std::vector<unsigned char> Image::download() const
    {
        std::vector<unsigned char> out(_spec.image_bytes());
        download(out.data());
        return out;
    }
}
