// Device-resident image: what flows between graph nodes in place of the reference's
// OIIO::ImageBuf (lib/toucanRender/ImageNode.h:43 `virtual OIIO::ImageBuf exec()`).
// Frames stay in HBM from upload to the final download; buffers come from the
// stream-ordered pool of libtoucan_cuda.so and are returned to it when the last handle
// goes away (no cudaMalloc / memset per node, unlike ImageEffect.cpp:93,108,126).
#pragma once

#include <toucan_cuda.h>

#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace toucan
{
    //! Per-thread device binding: one CUDA device + stream per render thread
    //! (one thread per GPU in the frame scheduler).
    class DeviceContext
    {
    public:
        //! Bind the calling thread to a device. stream == nullptr creates one.
        static void bind(int device, tc_stream stream = nullptr);
        //! Wait for the thread's stream and destroy it if bind() created it (end of a worker thread).
        static void unbind();
        static bool isBound();
        static int device();
        static tc_stream stream();
        //! Whether bind() created the thread's stream (it is destroyed with the binding) or the caller provided it.
        static bool streamIsOwned();
        static void synchronize();
    };

    //! Image description (the subset of OIIO::ImageSpec used on the path).
    struct ImageSpec
    {
        ImageSpec() = default;
        ImageSpec(int w, int h, int nch, int fmt = TC_U8) : width(w), height(h), nchannels(nch), format(fmt) {}
        int width = 0;
        int height = 0;
        int nchannels = 4;
        int format = TC_U8; // tc_type; ImageSpec(w,h,n) without a type is UINT8 like OIIO
        size_t pixel_bytes() const { return static_cast<size_t>(nchannels) * tc_type_size(format); }
        size_t scanline_bytes() const { return pixel_bytes() * static_cast<size_t>(width); }
        size_t image_bytes() const { return scanline_bytes() * static_cast<size_t>(height > 0 ? height : 0); }
    };

    //! Get a label for a pixel type: "u8", "u16", "half", "float".
    std::string toImageDataType(int format);

    class Image
    {
    public:
        Image() = default;
        //! Allocate an uninitialised device image on the calling thread's stream.
        explicit Image(const ImageSpec&);
        //! Wrap device memory owned by someone else (resident sources).
        static Image wrap(void* devicePtr, const ImageSpec&);
        //! Upload host pixels (tightly packed) on the calling thread's stream.
        static Image upload(const void* hostPixels, const ImageSpec&);

        //! Tie the lifetime of the host memory an upload reads to this image: uploads from page-locked memory are
        //! asynchronous, so the source must outlive the copy (it is released with the device buffer).
        void keepAlive(const std::shared_ptr<const void>& hostSource);

        //! Sharing one image between render threads (the device source cache): every thread has its own stream.
        //! publish(): record "contents complete" on the calling thread's stream -- call after the work that fills the
        //! image was enqueued, before other threads can see it.  acquire(): make the calling thread's stream wait for
        //! that point; a buffer acquired from a second stream is released with a device-wide wait instead of a
        //! stream-ordered free, because its readers are no longer ordered by the allocating stream alone.
        void publish();
        void acquire() const;

        const ImageSpec& spec() const { return _spec; }
        bool initialized() const { return _data != nullptr; }
        void* data() const { return _data; }
        tc_image view() const { return tc_image{ _data, _spec.width, _spec.height, _spec.format }; }

        //! Copy to host memory (blocks until the frame is complete).
        void download(void* hostPixels) const;
        std::vector<unsigned char> download() const;

    private:
        struct Buffer;
        ImageSpec _spec;
        void* _data = nullptr;
        std::shared_ptr<Buffer> _buffer;
    };

    //! Throw std::runtime_error with tc_last_error() if a C-ABI call failed.
    void tcCheck(int status, const char* what);
}
