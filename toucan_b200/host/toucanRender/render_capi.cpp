// C ABI over the C++ host (include/toucan_render.h).
#include <toucan_render.h>

#include "ImageEffectHost.h"
#include "FrameScheduler.h"
#include "ImageGraph.h"
#include "ExrRead.h"
#include "JpegRead.h"
#include "PngRead.h"
#include "TiffRead.h"
#include "StackFusion.h"
#include "TimelineWrapper.h"
#include "Util.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <sstream>

using namespace toucan;

struct tr_host
{
    std::shared_ptr<ftk::Context> context;
    std::shared_ptr<ImageEffectHost> host;
    std::vector<std::string> identifiers;
};

struct tr_timeline
{
    std::shared_ptr<ftk::Context> context;
    std::shared_ptr<MediaStore> store;
    std::shared_ptr<TimelineWrapper> wrapper;
    std::shared_ptr<ImageGraph> graph;
    Image resident; // the frame handed out by tr_timeline_render_resident
};

namespace
{
    std::atomic<bool> interleaveFrames{ true };
    thread_local std::string lastError;

    int fail(const std::string& what)
    {
        lastError = what;
        return -1;
    }

    template <class F>
    int guarded(F&& f)
    {
        try
        {
            return f();
        }
        catch (const std::exception& e)
        {
            return fail(e.what());
        }
    }

    std::filesystem::path libraryDirectory()
    {
        Dl_info info;
        if (dladdr(reinterpret_cast<const void*>(&tr_last_error), &info) && info.dli_fname)
        {
            return std::filesystem::path(info.dli_fname).parent_path();
        }
        return std::filesystem::current_path();
    }

    int copyOut(const std::string& s, char* out, size_t capacity)
    {
        if (!out || capacity < s.size() + 1)
        {
            return fail("output buffer too small");
        }
        memcpy(out, s.c_str(), s.size() + 1);
        return 0;
    }

    int ensureGraph(tr_timeline* t)
    {
        if (!t->graph)
        {
            t->graph = std::make_shared<ImageGraph>(t->context, t->wrapper->getPath().parent_path(), t->wrapper);
        }
        return 0;
    }

    OTIO_NS::RationalTime timeOf(const tr_timeline* t, double value)
    {
        return OTIO_NS::RationalTime(value, t->wrapper->getTimeRange().duration().rate());
    }

    Image renderImage(tr_timeline* t, tr_host* h, double value)
    {
        ensureGraph(t);
        auto node = t->graph->exec(h->host, timeOf(t, value));
        return node ? node->exec() : Image();
    }
}

extern "C"
{
    const char* tr_last_error(void) { return lastError.c_str(); }

    int tr_bind_device(int device)
    {
        return guarded([&] { DeviceContext::bind(device); return 0; });
    }

    int tr_bind_stream(int device, void* stream)
    {
        return guarded([&] { DeviceContext::bind(device, stream); return 0; });
    }

    int tr_sync(void)
    {
        return guarded([&] { DeviceContext::synchronize(); return 0; });
    }

    int tr_host_create(const char* plugin_dir, tr_host** out)
    {
        if (!out) return fail("tr_host_create: null");
        *out = nullptr;
        return guarded([&] {
            auto h = std::make_unique<tr_host>();
            h->context = ftk::Context::create();
            std::vector<std::filesystem::path> searchPath;
            if (plugin_dir && *plugin_dir)
            {
                searchPath.push_back(plugin_dir);
            }
            else
            {
                searchPath = getOpenFXPluginPaths(libraryDirectory() / "toucan-render");
            }
            h->host = std::make_shared<ImageEffectHost>(h->context, searchPath);
            h->identifiers = h->host->getPluginIdentifiers();
            *out = h.release();
            return 0;
        });
    }

    void tr_host_destroy(tr_host* host) { delete host; }

    int tr_host_plugin_count(const tr_host* host) { return host ? static_cast<int>(host->identifiers.size()) : 0; }

    const char* tr_host_plugin_identifier(const tr_host* host, int index)
    {
        if (!host || index < 0 || index >= static_cast<int>(host->identifiers.size())) return "";
        return host->identifiers[static_cast<size_t>(index)].c_str();
    }

    int tr_timeline_open(const char* otio_path, tr_timeline** out)
    {
        if (!out || !otio_path) return fail("tr_timeline_open: null");
        *out = nullptr;
        return guarded([&] {
            auto t = std::make_unique<tr_timeline>();
            t->context = ftk::Context::create();
            t->store = std::make_shared<MediaStore>();
            t->wrapper = std::make_shared<TimelineWrapper>(std::filesystem::path(otio_path));
            t->wrapper->setMediaStore(t->store);
            *out = t.release();
            return 0;
        });
    }

    int tr_timeline_open_json(const char* json, const char* directory, tr_timeline** out)
    {
        if (!out || !json) return fail("tr_timeline_open_json: null");
        *out = nullptr;
        return guarded([&] {
            auto t = std::make_unique<tr_timeline>();
            t->context = ftk::Context::create();
            t->store = std::make_shared<MediaStore>();
            t->wrapper = std::make_shared<TimelineWrapper>(std::string(json), std::filesystem::path(directory ? directory : ""));
            t->wrapper->setMediaStore(t->store);
            *out = t.release();
            return 0;
        });
    }

    void tr_timeline_close(tr_timeline* timeline) { delete timeline; }

    int tr_timeline_media_path(const tr_timeline* timeline, const char* url, char* out, size_t capacity)
    {
        if (!timeline || !url) return fail("tr_timeline_media_path: null");
        return guarded([&] { return copyOut(timeline->wrapper->getMediaPath(url), out, capacity); });
    }

    int tr_timeline_set_media(tr_timeline* timeline, const char* path, const void* pixels, int device, int width, int height, int type)
    {
        if (!timeline || !path || !pixels) return fail("tr_timeline_set_media: null");
        if (width <= 0 || height <= 0 || type < TC_U8 || type > TC_F32) return fail("tr_timeline_set_media: bad frame description");
        MediaFrame frame;
        if (device) frame.device = const_cast<void*>(pixels);
        else frame.host = pixels;
        frame.width = width;
        frame.height = height;
        frame.nchannels = 4;
        frame.format = type;
        timeline->store->set(path, frame);
        return 0;
    }

    int tr_timeline_set_source_cache(tr_timeline* timeline, size_t bytes)
    {
        if (!timeline) return fail("tr_timeline_set_source_cache: null");
        timeline->store->setDeviceCacheBytes(bytes);
        return 0;
    }

    int tr_timeline_upload_stats(const tr_timeline* timeline, unsigned long long* uploads, unsigned long long* bytes)
    {
        if (!timeline) return fail("tr_timeline_upload_stats: null");
        if (uploads) *uploads = timeline->store->getUploadCount();
        if (bytes) *bytes = timeline->store->getUploadBytes();
        return 0;
    }

    int tr_timeline_drop_source_cache(tr_timeline* timeline)
    {
        if (!timeline) return fail("tr_timeline_drop_source_cache: null");
        return guarded([&] { timeline->store->dropCached(); return 0; });
    }

    int tr_timeline_prepare(tr_timeline* timeline)
    {
        if (!timeline) return fail("tr_timeline_prepare: null");
        return guarded([&] { return ensureGraph(timeline); });
    }

    int tr_timeline_info(const tr_timeline* timeline, double* start_value, double* rate, int* frames, int* width, int* height, int* type)
    {
        if (!timeline) return fail("tr_timeline_info: null");
        return guarded([&] {
            const OTIO_NS::TimeRange& range = timeline->wrapper->getTimeRange();
            if (start_value) *start_value = range.start_time().value();
            if (rate) *rate = range.duration().rate();
            if (frames)
            {
                // the frame loop of App.cpp:222-224
                int n = 0;
                const OTIO_NS::RationalTime inc(1.0, range.duration().rate());
                for (OTIO_NS::RationalTime t = range.start_time(); t <= range.end_time_inclusive(); t += inc) ++n;
                *frames = n;
            }
            if (timeline->graph)
            {
                if (width) *width = timeline->graph->getImageSize().x;
                if (height) *height = timeline->graph->getImageSize().y;
                if (type)
                {
                    const std::string& s = timeline->graph->getImageDataType();
                    *type = s == "u8" ? TC_U8 : s == "u16" ? TC_U16 : s == "half" ? TC_F16 : s == "float" ? TC_F32 : -1;
                }
            }
            return 0;
        });
    }

    int tr_timeline_describe(tr_timeline* timeline, tr_host* host, double time_value, char* out, size_t capacity)
    {
        if (!timeline || !host) return fail("tr_timeline_describe: null");
        return guarded([&] {
            ensureGraph(timeline);
            auto node = timeline->graph->exec(host->host, timeOf(timeline, time_value));
            std::vector<std::string> lines;
            if (node) node->describeTree(lines);
            std::stringstream ss;
            for (const auto& line : lines) ss << line << "\n";
            return copyOut(ss.str(), out, capacity);
        });
    }

    int tr_timeline_render(tr_timeline* timeline, tr_host* host, double time_value, void* out, size_t capacity, int* width, int* height, int* type)
    {
        if (!timeline || !host || !out) return fail("tr_timeline_render: null");
        return guarded([&] {
            const Image image = renderImage(timeline, host, time_value);
            const ImageSpec& spec = image.spec();
            if (width) *width = spec.width;
            if (height) *height = spec.height;
            if (type) *type = spec.format;
            if (spec.image_bytes() > capacity) return fail("tr_timeline_render: output buffer too small");
            image.download(out);
            return 0;
        });
    }

    int tr_timeline_render_resident(tr_timeline* timeline, tr_host* host, double time_value, void** out, int* width, int* height, int* type)
    {
        if (!timeline || !host || !out) return fail("tr_timeline_render_resident: null");
        return guarded([&] {
            timeline->resident = renderImage(timeline, host, time_value);
            const ImageSpec& spec = timeline->resident.spec();
            *out = timeline->resident.data();
            if (width) *width = spec.width;
            if (height) *height = spec.height;
            if (type) *type = spec.format;
            return 0;
        });
    }

    int tr_render_range(tr_timeline* timeline, const char* plugin_dir, const int* devices, int n_devices, int first, int last,
                        int frames_in_flight, tr_frame_callback callback, void* user)
    {
        return tr_render_range_raw(timeline, plugin_dir, devices, n_devices, first, last, frames_in_flight, nullptr, callback, user);
    }

    int tr_render_range_raw(tr_timeline* timeline, const char* plugin_dir, const int* devices, int n_devices, int first, int last,
                            int frames_in_flight, const char* raw_format, tr_frame_callback callback, void* user)
    {
        if (!timeline || !callback) return fail("tr_render_range: null");
        return guarded([&] {
            FrameSchedulerOptions options;
            if (raw_format && *raw_format && !parseRawFormat(raw_format, options.rawType, options.rawChannels))
            {
                return fail("Cannot find the given raw format");
            }
            options.devices.clear();
            for (int i = 0; i < n_devices; ++i) options.devices.push_back(devices ? devices[i] : i);
            if (options.devices.empty()) options.devices.push_back(0);
            if (plugin_dir && *plugin_dir) options.pluginSearchPath.push_back(plugin_dir);
            else options.pluginSearchPath = getOpenFXPluginPaths(libraryDirectory() / "toucan-render");
            if (frames_in_flight > 0) options.framesInFlight = frames_in_flight;
            options.interleave = interleaveFrames;
            FrameScheduler scheduler(timeline->context, timeline->wrapper, options);
            scheduler.run(first, last, [&](int frame, const void* pixels, const ImageSpec& spec, size_t) {
                callback(user, frame, pixels, spec.width, spec.height, spec.format | (spec.nchannels == 3 ? 0x100 : 0));
            });
            return 0;
        });
    }

    int tr_render_range_y4m(tr_timeline* timeline, const char* plugin_dir, const int* devices, int n_devices, int first, int last,
                            int frames_in_flight, const char* y4m_format, tr_frame_callback callback, void* user)
    {
        if (!timeline || !callback || !y4m_format) return fail("tr_render_range_y4m: null");
        return guarded([&] {
            FrameSchedulerOptions options;
            if (!parseY4mFormat(y4m_format, options.y4mFormat))
            {
                return fail("Cannot find the given y4m format");
            }
            options.devices.clear();
            for (int i = 0; i < n_devices; ++i) options.devices.push_back(devices ? devices[i] : i);
            if (options.devices.empty()) options.devices.push_back(0);
            if (plugin_dir && *plugin_dir) options.pluginSearchPath.push_back(plugin_dir);
            else options.pluginSearchPath = getOpenFXPluginPaths(libraryDirectory() / "toucan-render");
            if (frames_in_flight > 0) options.framesInFlight = frames_in_flight;
            options.interleave = interleaveFrames;
            const int format = options.y4mFormat;
            FrameScheduler scheduler(timeline->context, timeline->wrapper, options);
            scheduler.run(first, last, [&](int frame, const void* pixels, const ImageSpec& spec, size_t) {
                callback(user, frame, pixels, spec.width, spec.height, 0x200 | format);
            });
            return 0;
        });
    }

    int tr_y4m_header(int width, int height, double rate, const char* y4m_format, char* out, size_t capacity)
    {
        if (!y4m_format || !out) return fail("tr_y4m_header: null");
        return guarded([&] {
            const std::string header = y4mHeader(width, height, rate, y4m_format);
            if (header.size() + 1 > capacity) return fail("tr_y4m_header: buffer too small");
            memcpy(out, header.c_str(), header.size() + 1);
            return 0;
        });
    }

    void tr_fit(int aw, int ah, int bw, int bh, int out[4])
    {
        const IMATH_NAMESPACE::Box2i box = toucan::fit(IMATH_NAMESPACE::V2i(aw, ah), IMATH_NAMESPACE::V2i(bw, bh));
        out[0] = box.min.x;
        out[1] = box.min.y;
        out[2] = box.max.x - box.min.x + 1;
        out[3] = box.max.y - box.min.y + 1;
    }

    void tr_frame_block(int first, int last, int index, int count, int* lo, int* hi)
    {
        const auto block = frameBlock(first, last, index, count);
        if (lo) *lo = block.first;
        if (hi) *hi = block.second;
    }

    void tr_set_unassociated_alpha(int keep) { setKeepUnassociatedAlpha(keep != 0); }
    int tr_get_unassociated_alpha(void) { return getKeepUnassociatedAlpha() ? 1 : 0; }

    int tr_decode_png(const char* path, void* out, size_t capacity, int* width, int* height, int* type)
    {
        if (!path) return fail("tr_decode_png: null");
        return guarded([&] {
            MediaFrame frame;
            if (!readPng(path, frame)) return fail(std::string("tr_decode_png: cannot read ") + path);
            if (width) *width = frame.width;
            if (height) *height = frame.height;
            if (type) *type = frame.format;
            const size_t bytes = tc_image_bytes(frame.width, frame.height, frame.format);
            if (out)
            {
                if (bytes > capacity) return fail("tr_decode_png: output buffer too small");
                memcpy(out, frame.host, bytes);
            }
            return 0;
        });
    }

    int tr_decode_jpeg(const char* path, void* out, size_t capacity, int* width, int* height, int* type)
    {
        if (!path) return fail("tr_decode_jpeg: null");
        return guarded([&] {
            MediaFrame frame;
            if (!readJpeg(path, frame)) return fail(std::string("tr_decode_jpeg: cannot read ") + path);
            if (width) *width = frame.width;
            if (height) *height = frame.height;
            if (type) *type = frame.format;
            const size_t bytes = tc_image_bytes(frame.width, frame.height, frame.format);
            if (out)
            {
                if (bytes > capacity) return fail("tr_decode_jpeg: output buffer too small");
                memcpy(out, frame.host, bytes);
            }
            return 0;
        });
    }

    int tr_decode_tiff(const char* path, void* out, size_t capacity, int* width, int* height, int* type)
    {
        if (!path) return fail("tr_decode_tiff: null");
        return guarded([&] {
            MediaFrame frame;
            if (!readTiff(path, frame)) return fail(std::string("tr_decode_tiff: cannot read ") + path);
            if (width) *width = frame.width;
            if (height) *height = frame.height;
            if (type) *type = frame.format;
            const size_t bytes = tc_image_bytes(frame.width, frame.height, frame.format);
            if (out)
            {
                if (bytes > capacity) return fail("tr_decode_tiff: output buffer too small");
                memcpy(out, frame.host, bytes);
            }
            return 0;
        });
    }

    int tr_write_exr(const char* path, const void* pixels, int width, int height, int type, int zip)
    {
        if (!path || !pixels) return fail("tr_write_exr: null");
        return guarded([&] {
            if (!writeExr(path, pixels, width, height, type, zip != 0)) return fail(std::string("tr_write_exr: cannot write ") + path);
            return 0;
        });
    }

    int tr_decode_exr(const char* path, void* out, size_t capacity, int* width, int* height, int* type)
    {
        if (!path) return fail("tr_decode_exr: null");
        return guarded([&] {
            MediaFrame frame;
            if (!readExr(path, frame)) return fail(std::string("tr_decode_exr: cannot read ") + path);
            if (width) *width = frame.width;
            if (height) *height = frame.height;
            if (type) *type = frame.format;
            const size_t bytes = tc_image_bytes(frame.width, frame.height, frame.format);
            if (out)
            {
                if (bytes > capacity) return fail("tr_decode_exr: output buffer too small");
                memcpy(out, frame.host, bytes);
            }
            return 0;
        });
    }

    void tr_set_fusion(int enabled) { setStackFusionEnabled(enabled != 0); }

    void tr_set_interleave(int enabled) { interleaveFrames = enabled != 0; }

    void tr_set_node_timers(int enabled) { NodeTimers::setEnabled(enabled != 0); }

    size_t tr_node_timers_report(char* out, size_t capacity)
    {
        const std::string text = NodeTimers::report();
        if (out && capacity > 0)
        {
            const size_t n = std::min(capacity - 1, text.size());
            memcpy(out, text.data(), n);
            out[n] = 0;
        }
        return text.size() + 1;
    }
}
