#include "FrameScheduler.h"

#include <atomic>

#include "ImageEffectHost.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <condition_variable>
#include <deque>
#include <exception>
#include <mutex>
#include <set>
#include <sstream>
#include <thread>

namespace toucan
{
    bool parseRawFormat(const std::string& name, int& type, int& channels)
    {
        static const struct { const char* name; int type; int channels; } specs[] = {
            { "rgb24", TC_U8, 3 }, { "rgba", TC_U8, 4 }, { "rgb48", TC_U16, 3 }, { "rgba64", TC_U16, 4 },
            { "rgbaf16", TC_F16, 4 }, { "rgbf32", TC_F32, 3 }, { "rgbaf32", TC_F32, 4 } };
        for (const auto& spec : specs)
        {
            if (name == spec.name)
            {
                type = spec.type;
                channels = spec.channels;
                return true;
            }
        }
        return false;
    }

    bool parseY4mFormat(const std::string& name, int& format)
    {
        static const struct { const char* name; int format; } specs[] = {
            { "422", TC_Y4M_422 }, { "444", TC_Y4M_444 }, { "444alpha", TC_Y4M_444ALPHA }, { "444p16", TC_Y4M_444P16 } };
        for (const auto& spec : specs)
        {
            if (name == spec.name)
            {
                format = spec.format;
                return true;
            }
        }
        return false;
    }

    std::string y4mHeader(int width, int height, double rate, const std::string& format)
    {
        // ftk::toRational (App.cpp:323): the common video rates as exact fractions, else (int(rate), 1)
        static const std::pair<int, int> common[] = { { 24, 1 }, { 30, 1 }, { 60, 1 }, { 24000, 1001 }, { 30000, 1001 }, { 60000, 1001 } };
        std::pair<int, int> r(static_cast<int>(rate), 1);
        for (const auto& c : common)
        {
            if (std::fabs(rate - c.first / static_cast<double>(c.second)) < 0.01)
            {
                r = c;
                break;
            }
        }
        std::stringstream ss;
        ss << "YUV4MPEG2 W" << width << " H" << height << " F" << r.first << ":" << r.second << " C" << format << "\n";
        return ss.str();
    }

    std::pair<int, int> frameBlock(int first, int last, int index, int count)
    {
        const int n = last > first ? last - first : 0;
        const int per = count > 0 ? (n + count - 1) / count : n;
        const int lo = std::min(first + index * per, last);
        return { lo, std::min(lo + per, last) };
    }

    namespace
    {
        // Pinned staging buffers are expensive to create (cudaHostAlloc of a 531 MB 8K frame takes
        // ~0.1 s), so slots draw from a process-wide free list that outlives a scheduler run.
        class PinnedPool
        {
        public:
            void* acquire(size_t bytes, size_t& capacity)
            {
                {
                    std::lock_guard<std::mutex> lock(_mutex);
                    for (size_t i = 0; i < _free.size(); ++i)
                    {
                        if (_free[i].second >= bytes)
                        {
                            void* p = _free[i].first;
                            capacity = _free[i].second;
                            _free.erase(_free.begin() + static_cast<long>(i));
                            return p;
                        }
                    }
                }
                void* p = nullptr;
                tcCheck(tc_host_alloc(&p, bytes), "tc_host_alloc");
                capacity = bytes;
                return p;
            }
            void release(void* p, size_t capacity)
            {
                if (!p) return;
                std::lock_guard<std::mutex> lock(_mutex);
                _free.emplace_back(p, capacity);
            }

        private:
            std::mutex _mutex;
            std::vector<std::pair<void*, size_t> > _free;
        };
        PinnedPool pinnedPool;

        struct Slot
        {
            void* pinned = nullptr;
            size_t capacity = 0;
            tc_event done = nullptr;
            Image image; // keeps the device frame alive until its download has been consumed
            ImageSpec spec;
            size_t bytes = 0;
            int frame = -1;
        };

        struct Worker
        {
            int device = 0;
            int first = 0, last = 0, step = 1;
            std::vector<Slot> slots;
            std::mutex mutex;
            std::condition_variable cv;
            std::deque<int> ready;  // slot indices, in frame order
            std::deque<int> free_;  // slot indices the writer has released
            std::exception_ptr error;
            bool finished = false;
            bool writerDone = false; // the writer will not touch this worker's slots again
            std::thread thread;
        };
    }

    FrameScheduler::FrameScheduler(
        const std::shared_ptr<ftk::Context>& context,
        const std::shared_ptr<TimelineWrapper>& timelineWrapper,
        const FrameSchedulerOptions& options) :
        _context(context),
        _timelineWrapper(timelineWrapper),
        _options(options)
    {
        if (_options.devices.empty()) _options.devices = { 0 };
        if (_options.framesInFlight < 1) _options.framesInFlight = 1;
    }

    void FrameScheduler::run(int first, int last, const Writer& writer)
    {
        const OTIO_NS::TimeRange& timeRange = _timelineWrapper->getTimeRange();
        const double rate = timeRange.duration().rate();
        const int count = static_cast<int>(_options.devices.size());
        std::vector<std::unique_ptr<Worker> > workers;
        for (int i = 0; i < count; ++i)
        {
            auto w = std::make_unique<Worker>();
            w->device = _options.devices[static_cast<size_t>(i)];
            if (_options.interleave)
            {
                // frame f belongs to worker (f - first) % count: the ordered writer visits the workers round-robin, so
                // every GPU renders all the time with a ring of only framesInFlight slots
                w->first = first + i;
                w->last = last;
                w->step = count;
            }
            else
            {
                const auto block = frameBlock(first, last, i, count);
                w->first = block.first;
                w->last = block.second;
            }
            w->slots.resize(static_cast<size_t>(_options.framesInFlight));
            for (int s = 0; s < _options.framesInFlight; ++s) w->free_.push_back(s);
            workers.push_back(std::move(w));
        }

        // set on any error, in a worker or in the writer: every worker stops at its next frame
        std::atomic<bool> cancel{ false };

        auto body = [this, rate, &timeRange, &cancel](Worker& w)
        {
            tc_stream copyStream = nullptr;
            tc_event rendered = nullptr;
            try
            {
                DeviceContext::bind(w.device);
                tcCheck(tc_stream_create(&copyStream), "tc_stream_create");
                tcCheck(tc_event_create(&rendered), "tc_event_create");
                for (Slot& slot : w.slots) tcCheck(tc_event_create(&slot.done), "tc_event_create");
                std::shared_ptr<ImageEffectHost> host;
                {
                    // the OpenFX modules keep module-wide state while they load: one host at a time
                    static std::mutex hostMutex;
                    std::lock_guard<std::mutex> lock(hostMutex);
                    host = std::make_shared<ImageEffectHost>(_context, _options.pluginSearchPath);
                }
                auto graph = std::make_shared<ImageGraph>(_context, _timelineWrapper->getPath().parent_path(), _timelineWrapper);
                for (int frame = w.first; frame < w.last && !cancel; frame += w.step)
                {
                    int index = -1;
                    {
                        std::unique_lock<std::mutex> lock(w.mutex);
                        w.cv.wait(lock, [&w, &cancel] { return !w.free_.empty() || cancel; });
                        if (cancel) break;
                        index = w.free_.front();
                        w.free_.pop_front();
                    }
                    Slot& slot = w.slots[static_cast<size_t>(index)];
                    slot.image = Image(); // released here, on the thread bound to the device
                    const OTIO_NS::RationalTime time = timeRange.start_time() + OTIO_NS::RationalTime(frame, rate);
                    if (auto node = graph->exec(host, time))
                    {
                        slot.image = node->exec();
                    }
                    slot.spec = slot.image.spec();
                    slot.frame = frame;
                    void* packed = nullptr;
                    size_t bytes = slot.spec.image_bytes();
                    if (_options.y4mFormat >= 0 && bytes > 0)
                    {
                        // paste + swscale of App.cpp:340-412 as one kernel; the download moves the planar bytes
                        const tc_image view = slot.image.view();
                        bytes = tc_y4m_frame_bytes(_options.y4mFormat, slot.spec.width, slot.spec.height);
                        slot.spec = ImageSpec(slot.spec.width, slot.spec.height, TC_Y4M_444ALPHA == _options.y4mFormat ? 4 : 3,
                            TC_Y4M_444P16 == _options.y4mFormat ? TC_U16 : TC_U8);
                        tcCheck(tc_malloc(&packed, bytes, DeviceContext::stream()), "tc_malloc");
                        tcCheck(tc_rgb_to_y4m(packed, &view, _options.y4mFormat, DeviceContext::stream()), "tc_rgb_to_y4m");
                    }
                    else if (_options.rawType >= 0 && slot.spec.format != _options.rawType && bytes > 0)
                    {
                        // output conversion on the device: the download moves the writer's bytes
                        const tc_image view = slot.image.view();
                        slot.spec = ImageSpec(slot.spec.width, slot.spec.height, _options.rawChannels, _options.rawType);
                        bytes = slot.spec.image_bytes();
                        tcCheck(tc_malloc(&packed, bytes, DeviceContext::stream()), "tc_malloc");
                        tcCheck(tc_pack_raw(packed, &view, _options.rawType, _options.rawChannels, DeviceContext::stream()), "tc_pack_raw");
                    }
                    slot.bytes = bytes;
                    if (bytes > slot.capacity)
                    {
                        pinnedPool.release(slot.pinned, slot.capacity);
                        slot.pinned = nullptr;
                        slot.capacity = 0;
                        slot.pinned = pinnedPool.acquire(bytes, slot.capacity);
                    }
                    if (bytes)
                    {
                        // download on the copy stream so the next frame's uploads and kernels overlap it
                        tcCheck(tc_event_record(rendered, DeviceContext::stream()), "tc_event_record");
                        tcCheck(tc_stream_wait_event(copyStream, rendered), "tc_stream_wait_event");
                        tcCheck(tc_download(slot.pinned, packed ? packed : slot.image.data(), bytes, copyStream), "tc_download");
                    }
                    if (packed)
                    {
                        tcCheck(tc_free(packed, copyStream), "tc_free"); // stream-ordered after the download
                    }
                    tcCheck(tc_event_record(slot.done, copyStream), "tc_event_record");
                    {
                        std::lock_guard<std::mutex> lock(w.mutex);
                        w.ready.push_back(index);
                    }
                    w.cv.notify_all();
                }
            }
            catch (...)
            {
                cancel = true;
                std::lock_guard<std::mutex> lock(w.mutex);
                w.error = std::current_exception();
            }
            w.cv.notify_all();
            // Drain before anything is released: every handed-out slot comes back, or -- after an error anywhere -- the
            // writer has declared it will not touch this worker's slots again; then both streams are idle (no copy is
            // still writing a pinned buffer that goes back to the process-wide pool).
            {
                std::unique_lock<std::mutex> lock(w.mutex);
                w.cv.wait(lock, [&w] { return w.free_.size() == w.slots.size() || w.writerDone; });
            }
            if (copyStream) tc_stream_sync(copyStream);
            if (DeviceContext::isBound()) tc_stream_sync(DeviceContext::stream());
            for (Slot& slot : w.slots)
            {
                slot.image = Image();
                pinnedPool.release(slot.pinned, slot.capacity);
                if (slot.done) tc_event_destroy(slot.done);
                slot.pinned = nullptr;
                slot.done = nullptr;
            }
            if (rendered) tc_event_destroy(rendered);
            if (copyStream) tc_stream_destroy(copyStream);
            DeviceContext::unbind(); // the thread's own stream goes with the thread
            {
                std::lock_guard<std::mutex> lock(w.mutex);
                w.finished = true;
            }
            w.cv.notify_all();
        };

        // Upload ahead: with the device source cache on, one thread per GPU copies the timeline's stills to the device in the order
        // the frames first touch them, on its own stream -- the frame that needs a still finds it cached (or waits for the copy's event
        // on the GPU, never on the host), and the host -> device copies of later clips overlap the device -> host copies of finished
        // frames (PCIe is full duplex; uploading at first use serialised the two directions).
        std::vector<std::thread> uploaders;
        const auto store = _timelineWrapper->getMediaStore();
        if (_options.uploadAhead && store && store->getDeviceCacheBytes() > 0)
        {
            std::set<int> devices(_options.devices.begin(), _options.devices.end());
            const auto stills = _timelineWrapper->getStillMediaByFirstUse(OTIO_NS::RationalTime(first, rate));
            for (const int device : devices)
            {
                uploaders.emplace_back([device, stills, store, &cancel]
                {
                    try
                    {
                        DeviceContext::bind(device);
                        size_t budget = store->getDeviceCacheBytes();
                        for (const auto& still : stills)
                        {
                            if (cancel) break;
                            MediaFrame frame;
                            if (!readMedia(store, still.first, frame, still.second) || frame.device || !frame.host || 4 != frame.nchannels) continue;
                            const size_t bytes = tc_image_bytes(frame.width, frame.height, frame.format);
                            if (bytes > budget) break; // the cache is full: what is needed now must not be evicted for what is needed later
                            budget -= bytes;
                            uploadToCache(frame, store, device, still.first);
                        }
                    }
                    catch (...)
                    {
                        // best effort: a still that cannot be read here fails where the frame that needs it is rendered
                    }
                    DeviceContext::unbind();
                });
            }
        }

        for (auto& w : workers)
        {
            Worker* wp = w.get();
            wp->thread = std::thread([body, wp] { body(*wp); });
        }

        // The writer: frames in timeline order, each from the worker that owns it.
        std::exception_ptr error;
        auto owner = [&](int frame) -> Worker*
        {
            if (_options.interleave)
            {
                return workers[static_cast<size_t>((frame - first) % count)].get();
            }
            for (auto& w : workers)
            {
                if (frame >= w->first && frame < w->last) return w.get();
            }
            return nullptr;
        };
        bool stopped = false;
        for (int frame = first; frame < last && !error && !stopped; ++frame)
        {
            Worker* w = owner(frame);
            if (!w)
            {
                error = std::make_exception_ptr(std::runtime_error("frame scheduler: frame without a worker"));
                break;
            }
            int index = -1;
            {
                std::unique_lock<std::mutex> lock(w->mutex);
                // another worker's failure sets `cancel` without signalling this worker's condition: poll for it
                while (!w->cv.wait_for(lock, std::chrono::milliseconds(20), [w] { return !w->ready.empty() || w->error || w->finished; }) && !cancel) {}
                if (w->error) { error = w->error; break; }
                if (w->ready.empty() && cancel) { stopped = true; break; } // the failing worker's exception is collected after the join
                if (w->ready.empty()) { error = std::make_exception_ptr(std::runtime_error("frame scheduler: worker ended early")); break; }
                index = w->ready.front();
                w->ready.pop_front();
            }
            Slot& slot = w->slots[static_cast<size_t>(index)];
            try
            {
                tcCheck(tc_event_sync(slot.done), "tc_event_sync");
                writer(slot.frame, slot.pinned, slot.spec, slot.bytes);
            }
            catch (...)
            {
                error = std::current_exception();
            }
            {
                std::lock_guard<std::mutex> lock(w->mutex);
                w->free_.push_back(index);
            }
            w->cv.notify_all();
        }
        if (error) cancel = true; // a failed writer (a closed pipe): the workers stop instead of rendering the rest
        for (auto& w : workers)
        {
            // the writer is done with every slot: workers waiting for one (or draining) may tear down
            {
                std::lock_guard<std::mutex> lock(w->mutex);
                w->writerDone = true;
            }
            w->cv.notify_all();
        }
        for (auto& w : workers)
        {
            if (w->thread.joinable()) w->thread.join();
            if (!error && w->error) error = w->error;
        }
        for (auto& u : uploaders)
        {
            if (u.joinable()) u.join();
        }
        if (error)
        {
            std::rethrow_exception(error);
        }
    }
}
