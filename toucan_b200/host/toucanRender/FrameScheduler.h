// Frame scheduler: renders a range of timeline frames on one or more GPUs and hands the
// finished frames to a writer in timeline order.
//
// Replaces the serial loop of bin/toucan-render/App.cpp:222-264 (one frame in flight, one
// node at a time, the writer blocking the renderer).  Frames are independent
// (lib/toucanRender/ImageGraph.cpp:144-215 builds a fresh tree per time), so:
//   * the range is cut into contiguous blocks, one per GPU -- no collective;
//   * each GPU has a worker thread with its own device binding, ImageEffectHost and
//     ImageGraph (ImageGraph::exec is not re-entrant), a render stream and a copy stream;
//   * a finished frame is downloaded on the copy stream into a pinned ring slot while the
//     next frame renders; the calling thread is the writer and receives frames in order.
#pragma once

#include <toucanRender/ImageGraph.h>

#include <functional>

namespace toucan
{
    struct FrameSchedulerOptions
    {
        std::vector<int> devices = { 0 };
        std::vector<std::filesystem::path> pluginSearchPath;
        int framesInFlight = 2; //!< pinned ring slots per GPU
        //! Raw output format (bin/toucan-render/App.cpp:26-35 rawSpecs): when rawType >= 0 and the
        //! frame's storage type differs from it, the frame is converted ON THE DEVICE to
        //! rawChannels x rawType before the download (the reference converts on the host with
        //! ImageBufAlgo::paste, and -- like here -- writes the frame as is when the types match).
        int rawType = -1;
        int rawChannels = 4;
    };

    //! "rgb24" | "rgba" | "rgb48" | "rgba64" | "rgbaf16" | "rgbf32" | "rgbaf32" -> (tc_type, channels)
    bool parseRawFormat(const std::string&, int& type, int& channels);

    //! Contiguous block of frames [first, last) owned by worker `index` of `count`.
    std::pair<int, int> frameBlock(int first, int last, int index, int count);

    class FrameScheduler
    {
    public:
        using Writer = std::function<void(int frame, const void* pixels, const ImageSpec&)>;

        FrameScheduler(
            const std::shared_ptr<ftk::Context>&,
            const std::shared_ptr<TimelineWrapper>&,
            const FrameSchedulerOptions&);

        //! Render frames [first, last) (0 = the timeline's first frame).  `writer` runs on the
        //! calling thread, once per frame, in increasing frame order; `pixels` is pinned host
        //! memory valid for the duration of the call.  Throws what a worker threw.
        void run(int first, int last, const Writer& writer);

    private:
        std::shared_ptr<ftk::Context> _context;
        std::shared_ptr<TimelineWrapper> _timelineWrapper;
        FrameSchedulerOptions _options;
    };
}
