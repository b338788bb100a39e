// Frame scheduler: renders a range of timeline frames on one or more GPUs and hands the
// finished frames to a writer in timeline order.
//
// Replaces the serial loop of bin/toucan-render/App.cpp:222-264 (one frame in flight, one
// node at a time, the writer blocking the renderer).  Frames are independent
// (lib/toucanRender/ImageGraph.cpp:144-215 builds a fresh tree per time), so:
//   * frames are dealt to the GPUs round-robin (or in contiguous blocks, SURVEY 8e, for writers that do not need
//     timeline order across GPUs) -- no collective;
//   * each GPU has a worker thread with its own device binding, ImageEffectHost and
//     ImageGraph (ImageGraph::exec is not re-entrant), a render stream and a copy stream;
//   * a finished frame is downloaded on the copy stream into a pinned ring slot while the
//     next frame renders; the calling thread is the writer and receives frames in order.
#pragma once

#include <cstdlib>
#include <toucanRender/ImageGraph.h>

#include <functional>

namespace toucan
{
    struct FrameSchedulerOptions
    {
        std::vector<int> devices = { 0 };
        std::vector<std::filesystem::path> pluginSearchPath;
        int framesInFlight = 4; //!< pinned ring slots per GPU (behind one ordered writer, two slots leave ~30 % of four GPUs idle)
        //! Frame -> worker assignment.  Interleaved (frame f to worker f % G): the ordered writer drains the workers
        //! round-robin, so all GPUs render concurrently with small rings.  Contiguous blocks (frameBlock) keep each
        //! worker on consecutive frames -- what separate processes use (bench.py, one rank per GPU, no shared writer) --
        //! but behind ONE ordered writer a worker can only run framesInFlight frames ahead of its turn.
        bool interleave = true;
        //! Raw output format (bin/toucan-render/App.cpp:26-35 rawSpecs): when rawType >= 0 and the
        //! frame's storage type differs from it, the frame is converted ON THE DEVICE to
        //! rawChannels x rawType before the download (the reference converts on the host with
        //! ImageBufAlgo::paste, and -- like here -- writes the frame as is when the types match).
        int rawType = -1;
        int rawChannels = 4;
        //! y4m output format (bin/toucan-render/App.cpp:37-43 y4mSpecs; TC_Y4M_*), -1 = off.  Every frame is
        //! converted ON THE DEVICE to the planar YUV bytes the reference's paste + libswscale chain
        //! (App.cpp:340-485) writes after "FRAME\n"; only those bytes are downloaded.
        int y4mFormat = -1;
        //! With the device source cache on, copy the timeline's stills to each GPU ahead of the frames that read them (one
        //! thread and stream per GPU, first-use order).  TOUCAN_NO_UPLOAD_AHEAD=1 turns it off.
        bool uploadAhead = getenv("TOUCAN_NO_UPLOAD_AHEAD") == nullptr;
    };

    //! "rgb24" | "rgba" | "rgb48" | "rgba64" | "rgbaf16" | "rgbf32" | "rgbaf32" -> (tc_type, channels)
    bool parseRawFormat(const std::string&, int& type, int& channels);

    //! "422" | "444" | "444alpha" | "444p16" -> TC_Y4M_*
    bool parseY4mFormat(const std::string&, int& format);

    //! The stream header of App.cpp:302-338: "YUV4MPEG2 W<w> H<h> F<num>:<den> C<format>\n".
    std::string y4mHeader(int width, int height, double rate, const std::string& format);

    //! Contiguous block of frames [first, last) owned by worker `index` of `count`.
    std::pair<int, int> frameBlock(int first, int last, int index, int count);

    class FrameScheduler
    {
    public:
        //! `bytes` is the size of `pixels`: spec.image_bytes() for interleaved frames, the planar frame size in y4m mode
        //! (where spec carries the frame size and the sample type only).
        using Writer = std::function<void(int frame, const void* pixels, const ImageSpec&, size_t bytes)>;

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
