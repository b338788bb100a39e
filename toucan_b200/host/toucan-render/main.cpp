// toucan-render: render a timeline to raw frames.  Mirrors bin/toucan-render/App.cpp:152-265
// for the in-scope outputs: raw frames in the timeline's working format to a file or stdout
// ("-").  The serial frame loop of App.cpp:222-264 is replaced by toucan::FrameScheduler:
// frames are sharded in contiguous blocks over the GPUs and written in order.  Both output
// conversions of App.cpp:267-485 (-raw: paste to the raw spec; -y4m: paste + libswscale to
// planar YUV) run on the device, so only the writer's bytes cross PCIe.  An output name ending in
// .png or .exr writes a numbered image sequence like App.cpp:243-252 (getSequenceFrame of the
// name's prefix, start number and padding): .png from 8/16-bit frames (wider frames are converted to
// 16 bit on the device), .exr from half / float frames (8/16-bit frames are converted to half).
// Movie encoders and the other OIIO writers are outside the hot path ("Unsupported output").
//
//   toucan-render input.otio output|- [-raw rgb24|rgba|rgb48|rgba64|rgbaf16|rgbf32|rgbaf32]
//                 [-y4m 422|444|444alpha|444p16] [-gpus N] [-workers K] [-cache GiB] [-inflight N] [-v] [-unassociated_alpha]
//                 [-print_start] [-print_duration] [-print_rate] [-print_size]
#include <toucanRender/ExrRead.h>
#include <toucanRender/FrameScheduler.h>
#include <toucanRender/ImageNode.h>
#include <toucanRender/PngRead.h>
#include <toucanRender/TimelineWrapper.h>
#include <toucanRender/Util.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>

using namespace toucan;

int main(int argc, char** argv)
{
    try
    {
        std::string input, output, raw, y4m;
        int gpus = 1;
        double cacheGiB = 64.0; // device source cache per GPU (a B200 has 180 GB): stills and sequence frames are uploaded once
        int inflight = 0; // pinned ring slots per worker (0 = the scheduler's default)
        int workers = 1; // render threads per GPU: small frames are bound by the host-side graph build, which overlaps across workers
        bool printStart = false, printDuration = false, printRate = false, printSize = false, verbose = false;
        int firstFrame = 0, lastFrame = -1; // -frames a:b: frame indices counted from the timeline's start, both inclusive
        for (int i = 1; i < argc; ++i)
        {
            const std::string a = argv[i];
            if (a == "-gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
            else if (a == "-workers" && i + 1 < argc) workers = std::atoi(argv[++i]);
            else if (a == "-cache" && i + 1 < argc) cacheGiB = std::atof(argv[++i]);
            else if (a == "-inflight" && i + 1 < argc) inflight = std::atoi(argv[++i]);
            else if (a == "-raw" && i + 1 < argc) raw = argv[++i];
            else if (a == "-y4m" && i + 1 < argc) y4m = argv[++i];
            else if (a == "-frames" && i + 1 < argc)
            {
                const std::string r = argv[++i];
                const size_t colon = r.find(':');
                firstFrame = std::atoi(r.substr(0, colon).c_str());
                lastFrame = colon == std::string::npos ? firstFrame : (colon + 1 < r.size() ? std::atoi(r.substr(colon + 1).c_str()) : -1);
            }
            else if (a == "-timers") NodeTimers::setEnabled(true);
            else if (a == "-v") verbose = true;
            else if (a == "-unassociated_alpha") setKeepUnassociatedAlpha(true); // OIIO's "oiio:UnassociatedAlpha" reader hint

            else if (a == "-print_start") printStart = true;
            else if (a == "-print_duration") printDuration = true;
            else if (a == "-print_rate") printRate = true;
            else if (a == "-print_size") printSize = true;
            else if (input.empty()) input = a;
            else if (output.empty()) output = a;
        }
        if (input.empty() || (output.empty() && !(printStart || printDuration || printRate || printSize)))
        {
            std::cerr << "usage: toucan-render input.otio output.raw|- [-raw format | -y4m format] [-gpus N] [-workers K] [-cache GiB] [-frames a:b] [-timers] [-v] [-unassociated_alpha] [-print_start|-print_duration|-print_rate|-print_size]" << std::endl;
            return 1;
        }

        auto context = ftk::Context::create();
        auto timelineWrapper = std::make_shared<TimelineWrapper>(std::filesystem::path(input));
        const OTIO_NS::TimeRange& timeRange = timelineWrapper->getTimeRange();
        const OTIO_NS::RationalTime timeInc(1.0, timeRange.duration().rate());

        if (printStart) { std::cout << timeRange.start_time().value() << std::endl; return 0; }
        if (printDuration) { std::cout << timeRange.duration().value() << std::endl; return 0; }
        if (printRate) { std::cout << timeRange.duration().rate() << std::endl; return 0; }
        if (printSize)
        {
            ImageGraph graph(context, std::filesystem::path(input).parent_path(), timelineWrapper);
            std::cout << graph.getImageSize().x << "x" << graph.getImageSize().y << std::endl;
            return 0;
        }

        // the frame loop of App.cpp:222-224: start .. end_time_inclusive in steps of 1/rate
        int frames = 0;
        for (OTIO_NS::RationalTime time = timeRange.start_time(); time <= timeRange.end_time_inclusive(); time += timeInc) ++frames;

        // output kind: "-" / anything else = raw stream, *.png / *.exr = numbered image sequence
        const std::filesystem::path outputPath(output);
        const std::string outputExtension = toLower(outputPath.extension().string());
        const bool stream = "-" == output || !raw.empty() || !y4m.empty(); // -raw / -y4m: a byte stream whatever the name
        const bool toPng = !stream && ".png" == outputExtension;
        const bool toExr = !stream && ".exr" == outputExtension;
        const bool toSequence = toPng || toExr;
        if (output != "-" && !toSequence && raw.empty() && y4m.empty() &&
            (hasExtension(outputExtension, { ".mov", ".mp4", ".m4v", ".y4m", ".jpg", ".jpeg", ".tif", ".tiff" })))
        {
            throw std::runtime_error("Unsupported output (raw stream, -raw, -y4m, .png and .exr sequences are built in): " + output);
        }
        const auto outputSplit = splitFileNameNumber(outputPath.stem().string());
        const int outputStartFrame = std::atoi(outputSplit.second.c_str());
        const int outputNumberPadding = static_cast<int>(getNumberPadding(outputSplit.second));
        FILE* f = nullptr;
        if (!toSequence)
        {
            f = output == "-" ? stdout : fopen(output.c_str(), "wb");
            if (!f)
            {
                throw std::runtime_error("Cannot open: " + output);
            }
        }
        // stills and sequence frames are decoded once and kept resident (-cache GiB per GPU, least recently used out first)
        timelineWrapper->getMediaStore()->setDeviceCacheBytes(static_cast<size_t>(std::max(0.0, cacheGiB) * 1073741824.0));
        FrameSchedulerOptions options;
        options.devices.clear();
        for (int i = 0; i < std::max(1, gpus); ++i)
        {
            for (int k = 0; k < std::max(1, workers); ++k) options.devices.push_back(i);
        }
        options.pluginSearchPath = getOpenFXPluginPaths(argv[0]);
        if (inflight > 0) options.framesInFlight = inflight;
        if (!raw.empty() && !parseRawFormat(raw, options.rawType, options.rawChannels))
        {
            throw std::runtime_error("Cannot find the given raw format");
        }
        if (toSequence)
        {
            // the file format decides the sample type; frames of another type are converted on the device
            ImageGraph graph(context, std::filesystem::path(input).parent_path(), timelineWrapper);
            const std::string type = graph.getImageDataType();
            options.rawChannels = 4;
            if (toPng) options.rawType = ("u8" == type) ? TC_U8 : TC_U16;
            else options.rawType = ("float" == type) ? TC_F32 : TC_F16;
        }
        if (!y4m.empty())
        {
            if (!parseY4mFormat(y4m, options.y4mFormat))
            {
                throw std::runtime_error("Cannot find the given y4m format");
            }
            // App.cpp:218-221: the stream header goes out before the first frame
            ImageGraph graph(context, std::filesystem::path(input).parent_path(), timelineWrapper);
            const std::string header = y4mHeader(graph.getImageSize().x, graph.getImageSize().y, timeRange.duration().rate(), y4m);
            fwrite(header.c_str(), header.size(), 1, f);
        }
        const auto wallStart = std::chrono::steady_clock::now();
        // decode the timeline's stills in the background, in clip order, while the first frames render
        MediaPrefetch prefetch(timelineWrapper->getMediaStore(), timelineWrapper->getStillMedia(), 2);
        FrameScheduler scheduler(context, timelineWrapper, options);
        double firstFrameSeconds = 0.0, halfWaySeconds = 0.0;
        firstFrame = std::max(0, std::min(firstFrame, frames));
        const int endFrame = lastFrame < 0 ? frames : std::max(firstFrame, std::min(lastFrame + 1, frames));
        scheduler.run(firstFrame, endFrame, [&](int frame, const void* pixels, const ImageSpec& spec, size_t bytes)
        {
            const int halfWay = firstFrame + (endFrame - firstFrame) / 2;
            if (verbose && (firstFrame == frame || halfWay == frame))
            {
                const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - wallStart).count();
                if (firstFrame == frame) firstFrameSeconds = t;
                if (halfWay == frame) halfWaySeconds = t;
            }
            if (output != "-")
            {
                std::cout << frame << "/" << frames << std::endl;
            }
            if (toSequence && (spec.width <= 0 || spec.height <= 0))
            {
                return; // no image at this time: nothing to write
            }
            if (toSequence)
            {
                // App.cpp:243-252: prefix + (start number + frame number of the absolute time) + extension
                const int number = outputStartFrame + static_cast<int>(timeRange.start_time().value()) + frame;
                const std::string fileName = getSequenceFrame(
                    outputPath.parent_path().string(), outputSplit.first, number, outputNumberPadding, outputPath.extension().string());
                const bool ok = toPng ?
                    writePngFrame(fileName, pixels, spec.width, spec.height, TC_U8 == spec.format ? 8 : 16, false, true) :
                    writeExr(fileName, pixels, spec.width, spec.height, spec.format);
                if (!ok)
                {
                    throw std::runtime_error("Cannot write: " + fileName);
                }
                return;
            }
            if (options.y4mFormat >= 0)
            {
                fwrite("FRAME\n", 6, 1, f);
            }
            fwrite(pixels, 1, bytes, f);
        });
        if (f && f != stdout) fclose(f);
        if (NodeTimers::getEnabled())
        {
            // -timers: device time per node label over the whole run, on stderr
            std::cerr << NodeTimers::report();
        }
        if (verbose)
        {
            // -v (App.cpp:98-100): run statistics on stderr, so they never mix with a stream on stdout
            const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wallStart).count();
            unsigned long long allocCalls = 0, allocNs = 0;
            tc_alloc_stats(&allocCalls, &allocNs);
            const double secondHalf = seconds - halfWaySeconds;
            const int rendered = endFrame - firstFrame;
            std::cerr << "frames: " << rendered << "  seconds: " << seconds << "  frames/s: " << (seconds > 0.0 ? rendered / seconds : 0.0)
                      << "  first frame after: " << firstFrameSeconds << " s  second half: "
                      << (secondHalf > 0.0 ? (rendered - rendered / 2) / secondHalf : 0.0) << " frames/s"
                      << "  kernel launches: " << tc_launch_count() << "  allocator calls: " << allocCalls << " (" << allocNs * 1e-9 << " s)"
                      << "  source uploads: " << timelineWrapper->getMediaStore()->getUploadCount() << " ("
                      << timelineWrapper->getMediaStore()->getUploadBytes() / 1.0e9 << " GB)" << std::endl;
        }
    }
    catch (const std::exception& e)
    {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
