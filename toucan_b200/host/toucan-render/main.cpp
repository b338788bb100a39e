// toucan-render: render a timeline to raw frames.  Mirrors bin/toucan-render/App.cpp:152-265
// for the in-scope outputs: raw frames in the timeline's working format to a file or stdout
// ("-").  The serial frame loop of App.cpp:222-264 is replaced by toucan::FrameScheduler:
// frames are sharded in contiguous blocks over the GPUs and written in order.  Both output
// conversions of App.cpp:267-485 (-raw: paste to the raw spec; -y4m: paste + libswscale to
// planar YUV) run on the device, so only the writer's bytes cross PCIe.  Movie and image-file
// encoders (FFmpeg, OIIO writers) are host-side writers outside the hot path.
//
//   toucan-render input.otio output|- [-raw rgb24|rgba|rgb48|rgba64|rgbaf16|rgbf32|rgbaf32]
//                 [-y4m 422|444|444alpha|444p16] [-gpus N] [-workers K]
//                 [-print_start] [-print_duration] [-print_rate] [-print_size]
#include <toucanRender/FrameScheduler.h>
#include <toucanRender/TimelineWrapper.h>
#include <toucanRender/Util.h>

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
        int workers = 1; // render threads per GPU: small frames are bound by the host-side graph build, which overlaps across workers
        bool printStart = false, printDuration = false, printRate = false, printSize = false;
        for (int i = 1; i < argc; ++i)
        {
            const std::string a = argv[i];
            if (a == "-gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
            else if (a == "-workers" && i + 1 < argc) workers = std::atoi(argv[++i]);
            else if (a == "-raw" && i + 1 < argc) raw = argv[++i];
            else if (a == "-y4m" && i + 1 < argc) y4m = argv[++i];
            else if (a == "-print_start") printStart = true;
            else if (a == "-print_duration") printDuration = true;
            else if (a == "-print_rate") printRate = true;
            else if (a == "-print_size") printSize = true;
            else if (input.empty()) input = a;
            else if (output.empty()) output = a;
        }
        if (input.empty() || (output.empty() && !(printStart || printDuration || printRate || printSize)))
        {
            std::cerr << "usage: toucan-render input.otio output.raw|- [-raw format | -y4m format] [-gpus N] [-workers K] [-print_start|-print_duration|-print_rate|-print_size]" << std::endl;
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

        FILE* f = output == "-" ? stdout : fopen(output.c_str(), "wb");
        if (!f)
        {
            throw std::runtime_error("Cannot open: " + output);
        }
        // stills and sequence frames are decoded once and kept resident (up to 8 GiB per GPU)
        timelineWrapper->getMediaStore()->setDeviceCacheBytes(size_t(8) << 30);
        FrameSchedulerOptions options;
        options.devices.clear();
        for (int i = 0; i < std::max(1, gpus); ++i)
        {
            for (int k = 0; k < std::max(1, workers); ++k) options.devices.push_back(i);
        }
        options.pluginSearchPath = getOpenFXPluginPaths(argv[0]);
        if (!raw.empty() && !parseRawFormat(raw, options.rawType, options.rawChannels))
        {
            throw std::runtime_error("Cannot find the given raw format");
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
        FrameScheduler scheduler(context, timelineWrapper, options);
        scheduler.run(0, frames, [&](int frame, const void* pixels, const ImageSpec&, size_t bytes)
        {
            if (output != "-")
            {
                std::cout << frame << "/" << frames << std::endl;
            }
            if (options.y4mFormat >= 0)
            {
                fwrite("FRAME\n", 6, 1, f);
            }
            fwrite(pixels, 1, bytes, f);
        });
        if (f != stdout) fclose(f);
    }
    catch (const std::exception& e)
    {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
