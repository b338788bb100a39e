// toucan-render: render a timeline to raw frames.  Mirrors the frame loop of
// bin/toucan-render/App.cpp:152-265 for the in-scope outputs: raw frames in the timeline's
// working format to a file or stdout ("-").  Encoders (FFmpeg, PNG, y4m) are host-side
// writers outside the hot path.
//
//   toucan-render input.otio output.raw|- [-gpu N] [-print_start] [-print_duration] [-print_rate] [-print_size]
#include <toucanRender/ImageEffectHost.h>
#include <toucanRender/ImageGraph.h>
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
        std::string input, output;
        int gpu = 0;
        bool printStart = false, printDuration = false, printRate = false, printSize = false;
        for (int i = 1; i < argc; ++i)
        {
            const std::string a = argv[i];
            if (a == "-gpu" && i + 1 < argc) gpu = std::atoi(argv[++i]);
            else if (a == "-print_start") printStart = true;
            else if (a == "-print_duration") printDuration = true;
            else if (a == "-print_rate") printRate = true;
            else if (a == "-print_size") printSize = true;
            else if (input.empty()) input = a;
            else if (output.empty()) output = a;
        }
        if (input.empty() || (output.empty() && !(printStart || printDuration || printRate || printSize)))
        {
            std::cerr << "usage: toucan-render input.otio output.raw|- [-gpu N] [-print_start|-print_duration|-print_rate|-print_size]" << std::endl;
            return 1;
        }

        auto context = ftk::Context::create();
        auto timelineWrapper = std::make_shared<TimelineWrapper>(std::filesystem::path(input));
        const OTIO_NS::TimeRange& timeRange = timelineWrapper->getTimeRange();
        const OTIO_NS::RationalTime timeInc(1.0, timeRange.duration().rate());
        const int frames = timeRange.duration().to_frames();
        auto graph = std::make_shared<ImageGraph>(context, std::filesystem::path(input).parent_path(), timelineWrapper);

        if (printStart) { std::cout << timeRange.start_time().value() << std::endl; return 0; }
        if (printDuration) { std::cout << timeRange.duration().value() << std::endl; return 0; }
        if (printRate) { std::cout << timeRange.duration().rate() << std::endl; return 0; }
        if (printSize) { std::cout << graph->getImageSize().x << "x" << graph->getImageSize().y << std::endl; return 0; }

        DeviceContext::bind(gpu);
        auto host = std::make_shared<ImageEffectHost>(context, getOpenFXPluginPaths(argv[0]));

        FILE* f = output == "-" ? stdout : fopen(output.c_str(), "wb");
        if (!f)
        {
            throw std::runtime_error("Cannot open: " + output);
        }
        std::vector<unsigned char> pixels;
        int n = 0;
        for (OTIO_NS::RationalTime time = timeRange.start_time(); time <= timeRange.end_time_inclusive(); time += timeInc, ++n)
        {
            if (output != "-")
            {
                std::cout << (time - timeRange.start_time()).value() << "/" << frames << std::endl;
            }
            if (auto node = graph->exec(host, time))
            {
                const Image image = node->exec();
                pixels.resize(image.spec().image_bytes());
                image.download(pixels.data());
                fwrite(pixels.data(), 1, pixels.size(), f);
            }
        }
        if (f != stdout) fclose(f);
    }
    catch (const std::exception& e)
    {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
