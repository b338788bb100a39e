#include "Util.h"

#include <map>

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <iomanip>
#include <sstream>

namespace toucan
{
    std::string toLower(const std::string& value)
    {
        std::string out(value);
        std::transform(out.begin(), out.end(), out.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
        return out;
    }

    bool hasExtension(const std::string& extension, const std::vector<std::string>& extensions)
    {
        const std::string lower = toLower(extension);
        return std::find(extensions.begin(), extensions.end(), lower) != extensions.end();
    }

    std::pair<std::string, std::string> splitURLProtocol(const std::string& url)
    {
        const size_t pos = url.find("://");
        if (pos == std::string::npos)
        {
            return { std::string(), url };
        }
        return { url.substr(0, pos + 2), url.substr(pos + 3) };
    }

    std::string getSequenceFrame(
        const std::filesystem::path& path,
        const std::string& namePrefix,
        int frame,
        int padding,
        const std::string& nameSuffix)
    {
        std::stringstream ss;
        ss << namePrefix << std::setw(padding) << std::setfill('0') << frame << nameSuffix;
        return (path / ss.str()).string();
    }

    std::pair<std::string, std::string> splitFileNameNumber(const std::string& stem)
    {
        size_t digits = 0;
        while (digits < stem.size() && std::isdigit(static_cast<unsigned char>(stem[stem.size() - 1 - digits])))
        {
            ++digits;
        }
        return { stem.substr(0, stem.size() - digits), stem.substr(stem.size() - digits) };
    }

    size_t getNumberPadding(const std::string& value)
    {
        return (!value.empty() && '0' == value[0]) ? value.size() : 0;
    }

    std::vector<std::filesystem::path> getSequence(const std::filesystem::path& path)
    {
        // Util.cpp:88-121: same prefix, same zero padding, same extension; ordered by number
        const auto split = splitFileNameNumber(path.stem().string());
        if (split.second.empty())
        {
            return { path };
        }
        std::map<int64_t, std::filesystem::path> byNumber;
        const size_t padding = getNumberPadding(split.second);
        const std::filesystem::path directory = path.parent_path().empty() ? std::filesystem::path(".") : path.parent_path();
        for (const auto& entry : std::filesystem::directory_iterator(directory))
        {
            if (!entry.is_regular_file()) continue;
            const auto other = splitFileNameNumber(entry.path().stem().string());
            if (other.first == split.first && !other.second.empty() && getNumberPadding(other.second) == padding &&
                entry.path().extension() == path.extension())
            {
                byNumber[std::atoi(other.second.c_str())] = entry.path();
            }
        }
        std::vector<std::filesystem::path> out;
        for (const auto& i : byNumber) out.push_back(i.second);
        if (out.empty()) out.push_back(path);
        return out;
    }

    IMATH_NAMESPACE::Box2i fit(const IMATH_NAMESPACE::V2i& a, const IMATH_NAMESPACE::V2i& b)
    {
        // The box decides which pixels the fit-resize writes, so the arithmetic is the reference's to the last
        // truncation (Util.cpp:123-147): width / height ratios in double (1 for an empty box), the side that
        // touches the frame copied, the other one scaled and truncated toward zero, halves taken before subtracting.
        auto aspect = [](const IMATH_NAMESPACE::V2i& v) { return v.y > 0 ? v.x / static_cast<double>(v.y) : 1.0; };
        const double target = aspect(b);
        const bool fullWidth = target > aspect(a);
        const int w = fullWidth ? a.x : static_cast<int>(a.y * target);
        const int h = fullWidth ? static_cast<int>(a.x / target) : a.y;
        const IMATH_NAMESPACE::V2i origin(a.x / 2 - w / 2, a.y / 2 - h / 2);
        IMATH_NAMESPACE::Box2i box;
        box.min = origin;
        box.max = IMATH_NAMESPACE::V2i(origin.x + w - 1, origin.y + h - 1);
        return box;
    }

    namespace
    {
        // OTIO stores numbers as int64_t or double and nothing else; any_cast is strict (a JSON real where an
        // integer is expected throws bad_any_cast, as in the reference)
        template <class Stored, class Component, int N>
        OTIO_NS::AnyVector pack(const Component (&c)[N])
        {
            OTIO_NS::AnyVector out;
            for (const Component& v : c) out.push_back(static_cast<Stored>(v));
            return out;
        }
        template <class Stored, class Component, int N>
        void unpack(const OTIO_NS::AnyVector& any, Component* const (&c)[N])
        {
            if (any.size() != static_cast<size_t>(N)) return; // any other length leaves the vector as it was
            for (int i = 0; i < N; ++i) *c[i] = static_cast<Component>(std::any_cast<Stored>(any[static_cast<size_t>(i)]));
        }
    }

    OTIO_NS::AnyVector vecToAny(const IMATH_NAMESPACE::V2i& vec)
    {
        const int c[2] = { vec.x, vec.y };
        return pack<int64_t>(c);
    }

    OTIO_NS::AnyVector vecToAny(const IMATH_NAMESPACE::V4f& vec)
    {
        const float c[4] = { vec.x, vec.y, vec.z, vec.w };
        return pack<double>(c);
    }

    void anyToVec(const OTIO_NS::AnyVector& any, IMATH_NAMESPACE::V2i& out)
    {
        int* const c[2] = { &out.x, &out.y };
        unpack<int64_t>(any, c);
    }

    void anyToVec(const OTIO_NS::AnyVector& any, IMATH_NAMESPACE::V4f& out)
    {
        float* const c[4] = { &out.x, &out.y, &out.z, &out.w };
        unpack<double>(any, c);
    }

    std::vector<std::filesystem::path> getOpenFXPluginPaths(const std::filesystem::path& executablePath)
    {
        const std::filesystem::path dir = executablePath.parent_path();
        std::vector<std::filesystem::path> out = { dir, dir / ".." / "..", "/usr/OFX/Plugins" };
        if (const char* env = std::getenv("OFX_PLUGIN_PATH"))
        {
            std::stringstream ss(env);
            std::string token;
            while (std::getline(ss, token, ':'))
            {
                if (!token.empty())
                {
                    out.push_back(token);
                }
            }
        }
        return out;
    }
}
