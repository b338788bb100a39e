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
        // Same arithmetic as the reference (Util.cpp:123-147): double aspect ratios and
        // truncation to int, because the box decides which pixels the resize writes.
        const double aAspect = a.y > 0 ? (a.x / static_cast<double>(a.y)) : 1.0;
        const double bAspect = b.y > 0 ? (b.x / static_cast<double>(b.y)) : 1.0;
        int w = 0;
        int h = 0;
        if (bAspect > aAspect)
        {
            w = a.x;
            h = static_cast<int>(w / bAspect);
        }
        else
        {
            h = a.y;
            w = static_cast<int>(h * bAspect);
        }
        IMATH_NAMESPACE::Box2i out;
        out.min.x = a.x / 2 - w / 2;
        out.min.y = a.y / 2 - h / 2;
        out.max.x = out.min.x + w - 1;
        out.max.y = out.min.y + h - 1;
        return out;
    }

    OTIO_NS::AnyVector vecToAny(const IMATH_NAMESPACE::V2i& vec)
    {
        return OTIO_NS::AnyVector{ static_cast<int64_t>(vec.x), static_cast<int64_t>(vec.y) };
    }

    OTIO_NS::AnyVector vecToAny(const IMATH_NAMESPACE::V4f& vec)
    {
        return OTIO_NS::AnyVector{
            static_cast<double>(vec.x), static_cast<double>(vec.y), static_cast<double>(vec.z), static_cast<double>(vec.w) };
    }

    void anyToVec(const OTIO_NS::AnyVector& any, IMATH_NAMESPACE::V2i& out)
    {
        // strict any_cast like the reference: a JSON real in "size" throws bad_any_cast
        if (2 == any.size())
        {
            out.x = static_cast<int>(std::any_cast<int64_t>(any[0]));
            out.y = static_cast<int>(std::any_cast<int64_t>(any[1]));
        }
    }

    void anyToVec(const OTIO_NS::AnyVector& any, IMATH_NAMESPACE::V4f& out)
    {
        if (4 == any.size())
        {
            out.x = static_cast<float>(std::any_cast<double>(any[0]));
            out.y = static_cast<float>(std::any_cast<double>(any[1]));
            out.z = static_cast<float>(std::any_cast<double>(any[2]));
            out.w = static_cast<float>(std::any_cast<double>(any[3]));
        }
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
