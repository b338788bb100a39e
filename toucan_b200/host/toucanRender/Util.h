// Small helpers of the render library, under the names the reference's callers use (lib/toucanRender/Util.h):
// file-name handling for image sequences, the fit box, OTIO any <-> Imath vector conversion, the plug-in
// search path.
#pragma once

#include <Imath/ImathVec.h>
#include <opentimelineio/otio.h>

#include <filesystem>
#include <string>
#include <utility>
#include <vector>

namespace toucan
{
    namespace util
    {
        using Path = std::filesystem::path;
        using Size = IMATH_NAMESPACE::V2i;
        using Color = IMATH_NAMESPACE::V4f;
        using Box = IMATH_NAMESPACE::Box2i;
        using StringPair = std::pair<std::string, std::string>;
    }

    // ---- strings and URLs ----
    std::string toLower(const std::string& text);
    //! True if `extension` (any case, with its dot) is one of `extensions` (lower case).
    bool hasExtension(const std::string& extension, const std::vector<std::string>& extensions);
    //! "proto://rest" -> ("proto:/", "rest"); without a protocol -> ("", url).
    util::StringPair splitURLProtocol(const std::string& url);

    // ---- numbered image sequences ----
    //! directory / (prefix + frame zero-padded to `padding` digits + suffix)
    std::string getSequenceFrame(const util::Path& directory, const std::string& namePrefix, int frame, int padding, const std::string& nameSuffix);
    //! A file stem split before its trailing digits: ("render.", "0042"); no digits -> (stem, "").
    util::StringPair splitFileNameNumber(const std::string& stem);
    //! Digits of a zero-padded number ("0042" -> 4); 0 for a number written without padding.
    size_t getNumberPadding(const std::string& number);
    //! The files of the numbered sequence `path` belongs to, in frame order (the path alone if its stem has no number).
    std::vector<util::Path> getSequence(const util::Path& path);

    // ---- geometry and OTIO values ----
    //! The box inside a frame of size `a` that an image of size `b` is fitted into: aspect preserved, integer
    //! truncation, centred (inclusive max corner).
    util::Box fit(const util::Size& a, const util::Size& b);
    //! {x, y} as int64_t / {r, g, b, a} as double: the element types OTIO metadata holds.
    OTIO_NS::AnyVector vecToAny(const util::Size&);
    OTIO_NS::AnyVector vecToAny(const util::Color&);
    //! The reverse; a vector of any other length leaves `out` untouched, an element of another type throws.
    void anyToVec(const OTIO_NS::AnyVector&, util::Size& out);
    void anyToVec(const OTIO_NS::AnyVector&, util::Color& out);

    // ---- plug-ins ----
    //! Where *.ofx modules are looked for: the executable's directory, two levels above it, the system OpenFX
    //! directory, then every entry of $OFX_PLUGIN_PATH.
    std::vector<util::Path> getOpenFXPluginPaths(const util::Path& executablePath);
}
