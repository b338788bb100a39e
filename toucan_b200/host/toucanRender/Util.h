// Helpers of the render library: sequence file names, fit(), any <-> vector, plug-in
// search paths.  Interface parity: lib/toucanRender/Util.h.
#pragma once

#include <Imath/ImathVec.h>
#include <opentimelineio/otio.h>

#include <filesystem>
#include <string>
#include <vector>

namespace toucan
{
    std::string toLower(const std::string&);
    bool hasExtension(const std::string& extension, const std::vector<std::string>& extensions);

    //! Split "proto://rest" into ("proto:/", "rest"); no protocol -> ("", url).
    std::pair<std::string, std::string> splitURLProtocol(const std::string&);

    //! prefix + zero padded frame + suffix, joined to a directory.
    std::string getSequenceFrame(
        const std::filesystem::path&,
        const std::string& namePrefix,
        int frame,
        int padding,
        const std::string& nameSuffix);

    //! Split trailing digits off a file stem.
    std::pair<std::string, std::string> splitFileNameNumber(const std::string&);
    size_t getNumberPadding(const std::string&);

    //! The files of the numbered sequence `path` belongs to, in frame order (the path alone if its stem has no number).
    std::vector<std::filesystem::path> getSequence(const std::filesystem::path&);

    //! Fit box b inside a (aspect preserved, integer truncation, centred).
    IMATH_NAMESPACE::Box2i fit(const IMATH_NAMESPACE::V2i& a, const IMATH_NAMESPACE::V2i& b);

    OTIO_NS::AnyVector vecToAny(const IMATH_NAMESPACE::V2i&);
    OTIO_NS::AnyVector vecToAny(const IMATH_NAMESPACE::V4f&);
    void anyToVec(const OTIO_NS::AnyVector&, IMATH_NAMESPACE::V2i&);
    void anyToVec(const OTIO_NS::AnyVector&, IMATH_NAMESPACE::V4f&);

    //! Plug-in search path: executable directory, two levels up, the system OpenFX
    //! directory and $OFX_PLUGIN_PATH.
    std::vector<std::filesystem::path> getOpenFXPluginPaths(const std::filesystem::path& executablePath);
}
