// Image property set -> device image view (the inverse of the host's bufToPropSet).
// Plays the role of plugins/Util.cpp:10-72 (propSetToBuf, a non-owning OIIO::ImageBuf over
// host memory) with a non-owning tc_image over DEVICE memory.
#pragma once

#include <OpenFX/ofxImageEffect.h>

#include <toucan_cuda.h>

namespace toucan_ofx
{
    //! Returns false if the property set does not describe a tightly packed 4-channel
    //! image of a supported depth (the render is then skipped, like the reference's
    //! UNKNOWN-format ImageBuf).
    bool propSetToImage(const OfxPropertySuiteV1&, OfxPropertySetHandle, tc_image&);
}
