#include "Util.h"

#include <cstring>

namespace toucan_ofx
{
    bool propSetToImage(const OfxPropertySuiteV1& props, OfxPropertySetHandle handle, tc_image& out)
    {
        out = tc_image{ nullptr, 0, 0, TC_U8 };
        if (!handle)
        {
            return false;
        }
        int bounds[4] = { 0, 0, 0, 0 };
        props.propGetIntN(handle, kOfxImagePropBounds, 4, bounds);
        char* components = nullptr;
        props.propGetString(handle, kOfxImageEffectPropComponents, 0, &components);
        char* depth = nullptr;
        props.propGetString(handle, kOfxImageEffectPropPixelDepth, 0, &depth);
        int rowBytes = 0;
        props.propGetInt(handle, kOfxImagePropRowBytes, 0, &rowBytes);
        void* data = nullptr;
        props.propGetPointer(handle, kOfxImagePropData, 0, &data);

        if (!components || 0 != strcmp(components, kOfxImageComponentRGBA) || !depth)
        {
            return false;
        }
        int type = -1;
        if (0 == strcmp(depth, kOfxBitDepthByte)) type = TC_U8;
        else if (0 == strcmp(depth, kOfxBitDepthShort)) type = TC_U16;
        else if (0 == strcmp(depth, kOfxBitDepthHalf)) type = TC_F16;
        else if (0 == strcmp(depth, kOfxBitDepthFloat)) type = TC_F32;
        if (type < 0)
        {
            return false;
        }
        out.data = data;
        out.w = bounds[2] - bounds[0];
        out.h = bounds[3] - bounds[1];
        out.type = type;
        // kernels address tightly packed rows
        return data && out.w > 0 && out.h > 0 &&
            static_cast<size_t>(rowBytes) == static_cast<size_t>(out.w) * 4 * tc_type_size(type);
    }
}
