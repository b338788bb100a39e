// toucanFilterPlugin.ofx: toucan:Blur, :ColorMap, :Invert, :Pow, :Saturate, :UnsharpMask.
// Parameter sets and local initialisers: plugins/FilterPlugin.cpp:160-162,185-186,243-245,
// 269-270,402-404,427-428,479-481,504-505,557-571,597-604.
#include "Plugin.h"

#include <cstring>

namespace toucan_ofx
{
    namespace
    {
        int blur(const RenderArgs& a, const Params& p)
        {
            const double radius = p.getDouble("radius", 0.0);
            // in the host's chain fusion a convolution node may START a chain (its followers become the kernel's epilogue)
            if (a.chain) return chainAppend(a, chainOp(TC_CHAIN_BLUR, static_cast<float>(radius)));
            return tc_blur(&a.output, &a.source[0], static_cast<float>(radius), a.stream);
        }

        int colorMap(const RenderArgs& a, const Params& p)
        {
            // color_map on RGB + copy of the alpha channel: one pass
            const std::string mapName = p.getString("map_name", "plasma");
            if (a.chain)
            {
                tc_chain_op op = chainOp(TC_CHAIN_COLOR_MAP);
                strncpy(op.name_a, mapName.c_str(), sizeof(op.name_a) - 1);
                return chainAppend(a, op);
            }
            return tc_color_map(&a.output, &a.source[0], mapName.c_str(), a.stream);
        }

        int invert(const RenderArgs& a, const Params&)
        {
            // invert on channels [0,3) + copy of the alpha channel: one pass
            if (a.chain) return chainAppend(a, chainOp(TC_CHAIN_INVERT));
            return tc_invert(&a.output, &a.source[0], a.stream);
        }

        int pow(const RenderArgs& a, const Params& p)
        {
            const double value = p.getDouble("value", 1.0);
            if (a.chain) return chainAppend(a, chainOp(TC_CHAIN_POW, static_cast<float>(value)));
            return tc_pow(&a.output, &a.source[0], static_cast<float>(value), a.stream);
        }

        int saturate(const RenderArgs& a, const Params& p)
        {
            const double value = p.getDouble("value", 1.0);
            if (a.chain) return chainAppend(a, chainOp(TC_CHAIN_SATURATE, static_cast<float>(value)));
            return tc_saturate(&a.output, &a.source[0], static_cast<float>(value), a.stream);
        }

        int unsharpMask(const RenderArgs& a, const Params& p)
        {
            const std::string kernel = p.getString("kernel", "gaussian");
            const double width = p.getDouble("width", 3.0);
            const double contrast = p.getDouble("contrast", 1.0);
            const double threshold = p.getDouble("threshold", 0.0);
            if (a.chain)
            {
                tc_chain_op op = chainOp(TC_CHAIN_UNSHARP, static_cast<float>(width));
                strncpy(op.name_a, kernel.c_str(), sizeof(op.name_a) - 1);
                op.color[0] = static_cast<float>(contrast);
                op.color[1] = static_cast<float>(threshold);
                return chainAppend(a, op);
            }
            return tc_unsharp_mask(&a.output, &a.source[0], kernel.c_str(), static_cast<float>(width), static_cast<float>(contrast),
                                   static_cast<float>(threshold), a.stream);
        }
    }

    const std::vector<EffectDef>& effects()
    {
        static const std::vector<EffectDef> table = {
            { "toucan:Blur", "Blur", Context::Filter, { paramDouble("radius", "Radius", 10.0) }, &blur, true },
            { "toucan:ColorMap", "ColorMap", Context::Filter, { paramString("map_name", "Map Name", "plasma") }, &colorMap, true },
            { "toucan:Invert", "Invert", Context::Filter, {}, &invert, true },
            { "toucan:Pow", "Pow", Context::Filter, { paramDouble("value", "Value", 1.0) }, &pow, true },
            { "toucan:Saturate", "Saturate", Context::Filter, { paramDouble("value", "Value", 1.0) }, &saturate, true },
            { "toucan:UnsharpMask", "UnsharpMask", Context::Filter,
              { paramString("kernel", "Kernel", "gaussian"), paramDouble("width", "Width", 1.0), paramDouble("contrast", "Contrast", 1.0),
                paramDouble("threshold", "Threshold", 1.0) },
              &unsharpMask, true },
        };
        return table;
    }
}
