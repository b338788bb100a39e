// toucanTransformPlugin.ofx: toucan:Crop, :Flip, :Flop, :Resize, :Rotate.
// Parameter sets and local initialisers: plugins/TransformPlugin.cpp:160-168,192-195,331-342,
// 367-372,420-430,455-460.
#include "Plugin.h"

#include <cmath>

namespace toucan_ofx
{
    namespace
    {
        int crop(const RenderArgs& a, const Params& p)
        {
            int64_t pos[2] = { 0, 0 };
            int64_t size[2] = { 0, 0 };
            p.getInt2("pos", pos);
            p.getInt2("size", size);
            // cut(src, ROI(pos, pos + size)) + copy into the output: addressing only
            return tc_crop(&a.output, &a.source[0], static_cast<int>(pos[0]), static_cast<int>(pos[1]), static_cast<int>(size[0]),
                           static_cast<int>(size[1]), a.stream);
        }

        // Flip / Flop only change where the source pixel is read: in the host's chain fusion they are position ops
        int flip(const RenderArgs& a, const Params&)
        {
            if (a.chain) return chainAppend(a, chainOp(TC_CHAIN_FLIP));
            return tc_flip(&a.output, &a.source[0], a.stream);
        }

        int flop(const RenderArgs& a, const Params&)
        {
            if (a.chain) return chainAppend(a, chainOp(TC_CHAIN_FLOP));
            return tc_flop(&a.output, &a.source[0], a.stream);
        }

        int resize(const RenderArgs& a, const Params& p)
        {
            int64_t size[2] = { 0, 0 };
            p.getInt2("size", size);
            const std::string filterName = p.getString("filter_name", "");
            const double filterWidth = p.getDouble("filter_width", 0.0);
            if (size[0] <= 0 || size[1] <= 0)
            {
                // resize(..., ROI(0, 0, 0, 0)) renders nothing: the output keeps its zero fill
                return tc_memset(a.output.data, 0, tc_image_bytes(a.output.w, a.output.h, a.output.type), a.stream);
            }
            return tc_resize(&a.output, &a.source[0], filterName.c_str(), static_cast<float>(filterWidth), a.stream);
        }

        int rotate(const RenderArgs& a, const Params& p)
        {
            const double angle = p.getDouble("angle", 0.0);
            const std::string filterName = p.getString("filter_name", "");
            const double filterWidth = p.getDouble("filter_width", 0.0);
            // degrees -> radians in double, narrowed by rotate(float angle) (TransformPlugin.cpp:465)
            const float radians = static_cast<float>(angle / 360.F * 2.F * M_PI);
            return tc_rotate(&a.output, &a.source[0], radians, filterName.c_str(), static_cast<float>(filterWidth), a.stream);
        }
    }

    const std::vector<EffectDef>& effects()
    {
        static const std::vector<EffectDef> table = {
            { "toucan:Crop", "Crop", Context::Filter, { paramInt2("pos", "Position", 0, 0), paramInt2("size", "Size", 0, 0) }, &crop },
            { "toucan:Flip", "Flip", Context::Filter, {}, &flip, true },
            { "toucan:Flop", "Flop", Context::Filter, {}, &flop, true },
            { "toucan:Resize", "Resize", Context::Filter,
              { paramInt2("size", "Size", 0, 0), paramString("filter_name", "Filter name", ""), paramDouble("filter_width", "Filter width", 0.0) },
              &resize },
            { "toucan:Rotate", "Rotate", Context::Filter,
              { paramDouble("angle", "Angle", 0.0), paramString("filter_name", "Filter name", ""),
                paramDouble("filter_width", "Filter width", 0.0) },
              &rotate },
        };
        return table;
    }
}
