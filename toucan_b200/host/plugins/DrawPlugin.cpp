// toucanDrawPlugin.ofx: toucan:Box, :Line, :Text.
// Parameter sets and local initialisers: plugins/DrawPlugin.cpp:150-181,198-212 (Box),
// :258-289,306-320 (Line), :366-400,418-434 (Text).
//
// Text: the reference rasterises glyphs with ImageBufAlgo::render_text (FreeType + a font file
// found on the machine).  Neither is part of this build, so toucan:Text copies its source --
// exactly what the reference produces when render_text finds no font (it returns false and
// the plug-in ignores the status, DrawPlugin.cpp:436-452).
#include "Plugin.h"

namespace toucan_ofx
{
    namespace
    {
        void colorOf(const Params& p, float out[4])
        {
            double color[4] = { 0.0, 0.0, 0.0, 0.0 };
            p.getRGBA("color", color);
            for (int i = 0; i < 4; ++i) out[i] = static_cast<float>(color[i]);
        }

        int box(const RenderArgs& a, const Params& p)
        {
            int64_t pos1[2] = { 0, 0 };
            int64_t pos2[2] = { 0, 0 };
            float color[4];
            p.getInt2("pos1", pos1);
            p.getInt2("pos2", pos2);
            colorOf(p, color);
            const int fill = p.getBoolAsInt("fill", 0);
            return tc_draw_box(&a.output, &a.source[0], static_cast<int>(pos1[0]), static_cast<int>(pos1[1]), static_cast<int>(pos2[0]),
                               static_cast<int>(pos2[1]), color, fill, a.stream);
        }

        int line(const RenderArgs& a, const Params& p)
        {
            int64_t pos1[2] = { 0, 0 };
            int64_t pos2[2] = { 0, 0 };
            float color[4];
            p.getInt2("pos1", pos1);
            p.getInt2("pos2", pos2);
            colorOf(p, color);
            const int skipFirstPoint = p.getBoolAsInt("skip_first_point", 0);
            return tc_draw_line(&a.output, &a.source[0], static_cast<int>(pos1[0]), static_cast<int>(pos1[1]), static_cast<int>(pos2[0]),
                                static_cast<int>(pos2[1]), color, skipFirstPoint, a.stream);
        }

        int text(const RenderArgs& a, const Params&)
        {
            return tc_convert(&a.output, &a.source[0], a.stream);
        }
    }

    const std::vector<EffectDef>& effects()
    {
        static const std::vector<EffectDef> table = {
            { "toucan:Box", "Box", Context::Filter,
              { paramInt2("pos1", "Position 1", 0, 0), paramInt2("pos2", "Position 2", 0, 0), paramRGBA("color", "Color", 1.0, 1.0, 1.0, 1.0),
                paramBool("fill", "Fill", 1) },
              &box },
            { "toucan:Line", "Line", Context::Filter,
              { paramInt2("pos1", "Position 1", 0, 0), paramInt2("pos2", "Position 2", 0, 0), paramRGBA("color", "Color", 1.0, 1.0, 1.0, 1.0),
                paramBool("skip_first_point", "Skip first point", 0) },
              &line },
            { "toucan:Text", "Text", Context::Filter,
              { paramInt2("pos", "Position", 0, 0), paramString("text", "Text", ""), paramInt("font_size", "Font size", 16),
                paramString("font_name", "Font name", ""), paramRGBA("color", "Color", 1.0, 1.0, 1.0, 1.0) },
              &text },
        };
        return table;
    }
}
