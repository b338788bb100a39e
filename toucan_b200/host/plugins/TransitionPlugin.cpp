// toucanTransitionPlugin.ofx: toucan:Dissolve, :HorizontalWipe, :VerticalWipe.
// `value` is injected by the graph builder (lib/toucanRender/ImageGraph.cpp:280-281,312-313).
// The reference builds each transition from full-frame temporaries (fill, fill, mul, mul,
// add; the wipes add two UINT8 mattes via fill/paste/rotate): plugins/TransitionPlugin.cpp:
// 192-267, 296-356, 385-446.  Here each is ONE pass that replays the typed temporaries'
// quantisation in registers.
#include "Plugin.h"

namespace toucan_ofx
{
    namespace
    {
        int dissolve(const RenderArgs& a, const Params& p)
        {
            const double value = p.getDouble("value", 0.0);
            const tc_image& from = a.source[0];
            const tc_image& to = a.source[1];
            // The reference's test reads `A && B && C && D && (w differs) || (h differs)`
            // (TransitionPlugin.cpp:203-206): kept as written.
            if ((from.w > 0 && from.h > 0 && to.w > 0 && to.h > 0 && to.w != from.w) || to.h != from.h)
            {
                // fit `to` into `from`'s frame: resize + paste on a transparent buffer of from's spec
                int width = to.w;
                int height = to.h;
                const double fgAspect = to.w / static_cast<double>(to.h);
                const double bgAspect = from.w / static_cast<double>(from.h);
                if (fgAspect > bgAspect)
                {
                    width = from.w;
                    height = static_cast<int>(width / fgAspect);
                }
                else
                {
                    height = from.h;
                    width = static_cast<int>(height * fgAspect);
                }
                tc_image resized = { nullptr, width, height, to.type };
                tc_image pasted = { nullptr, from.w, from.h, from.type };
                int st = tc_malloc(&pasted.data, tc_image_bytes(pasted.w, pasted.h, pasted.type), a.stream);
                if (TC_OK == st && width > 0 && height > 0)
                {
                    st = tc_malloc(&resized.data, tc_image_bytes(width, height, to.type), a.stream);
                    if (TC_OK == st) st = tc_resize(&resized, &to, "", 0.f, a.stream);
                    if (TC_OK == st) st = tc_paste_on_black(&pasted, from.w / 2 - width / 2, from.h / 2 - height / 2, &resized, a.stream);
                }
                else if (TC_OK == st)
                {
                    st = tc_memset(pasted.data, 0, tc_image_bytes(pasted.w, pasted.h, pasted.type), a.stream);
                }
                if (TC_OK == st) st = tc_dissolve(&a.output, &from, &pasted, value, a.stream);
                tc_free(resized.data, a.stream);
                tc_free(pasted.data, a.stream);
                return st;
            }
            return tc_dissolve(&a.output, &from, &to, value, a.stream);
        }

        int horizontalWipe(const RenderArgs& a, const Params& p)
        {
            return tc_wipe(&a.output, &a.source[0], &a.source[1], p.getDouble("value", 0.0), 0, a.stream);
        }

        int verticalWipe(const RenderArgs& a, const Params& p)
        {
            return tc_wipe(&a.output, &a.source[0], &a.source[1], p.getDouble("value", 0.0), 1, a.stream);
        }
    }

    const std::vector<EffectDef>& effects()
    {
        static const std::vector<EffectDef> table = {
            { "toucan:Dissolve", "Dissolve", Context::Transition, { paramDouble("value", "Value", 0.0) }, &dissolve },
            { "toucan:HorizontalWipe", "HorizontalWipe", Context::Transition, { paramDouble("value", "Value", 0.0) }, &horizontalWipe },
            { "toucan:VerticalWipe", "VerticalWipe", Context::Transition, { paramDouble("value", "Value", 0.0) }, &verticalWipe },
        };
        return table;
    }
}
