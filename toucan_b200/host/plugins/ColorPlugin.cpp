// toucanColorPlugin.ofx: toucan:ColorConvert, :Premult, :Unpremult.
// Parameter sets and local initialisers: plugins/ColorPlugin.cpp:157-179,207-218.
#include "Plugin.h"

#include <cstring>

namespace toucan_ofx
{
    namespace
    {
        int colorConvert(const RenderArgs& a, const Params& p)
        {
            const std::string fromSpace = p.getString("fromspace", "");
            const std::string toSpace = p.getString("tospace", "");
            // `int premult = 0` receives the described default `true` as a plain int, which
            // the host never delivers: the effective flag is what the metadata says, else 0.
            const int premult = p.getBoolAsInt("premult", 0);
            const std::string contextKey = p.getString("context_key", "");
            const std::string contextValue = p.getString("context_value", "");
            const std::string colorConfig = p.getString("color_config", "");
            (void)contextValue;
            if (!colorConfig.empty() || !contextKey.empty())
            {
                // Only OCIO's built-in config, baked per (from, to) pair, is in scope.
                return TC_ERR_UNSUPPORTED;
            }
            if (a.chain)
            {
                tc_chain_op op = chainOp(TC_CHAIN_COLOR_CONVERT);
                op.flag = premult;
                strncpy(op.name_a, fromSpace.c_str(), sizeof(op.name_a) - 1);
                strncpy(op.name_b, toSpace.c_str(), sizeof(op.name_b) - 1);
                return chainAppend(a, op);
            }
            return tc_color_convert2(&a.output, &a.source[0], fromSpace.c_str(), toSpace.c_str(), premult, a.stream);
        }

        int premult(const RenderArgs& a, const Params&)
        {
            if (a.chain) return chainAppend(a, chainOp(TC_CHAIN_PREMULT));
            return tc_premult(&a.output, &a.source[0], a.stream);
        }

        int unpremult(const RenderArgs& a, const Params&)
        {
            if (a.chain) return chainAppend(a, chainOp(TC_CHAIN_UNPREMULT));
            return tc_unpremult(&a.output, &a.source[0], a.stream);
        }
    }

    const std::vector<EffectDef>& effects()
    {
        static const std::vector<EffectDef> table = {
            { "toucan:ColorConvert", "ColorConvert", Context::Filter,
              { paramString("fromspace", "From Space", ""), paramString("tospace", "To Space", ""), paramBool("premult", "Premultiplied", 1),
                paramString("context_key", "Context Key", ""), paramString("context_value", "Context Value", ""),
                paramString("color_config", "ColorConfig", "") },
              &colorConvert, true },
            { "toucan:Premult", "Premult", Context::Filter, {}, &premult, true },
            { "toucan:Unpremult", "Unpremult", Context::Filter, {}, &unpremult, true },
        };
        return table;
    }
}
