// toucanGeneratorPlugin.ofx: toucan:Checkers, toucan:Fill, toucan:Gradient, toucan:Noise.
// Parameter sets and local initialisers: plugins/GeneratorPlugin.cpp:72-75,157-174,198-203,
// 262-267,289-295,344-360,384-389,472-490,514-523.  Render bodies call the sm_100a kernels.
#include "Plugin.h"

namespace toucan_ofx
{
    namespace
    {
        void toFloat(const double in[4], float out[4])
        {
            for (int i = 0; i < 4; ++i) out[i] = static_cast<float>(in[i]);
        }

        int checkers(const RenderArgs& a, const Params& p)
        {
            int64_t checkerSize[2] = { 0, 0 };
            double color1[4] = { 0.0, 0.0, 0.0, 0.0 };
            double color2[4] = { 0.0, 0.0, 0.0, 0.0 };
            p.getInt2("checkerSize", checkerSize);
            p.getRGBA("color1", color1);
            p.getRGBA("color2", color2);
            float c1[4], c2[4];
            toFloat(color1, c1);
            toFloat(color2, c2);
            return tc_checkers(&a.output, static_cast<int>(checkerSize[0]), static_cast<int>(checkerSize[1]), c1, c2, a.stream);
        }

        int fill(const RenderArgs& a, const Params& p)
        {
            double color[4] = { 0.0, 0.0, 0.0, 0.0 };
            p.getRGBA("color", color);
            float c[4];
            toFloat(color, c);
            return tc_fill(&a.output, c, a.stream);
        }

        int gradient(const RenderArgs& a, const Params& p)
        {
            double color1[4] = { 0.0, 0.0, 0.0, 0.0 };
            double color2[4] = { 0.0, 0.0, 0.0, 0.0 };
            p.getRGBA("color1", color1);
            p.getRGBA("color2", color2);
            const bool vertical = p.getBool("vertical", false);
            float c1[4], c2[4];
            toFloat(color1, c1);
            toFloat(color2, c2);
            return tc_gradient(&a.output, c1, c2, vertical ? 1 : 0, a.stream);
        }

        int noise(const RenderArgs& a, const Params& p)
        {
            const std::string type = p.getString("type", "");
            const double va = p.getDouble("a", 0.0);
            const double vb = p.getDouble("b", 0.0);
            const int mono = p.getBoolAsInt("mono", 0);
            const int64_t seed = p.getInt("seed", 0);
            // noise() ADDS to the output; generator outputs start zero-filled in the
            // reference (ImageEffect.cpp:93), which the kernel folds in.
            return tc_noise(&a.output, type.c_str(), static_cast<float>(va), static_cast<float>(vb), mono, static_cast<int>(seed), a.stream);
        }

        const ParamDef size = paramInt2("size", "Size", 1280, 720);
    }

    const std::vector<EffectDef>& effects()
    {
        static const std::vector<EffectDef> table = {
            { "toucan:Checkers", "Checkers", Context::Generator,
              { size, paramInt2("checkerSize", "Checker Size", 100, 100), paramRGBA("color1", "Color 1", 0, 0, 0, 1),
                paramRGBA("color2", "Color 2", 1, 1, 1, 1) },
              &checkers },
            { "toucan:Fill", "Fill", Context::Generator, { size, paramRGBA("color", "Color", 0, 0, 0, 0) }, &fill },
            { "toucan:Gradient", "Gradient", Context::Generator,
              { size, paramRGBA("color1", "Color 1", 0, 0, 0, 1), paramRGBA("color2", "Color 2", 1, 1, 1, 1),
                paramBool("vertical", "Vertical", 0) },
              &gradient },
            { "toucan:Noise", "Noise", Context::Generator,
              { size, paramString("type", "Type", "gaussian"), paramDouble("a", "A", 0.0), paramDouble("b", "B", 0.0),
                paramBool("mono", "Mono", 0), paramInt("seed", "Seed", 0) },
              &noise },
        };
        return table;
    }
}
