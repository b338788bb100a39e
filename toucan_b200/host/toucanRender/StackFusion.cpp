#include "StackFusion.h"

#include "Comp.h"
#include "ImageEffect.h"
#include "Read.h"
#include "Util.h"

#include <atomic>
#include <cmath>
#include <cstdlib>

namespace toucan
{
    namespace
    {
        std::atomic<int> fusionEnabled{ -1 };

        constexpr int maxLayers = 32;

        const ImageEffectNode* asEffect(const std::shared_ptr<IImageNode>& node, const char* name)
        {
            auto effect = dynamic_cast<const ImageEffectNode*>(node.get());
            return (effect && effect->getName() == name) ? effect : nullptr;
        }

        bool hasSizeOverride(const ImageEffectNode* node)
        {
            const auto& metaData = node->getMetaData();
            const auto i = metaData.find("size");
            return i != metaData.end() && i->second.has_value();
        }

        bool getDouble(const ImageEffectNode* node, const char* name, double& out)
        {
            const auto& params = node->getParams();
            const auto i = params.find(name);
            if (i == params.end())
            {
                return false;
            }
            if (auto v = std::any_cast<double>(&i->second))
            {
                out = *v;
                return true;
            }
            return false;
        }

        // [Blur of] a read node
        struct Source
        {
            IReadNode* read = nullptr;
            float radius = 0.F;
        };

        bool matchSource(const std::shared_ptr<IImageNode>& node, Source& out)
        {
            std::shared_ptr<IImageNode> n = node;
            if (auto blur = asEffect(n, "toucan:Blur"))
            {
                double radius = 0.0;
                if (hasSizeOverride(blur) || blur->getInputs().size() != 1 || !getDouble(blur, "radius", radius) || !(radius > 0.0))
                {
                    return false;
                }
                out.radius = static_cast<float>(radius);
                n = blur->getInputs()[0];
            }
            out.read = dynamic_cast<IReadNode*>(n.get());
            return out.read != nullptr;
        }

        struct Layer
        {
            Source a, b;
            bool transition = false;
            double value = 0.0;
        };

        bool matchLayer(const std::shared_ptr<IImageNode>& node, Layer& out)
        {
            if (auto dissolve = asEffect(node, "toucan:Dissolve"))
            {
                if (hasSizeOverride(dissolve) || dissolve->getInputs().size() != 2 || !getDouble(dissolve, "value", out.value))
                {
                    return false;
                }
                out.transition = true;
                return matchSource(dissolve->getInputs()[0], out.a) && matchSource(dissolve->getInputs()[1], out.b);
            }
            return matchSource(node, out.a);
        }

        // OIIO's float -> UINT8 -> float round trip: generator outputs are UINT8 (ImageEffect.cpp:93)
        float quantU8(float v)
        {
            float s = v * 255.0F;
            s += s < 0.0F ? -0.5F : 0.5F;
            s = s < 0.0F ? 0.0F : (s > 255.0F ? 255.0F : s);
            return static_cast<float>(static_cast<unsigned>(s)) * (1.0F / 255.0F);
        }
    }

    void setStackFusionEnabled(bool value) { fusionEnabled = value ? 1 : 0; }

    bool getStackFusionEnabled()
    {
        int v = fusionEnabled.load();
        if (v < 0)
        {
            const char* env = std::getenv("TOUCAN_NO_FUSION");
            v = (env && *env && *env != '0') ? 0 : 1;
            fusionEnabled = v;
        }
        return v != 0;
    }

    bool matchFillBackground(const std::shared_ptr<IImageNode>& node, IMATH_NAMESPACE::V2i& size, float color[4])
    {
        const ImageEffectNode* fill = asEffect(node, "toucan:Fill");
        if (!fill || !fill->getInputs().empty())
        {
            return false;
        }
        const auto& metaData = fill->getMetaData();
        const auto sizeIt = metaData.find("size");
        if (sizeIt == metaData.end() || !sizeIt->second.has_value())
        {
            return false;
        }
        anyToVec(std::any_cast<OTIO_NS::AnyVector>(sizeIt->second), size);
        if (size.x <= 0 || size.y <= 0)
        {
            return false;
        }
        // what Fill's render body receives: a vector of doubles from the metadata, else only element 0 of the
        // described default (= 0)
        for (int c = 0; c < 4; ++c) color[c] = 0.F;
        const auto& params = fill->getParams();
        const auto i = params.find("color");
        if (i != params.end())
        {
            if (auto v = std::any_cast<OTIO_NS::AnyVector>(&i->second))
            {
                for (size_t c = 0; c < v->size() && c < 4; ++c)
                {
                    auto d = std::any_cast<double>(&(*v)[c]);
                    if (!d) return false;
                    color[c] = static_cast<float>(*d);
                }
            }
            else if (auto d = std::any_cast<double>(&i->second))
            {
                color[0] = static_cast<float>(*d);
            }
            else
            {
                return false;
            }
        }
        return true;
    }

    bool tryFusedStack(const CompNode& top, Image& out)
    {
        if (!getStackFusionEnabled())
        {
            return false;
        }

        // Walk down the chain: Comp(premult){ layer, below }.
        std::vector<Layer> layers; // top first
        const IImageNode* node = &top;
        const ImageEffectNode* background = nullptr;
        std::shared_ptr<IImageNode> backgroundNode;
        while (true)
        {
            auto comp = dynamic_cast<const CompNode*>(node);
            if (!comp || !comp->getPremult() || comp->getInputs().size() != 2 || !comp->getInputs()[0] || !comp->getInputs()[1])
            {
                return false;
            }
            Layer layer;
            if (!matchLayer(comp->getInputs()[0], layer) || static_cast<int>(layers.size()) >= maxLayers)
            {
                return false;
            }
            layers.push_back(layer);
            const auto& below = comp->getInputs()[1];
            if ((background = asEffect(below, "toucan:Fill")))
            {
                backgroundNode = below;
                break;
            }
            node = below.get();
        }
        if (layers.empty() || !background->getInputs().empty())
        {
            return false;
        }

        // Background: a Fill generator of the output size.
        IMATH_NAMESPACE::V2i size(0, 0);
        float color[4] = { 0.F, 0.F, 0.F, 0.F };
        if (!matchFillBackground(backgroundNode, size, color))
        {
            return false;
        }
        for (float& c : color) c = quantU8(c);

        // All sources must be float32 RGBA frames of the output size (checked on the read
        // nodes' specs before anything is evaluated, and again on the evaluated frames).
        auto specOk = [&](const ImageSpec& spec)
        {
            return spec.width == size.x && spec.height == size.y && spec.nchannels == 4 && spec.format == TC_F32;
        };
        for (const Layer& layer : layers)
        {
            if (!specOk(layer.a.read->getSpec()) || (layer.transition && !specOk(layer.b.read->getSpec())))
            {
                return false;
            }
        }

        std::vector<Image> frames;
        std::vector<tc_stack_layer> desc(layers.size());
        for (size_t i = 0; i < layers.size(); ++i)
        {
            const Layer& layer = layers[layers.size() - 1 - i]; // bottom track first
            tc_stack_layer& d = desc[i];
            frames.push_back(layer.a.read->exec());
            if (!specOk(frames.back().spec())) return false;
            d.src_a = frames.back().data();
            d.radius_a = layer.a.radius;
            d.src_b = nullptr;
            d.radius_b = 0.F;
            d.value = layer.value;
            if (layer.transition)
            {
                frames.push_back(layer.b.read->exec());
                if (!specOk(frames.back().spec())) return false;
                d.src_b = frames.back().data();
                d.radius_b = layer.b.radius;
            }
        }

        Image result(ImageSpec(size.x, size.y, 4, TC_F32));
        const tc_image view = result.view();
        NodeTimers::Scope timer("Comp (fused stack)");
        const int status = tc_stack_f32(&view, color, desc.data(), static_cast<int>(desc.size()), DeviceContext::stream());
        if (TC_ERR_UNSUPPORTED == status)
        {
            return false; // e.g. a blur radius the fused kernel does not cover
        }
        tcCheck(status, "tc_stack_f32");
        out = result;
        return true;
    }
}
