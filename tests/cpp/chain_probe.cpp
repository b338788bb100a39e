// Host-side probe (no GPU): load the OpenFX modules, create effect nodes the way ImageGraph does and ask each to
// DESCRIBE its pixel operation (the Render action in pointwise-chain mode).  That runs the plug-ins' parameter
// getters through the host's parameter suite -- the same calls a real render makes -- without launching a kernel.
// Prints one line per effect: name ok n kind value flag name_a name_b
#include <toucanRender/ImageEffect.h>
#include <toucanRender/ImageEffectHost.h>

#include <cstdio>
#include <string>

namespace
{
    struct Leaf : toucan::IImageNode
    {
        Leaf() : toucan::IImageNode("leaf") {}
        toucan::Image exec() override { return toucan::Image(); }
    };

    void probe(const std::shared_ptr<toucan::ImageEffectHost>& host, const std::string& name, const OTIO_NS::AnyDictionary& metaData)
    {
        auto node = host->createNode(metaData, name, { std::make_shared<Leaf>() });
        auto effect = std::dynamic_pointer_cast<toucan::ImageEffectNode>(node);
        tc_chain chain;
        chain.n = 0;
        const bool ok = effect && effect->describePointwise(chain);
        const tc_chain_op op = ok ? chain.ops[0] : tc_chain_op{ -1, 0.f, 0, "", "" };
        std::printf("%s %d %d %d %.9g %d [%s] [%s]\n", name.c_str(), ok ? 1 : 0, chain.n, op.kind, op.value, op.flag, op.name_a, op.name_b);
    }
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    auto context = ftk::Context::create();
    auto host = std::make_shared<toucan::ImageEffectHost>(context, std::vector<std::filesystem::path>{ argv[1] });
    OTIO_NS::AnyDictionary none;
    OTIO_NS::AnyDictionary pow{ { "value", 2.2 } };
    OTIO_NS::AnyDictionary saturate{ { "value", 0.25 } };
    OTIO_NS::AnyDictionary colorMap{ { "map_name", std::string("plasma") } };
    OTIO_NS::AnyDictionary convert{ { "fromspace", std::string("sRGB - Texture") }, { "tospace", std::string("ACEScg") }, { "premult", true } };
    OTIO_NS::AnyDictionary blur{ { "radius", 10.0 } };
    OTIO_NS::AnyDictionary sized{ { "value", 2.0 }, { "size", OTIO_NS::AnyVector{ int64_t(64), int64_t(48) } } };
    probe(host, "toucan:Invert", none);
    probe(host, "toucan:Pow", pow);
    probe(host, "toucan:Pow", none);          // described default
    probe(host, "toucan:Saturate", saturate);
    probe(host, "toucan:ColorMap", colorMap);
    probe(host, "toucan:ColorConvert", convert);
    probe(host, "toucan:ColorConvert", none); // effective defaults (premult -> 0)
    probe(host, "toucan:Premult", none);
    probe(host, "toucan:Unpremult", none);
    probe(host, "toucan:Blur", blur);         // a convolution node describes itself as the head of a chain
    probe(host, "toucan:Pow", sized);         // output-size override: refuses
    return 0;
}
