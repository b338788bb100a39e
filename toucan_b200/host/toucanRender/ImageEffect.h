// An OpenFX effect as a graph node: parameter marshalling, output allocation and the
// Render action.  Interface parity: lib/toucanRender/ImageEffect.h:17-60.
#pragma once

#include <toucanRender/ImageNode.h>
#include <toucanRender/Plugin.h>
#include <toucanRender/PropertySet.h>

#include <OpenFX/ofxImageEffect.h>

#include <any>
#include <map>

namespace toucan
{
    //! Image effect plugin.
    struct ImageEffectPlugin
    {
        std::shared_ptr<Plugin> plugin;
        OfxPlugin* ofxPlugin = nullptr;
        PropertySet propSet;
        std::map<std::string, PropertySet> clipPropSets;
        std::map<std::string, std::string> paramTypes;
        std::map<std::string, PropertySet> paramDefs;
        //! The described defaults as the node constructor harvests them (computed once, on first use; a host and its
        //! plug-in records belong to one thread).
        std::map<std::string, std::any> harvestedDefaults;
        bool defaultsHarvested = false;
    };

    //! Image effect instance.
    struct ImageEffectInstance
    {
        std::map<std::string, std::any> params;
        std::map<std::string, PropertySet> images;
    };

    //! Image effect handle.
    struct ImageEffectHandle
    {
        ImageEffectPlugin* plugin = nullptr;
        ImageEffectInstance* instance = nullptr;
    };

    //! Image effect node.
    class ImageEffectNode : public IImageNode
    {
    public:
        ImageEffectNode(
            ImageEffectPlugin&,
            const OTIO_NS::AnyDictionary& metaData,
            const std::string& name,
            const std::vector<std::shared_ptr<IImageNode> >& = {});

        virtual ~ImageEffectNode();

        Image exec() override;

        std::string describe() const override;

        //! The effect parameters after defaults + metadata (used by the cross-node fusion
        //! pass and by tests).
        const std::map<std::string, std::any>& getParams() const;
        const OTIO_NS::AnyDictionary& getMetaData() const;

        //! Render into `out` from already evaluated inputs (the per-node OpenFX path).
        Image render(const std::vector<Image>& inputs);

        //! Cross-node fusion of pointwise effects: ask the plug-in to append its pixel operation (with the
        //! parameters as IT resolves them) to `chain` instead of launching it.  False if the effect is not a
        //! same-size, single-input pointwise effect, or the plug-in does not take part.
        bool describePointwise(tc_chain& chain) const;

    private:
        bool _tryPointwiseChain(Image& out);

        ImageEffectPlugin& _plugin;
        std::unique_ptr<ImageEffectInstance> _instance;
        ImageEffectHandle _handle;
        OTIO_NS::AnyDictionary _metaData;
    };
}
