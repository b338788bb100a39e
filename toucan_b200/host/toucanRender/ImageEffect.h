// An OpenFX effect as a graph node: parameter marshalling, output allocation and the Render action
// (the node of lib/toucanRender/ImageEffect.h:42-60), plus the records the host keeps per loaded effect and per
// live instance.  A plug-in sees none of these types: OfxImageEffectHandle / OfxParamSetHandle are an
// ImageEffectHandle* in disguise, OfxParamHandle is a std::any* into ImageEffectInstance::params, and
// OfxImageClipHandle is the PropertySet* of that clip's image for the current render.
#pragma once

#include <toucanRender/ImageNode.h>
#include <toucanRender/Plugin.h>
#include <toucanRender/PropertySet.h>

#include <OpenFX/ofxImageEffect.h>

#include <any>
#include <map>

namespace toucan
{
    using PropertySetMap = std::map<std::string, PropertySet>;
    using ParamValues = std::map<std::string, std::any>;

    //! What one render call works on: parameter values by name, and one image description per clip name
    //! ("Output", "Source", "SourceFrom", "SourceTo").
    struct ImageEffectInstance
    {
        ParamValues params;
        PropertySetMap images;
    };

    //! One loaded effect: the module it came from, its OfxPlugin entry, and everything its Describe /
    //! DescribeInContext actions declared.
    struct ImageEffectPlugin
    {
        OfxPlugin* ofxPlugin = nullptr;
        std::shared_ptr<Plugin> plugin;                   // keeps the module loaded
        PropertySet propSet;                              // effect descriptor properties
        PropertySetMap clipPropSets;                      // per clip name
        PropertySetMap paramDefs;                         // per parameter name: OfxParamPropDefault and friends
        std::map<std::string, std::string> paramTypes;    // per parameter name: kOfxParamType*
        //! The described defaults as the node constructor harvests them (computed once, on first use; a host and its
        //! plug-in records belong to one thread).
        ParamValues harvestedDefaults;
        bool defaultsHarvested = false;
    };

    //! The object behind the opaque effect handle.  `instance` is null while the effect is being described.
    struct ImageEffectHandle
    {
        ImageEffectPlugin* plugin = nullptr;
        ImageEffectInstance* instance = nullptr;
    };

    class ImageEffectNode : public IImageNode
    {
    public:
        //! `metaData` carries the effect's parameters (OTIO effect metadata or generator parameters); a parameter absent
        //! from it takes the plug-in's described default as the reference delivers it ("effective defaults").
        ImageEffectNode(
            ImageEffectPlugin&,
            const OTIO_NS::AnyDictionary& metaData,
            const std::string& name,
            const std::vector<std::shared_ptr<IImageNode> >& = {});
        virtual ~ImageEffectNode();

        Image exec() override;
        std::string describe() const override;

        //! Render from already evaluated inputs (the per-node OpenFX path).
        Image render(const std::vector<Image>& inputs);

        //! The effect parameters after defaults + metadata (used by the cross-node fusion pass and by tests).
        const ParamValues& getParams() const;
        const OTIO_NS::AnyDictionary& getMetaData() const;

        //! Cross-node fusion of pointwise effects: ask the plug-in to append its pixel operation (with the
        //! parameters as IT resolves them) to `chain` instead of launching it.  False if the effect is not a
        //! same-size, single-input pointwise effect, or the plug-in does not take part.
        bool describePointwise(tc_chain& chain) const;
        bool describeChain(tc_chain& chain, std::shared_ptr<IImageNode>& input) const override;
        bool sizeHint(int& width, int& height) const override;

    private:

        ImageEffectPlugin& _effect;                  // the loaded effect this node instantiates
        OTIO_NS::AnyDictionary _settings;            // the metadata the node was created with
        std::unique_ptr<ImageEffectInstance> _live;  // parameters now, images during a render
        ImageEffectHandle _self;                     // what the plug-in receives as its effect handle
    };
}
