#include "ImageEffect.h"

#include "StackFusion.h"

#include <toucanRender/Util.h>

#include <algorithm>
#include <cstring>
#include <sstream>

namespace toucan
{
    namespace
    {
        bool hasProperty(const std::vector<std::string>& names, const char* name)
        {
            return std::find(names.begin(), names.end(), name) != names.end();
        }

        // Text form of a parameter value for describe().
        void printAny(std::ostream& os, const std::any& a)
        {
            if (auto v = std::any_cast<bool>(&a)) os << (*v ? "true" : "false");
            else if (auto v = std::any_cast<int64_t>(&a)) os << *v;
            else if (auto v = std::any_cast<int>(&a)) os << "int(" << *v << ")";
            else if (auto v = std::any_cast<double>(&a)) { os.precision(17); os << *v; if (*v == static_cast<int64_t>(*v)) os << ".0"; }
            else if (auto v = std::any_cast<std::string>(&a)) os << '"' << *v << '"';
            else if (auto v = std::any_cast<OTIO_NS::AnyVector>(&a))
            {
                os << "[";
                for (size_t i = 0; i < v->size(); ++i)
                {
                    if (i) os << ",";
                    printAny(os, (*v)[i]);
                }
                os << "]";
            }
            else if (!a.has_value()) os << "null";
            else os << "?";
        }
    }

    ImageEffectNode::ImageEffectNode(
        ImageEffectPlugin& plugin,
        const OTIO_NS::AnyDictionary& metaData,
        const std::string& name,
        const std::vector<std::shared_ptr<IImageNode> >& inputs) :
        IImageNode(name, inputs),
        _effect(plugin),
        _settings(metaData),
        _live(new ImageEffectInstance),
        _self{ &plugin, _live.get() }
    {
        // Harvest described defaults the way the reference does (ImageEffect.cpp:21-51):
        // only element 0 of OfxParamPropDefault, typed by the property table it sits in.
        // NOTE the int case stores a plain `int`, which paramGetValue() never delivers
        // (it matches bool / int64_t / double / string / AnyVector) -- so Integer,
        // Integer2D and Boolean defaults silently fall back to the plug-in's local
        // initialisers.  Kept on purpose: it defines the effective parameter defaults.
        if (!_effect.defaultsHarvested)
        {
            for (const auto& def : _effect.paramDefs)
            {
                if (hasProperty(def.second.getStringProperties(), kOfxParamPropDefault))
                {
                    char* s = nullptr;
                    def.second.getString(kOfxParamPropDefault, 0, &s);
                    if (s) _effect.harvestedDefaults[def.first] = std::string(s);
                }
                if (hasProperty(def.second.getDoubleProperties(), kOfxParamPropDefault))
                {
                    double d = 0.0;
                    def.second.getDouble(kOfxParamPropDefault, 0, &d);
                    _effect.harvestedDefaults[def.first] = d;
                }
                if (hasProperty(def.second.getIntProperties(), kOfxParamPropDefault))
                {
                    int i = 0;
                    def.second.getInt(kOfxParamPropDefault, 0, &i);
                    _effect.harvestedDefaults[def.first] = i;
                }
            }
            _effect.defaultsHarvested = true;
        }
        _live->params = _effect.harvestedDefaults;
        for (const auto& kv : metaData)
        {
            _live->params[kv.first] = kv.second;
        }

        _effect.ofxPlugin->mainEntry(kOfxActionCreateInstance, &_self, nullptr, nullptr);
    }

    ImageEffectNode::~ImageEffectNode()
    {
        _effect.ofxPlugin->mainEntry(kOfxActionDestroyInstance, &_self, nullptr, nullptr);
    }

    const std::map<std::string, std::any>& ImageEffectNode::getParams() const { return _live->params; }

    const OTIO_NS::AnyDictionary& ImageEffectNode::getMetaData() const { return _settings; }

    std::string ImageEffectNode::describe() const
    {
        std::stringstream ss;
        ss << _name << " {";
        bool first = true;
        for (const auto& kv : _settings)
        {
            if (!first) ss << ", ";
            first = false;
            ss << kv.first << "=";
            printAny(ss, kv.second);
        }
        ss << "}";
        return ss.str();
    }

    bool ImageEffectNode::describePointwise(tc_chain& chain) const
    {
        char* context = nullptr;
        _effect.propSet.getString(kOfxImageEffectPropSupportedContexts, 0, &context);
        if (!context || 0 != strcmp(context, kOfxImageEffectContextFilter) || _inputs.size() != 1 || !_inputs[0])
        {
            return false;
        }
        const auto sizeIt = _settings.find("size");
        if (sizeIt != _settings.end() && sizeIt->second.has_value())
        {
            return false; // an output-size override is not a same-position operation
        }
        const int before = chain.n;
        if (before >= TC_CHAIN_MAX)
        {
            return false;
        }
        // The plug-in resolves its parameters exactly as in a real render (same getters, same local initialisers)
        // and appends one tc_chain_op; effects that are not pointwise answer kOfxStatReplyDefault.
        PropertySet args;
        args.setDouble(kOfxPropTime, 0, _time.value());
        const int window[4] = { 0, 0, 0, 0 };
        args.setIntN(kOfxImageEffectPropRenderWindow, 4, window);
        args.setInt(kOfxImageEffectPropCudaEnabled, 0, 1);
        // no kernel is launched in this mode: the stream is passed along if the thread has one, never created for it
        args.setPointer(kOfxImageEffectPropCudaStream, 0, DeviceContext::isBound() ? DeviceContext::stream() : nullptr);
        args.setPointer("ToucanPropPointwiseChain", 0, &chain);
        _live->images.clear();
        const OfxStatus status = _effect.ofxPlugin->mainEntry(
            kOfxImageEffectActionRender,
            &_self,
            reinterpret_cast<OfxPropertySetHandle>(&args),
            nullptr);
        if (status != kOfxStatOK || chain.n != before + 1)
        {
            chain.n = before;
            return false;
        }
        return true;
    }

    bool ImageEffectNode::describeChain(tc_chain& chain, std::shared_ptr<IImageNode>& input) const
    {
        if (!describePointwise(chain))
        {
            return false;
        }
        input = _inputs[0];
        return true;
    }

    bool ImageEffectNode::sizeHint(int& width, int& height) const
    {
        const auto sizeIt = _settings.find("size");
        if (sizeIt != _settings.end() && sizeIt->second.has_value())
        {
            IMATH_NAMESPACE::V2i size(0, 0);
            try
            {
                anyToVec(std::any_cast<OTIO_NS::AnyVector>(sizeIt->second), size);
            }
            catch (const std::exception&)
            {
                return false;
            }
            width = size.x;
            height = size.y;
            return size.x > 0 && size.y > 0;
        }
        char* context = nullptr;
        _effect.propSet.getString(kOfxImageEffectPropSupportedContexts, 0, &context);
        if (context && 0 != strcmp(context, kOfxImageEffectContextGenerator) && !_inputs.empty() && _inputs[0])
        {
            return _inputs[0]->sizeHint(width, height); // filters and transitions render at their (first) input's size
        }
        return false;
    }

    Image ImageEffectNode::exec()
    {
        // Evaluate the inputs the effect's context consumes, then render.
        char* context = nullptr;
        _effect.propSet.getString(kOfxImageEffectPropSupportedContexts, 0, &context);
        std::vector<Image> inputs;
        if (context && 0 == strcmp(context, kOfxImageEffectContextFilter))
        {
            Image fused;
            if (tryNodeChain(*this, fused))
            {
                return fused;
            }
            if (!_inputs.empty() && _inputs[0]) inputs.push_back(_inputs[0]->exec());
        }
        else if (context && 0 == strcmp(context, kOfxImageEffectContextTransition))
        {
            if (_inputs.size() > 1 && _inputs[0] && _inputs[1])
            {
                inputs.push_back(_inputs[0]->exec());
                inputs.push_back(_inputs[1]->exec());
            }
        }
        return render(inputs);
    }

    Image ImageEffectNode::render(const std::vector<Image>& inputs)
    {
        Image out;

        // Output size override: read straight from the node metadata (ImageEffect.cpp:84-88).
        IMATH_NAMESPACE::V2i size(0, 0);
        const auto sizeIt = _settings.find("size");
        if (sizeIt != _settings.end() && sizeIt->second.has_value())
        {
            anyToVec(std::any_cast<OTIO_NS::AnyVector>(sizeIt->second), size);
        }

        char* context = nullptr;
        _effect.propSet.getString(kOfxImageEffectPropSupportedContexts, 0, &context);
        if (!context)
        {
            return out;
        }
        _live->images.clear();
        if (0 == strcmp(context, kOfxImageEffectContextGenerator))
        {
            // generators are hard-wired to 4-channel UINT8 (ImageEffect.cpp:93)
            out = Image(ImageSpec(size.x, size.y, 4));
            _live->images["Output"] = bufToPropSet(out);
        }
        else if (0 == strcmp(context, kOfxImageEffectContextFilter) && inputs.size() >= 1)
        {
            ImageSpec spec = inputs[0].spec();
            if (spec.width <= 0 || spec.height <= 0)
            {
                // No image at this time (missing media, a read past the available range).  The reference would give a
                // sized effect (Crop, Resize: "size" in the metadata) an output of that size with the empty input's ZERO
                // channels and call the plug-in on it -- nothing defined comes out.  Here the absence propagates: an
                // effect over no image yields no image (the oracle's choice as well), never an uninitialised frame.
                return out;
            }
            if (size.x > 0 && size.y > 0)
            {
                spec.width = size.x;
                spec.height = size.y;
            }
            out = Image(spec);
            _live->images["Source"] = bufToPropSet(inputs[0]);
            _live->images["Output"] = bufToPropSet(out);
        }
        else if (0 == strcmp(context, kOfxImageEffectContextTransition) && inputs.size() >= 2)
        {
            ImageSpec spec = inputs[0].spec();
            if (spec.width <= 0 || spec.height <= 0)
            {
                return out; // no "from" image: no image (see the filter case)
            }
            if (size.x > 0 && size.y > 0)
            {
                spec.width = size.x;
                spec.height = size.y;
            }
            out = Image(spec);
            _live->images["SourceFrom"] = bufToPropSet(inputs[0]);
            _live->images["SourceTo"] = bufToPropSet(inputs[1]);
            _live->images["Output"] = bufToPropSet(out);
        }

        const ImageSpec& spec = out.spec();
        if (spec.width > 0 && spec.height > 0)
        {
            PropertySet args;
            args.setDouble(kOfxPropTime, 0, _time.value());
            const int window[4] = { 0, 0, spec.width, spec.height };
            args.setIntN(kOfxImageEffectPropRenderWindow, 4, window);
            // OpenFX CUDA render: image data pointers are device pointers, work goes on this stream
            args.setInt(kOfxImageEffectPropCudaEnabled, 0, 1);
            args.setPointer(kOfxImageEffectPropCudaStream, 0, DeviceContext::stream());

            NodeTimers::Scope timer(_name);
            const OfxStatus status = _effect.ofxPlugin->mainEntry(
                kOfxImageEffectActionRender,
                &_self,
                reinterpret_cast<OfxPropertySetHandle>(&args),
                nullptr);
            // The reference ignores the render status (ImageEffect.cpp:145); a device
            // failure must not be silent, so kOfxStatErrFatal/Failed raise here.
            if (status == kOfxStatFailed || status == kOfxStatErrFatal || status == kOfxStatErrMemory)
            {
                std::stringstream ss;
                ss << "Render failed in " << _name << ": " << tc_last_error();
                throw std::runtime_error(ss.str());
            }
        }
        return out;
    }
}
