#include "Plugin.h"

#include "Util.h"

#include <cstring>
#include <utility>

namespace toucan_ofx
{
    namespace
    {
        // Suites are copied BY VALUE at Load: several hosts (one per GPU thread) may load
        // the same module, and a host's suite tables die with it.
        struct Suites
        {
            bool loaded = false;
            OfxPropertySuiteV1 prop;
            OfxParameterSuiteV1 param;
            OfxImageEffectSuiteV1 effect;
        };
        Suites suites;
        OfxHost* host = nullptr;

        const char* contextName(Context c)
        {
            switch (c)
            {
            case Context::Generator: return kOfxImageEffectContextGenerator;
            case Context::Filter: return kOfxImageEffectContextFilter;
            default: return kOfxImageEffectContextTransition;
            }
        }

        OfxStatus load()
        {
            if (!host || !host->fetchSuite)
            {
                return kOfxStatErrMissingHostFeature;
            }
            auto p = static_cast<const OfxPropertySuiteV1*>(host->fetchSuite(host->host, kOfxPropertySuite, 1));
            auto q = static_cast<const OfxParameterSuiteV1*>(host->fetchSuite(host->host, kOfxParameterSuite, 1));
            auto e = static_cast<const OfxImageEffectSuiteV1*>(host->fetchSuite(host->host, kOfxImageEffectSuite, 1));
            if (!p || !q || !e)
            {
                return kOfxStatErrMissingHostFeature;
            }
            suites.prop = *p;
            suites.param = *q;
            suites.effect = *e;
            suites.loaded = true;
            return kOfxStatOK;
        }

        OfxStatus describe(const EffectDef& fx, OfxImageEffectHandle handle)
        {
            OfxPropertySetHandle props = nullptr;
            suites.effect.getPropertySet(handle, &props);
            suites.prop.propSetString(props, kOfxPropLabel, 0, fx.label);
            suites.prop.propSetString(props, kOfxImageEffectPluginPropGrouping, 0, "toucan");
            suites.prop.propSetString(props, kOfxImageEffectPropSupportedContexts, 0, contextName(fx.context));
            // frames arrive as device pointers (OpenFX CUDA render)
            suites.prop.propSetString(props, kOfxImageEffectPropCudaRenderSupported, 0, "true");
            suites.prop.propSetString(props, kOfxImageEffectPropCudaStreamSupported, 0, "true");
            return kOfxStatOK;
        }

        void defineClip(OfxImageEffectHandle handle, const char* name)
        {
            OfxPropertySetHandle props = nullptr;
            suites.effect.clipDefine(handle, name, &props);
            const char* components[] = { kOfxImageComponentAlpha, kOfxImageComponentRGB, kOfxImageComponentRGBA };
            const char* depths[] = { kOfxBitDepthByte, kOfxBitDepthShort, kOfxBitDepthHalf, kOfxBitDepthFloat };
            for (int i = 0; i < 3; ++i) suites.prop.propSetString(props, kOfxImageEffectPropSupportedComponents, i, components[i]);
            for (int i = 0; i < 4; ++i) suites.prop.propSetString(props, kOfxImageEffectPropSupportedPixelDepths, i, depths[i]);
        }

        OfxStatus describeInContext(const EffectDef& fx, OfxImageEffectHandle handle)
        {
            switch (fx.context)
            {
            case Context::Generator: break;
            case Context::Filter: defineClip(handle, "Source"); break;
            case Context::Transition:
                defineClip(handle, "SourceFrom");
                defineClip(handle, "SourceTo");
                break;
            }
            defineClip(handle, "Output");

            OfxParamSetHandle paramSet = nullptr;
            suites.effect.getParamSet(handle, &paramSet);
            for (const ParamDef& p : fx.params)
            {
                OfxPropertySetHandle props = nullptr;
                switch (p.kind)
                {
                case ParamDef::Integer:
                    suites.param.paramDefine(paramSet, kOfxParamTypeInteger, p.name, &props);
                    suites.prop.propSetInt(props, kOfxParamPropDefault, 0, p.i[0]);
                    break;
                case ParamDef::Boolean:
                    suites.param.paramDefine(paramSet, kOfxParamTypeBoolean, p.name, &props);
                    suites.prop.propSetInt(props, kOfxParamPropDefault, 0, p.i[0]);
                    break;
                case ParamDef::Integer2D:
                    suites.param.paramDefine(paramSet, kOfxParamTypeInteger2D, p.name, &props);
                    suites.prop.propSetInt(props, kOfxParamPropDefault, 0, p.i[0]);
                    suites.prop.propSetInt(props, kOfxParamPropDefault, 1, p.i[1]);
                    break;
                case ParamDef::Double:
                    suites.param.paramDefine(paramSet, kOfxParamTypeDouble, p.name, &props);
                    suites.prop.propSetDouble(props, kOfxParamPropDefault, 0, p.d[0]);
                    break;
                case ParamDef::String:
                    suites.param.paramDefine(paramSet, kOfxParamTypeString, p.name, &props);
                    suites.prop.propSetString(props, kOfxParamPropDefault, 0, p.s);
                    break;
                case ParamDef::RGBA:
                    suites.param.paramDefine(paramSet, kOfxParamTypeRGBA, p.name, &props);
                    for (int i = 0; i < 4; ++i) suites.prop.propSetDouble(props, kOfxParamPropDefault, i, p.d[i]);
                    break;
                }
                if (props)
                {
                    suites.prop.propSetString(props, kOfxPropLabel, 0, p.label);
                }
            }
            return kOfxStatOK;
        }

        bool clipImage(OfxImageEffectHandle handle, const char* name, OfxTime time, OfxPropertySetHandle& image, tc_image& out)
        {
            OfxImageClipHandle clip = nullptr;
            image = nullptr;
            suites.effect.clipGetHandle(handle, name, &clip, nullptr);
            if (!clip)
            {
                return false;
            }
            suites.effect.clipGetImage(clip, time, nullptr, &image);
            return propSetToImage(suites.prop, image, out);
        }

        OfxStatus render(const EffectDef& fx, OfxImageEffectHandle handle, OfxPropertySetHandle inArgs)
        {
            RenderArgs args;
            memset(&args, 0, sizeof(args));
            suites.prop.propGetDouble(inArgs, kOfxPropTime, 0, &args.time);
            suites.prop.propGetIntN(inArgs, kOfxImageEffectPropRenderWindow, 4, &args.window.x1);

            // This build renders on the device only: a host that cannot pass device
            // pointers (no OpenFX CUDA render support) is refused, never served from the CPU.
            int cudaEnabled = 0;
            suites.prop.propGetInt(inArgs, kOfxImageEffectPropCudaEnabled, 0, &cudaEnabled);
            if (!cudaEnabled)
            {
                return kOfxStatErrUnsupported;
            }
            void* stream = nullptr;
            suites.prop.propGetPointer(inArgs, kOfxImageEffectPropCudaStream, 0, &stream);
            args.stream = stream;

            // Describe-only pass of the host's pointwise-chain fusion: no images, nothing is launched.
            void* chain = nullptr;
            if (kOfxStatOK == suites.prop.propGetPointer(inArgs, kToucanPropPointwiseChain, 0, &chain) && chain)
            {
                if (!fx.pointwise)
                {
                    return kOfxStatReplyDefault;
                }
                args.chain = static_cast<tc_chain*>(chain);
                OfxParamSetHandle paramSet = nullptr;
                suites.effect.getParamSet(handle, &paramSet);
                const int st = fx.render(args, Params(suites.param, paramSet));
                if (TC_ERR_UNSUPPORTED == st || TC_ERR_INVALID == st)
                {
                    // the per-node path clears the output for such arguments
                    tc_chain_op zero;
                    memset(&zero, 0, sizeof(zero));
                    zero.kind = TC_CHAIN_ZERO;
                    return TC_OK == chainAppend(args, zero) ? kOfxStatOK : kOfxStatReplyDefault;
                }
                return TC_OK == st ? kOfxStatOK : kOfxStatReplyDefault;
            }

            OfxPropertySetHandle images[3] = { nullptr, nullptr, nullptr };
            bool ok = clipImage(handle, "Output", args.time, images[0], args.output);
            if (Context::Filter == fx.context)
            {
                ok = clipImage(handle, "Source", args.time, images[1], args.source[0]) && ok;
            }
            else if (Context::Transition == fx.context)
            {
                ok = clipImage(handle, "SourceFrom", args.time, images[1], args.source[0]) && ok;
                ok = clipImage(handle, "SourceTo", args.time, images[2], args.source[1]) && ok;
            }

            OfxStatus status = kOfxStatOK;
            if (ok)
            {
                OfxParamSetHandle paramSet = nullptr;
                suites.effect.getParamSet(handle, &paramSet);
                const int st = fx.render(args, Params(suites.param, paramSet));
                if (TC_ERR_UNSUPPORTED == st || TC_ERR_INVALID == st)
                {
                    // arguments the reference would have handed to OIIO, which fails and
                    // leaves the output untouched; here the output is cleared to zero
                    tc_memset(args.output.data, 0, tc_image_bytes(args.output.w, args.output.h, args.output.type), args.stream);
                    status = kOfxStatOK;
                }
                else if (TC_OK != st)
                {
                    status = kOfxStatFailed;
                }
            }
            for (OfxPropertySetHandle image : images)
            {
                if (image)
                {
                    suites.effect.clipReleaseImage(image);
                }
            }
            return status;
        }

        OfxStatus entryPoint(size_t index, const char* action, const void* handle, OfxPropertySetHandle inArgs, OfxPropertySetHandle)
        {
            const EffectDef& fx = effects()[index];
            OfxImageEffectHandle effectHandle = reinterpret_cast<OfxImageEffectHandle>(const_cast<void*>(handle));
            if (0 == strcmp(action, kOfxActionLoad)) return load();
            if (0 == strcmp(action, kOfxActionUnload)) return kOfxStatOK;
            if (!suites.loaded) return kOfxStatErrBadHandle;
            if (0 == strcmp(action, kOfxActionDescribe)) return describe(fx, effectHandle);
            if (0 == strcmp(action, kOfxImageEffectActionDescribeInContext)) return describeInContext(fx, effectHandle);
            if (0 == strcmp(action, kOfxActionCreateInstance)) return kOfxStatOK;
            if (0 == strcmp(action, kOfxActionDestroyInstance)) return kOfxStatOK;
            if (0 == strcmp(action, kOfxImageEffectActionRender)) return render(fx, effectHandle, inArgs);
            return kOfxStatReplyDefault;
        }

        // OfxPlugin carries no user pointer, so each table row gets its own pair of
        // trampolines, stamped out at compile time.
        constexpr size_t maxEffects = 8;

        void setHostFunc(OfxHost* value) { host = value; }

        template <size_t I>
        OfxStatus mainEntry(const char* action, const void* handle, OfxPropertySetHandle inArgs, OfxPropertySetHandle outArgs)
        {
            return entryPoint(I, action, handle, inArgs, outArgs);
        }

        template <size_t... I>
        std::vector<OfxPlugin> makePlugins(std::index_sequence<I...>)
        {
            OfxPluginEntryPoint* entries[] = { &mainEntry<I>... };
            std::vector<OfxPlugin> out;
            const auto& table = effects();
            for (size_t i = 0; i < table.size() && i < maxEffects; ++i)
            {
                out.push_back(OfxPlugin{ kOfxImageEffectPluginApi, kOfxImageEffectPluginApiVersion, table[i].identifier, 1, 0,
                                         &setHostFunc, entries[i] });
            }
            return out;
        }

        std::vector<OfxPlugin>& plugins()
        {
            static std::vector<OfxPlugin> p = makePlugins(std::make_index_sequence<maxEffects>());
            return p;
        }
    }

    int chainAppend(const RenderArgs& a, const tc_chain_op& op)
    {
        if (!a.chain || a.chain->n >= TC_CHAIN_MAX)
        {
            return TC_ERR_CUDA;
        }
        a.chain->ops[a.chain->n++] = op;
        return TC_OK;
    }

    // ---- Params ---------------------------------------------------------------------------
    // The host writes through the pointers by the DYNAMIC type of the stored value
    // (lib/toucanRender/ImageEffectHost.cpp:277-322): vectors get one write per element, so
    // every variadic call below passes spare slots behind the ones it uses.
    Params::Params(const OfxParameterSuiteV1& suite, OfxParamSetHandle set) : _suite(suite), _set(set) {}

    OfxParamHandle Params::_handle(const char* name) const
    {
        OfxParamHandle h = nullptr;
        _suite.paramGetHandle(_set, name, &h, nullptr);
        return h;
    }

    namespace
    {
        union Slot { double d; int64_t i; bool b; };
    }

    double Params::getDouble(const char* name, double local) const
    {
        Slot v[8];
        v[0].d = local;
        if (OfxParamHandle h = _handle(name)) _suite.paramGetValue(h, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
        return v[0].d;
    }

    int64_t Params::getInt(const char* name, int64_t local) const
    {
        Slot v[8];
        v[0].i = local;
        if (OfxParamHandle h = _handle(name)) _suite.paramGetValue(h, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
        return v[0].i;
    }

    bool Params::getBool(const char* name, bool local) const
    {
        Slot v[8];
        memset(v, 0, sizeof(v));
        v[0].b = local;
        if (OfxParamHandle h = _handle(name)) _suite.paramGetValue(h, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
        return v[0].b;
    }

    int Params::getBoolAsInt(const char* name, int local) const
    {
        // `int x = local; paramGetValue(h, &x)` with a bool write: only the low byte changes
        Slot v[8];
        memset(v, 0, sizeof(v));
        int x = local;
        memcpy(&v[0], &x, sizeof(x));
        if (OfxParamHandle h = _handle(name)) _suite.paramGetValue(h, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
        memcpy(&x, &v[0], sizeof(x));
        return x;
    }

    std::string Params::getString(const char* name, const std::string& local) const
    {
        // A string parameter is written as a std::string assignment through the pointer
        // (toucan's convention; both sides share one libstdc++).
        std::string out = local;
        if (OfxParamHandle h = _handle(name))
        {
            Slot spare[7];
            _suite.paramGetValue(h, &out, &spare[0], &spare[1], &spare[2], &spare[3], &spare[4], &spare[5], &spare[6]);
        }
        return out;
    }

    void Params::getInt2(const char* name, int64_t out[2]) const
    {
        Slot v[8];
        memset(v, 0, sizeof(v));
        v[0].i = out[0];
        v[1].i = out[1];
        if (OfxParamHandle h = _handle(name)) _suite.paramGetValue(h, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
        out[0] = v[0].i;
        out[1] = v[1].i;
    }

    void Params::getRGBA(const char* name, double out[4]) const
    {
        Slot v[8];
        memset(v, 0, sizeof(v));
        for (int i = 0; i < 4; ++i) v[i].d = out[i];
        if (OfxParamHandle h = _handle(name)) _suite.paramGetValue(h, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
        for (int i = 0; i < 4; ++i) out[i] = v[i].d;
    }
}

extern "C"
{
    OfxExport int OfxGetNumberOfPlugins(void)
    {
        return static_cast<int>(toucan_ofx::plugins().size());
    }

    OfxExport OfxPlugin* OfxGetPlugin(int index)
    {
        auto& p = toucan_ofx::plugins();
        return (index >= 0 && index < static_cast<int>(p.size())) ? &p[static_cast<size_t>(index)] : nullptr;
    }
}
