// The OpenFX host.  All of its state sits in ImageEffectHost::State (the header only shows the interface the
// reference's callers use: construct with a search path, createNode by identifier); the three suites it serves are
// file-local functions below, the property suite comes from PropertySet.cpp.
#include "ImageEffectHost.h"

#include <ftk/Core/Context.h>

#include <cstdarg>
#include <cstring>
#include <iostream>
#include <sstream>

namespace toucan
{
    namespace
    {
        const std::string logPrefix = "toucan::ImageEffectHost";

        // ---- parameter and image-effect suites (ImageEffectHost.cpp:86-97, 245-351 of the reference list the live
        // slots; every other slot stays null here instead of uninitialised) -------------------------------------------
        OfxStatus suite_getPropertySet(OfxImageEffectHandle effectHandle, OfxPropertySetHandle* propHandle)
        {
            auto handle = reinterpret_cast<ImageEffectHandle*>(effectHandle);
            *propHandle = reinterpret_cast<OfxPropertySetHandle>(&handle->plugin->propSet);
            return kOfxStatOK;
        }

        OfxStatus suite_getParamSet(OfxImageEffectHandle effectHandle, OfxParamSetHandle* paramHandle)
        {
            // the parameter set handle IS the effect handle
            *paramHandle = reinterpret_cast<OfxParamSetHandle>(effectHandle);
            return kOfxStatOK;
        }

        OfxStatus suite_paramDefine(OfxParamSetHandle effectHandle, const char* paramType, const char* name, OfxPropertySetHandle* propHandle)
        {
            auto handle = reinterpret_cast<ImageEffectHandle*>(effectHandle);
            handle->plugin->paramTypes[name] = paramType;
            *propHandle = reinterpret_cast<OfxPropertySetHandle>(&handle->plugin->paramDefs[name]);
            return kOfxStatOK;
        }

        OfxStatus suite_paramGetHandle(OfxParamSetHandle effectHandle, const char* name, OfxParamHandle* paramHandle, OfxPropertySetHandle*)
        {
            // A parameter handle is the std::any inside the instance's parameter map; a
            // parameter that is neither described with a usable default nor present in the
            // metadata leaves *paramHandle untouched (callers pre-set it to null).
            auto handle = reinterpret_cast<ImageEffectHandle*>(effectHandle);
            if (handle->instance)
            {
                auto i = handle->instance->params.find(name);
                if (i != handle->instance->params.end())
                {
                    *paramHandle = reinterpret_cast<OfxParamHandle>(&i->second);
                }
            }
            return kOfxStatOK;
        }

        OfxStatus suite_paramGetValue(OfxParamHandle handle, ...)
        {
            // toucan's (non-standard) convention: the out-pointer types follow the DYNAMIC
            // type of the stored value -- bool*, int64_t*, double*, std::string*, or one
            // pointer per element for vectors of double / int64_t.  Anything else (notably a
            // plain `int` default) writes nothing.
            if (!handle)
            {
                return kOfxStatErrBadHandle;
            }
            const std::any& a = *reinterpret_cast<const std::any*>(handle);
            va_list args;
            va_start(args, handle);
            if (auto v = std::any_cast<bool>(&a)) *va_arg(args, bool*) = *v;
            else if (auto v = std::any_cast<int64_t>(&a)) *va_arg(args, int64_t*) = *v;
            else if (auto v = std::any_cast<double>(&a)) *va_arg(args, double*) = *v;
            else if (auto v = std::any_cast<std::string>(&a)) *va_arg(args, std::string*) = *v;
            else if (auto v = std::any_cast<OTIO_NS::AnyVector>(&a))
            {
                if (!v->empty() && v->front().type() == typeid(double))
                {
                    for (const auto& e : *v) *va_arg(args, double*) = std::any_cast<double>(e);
                }
                else if (!v->empty() && v->front().type() == typeid(int64_t))
                {
                    for (const auto& e : *v) *va_arg(args, int64_t*) = std::any_cast<int64_t>(e);
                }
            }
            va_end(args);
            return kOfxStatOK;
        }

        OfxStatus suite_clipDefine(OfxImageEffectHandle effectHandle, const char* name, OfxPropertySetHandle* propHandle)
        {
            auto handle = reinterpret_cast<ImageEffectHandle*>(effectHandle);
            *propHandle = reinterpret_cast<OfxPropertySetHandle>(&handle->plugin->clipPropSets[name]);
            return kOfxStatOK;
        }

        OfxStatus suite_clipGetHandle(OfxImageEffectHandle effectHandle, const char* name, OfxImageClipHandle* clip, OfxPropertySetHandle* propHandle)
        {
            // A clip handle is the image property set of that clip for the current render.
            auto handle = reinterpret_cast<ImageEffectHandle*>(effectHandle);
            *clip = nullptr;
            auto i = handle->instance->images.find(name);
            if (i != handle->instance->images.end())
            {
                *clip = reinterpret_cast<OfxImageClipHandle>(&i->second);
            }
            if (propHandle)
            {
                *propHandle = reinterpret_cast<OfxPropertySetHandle>(&handle->plugin->clipPropSets[name]);
            }
            return kOfxStatOK;
        }

        OfxStatus suite_clipGetImage(OfxImageClipHandle handle, OfxTime, const OfxRectD*, OfxPropertySetHandle* propHandle)
        {
            *propHandle = reinterpret_cast<OfxPropertySetHandle>(handle);
            return handle ? kOfxStatOK : kOfxStatFailed;
        }

        OfxStatus suite_clipReleaseImage(OfxPropertySetHandle)
        {
            return kOfxStatOK;
        }
    }

    struct ImageEffectHost::State
    {
        std::weak_ptr<ftk::Context> context;
        PropertySet hostProps;   // what OfxHost::host points at; carries the back pointer for fetchSuite
        OfxHost ofxHost;
        OfxPropertySuiteV1 propertySuite;
        OfxParameterSuiteV1 parameterSuite;
        OfxImageEffectSuiteV1 effectSuite;
        // nodes keep ImageEffectPlugin& references: the elements must not move
        std::vector<std::unique_ptr<ImageEffectPlugin> > plugins;

        void fillSuites();
        void loadModules(const std::vector<std::filesystem::path>& searchPath);
        static const void* fetchSuite(OfxPropertySetHandle, const char* suiteName, int suiteVersion);
    };

    ImageEffectHost::ImageEffectHost(
        const std::shared_ptr<ftk::Context>& context,
        const std::vector<std::filesystem::path>& searchPath) :
        _state(new State)
    {
        _state->context = context;
        // The OfxHost struct has no user-data slot: the host property set carries the way back to the state.
        _state->hostProps.setPointer("host", 0, _state.get());
        _state->hostProps.setString(kOfxPropName, 0, "toucan.b200");
        _state->hostProps.setString(kOfxImageEffectPropCudaRenderSupported, 0, "true");
        _state->hostProps.setString(kOfxImageEffectPropCudaStreamSupported, 0, "true");
        _state->ofxHost.host = reinterpret_cast<OfxPropertySetHandle>(&_state->hostProps);
        _state->ofxHost.fetchSuite = &State::fetchSuite;
        _state->fillSuites();
        _state->loadModules(searchPath);
    }

    ImageEffectHost::~ImageEffectHost()
    {
        for (const auto& plugin : _state->plugins)
        {
            plugin->ofxPlugin->mainEntry(kOfxActionUnload, nullptr, nullptr, nullptr);
        }
    }

    std::shared_ptr<IImageNode> ImageEffectHost::createNode(
        const OTIO_NS::AnyDictionary& metaData,
        const std::string& name,
        const std::vector<std::shared_ptr<IImageNode> >& inputs)
    {
        for (auto& plugin : _state->plugins)
        {
            if (name == plugin->ofxPlugin->pluginIdentifier)
            {
                return std::make_shared<ImageEffectNode>(*plugin, metaData, name, inputs);
            }
        }
        return nullptr;
    }

    std::vector<std::string> ImageEffectHost::getPluginIdentifiers() const
    {
        std::vector<std::string> out;
        for (const auto& plugin : _state->plugins)
        {
            out.push_back(plugin->ofxPlugin->pluginIdentifier);
        }
        return out;
    }

    void ImageEffectHost::State::fillSuites()
    {
        memset(&propertySuite, 0, sizeof(propertySuite));
        memset(&parameterSuite, 0, sizeof(parameterSuite));
        memset(&effectSuite, 0, sizeof(effectSuite));
        fillPropertySuite(propertySuite);
        parameterSuite.paramDefine = &suite_paramDefine;
        parameterSuite.paramGetHandle = &suite_paramGetHandle;
        parameterSuite.paramGetValue = &suite_paramGetValue;
        effectSuite.getPropertySet = &suite_getPropertySet;
        effectSuite.getParamSet = &suite_getParamSet;
        effectSuite.clipDefine = &suite_clipDefine;
        effectSuite.clipGetHandle = &suite_clipGetHandle;
        effectSuite.clipGetImage = &suite_clipGetImage;
        effectSuite.clipReleaseImage = &suite_clipReleaseImage;
    }

    const void* ImageEffectHost::State::fetchSuite(OfxPropertySetHandle handle, const char* suiteName, int suiteVersion)
    {
        State* state = nullptr;
        reinterpret_cast<PropertySet*>(handle)->getPointer("host", 0, reinterpret_cast<void**>(&state));
        if (!state || suiteVersion != 1)
        {
            return nullptr;
        }
        if (0 == strcmp(suiteName, kOfxPropertySuite)) return &state->propertySuite;
        if (0 == strcmp(suiteName, kOfxParameterSuite)) return &state->parameterSuite;
        if (0 == strcmp(suiteName, kOfxImageEffectSuite)) return &state->effectSuite;
        return nullptr;
    }

    void ImageEffectHost::State::loadModules(const std::vector<std::filesystem::path>& searchPath)
    {
        auto logSystem = context.lock()->getSystem<ftk::LogSystem>();

        std::vector<std::filesystem::path> modules;
        for (const auto& path : searchPath)
        {
            logSystem->print(logPrefix, "Search path: " + path.string());
            findPlugins(path, modules);
        }
        if (modules.empty())
        {
            logSystem->print(logPrefix, "No plugins found", ftk::LogType::Warning);
        }

        // Load: setHost + OfxActionLoad.  OK / ReplyDefault keep the effect, ErrFatal is an
        // error for the whole module, anything else drops the effect.
        for (const auto& path : modules)
        {
            logSystem->print(logPrefix, "Loading: " + path.string());
            try
            {
                auto module = std::make_shared<Plugin>(path);
                const int count = module->getCount();
                for (int i = 0; i < count; ++i)
                {
                    OfxPlugin* ofxPlugin = module->getPlugin(i);
                    if (!ofxPlugin || 0 != strcmp(ofxPlugin->pluginApi, kOfxImageEffectPluginApi))
                    {
                        continue;
                    }
                    ofxPlugin->setHost(&ofxHost);
                    const OfxStatus status = ofxPlugin->mainEntry(kOfxActionLoad, nullptr, nullptr, nullptr);
                    if (kOfxStatOK == status || kOfxStatReplyDefault == status)
                    {
                        auto p = std::make_unique<ImageEffectPlugin>();
                        p->plugin = module;
                        p->ofxPlugin = ofxPlugin;
                        plugins.push_back(std::move(p));
                    }
                    else if (kOfxStatErrFatal == status)
                    {
                        throw std::runtime_error("Fatal error in plugin: " + path.string());
                    }
                }
            }
            catch (const std::exception& e)
            {
                std::cout << "ERROR: " << e.what() << std::endl;
            }
        }

        // Describe, then DescribeInContext once per supported context.
        for (auto& plugin : plugins)
        {
            logSystem->print(logPrefix, std::string("Plugin: ") + plugin->ofxPlugin->pluginIdentifier);
            ImageEffectHandle handle = { plugin.get(), nullptr };
            plugin->ofxPlugin->mainEntry(kOfxActionDescribe, &handle, nullptr, nullptr);
            int contextCount = 0;
            plugin->propSet.getDimension(kOfxImageEffectPropSupportedContexts, &contextCount);
            for (int i = 0; i < contextCount; ++i)
            {
                char* context = nullptr;
                plugin->propSet.getString(kOfxImageEffectPropSupportedContexts, i, &context);
                if (context)
                {
                    PropertySet inArgs;
                    inArgs.setString(kOfxImageEffectPropContext, 0, context);
                    plugin->ofxPlugin->mainEntry(
                        kOfxImageEffectActionDescribeInContext,
                        &handle,
                        reinterpret_cast<OfxPropertySetHandle>(&inArgs),
                        nullptr);
                }
            }
            for (const auto& param : plugin->paramTypes)
            {
                logSystem->print(logPrefix, "    \"" + param.first + "\": " + param.second);
            }
        }
    }
}
