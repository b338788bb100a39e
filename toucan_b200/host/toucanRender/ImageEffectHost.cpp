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
    }

    ImageEffectHost::ImageEffectHost(
        const std::shared_ptr<ftk::Context>& context,
        const std::vector<std::filesystem::path>& searchPath) :
        _context(context)
    {
        // The host property set carries a back pointer so the static fetchSuite can
        // find this object (the OfxHost struct has no user data slot).
        _propSet.setPointer("host", 0, this);
        _propSet.setString(kOfxPropName, 0, "toucan.b200");
        _propSet.setString(kOfxImageEffectPropCudaRenderSupported, 0, "true");
        _propSet.setString(kOfxImageEffectPropCudaStreamSupported, 0, "true");
        _host.host = reinterpret_cast<OfxPropertySetHandle>(&_propSet);
        _host.fetchSuite = &_fetchSuite;
        _suiteInit();
        _pluginInit(searchPath);
    }

    ImageEffectHost::~ImageEffectHost()
    {
        for (const auto& plugin : _plugins)
        {
            plugin->ofxPlugin->mainEntry(kOfxActionUnload, nullptr, nullptr, nullptr);
        }
    }

    std::shared_ptr<IImageNode> ImageEffectHost::createNode(
        const OTIO_NS::AnyDictionary& metaData,
        const std::string& name,
        const std::vector<std::shared_ptr<IImageNode> >& inputs)
    {
        for (auto& plugin : _plugins)
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
        for (const auto& plugin : _plugins)
        {
            out.push_back(plugin->ofxPlugin->pluginIdentifier);
        }
        return out;
    }

    void ImageEffectHost::_suiteInit()
    {
        // Unserved slots are null (the reference leaves them uninitialised).
        memset(&_propertySuite, 0, sizeof(_propertySuite));
        memset(&_parameterSuite, 0, sizeof(_parameterSuite));
        memset(&_effectSuite, 0, sizeof(_effectSuite));

        _propertySuite.propSetPointer = &PropertySet::setPointer;
        _propertySuite.propSetString = &PropertySet::setString;
        _propertySuite.propSetDouble = &PropertySet::setDouble;
        _propertySuite.propSetInt = &PropertySet::setInt;
        _propertySuite.propSetPointerN = &PropertySet::setPointerN;
        _propertySuite.propSetStringN = &PropertySet::setStringN;
        _propertySuite.propSetDoubleN = &PropertySet::setDoubleN;
        _propertySuite.propSetIntN = &PropertySet::setIntN;
        _propertySuite.propGetPointer = &PropertySet::getPointer;
        _propertySuite.propGetString = &PropertySet::getString;
        _propertySuite.propGetDouble = &PropertySet::getDouble;
        _propertySuite.propGetInt = &PropertySet::getInt;
        _propertySuite.propGetPointerN = &PropertySet::getPointerN;
        _propertySuite.propGetStringN = &PropertySet::getStringN;
        _propertySuite.propGetDoubleN = &PropertySet::getDoubleN;
        _propertySuite.propGetIntN = &PropertySet::getIntN;
        _propertySuite.propReset = &PropertySet::reset;
        _propertySuite.propGetDimension = &PropertySet::getDimension;

        _parameterSuite.paramDefine = &_paramDefine;
        _parameterSuite.paramGetHandle = &_paramGetHandle;
        _parameterSuite.paramGetValue = &_paramGetValue;

        _effectSuite.getPropertySet = &_getPropertySet;
        _effectSuite.getParamSet = &_getParamSet;
        _effectSuite.clipDefine = &_clipDefine;
        _effectSuite.clipGetHandle = &_clipGetHandle;
        _effectSuite.clipGetImage = &_clipGetImage;
        _effectSuite.clipReleaseImage = &_clipReleaseImage;
    }

    void ImageEffectHost::_pluginInit(const std::vector<std::filesystem::path>& searchPath)
    {
        auto logSystem = _context.lock()->getSystem<ftk::LogSystem>();

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
                    ofxPlugin->setHost(&_host);
                    const OfxStatus status = ofxPlugin->mainEntry(kOfxActionLoad, nullptr, nullptr, nullptr);
                    if (kOfxStatOK == status || kOfxStatReplyDefault == status)
                    {
                        auto p = std::make_unique<ImageEffectPlugin>();
                        p->plugin = module;
                        p->ofxPlugin = ofxPlugin;
                        _plugins.push_back(std::move(p));
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
        for (auto& plugin : _plugins)
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

    const void* ImageEffectHost::_fetchSuite(OfxPropertySetHandle handle, const char* suiteName, int suiteVersion)
    {
        ImageEffectHost* host = nullptr;
        reinterpret_cast<PropertySet*>(handle)->getPointer("host", 0, reinterpret_cast<void**>(&host));
        if (!host || suiteVersion != 1)
        {
            return nullptr;
        }
        if (0 == strcmp(suiteName, kOfxPropertySuite)) return &host->_propertySuite;
        if (0 == strcmp(suiteName, kOfxParameterSuite)) return &host->_parameterSuite;
        if (0 == strcmp(suiteName, kOfxImageEffectSuite)) return &host->_effectSuite;
        return nullptr;
    }

    OfxStatus ImageEffectHost::_getPropertySet(OfxImageEffectHandle effectHandle, OfxPropertySetHandle* propHandle)
    {
        auto handle = reinterpret_cast<ImageEffectHandle*>(effectHandle);
        *propHandle = reinterpret_cast<OfxPropertySetHandle>(&handle->plugin->propSet);
        return kOfxStatOK;
    }

    OfxStatus ImageEffectHost::_getParamSet(OfxImageEffectHandle effectHandle, OfxParamSetHandle* paramHandle)
    {
        // the parameter set handle IS the effect handle
        *paramHandle = reinterpret_cast<OfxParamSetHandle>(effectHandle);
        return kOfxStatOK;
    }

    OfxStatus ImageEffectHost::_paramDefine(OfxParamSetHandle effectHandle, const char* paramType, const char* name, OfxPropertySetHandle* propHandle)
    {
        auto handle = reinterpret_cast<ImageEffectHandle*>(effectHandle);
        handle->plugin->paramTypes[name] = paramType;
        *propHandle = reinterpret_cast<OfxPropertySetHandle>(&handle->plugin->paramDefs[name]);
        return kOfxStatOK;
    }

    OfxStatus ImageEffectHost::_paramGetHandle(OfxParamSetHandle effectHandle, const char* name, OfxParamHandle* paramHandle, OfxPropertySetHandle*)
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

    OfxStatus ImageEffectHost::_paramGetValue(OfxParamHandle handle, ...)
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

    OfxStatus ImageEffectHost::_clipDefine(OfxImageEffectHandle effectHandle, const char* name, OfxPropertySetHandle* propHandle)
    {
        auto handle = reinterpret_cast<ImageEffectHandle*>(effectHandle);
        *propHandle = reinterpret_cast<OfxPropertySetHandle>(&handle->plugin->clipPropSets[name]);
        return kOfxStatOK;
    }

    OfxStatus ImageEffectHost::_clipGetHandle(OfxImageEffectHandle effectHandle, const char* name, OfxImageClipHandle* clip, OfxPropertySetHandle* propHandle)
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

    OfxStatus ImageEffectHost::_clipGetImage(OfxImageClipHandle handle, OfxTime, const OfxRectD*, OfxPropertySetHandle* propHandle)
    {
        *propHandle = reinterpret_cast<OfxPropertySetHandle>(handle);
        return handle ? kOfxStatOK : kOfxStatFailed;
    }

    OfxStatus ImageEffectHost::_clipReleaseImage(OfxPropertySetHandle)
    {
        return kOfxStatOK;
    }
}
