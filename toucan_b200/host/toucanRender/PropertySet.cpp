#include "PropertySet.h"

#include <OpenFX/ofxImageEffect.h>

namespace toucan
{
    namespace
    {
        // Shared helpers over one typed table.  Statuses follow the reference's
        // PropertySet: unknown property or bad index/count -> kOfxStatFailed.
        template <class T, class V>
        OfxStatus put(std::map<std::string, std::vector<T> >& table, const char* property, int index, const V& value)
        {
            if (index < 0) return kOfxStatFailed;
            auto& slot = table[property];
            if (static_cast<size_t>(index) >= slot.size()) slot.resize(static_cast<size_t>(index) + 1);
            slot[static_cast<size_t>(index)] = value;
            return kOfxStatOK;
        }

        template <class T>
        const std::vector<T>* lookup(const std::map<std::string, std::vector<T> >& table, const char* property)
        {
            const auto i = table.find(property);
            return i != table.end() ? &i->second : nullptr;
        }

        template <class T>
        std::vector<std::string> names(const std::map<std::string, std::vector<T> >& table)
        {
            std::vector<std::string> out;
            out.reserve(table.size());
            for (const auto& kv : table) out.push_back(kv.first);
            return out;
        }
    }

    OfxStatus PropertySet::setPointer(const char* property, int index, void* value) { return put(_pointers, property, index, value); }
    OfxStatus PropertySet::setString(const char* property, int index, const char* value) { return put(_strings, property, index, std::string(value ? value : "")); }
    OfxStatus PropertySet::setDouble(const char* property, int index, double value) { return put(_doubles, property, index, value); }
    OfxStatus PropertySet::setInt(const char* property, int index, int value) { return put(_ints, property, index, value); }

    OfxStatus PropertySet::setPointerN(const char* property, int count, void* const* value)
    {
        auto& slot = _pointers[property];
        slot.resize(static_cast<size_t>(count > 0 ? count : 0));
        for (int i = 0; i < count; ++i) slot[static_cast<size_t>(i)] = value[i];
        return kOfxStatOK;
    }
    OfxStatus PropertySet::setStringN(const char* property, int count, const char* const* value)
    {
        auto& slot = _strings[property];
        slot.resize(static_cast<size_t>(count > 0 ? count : 0));
        for (int i = 0; i < count; ++i) slot[static_cast<size_t>(i)] = std::string(value[i] ? value[i] : "");
        return kOfxStatOK;
    }
    OfxStatus PropertySet::setDoubleN(const char* property, int count, const double* value)
    {
        auto& slot = _doubles[property];
        slot.resize(static_cast<size_t>(count > 0 ? count : 0));
        for (int i = 0; i < count; ++i) slot[static_cast<size_t>(i)] = value[i];
        return kOfxStatOK;
    }
    OfxStatus PropertySet::setIntN(const char* property, int count, const int* value)
    {
        auto& slot = _ints[property];
        slot.resize(static_cast<size_t>(count > 0 ? count : 0));
        for (int i = 0; i < count; ++i) slot[static_cast<size_t>(i)] = value[i];
        return kOfxStatOK;
    }

    OfxStatus PropertySet::getPointer(const char* property, int index, void** value) const
    {
        const auto* v = lookup(_pointers, property);
        if (!v) return kOfxStatFailed;
        if (index < 0 || static_cast<size_t>(index) >= v->size()) return kOfxStatFailed;
        *value = (*v)[static_cast<size_t>(index)];
        return kOfxStatOK;
    }
    OfxStatus PropertySet::getString(const char* property, int index, char** value) const
    {
        const auto* v = lookup(_strings, property);
        if (!v) return kOfxStatFailed;
        if (index < 0 || static_cast<size_t>(index) >= v->size()) return kOfxStatFailed;
        // the string storage lives as long as the property set and is stable until the
        // property is written again, as OpenFX requires
        *value = const_cast<char*>((*v)[static_cast<size_t>(index)].c_str());
        return kOfxStatOK;
    }
    OfxStatus PropertySet::getDouble(const char* property, int index, double* value) const
    {
        const auto* v = lookup(_doubles, property);
        if (!v) return kOfxStatFailed;
        if (index < 0 || static_cast<size_t>(index) >= v->size()) return kOfxStatFailed;
        *value = (*v)[static_cast<size_t>(index)];
        return kOfxStatOK;
    }
    OfxStatus PropertySet::getInt(const char* property, int index, int* value) const
    {
        const auto* v = lookup(_ints, property);
        if (!v) return kOfxStatFailed;
        if (index < 0 || static_cast<size_t>(index) >= v->size()) return kOfxStatFailed;
        *value = (*v)[static_cast<size_t>(index)];
        return kOfxStatOK;
    }

    OfxStatus PropertySet::getPointerN(const char* property, int count, void** value) const
    {
        const auto* v = lookup(_pointers, property);
        if (!v) return kOfxStatFailed;
        if (count < 0 || static_cast<size_t>(count) != v->size()) return kOfxStatFailed;
        for (int i = 0; i < count; ++i) value[i] = (*v)[static_cast<size_t>(i)];
        return kOfxStatOK;
    }
    OfxStatus PropertySet::getStringN(const char* property, int count, char** value) const
    {
        const auto* v = lookup(_strings, property);
        if (!v) return kOfxStatFailed;
        if (count < 0 || static_cast<size_t>(count) != v->size()) return kOfxStatFailed;
        for (int i = 0; i < count; ++i) value[i] = const_cast<char*>((*v)[static_cast<size_t>(i)].c_str());
        return kOfxStatOK;
    }
    OfxStatus PropertySet::getDoubleN(const char* property, int count, double* value) const
    {
        const auto* v = lookup(_doubles, property);
        if (!v) return kOfxStatFailed;
        if (count < 0 || static_cast<size_t>(count) != v->size()) return kOfxStatFailed;
        for (int i = 0; i < count; ++i) value[i] = (*v)[static_cast<size_t>(i)];
        return kOfxStatOK;
    }
    OfxStatus PropertySet::getIntN(const char* property, int count, int* value) const
    {
        const auto* v = lookup(_ints, property);
        if (!v) return kOfxStatFailed;
        if (count < 0 || static_cast<size_t>(count) != v->size()) return kOfxStatFailed;
        for (int i = 0; i < count; ++i) value[i] = (*v)[static_cast<size_t>(i)];
        return kOfxStatOK;
    }

    OfxStatus PropertySet::reset(const char* property)
    {
        size_t n = 0;
        n += _pointers.erase(property);
        n += _strings.erase(property);
        n += _doubles.erase(property);
        n += _ints.erase(property);
        return n ? kOfxStatOK : kOfxStatFailed;
    }

    OfxStatus PropertySet::getDimension(const char* property, int* count) const
    {
        if (const auto* v = lookup(_pointers, property)) { *count = static_cast<int>(v->size()); return kOfxStatOK; }
        if (const auto* v = lookup(_strings, property)) { *count = static_cast<int>(v->size()); return kOfxStatOK; }
        if (const auto* v = lookup(_doubles, property)) { *count = static_cast<int>(v->size()); return kOfxStatOK; }
        if (const auto* v = lookup(_ints, property)) { *count = static_cast<int>(v->size()); return kOfxStatOK; }
        *count = 0;
        return kOfxStatFailed;
    }

    std::vector<std::string> PropertySet::getPointerProperties() const { return names(_pointers); }
    std::vector<std::string> PropertySet::getStringProperties() const { return names(_strings); }
    std::vector<std::string> PropertySet::getDoubleProperties() const { return names(_doubles); }
    std::vector<std::string> PropertySet::getIntProperties() const { return names(_ints); }

    // The property suite handed to plug-ins: every slot is a thin adaptor from the opaque handle (a PropertySet*) to
    // the typed tables above.  Eighteen functions, all live (ImageEffectHost.cpp:67-84 of the reference fills the same slots).
    void fillPropertySuite(OfxPropertySuiteV1& suite)
    {
        suite.propSetPointer = [](OfxPropertySetHandle h, const char* name, int index, void* value) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->setPointer(name, index, value); };
        suite.propSetString = [](OfxPropertySetHandle h, const char* name, int index, const char* value) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->setString(name, index, value); };
        suite.propSetDouble = [](OfxPropertySetHandle h, const char* name, int index, double value) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->setDouble(name, index, value); };
        suite.propSetInt = [](OfxPropertySetHandle h, const char* name, int index, int value) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->setInt(name, index, value); };
        suite.propSetPointerN = [](OfxPropertySetHandle h, const char* name, int count, void* const* values) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->setPointerN(name, count, values); };
        suite.propSetStringN = [](OfxPropertySetHandle h, const char* name, int count, const char* const* values) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->setStringN(name, count, values); };
        suite.propSetDoubleN = [](OfxPropertySetHandle h, const char* name, int count, const double* values) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->setDoubleN(name, count, values); };
        suite.propSetIntN = [](OfxPropertySetHandle h, const char* name, int count, const int* values) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->setIntN(name, count, values); };
        suite.propGetPointer = [](OfxPropertySetHandle h, const char* name, int index, void** value) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getPointer(name, index, value); };
        suite.propGetString = [](OfxPropertySetHandle h, const char* name, int index, char** value) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getString(name, index, value); };
        suite.propGetDouble = [](OfxPropertySetHandle h, const char* name, int index, double* value) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getDouble(name, index, value); };
        suite.propGetInt = [](OfxPropertySetHandle h, const char* name, int index, int* value) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getInt(name, index, value); };
        suite.propGetPointerN = [](OfxPropertySetHandle h, const char* name, int count, void** values) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getPointerN(name, count, values); };
        suite.propGetStringN = [](OfxPropertySetHandle h, const char* name, int count, char** values) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getStringN(name, count, values); };
        suite.propGetDoubleN = [](OfxPropertySetHandle h, const char* name, int count, double* values) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getDoubleN(name, count, values); };
        suite.propGetIntN = [](OfxPropertySetHandle h, const char* name, int count, int* values) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getIntN(name, count, values); };
        suite.propReset = [](OfxPropertySetHandle h, const char* name) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->reset(name); };
        suite.propGetDimension = [](OfxPropertySetHandle h, const char* name, int* count) -> OfxStatus {
            return reinterpret_cast<PropertySet*>(h)->getDimension(name, count); };
    }

    PropertySet bufToPropSet(const Image& image)
    {
        PropertySet out;
        const ImageSpec& spec = image.spec();
        const int bounds[4] = { 0, 0, spec.width, spec.height };
        out.setIntN(kOfxImagePropBounds, 4, bounds);

        const char* components = "";
        switch (spec.nchannels)
        {
        case 1: components = kOfxImageComponentAlpha; break;
        case 3: components = kOfxImageComponentRGB; break;
        case 4: components = kOfxImageComponentRGBA; break;
        default: break;
        }
        out.setString(kOfxImageEffectPropComponents, 0, components);

        // The reference has no HALF case (PropertySet.cpp:452-466) so half images reach its
        // plug-ins with an empty depth; this host names the depth so half frames can
        // traverse the graph (documented extension).
        const char* depth = "";
        switch (spec.format)
        {
        case TC_U8: depth = kOfxBitDepthByte; break;
        case TC_U16: depth = kOfxBitDepthShort; break;
        case TC_F16: depth = kOfxBitDepthHalf; break;
        case TC_F32: depth = kOfxBitDepthFloat; break;
        default: break;
        }
        out.setString(kOfxImageEffectPropPixelDepth, 0, depth);
        out.setInt(kOfxImagePropRowBytes, 0, static_cast<int>(spec.scanline_bytes()));
        out.setPointer(kOfxImagePropData, 0, image.data());
        return out;
    }
}
