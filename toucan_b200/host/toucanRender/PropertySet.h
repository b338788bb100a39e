// OpenFX property storage.  Plug-ins never see this class: they reach it through the C property suite
// (fillPropertySuite), with a PropertySet* travelling as the opaque OfxPropertySetHandle.  The host side uses the
// typed accessors directly.  Method names and statuses follow the reference's class of the same name
// (lib/toucanRender/PropertySet.h:16-72) so host code reads the same; the image marshalling is
// PropertySet.cpp:431-471 of the reference restated for device images.
#pragma once

#include <toucanRender/Image.h>

#include <OpenFX/ofxProperty.h>

#include <map>
#include <string>
#include <vector>

namespace toucan
{
    class PropertySet
    {
    public:
        // ---- one value of a property (the array grows to hold `index`) ----
        OfxStatus setInt(const char* property, int index, int value);
        OfxStatus setDouble(const char* property, int index, double value);
        OfxStatus setString(const char* property, int index, const char* value);
        OfxStatus setPointer(const char* property, int index, void* value);
        OfxStatus getInt(const char* property, int index, int* value) const;
        OfxStatus getDouble(const char* property, int index, double* value) const;
        // the returned characters belong to the set and stay valid until the property is written again
        OfxStatus getString(const char* property, int index, char** value) const;
        OfxStatus getPointer(const char* property, int index, void** value) const;

        // ---- the first `count` values of a property ----
        OfxStatus setIntN(const char* property, int count, const int* value);
        OfxStatus setDoubleN(const char* property, int count, const double* value);
        OfxStatus setStringN(const char* property, int count, const char* const* value);
        OfxStatus setPointerN(const char* property, int count, void* const* value);
        OfxStatus getIntN(const char* property, int count, int* value) const;
        OfxStatus getDoubleN(const char* property, int count, double* value) const;
        OfxStatus getStringN(const char* property, int count, char** value) const;
        OfxStatus getPointerN(const char* property, int count, void** value) const;

        // ---- whole properties ----
        OfxStatus getDimension(const char* property, int* count) const;
        OfxStatus reset(const char* property);

        // names present in each typed table (diagnostics, parameter harvesting)
        std::vector<std::string> getIntProperties() const;
        std::vector<std::string> getDoubleProperties() const;
        std::vector<std::string> getStringProperties() const;
        std::vector<std::string> getPointerProperties() const;

    private:
        template <class T>
        using Table = std::map<std::string, std::vector<T> >;
        Table<void*> _pointers;
        Table<std::string> _strings;
        Table<double> _doubles;
        Table<int> _ints;
    };

    //! Fill every slot of the OpenFX property suite with adaptors from the opaque handle to PropertySet.
    void fillPropertySuite(OfxPropertySuiteV1&);

    //! Describe a device image to a plug-in: bounds, components, depth, row bytes and the
    //! DEVICE pointer in OfxImagePropData (OpenFX CUDA render convention).
    PropertySet bufToPropSet(const Image&);
}
