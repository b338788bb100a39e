// The OpenFX host: finds and loads the *.ofx modules under a search path, drives Load / Describe /
// DescribeInContext on every image-effect plug-in in them, serves the property / parameter / image-effect
// suites, and creates effect nodes by plug-in identifier.
// The public interface is the reference's (lib/toucanRender/ImageEffectHost.h:17-30): same constructor,
// same createNode; everything else lives in the implementation file.
#pragma once

#include <toucanRender/ImageEffect.h>

#include <ftk/Core/Context.h>

#include <filesystem>
#include <memory>
#include <string>
#include <vector>

namespace toucan
{
    class ImageEffectHost : public std::enable_shared_from_this<ImageEffectHost>
    {
    public:
        ImageEffectHost(
            const std::shared_ptr<ftk::Context>&,
            const std::vector<std::filesystem::path>& searchPath);
        ~ImageEffectHost();

        ImageEffectHost(const ImageEffectHost&) = delete;
        ImageEffectHost& operator=(const ImageEffectHost&) = delete;

        //! A node running the effect `name` (a plug-in identifier such as "toucan:Blur") with the parameters in
        //! `metaData` on `inputs`; nullptr when no loaded plug-in has that identifier.
        std::shared_ptr<IImageNode> createNode(
            const OTIO_NS::AnyDictionary& metaData,
            const std::string& name,
            const std::vector<std::shared_ptr<IImageNode> >& inputs = {});

        //! Identifiers of the loaded effects, in load order.
        std::vector<std::string> getPluginIdentifiers() const;

    private:
        struct State;
        std::unique_ptr<State> _state;
    };
}
