// Plug-in module loader: dlopen + the two OpenFX entry symbols, and the recursive
// search for *.ofx modules.  Interface parity: lib/toucanRender/Plugin.h,
// PluginUnix.cpp:18-40, Plugin.cpp:20-53.
#pragma once

#include <OpenFX/ofxCore.h>

#include <filesystem>
#include <vector>

namespace toucan
{
    //! A loaded plug-in module (one .ofx shared object).
    class Plugin
    {
    public:
        Plugin(const std::filesystem::path&);
        ~Plugin();
        Plugin(const Plugin&) = delete;
        Plugin& operator=(const Plugin&) = delete;

        int getCount();
        OfxPlugin* getPlugin(int);

    private:
        void* _handle = nullptr;
        int (*_getNumberOfPlugins)(void) = nullptr;
        OfxPlugin* (*_getPlugin)(int) = nullptr;
    };

    //! Find plug-in modules (extension ".ofx"), descending at most maxDepth directories.
    void findPlugins(
        const std::filesystem::path&,
        std::vector<std::filesystem::path>&,
        int depth = 0,
        int maxDepth = 10);
}
