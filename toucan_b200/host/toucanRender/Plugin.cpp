#include "Plugin.h"

#include <dlfcn.h>

#include <sstream>
#include <stdexcept>

namespace toucan
{
    Plugin::Plugin(const std::filesystem::path& path)
    {
        _handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!_handle)
        {
            std::stringstream ss;
            ss << "Cannot load library: " << path.string();
            if (const char* why = dlerror()) ss << " (" << why << ")";
            throw std::runtime_error(ss.str());
        }
        _getNumberOfPlugins = reinterpret_cast<int (*)(void)>(dlsym(_handle, "OfxGetNumberOfPlugins"));
        _getPlugin = reinterpret_cast<OfxPlugin* (*)(int)>(dlsym(_handle, "OfxGetPlugin"));
    }

    Plugin::~Plugin()
    {
        if (_handle)
        {
            dlclose(_handle);
        }
    }

    int Plugin::getCount()
    {
        return _getNumberOfPlugins ? _getNumberOfPlugins() : 0;
    }

    OfxPlugin* Plugin::getPlugin(int index)
    {
        return _getPlugin ? _getPlugin(index) : nullptr;
    }

    void findPlugins(
        const std::filesystem::path& path,
        std::vector<std::filesystem::path>& out,
        int depth,
        int maxDepth)
    {
        std::error_code ec;
        if (depth > maxDepth || !std::filesystem::is_directory(path, ec))
        {
            return;
        }
        for (const auto& entry : std::filesystem::directory_iterator(path, ec))
        {
            if (entry.is_regular_file(ec) && entry.path().extension() == ".ofx")
            {
                out.push_back(entry.path());
            }
            else if (entry.is_directory(ec))
            {
                findPlugins(entry.path(), out, depth + 1, maxDepth);
            }
        }
    }
}
