// Read-only ZIP access for .otioz bundles.  The reference memory-maps the
// archive and hands the STORED media entries to the readers as in-memory files
// (lib/toucanRender/TimelineWrapper.cpp:146-246, through minizip-ng); this is the same with a
// small central-directory parser: the archive is mapped once, content.otio is inflated with
// zlib, and media entries are views into the mapping (they must be stored uncompressed, as
// the reference requires).
#pragma once

#include <cstdint>
#include <filesystem>
#include <map>
#include <string>
#include <vector>

namespace toucan
{
    class ZipArchive
    {
    public:
        struct Entry
        {
            int method = 0;              //!< 0 = stored, 8 = deflate
            uint64_t compressedSize = 0;
            uint64_t size = 0;
            uint64_t dataOffset = 0;     //!< of the entry's bytes inside the archive
        };

        explicit ZipArchive(const std::filesystem::path&);
        ~ZipArchive();
        ZipArchive(const ZipArchive&) = delete;
        ZipArchive& operator=(const ZipArchive&) = delete;

        const Entry* find(const std::string& name) const;
        const std::map<std::string, Entry>& entries() const { return _entries; }

        //! The entry's contents (inflated when needed).
        std::vector<unsigned char> read(const Entry&) const;

        const unsigned char* data() const { return _data; }
        size_t size() const { return _size; }

    private:
        const unsigned char* _data = nullptr;
        size_t _size = 0;
        std::map<std::string, Entry> _entries;
    };
}
