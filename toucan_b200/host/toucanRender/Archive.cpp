#include "Archive.h"

#include <algorithm>
#include <zlib.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <stdexcept>

namespace toucan
{
    namespace
    {
        uint16_t le16(const unsigned char* p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }
        uint32_t le32(const unsigned char* p) { return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) | (static_cast<uint32_t>(p[3]) << 24); }
        uint64_t le64(const unsigned char* p) { return static_cast<uint64_t>(le32(p)) | (static_cast<uint64_t>(le32(p + 4)) << 32); }
    }

    ZipArchive::ZipArchive(const std::filesystem::path& path)
    {
        const int fd = open(path.string().c_str(), O_RDONLY);
        if (fd < 0)
        {
            throw std::runtime_error("Cannot open zip file: " + path.string());
        }
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < 22)
        {
            close(fd);
            throw std::runtime_error("Cannot open zip file: " + path.string());
        }
        _size = static_cast<size_t>(st.st_size);
        void* map = mmap(nullptr, _size, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (MAP_FAILED == map)
        {
            throw std::runtime_error("Cannot memory-map zip file: " + path.string());
        }
        _data = static_cast<const unsigned char*>(map);

        // end of central directory: the last 22 .. 22 + 65535 bytes
        size_t eocd = _size;
        for (size_t i = _size - 22, n = 0; n <= 65535; --i, ++n)
        {
            if (0x06054b50 == le32(_data + i))
            {
                eocd = i;
                break;
            }
            if (0 == i) break;
        }
        if (eocd == _size)
        {
            throw std::runtime_error("Cannot open zip file: " + path.string());
        }
        uint64_t count = le16(_data + eocd + 10);
        uint64_t cdOffset = le32(_data + eocd + 16);
        if (eocd >= 20 && 0x07064b50 == le32(_data + eocd - 20))
        {
            // ZIP64: the locator points at the 64-bit end-of-central-directory record
            const uint64_t eocd64 = le64(_data + eocd - 20 + 8);
            if (eocd64 <= _size && _size - eocd64 >= 56 && 0x06064b50 == le32(_data + eocd64))
            {
                count = le64(_data + eocd64 + 32);
                cdOffset = le64(_data + eocd64 + 48);
            }
        }
        uint64_t p = cdOffset;
        for (uint64_t i = 0; i < count; ++i)
        {
            // every bound below is written as "offset <= size && size - offset >= need": the 64-bit ZIP64 fields of a crafted
            // archive must not wrap the sum
            if (p > _size || _size - p < 46 || le32(_data + p) != 0x02014b50)
            {
                throw std::runtime_error("Corrupt zip central directory: " + path.string());
            }
            Entry e;
            e.method = le16(_data + p + 10);
            e.compressedSize = le32(_data + p + 20);
            e.size = le32(_data + p + 24);
            const size_t nameLen = le16(_data + p + 28), extraLen = le16(_data + p + 30), commentLen = le16(_data + p + 32);
            uint64_t local = le32(_data + p + 42);
            if (_size - p - 46 < nameLen + extraLen)
            {
                throw std::runtime_error("Corrupt zip central directory: " + path.string());
            }
            const std::string name(reinterpret_cast<const char*>(_data + p + 46), nameLen);
            // ZIP64 extended information: the fields that overflowed, in this order
            for (size_t x = 0; x + 4 <= extraLen;)
            {
                const unsigned char* f = _data + p + 46 + nameLen + x;
                const size_t id = le16(f);
                const size_t len = std::min<size_t>(le16(f + 2), extraLen - x - 4); // a field cannot reach past the extra area
                if (0x0001 == id)
                {
                    size_t o = 4;
                    if (0xffffffffu == e.size && o + 8 <= 4 + len) { e.size = le64(f + o); o += 8; }
                    if (0xffffffffu == e.compressedSize && o + 8 <= 4 + len) { e.compressedSize = le64(f + o); o += 8; }
                    if (0xffffffffu == local && o + 8 <= 4 + len) { local = le64(f + o); }
                }
                x += 4 + len;
            }
            if (local > _size || _size - local < 30 || le32(_data + local) != 0x04034b50)
            {
                throw std::runtime_error("Corrupt zip local header: " + name);
            }
            // TimelineWrapper.cpp:205-211: 30 bytes + file name + extra field of the LOCAL header
            e.dataOffset = local + 30 + le16(_data + local + 26) + le16(_data + local + 28);
            if (e.dataOffset > _size || _size - e.dataOffset < e.compressedSize)
            {
                throw std::runtime_error("Corrupt zip entry: " + name);
            }
            // a stored entry IS its data: a size other than the stored size would read past the entry (media of an
            // .otioz are referenced in place with this size, TimelineWrapper.cpp); deflate cannot expand beyond 1032:1
            if ((0 == e.method && e.size != e.compressedSize) || (8 == e.method && e.size / 1032 > e.compressedSize + 1))
            {
                throw std::runtime_error("Corrupt zip entry: " + name);
            }
            _entries[name] = e;
            p += 46 + nameLen + extraLen + commentLen;
        }
    }

    ZipArchive::~ZipArchive()
    {
        if (_data)
        {
            munmap(const_cast<unsigned char*>(_data), _size);
        }
    }

    const ZipArchive::Entry* ZipArchive::find(const std::string& name) const
    {
        const auto i = _entries.find(name);
        return i != _entries.end() ? &i->second : nullptr;
    }

    std::vector<unsigned char> ZipArchive::read(const Entry& e) const
    {
        std::vector<unsigned char> out(e.size);
        if (0 == e.method)
        {
            memcpy(out.data(), _data + e.dataOffset, e.size);
            return out;
        }
        if (e.method != 8)
        {
            throw std::runtime_error("Unsupported zip compression method");
        }
        z_stream z;
        memset(&z, 0, sizeof(z));
        if (inflateInit2(&z, -MAX_WBITS) != Z_OK)
        {
            throw std::runtime_error("Cannot initialise zlib");
        }
        z.next_in = const_cast<unsigned char*>(_data + e.dataOffset);
        z.avail_in = static_cast<uInt>(e.compressedSize);
        z.next_out = out.data();
        z.avail_out = static_cast<uInt>(out.size());
        const int r = inflate(&z, Z_FINISH);
        inflateEnd(&z);
        if (r != Z_STREAM_END)
        {
            throw std::runtime_error("Cannot inflate zip entry");
        }
        return out;
    }
}
