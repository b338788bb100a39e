#include "ExrRead.h"

#include <zlib.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <thread>

namespace toucan
{
    namespace
    {
        [[noreturn]] void bad(const std::string& what) { throw std::runtime_error("EXR: " + what); }

        struct Reader
        {
            const unsigned char* p;
            size_t size;
            size_t pos = 0;

            void need(size_t n) const
            {
                if (pos + n > size) bad("truncated");
            }
            uint32_t u32()
            {
                need(4);
                const uint32_t v = static_cast<uint32_t>(p[pos]) | (static_cast<uint32_t>(p[pos + 1]) << 8) |
                    (static_cast<uint32_t>(p[pos + 2]) << 16) | (static_cast<uint32_t>(p[pos + 3]) << 24);
                pos += 4;
                return v;
            }
            int32_t i32() { return static_cast<int32_t>(u32()); }
            uint64_t u64()
            {
                const uint64_t lo = u32();
                return lo | (static_cast<uint64_t>(u32()) << 32);
            }
            std::string str()
            {
                std::string s;
                for (;;)
                {
                    need(1);
                    const char c = static_cast<char>(p[pos++]);
                    if (!c) break;
                    s += c;
                    if (s.size() > 255) bad("bad string");
                }
                return s;
            }
        };

        struct Channel
        {
            std::string name;
            int type = 1; // 0 uint, 1 half, 2 float
            size_t offset = 0; // byte offset of this channel's samples inside one scanline
        };

        // undo OpenEXR's byte predictor and the split into even / odd bytes (ImfZip.cpp / ImfRle.cpp)
        void unpredict(std::vector<unsigned char>& t, std::vector<unsigned char>& out)
        {
            const size_t n = t.size();
            for (size_t i = 1; i < n; ++i)
            {
                t[i] = static_cast<unsigned char>(t[i - 1] + t[i] - 128);
            }
            out.resize(n);
            const size_t half = (n + 1) / 2;
            for (size_t i = 0, a = 0, b = half; i < n;)
            {
                out[i++] = t[a++];
                if (i < n) out[i++] = t[b++];
            }
        }

        void unrle(const unsigned char* in, size_t inSize, std::vector<unsigned char>& out, size_t expected)
        {
            out.clear();
            out.reserve(expected);
            size_t i = 0;
            while (i < inSize)
            {
                const int count = static_cast<signed char>(in[i++]);
                if (count < 0)
                {
                    const size_t n = static_cast<size_t>(-count);
                    if (i + n > inSize || out.size() + n > expected) bad("bad RLE data");
                    out.insert(out.end(), in + i, in + i + n);
                    i += n;
                }
                else
                {
                    const size_t n = static_cast<size_t>(count) + 1;
                    if (i >= inSize || out.size() + n > expected) bad("bad RLE data");
                    out.insert(out.end(), n, in[i++]);
                }
            }
            if (out.size() != expected) bad("bad RLE data");
        }

        float halfToFloat(uint16_t h)
        {
            const uint32_t sign = static_cast<uint32_t>(h & 0x8000u) << 16;
            uint32_t exp = (h >> 10) & 0x1fu;
            uint32_t man = h & 0x3ffu;
            uint32_t bits;
            if (0 == exp)
            {
                if (0 == man)
                {
                    bits = sign;
                }
                else
                {
                    // subnormal half -> normal float
                    exp = 127 - 15 + 1;
                    while (!(man & 0x400u))
                    {
                        man <<= 1;
                        --exp;
                    }
                    bits = sign | (exp << 23) | ((man & 0x3ffu) << 13);
                }
            }
            else if (31 == exp)
            {
                bits = sign | 0x7f800000u | (man << 13);
            }
            else
            {
                bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
            }
            float f;
            memcpy(&f, &bits, 4);
            return f;
        }
    }

    bool readExrMemory(const unsigned char* data, size_t size, MediaFrame& out)
    {
        Reader r{ data, size };
        if (r.u32() != 20000630u) bad("not an OpenEXR file");
        const uint32_t version = r.u32();
        if ((version & 0xffu) != 2) bad("unsupported file version");
        if (version & 0x1a00u) bad("tiled, deep and multi-part files are not supported"); // 0x200 tiled, 0x800 deep, 0x1000 multipart

        std::vector<Channel> channels;
        int compression = 0;
        int dw[4] = { 0, 0, -1, -1 };
        bool haveChannels = false, haveWindow = false;
        for (;;)
        {
            const std::string name = r.str();
            if (name.empty()) break;
            const std::string type = r.str();
            const uint32_t len = r.u32();
            r.need(len);
            const size_t end = r.pos + len;
            if ("channels" == name && "chlist" == type)
            {
                for (;;)
                {
                    const std::string ch = r.str();
                    if (ch.empty()) break;
                    Channel c;
                    c.name = ch;
                    c.type = r.i32();
                    r.need(4);
                    r.pos += 4; // pLinear + reserved
                    const int xs = r.i32(), ys = r.i32();
                    if (xs != 1 || ys != 1) bad("sub-sampled channels are not supported");
                    if (c.type < 0 || c.type > 2) bad("bad channel type");
                    channels.push_back(c);
                }
                haveChannels = true;
            }
            else if ("compression" == name)
            {
                compression = data[r.pos];
            }
            else if ("dataWindow" == name && "box2i" == type)
            {
                for (int& v : dw) v = r.i32();
                haveWindow = true;
            }
            r.pos = end;
        }
        if (!haveChannels || !haveWindow) bad("missing header attributes");
        const int width = dw[2] - dw[0] + 1, height = dw[3] - dw[1] + 1;
        if (width <= 0 || height <= 0 || width > 65536 || height > 65536) bad("bad data window");
        int linesPerChunk = 1;
        switch (compression)
        {
        case 0: case 1: case 2: linesPerChunk = 1; break; // none, RLE, ZIPS
        case 3: linesPerChunk = 16; break;                 // ZIP
        default: bad("only uncompressed, RLE, ZIPS and ZIP files are supported (PIZ, PXR24, B44, DWA are not)");
        }

        // scanline layout: channels in file (alphabetical) order, each `width` samples
        size_t lineBytes = 0;
        const Channel* rgba[4] = { nullptr, nullptr, nullptr, nullptr };
        for (Channel& c : channels)
        {
            c.offset = lineBytes;
            lineBytes += static_cast<size_t>(width) * (1 == c.type ? 2 : 4);
            if ("R" == c.name) rgba[0] = &c;
            else if ("G" == c.name) rgba[1] = &c;
            else if ("B" == c.name) rgba[2] = &c;
            else if ("A" == c.name) rgba[3] = &c;
        }
        if (!rgba[0] || !rgba[1] || !rgba[2]) bad("no R, G, B channels (luminance / arbitrary-channel files are not RGBA frames)");
        // OIIO's spec.format is the widest channel type
        bool asFloat = false;
        for (const Channel* c : rgba)
        {
            if (c && 0 == c->type) bad("UINT channels are not supported");
            if (c && 2 == c->type) asFloat = true;
        }

        const int chunks = (height + linesPerChunk - 1) / linesPerChunk;
        std::vector<uint64_t> offsets(static_cast<size_t>(chunks));
        for (uint64_t& o : offsets) o = r.u64();

        const size_t pixelBytes = asFloat ? 16 : 8;
        // uninitialised storage: value-initialising (and page-faulting) half a gigabyte twice is most of a naive decode
        std::shared_ptr<unsigned char> pixels(new unsigned char[static_cast<size_t>(width) * height * pixelBytes], std::default_delete<unsigned char[]>());

        // one planar scanline -> interleaved RGBA; typed loops the compiler can keep in registers
        auto interleave = [&](const unsigned char* line, unsigned char* dst)
        {
            for (int ch = 0; ch < 4; ++ch)
            {
                const Channel* cn = rgba[ch];
                if (asFloat)
                {
                    float* o = reinterpret_cast<float*>(dst) + ch;
                    if (!cn)
                    {
                        for (int x = 0; x < width; ++x) o[static_cast<size_t>(x) * 4] = 1.F; // Read.cpp:86-94: A = 1
                    }
                    else if (2 == cn->type)
                    {
                        const unsigned char* in = line + cn->offset;
                        for (int x = 0; x < width; ++x)
                        {
                            float v;
                            memcpy(&v, in + static_cast<size_t>(x) * 4, 4);
                            o[static_cast<size_t>(x) * 4] = v;
                        }
                    }
                    else
                    {
                        const unsigned char* in = line + cn->offset;
                        for (int x = 0; x < width; ++x)
                        {
                            uint16_t hv;
                            memcpy(&hv, in + static_cast<size_t>(x) * 2, 2);
                            o[static_cast<size_t>(x) * 4] = halfToFloat(hv);
                        }
                    }
                }
                else
                {
                    uint16_t* o = reinterpret_cast<uint16_t*>(dst) + ch;
                    if (!cn)
                    {
                        for (int x = 0; x < width; ++x) o[static_cast<size_t>(x) * 4] = 0x3c00; // half 1.0
                    }
                    else
                    {
                        const unsigned char* in = line + cn->offset;
                        for (int x = 0; x < width; ++x)
                        {
                            uint16_t hv;
                            memcpy(&hv, in + static_cast<size_t>(x) * 2, 2);
                            o[static_cast<size_t>(x) * 4] = hv;
                        }
                    }
                }
            }
        };

        // chunks are independent: decode them on several threads (a 4K float frame is 133 MB, an 8K one 531 MB)
        auto decodeChunks = [&](int first, int step, std::string& error)
        {
            std::vector<unsigned char> tmp, raw;
            try
            {
                for (int k = first; k < chunks; k += step)
                {
                    Reader c{ data, size, static_cast<size_t>(offsets[static_cast<size_t>(k)]) };
                    const int y0 = c.i32() - dw[1];
                    const uint32_t packed = c.u32();
                    c.need(packed);
                    if (y0 < 0 || y0 >= height) bad("bad chunk");
                    const int lines = std::min(linesPerChunk, height - y0);
                    const size_t expected = lineBytes * static_cast<size_t>(lines);
                    const unsigned char* src = data + c.pos;
                    if (packed < expected)
                    {
                        if (1 == compression)
                        {
                            unrle(src, packed, tmp, expected);
                        }
                        else if (compression >= 2)
                        {
                            tmp.resize(expected);
                            uLongf n = static_cast<uLongf>(expected);
                            if (uncompress(tmp.data(), &n, src, packed) != Z_OK || n != expected) bad("bad zlib data");
                        }
                        else
                        {
                            bad("bad chunk size");
                        }
                        unpredict(tmp, raw);
                        src = raw.data();
                    }
                    else if (packed != expected)
                    {
                        bad("bad chunk size");
                    }
                    // (a chunk that did not shrink is stored as it is)
                    for (int l = 0; l < lines; ++l)
                    {
                        interleave(src + lineBytes * static_cast<size_t>(l), pixels.get() + (static_cast<size_t>(y0 + l) * width) * pixelBytes);
                    }
                }
            }
            catch (const std::exception& e)
            {
                error = e.what();
            }
        };
        const size_t totalBytes = static_cast<size_t>(width) * height * pixelBytes;
        const int threads = static_cast<int>(std::max<size_t>(1, std::min<size_t>({ 8, std::thread::hardware_concurrency() / 2, totalBytes >> 22,
                                                                                    static_cast<size_t>(chunks) })));
        std::vector<std::string> errors(static_cast<size_t>(threads));
        if (1 == threads)
        {
            decodeChunks(0, 1, errors[0]);
        }
        else
        {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t) pool.emplace_back(decodeChunks, t, threads, std::ref(errors[static_cast<size_t>(t)]));
            for (auto& t : pool) t.join();
        }
        for (const std::string& e : errors)
        {
            if (!e.empty()) throw std::runtime_error(e);
        }

        out = MediaFrame();
        out.width = width;
        out.height = height;
        out.nchannels = 4;
        out.format = asFloat ? TC_F32 : TC_F16;
        out.storage = pixels;
        out.host = pixels.get();
        return true;
    }

    bool writeExr(const std::string& path, const void* pixels, int width, int height, int type, bool zip)
    {
        if ((type != TC_F16 && type != TC_F32) || width <= 0 || height <= 0) return false;
        const size_t bps = TC_F16 == type ? 2 : 4;
        std::vector<unsigned char> file;
        auto put32 = [&file](uint32_t v) { for (int i = 0; i < 4; ++i) file.push_back(static_cast<unsigned char>(v >> (8 * i))); };
        auto putStr = [&file](const char* s) { file.insert(file.end(), s, s + strlen(s) + 1); };
        auto attr = [&](const char* name, const char* typeName, uint32_t len) { putStr(name); putStr(typeName); put32(len); };
        put32(20000630u);
        put32(2);
        attr("channels", "chlist", 4 * 18 + 1);
        for (const char* ch : { "A", "B", "G", "R" })
        {
            putStr(ch);
            put32(TC_F16 == type ? 1u : 2u);
            put32(0); // pLinear + reserved
            put32(1);
            put32(1);
        }
        file.push_back(0);
        attr("compression", "compression", 1);
        file.push_back(zip ? 3 : 0); // ZIP (blocks of 16 lines) or NONE (one line per chunk)
        for (const char* name : { "dataWindow", "displayWindow" })
        {
            attr(name, "box2i", 16);
            put32(0);
            put32(0);
            put32(static_cast<uint32_t>(width - 1));
            put32(static_cast<uint32_t>(height - 1));
        }
        attr("lineOrder", "lineOrder", 1);
        file.push_back(0);
        const float one = 1.F, zero = 0.F;
        uint32_t oneBits, zeroBits;
        memcpy(&oneBits, &one, 4);
        memcpy(&zeroBits, &zero, 4);
        attr("pixelAspectRatio", "float", 4);
        put32(oneBits);
        attr("screenWindowCenter", "v2f", 8);
        put32(zeroBits);
        put32(zeroBits);
        attr("screenWindowWidth", "float", 4);
        put32(oneBits);
        file.push_back(0);

        const int linesPerChunk = zip ? 16 : 1;
        const int chunks = (height + linesPerChunk - 1) / linesPerChunk;
        const size_t tableAt = file.size();
        file.resize(tableAt + static_cast<size_t>(chunks) * 8);
        const size_t lineBytes = static_cast<size_t>(width) * 4 * bps;
        std::vector<unsigned char> raw, split, packed;
        static const int order[4] = { 3, 2, 1, 0 }; // file order A, B, G, R
        for (int k = 0; k < chunks; ++k)
        {
            const int y0 = k * linesPerChunk, lines = std::min(linesPerChunk, height - y0);
            raw.resize(lineBytes * static_cast<size_t>(lines));
            for (int l = 0; l < lines; ++l)
            {
                const unsigned char* src = static_cast<const unsigned char*>(pixels) + static_cast<size_t>(y0 + l) * lineBytes;
                unsigned char* dst = raw.data() + static_cast<size_t>(l) * lineBytes;
                for (int c = 0; c < 4; ++c)
                {
                    for (int x = 0; x < width; ++x)
                    {
                        memcpy(dst + (static_cast<size_t>(c) * width + x) * bps, src + (static_cast<size_t>(x) * 4 + order[c]) * bps, bps);
                    }
                }
            }
            const size_t n = raw.size();
            if (!zip)
            {
                const uint64_t at = file.size();
                for (int i = 0; i < 8; ++i) file[tableAt + static_cast<size_t>(k) * 8 + static_cast<size_t>(i)] = static_cast<unsigned char>(at >> (8 * i));
                put32(static_cast<uint32_t>(y0));
                put32(static_cast<uint32_t>(n));
                file.insert(file.end(), raw.begin(), raw.end());
                continue;
            }
            // even / odd byte split, then the byte predictor (ImfZip.cpp)
            split.resize(n);
            {
                size_t a = 0, b = (n + 1) / 2;
                for (size_t i = 0; i < n;)
                {
                    split[a++] = raw[i++];
                    if (i < n) split[b++] = raw[i++];
                }
                int prev = split[0];
                for (size_t i = 1; i < n; ++i)
                {
                    const int cur = split[i];
                    split[i] = static_cast<unsigned char>(cur - prev + 128);
                    prev = cur;
                }
            }
            uLongf zlen = compressBound(static_cast<uLong>(n));
            packed.resize(zlen);
            if (compress2(packed.data(), &zlen, split.data(), static_cast<uLong>(n), 4) != Z_OK) return false;
            const bool stored = zlen >= n;
            const uint64_t at = file.size();
            for (int i = 0; i < 8; ++i) file[tableAt + static_cast<size_t>(k) * 8 + static_cast<size_t>(i)] = static_cast<unsigned char>(at >> (8 * i));
            put32(static_cast<uint32_t>(y0));
            put32(static_cast<uint32_t>(stored ? n : zlen));
            if (stored) file.insert(file.end(), raw.begin(), raw.end());
            else file.insert(file.end(), packed.begin(), packed.begin() + static_cast<long>(zlen));
        }
        std::ofstream f(path, std::ios::binary);
        if (!f) return false;
        f.write(reinterpret_cast<const char*>(file.data()), static_cast<std::streamsize>(file.size()));
        return static_cast<bool>(f);
    }

    bool readExr(const std::string& path, MediaFrame& out)
    {
        // map the file: the decoder reads every byte once
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size <= 0)
        {
            close(fd);
            return false;
        }
        const size_t size = static_cast<size_t>(st.st_size);
        void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (MAP_FAILED == map) return false;
        madvise(map, size, MADV_SEQUENTIAL);
        try
        {
            const bool ok = readExrMemory(static_cast<const unsigned char*>(map), size, out);
            munmap(map, size);
            return ok;
        }
        catch (...)
        {
            munmap(map, size);
            throw;
        }
    }
}
