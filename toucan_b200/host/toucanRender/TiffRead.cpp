#include "TiffRead.h"

#include <zlib.h>

#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>

namespace toucan
{
    namespace
    {
        [[noreturn]] void bad(const std::string& what) { throw std::runtime_error("TIFF: " + what); }

        struct File
        {
            const unsigned char* p;
            size_t size;
            bool big = false; // "MM": big-endian

            void need(size_t at, size_t n) const
            {
                if (at + n > size || at + n < at) bad("truncated");
            }
            uint16_t u16(size_t at) const
            {
                need(at, 2);
                return big ? static_cast<uint16_t>((p[at] << 8) | p[at + 1]) : static_cast<uint16_t>(p[at] | (p[at + 1] << 8));
            }
            uint32_t u32(size_t at) const
            {
                need(at, 4);
                return big ? (static_cast<uint32_t>(p[at]) << 24) | (static_cast<uint32_t>(p[at + 1]) << 16) | (static_cast<uint32_t>(p[at + 2]) << 8) | p[at + 3]
                           : static_cast<uint32_t>(p[at]) | (static_cast<uint32_t>(p[at + 1]) << 8) | (static_cast<uint32_t>(p[at + 2]) << 16) |
                        (static_cast<uint32_t>(p[at + 3]) << 24);
            }
        };

        struct Entry
        {
            int type = 0;
            uint32_t count = 0;
            size_t at = 0; // where the values are
        };

        size_t typeSize(int type)
        {
            switch (type)
            {
            case 1: case 2: case 6: case 7: return 1;
            case 3: case 8: return 2;
            case 4: case 9: case 11: return 4;
            case 5: case 10: case 12: return 8;
            default: return 0;
            }
        }

        uint32_t value(const File& f, const Entry& e, uint32_t i = 0)
        {
            if (i >= e.count) bad("bad tag");
            switch (e.type)
            {
            case 1: f.need(e.at + i, 1); return f.p[e.at + i];
            case 3: return f.u16(e.at + 2 * i);
            case 4: return f.u32(e.at + 4 * i);
            default: bad("unexpected tag type");
            }
        }

        // TIFF's LZW: MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, width grows one code early
        void lzw(const unsigned char* in, size_t inSize, std::vector<unsigned char>& out, size_t expected)
        {
            out.clear();
            out.reserve(expected);
            struct Item { int prefix; unsigned char last; unsigned char first; uint16_t length; };
            std::vector<Item> table(4096);
            for (int i = 0; i < 256; ++i) table[static_cast<size_t>(i)] = Item{ -1, static_cast<unsigned char>(i), static_cast<unsigned char>(i), 1 };
            int next = 258, width = 9, prev = -1;
            uint32_t acc = 0;
            int bits = 0;
            size_t pos = 0;
            std::vector<unsigned char> scratch;
            for (;;)
            {
                while (bits < width)
                {
                    if (pos >= inSize) return; // no EndOfInformation: take what there is
                    acc = (acc << 8) | in[pos++];
                    bits += 8;
                }
                const int code = static_cast<int>((acc >> (bits - width)) & ((1u << width) - 1));
                bits -= width;
                if (257 == code) return;
                if (256 == code)
                {
                    next = 258;
                    width = 9;
                    prev = -1;
                    continue;
                }
                if (out.size() >= expected) return;
                int emit = code;
                if (code >= next)
                {
                    if (code != next || prev < 0) bad("bad LZW data");
                    emit = prev; // the KwKwK case: previous string + its first byte
                }
                const Item& item = table[static_cast<size_t>(emit)];
                const size_t start = out.size();
                const size_t len = item.length + (code >= next ? 1u : 0u);
                out.resize(start + len);
                {
                    size_t k = start + item.length;
                    int c = emit;
                    while (c >= 0)
                    {
                        out[--k] = table[static_cast<size_t>(c)].last;
                        c = table[static_cast<size_t>(c)].prefix;
                    }
                    if (code >= next) out[start + len - 1] = item.first;
                }
                if (prev >= 0 && next < 4096)
                {
                    const Item& p = table[static_cast<size_t>(prev)];
                    table[static_cast<size_t>(next)] = Item{ prev, out[start], p.first, static_cast<uint16_t>(p.length + 1) };
                    ++next;
                    if (next + 1 >= (1 << width) && width < 12) ++width; // "early change"
                }
                prev = code;
            }
        }

        void packBits(const unsigned char* in, size_t inSize, std::vector<unsigned char>& out, size_t expected)
        {
            out.clear();
            out.reserve(expected);
            size_t i = 0;
            while (i < inSize && out.size() < expected)
            {
                const int n = static_cast<signed char>(in[i++]);
                if (n >= 0)
                {
                    const size_t len = static_cast<size_t>(n) + 1;
                    if (i + len > inSize) bad("bad PackBits data");
                    out.insert(out.end(), in + i, in + i + len);
                    i += len;
                }
                else if (n != -128)
                {
                    if (i >= inSize) bad("bad PackBits data");
                    out.insert(out.end(), static_cast<size_t>(1 - n), in[i++]);
                }
            }
        }

        template <class T>
        T toUnsigned(float f, float maxv)
        {
            float s = f * maxv + 0.5f;
            if (s < 0.f) s = 0.f;
            if (s > maxv) s = maxv;
            return static_cast<T>(s);
        }
    }

    bool readTiffMemory(const unsigned char* data, size_t size, MediaFrame& out)
    {
        File f{ data, size };
        if (size < 8) bad("not a TIFF file");
        if ('I' == data[0] && 'I' == data[1]) f.big = false;
        else if ('M' == data[0] && 'M' == data[1]) f.big = true;
        else bad("not a TIFF file");
        if (f.u16(2) != 42) bad("not a classic TIFF file (BigTIFF is not supported)");
        const size_t ifd = f.u32(4);
        const int n = f.u16(ifd);
        std::map<int, Entry> tags;
        for (int i = 0; i < n; ++i)
        {
            const size_t at = ifd + 2 + static_cast<size_t>(i) * 12;
            Entry e;
            const int tag = f.u16(at);
            e.type = f.u16(at + 2);
            e.count = f.u32(at + 4);
            const size_t bytes = typeSize(e.type) * e.count;
            e.at = bytes <= 4 ? at + 8 : f.u32(at + 8);
            f.need(e.at, bytes);
            tags[tag] = e;
        }
        auto get = [&](int tag, uint32_t fallback, uint32_t i = 0)
        {
            const auto it = tags.find(tag);
            return it != tags.end() && i < it->second.count ? value(f, it->second, i) : fallback;
        };
        const int width = static_cast<int>(get(256, 0)), height = static_cast<int>(get(257, 0));
        const int samples = static_cast<int>(get(277, 1));
        const int bits = static_cast<int>(get(258, 1));
        const int compression = static_cast<int>(get(259, 1));
        const int photometric = static_cast<int>(get(262, 2));
        const int planar = static_cast<int>(get(284, 1));
        const int predictor = static_cast<int>(get(317, 1));
        const int sampleFormat = static_cast<int>(get(339, 1));
        if (width <= 0 || height <= 0 || width > 65536 || height > 65536) bad("bad image size");
        if (photometric != 2 || (samples != 3 && samples != 4)) bad("only RGB / RGBA images are supported (the effects take RGBA)");
        for (int c = 1; c < samples; ++c)
        {
            if (static_cast<int>(get(258, static_cast<uint32_t>(bits), static_cast<uint32_t>(c))) != bits) bad("mixed sample depths");
        }
        const bool isFloat = 3 == sampleFormat;
        if (!((8 == bits || 16 == bits) && 1 == sampleFormat) && !(32 == bits && isFloat)) bad("only 8/16-bit unsigned and 32-bit float samples are supported");
        if (planar != 1) bad("planar files are not supported");
        if (predictor != 1 && !(2 == predictor && !isFloat)) bad("unsupported predictor");
        // ExtraSamples: 1 = associated alpha, 2 = unassociated (associated on read, like OIIO's reader), 0 / absent = unspecified
        const int extra = (4 == samples && !getKeepUnassociatedAlpha()) ? static_cast<int>(get(338, 1)) : 0;

        const size_t bps = static_cast<size_t>(bits / 8);
        const size_t lineBytes = static_cast<size_t>(width) * samples * bps;
        std::vector<unsigned char> pix(lineBytes * static_cast<size_t>(height));
        std::vector<unsigned char> tmp;
        // one compressed block (a strip or a tile) -> `expected` bytes at dst
        auto decode = [&](size_t at, size_t len, unsigned char* dst, size_t expected)
        {
            f.need(at, len);
            switch (compression)
            {
            case 1:
                if (len < expected) bad("short block");
                memcpy(dst, data + at, expected);
                break;
            case 5:
                lzw(data + at, len, tmp, expected);
                if (tmp.size() < expected) bad("short LZW block");
                memcpy(dst, tmp.data(), expected);
                break;
            case 8: case 32946:
            {
                uLongf got = static_cast<uLongf>(expected);
                if (uncompress(dst, &got, data + at, static_cast<uLong>(len)) != Z_OK || got != expected) bad("bad deflate block");
                break;
            }
            case 32773:
                packBits(data + at, len, tmp, expected);
                if (tmp.size() < expected) bad("short PackBits block");
                memcpy(dst, tmp.data(), expected);
                break;
            default: bad("only uncompressed, LZW, deflate and PackBits files are supported");
            }
        };
        // horizontal differencing: per row of the block, per sample, in the file's byte order
        auto undoPredictor = [&](unsigned char* block, size_t rows, size_t rowPixels)
        {
            // libtiff runs the predictor inside its LZW and deflate codecs only: with any other scheme the tag is inert
            if (predictor != 2 || !(5 == compression || 8 == compression || 32946 == compression)) return;
            const size_t rowBytes = rowPixels * samples * bps;
            for (size_t r = 0; r < rows; ++r)
            {
                unsigned char* line = block + r * rowBytes;
                if (8 == bits)
                {
                    for (size_t i = static_cast<size_t>(samples); i < rowBytes; ++i) line[i] = static_cast<unsigned char>(line[i] + line[i - samples]);
                }
                else
                {
                    for (size_t i = static_cast<size_t>(samples); i < rowPixels * samples; ++i)
                    {
                        unsigned char* a = line + i * 2;
                        const unsigned char* b = line + (i - samples) * 2;
                        const unsigned va = f.big ? (a[0] << 8 | a[1]) : (a[0] | a[1] << 8), vb = f.big ? (b[0] << 8 | b[1]) : (b[0] | b[1] << 8);
                        const unsigned v = (va + vb) & 0xffffu;
                        if (f.big) { a[0] = static_cast<unsigned char>(v >> 8); a[1] = static_cast<unsigned char>(v); }
                        else { a[0] = static_cast<unsigned char>(v); a[1] = static_cast<unsigned char>(v >> 8); }
                    }
                }
            }
        };

        if (tags.count(322))
        {
            // Tiles: TileWidth x TileLength blocks in row-major order, every one stored at full size (the right and bottom
            // edges are padded); the predictor runs along a tile's own rows.
            const uint32_t tw = get(322, 0), th = get(323, 0);
            const auto offsets = tags.find(324), counts = tags.find(325);
            if (0 == tw || 0 == th || tw > 65536 || th > 65536 || offsets == tags.end() || counts == tags.end()) bad("missing tiles");
            const uint32_t across = (static_cast<uint32_t>(width) + tw - 1) / tw, down = (static_cast<uint32_t>(height) + th - 1) / th;
            if (offsets->second.count < across * down || counts->second.count < across * down) bad("missing tiles");
            const size_t tileRowBytes = static_cast<size_t>(tw) * samples * bps;
            std::vector<unsigned char> tile(tileRowBytes * th);
            for (uint32_t ty = 0; ty < down; ++ty)
            {
                for (uint32_t tx = 0; tx < across; ++tx)
                {
                    const uint32_t t = ty * across + tx;
                    decode(value(f, offsets->second, t), value(f, counts->second, t), tile.data(), tile.size());
                    undoPredictor(tile.data(), th, tw);
                    const size_t x0 = static_cast<size_t>(tx) * tw, y0 = static_cast<size_t>(ty) * th;
                    const size_t cols = std::min<size_t>(tw, static_cast<size_t>(width) - x0), rows = std::min<size_t>(th, static_cast<size_t>(height) - y0);
                    for (size_t r = 0; r < rows; ++r)
                    {
                        memcpy(&pix[(y0 + r) * lineBytes + x0 * samples * bps], &tile[r * tileRowBytes], cols * samples * bps);
                    }
                }
            }
        }
        else
        {
            const uint32_t rowsPerStrip = std::min<uint32_t>(get(278, static_cast<uint32_t>(height)), static_cast<uint32_t>(height));
            const auto offsets = tags.find(273), counts = tags.find(279);
            if (offsets == tags.end() || counts == tags.end() || 0 == rowsPerStrip) bad("missing strips");
            const uint32_t strips = (static_cast<uint32_t>(height) + rowsPerStrip - 1) / rowsPerStrip;
            if (offsets->second.count < strips || counts->second.count < strips) bad("missing strips");
            for (uint32_t s = 0; s < strips; ++s)
            {
                const size_t row0 = static_cast<size_t>(s) * rowsPerStrip;
                const size_t rows = std::min<size_t>(rowsPerStrip, static_cast<size_t>(height) - row0);
                unsigned char* dst = &pix[row0 * lineBytes];
                decode(value(f, offsets->second, s), value(f, counts->second, s), dst, rows * lineBytes);
                undoPredictor(dst, rows, static_cast<size_t>(width));
            }
        }

        const size_t count = static_cast<size_t>(width) * height;
        out = MediaFrame();
        out.width = width;
        out.height = height;
        out.nchannels = 4;
        out.format = 8 == bits ? TC_U8 : (16 == bits ? TC_U16 : TC_F32);
        out.owned = std::make_shared<std::vector<unsigned char> >(count * 4 * bps);
        for (size_t i = 0; i < count; ++i)
        {
            const unsigned char* p = &pix[i * samples * bps];
            if (8 == bits)
            {
                unsigned char v[4] = { p[0], p[1], p[2], static_cast<unsigned char>(4 == samples ? p[3] : 255) };
                if (2 == extra && v[3] != 255)
                {
                    const float af = v[3] * (1.0f / 255.0f);
                    for (int c = 0; c < 3; ++c) v[c] = toUnsigned<unsigned char>(af * (v[c] * (1.0f / 255.0f)), 255.f);
                }
                memcpy(out.owned->data() + i * 4, v, 4);
            }
            else if (16 == bits)
            {
                uint16_t v[4] = { 0, 0, 0, 65535 };
                for (int c = 0; c < samples; ++c) v[c] = f.big ? static_cast<uint16_t>(p[c * 2] << 8 | p[c * 2 + 1]) : static_cast<uint16_t>(p[c * 2] | p[c * 2 + 1] << 8);
                if (2 == extra && v[3] != 65535)
                {
                    const float af = v[3] * (1.0f / 65535.0f);
                    for (int c = 0; c < 3; ++c) v[c] = toUnsigned<uint16_t>(af * (v[c] * (1.0f / 65535.0f)), 65535.f);
                }
                memcpy(out.owned->data() + i * 8, v, 8);
            }
            else
            {
                float v[4] = { 0.F, 0.F, 0.F, 1.F };
                for (int c = 0; c < samples; ++c)
                {
                    uint32_t b = f.big ? (static_cast<uint32_t>(p[c * 4]) << 24 | static_cast<uint32_t>(p[c * 4 + 1]) << 16 | static_cast<uint32_t>(p[c * 4 + 2]) << 8 | p[c * 4 + 3])
                                       : (static_cast<uint32_t>(p[c * 4]) | static_cast<uint32_t>(p[c * 4 + 1]) << 8 | static_cast<uint32_t>(p[c * 4 + 2]) << 16 |
                                          static_cast<uint32_t>(p[c * 4 + 3]) << 24);
                    memcpy(&v[c], &b, 4);
                }
                if (2 == extra && v[3] != 1.F)
                {
                    for (int c = 0; c < 3; ++c) v[c] = v[c] * v[3];
                }
                memcpy(out.owned->data() + i * 16, v, 16);
            }
        }
        out.host = out.owned->data();
        return true;
    }

    bool readTiff(const std::string& path, MediaFrame& out)
    {
        std::ifstream file(path, std::ios::binary | std::ios::ate);
        if (!file) return false;
        const std::streamsize n = file.tellg();
        if (n <= 0) return false;
        std::vector<unsigned char> bytes(static_cast<size_t>(n));
        file.seekg(0);
        if (!file.read(reinterpret_cast<char*>(bytes.data()), n)) return false;
        return readTiffMemory(bytes.data(), bytes.size(), out);
    }
}
