#include "PngRead.h"

#include <atomic>
#include <cstdlib>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace toucan
{
    // the reader option declared in Read.h (lives here so the readers link on their own, e.g. in tests/cpp/reader_fuzz.cpp)
    namespace
    {
        std::atomic<bool>& keepUnassociatedAlpha()
        {
            static std::atomic<bool> value([] {
                const char* e = getenv("TOUCAN_UNASSOCIATED_ALPHA");
                return e && *e && *e != '0';
            }());
            return value;
        }
    }

    void setKeepUnassociatedAlpha(bool value) { keepUnassociatedAlpha() = value; }
    bool getKeepUnassociatedAlpha() { return keepUnassociatedAlpha(); }

    namespace
    {
        uint32_t be32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

        int paeth(int a, int b, int c)
        {
            const int p = a + b - c;
            const int pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
            if (pa <= pb && pa <= pc) return a;
            return pb <= pc ? b : c;
        }

        // OIIO convert_type float -> unsigned: clamp(f*max + 0.5) truncated
        template <class T>
        T toUnsigned(float f, float maxv)
        {
            float s = f * maxv + 0.5f;
            if (s < 0.f) s = 0.f;
            if (s > maxv) s = maxv;
            return static_cast<T>(s);
        }
    }

    bool readPng(const std::string& path, MediaFrame& out)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        std::vector<unsigned char> file;
        unsigned char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
        fclose(f);
        return readPngMemory(file.data(), file.size(), out);
    }

    bool readPngMemory(const unsigned char* bytes, size_t size, MediaFrame& out)
    {
        // the chunk walk below indexes `file`
        struct View
        {
            const unsigned char* p;
            size_t n;
            const unsigned char* data() const { return p; }
            size_t size() const { return n; }
            const unsigned char& operator[](size_t i) const { return p[i]; }
        };
        const View file{ bytes, size };
        static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
        if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) return false;

        uint32_t width = 0, height = 0;
        int depth = 0, colorType = 0, interlace = 0;
        std::vector<unsigned char> idat, palette, trns;
        size_t pos = 8;
        while (pos + 12 <= file.size())
        {
            const uint32_t len = be32(&file[pos]);
            const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
            const unsigned char* data = &file[pos + 8];
            if (pos + 12 + len > file.size()) return false;
            if (!memcmp(type, "IHDR", 4))
            {
                width = be32(data);
                height = be32(data + 4);
                depth = data[8];
                colorType = data[9];
                interlace = data[12];
            }
            else if (!memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
            else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
            else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
            else if (!memcmp(type, "IEND", 4)) break;
            pos += 12 + len;
        }
        if (!width || !height || width > 65536 || height > 65536 || interlace > 1) return false;
        int channels = 0;
        switch (colorType)
        {
        case 2: channels = 3; break;
        case 6: channels = 4; break;
        case 3: channels = 1; break;
        default: return false; // grey images are 1-channel in the reference: outside the RGBA path
        }
        const bool packed = 3 == colorType && (1 == depth || 2 == depth || 4 == depth); // several palette indices per byte
        if (!(depth == 8 || (depth == 16 && colorType != 3) || packed)) return false;
        const size_t bpp = packed ? 1 : static_cast<size_t>(channels) * depth / 8; // bytes per pixel in `pix`, and the filters' pixel distance
        const size_t stride = bpp * width;
        auto rowBytes = [&](size_t pixels) { return packed ? (pixels * static_cast<size_t>(depth) + 7) / 8 : pixels * bpp; };

        // The image is one filtered sub-image, or -- Adam7 interlace -- seven of them, each covering the pixels
        // (x0 + i dx, y0 + j dy) and filtered on its own; empty passes (narrow images) are not stored at all.
        struct Pass { uint32_t x0, y0, dx, dy; };
        static const Pass adam7[7] = { { 0, 0, 8, 8 }, { 4, 0, 8, 8 }, { 0, 4, 4, 8 }, { 2, 0, 4, 4 }, { 0, 2, 2, 4 }, { 1, 0, 2, 2 }, { 0, 1, 1, 2 } };
        static const Pass whole = { 0, 0, 1, 1 };
        const Pass* passes = interlace ? adam7 : &whole;
        const int nPasses = interlace ? 7 : 1;
        size_t rawBytes = 0;
        for (int k = 0; k < nPasses; ++k)
        {
            const Pass& ps = passes[k];
            const size_t pw = width > ps.x0 ? (width - ps.x0 + ps.dx - 1) / ps.dx : 0, ph = height > ps.y0 ? (height - ps.y0 + ps.dy - 1) / ps.dy : 0;
            if (pw && ph) rawBytes += (rowBytes(pw) + 1) * ph;
        }
        // deflate cannot expand by more than ~1032:1: a header that promises more than the data can hold is damaged
        if (rawBytes / 1032 > idat.size() + 1) return false;
        std::vector<unsigned char> raw(rawBytes);
        uLongf rawLen = static_cast<uLongf>(raw.size());
        if (uncompress(raw.data(), &rawLen, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || rawLen != raw.size()) return false;

        std::vector<unsigned char> pix(stride * height);
        std::vector<unsigned char> sub; // a pass's unfiltered rows (interlaced files only)
        const unsigned char* in = raw.data();
        for (int k = 0; k < nPasses; ++k)
        {
            const Pass& ps = passes[k];
            const size_t pw = width > ps.x0 ? (width - ps.x0 + ps.dx - 1) / ps.dx : 0, ph = height > ps.y0 ? (height - ps.y0 + ps.dy - 1) / ps.dy : 0;
            if (!pw || !ph) continue;
            const size_t pstride = rowBytes(pw);
            unsigned char* rows = pix.data();
            const bool scatter = interlace || packed; // the pass is unfiltered on its own, then its pixels are placed
            if (scatter)
            {
                sub.resize(pstride * ph);
                rows = sub.data();
            }
            // undo the scanline filters
            for (size_t y = 0; y < ph; ++y)
            {
                const int ft = in[0];
                ++in;
                unsigned char* cur = rows + pstride * y;
                const unsigned char* up = y ? rows + pstride * (y - 1) : nullptr;
                for (size_t i = 0; i < pstride; ++i)
                {
                    const int a = i >= bpp ? cur[i - bpp] : 0;
                    const int b = up ? up[i] : 0;
                    const int c = (up && i >= bpp) ? up[i - bpp] : 0;
                    int v = in[i];
                    switch (ft)
                    {
                    case 1: v += a; break;
                    case 2: v += b; break;
                    case 3: v += (a + b) / 2; break;
                    case 4: v += paeth(a, b, c); break;
                    default: break;
                    }
                    cur[i] = static_cast<unsigned char>(v);
                }
                in += pstride;
            }
            if (scatter)
            {
                for (size_t y = 0; y < ph; ++y)
                {
                    for (size_t x = 0; x < pw; ++x)
                    {
                        unsigned char* dst = &pix[(ps.y0 + y * ps.dy) * stride + (ps.x0 + x * ps.dx) * bpp];
                        if (packed)
                        {
                            // indices are packed most significant bits first
                            const size_t bit = x * static_cast<size_t>(depth);
                            *dst = static_cast<unsigned char>((sub[y * pstride + (bit >> 3)] >> (8 - depth - static_cast<int>(bit & 7))) & ((1 << depth) - 1));
                        }
                        else
                        {
                            memcpy(dst, &sub[y * pstride + x * bpp], bpp);
                        }
                    }
                }
            }
        }

        const size_t count = static_cast<size_t>(width) * height;
        const bool associate = !getKeepUnassociatedAlpha(); // OIIO's pnginput associates unless "oiio:UnassociatedAlpha" is set
        out = MediaFrame();
        out.width = static_cast<int>(width);
        out.height = static_cast<int>(height);
        out.nchannels = 4;
        if (depth <= 8)
        {
            out.format = TC_U8;
            out.owned = std::make_shared<std::vector<unsigned char> >(count * 4);
            unsigned char* o = out.owned->data();
            for (size_t i = 0; i < count; ++i)
            {
                unsigned char r, g, b, a = 255;
                if (colorType == 3)
                {
                    const unsigned idx = pix[i];
                    if (idx * 3 + 2 >= palette.size()) return false;
                    r = palette[idx * 3];
                    g = palette[idx * 3 + 1];
                    b = palette[idx * 3 + 2];
                    if (idx < trns.size()) a = trns[idx];
                }
                else
                {
                    const unsigned char* p = &pix[i * channels];
                    r = p[0];
                    g = p[1];
                    b = p[2];
                    if (channels == 4) a = p[3];
                }
                if (a != 255 && associate)
                {
                    const float af = a * (1.0f / 255.0f);
                    r = toUnsigned<unsigned char>(af * (r * (1.0f / 255.0f)), 255.f);
                    g = toUnsigned<unsigned char>(af * (g * (1.0f / 255.0f)), 255.f);
                    b = toUnsigned<unsigned char>(af * (b * (1.0f / 255.0f)), 255.f);
                }
                o[i * 4 + 0] = r;
                o[i * 4 + 1] = g;
                o[i * 4 + 2] = b;
                o[i * 4 + 3] = a;
            }
        }
        else
        {
            out.format = TC_U16;
            out.owned = std::make_shared<std::vector<unsigned char> >(count * 8);
            uint16_t* o = reinterpret_cast<uint16_t*>(out.owned->data());
            for (size_t i = 0; i < count; ++i)
            {
                const unsigned char* p = &pix[i * channels * 2];
                uint16_t v[4] = { 0, 0, 0, 65535 };
                for (int c = 0; c < channels; ++c) v[c] = static_cast<uint16_t>((p[c * 2] << 8) | p[c * 2 + 1]);
                if (v[3] != 65535 && associate)
                {
                    const float af = v[3] * (1.0f / 65535.0f);
                    for (int c = 0; c < 3; ++c) v[c] = toUnsigned<uint16_t>(af * (v[c] * (1.0f / 65535.0f)), 65535.f);
                }
                memcpy(o + i * 4, v, 8);
            }
        }
        out.host = out.owned->data();
        return true;
    }

    bool writePng(const std::string& path, const unsigned char* rgba, int width, int height, bool rgb)
    {
        return writePngFrame(path, rgba, width, height, 8, rgb, false);
    }

    bool writePngFrame(const std::string& path, const void* pixels, int width, int height, int depth, bool rgb, bool unassociate)
    {
        if (depth != 8 && depth != 16) return false;
        const int ch = rgb ? 3 : 4;
        const size_t bps = static_cast<size_t>(depth / 8);
        const size_t stride = static_cast<size_t>(width) * ch * bps + 1;
        const unsigned maxv = 16 == depth ? 65535u : 255u;
        std::vector<unsigned char> raw(stride * height);
        for (int y = 0; y < height; ++y)
        {
            unsigned char* row = &raw[stride * y];
            row[0] = 0;
            for (int x = 0; x < width; ++x)
            {
                unsigned v[4];
                const size_t at = (static_cast<size_t>(y) * width + x) * 4;
                for (int c = 0; c < 4; ++c)
                {
                    v[c] = 16 == depth ? static_cast<const uint16_t*>(pixels)[at + c] : static_cast<const unsigned char*>(pixels)[at + c];
                }
                if (unassociate && !rgb && v[3] != 0 && v[3] != maxv)
                {
                    // PNG stores unassociated alpha: colour * max / alpha, clamped (OIIO's png writer does the same on write)
                    for (int c = 0; c < 3; ++c)
                    {
                        const unsigned long long f = static_cast<unsigned long long>(v[c]) * maxv / v[3];
                        v[c] = f > maxv ? maxv : static_cast<unsigned>(f);
                    }
                }
                for (int c = 0; c < ch; ++c)
                {
                    unsigned char* o = row + 1 + (static_cast<size_t>(x) * ch + c) * bps;
                    if (16 == depth)
                    {
                        o[0] = static_cast<unsigned char>(v[c] >> 8); // PNG samples are big-endian
                        o[1] = static_cast<unsigned char>(v[c]);
                    }
                    else
                    {
                        o[0] = static_cast<unsigned char>(v[c]);
                    }
                }
            }
        }
        uLongf zlen = compressBound(static_cast<uLong>(raw.size()));
        std::vector<unsigned char> z(zlen);
        if (compress2(z.data(), &zlen, raw.data(), static_cast<uLong>(raw.size()), 3) != Z_OK) return false;
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) return false;
        auto chunk = [&](const char* type, const unsigned char* data, uint32_t len) {
            unsigned char hdr[8] = { static_cast<unsigned char>(len >> 24), static_cast<unsigned char>(len >> 16),
                                     static_cast<unsigned char>(len >> 8), static_cast<unsigned char>(len),
                                     static_cast<unsigned char>(type[0]), static_cast<unsigned char>(type[1]),
                                     static_cast<unsigned char>(type[2]), static_cast<unsigned char>(type[3]) };
            fwrite(hdr, 1, 8, f);
            if (len) fwrite(data, 1, len, f);
            uLong crc = crc32(0L, hdr + 4, 4);
            if (len) crc = crc32(crc, data, len);
            const unsigned char c[4] = { static_cast<unsigned char>(crc >> 24), static_cast<unsigned char>(crc >> 16),
                                         static_cast<unsigned char>(crc >> 8), static_cast<unsigned char>(crc) };
            fwrite(c, 1, 4, f);
        };
        static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
        fwrite(sig, 1, 8, f);
        unsigned char ihdr[13] = { static_cast<unsigned char>(width >> 24), static_cast<unsigned char>(width >> 16),
                                   static_cast<unsigned char>(width >> 8), static_cast<unsigned char>(width),
                                   static_cast<unsigned char>(height >> 24), static_cast<unsigned char>(height >> 16),
                                   static_cast<unsigned char>(height >> 8), static_cast<unsigned char>(height),
                                   static_cast<unsigned char>(depth), static_cast<unsigned char>(rgb ? 2 : 6), 0, 0, 0 };
        chunk("IHDR", ihdr, 13);
        chunk("IDAT", z.data(), static_cast<uint32_t>(zlen));
        chunk("IEND", nullptr, 0);
        fclose(f);
        return true;
    }
}
