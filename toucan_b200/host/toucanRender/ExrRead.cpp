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
            bool pLinear = false; // "perceptually linear" hint: B44 then codes the samples through a log table
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


        // ---- PIZ: range table + Huffman + 2-D wavelet (OpenEXR ImfPizCompressor / ImfHuf / ImfWav) ----
        constexpr int HUF_ENCBITS = 16, HUF_DECBITS = 14;
        constexpr int HUF_ENCSIZE = (1 << HUF_ENCBITS) + 1, HUF_DECSIZE = 1 << HUF_DECBITS, HUF_DECMASK = HUF_DECSIZE - 1;
        constexpr int SHORT_ZEROCODE_RUN = 59, LONG_ZEROCODE_RUN = 63, SHORTEST_LONG_RUN = 2 + LONG_ZEROCODE_RUN - SHORT_ZEROCODE_RUN;

        struct HufDec
        {
            int len = 0;              // short code: its length
            int lit = 0;              // short code: the symbol; long codes: how many share this prefix
            std::vector<int> longs;   // symbols whose codes are longer than HUF_DECBITS and start with this prefix
        };

        struct BitIn
        {
            const unsigned char* p;
            const unsigned char* end;
            uint64_t c = 0;
            int lc = 0;
            void byte()
            {
                if (p >= end) bad("truncated PIZ data");
                c = (c << 8) | *p++;
                lc += 8;
            }
            uint64_t bits(int n)
            {
                while (lc < n) byte();
                lc -= n;
                return (c >> lc) & ((uint64_t(1) << n) - 1);
            }
        };

        void hufUncompress(const unsigned char* in, size_t inSize, uint16_t* out, size_t count)
        {
            if (0 == inSize)
            {
                if (count) bad("bad PIZ data");
                return;
            }
            if (inSize < 20) bad("truncated PIZ data");
            auto u32 = [in](size_t at) { return uint32_t(in[at]) | uint32_t(in[at + 1]) << 8 | uint32_t(in[at + 2]) << 16 | uint32_t(in[at + 3]) << 24; };
            const uint32_t im = u32(0), iM = u32(4), nBits = u32(12);
            if (im >= uint32_t(HUF_ENCSIZE) || iM >= uint32_t(HUF_ENCSIZE)) bad("bad PIZ Huffman table");
            // code lengths, run-length packed in 6-bit fields
            std::vector<uint64_t> code(HUF_ENCSIZE, 0);
            BitIn table{ in + 20, in + inSize };
            for (uint32_t i = im; i <= iM; ++i)
            {
                const uint64_t l = code[i] = table.bits(6);
                if (LONG_ZEROCODE_RUN == l || l >= uint64_t(SHORT_ZEROCODE_RUN))
                {
                    uint64_t run = LONG_ZEROCODE_RUN == l ? table.bits(8) + SHORTEST_LONG_RUN : l - SHORT_ZEROCODE_RUN + 2;
                    if (i + run > uint64_t(iM) + 1) bad("bad PIZ Huffman table");
                    while (run--) code[i++] = 0;
                    --i;
                }
            }
            const unsigned char* data = table.p;
            if (nBits > 8 * uint64_t(in + inSize - data)) bad("bad PIZ data");
            // canonical codes from the lengths: code[i] = length | code << 6
            {
                uint64_t n[59] = {};
                for (int i = 0; i < HUF_ENCSIZE; ++i) n[code[size_t(i)]] += 1;
                uint64_t c = 0;
                for (int i = 58; i > 0; --i)
                {
                    const uint64_t nc = (c + n[i]) >> 1;
                    n[i] = c;
                    c = nc;
                }
                for (int i = 0; i < HUF_ENCSIZE; ++i)
                {
                    const uint64_t l = code[size_t(i)];
                    if (l > 0) code[size_t(i)] = l | (n[l]++ << 6);
                }
            }
            // decoding table: HUF_DECBITS-bit prefixes
            std::vector<HufDec> dec(HUF_DECSIZE);
            for (uint32_t i = im; i <= iM; ++i)
            {
                const uint64_t c = code[i] >> 6;
                const int l = int(code[i] & 63);
                if (c >> l) bad("bad PIZ Huffman table");
                if (l > HUF_DECBITS)
                {
                    HufDec& d = dec[size_t(c >> (l - HUF_DECBITS))];
                    if (d.len) bad("bad PIZ Huffman table");
                    d.lit++;
                    d.longs.push_back(int(i));
                }
                else if (l)
                {
                    HufDec* d = &dec[size_t(c << (HUF_DECBITS - l))];
                    for (uint64_t k = uint64_t(1) << (HUF_DECBITS - l); k > 0; --k, ++d)
                    {
                        if (d->len || !d->longs.empty()) bad("bad PIZ Huffman table");
                        d->len = l;
                        d->lit = int(i);
                    }
                }
            }
            // decode; symbol iM is the run-length escape (repeat the previous output `next byte` times)
            const unsigned char* p = data;
            const unsigned char* pe = data + (nBits + 7) / 8;
            uint64_t c = 0;
            int lc = 0;
            size_t o = 0;
            auto emit = [&](int symbol)
            {
                if (uint32_t(symbol) == iM)
                {
                    if (lc < 8)
                    {
                        if (p >= in + inSize) bad("truncated PIZ data");
                        c = (c << 8) | *p++;
                        lc += 8;
                    }
                    lc -= 8;
                    size_t run = size_t((c >> lc) & 0xff);
                    if (o + run > count || 0 == o) bad("bad PIZ run");
                    const uint16_t v = out[o - 1];
                    while (run--) out[o++] = v;
                }
                else
                {
                    if (o >= count) bad("bad PIZ data");
                    out[o++] = uint16_t(symbol);
                }
            };
            while (p < pe)
            {
                c = (c << 8) | *p++;
                lc += 8;
                while (lc >= HUF_DECBITS)
                {
                    const HufDec& d = dec[size_t((c >> (lc - HUF_DECBITS)) & HUF_DECMASK)];
                    if (d.len)
                    {
                        lc -= d.len;
                        emit(d.lit);
                    }
                    else
                    {
                        if (d.longs.empty()) bad("bad PIZ code");
                        size_t j = 0;
                        for (; j < d.longs.size(); ++j)
                        {
                            const int l = int(code[size_t(d.longs[j])] & 63);
                            while (lc < l && p < pe)
                            {
                                c = (c << 8) | *p++;
                                lc += 8;
                            }
                            if (lc >= l && (code[size_t(d.longs[j])] >> 6) == ((c >> (lc - l)) & ((uint64_t(1) << l) - 1)))
                            {
                                lc -= l;
                                emit(d.longs[j]);
                                break;
                            }
                        }
                        if (j == d.longs.size()) bad("bad PIZ code");
                    }
                }
            }
            const int pad = int((8 - nBits) & 7);
            c >>= pad;
            lc -= pad;
            while (lc > 0)
            {
                const HufDec& d = dec[size_t((c << (HUF_DECBITS - lc)) & HUF_DECMASK)];
                if (!d.len) bad("bad PIZ code");
                lc -= d.len;
                emit(d.lit);
            }
            if (o != count) bad("bad PIZ data");
        }

        inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
        {
            const int ls = int16_t(l), hi = int16_t(h);
            const int ai = ls + (hi & 1) + (hi >> 1);
            a = uint16_t(int16_t(ai));
            b = uint16_t(int16_t(ai - hi));
        }

        inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
        {
            const int m = l, d = h;
            const int bb = (m - (d >> 1)) & 0xffff;
            a = uint16_t((d + bb - 0x8000) & 0xffff);
            b = uint16_t(bb);
        }

        void wav2Decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx)
        {
            const bool w14 = mx < (1 << 14);
            const int n = nx > ny ? ny : nx;
            int p = 1;
            while (p <= n) p <<= 1;
            p >>= 1;
            int p2 = p;
            p >>= 1;
            auto dec = [w14](uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) { if (w14) wdec14(l, h, a, b); else wdec16(l, h, a, b); };
            while (p >= 1)
            {
                uint16_t* py = in;
                uint16_t* const ey = in + ptrdiff_t(oy) * (ny - p2);
                const ptrdiff_t oy1 = ptrdiff_t(oy) * p, oy2 = ptrdiff_t(oy) * p2, ox1 = ptrdiff_t(ox) * p, ox2 = ptrdiff_t(ox) * p2;
                uint16_t i00, i01, i10, i11;
                for (; py <= ey; py += oy2)
                {
                    uint16_t* px = py;
                    uint16_t* const ex = py + ptrdiff_t(ox) * (nx - p2);
                    for (; px <= ex; px += ox2)
                    {
                        uint16_t* p01 = px + ox1;
                        uint16_t* p10 = px + oy1;
                        uint16_t* p11 = p10 + ox1;
                        dec(*px, *p10, i00, i10);
                        dec(*p01, *p11, i01, i11);
                        dec(i00, i01, *px, *p01);
                        dec(i10, i11, *p10, *p11);
                    }
                    if (nx & p)
                    {
                        uint16_t* p10 = px + oy1;
                        dec(*px, *p10, i00, *p10);
                        *px = i00;
                    }
                }
                if (ny & p)
                {
                    uint16_t* px = py;
                    uint16_t* const ex = py + ptrdiff_t(ox) * (nx - p2);
                    for (; px <= ex; px += ox2)
                    {
                        uint16_t* p01 = px + ox1;
                        dec(*px, *p01, i00, *p01);
                        *px = i00;
                    }
                }
                p2 = p;
                p >>= 1;
            }
        }

        // One PXR24 chunk -> the uncompressed scanline layout.  The chunk is a zlib stream of byte PLANES: for every line and
        // channel, the most significant bytes of all samples, then the next bytes, ...; a sample is the running sum of the
        // differences those bytes spell (OpenEXR ImfPxr24Compressor).  HALF samples are whole (2 planes), FLOAT samples were
        // rounded to 24 bits by the writer (3 planes, low byte zero), UINT samples have 4 planes.
        void unpxr24(const unsigned char* in, size_t inSize, const std::vector<Channel>& channels, int width, int lines, std::vector<unsigned char>& planes,
                     std::vector<unsigned char>& out)
        {
            size_t planeBytes = 0, outBytes = 0;
            for (const Channel& c : channels)
            {
                planeBytes += size_t(width) * (1 == c.type ? 2 : (2 == c.type ? 3 : 4));
                outBytes += size_t(width) * (1 == c.type ? 2 : 4);
            }
            planeBytes *= size_t(lines);
            planes.resize(planeBytes);
            uLongf n = static_cast<uLongf>(planeBytes);
            if (uncompress(planes.data(), &n, in, static_cast<uLong>(inSize)) != Z_OK || n != planeBytes) bad("bad PXR24 data");
            out.resize(outBytes * size_t(lines));
            const unsigned char* p = planes.data();
            unsigned char* w = out.data();
            for (int l = 0; l < lines; ++l)
            {
                for (const Channel& c : channels)
                {
                    const size_t count = size_t(width);
                    uint32_t value = 0;
                    if (1 == c.type)
                    {
                        for (size_t j = 0; j < count; ++j)
                        {
                            value += (uint32_t(p[j]) << 8) | uint32_t(p[count + j]);
                            w[0] = static_cast<unsigned char>(value);
                            w[1] = static_cast<unsigned char>(value >> 8);
                            w += 2;
                        }
                        p += 2 * count;
                    }
                    else
                    {
                        const int nPlanes = 2 == c.type ? 3 : 4;
                        for (size_t j = 0; j < count; ++j)
                        {
                            uint32_t diff = (uint32_t(p[j]) << 24) | (uint32_t(p[count + j]) << 16) | (uint32_t(p[2 * count + j]) << 8);
                            if (4 == nPlanes) diff |= uint32_t(p[3 * count + j]);
                            value += diff;
                            for (int k = 0; k < 4; ++k) w[k] = static_cast<unsigned char>(value >> (8 * k));
                            w += 4;
                        }
                        p += size_t(nPlanes) * count;
                    }
                }
            }
        }

        // One B44 / B44A chunk -> the uncompressed scanline layout.  HALF channels are coded in 4 x 4 blocks of 14 bytes
        // (the first sample, then 6-bit differences down the first column and along the rows, all shifted by a common
        // amount) or, B44A only, 3 bytes for a block of one value; FLOAT and UINT channels are stored as they are.  The
        // chunk holds one channel after the other, each as a plane of `lines` rows (OpenEXR ImfB44Compressor).  Fixed
        // rate and lossy at the writer; decoding is exact.
        void unb44(const unsigned char* in, size_t inSize, const std::vector<Channel>& channels, int width, int lines, std::vector<unsigned char>& out)
        {
            const size_t nx = size_t(width), ny = size_t(lines);
            std::vector<std::vector<uint16_t> > planes(channels.size());
            const unsigned char* p = in;
            const unsigned char* const end = in + inSize;
            auto ordered = [](uint16_t v) -> uint16_t { return (v & 0x8000u) ? uint16_t(v & 0x7fffu) : uint16_t(~v); };
            for (size_t ci = 0; ci < channels.size(); ++ci)
            {
                const Channel& c = channels[ci];
                std::vector<uint16_t>& plane = planes[ci];
                if (c.type != 1)
                {
                    const size_t bytes = nx * ny * 4;
                    if (size_t(end - p) < bytes) bad("truncated B44 data");
                    plane.resize(nx * ny * 2);
                    memcpy(plane.data(), p, bytes);
                    p += bytes;
                    continue;
                }
                if (c.pLinear) bad("B44 channels flagged perceptually linear are not supported");
                plane.resize(nx * ny);
                for (size_t y = 0; y < ny; y += 4)
                {
                    for (size_t x = 0; x < nx; x += 4)
                    {
                        uint16_t s[16];
                        if (end - p < 3) bad("truncated B44 data");
                        if (0xfc == p[2])
                        {
                            const uint16_t v = ordered(uint16_t((p[0] << 8) | p[1]));
                            for (uint16_t& e : s) e = v;
                            p += 3;
                        }
                        else
                        {
                            if (end - p < 14) bad("truncated B44 data");
                            // (a damaged block can name a shift of up to 62: 64-bit arithmetic keeps that defined, the
                            // 16-bit results wrap exactly as they would for a valid one)
                            const unsigned shift = p[2] >> 2;
                            const uint64_t bias = uint64_t(0x20) << shift;
                            // six-bit fields, most significant bit first, starting at bit 6 of byte 2
                            auto field = [&](int k) -> unsigned
                            {
                                const int bit = 22 + 6 * k; // bit offset from the start of the block
                                const int at = bit >> 3; // a field spans at most two of the block's 14 bytes
                                const unsigned w16 = (unsigned(p[at]) << 8) | unsigned(at + 1 < 14 ? p[at + 1] : 0);
                                return (w16 >> (10 - (bit & 7))) & 0x3fu;
                            };
                            uint16_t t[16];
                            t[0] = uint16_t((p[0] << 8) | p[1]);
                            // column 0 downwards (fields 0-2), then each further column from the one on its left
                            t[4] = uint16_t(t[0] + (uint64_t(field(0)) << shift) - bias);
                            t[8] = uint16_t(t[4] + (uint64_t(field(1)) << shift) - bias);
                            t[12] = uint16_t(t[8] + (uint64_t(field(2)) << shift) - bias);
                            for (int col = 1; col < 4; ++col)
                            {
                                for (int row = 0; row < 4; ++row)
                                {
                                    t[row * 4 + col] = uint16_t(t[row * 4 + col - 1] + (uint64_t(field(3 + (col - 1) * 4 + row)) << shift) - bias);
                                }
                            }
                            for (int i = 0; i < 16; ++i) s[i] = ordered(t[i]);
                            p += 14;
                        }
                        for (size_t j = 0; j < 4 && y + j < ny; ++j)
                        {
                            for (size_t i = 0; i < 4 && x + i < nx; ++i) plane[(y + j) * nx + x + i] = s[j * 4 + i];
                        }
                    }
                }
            }
            size_t lineBytes = 0;
            for (const Channel& c : channels) lineBytes += nx * (1 == c.type ? 2 : 4);
            out.resize(lineBytes * ny);
            unsigned char* w = out.data();
            for (size_t y = 0; y < ny; ++y)
            {
                for (size_t ci = 0; ci < channels.size(); ++ci)
                {
                    const size_t halves = nx * (1 == channels[ci].type ? 1 : 2);
                    memcpy(w, planes[ci].data() + y * halves, halves * 2);
                    w += halves * 2;
                }
            }
        }

        // One PIZ chunk -> the uncompressed scanline layout (per line: each channel's samples, little-endian)
        void unpiz(const unsigned char* in, size_t inSize, const std::vector<Channel>& channels, int width, int lines, std::vector<unsigned char>& out)
        {
            size_t total = 0; // in 16-bit units
            for (const Channel& c : channels) total += size_t(width) * lines * (1 == c.type ? 1 : 2);
            std::vector<uint16_t> tmp(total);
            if (inSize < 4) bad("truncated PIZ data");
            const int minNonZero = in[0] | in[1] << 8, maxNonZero = in[2] | in[3] << 8;
            size_t at = 4;
            std::vector<unsigned char> bitmap(8192, 0);
            if (maxNonZero >= 8192) bad("bad PIZ bitmap");
            if (minNonZero <= maxNonZero)
            {
                const size_t n = size_t(maxNonZero - minNonZero + 1);
                if (at + n > inSize) bad("truncated PIZ data");
                memcpy(&bitmap[size_t(minNonZero)], in + at, n);
                at += n;
            }
            std::vector<uint16_t> lut(65536, 0);
            int k = 0;
            for (int i = 0; i < 65536; ++i)
            {
                if (0 == i || (bitmap[size_t(i >> 3)] & (1 << (i & 7)))) lut[size_t(k++)] = uint16_t(i);
            }
            const uint16_t maxValue = uint16_t(k - 1);
            if (at + 4 > inSize) bad("truncated PIZ data");
            const uint32_t length = uint32_t(in[at]) | uint32_t(in[at + 1]) << 8 | uint32_t(in[at + 2]) << 16 | uint32_t(in[at + 3]) << 24;
            at += 4;
            if (at + length > inSize) bad("truncated PIZ data");
            hufUncompress(in + at, length, tmp.data(), total);
            // wavelet: every channel is a (lines x width) plane of `size` interleaved 16-bit words
            std::vector<size_t> start(channels.size());
            size_t offset = 0;
            for (size_t c = 0; c < channels.size(); ++c)
            {
                const int size = 1 == channels[c].type ? 1 : 2;
                start[c] = offset;
                for (int j = 0; j < size; ++j) wav2Decode(tmp.data() + offset + j, width, size, lines, width * size, maxValue);
                offset += size_t(width) * lines * size;
            }
            for (uint16_t& v : tmp) v = lut[v];
            out.resize(total * 2);
            unsigned char* o = out.data();
            for (int y = 0; y < lines; ++y)
            {
                for (size_t c = 0; c < channels.size(); ++c)
                {
                    const size_t n = size_t(width) * (1 == channels[c].type ? 1 : 2);
                    const uint16_t* src = tmp.data() + start[c] + size_t(y) * n;
                    for (size_t i = 0; i < n; ++i)
                    {
                        *o++ = static_cast<unsigned char>(src[i]);
                        *o++ = static_cast<unsigned char>(src[i] >> 8);
                    }
                }
            }
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
        if (version & 0x1800u) bad("deep and multi-part files are not supported"); // 0x800 deep, 0x1000 multipart
        const bool tiled = (version & 0x200u) != 0;

        std::vector<Channel> channels;
        int compression = 0;
        int dw[4] = { 0, 0, -1, -1 };
        uint32_t tileW = 0, tileH = 0;
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
                    c.pLinear = data[r.pos] != 0;
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
            else if ("tiles" == name && "tiledesc" == type)
            {
                // x size, y size, then level mode | rounding mode << 4: every mode stores the full-resolution level first,
                // which is the only one read here
                tileW = r.u32();
                tileH = r.u32();
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
        case 4: linesPerChunk = 32; break;                 // PIZ
        case 5: linesPerChunk = 16; break;                 // PXR24
        case 6: case 7: linesPerChunk = 32; break;         // B44, B44A
        default: bad("only uncompressed, RLE, ZIPS, ZIP, PIZ, PXR24 and B44 files are supported (DWA is not)");
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

        if (tiled && (0 == tileW || 0 == tileH || tileW > 65536 || tileH > 65536)) bad("missing tile description");
        const int tilesX = tiled ? (width + static_cast<int>(tileW) - 1) / static_cast<int>(tileW) : 1;
        const int tilesY = tiled ? (height + static_cast<int>(tileH) - 1) / static_cast<int>(tileH) : 1;
        // scanline files: blocks of linesPerChunk lines; tiled files: the tiles of level 0 (first in the offset table)
        const int chunks = tiled ? tilesX * tilesY : (height + linesPerChunk - 1) / linesPerChunk;
        std::vector<uint64_t> offsets(static_cast<size_t>(chunks));
        for (uint64_t& o : offsets) o = r.u64();

        const size_t pixelBytes = asFloat ? 16 : 8;
        // uninitialised storage: value-initialising (and page-faulting) half a gigabyte twice is most of a naive decode
        std::shared_ptr<unsigned char> pixels(new unsigned char[static_cast<size_t>(width) * height * pixelBytes], std::default_delete<unsigned char[]>());

        // one planar scanline -> interleaved RGBA; typed loops the compiler can keep in registers
        // (`width` below is the number of samples per channel in this line: the frame's width, or a tile's)
        auto interleave = [&, frameWidth = width](const unsigned char* line, unsigned char* dst, int width)
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
                        const unsigned char* in = line + cn->offset / static_cast<size_t>(frameWidth) * static_cast<size_t>(width);
                        for (int x = 0; x < width; ++x)
                        {
                            float v;
                            memcpy(&v, in + static_cast<size_t>(x) * 4, 4);
                            o[static_cast<size_t>(x) * 4] = v;
                        }
                    }
                    else
                    {
                        const unsigned char* in = line + cn->offset / static_cast<size_t>(frameWidth) * static_cast<size_t>(width);
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
                        const unsigned char* in = line + cn->offset / static_cast<size_t>(frameWidth) * static_cast<size_t>(width);
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
                    int x0 = 0, y0 = 0, cols = width, lines = 0;
                    if (tiled)
                    {
                        const int tx = c.i32(), ty = c.i32(), lx = c.i32(), ly = c.i32();
                        if (tx < 0 || tx >= tilesX || ty < 0 || ty >= tilesY || lx != 0 || ly != 0) bad("bad tile");
                        x0 = tx * static_cast<int>(tileW);
                        y0 = ty * static_cast<int>(tileH);
                        cols = std::min(static_cast<int>(tileW), width - x0); // edge tiles are clipped, not padded
                        lines = std::min(static_cast<int>(tileH), height - y0);
                    }
                    else
                    {
                        y0 = c.i32() - dw[1];
                        if (y0 < 0 || y0 >= height) bad("bad chunk");
                        lines = std::min(linesPerChunk, height - y0);
                    }
                    const uint32_t packed = c.u32();
                    c.need(packed);
                    const size_t blockLineBytes = lineBytes / static_cast<size_t>(width) * static_cast<size_t>(cols);
                    const size_t expected = blockLineBytes * static_cast<size_t>(lines);
                    const unsigned char* src = data + c.pos;
                    if (packed < expected)
                    {
                        if (4 == compression)
                        {
                            unpiz(src, packed, channels, cols, lines, raw);
                            if (raw.size() != expected) bad("bad PIZ chunk");
                            src = raw.data();
                        }
                        else if (6 == compression || 7 == compression)
                        {
                            unb44(src, packed, channels, cols, lines, raw);
                            if (raw.size() != expected) bad("bad B44 chunk");
                            src = raw.data();
                        }
                        else if (5 == compression)
                        {
                            unpxr24(src, packed, channels, cols, lines, tmp, raw);
                            if (raw.size() != expected) bad("bad PXR24 chunk");
                            src = raw.data();
                        }
                        else if (1 == compression)
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
                        if (compression < 4)
                        {
                            unpredict(tmp, raw);
                            src = raw.data();
                        }
                    }
                    else if (packed != expected)
                    {
                        bad("bad chunk size");
                    }
                    // (a chunk that did not shrink is stored as it is)
                    for (int l = 0; l < lines; ++l)
                    {
                        interleave(src + blockLineBytes * static_cast<size_t>(l),
                                   pixels.get() + (static_cast<size_t>(y0 + l) * width + static_cast<size_t>(x0)) * pixelBytes, cols);
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
