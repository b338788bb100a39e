#include "JpegRead.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace toucan
{
    namespace
    {
        // zig-zag position -> natural (row-major) position inside the 8x8 block
        const unsigned char zigzag[64] = {
            0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
            35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

        [[noreturn]] void bad(const std::string& what) { throw std::runtime_error("JPEG: " + what); }

        struct HuffmanTable
        {
            bool defined = false;
            unsigned char values[256] = {};
            int maxcode[18] = {};  // largest code of each length, -1 if none
            int valptr[17] = {};   // index of the first value of each length
            int mincode[17] = {};
            // 9-bit prefix lookup: (length << 8) | symbol, 0 = longer than 9 bits
            unsigned short fast[512] = {};

            void build(const unsigned char counts[16], const unsigned char* symbols, int n)
            {
                memcpy(values, symbols, static_cast<size_t>(n));
                int code = 0, k = 0;
                memset(fast, 0, sizeof(fast));
                for (int len = 1; len <= 16; ++len)
                {
                    valptr[len] = k;
                    mincode[len] = code;
                    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code)
                    {
                        if (len <= 9)
                        {
                            const int first = code << (9 - len);
                            for (int j = 0; j < (1 << (9 - len)); ++j)
                            {
                                fast[first + j] = static_cast<unsigned short>((len << 8) | values[k]);
                            }
                        }
                    }
                    maxcode[len] = counts[len - 1] ? code - 1 : -1;
                    if (code > (1 << len)) bad("bad Huffman table");
                    code <<= 1;
                }
                maxcode[17] = 0x7fffffff;
                defined = true;
            }
        };

        // MSB-first bit reader over entropy-coded data: removes the 0x00 stuffed after 0xFF and
        // stops (feeding zero bits) at a marker, which it leaves in place for the caller.
        struct BitReader
        {
            const unsigned char* p = nullptr;
            const unsigned char* end = nullptr;
            uint32_t acc = 0;
            int nbits = 0;
            bool marker = false;

            void fill()
            {
                while (nbits <= 24)
                {
                    unsigned b = 0;
                    if (!marker && p < end)
                    {
                        b = *p;
                        if (0xFF == b)
                        {
                            if (p + 1 < end && 0 == p[1])
                            {
                                p += 2;
                            }
                            else
                            {
                                marker = true;
                                b = 0;
                            }
                        }
                        else
                        {
                            ++p;
                        }
                    }
                    acc |= b << (24 - nbits);
                    nbits += 8;
                }
            }
            int bits(int n)
            {
                if (0 == n) return 0;
                fill();
                const int v = static_cast<int>(acc >> (32 - n));
                acc <<= n;
                nbits -= n;
                return v;
            }
            int bit() { return bits(1); }
            void reset()
            {
                acc = 0;
                nbits = 0;
                marker = false;
            }
            int decode(const HuffmanTable& t)
            {
                fill();
                const unsigned short f = t.fast[acc >> 23];
                if (f)
                {
                    const int len = f >> 8;
                    acc <<= len;
                    nbits -= len;
                    return f & 0xff;
                }
                for (int len = 10; len <= 16; ++len)
                {
                    const int code = static_cast<int>(acc >> (32 - len));
                    if (code <= t.maxcode[len])
                    {
                        acc <<= len;
                        nbits -= len;
                        return t.values[t.valptr[len] + code - t.mincode[len]];
                    }
                }
                bad("bad Huffman code");
            }
        };

        inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

        struct ScanComponent
        {
            JpegComponent* c = nullptr;
            int td = 0, ta = 0;
            int pred = 0;
        };

        struct Decoder
        {
            const unsigned char* data;
            size_t size;
            size_t pos = 0;
            JpegCoefficients& out;
            HuffmanTable dc[4], ac[4];
            bool progressive = false;
            bool sawSof = false, sawJfif = false, sawAdobe = false;
            int adobeTransform = 0;
            int restartInterval = 0;

            Decoder(const unsigned char* d, size_t n, JpegCoefficients& o) : data(d), size(n), out(o) {}

            int u8()
            {
                if (pos >= size) bad("truncated");
                return data[pos++];
            }
            int u16()
            {
                const int a = u8();
                return (a << 8) | u8();
            }

            void dqt(size_t end)
            {
                while (pos < end)
                {
                    const int pq = u8();
                    const int t = pq & 15;
                    if (t > 3) bad("bad quantisation table index");
                    for (int i = 0; i < 64; ++i)
                    {
                        out.quant[t][zigzag[i]] = static_cast<uint16_t>((pq >> 4) ? u16() : u8());
                    }
                }
            }

            void dht(size_t end)
            {
                while (pos < end)
                {
                    const int tc = u8();
                    unsigned char counts[16];
                    int n = 0;
                    for (int i = 0; i < 16; ++i)
                    {
                        counts[i] = static_cast<unsigned char>(u8());
                        n += counts[i];
                    }
                    if (n > 256 || pos + static_cast<size_t>(n) > size || (tc & 15) > 3) bad("bad Huffman table");
                    ((tc >> 4) ? ac : dc)[tc & 15].build(counts, data + pos, n);
                    pos += static_cast<size_t>(n);
                }
            }

            void sof(int marker)
            {
                if (sawSof) bad("more than one frame");
                progressive = 0xC2 == marker;
                if (u8() != 8) bad("only 8-bit samples are supported");
                out.height = u16();
                out.width = u16();
                const int n = u8();
                if (out.width <= 0 || out.height <= 0) bad("empty image");
                if (1 == n) bad("grayscale images are not supported (the effects take RGBA)");
                if (n != 3) bad("only 3-component images are supported");
                out.components.resize(3);
                for (auto& c : out.components)
                {
                    c.id = u8();
                    const int hv = u8();
                    c.h = hv >> 4;
                    c.v = hv & 15;
                    c.tq = u8();
                    if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) bad("bad component");
                    out.maxH = std::max(out.maxH, c.h);
                    out.maxV = std::max(out.maxV, c.v);
                }
                const int mcusX = (out.width + 8 * out.maxH - 1) / (8 * out.maxH);
                const int mcusY = (out.height + 8 * out.maxV - 1) / (8 * out.maxV);
                for (auto& c : out.components)
                {
                    c.blocksW = mcusX * c.h;
                    c.blocksH = mcusY * c.v;
                    c.width = (out.width * c.h + out.maxH - 1) / out.maxH;
                    c.height = (out.height * c.v + out.maxV - 1) / out.maxV;
                    c.coef.assign(static_cast<size_t>(c.blocksW) * c.blocksH * 64, 0);
                }
                sawSof = true;
            }

            // ---- block decoders ----
            void baselineBlock(BitReader& br, ScanComponent& sc, int16_t* block)
            {
                const int t = br.decode(dc[sc.td]);
                if (t > 15) bad("bad DC difference");
                sc.pred += t ? extend(br.bits(t), t) : 0;
                block[0] = static_cast<int16_t>(sc.pred);
                const HuffmanTable& table = ac[sc.ta];
                for (int k = 1; k < 64;)
                {
                    const int rs = br.decode(table);
                    const int r = rs >> 4, s = rs & 15;
                    if (0 == s)
                    {
                        if (r != 15) break;
                        k += 16;
                        continue;
                    }
                    k += r;
                    if (k > 63) bad("bad coefficient run");
                    block[zigzag[k]] = static_cast<int16_t>(extend(br.bits(s), s));
                    ++k;
                }
            }

            void dcFirst(BitReader& br, ScanComponent& sc, int16_t* block, int al)
            {
                const int t = br.decode(dc[sc.td]);
                if (t > 15) bad("bad DC difference");
                sc.pred += t ? extend(br.bits(t), t) : 0;
                block[0] = static_cast<int16_t>(sc.pred * (1 << al));
            }

            void acFirst(BitReader& br, const HuffmanTable& table, int16_t* block, int ss, int se, int al, int& eobrun)
            {
                if (eobrun > 0)
                {
                    --eobrun;
                    return;
                }
                for (int k = ss; k <= se;)
                {
                    const int rs = br.decode(table);
                    const int r = rs >> 4, s = rs & 15;
                    if (s)
                    {
                        k += r;
                        if (k > 63) bad("bad coefficient run");
                        block[zigzag[k]] = static_cast<int16_t>(extend(br.bits(s), s) * (1 << al));
                        ++k;
                    }
                    else if (15 == r)
                    {
                        k += 16;
                    }
                    else
                    {
                        eobrun = (1 << r) - 1;
                        if (r) eobrun += br.bits(r);
                        break;
                    }
                }
            }

            // one correction bit for a coefficient that is already non-zero
            static void refine(BitReader& br, int16_t& c, int plus, int minus)
            {
                if (br.bit() && 0 == (c & plus))
                {
                    c = static_cast<int16_t>(c + (c >= 0 ? plus : minus));
                }
            }

            void acRefine(BitReader& br, const HuffmanTable& table, int16_t* block, int ss, int se, int al, int& eobrun)
            {
                const int plus = 1 << al, minus = -(1 << al);
                int k = ss;
                if (0 == eobrun)
                {
                    while (k <= se)
                    {
                        const int rs = br.decode(table);
                        int r = rs >> 4;
                        const int s = rs & 15;
                        int value = 0;
                        if (s)
                        {
                            value = br.bit() ? plus : minus; // s is 1: a newly non-zero coefficient
                        }
                        else if (r != 15)
                        {
                            eobrun = 1 << r;
                            if (r) eobrun += br.bits(r);
                            break;
                        }
                        // skip r zero-history coefficients, correcting the non-zero ones passed on the way
                        while (k <= se)
                        {
                            int16_t& c = block[zigzag[k]];
                            if (c != 0)
                            {
                                refine(br, c, plus, minus);
                            }
                            else if (--r < 0)
                            {
                                break;
                            }
                            ++k;
                        }
                        if (value && k <= se)
                        {
                            block[zigzag[k]] = static_cast<int16_t>(value);
                        }
                        ++k;
                    }
                }
                if (eobrun > 0)
                {
                    // the rest of the band: only correction bits
                    for (; k <= se; ++k)
                    {
                        int16_t& c = block[zigzag[k]];
                        if (c != 0) refine(br, c, plus, minus);
                    }
                    --eobrun;
                }
            }

            void sos()
            {
                if (!sawSof) bad("scan before frame header");
                const size_t len = static_cast<size_t>(u16());
                const size_t end = pos + len - 2;
                const int n = u8();
                if (n < 1 || n > 3) bad("bad scan");
                std::vector<ScanComponent> scan(static_cast<size_t>(n));
                for (auto& sc : scan)
                {
                    const int id = u8();
                    for (auto& c : out.components)
                    {
                        if (c.id == id) sc.c = &c;
                    }
                    if (!sc.c) bad("scan names an unknown component");
                    const int t = u8();
                    sc.td = t >> 4;
                    sc.ta = t & 15;
                    if (sc.td > 3 || sc.ta > 3) bad("bad table selector");
                }
                const int ss = u8(), se = u8(), a = u8();
                const int ah = a >> 4, al = a & 15;
                pos = end;
                if (progressive)
                {
                    if (ss > se || se > 63 || (0 == ss && se != 0) || (ss > 0 && n != 1) || al > 13) bad("bad progressive scan");
                }
                for (const auto& sc : scan)
                {
                    const bool needDc = !progressive || (0 == ss && 0 == ah);
                    const bool needAc = !progressive || ss > 0;
                    if (needDc && !dc[sc.td].defined) bad("missing DC Huffman table");
                    if (needAc && !ac[sc.ta].defined) bad("missing AC Huffman table");
                }

                BitReader br;
                br.p = data + pos;
                br.end = data + size;
                int eobrun = 0;
                int untilRestart = restartInterval;
                int nextRestart = 0;

                auto block = [&](ScanComponent& sc, int bx, int by)
                {
                    int16_t* b = &sc.c->coef[(static_cast<size_t>(by) * sc.c->blocksW + bx) * 64];
                    if (!progressive) baselineBlock(br, sc, b);
                    else if (0 == ss)
                    {
                        if (0 == ah) dcFirst(br, sc, b, al);
                        else if (br.bit()) b[0] = static_cast<int16_t>(b[0] | (1 << al));
                    }
                    else if (0 == ah) acFirst(br, ac[sc.ta], b, ss, se, al, eobrun);
                    else acRefine(br, ac[sc.ta], b, ss, se, al, eobrun);
                };
                auto restart = [&]
                {
                    // byte-align, expect RSTn, reset the predictors
                    br.reset();
                    const unsigned char* q = br.p;
                    while (q + 1 < br.end && !(0xFF == q[0] && q[1] != 0 && q[1] != 0xFF)) ++q;
                    if (q + 1 < br.end && q[1] == 0xD0 + nextRestart)
                    {
                        br.p = q + 2;
                    }
                    else
                    {
                        br.p = q; // missing restart marker: carry on at the next marker like libjpeg's resync
                    }
                    nextRestart = (nextRestart + 1) & 7;
                    for (auto& sc : scan) sc.pred = 0;
                    eobrun = 0;
                    untilRestart = restartInterval;
                };

                if (1 == n)
                {
                    // non-interleaved: the component's real blocks, row by row
                    ScanComponent& sc = scan[0];
                    const int bw = (sc.c->width + 7) / 8, bh = (sc.c->height + 7) / 8;
                    for (int by = 0; by < bh; ++by)
                    {
                        for (int bx = 0; bx < bw; ++bx)
                        {
                            if (restartInterval && 0 == untilRestart) restart();
                            block(sc, bx, by);
                            --untilRestart;
                        }
                    }
                }
                else
                {
                    const int mcusX = out.components[0].blocksW / out.components[0].h;
                    const int mcusY = out.components[0].blocksH / out.components[0].v;
                    for (int my = 0; my < mcusY; ++my)
                    {
                        for (int mx = 0; mx < mcusX; ++mx)
                        {
                            if (restartInterval && 0 == untilRestart) restart();
                            for (auto& sc : scan)
                            {
                                for (int v = 0; v < sc.c->v; ++v)
                                {
                                    for (int h = 0; h < sc.c->h; ++h)
                                    {
                                        block(sc, mx * sc.c->h + h, my * sc.c->v + v);
                                    }
                                }
                            }
                            --untilRestart;
                        }
                    }
                }
                // continue at the marker that ends the entropy-coded segment
                const unsigned char* q = br.p;
                while (q + 1 < br.end && !(0xFF == q[0] && q[1] != 0 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
                pos = static_cast<size_t>(q - data);
            }

            void run()
            {
                if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) bad("not a JPEG stream");
                pos = 2;
                bool sawScan = false;
                for (;;)
                {
                    if (pos + 1 >= size)
                    {
                        if (sawScan) break; // missing EOI: keep what was decoded, like libjpeg
                        bad("truncated");
                    }
                    if (data[pos] != 0xFF)
                    {
                        ++pos;
                        continue;
                    }
                    const int m = data[pos + 1];
                    if (0xFF == m)
                    {
                        ++pos;
                        continue;
                    }
                    pos += 2;
                    if (0xD9 == m) break;
                    if (0x01 == m || (m >= 0xD0 && m <= 0xD7)) continue;
                    if (0xDA == m)
                    {
                        sos();
                        sawScan = true;
                        continue;
                    }
                    const size_t len = static_cast<size_t>(u16());
                    if (len < 2 || pos + len - 2 > size) bad("bad segment length");
                    const size_t end = pos + len - 2;
                    switch (m)
                    {
                    case 0xC0: case 0xC1: case 0xC2: sof(m); break;
                    case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                        bad("lossless, hierarchical and arithmetic-coded JPEG are not supported");
                    case 0xC4: dht(end); break;
                    case 0xDB: dqt(end); break;
                    case 0xDD: restartInterval = u16(); break;
                    case 0xE0:
                        if (len >= 7 && 0 == memcmp(data + pos, "JFIF", 5)) sawJfif = true;
                        break;
                    case 0xEE:
                        if (len >= 14 && 0 == memcmp(data + pos, "Adobe", 5))
                        {
                            sawAdobe = true;
                            adobeTransform = data[pos + 11];
                        }
                        break;
                    default: break;
                    }
                    pos = end;
                }
                if (!sawSof || !sawScan) bad("no image data");
                // libjpeg's default_decompress_parms for three components
                if (sawJfif) out.ycc = true;
                else if (sawAdobe) out.ycc = adobeTransform != 0;
                else
                {
                    const auto& c = out.components;
                    out.ycc = !('R' == c[0].id && 'G' == c[1].id && 'B' == c[2].id);
                }
            }
        };

        // ---- reconstruction: libjpeg's jidctint.c (islow) ----
        constexpr int CONST_BITS = 13, PASS1_BITS = 2;
        constexpr int F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
                      F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;

        inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

        // the post-IDCT range limit: the 10-bit wrapped value, re-centred and clamped
        inline unsigned char limit(int v)
        {
            v = ((v + 512) & 1023) - 512 + 128;
            return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
        }

        inline void idct1d(const int in[8], int out[8])
        {
            // even part
            int z1 = (in[2] + in[6]) * F_0_541;
            const int tmp2 = z1 + in[6] * -F_1_847;
            const int tmp3 = z1 + in[2] * F_0_765;
            const int tmp0 = (in[0] + in[4]) * (1 << CONST_BITS);
            const int tmp1 = (in[0] - in[4]) * (1 << CONST_BITS);
            const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            // odd part
            int t0 = in[7], t1 = in[5], t2 = in[3], t3 = in[1];
            z1 = t0 + t3;
            int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
            const int z5 = (z3 + z4) * F_1_175;
            t0 *= F_0_298;
            t1 *= F_2_053;
            t2 *= F_3_072;
            t3 *= F_1_501;
            z1 *= -F_0_899;
            z2 *= -F_2_562;
            z3 = z3 * -F_1_961 + z5;
            z4 = z4 * -F_0_390 + z5;
            t0 += z1 + z3;
            t1 += z2 + z4;
            t2 += z2 + z3;
            t3 += z1 + z4;
            out[0] = tmp10 + t3;
            out[7] = tmp10 - t3;
            out[1] = tmp11 + t2;
            out[6] = tmp11 - t2;
            out[2] = tmp12 + t1;
            out[5] = tmp12 - t1;
            out[3] = tmp13 + t0;
            out[4] = tmp13 - t0;
        }

        void idctBlock(const int16_t* coef, const uint16_t* quant, unsigned char* dst, size_t stride)
        {
            int ws[64];
            for (int x = 0; x < 8; ++x)
            {
                int in[8], o[8];
                for (int y = 0; y < 8; ++y) in[y] = coef[y * 8 + x] * quant[y * 8 + x];
                idct1d(in, o);
                for (int y = 0; y < 8; ++y) ws[y * 8 + x] = descale(o[y], CONST_BITS - PASS1_BITS);
            }
            for (int y = 0; y < 8; ++y)
            {
                int o[8];
                idct1d(&ws[y * 8], o);
                for (int x = 0; x < 8; ++x) dst[y * stride + x] = limit(descale(o[x], CONST_BITS + PASS1_BITS + 3));
            }
        }

        // ---- chroma upsampling: libjpeg-turbo's jdsample.c defaults (do_fancy_upsampling) ----
        struct Plane
        {
            int w = 0, h = 0;   // real samples
            size_t stride = 0;
            std::vector<unsigned char> px;
            const unsigned char* row(int y) const { return &px[static_cast<size_t>(std::min(std::max(y, 0), h - 1)) * stride]; }
        };

        void upsampleRowH2(const unsigned char* in, int w, unsigned char* out)
        {
            // h2v1 "triangle" filter: 3/4 nearer + 1/4 further, alternating rounding
            out[0] = in[0];
            out[1] = static_cast<unsigned char>((in[0] * 3 + in[1] + 2) >> 2);
            for (int i = 1; i < w - 1; ++i)
            {
                const int v = in[i] * 3;
                out[2 * i] = static_cast<unsigned char>((v + in[i - 1] + 1) >> 2);
                out[2 * i + 1] = static_cast<unsigned char>((v + in[i + 1] + 2) >> 2);
            }
            out[2 * w - 2] = static_cast<unsigned char>((in[w - 1] * 3 + in[w - 2] + 1) >> 2);
            out[2 * w - 1] = in[w - 1];
        }

        void upsampleRowH2V2(const unsigned char* near, const unsigned char* far, int w, unsigned char* out)
        {
            int cur = near[0] * 3 + far[0];
            int next = near[1] * 3 + far[1];
            out[0] = static_cast<unsigned char>((cur * 4 + 8) >> 4);
            out[1] = static_cast<unsigned char>((cur * 3 + next + 7) >> 4);
            int last = cur;
            cur = next;
            for (int i = 1; i < w - 1; ++i)
            {
                next = near[i + 1] * 3 + far[i + 1];
                out[2 * i] = static_cast<unsigned char>((cur * 3 + last + 8) >> 4);
                out[2 * i + 1] = static_cast<unsigned char>((cur * 3 + next + 7) >> 4);
                last = cur;
                cur = next;
            }
            out[2 * w - 2] = static_cast<unsigned char>((cur * 3 + last + 8) >> 4);
            out[2 * w - 1] = static_cast<unsigned char>((cur * 4 + 7) >> 4);
        }

        // One full-resolution row `y` of a component (at least `width` samples) into `out`.
        void upsampledRow(const Plane& p, int hs, int vs, int y, int width, unsigned char* out, std::vector<unsigned char>& tmp)
        {
            if (1 == hs && 1 == vs)
            {
                memcpy(out, p.row(y), static_cast<size_t>(width));
                return;
            }
            const bool fancyH = p.w > 2; // jinit_upsampler: the fancy variants need more than two columns
            if (2 == hs && 1 == vs)
            {
                if (fancyH)
                {
                    upsampleRowH2(p.row(y), p.w, out);
                    return;
                }
            }
            else if (2 == hs && 2 == vs)
            {
                if (fancyH)
                {
                    const int r = y >> 1;
                    upsampleRowH2V2(p.row(r), p.row((y & 1) ? r + 1 : r - 1), p.w, out);
                    return;
                }
            }
            else if (1 == hs && 2 == vs)
            {
                const int r = y >> 1;
                const unsigned char* near = p.row(r);
                const unsigned char* far = p.row((y & 1) ? r + 1 : r - 1);
                const int bias = (y & 1) ? 2 : 1;
                for (int x = 0; x < p.w; ++x) out[x] = static_cast<unsigned char>((near[x] * 3 + far[x] + bias) >> 2);
                return;
            }
            // any other integer ratio (and the narrow cases above): sample replication
            (void)tmp;
            const unsigned char* in = p.row(y / vs);
            for (int x = 0; x < width; ++x) out[x] = in[std::min(x / hs, p.w - 1)];
        }
    }

    void parseJpeg(const unsigned char* data, size_t size, JpegCoefficients& out)
    {
        out = JpegCoefficients();
        Decoder decoder(data, size, out);
        decoder.run();
        for (const auto& c : out.components)
        {
            if (out.maxH % c.h || out.maxV % c.v) bad("fractional sampling ratios are not supported");
        }
    }

    void reconstructJpeg(const JpegCoefficients& jpeg, unsigned char* rgba)
    {
        Plane planes[3];
        for (int i = 0; i < 3; ++i)
        {
            const JpegComponent& c = jpeg.components[static_cast<size_t>(i)];
            Plane& p = planes[i];
            p.w = c.width;
            p.h = c.height;
            p.stride = static_cast<size_t>(c.blocksW) * 8;
            p.px.resize(p.stride * static_cast<size_t>(c.blocksH) * 8);
            for (int by = 0; by < c.blocksH; ++by)
            {
                for (int bx = 0; bx < c.blocksW; ++bx)
                {
                    idctBlock(&c.coef[(static_cast<size_t>(by) * c.blocksW + bx) * 64], jpeg.quant[c.tq],
                        &p.px[static_cast<size_t>(by) * 8 * p.stride + static_cast<size_t>(bx) * 8], p.stride);
                }
            }
        }
        // libjpeg's jdcolor.c tables, SCALEBITS = 16
        int crR[256], cbB[256], crG[256], cbG[256];
        for (int i = 0; i < 256; ++i)
        {
            const int x = i - 128;
            crR[i] = (91881 * x + 32768) >> 16;
            cbB[i] = (116130 * x + 32768) >> 16;
            crG[i] = -46802 * x;
            cbG[i] = -22554 * x + 32768;
        }
        const int w = jpeg.width, h = jpeg.height;
        const size_t rowSize = static_cast<size_t>(w) + 16 + 2 * 8 * static_cast<size_t>(jpeg.maxH);
        std::vector<unsigned char> rows[3], tmp;
        for (int i = 0; i < 3; ++i)
        {
            rows[i].resize(std::max(rowSize, planes[i].stride * static_cast<size_t>(jpeg.maxH)));
        }
        auto clamp = [](int v) { return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v)); };
        for (int y = 0; y < h; ++y)
        {
            for (int i = 0; i < 3; ++i)
            {
                const JpegComponent& c = jpeg.components[static_cast<size_t>(i)];
                upsampledRow(planes[i], jpeg.maxH / c.h, jpeg.maxV / c.v, y, w, rows[i].data(), tmp);
            }
            unsigned char* o = rgba + static_cast<size_t>(y) * w * 4;
            for (int x = 0; x < w; ++x, o += 4)
            {
                const int a = rows[0][static_cast<size_t>(x)], b = rows[1][static_cast<size_t>(x)], c = rows[2][static_cast<size_t>(x)];
                if (jpeg.ycc)
                {
                    o[0] = clamp(a + crR[c]);
                    o[1] = clamp(a + ((cbG[b] + crG[c]) >> 16));
                    o[2] = clamp(a + cbB[b]);
                }
                else
                {
                    o[0] = static_cast<unsigned char>(a);
                    o[1] = static_cast<unsigned char>(b);
                    o[2] = static_cast<unsigned char>(c);
                }
                o[3] = 255;
            }
        }
    }

    bool readJpegMemory(const unsigned char* data, size_t size, MediaFrame& out)
    {
        JpegCoefficients jpeg;
        parseJpeg(data, size, jpeg);
        auto pixels = std::make_shared<std::vector<unsigned char> >(static_cast<size_t>(jpeg.width) * jpeg.height * 4);
        reconstructJpeg(jpeg, pixels->data());
        out = MediaFrame();
        out.width = jpeg.width;
        out.height = jpeg.height;
        out.nchannels = 4;
        out.format = TC_U8;
        out.owned = pixels;
        out.host = pixels->data();
        return true;
    }

    bool readJpeg(const std::string& path, MediaFrame& out)
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) return false;
        const std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        return readJpegMemory(bytes.data(), bytes.size(), out);
    }
}
