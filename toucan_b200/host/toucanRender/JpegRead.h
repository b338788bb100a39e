// JPEG reader for the read nodes (ingest).  The reference reads stills and
// image sequences through OpenImageIO (lib/toucanRender/Read.cpp:34-101,164-216), whose JPEG
// plug-in is libjpeg(-turbo) with its defaults: integer "islow" IDCT, "fancy" (triangle) chroma
// upsampling, integer YCbCr -> RGB tables.  This reader reproduces those defaults bit for bit
// (pinned against libjpeg-turbo through Pillow in tests/test_ingest.py) so a timeline that
// names a .jpg renders the same pixels as it does with the reference.
//
// Two stages, split where the data stops being serial:
//   parseJpeg        markers + Huffman entropy decoding (baseline, extended and progressive,
//                    restart intervals, any 1x/2x sampling) -> quantised DCT coefficients;
//   reconstructJpeg  dequantisation, IDCT, chroma upsampling, colour conversion -> RGBA u8
//                    (block-parallel; the stage a device decoder would take over).
#pragma once

#include <toucanRender/Read.h>

#include <cstdint>
#include <string>
#include <vector>

namespace toucan
{
    struct JpegComponent
    {
        int id = 0;
        int h = 1, v = 1;         //!< sampling factors
        int tq = 0;               //!< quantisation table
        int blocksW = 0;          //!< blocks per row of the coefficient array (padded to whole MCUs)
        int blocksH = 0;
        int width = 0;            //!< real samples per row: ceil(image width * h / maxH)
        int height = 0;
        std::vector<int16_t> coef; //!< blocksW * blocksH * 64, natural (row-major) order inside a block
    };

    struct JpegCoefficients
    {
        int width = 0, height = 0;
        int maxH = 1, maxV = 1;
        bool ycc = true;          //!< components are YCbCr (else RGB)
        std::vector<JpegComponent> components;
        uint16_t quant[4][64] = {}; //!< natural order
    };

    //! Entropy-decode a JPEG stream.  Throws std::runtime_error on malformed or unsupported data
    //! (arithmetic coding, lossless, 12-bit, CMYK, grayscale: the plug-ins take RGBA).
    void parseJpeg(const unsigned char* data, size_t size, JpegCoefficients& out);

    //! Coefficients -> RGBA u8 (A = 255: Read.cpp:86-94 pads RGB with A = 1).
    void reconstructJpeg(const JpegCoefficients&, unsigned char* rgba);

    bool readJpeg(const std::string& path, MediaFrame& out);
    bool readJpegMemory(const unsigned char* data, size_t size, MediaFrame& out);
}
