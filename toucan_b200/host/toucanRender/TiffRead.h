// TIFF reader for the read nodes (ingest; the reference lists .tif / .tiff among the extensions it
// opens through OpenImageIO, lib/toucanRender/Read.cpp:103-106).  Classic (not Big) TIFF, stripped, chunky RGB / RGBA
// of 8- or 16-bit unsigned or 32-bit float samples, either byte order; uncompressed, LZW, deflate or PackBits, with
// or without the horizontal-differencing predictor.  RGB gets A = 1 (Read.cpp:86-94); unassociated alpha
// (ExtraSamples = 2) is associated on read like OIIO's reader does.  Lossless: pinned against libtiff through Pillow and
// OpenCV in tests/test_ingest.py.  Anything else (tiles, planar, palettes, CMYK, JPEG-in-TIFF) throws.
#pragma once

#include <toucanRender/Read.h>

#include <string>

namespace toucan
{
    bool readTiff(const std::string& path, MediaFrame& out);
    bool readTiffMemory(const unsigned char* data, size_t size, MediaFrame& out);
}
