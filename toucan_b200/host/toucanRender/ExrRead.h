// OpenEXR reader for the read nodes (ingest): the half / float stills and sequences a
// VFX timeline is made of.  The reference reads them through OpenImageIO -> OpenEXR
// (lib/toucanRender/Read.cpp:34-101,164-216); this reader covers single-part scanline files and the
// full-resolution level of single-part tiled files -- R/G/B[/A] channels of type HALF or FLOAT, compression NONE,
// RLE, ZIPS, ZIP, PIZ, PXR24, B44 or B44A -- and yields what OIIO's read_image yields for them: the data window, RGBA order, the
// widest channel type as the frame's type, A = 1 when the file has no alpha (Read.cpp:86-94).  Every sample is the
// file's sample (pinned against OpenEXR through OpenCV in tests/test_ingest.py; PXR24 keeps 24 bits of a float
// sample and B44 is lossy at the writer, by design; decoding is exact).  Deep, multi-part and DWA files throw (the read node then yields no image, like a
// file OIIO cannot open).
#pragma once

#include <toucanRender/Read.h>

#include <string>

namespace toucan
{
    bool readExr(const std::string& path, MediaFrame& out);
    bool readExrMemory(const unsigned char* data, size_t size, MediaFrame& out);

    //! Write an RGBA half (TC_F16) or float (TC_F32) frame as a scanline OpenEXR file: ZIP-compressed in blocks of 16
    //! lines (OpenImageIO's default for .exr output), or uncompressed.
    bool writeExr(const std::string& path, const void* pixels, int width, int height, int type, bool zip = true);
}
