// Small PNG reader/writer over zlib for the command line renderer (8/16-bit RGB/RGBA,
// non-interlaced).  Replicates the two read-side conventions of the reference's OIIO
// reader that affect pixels: RGB gets A = 1 appended (Read.cpp:86-94) and PNG's
// unassociated alpha is associated on read.  Decode is outside the hot path.
#pragma once

#include <toucanRender/Read.h>

#include <string>

namespace toucan
{
    bool readPng(const std::string& path, MediaFrame& out);
    bool readPngMemory(const unsigned char* data, size_t size, MediaFrame& out);

    //! Write 8-bit RGBA (alpha dropped when `rgb` is true).
    bool writePng(const std::string& path, const unsigned char* rgba, int width, int height, bool rgb = false);

    //! Write an RGBA frame of 8- or 16-bit samples (host byte order in memory).  With `unassociate`, premultiplied
    //! colour is divided by alpha on the way out, as PNG stores unassociated alpha.
    bool writePngFrame(const std::string& path, const void* pixels, int width, int height, int depth, bool rgb, bool unassociate);
}
