// Robustness harness for the built-in readers (tests/test_ingest.py::test_readers_survive_damaged_files): every file given
// on the command line decoded whole, truncated at sixteen lengths and with flipped bytes (argv[1] mutations per file),
// meant to be built with AddressSanitizer + UBSan.  A reader may refuse (return false / throw std::exception) but must not crash.
#include <toucanRender/ExrRead.h>
#include <toucanRender/JpegRead.h>
#include <toucanRender/PngRead.h>
#include <toucanRender/TiffRead.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include <vector>

using namespace toucan;

static int decode(const std::string& path, const std::vector<unsigned char>& b)
{
    MediaFrame f;
    try
    {
        const auto ext = path.substr(path.rfind('.'));
        if (ext == ".exr") return readExrMemory(b.data(), b.size(), f) ? 1 : 0;
        if (ext == ".tif") return readTiffMemory(b.data(), b.size(), f) ? 1 : 0;
        if (ext == ".png") return readPngMemory(b.data(), b.size(), f) ? 1 : 0;
        if (ext == ".jpg") return readJpegMemory(b.data(), b.size(), f) ? 1 : 0;
    }
    catch (const std::exception&)
    {
        return -1;
    }
    return 0;
}

int main(int argc, char** argv)
{
    long ok = 0, refused = 0, thrown = 0;
    const int mutations = argc > 1 ? std::atoi(argv[1]) : 0;
    for (int i = 2; i < argc; ++i)
    {
        std::ifstream file(argv[i], std::ios::binary);
        std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
        if (bytes.empty()) continue;
        auto count = [&](int r) { (r > 0 ? ok : (r < 0 ? thrown : refused))++; };
        count(decode(argv[i], bytes));
        for (int k = 1; k < 16; ++k)
        {
            std::vector<unsigned char> cut(bytes.begin(), bytes.begin() + bytes.size() * k / 16);
            count(decode(argv[i], cut));
        }
        std::mt19937 rng(99u + i);
        for (int m = 0; m < mutations; ++m)
        {
            std::vector<unsigned char> mut = bytes;
            const int flips = 1 + int(rng() % 4);
            for (int q = 0; q < flips; ++q)
            {
                // mostly in the first kilobyte (headers, tables), sometimes anywhere
                const size_t at = (rng() % 3) ? rng() % std::min<size_t>(mut.size(), 1024) : rng() % mut.size();
                mut[at] = static_cast<unsigned char>(rng());
            }
            count(decode(argv[i], mut));
        }
    }
    std::printf("decoded %ld, refused %ld, threw %ld\n", ok, refused, thrown);
    return 0;
}
