// The JSON / OTIO document parser and the .otioz central-directory reader with damaged inputs, meant to be built with
// AddressSanitizer + UBSan (tests/test_host_graph.py::test_timeline_documents_survive_damage).
// usage: otio_fuzz <mutations per file> <scratch path for mutated archives> files...
#include <opentimelineio/otio.h>
#include <toucanRender/Archive.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include <vector>

int main(int argc, char** argv)
{
    const int mutations = std::atoi(argv[1]);
    long ok = 0, failed = 0;
    const std::string scratch = argv[2];
    for (int i = 3; i < argc; ++i)
    {
        std::ifstream file(argv[i], std::ios::binary);
        std::string bytes((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
        const std::string path = argv[i];
        const bool zip = path.size() > 6 && path.substr(path.size() - 6) == ".otioz";
        std::mt19937 rng(7u + i);
        for (int m = -1; m < mutations; ++m)
        {
            std::string mut = bytes;
            if (m >= 0)
            {
                if (m % 5 == 0) mut.resize(rng() % mut.size());
                const int flips = 1 + int(rng() % 3);
                for (int q = 0; q < flips && !mut.empty(); ++q)
                {
                    size_t at = rng() % mut.size();
                    if (zip && (rng() & 1)) at = mut.size() - 1 - rng() % std::min<size_t>(mut.size(), 400); // the central directory
                    mut[at] = char(rng());
                }
                if (zip && (m % 7 == 3 || m % 7 == 5))
                {
                    // crafted central-directory records: a stored entry that claims more bytes than it holds, sizes / offsets
                    // at the 32-bit sentinel (ZIP64 lookups) and offsets near 2^32 (sums that would wrap)
                    mut = bytes;
                    for (size_t at = 0; at + 46 < mut.size(); ++at)
                    {
                        if (mut[at] == 'P' && mut[at + 1] == 'K' && mut[at + 2] == 1 && mut[at + 3] == 2 && (rng() & 1))
                        {
                            const uint32_t v = m % 7 == 3 ? (rng() | 0x40000000u) : (0xfffffff0u + rng() % 16);
                            const size_t field = m % 7 == 3 ? (rng() & 1 ? 24 : 20) : 42;
                            for (int b = 0; b < 4; ++b) mut[at + field + b] = char(v >> (8 * b));
                        }
                    }
                }
            }
            try
            {
                if (zip)
                {
                    const std::string tmp = scratch;
                    std::ofstream(tmp, std::ios::binary).write(mut.data(), std::streamsize(mut.size()));
                    toucan::ZipArchive archive(tmp);
                    (void)archive.find("content.otio");
                    // every entry the directory accepted can be read without leaving the mapping
                    for (const auto& e : archive.entries())
                    {
                        if (e.second.size > (size_t(1) << 28)) continue;
                        try { (void)archive.read(e.second); } catch (const std::exception&) {}
                        if (0 == e.second.method && (e.second.dataOffset > archive.size() || archive.size() - e.second.dataOffset < e.second.size)) std::abort();
                    }
                    ++ok;
                }
                else
                {
                    OTIO_NS::ErrorStatus status;
                    OTIO_NS::SerializableObject* obj = OTIO_NS::Timeline::from_json_string(mut, &status);
                    OTIO_NS::SerializableObject::Retainer<OTIO_NS::SerializableObject> keep(obj);
                    (obj ? ok : failed)++;
                }
            }
            catch (const std::exception&)
            {
                ++failed;
            }
        }
    }
    std::printf("parsed %ld, refused %ld\n", ok, failed);
    return 0;
}
