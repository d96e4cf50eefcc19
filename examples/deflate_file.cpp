// C++ caller of include/deflate_b200.hpp: compresses a file with the crate-style API, one-shot and through a
// ZlibEncoder fed in pieces, and checks that both give the same bytes.
//   usage: deflate_file_cpp <input> <output.zlib>
#include <cstdio>
#include <fstream>
#include <iterator>

#include "deflate_b200.hpp"

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <input> <output.zlib>\n", argv[0]); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    try {
        namespace d = deflate_b200;
        const std::vector<uint8_t> one = d::deflate_bytes_zlib_conf(data.data(), data.size(), d::Compression::Default);
        std::vector<uint8_t> streamed;
        {
            d::ZlibEncoder enc([&](const uint8_t* p, size_t n) { streamed.insert(streamed.end(), p, p + n); return n; },
                               d::Compression::Default);
            for (size_t i = 0; i < data.size(); i += 50000) enc.write_all(data.data() + i, std::min<size_t>(50000, data.size() - i));
            enc.finish();
        }
        if (one != streamed) { std::fprintf(stderr, "one-shot and streamed output differ\n"); return 4; }
        std::ofstream o(argv[2], std::ios::binary);
        o.write(reinterpret_cast<const char*>(one.data()), (std::streamsize)one.size());
    } catch (const deflate_b200::Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    return 0;
}
