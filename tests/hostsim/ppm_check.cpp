// Test-only driver of the C++ host's P3 writer (host/rayz_host.hpp: Image::writePPM, restating image.zig:29-41).
//   ppm_check <w> <h> <seed> <out.ppm>     fills rgb8 with a seeded pattern that covers all 256 values, writes the file,
//                                          prints "<bytes> <best format ms of 5>" on stdout
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../../host/rayz_host.hpp"

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const size_t w = std::strtoull(argv[1], nullptr, 10), h = std::strtoull(argv[2], nullptr, 10);
    uint64_t s = std::strtoull(argv[3], nullptr, 10);
    rayz::Image img = rayz::Image::initEmpty(h, w);
    img.rgb8.resize(w * h * 3);
    for (size_t i = 0; i < img.rgb8.size(); i++) {   // the same LCG the Python test runs
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        img.rgb8[i] = (uint8_t)(s >> 56);
    }
    FILE *f = std::fopen(argv[4], "wb");
    if (!f) return 1;
    const size_t n = img.writePPM(f);
    std::fclose(f);
    std::vector<char> buf;
    double best = 1e30;
    for (int i = 0; i < 5; i++) {
        const auto t0 = std::chrono::steady_clock::now();
        img.formatPPM(buf);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < best) best = ms;
    }
    std::printf("%zu %.3f\n", n, best);
    return 0;
}
