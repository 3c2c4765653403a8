#include <cstdio>
#include <thread>
#include <vector>
#include "mosaic_jpeg.hpp"
#include "mosaic_png.hpp"
#include "mosaic_webp.hpp"
using namespace mosaic_b200::codec;
int main(int argc, char** argv) {
  std::vector<std::vector<uint8_t>> files;
  for (int a = 1; a < argc; ++a) files.push_back(detail::read_file(argv[a]));
  std::vector<std::thread> pool;
  std::vector<long> sums(8, 0);
  for (int t = 0; t < 8; ++t)
    pool.emplace_back([&, t] {
      for (int round = 0; round < 6; ++round)
        for (size_t i = 0; i < files.size(); ++i) {
          const auto& d = files[(i + t) % files.size()];
          try {
            long s = 0;
            if (d[0] == 0x89) { auto im = decode_png_memory(d.data(), d.size()); for (auto v : im.rgb) s += v; }
            else if (d[0] == 0xFF) { auto im = decode_jpeg_memory(d.data(), d.size()); for (auto v : im.rgb) s += v; }
            else { auto im = decode_webp_memory(d.data(), d.size()); for (auto v : im.rgb) s += v; }
            sums[t] += s;
          } catch (const std::exception&) {}
        }
    });
  for (auto& th : pool) th.join();
  for (int t = 1; t < 8; ++t) if (sums[t] != sums[0]) { std::printf("MISMATCH\n"); return 1; }
  std::printf("tsan_decode: 8 threads agree (%ld)\n", sums[0]);
  return 0;
}
