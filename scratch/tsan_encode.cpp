#include <cstdio>
#include <thread>
#include <vector>
#include "mosaic_jpeg.hpp"
#include "mosaic_png.hpp"
#include "mosaic_webp.hpp"
using namespace mosaic_b200::codec;
int main() {
  const int w = 97, h = 61;
  std::vector<uint8_t> img((size_t)w * h * 3);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 13);
  std::vector<std::thread> pool;
  std::vector<size_t> sizes(8, 0);
  for (int t = 0; t < 8; ++t)
    pool.emplace_back([&, t] {
      for (int round = 0; round < 10; ++round) {
        sizes[t] += encode_png_memory(img.data(), w, h).size();
        sizes[t] += encode_jpeg_memory(img.data(), w, h, 60 + round).size();
        sizes[t] += webp_detail::encode_vp8l(img.data(), w, h).size();
      }
    });
  for (auto& th : pool) th.join();
  for (int t = 1; t < 8; ++t) if (sizes[t] != sizes[0]) { std::printf("MISMATCH\n"); return 1; }
  std::printf("tsan_encode: 8 threads agree (%zu bytes)\n", sizes[0]);
  return 0;
}
