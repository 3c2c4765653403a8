// fuzz_codecs.cpp — mutation fuzzer for the three host-side decoders (include/mosaic_png.hpp, mosaic_jpeg.hpp, mosaic_webp.hpp),
// built with -fsanitize=address,undefined by tests/test_codec_fuzz.py.  Material directories are untrusted input: a corrupt
// file must end in an exception (the C ABI's error code), never in a crash, an out-of-bounds access or an absurd allocation.
//   usage: fuzz_codecs <trials per file> <file>...      (format by the file's first bytes)
// Mutations: byte flips anywhere, flips concentrated on the headers (first 160 bytes), truncation; PNG chunk CRCs are repaired
// afterwards in four of five trials so that the mutation reaches the decoder behind the CRC check.
#include <cstdio>
#include <random>

#include "mosaic_jpeg.hpp"
#include "mosaic_png.hpp"
#include "mosaic_webp.hpp"

using namespace mosaic_b200::codec;

static void repair_png_crcs(std::vector<uint8_t>& d) {
  size_t pos = 8;
  while (pos + 12 <= d.size()) {
    const uint32_t len = ((uint32_t)d[pos] << 24) | ((uint32_t)d[pos + 1] << 16) | ((uint32_t)d[pos + 2] << 8) | d[pos + 3];
    if (len > d.size() - pos - 12) break;
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), d.data() + pos + 4, len + 4);
    d[pos + 8 + len] = (uint8_t)(crc >> 24), d[pos + 9 + len] = (uint8_t)(crc >> 16), d[pos + 10 + len] = (uint8_t)(crc >> 8), d[pos + 11 + len] = (uint8_t)crc;
    pos += 12 + len;
  }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const int trials = std::atoi(argv[1]);
  std::mt19937 rng(20261018);
  long decoded = 0, refused = 0;
  for (int a = 2; a < argc; ++a) {
    const std::vector<uint8_t> data = detail::read_file(argv[a]);
    if (data.size() < 32) return 3;
    const bool png = data[0] == 0x89, jpeg = data[0] == 0xFF;
    for (int t = 0; t < trials; ++t) {
      std::vector<uint8_t> bad = data;
      const int flips = 1 + (int)(rng() % 4);
      const size_t span = (t & 1) ? std::min<size_t>(bad.size() - 2, 160) : bad.size() - 2;
      for (int i = 0; i < flips; ++i) {
        const size_t pos = 2 + rng() % span;
        bad[pos] = (rng() % 3 == 0) ? (uint8_t)rng() : (uint8_t)(bad[pos] ^ (1u << (rng() % 8)));
      }
      if (t % 7 == 0) bad.resize(16 + rng() % (bad.size() - 16));
      if (png && t % 5 != 0) repair_png_crcs(bad);
      try {
        if (png)
          (void)decode_png_memory(bad.data(), bad.size());
        else if (jpeg)
          (void)decode_jpeg_memory(bad.data(), bad.size());
        else
          (void)decode_webp_memory(bad.data(), bad.size());
        ++decoded;
      } catch (const std::exception&) {
        ++refused;
      }
    }
  }
  std::printf("fuzz_codecs: %ld decoded, %ld refused\n", decoded, refused);
  return 0;
}
