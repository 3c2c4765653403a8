// decode_files.cpp — decodes every file named on the command line ONCE with the host-side decoders (format by the first
// byte, like fuzz_codecs.cpp) and counts results; built with -fsanitize=address,undefined by tests/test_codec_hostile.py,
// which feeds it the structure-aware hostile files of tests/hostile_files.py.  A hostile file may decode or be refused —
// what it must not do is crash, read or write out of bounds, or allocate by a number it claims.
#include <cstdio>

#include "mosaic_jpeg.hpp"
#include "mosaic_png.hpp"
#include "mosaic_webp.hpp"

using namespace mosaic_b200::codec;

int main(int argc, char** argv) {
  long decoded = 0, refused = 0;
  for (int a = 1; a < argc; ++a) {
    const std::vector<uint8_t> d = detail::read_file(argv[a]);
    try {
      if (!d.empty() && d[0] == 0x89)
        (void)decode_png_memory(d.data(), d.size());
      else if (!d.empty() && d[0] == 0xFF)
        (void)decode_jpeg_memory(d.data(), d.size());
      else
        (void)decode_webp_memory(d.data(), d.size());
      ++decoded;
    } catch (const std::exception&) {
      ++refused;
    }
  }
  std::printf("decode_files: %ld decoded, %ld refused\n", decoded, refused);
  return 0;
}
