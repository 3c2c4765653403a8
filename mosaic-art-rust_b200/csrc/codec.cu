// codec.cu — host-only entry points of the C ABI: PNG files <-> interleaved RGB8 (include/mosaic_png.hpp over zlib).
// what `image::open(path)?.to_rgb8()` (main.rs:263-301, 484, 548, 632) and
// `output_img.save(path)` (main.rs:668-672) do for the one lossless format the reference accepts.  No device work.
#include "../../include/mosaic_png.hpp"
#include "common.cuh"

using namespace mosaic;

extern "C" {

int mosaic_png_decode_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height) {
  if (!file || !rgb_out || !width || !height) return fail(nullptr, MOSAIC_ERR_INVALID, "png_decode: bad arguments");
  *rgb_out = nullptr;
  try {
    const mosaic_b200::codec::Image img = mosaic_b200::codec::decode_png(file);
    uint8_t* buf = static_cast<uint8_t*>(std::malloc(img.rgb.size() ? img.rgb.size() : 1));
    if (!buf) return fail(nullptr, MOSAIC_ERR_INVALID, "png_decode: out of memory");
    std::memcpy(buf, img.rgb.data(), img.rgb.size());
    *rgb_out = buf;
    *width = img.width;
    *height = img.height;
    return MOSAIC_OK;
  } catch (const std::exception& e) {
    return fail(nullptr, MOSAIC_ERR_INVALID, std::string(e.what()) + " (" + file + ")");
  }
}

int mosaic_png_encode_rgb8(const char* file, const uint8_t* rgb, int width, int height) {
  if (!file || !rgb) return fail(nullptr, MOSAIC_ERR_INVALID, "png_encode: bad arguments");
  try {
    mosaic_b200::codec::encode_png(file, rgb, width, height);
    return MOSAIC_OK;
  } catch (const std::exception& e) {
    return fail(nullptr, MOSAIC_ERR_INVALID, std::string(e.what()) + " (" + file + ")");
  }
}

void mosaic_host_free(void* p) { std::free(p); }

}  // extern "C"
