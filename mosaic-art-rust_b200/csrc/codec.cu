// codec.cu — host-only entry points of the C ABI: PNG files <-> interleaved RGB8 (include/mosaic_png.hpp over zlib), JPEG files <->
// RGB8 (include/mosaic_jpeg.hpp), WebP files <-> RGB8 (include/mosaic_webp.hpp: VP8 and VP8L in, VP8L out).
// what `image::open(path)?.to_rgb8()` (main.rs:263-301, 484, 548, 632) and
// `output_img.save(path)` (main.rs:668-672) do for the one lossless format the reference accepts.  No device work.
#include "../../include/mosaic_jpeg.hpp"
#include "../../include/mosaic_png.hpp"
#include "../../include/mosaic_webp.hpp"
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

int mosaic_jpeg_decode_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height) {
  if (!file || !rgb_out || !width || !height) return fail(nullptr, MOSAIC_ERR_INVALID, "jpeg_decode: bad arguments");
  *rgb_out = nullptr;
  try {
    const mosaic_b200::codec::JpegImage img = mosaic_b200::codec::decode_jpeg(file);
    uint8_t* buf = static_cast<uint8_t*>(std::malloc(img.rgb.size() ? img.rgb.size() : 1));
    if (!buf) return fail(nullptr, MOSAIC_ERR_INVALID, "jpeg_decode: out of memory");
    std::memcpy(buf, img.rgb.data(), img.rgb.size());
    *rgb_out = buf;
    *width = img.width;
    *height = img.height;
    return MOSAIC_OK;
  } catch (const std::exception& e) {
    return fail(nullptr, MOSAIC_ERR_INVALID, std::string(e.what()) + " (" + file + ")");
  }
}

int mosaic_jpeg_encode_rgb8(const char* file, const uint8_t* rgb, int width, int height, int quality) {
  if (!file || !rgb) return fail(nullptr, MOSAIC_ERR_INVALID, "jpeg_encode: bad arguments");
  try {
    mosaic_b200::codec::encode_jpeg(file, rgb, width, height, quality <= 0 ? 75 : quality);
    return MOSAIC_OK;
  } catch (const std::exception& e) {
    return fail(nullptr, MOSAIC_ERR_INVALID, std::string(e.what()) + " (" + file + ")");
  }
}

int mosaic_webp_decode_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height) {
  if (!file || !rgb_out || !width || !height) return fail(nullptr, MOSAIC_ERR_INVALID, "webp_decode: bad arguments");
  *rgb_out = nullptr;
  try {
    const mosaic_b200::codec::Image img = mosaic_b200::codec::decode_webp(file);
    uint8_t* buf = static_cast<uint8_t*>(std::malloc(img.rgb.size() ? img.rgb.size() : 1));
    if (!buf) return fail(nullptr, MOSAIC_ERR_INVALID, "webp_decode: out of memory");
    std::memcpy(buf, img.rgb.data(), img.rgb.size());
    *rgb_out = buf;
    *width = img.width;
    *height = img.height;
    return MOSAIC_OK;
  } catch (const std::exception& e) {
    return fail(nullptr, MOSAIC_ERR_INVALID, std::string(e.what()) + " (" + file + ")");
  }
}

int mosaic_webp_encode_rgb8(const char* file, const uint8_t* rgb, int width, int height) {
  if (!file || !rgb) return fail(nullptr, MOSAIC_ERR_INVALID, "webp_encode: bad arguments");
  try {
    mosaic_b200::codec::encode_webp(file, rgb, width, height);
    return MOSAIC_OK;
  } catch (const std::exception& e) {
    return fail(nullptr, MOSAIC_ERR_INVALID, std::string(e.what()) + " (" + file + ")");
  }
}

// image::open(path)?.to_rgb8() for the formats this library decodes, told apart by the file's first bytes
int mosaic_image_decode_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height) {
  if (!file) return fail(nullptr, MOSAIC_ERR_INVALID, "image_decode: bad arguments");
  unsigned char head[12] = {};
  FILE* f = std::fopen(file, "rb");
  if (!f) return fail(nullptr, MOSAIC_ERR_INVALID, std::string("image_decode: cannot open ") + file);
  const size_t got = std::fread(head, 1, sizeof head, f);
  std::fclose(f);
  if (got >= 8 && head[0] == 0x89 && head[1] == 'P' && head[2] == 'N' && head[3] == 'G') return mosaic_png_decode_rgb8(file, rgb_out, width, height);
  if (got >= 3 && head[0] == 0xFF && head[1] == 0xD8 && head[2] == 0xFF) return mosaic_jpeg_decode_rgb8(file, rgb_out, width, height);
  if (got >= 12 && !std::memcmp(head, "RIFF", 4) && !std::memcmp(head + 8, "WEBP", 4)) return mosaic_webp_decode_rgb8(file, rgb_out, width, height);
  return fail(nullptr, MOSAIC_ERR_INVALID, std::string("image_decode: not a PNG, JPEG or WebP file (the formats src/main.rs:181 accepts): ") + file);
}

// image::open(path)?.to_rgb8() with image::open's own rule: "the image's format is determined from the path's file
// extension" (ImageFormat::from_path, lower-cased) — the content is not sniffed, so a JPEG file called x.png is an error
// of the PNG decoder and the tile is dropped with a message, as in the reference (main.rs:198-206)
int mosaic_image_open_rgb8(const char* file, uint8_t** rgb_out, int* width, int* height) {
  if (!file || !rgb_out || !width || !height) return fail(nullptr, MOSAIC_ERR_INVALID, "image_open: bad arguments");
  *rgb_out = nullptr;
  const char* slash = std::strrchr(file, '/');
  const char* dot = std::strrchr(slash ? slash + 1 : file, '.');
  std::string ext = dot && dot != (slash ? slash + 1 : file) ? dot + 1 : "";  // ".png" alone is a file name, not an extension
  for (char& ch : ext)
    if (ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
  if (ext == "png") return mosaic_png_decode_rgb8(file, rgb_out, width, height);
  if (ext == "jpg" || ext == "jpeg") return mosaic_jpeg_decode_rgb8(file, rgb_out, width, height);
  if (ext == "webp") return mosaic_webp_decode_rgb8(file, rgb_out, width, height);
  return fail(nullptr, MOSAIC_ERR_INVALID, std::string("image_open: the extension of ") + file +
                                               " is not one of png / jpg / jpeg / webp (the formats src/main.rs:181 accepts)");
}

void mosaic_host_free(void* p) { std::free(p); }

}  // extern "C"
