// mosaic_png.hpp — PNG in and out for the host side of the path (C++17 + zlib, header-only).
//
// the reference reads every material and the target with `image::open(path)` and hands the
// pixels on as `img.to_rgb8()` (src/main.rs:263-301, 484, 548, 632; src/lib.rs:94-96), and writes the mosaic with
// `output_img.save(output_path)` (src/main.rs:668-672).  PNG is the lossless member of the formats it accepts
// (main.rs:181: png / jpg / jpeg / webp) and the only one this image has a library for (zlib); JPEG and WebP are decoded by
// include/mosaic_jpeg.hpp and include/mosaic_webp.hpp.
//
//   decode_png(path)  = image::open(path)?.to_rgb8() for a PNG file [3P: image 0.25.6 / png 0.17.16, restated]:
//                       every colour type (grey, RGB, palette, grey+alpha, RGBA), bit depths 1/2/4/8/16, Adam7 interlace;
//                       grey is replicated, alpha and tRNS are dropped (to_rgb8 does not blend), 1/2/4-bit grey is scaled by
//                       255 / (2^d - 1) (png's EXPAND transformation), 16-bit samples become round(v * 255 / 65535)
//                       = (v + 128) / 257 (image's FromPrimitive<u16> for u8).  CRCs are checked.
//   encode_png(path)  = RgbImage::save(path) for a ".png" path: 8-bit RGB, no interlace.  The reference's encoder picks its
//                       own filters and deflate settings — the FILE BYTES are not part of the contract, the pixels are
//                       (a PNG is lossless): this writer chooses the filter of every row by the minimum sum of absolute
//                       differences and deflates at zlib's default level.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace mosaic_b200 {
namespace codec {

struct Image {  // interleaved RGB8, row-major (image::RgbImage)
  int width = 0, height = 0;
  std::vector<uint8_t> rgb;
};

struct PngError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

namespace detail {

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24), p[1] = (uint8_t)(v >> 16), p[2] = (uint8_t)(v >> 8), p[3] = (uint8_t)v; }

inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undo the filter of one scanline in place (`prev`: the reconstructed line above, or NULL for the first one of a pass)
inline void unfilter(int type, uint8_t* cur, const uint8_t* prev, size_t n, size_t bpp) {
  switch (type) {
    case 0: break;
    case 1:
      for (size_t i = bpp; i < n; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);
      break;
    case 2:
      if (prev)
        for (size_t i = 0; i < n; ++i) cur[i] = (uint8_t)(cur[i] + prev[i]);
      break;
    case 3:
      for (size_t i = 0; i < n; ++i) {
        const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0;
        cur[i] = (uint8_t)(cur[i] + ((a + b) >> 1));
      }
      break;
    case 4:
      for (size_t i = 0; i < n; ++i) {
        const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
        cur[i] = (uint8_t)(cur[i] + paeth(a, b, c));
      }
      break;
    default: throw PngError("png: unknown filter type");
  }
}

inline std::vector<uint8_t> read_file(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw PngError("png: cannot open " + path);
  std::vector<uint8_t> data;
  uint8_t buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + got);
  std::fclose(f);
  return data;
}

}  // namespace detail

inline Image decode_png_memory(const uint8_t* data, size_t size) {
  using namespace detail;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (size < 8 || std::memcmp(data, sig, 8) != 0) throw PngError("png: bad signature");
  uint32_t width = 0, height = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> palette, idat;
  bool have_ihdr = false, have_iend = false;
  size_t pos = 8;
  while (pos + 12 <= size && !have_iend) {
    const uint32_t len = be32(data + pos);
    if (len > size - pos - 12) throw PngError("png: truncated chunk");
    const uint8_t* type = data + pos + 4;
    const uint8_t* body = data + pos + 8;
    const uint32_t crc = be32(body + len);
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4) != crc) throw PngError("png: chunk CRC mismatch");
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len != 13) throw PngError("png: bad IHDR");
      width = be32(body), height = be32(body + 4);
      depth = body[8], ctype = body[9], interlace = body[12];
      if (body[10] != 0 || body[11] != 0 || interlace > 1) throw PngError("png: unsupported compression / filter / interlace method");
      have_ihdr = true;
    } else if (!std::memcmp(type, "PLTE", 4)) {
      palette.assign(body, body + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      have_iend = true;
    } else if (!(type[0] & 0x20)) {
      throw PngError("png: unknown critical chunk");
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || !have_iend) throw PngError("png: missing IHDR / IEND");
  if (width == 0 || height == 0 || width > (1u << 24) || height > (1u << 24)) throw PngError("png: bad dimensions");
  int channels;
  switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: throw PngError("png: bad colour type");
  }
  const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8) : (depth == 8 || depth == 16);
  if (!depth_ok) throw PngError("png: bad bit depth for the colour type");
  if (ctype == 3 && (palette.empty() || palette.size() % 3 != 0)) throw PngError("png: palette image without PLTE");
  const size_t bits_pp = (size_t)channels * depth;
  const size_t bpp = bits_pp >= 8 ? bits_pp / 8 : 1;  // filter distance in bytes
  auto line_bytes = [&](uint32_t w) { return ((size_t)w * bits_pp + 7) / 8; };

  // passes: the whole image, or Adam7's seven sub-images
  struct Pass { int x0, y0, dx, dy; };
  static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
  static const Pass whole = {0, 0, 1, 1};
  const int npass = interlace ? 7 : 1;
  size_t raw_size = 0;
  for (int ps = 0; ps < npass; ++ps) {
    const Pass& P = interlace ? adam7[ps] : whole;
    const uint32_t pw = (width + P.dx - 1 - P.x0) / P.dx, ph = (height + P.dy - 1 - P.y0) / P.dy;
    if (pw && ph) raw_size += (size_t)ph * (1 + line_bytes(pw));
  }
  // A header can claim 2^31 x 2^31 pixels in front of a few bytes of data: deflate expands by at most ~1032 : 1, so the IDAT
  // stream bounds the image it can hold — checked before anything of the claimed size is allocated
  if (raw_size > idat.size() * 1032 + 65536) throw PngError("png: the IDAT stream is too short for the dimensions in the header");
  if ((uint64_t)width * (uint64_t)height * 3u > ((uint64_t)1 << 32)) throw PngError("png: more than 4 GiB of pixels");
  std::vector<uint8_t> raw(raw_size);
  {
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) throw PngError("png: inflateInit failed");
    // (zlib counts in 32 bits: feed and drain in pieces)
    size_t in_pos = 0, out_pos = 0;
    int rc;
    do {
      if (zs.avail_in == 0 && in_pos < idat.size()) {
        const size_t piece = std::min<size_t>(idat.size() - in_pos, (size_t)1 << 30);
        zs.next_in = idat.data() + in_pos;
        zs.avail_in = (uInt)piece;
        in_pos += piece;
      }
      if (zs.avail_out == 0 && out_pos < raw.size()) {
        const size_t piece = std::min<size_t>(raw.size() - out_pos, (size_t)1 << 30);
        zs.next_out = raw.data() + out_pos;
        zs.avail_out = (uInt)piece;
        out_pos += piece;
      }
      rc = inflate(&zs, Z_NO_FLUSH);
    } while (rc == Z_OK && (zs.avail_in > 0 || in_pos < idat.size()) && (zs.avail_out > 0 || out_pos < raw.size()));
    const size_t produced = (size_t)zs.total_out;
    inflateEnd(&zs);
    if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || produced != raw.size()) throw PngError("png: bad image data");
  }

  Image img;
  img.width = (int)width, img.height = (int)height;
  img.rgb.assign((size_t)width * height * 3, 0);
  auto sample8 = [&](const uint8_t* line, size_t idx) -> uint8_t {  // sample `idx` of a line as an 8-bit value
    if (depth == 8) return line[idx];
    if (depth == 16) {
      const uint32_t v = ((uint32_t)line[2 * idx] << 8) | line[2 * idx + 1];
      return (uint8_t)((v + 128u) / 257u);
    }
    const size_t bit = idx * depth;
    const uint32_t v = (line[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
    return ctype == 3 ? (uint8_t)v : (uint8_t)(v * (255u / ((1u << depth) - 1u)));
  };
  size_t off = 0;
  for (int ps = 0; ps < npass; ++ps) {
    const Pass& P = interlace ? adam7[ps] : whole;
    const uint32_t pw = (width + P.dx - 1 - P.x0) / P.dx, ph = (height + P.dy - 1 - P.y0) / P.dy;
    if (!pw || !ph) continue;
    const size_t lb = line_bytes(pw);
    const uint8_t* prev = nullptr;
    for (uint32_t r = 0; r < ph; ++r) {
      uint8_t* line = raw.data() + off + 1;
      unfilter(raw[off], line, prev, lb, bpp);
      prev = line;
      off += 1 + lb;
      uint8_t* dst_row = img.rgb.data() + ((size_t)(P.y0 + r * P.dy) * width) * 3;
      for (uint32_t c = 0; c < pw; ++c) {
        uint8_t* d = dst_row + (size_t)(P.x0 + c * P.dx) * 3;
        if (ctype == 3) {
          const size_t e = sample8(line, c);
          if (3 * e + 2 >= palette.size()) throw PngError("png: palette index out of range");
          d[0] = palette[3 * e], d[1] = palette[3 * e + 1], d[2] = palette[3 * e + 2];
        } else if (channels <= 2) {  // grey (+ alpha): replicate, drop alpha
          d[0] = d[1] = d[2] = sample8(line, (size_t)c * channels);
        } else {  // RGB (+ alpha)
          d[0] = sample8(line, (size_t)c * channels), d[1] = sample8(line, (size_t)c * channels + 1);
          d[2] = sample8(line, (size_t)c * channels + 2);
        }
      }
    }
  }
  return img;
}

// width and height from the IHDR chunk alone (what `img.dimensions()` needs for the aspect filter, main.rs:269-270)
inline bool png_dimensions(const std::string& path, int* width, int* height) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  uint8_t head[24];
  const bool ok = std::fread(head, 1, sizeof head, f) == sizeof head;
  std::fclose(f);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (!ok || std::memcmp(head, sig, 8) != 0 || std::memcmp(head + 12, "IHDR", 4) != 0) return false;
  *width = (int)detail::be32(head + 16);
  *height = (int)detail::be32(head + 20);
  return *width > 0 && *height > 0;
}

inline Image decode_png(const std::string& path) {
  const std::vector<uint8_t> data = detail::read_file(path);
  return decode_png_memory(data.data(), data.size());
}

inline std::vector<uint8_t> encode_png_memory(const uint8_t* rgb, int width, int height) {
  using namespace detail;
  if (width <= 0 || height <= 0) throw PngError("png: empty image");
  const size_t lb = (size_t)width * 3;
  std::vector<uint8_t> raw((size_t)height * (1 + lb));
  std::vector<uint8_t> cand(lb), best(lb);
  for (int y = 0; y < height; ++y) {
    const uint8_t* cur = rgb + (size_t)y * lb;
    const uint8_t* prev = y ? cur - lb : nullptr;
    unsigned long best_cost = ~0ul;
    int best_type = 0;
    for (int type = 0; type < 5; ++type) {
      unsigned long cost = 0;
      for (size_t i = 0; i < lb; ++i) {
        const int a = i >= 3 ? cur[i - 3] : 0, b = prev ? prev[i] : 0, c = (prev && i >= 3) ? prev[i - 3] : 0;
        const int pred = type == 0 ? 0 : type == 1 ? a : type == 2 ? b : type == 3 ? ((a + b) >> 1) : paeth(a, b, c);
        const uint8_t v = (uint8_t)(cur[i] - pred);
        cand[i] = v;
        cost += v < 128 ? v : 256 - v;
      }
      if (cost < best_cost) best_cost = cost, best_type = type, best.swap(cand);
    }
    uint8_t* out = raw.data() + (size_t)y * (1 + lb);
    out[0] = (uint8_t)best_type;
    std::memcpy(out + 1, best.data(), lb);
  }
  uLongf zcap = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(zcap);
  if (compress2(z.data(), &zcap, raw.data(), (uLong)raw.size(), Z_DEFAULT_COMPRESSION) != Z_OK) throw PngError("png: deflate failed");
  z.resize(zcap);

  std::vector<uint8_t> file;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  file.insert(file.end(), sig, sig + 8);
  auto chunk = [&](const char* type, const uint8_t* body, size_t len) {
    uint8_t hdr[8];
    put_be32(hdr, (uint32_t)len);
    std::memcpy(hdr + 4, type, 4);
    file.insert(file.end(), hdr, hdr + 8);
    if (len) file.insert(file.end(), body, body + len);
    uLong crc = crc32(crc32(0L, Z_NULL, 0), (const Bytef*)type, 4);
    if (len) crc = crc32(crc, body, (uInt)len);
    uint8_t tail[4];
    put_be32(tail, (uint32_t)crc);
    file.insert(file.end(), tail, tail + 4);
  };
  uint8_t ihdr[13];
  put_be32(ihdr, (uint32_t)width);
  put_be32(ihdr + 4, (uint32_t)height);
  ihdr[8] = 8, ihdr[9] = 2, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;
  chunk("IHDR", ihdr, 13);
  for (size_t o = 0; o < z.size(); o += (size_t)1 << 30) chunk("IDAT", z.data() + o, std::min<size_t>((size_t)1 << 30, z.size() - o));
  chunk("IEND", nullptr, 0);
  return file;
}

inline void encode_png(const std::string& path, const uint8_t* rgb, int width, int height) {
  const std::vector<uint8_t> file = encode_png_memory(rgb, width, height);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw PngError("png: cannot create " + path);
  const bool ok = std::fwrite(file.data(), 1, file.size(), f) == file.size();
  if (std::fclose(f) != 0 || !ok) throw PngError("png: short write to " + path);
}
inline void encode_png(const std::string& path, const Image& img) { encode_png(path, img.rgb.data(), img.width, img.height); }

}  // namespace codec
}  // namespace mosaic_b200
