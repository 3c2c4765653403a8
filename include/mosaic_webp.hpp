// mosaic_webp.hpp — WebP in and out for the host side of the path (C++17, header-only, no library).
//
// the reference takes `.webp` materials (src/main.rs:181 lists png / jpg / jpeg / webp) and reads them
// like every other file with `image::open(path)?.to_rgb8()` (src/main.rs:263-301).  The `image` crate hands WebP to
// `image-webp` [3P, absent from /root/reference — restated from the published formats, not from that crate's source]:
//
//   * the RIFF container ("WebP Container Specification"): simple files (`VP8 ` / `VP8L`) and extended ones (`VP8X` with
//     ICCP / EXIF / XMP / ALPH chunks around the image chunk).  Animated files are refused.
//   * VP8L, the lossless bitstream ("WebP Lossless Bitstream Specification"): prefix codes (simple and normal code-length
//     codes), meta prefix image, colour cache, LZ77 backward references with the 120-entry distance map, and the four
//     transforms (predictor with its 14 modes, cross-colour, subtract-green, colour indexing with pixel bundling).
//     Lossless means what it says: the pixels are the file's, whatever decoder reads them.
//   * VP8, the lossy key frame (RFC 6386): boolean decoder, segment / filter / quantiser headers, token partitions,
//     intra prediction (16x16, chroma, 4x4 with its ten sub-block modes), inverse WHT / DCT, simple and normal loop
//     filters.  RFC 6386 fixes the YUV planes bit for bit; the way to RGB is the decoder's choice — this one takes
//     libwebp's defaults (the 9-3-3-1 "fancy" chroma upsampler and its 14-bit fixed-point BT.601 conversion), so a file
//     decodes to the pixels libwebp (and Pillow, which the tests compare with) gives.  image-webp follows the same
//     arithmetic from 0.2 on [3P-memory: parity with the reference's own decoder unpinned, like the JPEG decoder's].
//
// `to_rgb8()` drops alpha without blending (image's `Rgba -> Rgb` conversion), so the ALPH chunk of a lossy file and the
// alpha channel of a lossless one are not needed for the pixels the path consumes.
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mosaic_png.hpp"  // codec::Image, detail::read_file

namespace mosaic_b200 {
namespace codec {

struct WebpError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

namespace webp_detail {

inline uint32_t le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t le24(const uint8_t* p) { return le16(p) | ((uint32_t)p[2] << 16); }
inline uint32_t le32(const uint8_t* p) { return le24(p) | ((uint32_t)p[3] << 24); }
inline int ceil_shift(int v, int bits) { return (v + (1 << bits) - 1) >> bits; }

// ------------------------------------------------------------------------------------------------------------------------
// VP8L (lossless)
// ------------------------------------------------------------------------------------------------------------------------

// bits come least-significant first (spec §"Details of Reading Bits")
struct LsbReader {
  const uint8_t* p;
  size_t n, pos = 0;
  uint64_t acc = 0;
  int have = 0;
  bool past_end = false;
  LsbReader(const uint8_t* data, size_t size) : p(data), n(size) {}
  inline void fill() {
    while (have <= 56) {
      if (pos < n) {
        acc |= (uint64_t)p[pos++] << have;
      } else {
        if (pos >= n + 8) past_end = true;  // (a whole word of padding consumed: the stream ended early)
        ++pos;
      }
      have += 8;
    }
  }
  inline uint32_t peek(int nb) {
    if (have < nb) fill();
    return (uint32_t)(acc & ((1ull << nb) - 1));
  }
  inline void skip(int nb) { acc >>= nb, have -= nb; }
  inline uint32_t read(int nb) {
    if (nb == 0) return 0;
    const uint32_t v = peek(nb);
    skip(nb);
    return v;
  }
};

// canonical prefix code; the first bit read is the most significant bit of the code word
struct PrefixCode {
  static constexpr int FAST = 8;
  int single = -1;  // the only symbol of a one-symbol code (decoded without reading a bit), else -1
  uint16_t count[16] = {};
  std::vector<uint16_t> symbols;  // by (length, symbol)
  uint16_t fast[1 << FAST] = {};  // length << 12 | symbol for code words of <= FAST bits, 0: longer

  void build(const std::vector<uint8_t>& len) {
    int used = 0, last = -1;
    std::fill(count, count + 16, 0);
    for (size_t s = 0; s < len.size(); ++s) {
      if (len[s] > 15) throw WebpError("webp: prefix code length above 15");
      if (len[s]) ++count[len[s]], ++used, last = (int)s;
    }
    if (used == 0) throw WebpError("webp: prefix code without symbols");
    if (used == 1) {
      single = last;
      return;
    }
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) throw WebpError("webp: over-subscribed prefix code");
    }
    if (left != 0) throw WebpError("webp: incomplete prefix code");
    uint16_t offs[16] = {};
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    symbols.assign(used, 0);
    for (size_t s = 0; s < len.size(); ++s)
      if (len[s]) symbols[offs[len[s]]++] = (uint16_t)s;
    std::fill(fast, fast + (1 << FAST), 0);
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= FAST; ++l) {
      for (int i = 0; i < count[l]; ++i, ++code, ++idx) {
        uint32_t rev = 0;
        for (int b = 0; b < l; ++b) rev |= ((code >> b) & 1u) << (l - 1 - b);
        for (uint32_t f = rev; f < (1u << FAST); f += 1u << l) fast[f] = (uint16_t)((l << 12) | symbols[idx]);
      }
      code <<= 1;
    }
  }

  inline int decode(LsbReader& br) const {
    if (single >= 0) return single;
    const uint32_t look = br.peek(15);
    const uint16_t f = fast[look & ((1u << FAST) - 1)];
    if (f) {
      br.skip(f >> 12);
      return f & 0x0FFF;
    }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; ++l) {
      code |= (int)((look >> (l - 1)) & 1u);
      const int c = count[l];
      if (code - c < first) {
        br.skip(l);
        return symbols[index + (code - first)];
      }
      index += c, first += c;
      first <<= 1, code <<= 1;
    }
    throw WebpError("webp: invalid prefix code word");
  }
};

// §"Decoding of Meta Prefix Codes" / "Normal Code Length Code" / "Simple Code Length Code"
inline void read_prefix_code(LsbReader& br, int alphabet, PrefixCode& out) {
  std::vector<uint8_t> len((size_t)alphabet, 0);
  if (br.read(1)) {  // simple: one or two symbols
    const int nsym = (int)br.read(1) + 1;
    const int s0 = (int)br.read(br.read(1) ? 8 : 1);
    if (s0 >= alphabet) throw WebpError("webp: simple prefix code symbol out of range");
    len[s0] = 1;
    if (nsym == 2) {
      const int s1 = (int)br.read(8);
      if (s1 >= alphabet) throw WebpError("webp: simple prefix code symbol out of range");
      len[s1] = 1;
    }
  } else {
    static const uint8_t order[19] = {17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
    std::vector<uint8_t> cl(19, 0);
    const int ncl = 4 + (int)br.read(4);
    for (int i = 0; i < ncl; ++i) cl[order[i]] = (uint8_t)br.read(3);
    PrefixCode lengths;
    lengths.build(cl);
    int max_tokens = alphabet;
    if (br.read(1)) {
      const int nbits = 2 + 2 * (int)br.read(3);
      max_tokens = 2 + (int)br.read(nbits);
      if (max_tokens > alphabet) throw WebpError("webp: more code lengths than symbols");
    }
    int sym = 0, prev = 8;
    while (sym < alphabet) {
      if (max_tokens-- == 0) break;
      const int c = lengths.decode(br);
      if (c < 16) {
        len[sym++] = (uint8_t)c;
        if (c) prev = c;
      } else {
        static const int extra[3] = {2, 3, 7}, base[3] = {3, 3, 11};
        const int rep = (int)br.read(extra[c - 16]) + base[c - 16];
        if (sym + rep > alphabet) throw WebpError("webp: code length run past the alphabet");
        const uint8_t v = c == 16 ? (uint8_t)prev : (uint8_t)0;
        for (int i = 0; i < rep; ++i) len[sym++] = v;
      }
    }
  }
  if (br.past_end) throw WebpError("webp: truncated lossless stream");
  out.build(len);
}

// LZ77 prefix coding of lengths and distances
inline int prefixed_value(LsbReader& br, int sym) {
  if (sym < 4) return sym + 1;
  const int extra = (sym - 2) >> 1;
  const int offset = (2 + (sym & 1)) << extra;
  return offset + (int)br.read(extra) + 1;
}

// the 120 short distance codes: (xi, yi) neighbourhood in the order of the specification, packed as yi << 4 | (8 - xi)
inline int plane_code_to_distance(int xsize, int code) {
  static const uint8_t map[120] = {
      0x18, 0x07, 0x17, 0x19, 0x28, 0x06, 0x27, 0x29, 0x16, 0x1a, 0x26, 0x2a, 0x38, 0x05, 0x37, 0x39, 0x15, 0x1b, 0x36, 0x3a,
      0x25, 0x2b, 0x48, 0x04, 0x47, 0x49, 0x14, 0x1c, 0x35, 0x3b, 0x46, 0x4a, 0x24, 0x2c, 0x58, 0x45, 0x4b, 0x34, 0x3c, 0x03,
      0x57, 0x59, 0x13, 0x1d, 0x56, 0x5a, 0x23, 0x2d, 0x44, 0x4c, 0x55, 0x5b, 0x33, 0x3d, 0x68, 0x02, 0x67, 0x69, 0x12, 0x1e,
      0x66, 0x6a, 0x22, 0x2e, 0x54, 0x5c, 0x43, 0x4d, 0x65, 0x6b, 0x32, 0x3e, 0x78, 0x01, 0x77, 0x79, 0x53, 0x5d, 0x11, 0x1f,
      0x64, 0x6c, 0x42, 0x4e, 0x76, 0x7a, 0x21, 0x2f, 0x75, 0x7b, 0x31, 0x3f, 0x63, 0x6d, 0x52, 0x5e, 0x00, 0x74, 0x7c, 0x41,
      0x4f, 0x10, 0x20, 0x62, 0x6e, 0x30, 0x73, 0x7d, 0x51, 0x5f, 0x40, 0x72, 0x7e, 0x61, 0x6f, 0x50, 0x71, 0x7f, 0x60, 0x70};
  if (code > 120) return code - 120;
  const int m = map[code - 1];
  const int dist = (m >> 4) * xsize + (8 - (m & 15));
  return dist >= 1 ? dist : 1;
}

// an entropy-coded ARGB image: the main image (`main`: may carry a meta prefix image) or one of the sub-images
inline std::vector<uint32_t> read_entropy_image(LsbReader& br, int w, int h, bool main) {
  int cache_bits = 0;
  if (br.read(1)) {
    cache_bits = (int)br.read(4);
    if (cache_bits < 1 || cache_bits > 11) throw WebpError("webp: colour cache size out of range");
  }
  int meta_bits = 0, meta_w = 0, ngroups = 1;
  std::vector<uint32_t> meta;
  if (main && br.read(1)) {
    meta_bits = (int)br.read(3) + 2;
    meta_w = ceil_shift(w, meta_bits);
    meta = read_entropy_image(br, meta_w, ceil_shift(h, meta_bits), false);
    for (uint32_t& v : meta) {
      v = (v >> 8) & 0xFFFF;
      ngroups = std::max(ngroups, (int)v + 1);
    }
  }
  const int alphabets[5] = {256 + 24 + (cache_bits ? 1 << cache_bits : 0), 256, 256, 256, 40};
  std::vector<std::array<PrefixCode, 5>> groups((size_t)ngroups);
  for (auto& g : groups)
    for (int j = 0; j < 5; ++j) read_prefix_code(br, alphabets[j], g[j]);

  std::vector<uint32_t> cache(cache_bits ? (size_t)1 << cache_bits : 0, 0);
  const int cache_shift = 32 - cache_bits;
  auto remember = [&](uint32_t px) {
    if (cache_bits) cache[(uint32_t)(px * 0x1e35a7bdu) >> cache_shift] = px;
  };

  const size_t total = (size_t)w * h;
  std::vector<uint32_t> px(total);
  size_t pos = 0;
  int x = 0, y = 0;
  const std::array<PrefixCode, 5>* g = &groups[0];
  const int meta_mask = meta_bits ? (1 << meta_bits) - 1 : ~0;
  bool pick = true;
  while (pos < total) {
    if (pick) {
      if (meta_bits) g = &groups[meta[(size_t)(y >> meta_bits) * meta_w + (x >> meta_bits)]];
      pick = false;
    }
    const int code = (*g)[0].decode(br);
    if (code < 256) {
      const uint32_t red = (uint32_t)(*g)[1].decode(br), blue = (uint32_t)(*g)[2].decode(br), alpha = (uint32_t)(*g)[3].decode(br);
      const uint32_t v = (alpha << 24) | (red << 16) | ((uint32_t)code << 8) | blue;
      px[pos++] = v;
      remember(v);
      if (++x == w) x = 0, ++y;
      if ((x & meta_mask) == 0) pick = true;
    } else if (code < 256 + 24) {
      const int length = prefixed_value(br, code - 256);
      const int dist = plane_code_to_distance(w, prefixed_value(br, (*g)[4].decode(br)));
      if ((size_t)dist > pos || total - pos < (size_t)length) throw WebpError("webp: backward reference outside the image");
      for (int i = 0; i < length; ++i, ++pos) {
        px[pos] = px[pos - dist];
        remember(px[pos]);
      }
      x += length;
      while (x >= w) x -= w, ++y;
      pick = true;
    } else {
      const int key = code - (256 + 24);
      if (key >= (int)cache.size()) throw WebpError("webp: colour cache index out of range");
      px[pos++] = cache[key];
      if (++x == w) x = 0, ++y;
      if ((x & meta_mask) == 0) pick = true;
    }
    if (br.past_end) throw WebpError("webp: truncated lossless stream");
  }
  return px;
}

inline uint32_t avg2(uint32_t a, uint32_t b) { return (((a ^ b) & 0xFEFEFEFEu) >> 1) + (a & b); }
inline uint32_t add_px(uint32_t a, uint32_t b) {
  const uint32_t ag = (a & 0xFF00FF00u) + (b & 0xFF00FF00u), rb = (a & 0x00FF00FFu) + (b & 0x00FF00FFu);
  return (ag & 0xFF00FF00u) | (rb & 0x00FF00FFu);
}
inline int clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

inline uint32_t select_px(uint32_t L, uint32_t T, uint32_t TL) {
  int pl = 0, pt = 0;  // Manhattan distance of the estimate L + T - TL to L (= |T - TL|) and to T (= |L - TL|)
  for (int s = 0; s < 32; s += 8) {
    const int l = (L >> s) & 255, t = (T >> s) & 255, tl = (TL >> s) & 255;
    pl += std::abs(t - tl), pt += std::abs(l - tl);
  }
  return pl < pt ? L : T;
}
inline uint32_t clamp_add_sub_full(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r = 0;
  for (int s = 0; s < 32; s += 8) r |= (uint32_t)clip255((int)((a >> s) & 255) + (int)((b >> s) & 255) - (int)((c >> s) & 255)) << s;
  return r;
}
inline uint32_t clamp_add_sub_half(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int s = 0; s < 32; s += 8) {
    const int x = (a >> s) & 255, y = (b >> s) & 255;
    r |= (uint32_t)clip255(x + (x - y) / 2) << s;
  }
  return r;
}

inline uint32_t predict(int mode, uint32_t L, uint32_t T, uint32_t TL, uint32_t TR) {
  switch (mode) {
    case 1: return L;
    case 2: return T;
    case 3: return TR;
    case 4: return TL;
    case 5: return avg2(avg2(L, TR), T);
    case 6: return avg2(L, TL);
    case 7: return avg2(L, T);
    case 8: return avg2(TL, T);
    case 9: return avg2(T, TR);
    case 10: return avg2(avg2(L, TL), avg2(T, TR));
    case 11: return select_px(L, T, TL);
    case 12: return clamp_add_sub_full(L, T, TL);
    case 13: return clamp_add_sub_half(avg2(L, T), TL);
    default: return 0xFF000000u;  // 0, and the two unused codes
  }
}

struct Transform {
  int type = 0, bits = 0, xsize = 0;  // xsize: the image width when the transform was read
  std::vector<uint32_t> data;
};

inline void undo_predictor(const Transform& t, std::vector<uint32_t>& px, int h) {
  const int w = t.xsize, bw = ceil_shift(w, t.bits);
  if (w == 0 || h == 0) return;
  px[0] = add_px(px[0], 0xFF000000u);
  for (int x = 1; x < w; ++x) px[x] = add_px(px[x], px[x - 1]);
  for (int y = 1; y < h; ++y) {
    uint32_t* cur = &px[(size_t)y * w];
    const uint32_t* up = cur - w;
    cur[0] = add_px(cur[0], up[0]);
    const uint32_t* modes = &t.data[(size_t)(y >> t.bits) * bw];
    for (int x = 1; x < w; ++x) {
      const int mode = (modes[x >> t.bits] >> 8) & 15;
      // (up[x + 1] of the last column is the first pixel of the current row: the memory order the format defines)
      cur[x] = add_px(cur[x], predict(mode, cur[x - 1], up[x], up[x - 1], up[x + 1]));
    }
  }
}

inline int colour_delta(int8_t t, int8_t c) { return ((int)t * (int)c) >> 5; }

inline void undo_cross_colour(const Transform& t, std::vector<uint32_t>& px, int h) {
  const int w = t.xsize, bw = ceil_shift(w, t.bits);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const uint32_t m = t.data[(size_t)(y >> t.bits) * bw + (x >> t.bits)];
      const int8_t g2r = (int8_t)(m & 255), g2b = (int8_t)((m >> 8) & 255), r2b = (int8_t)((m >> 16) & 255);
      uint32_t& v = px[(size_t)y * w + x];
      const int8_t green = (int8_t)((v >> 8) & 255);
      int red = (v >> 16) & 255, blue = v & 255;
      red = (red + colour_delta(g2r, green)) & 255;
      blue = (blue + colour_delta(g2b, green)) & 255;
      blue = (blue + colour_delta(r2b, (int8_t)red)) & 255;
      v = (v & 0xFF00FF00u) | ((uint32_t)red << 16) | (uint32_t)blue;
    }
}

inline void undo_subtract_green(std::vector<uint32_t>& px) {
  for (uint32_t& v : px) {
    const uint32_t g = (v >> 8) & 255;
    const uint32_t rb = ((v & 0x00FF00FFu) + ((g << 16) | g)) & 0x00FF00FFu;
    v = (v & 0xFF00FF00u) | rb;
  }
}

inline void undo_colour_index(const Transform& t, std::vector<uint32_t>& px, int h) {
  const int w = t.xsize, pw = ceil_shift(w, t.bits);
  const int bpp = 8 >> t.bits, per = 1 << t.bits;
  const uint32_t mask = (1u << bpp) - 1;
  std::vector<uint32_t> out((size_t)w * h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const uint32_t packed = (px[(size_t)y * pw + (x >> t.bits)] >> 8) & 255;
      const uint32_t idx = t.bits ? (packed >> ((x & (per - 1)) * bpp)) & mask : packed;
      out[(size_t)y * w + x] = idx < t.data.size() ? t.data[idx] : 0u;
    }
  px.swap(out);
}

inline void vp8l_dimensions(const uint8_t* d, size_t n, int* w, int* h) {
  if (n < 5 || d[0] != 0x2F) throw WebpError("webp: bad lossless signature");
  const uint32_t v = le32(d + 1);
  *w = (int)(v & 0x3FFF) + 1, *h = (int)((v >> 14) & 0x3FFF) + 1;
  if ((v >> 29) != 0) throw WebpError("webp: unknown lossless version");
}

inline Image decode_vp8l(const uint8_t* d, size_t n) {
  int width, height;
  vp8l_dimensions(d, n, &width, &height);
  LsbReader br(d + 5, n - 5);
  std::vector<Transform> transforms;
  int xsize = width;
  unsigned seen = 0;
  while (br.read(1)) {
    Transform t;
    t.type = (int)br.read(2);
    if (seen & (1u << t.type)) throw WebpError("webp: transform used twice");
    seen |= 1u << t.type;
    t.xsize = xsize;
    if (t.type == 0 || t.type == 1) {
      t.bits = (int)br.read(3) + 2;
      t.data = read_entropy_image(br, ceil_shift(xsize, t.bits), ceil_shift(height, t.bits), false);
    } else if (t.type == 3) {
      const int ncol = (int)br.read(8) + 1;
      t.bits = ncol > 16 ? 0 : ncol > 4 ? 1 : ncol > 2 ? 2 : 3;
      t.data = read_entropy_image(br, ncol, 1, false);
      for (int i = 1; i < ncol; ++i) t.data[i] = add_px(t.data[i], t.data[i - 1]);
      xsize = ceil_shift(xsize, t.bits);
    }
    transforms.push_back(std::move(t));
  }
  std::vector<uint32_t> px = read_entropy_image(br, xsize, height, true);
  for (size_t i = transforms.size(); i-- > 0;) {
    const Transform& t = transforms[i];
    switch (t.type) {
      case 0: undo_predictor(t, px, height); break;
      case 1: undo_cross_colour(t, px, height); break;
      case 2: undo_subtract_green(px); break;
      default: undo_colour_index(t, px, height); break;
    }
  }
  Image img;
  img.width = width, img.height = height;
  img.rgb.resize((size_t)width * height * 3);
  for (size_t i = 0; i < (size_t)width * height; ++i) {
    img.rgb[3 * i] = (uint8_t)(px[i] >> 16), img.rgb[3 * i + 1] = (uint8_t)(px[i] >> 8), img.rgb[3 * i + 2] = (uint8_t)px[i];
  }
  return img;
}

// ------------------------------------------------------------------------------------------------------------------------
// container
// ------------------------------------------------------------------------------------------------------------------------

struct Payload {
  const uint8_t* data = nullptr;
  size_t size = 0;
  bool lossless = false;
};

inline Payload find_image_chunk(const uint8_t* d, size_t n) {
  if (n < 12 || std::memcmp(d, "RIFF", 4) != 0 || std::memcmp(d + 8, "WEBP", 4) != 0) throw WebpError("webp: not a RIFF/WEBP file");
  const size_t riff_end = std::min(n, (size_t)le32(d + 4) + 8);
  size_t pos = 12;
  while (pos + 8 <= riff_end) {
    const uint8_t* tag = d + pos;
    const size_t len = le32(d + pos + 4);
    if (len > riff_end - pos - 8) throw WebpError("webp: truncated chunk");
    const uint8_t* body = d + pos + 8;
    if (!std::memcmp(tag, "VP8 ", 4)) return Payload{body, len, false};
    if (!std::memcmp(tag, "VP8L", 4)) return Payload{body, len, true};
    if (!std::memcmp(tag, "VP8X", 4)) {
      if (len < 10) throw WebpError("webp: bad VP8X chunk");
      if (body[0] & 0x02) throw WebpError("webp: animated files are not supported");
    }
    if (!std::memcmp(tag, "ANIM", 4) || !std::memcmp(tag, "ANMF", 4)) throw WebpError("webp: animated files are not supported");
    pos += 8 + len + (len & 1);
  }
  throw WebpError("webp: no image chunk");
}

}  // namespace webp_detail
}  // namespace codec
}  // namespace mosaic_b200

// ------------------------------------------------------------------------------------------------------------------------
// VP8 (lossy key frame, RFC 6386)
// ------------------------------------------------------------------------------------------------------------------------

namespace mosaic_b200 {
namespace codec {
namespace webp_detail {

// RFC 6386 §7: the boolean entropy decoder.  Past the end of its partition it reads zero bytes (§7.3: "the decoder
// must not read past the buffer" — what is returned there is unspecified; libwebp shifts in zeros).
struct BoolDecoder {
  const uint8_t* p = nullptr;
  const uint8_t* end = nullptr;
  uint32_t value = 0, range = 255;
  int bit_count = 0;
  int over = 0;  // bytes taken from behind the end: up to two are the decoder's own look-ahead, more means a truncated stream
  inline uint32_t next_byte() {
    if (p < end) return *p++;
    ++over;
    return 0u;
  }
  void init(const uint8_t* data, size_t size) {
    p = data, end = data + size;
    value = 0, over = 0;
    for (int i = 0; i < 2; ++i) value = (value << 8) | next_byte();
    range = 255, bit_count = 0;
  }
  bool truncated() const { return over > 2; }
  inline int bit(int prob) {
    const uint32_t split = 1 + (((range - 1) * (uint32_t)prob) >> 8);
    const uint32_t big = split << 8;
    int r;
    if (value >= big) {
      r = 1, range -= split, value -= big;
    } else {
      r = 0, range = split;
    }
    while (range < 128) {
      value <<= 1, range <<= 1;
      if (++bit_count == 8) {
        bit_count = 0;
        value |= next_byte();
      }
    }
    return r;
  }
  inline int literal(int n) {
    int v = 0;
    while (n-- > 0) v = (v << 1) | bit(128);
    return v;
  }
  inline int signed_literal(int n) {
    const int v = literal(n);
    return bit(128) ? -v : v;
  }
};

#include "mosaic_webp_tables.inc"

static const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
static const uint8_t kBands[16 + 1] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};  // (+ a sentinel for position 16)
// RFC 6386 §13.2: probabilities of the extra bits of the DCT value categories 3..6, zero-terminated
static const uint8_t kCat3[] = {173, 148, 140, 0};
static const uint8_t kCat4[] = {176, 155, 140, 135, 0};
static const uint8_t kCat5[] = {180, 157, 141, 134, 130, 0};
static const uint8_t kCat6[] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0};
static const uint8_t* const kCat3456[4] = {kCat3, kCat4, kCat5, kCat6};

// sub-block modes in the numbering of kBModeProbs; the four whole-block modes share the first four numbers
enum { B_DC = 0, B_TM, B_VE, B_HE, B_RD, B_VR, B_LD, B_VL, B_HD, B_HU };
// RFC 6386 §11.2 bmode_tree: entry pair 2i, 2i+1 = the children of node i (<= 0: minus a mode), node i uses probability i
static const int8_t kBModeTree[18] = {-B_DC, 1, -B_TM, 2, -B_VE, 3, 4, 6, -B_HE, 5, -B_RD, -B_VR, -B_LD, 7, -B_VL, 8, -B_HD, -B_HU};

struct Plane {  // one reconstructed plane with a border of one row above and one column left (the 127 / 129 of §12.2)
  int w = 0, h = 0, stride = 0;
  std::vector<uint8_t> buf;
  void init(int width, int height) {
    w = width, h = height, stride = width + 1;
    buf.assign((size_t)stride * (height + 1), 0);
    std::fill(buf.begin(), buf.begin() + stride, (uint8_t)127);
    for (int y = 0; y < height; ++y) buf[(size_t)(y + 1) * stride] = 129;
  }
  inline uint8_t* at(int x, int y) { return &buf[(size_t)(y + 1) * stride + (x + 1)]; }
};

inline uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

// §14.3 inverse Walsh-Hadamard transform of the Y2 block: its 16 outputs are the DC terms of the 16 luma blocks
inline void inverse_wht(const int16_t* in, int16_t* coeffs /* 16 blocks x 16 */) {
  int tmp[16];
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[0 + i] + in[12 + i], a1 = in[4 + i] + in[8 + i], a2 = in[4 + i] - in[8 + i], a3 = in[0 + i] - in[12 + i];
    tmp[0 + i] = a0 + a1, tmp[8 + i] = a0 - a1, tmp[4 + i] = a3 + a2, tmp[12 + i] = a3 - a2;
  }
  for (int i = 0; i < 4; ++i) {
    const int dc = tmp[0 + i * 4] + 3;
    const int a0 = dc + tmp[3 + i * 4], a1 = tmp[1 + i * 4] + tmp[2 + i * 4], a2 = tmp[1 + i * 4] - tmp[2 + i * 4], a3 = dc - tmp[3 + i * 4];
    coeffs[(i * 4 + 0) * 16] = (int16_t)((a0 + a1) >> 3);
    coeffs[(i * 4 + 1) * 16] = (int16_t)((a3 + a2) >> 3);
    coeffs[(i * 4 + 2) * 16] = (int16_t)((a0 - a1) >> 3);
    coeffs[(i * 4 + 3) * 16] = (int16_t)((a3 - a2) >> 3);
  }
}

// §14.3 / §14.4 inverse DCT of one 4x4 block, added to the prediction in place
inline void inverse_dct_add(const int16_t* in, uint8_t* dst, int stride) {
  auto mul1 = [](int a) { return ((a * 20091) >> 16) + a; };
  auto mul2 = [](int a) { return (a * 35468) >> 16; };
  int tmp[16];
  for (int i = 0; i < 4; ++i) {  // vertical pass
    const int a = in[i] + in[8 + i], b = in[i] - in[8 + i];
    const int c = mul2(in[4 + i]) - mul1(in[12 + i]), d = mul1(in[4 + i]) + mul2(in[12 + i]);
    tmp[4 * i + 0] = a + d, tmp[4 * i + 1] = b + c, tmp[4 * i + 2] = b - c, tmp[4 * i + 3] = a - d;
  }
  for (int i = 0; i < 4; ++i) {  // horizontal pass
    const int dc = tmp[i] + 4;
    const int a = dc + tmp[8 + i], b = dc - tmp[8 + i];
    const int c = mul2(tmp[4 + i]) - mul1(tmp[12 + i]), d = mul1(tmp[4 + i]) + mul2(tmp[12 + i]);
    uint8_t* o = dst + (size_t)i * stride;
    o[0] = clip8(o[0] + ((a + d) >> 3));
    o[1] = clip8(o[1] + ((b + c) >> 3));
    o[2] = clip8(o[2] + ((b - c) >> 3));
    o[3] = clip8(o[3] + ((a - d) >> 3));
  }
}

// §12.3: the ten 4x4 predictors.  top[-1] = above-left, top[0..7] = above and above-right, left[0..3].
inline void predict4(int mode, const uint8_t* top, const uint8_t* left, uint8_t* dst, int stride) {
  auto avg2 = [](int a, int b) { return (uint8_t)((a + b + 1) >> 1); };
  auto avg3 = [](int a, int b, int c) { return (uint8_t)((a + 2 * b + c + 2) >> 2); };
  const int X = top[-1], A = top[0], B = top[1], C = top[2], D = top[3], E = top[4], F = top[5], G = top[6], H = top[7];
  const int I = left[0], J = left[1], K = left[2], L = left[3];
  auto put = [&](int x, int y, uint8_t v) { dst[(size_t)y * stride + x] = v; };
  switch (mode) {
    case B_DC: {
      const int dc = (A + B + C + D + I + J + K + L + 4) >> 3;
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) put(x, y, (uint8_t)dc);
      break;
    }
    case B_TM:
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) put(x, y, clip8(top[x] + left[y] - X));
      break;
    case B_VE: {
      const uint8_t v[4] = {avg3(X, A, B), avg3(A, B, C), avg3(B, C, D), avg3(C, D, E)};
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) put(x, y, v[x]);
      break;
    }
    case B_HE: {
      const uint8_t v[4] = {avg3(X, I, J), avg3(I, J, K), avg3(J, K, L), avg3(K, L, L)};
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) put(x, y, v[y]);
      break;
    }
    case B_RD: {
      // the diagonal through (x, y) holds one value: index x - y + 3 into the edge L K J I X A B C D
      const int e[9] = {L, K, J, I, X, A, B, C, D};
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) {
          const int i = x - y + 3;  // 0..6, centre of the three taps at e[i + 1]
          put(x, y, avg3(e[i], e[i + 1], e[i + 2]));
        }
      break;
    }
    case B_VR:
      put(0, 0, avg2(X, A)), put(1, 2, avg2(X, A));
      put(1, 0, avg2(A, B)), put(2, 2, avg2(A, B));
      put(2, 0, avg2(B, C)), put(3, 2, avg2(B, C));
      put(3, 0, avg2(C, D));
      put(0, 3, avg3(K, J, I));
      put(0, 2, avg3(J, I, X));
      put(0, 1, avg3(I, X, A)), put(1, 3, avg3(I, X, A));
      put(1, 1, avg3(X, A, B)), put(2, 3, avg3(X, A, B));
      put(2, 1, avg3(A, B, C)), put(3, 3, avg3(A, B, C));
      put(3, 1, avg3(B, C, D));
      break;
    case B_LD: {
      const int e[9] = {A, B, C, D, E, F, G, H, H};
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) put(x, y, avg3(e[x + y], e[x + y + 1], e[x + y + 2]));
      break;
    }
    case B_VL:
      put(0, 0, avg2(A, B));
      put(1, 0, avg2(B, C)), put(0, 2, avg2(B, C));
      put(2, 0, avg2(C, D)), put(1, 2, avg2(C, D));
      put(3, 0, avg2(D, E)), put(2, 2, avg2(D, E));
      put(0, 1, avg3(A, B, C));
      put(1, 1, avg3(B, C, D)), put(0, 3, avg3(B, C, D));
      put(2, 1, avg3(C, D, E)), put(1, 3, avg3(C, D, E));
      put(3, 1, avg3(D, E, F)), put(2, 3, avg3(D, E, F));
      put(3, 2, avg3(E, F, G));
      put(3, 3, avg3(F, G, H));
      break;
    case B_HD:
      put(0, 0, avg2(I, X)), put(2, 1, avg2(I, X));
      put(0, 1, avg2(J, I)), put(2, 2, avg2(J, I));
      put(0, 2, avg2(K, J)), put(2, 3, avg2(K, J));
      put(0, 3, avg2(L, K));
      put(3, 0, avg3(A, B, C));
      put(2, 0, avg3(X, A, B));
      put(1, 0, avg3(I, X, A)), put(3, 1, avg3(I, X, A));
      put(1, 1, avg3(J, I, X)), put(3, 2, avg3(J, I, X));
      put(1, 2, avg3(K, J, I)), put(3, 3, avg3(K, J, I));
      put(1, 3, avg3(L, K, J));
      break;
    default:  // B_HU
      put(0, 0, avg2(I, J));
      put(2, 0, avg2(J, K)), put(0, 1, avg2(J, K));
      put(2, 1, avg2(K, L)), put(0, 2, avg2(K, L));
      put(1, 0, avg3(I, J, K));
      put(3, 0, avg3(J, K, L)), put(1, 1, avg3(J, K, L));
      put(3, 1, avg3(K, L, L)), put(1, 2, avg3(K, L, L));
      put(3, 2, (uint8_t)L), put(2, 2, (uint8_t)L), put(0, 3, (uint8_t)L), put(1, 3, (uint8_t)L), put(2, 3, (uint8_t)L), put(3, 3, (uint8_t)L);
      break;
  }
}

// §12.2: whole-block prediction (16x16 luma, 8x8 chroma) from the row above and the column left of the block
inline void predict_block(int mode, uint8_t* dst, int stride, int size, bool have_top, bool have_left) {
  const uint8_t* top = dst - stride;
  switch (mode) {
    case B_DC: {
      int dc = 128;
      const int shift = size == 16 ? 4 : 3;
      int st = 0, sl = 0;
      for (int i = 0; i < size; ++i) st += top[i], sl += dst[(size_t)i * stride - 1];
      if (have_top && have_left)
        dc = (st + sl + size) >> (shift + 1);
      else if (have_top)
        dc = (st + (size >> 1)) >> shift;
      else if (have_left)
        dc = (sl + (size >> 1)) >> shift;
      for (int y = 0; y < size; ++y) std::memset(dst + (size_t)y * stride, dc, (size_t)size);
      break;
    }
    case B_TM: {
      const int X = top[-1];
      for (int y = 0; y < size; ++y) {
        uint8_t* row = dst + (size_t)y * stride;
        const int l = row[-1];
        for (int x = 0; x < size; ++x) row[x] = clip8(top[x] + l - X);
      }
      break;
    }
    case B_VE:
      for (int y = 0; y < size; ++y) std::memcpy(dst + (size_t)y * stride, top, (size_t)size);
      break;
    default:  // B_HE
      for (int y = 0; y < size; ++y) {
        uint8_t* row = dst + (size_t)y * stride;
        std::memset(row, row[-1], (size_t)size);
      }
      break;
  }
}

// §15 loop filters, on one line of pixels across an edge: p[-4s] .. p[3s], the edge between p[-s] and p[0]
inline int sclamp(int v) { return v < -128 ? -128 : v > 127 ? 127 : v; }
inline int clamp5(int v) { return v < -16 ? -16 : v > 15 ? 15 : v; }
inline bool edge_over_limit(const uint8_t* p, int s, int t2) {  // t2 = 2 * limit + 1
  return 4 * std::abs(p[-s] - p[0]) + std::abs(p[-2 * s] - p[s]) > t2;
}
inline bool interior_smooth(const uint8_t* p, int s, int it) {
  return std::abs(p[-4 * s] - p[-3 * s]) <= it && std::abs(p[-3 * s] - p[-2 * s]) <= it && std::abs(p[-2 * s] - p[-s]) <= it &&
         std::abs(p[3 * s] - p[2 * s]) <= it && std::abs(p[2 * s] - p[s]) <= it && std::abs(p[s] - p[0]) <= it;
}
inline bool high_variance(const uint8_t* p, int s, int thresh) { return std::abs(p[-2 * s] - p[-s]) > thresh || std::abs(p[s] - p[0]) > thresh; }
inline void filter2(uint8_t* p, int s) {  // the two pixels at the edge, outer taps in
  const int p1 = p[-2 * s], p0 = p[-s], q0 = p[0], q1 = p[s];
  const int a = 3 * (q0 - p0) + sclamp(p1 - q1);
  const int a1 = clamp5((a + 4) >> 3), a2 = clamp5((a + 3) >> 3);
  p[-s] = clip8(p0 + a2), p[0] = clip8(q0 - a1);
}
inline void filter4(uint8_t* p, int s) {  // inner edges without high edge variance
  const int p1 = p[-2 * s], p0 = p[-s], q0 = p[0], q1 = p[s];
  const int a = 3 * (q0 - p0);
  const int a1 = clamp5((a + 4) >> 3), a2 = clamp5((a + 3) >> 3), a3 = (a1 + 1) >> 1;
  p[-2 * s] = clip8(p1 + a3), p[-s] = clip8(p0 + a2), p[0] = clip8(q0 - a1), p[s] = clip8(q1 - a3);
}
inline void filter6(uint8_t* p, int s) {  // macroblock edges without high edge variance
  const int p2 = p[-3 * s], p1 = p[-2 * s], p0 = p[-s], q0 = p[0], q1 = p[s], q2 = p[2 * s];
  const int a = sclamp(3 * (q0 - p0) + sclamp(p1 - q1));
  const int a1 = (27 * a + 63) >> 7, a2 = (18 * a + 63) >> 7, a3 = (9 * a + 63) >> 7;
  p[-3 * s] = clip8(p2 + a3), p[-2 * s] = clip8(p1 + a2), p[-s] = clip8(p0 + a1);
  p[0] = clip8(q0 - a1), p[s] = clip8(q1 - a2), p[2 * s] = clip8(q2 - a3);
}
// `across`: the step over the edge, `along`: the step to the next line, n lines
inline void simple_edge(uint8_t* p, int across, int along, int n, int limit) {
  for (int i = 0; i < n; ++i, p += along)
    if (!edge_over_limit(p, across, 2 * limit + 1)) filter2(p, across);
}
inline void normal_edge(uint8_t* p, int across, int along, int n, int limit, int ilimit, int hev, bool mb_edge) {
  for (int i = 0; i < n; ++i, p += along) {
    if (edge_over_limit(p, across, 2 * limit + 1) || !interior_smooth(p, across, ilimit)) continue;
    if (high_variance(p, across, hev))
      filter2(p, across);
    else if (mb_edge)
      filter6(p, across);
    else
      filter4(p, across);
  }
}

struct FilterInfo {
  uint8_t limit = 0, ilevel = 0, hev = 0, inner = 0;
};

inline void vp8_dimensions(const uint8_t* d, size_t n, int* w, int* h) {
  if (n < 10) throw WebpError("webp: truncated lossy header");
  if (d[0] & 1) throw WebpError("webp: not a key frame");
  if (d[3] != 0x9D || d[4] != 0x01 || d[5] != 0x2A) throw WebpError("webp: bad VP8 start code");
  *w = (int)(le16(d + 6) & 0x3FFF), *h = (int)(le16(d + 8) & 0x3FFF);
  if (*w == 0 || *h == 0) throw WebpError("webp: empty frame");
}

// libwebp's default way from the planes to RGB: 14-bit fixed-point BT.601 (limited range) ...
inline uint8_t yuv_clip(int v) { return (v & ~16383) == 0 ? (uint8_t)(v >> 6) : v < 0 ? 0 : 255; }
inline void yuv_to_rgb(int y, int u, int v, uint8_t* rgb) {
  const int yy = (y * 19077) >> 8;
  rgb[0] = yuv_clip(yy + ((v * 26149) >> 8) - 14234);
  rgb[1] = yuv_clip(yy - ((u * 6419) >> 8) - ((v * 13320) >> 8) + 8708);
  rgb[2] = yuv_clip(yy + ((u * 33050) >> 8) - 17685);
}

inline Image decode_vp8(const uint8_t* d, size_t n) {
  int width, height;
  vp8_dimensions(d, n, &width, &height);
  const uint32_t tag = le24(d);
  if (((tag >> 1) & 7) > 3) throw WebpError("webp: unknown VP8 profile");
  const size_t part0 = tag >> 5;
  if (part0 > n - 10) throw WebpError("webp: truncated first partition");
  BoolDecoder br;
  br.init(d + 10, part0);
  br.literal(1);  // colour space
  br.literal(1);  // clamping type (the reconstruction clamps either way)

  // §9.3 segments
  bool use_segment = false, update_map = false, absolute_delta = true;
  int seg_quant[4] = {}, seg_filter[4] = {};
  uint8_t seg_probs[3] = {255, 255, 255};
  use_segment = br.literal(1);
  if (use_segment) {
    update_map = br.literal(1);
    if (br.literal(1)) {
      absolute_delta = br.literal(1);
      for (int s = 0; s < 4; ++s) seg_quant[s] = br.literal(1) ? br.signed_literal(7) : 0;
      for (int s = 0; s < 4; ++s) seg_filter[s] = br.literal(1) ? br.signed_literal(6) : 0;
    }
    if (update_map)
      for (int s = 0; s < 3; ++s) seg_probs[s] = br.literal(1) ? (uint8_t)br.literal(8) : (uint8_t)255;
  }
  // §9.6 loop filter
  const bool simple_filter = br.literal(1);
  const int level = br.literal(6), sharpness = br.literal(3);
  int ref_delta[4] = {}, mode_delta[4] = {};
  const bool use_lf_delta = br.literal(1);
  if (use_lf_delta && br.literal(1)) {
    for (int i = 0; i < 4; ++i)
      if (br.literal(1)) ref_delta[i] = br.signed_literal(6);
    for (int i = 0; i < 4; ++i)
      if (br.literal(1)) mode_delta[i] = br.signed_literal(6);
  }
  // §9.5 token partitions
  const int nparts = 1 << br.literal(2);
  const uint8_t* sizes = d + 10 + part0;
  const uint8_t* end = d + n;
  if ((size_t)(end - sizes) < (size_t)3 * (nparts - 1)) throw WebpError("webp: truncated partition table");
  std::vector<BoolDecoder> parts((size_t)nparts);
  {
    const uint8_t* start = sizes + 3 * (nparts - 1);
    for (int i = 0; i < nparts; ++i) {
      size_t len = i + 1 < nparts ? le24(sizes + 3 * i) : (size_t)(end - start);
      if (len > (size_t)(end - start)) len = (size_t)(end - start);
      parts[i].init(start, len);
      start += len;
    }
  }
  // §9.6 / §14.1 quantiser
  const int base_q = br.literal(7);
  const int dqy1_dc = br.literal(1) ? br.signed_literal(4) : 0, dqy2_dc = br.literal(1) ? br.signed_literal(4) : 0,
            dqy2_ac = br.literal(1) ? br.signed_literal(4) : 0, dquv_dc = br.literal(1) ? br.signed_literal(4) : 0,
            dquv_ac = br.literal(1) ? br.signed_literal(4) : 0;
  struct Quant {
    int y1[2], y2[2], uv[2];
  } quant[4];
  auto clipq = [](int v, int hi) { return v < 0 ? 0 : v > hi ? hi : v; };
  for (int s = 0; s < 4; ++s) {
    int q = base_q;
    if (use_segment) q = absolute_delta ? seg_quant[s] : seg_quant[s] + base_q;
    Quant& m = quant[s];
    m.y1[0] = kDcQ[clipq(q + dqy1_dc, 127)], m.y1[1] = kAcQ[clipq(q, 127)];
    m.y2[0] = kDcQ[clipq(q + dqy2_dc, 127)] * 2;
    m.y2[1] = (kAcQ[clipq(q + dqy2_ac, 127)] * 101581) >> 16;  // x 155 / 100
    if (m.y2[1] < 8) m.y2[1] = 8;
    m.uv[0] = kDcQ[clipq(q + dquv_dc, 117)], m.uv[1] = kAcQ[clipq(q + dquv_ac, 127)];
  }
  // §15.2 filter parameters per segment and block kind
  FilterInfo finfo[4][2];
  for (int s = 0; s < 4; ++s) {
    int base = level;
    if (use_segment) base = absolute_delta ? seg_filter[s] : seg_filter[s] + level;
    for (int i4 = 0; i4 < 2; ++i4) {
      int lv = base;
      if (use_lf_delta) {
        lv += ref_delta[0];
        if (i4) lv += mode_delta[0];
      }
      lv = lv < 0 ? 0 : lv > 63 ? 63 : lv;
      FilterInfo& f = finfo[s][i4];
      if (lv > 0) {
        int il = lv;
        if (sharpness > 0) {
          il >>= sharpness > 4 ? 2 : 1;
          if (il > 9 - sharpness) il = 9 - sharpness;
        }
        if (il < 1) il = 1;
        f.ilevel = (uint8_t)il, f.limit = (uint8_t)(2 * lv + il), f.hev = (uint8_t)(lv >= 40 ? 2 : lv >= 15 ? 1 : 0);
      }
      f.inner = (uint8_t)i4;
    }
  }
  const int filter_type = level == 0 ? 0 : simple_filter ? 1 : 2;

  br.literal(1);  // refresh_entropy_probs: nothing follows a key frame here
  // §13.4 token probabilities
  uint8_t probs[4][8][3][11];
  for (int t = 0, i = 0; t < 4; ++t)
    for (int b = 0; b < 8; ++b)
      for (int c = 0; c < 3; ++c)
        for (int p = 0; p < 11; ++p, ++i) probs[t][b][c][p] = br.bit(kCoeffUpdateProbs[i]) ? (uint8_t)br.literal(8) : kCoeffProbs0[i];
  const bool use_skip = br.literal(1);
  const int skip_prob = use_skip ? br.literal(8) : 0;

  const int mbw = (width + 15) >> 4, mbh = (height + 15) >> 4;
  Plane Y, U, V;
  Y.init(mbw * 16, mbh * 16), U.init(mbw * 8, mbh * 8), V.init(mbw * 8, mbh * 8);
  std::vector<uint8_t> top_modes((size_t)mbw * 4, (uint8_t)B_DC);
  struct NzCtx {
    uint8_t y[4] = {}, u[2] = {}, v[2] = {}, dc = 0;
  };
  std::vector<NzCtx> top_nz((size_t)mbw);
  std::vector<FilterInfo> mb_filter((size_t)mbw * mbh);

  // §13: the coefficients of one block; returns the position after the last decoded coefficient
  auto read_coeffs = [&](BoolDecoder& tb, int type, int ctx, const int* dq, int nstart, int16_t* out) -> int {
    int nn = nstart;
    const uint8_t* p = probs[type][kBands[nn]][ctx];
    for (; nn < 16; ++nn) {
      if (!tb.bit(p[0])) return nn;  // end of block
      while (!tb.bit(p[1])) {        // zeros
        p = probs[type][kBands[++nn]][0];
        if (nn == 16) return 16;
      }
      int v;
      if (!tb.bit(p[2])) {
        v = 1;
        p = probs[type][kBands[nn + 1]][1];
      } else {
        if (!tb.bit(p[3])) {
          v = !tb.bit(p[4]) ? 2 : 3 + tb.bit(p[5]);
        } else if (!tb.bit(p[6])) {
          if (!tb.bit(p[7])) {
            v = 5 + tb.bit(159);
          } else {
            v = 7 + 2 * tb.bit(165);
            v += tb.bit(145);
          }
        } else {
          const int b1 = tb.bit(p[8]), b0 = tb.bit(p[9 + b1]);
          const int cat = 2 * b1 + b0;
          v = 0;
          for (const uint8_t* tab = kCat3456[cat]; *tab; ++tab) v += v + tb.bit(*tab);
          v += 3 + (8 << cat);
        }
        p = probs[type][kBands[nn + 1]][2];
      }
      out[kZigzag[nn]] = (int16_t)((tb.bit(128) ? -v : v) * dq[nn > 0]);
    }
    return 16;
  };

  for (int my = 0; my < mbh; ++my) {
    BoolDecoder& tb = parts[(size_t)(my & (nparts - 1))];
    uint8_t left_modes[4] = {B_DC, B_DC, B_DC, B_DC};
    NzCtx left_nz;
    for (int mx = 0; mx < mbw; ++mx) {
      // §9.3, §11.2-11.4: the macroblock header
      int segment = 0;
      if (update_map) segment = !br.bit(seg_probs[0]) ? br.bit(seg_probs[1]) : br.bit(seg_probs[2]) + 2;
      const bool skip_flag = use_skip ? br.bit(skip_prob) : false;
      const bool i4 = !br.bit(145);
      int ymode = B_DC;
      uint8_t bmodes[16];
      uint8_t* tm = &top_modes[(size_t)mx * 4];
      if (!i4) {
        ymode = br.bit(156) ? (br.bit(128) ? B_TM : B_HE) : (br.bit(163) ? B_VE : B_DC);
        std::memset(tm, ymode, 4), std::memset(left_modes, ymode, 4);
      } else {
        for (int by = 0; by < 4; ++by) {
          int m = left_modes[by];
          for (int bx = 0; bx < 4; ++bx) {
            const uint8_t* pr = &kBModeProbs[((size_t)tm[bx] * 10 + m) * 9];
            int i = kBModeTree[br.bit(pr[0])];
            while (i > 0) i = kBModeTree[2 * i + br.bit(pr[i])];
            m = -i;
            tm[bx] = (uint8_t)m;
            bmodes[by * 4 + bx] = (uint8_t)m;
          }
          left_modes[by] = (uint8_t)m;
        }
      }
      const int uvmode = !br.bit(142) ? B_DC : !br.bit(114) ? B_VE : br.bit(183) ? B_TM : B_HE;

      // §13: residual tokens
      int16_t coeffs[25 * 16];
      std::memset(coeffs, 0, sizeof coeffs);
      uint32_t nz_any = 0;  // (does any block carry a non-zero coefficient)
      uint16_t block_nz = 0;
      uint8_t chroma_nz = 0;
      NzCtx& tnz = top_nz[(size_t)mx];
      const Quant& q = quant[segment];
      if (!skip_flag) {
        int first = 0, ytype = 3;
        if (!i4) {
          int16_t y2[16] = {};
          const int nz = read_coeffs(tb, 1, tnz.dc + left_nz.dc, q.y2, 0, y2);
          tnz.dc = left_nz.dc = (uint8_t)(nz > 0);
          inverse_wht(y2, coeffs);
          first = 1, ytype = 0;
        }
        for (int by = 0; by < 4; ++by)
          for (int bx = 0; bx < 4; ++bx) {
            int16_t* c = coeffs + (by * 4 + bx) * 16;
            const int nz = read_coeffs(tb, ytype, tnz.y[bx] + left_nz.y[by], q.y1, first, c);
            tnz.y[bx] = left_nz.y[by] = (uint8_t)(nz > first);
            if (nz > 1 || c[0] != 0) block_nz |= (uint16_t)(1u << (by * 4 + bx));
          }
        for (int ch = 0; ch < 2; ++ch) {
          uint8_t* tc = ch ? tnz.v : tnz.u;
          uint8_t* lc = ch ? left_nz.v : left_nz.u;
          for (int by = 0; by < 2; ++by)
            for (int bx = 0; bx < 2; ++bx) {
              int16_t* c = coeffs + (16 + ch * 4 + by * 2 + bx) * 16;
              const int nz = read_coeffs(tb, 2, tc[bx] + lc[by], q.uv, 0, c);
              tc[bx] = lc[by] = (uint8_t)(nz > 0);
              if (nz > 0) chroma_nz |= (uint8_t)(1u << (ch * 4 + by * 2 + bx));
            }
        }
        nz_any = (uint32_t)block_nz | ((uint32_t)chroma_nz << 16);
      } else {
        std::memset(tnz.y, 0, 4), std::memset(left_nz.y, 0, 4);
        std::memset(tnz.u, 0, 2), std::memset(tnz.v, 0, 2), std::memset(left_nz.u, 0, 2), std::memset(left_nz.v, 0, 2);
        if (!i4) tnz.dc = left_nz.dc = 0;
      }
      FilterInfo fi = finfo[segment][i4 ? 1 : 0];
      fi.inner |= (uint8_t)(nz_any != 0);
      mb_filter[(size_t)my * mbw + mx] = fi;

      // §12, §14: prediction + residual (from the unfiltered neighbours)
      uint8_t* yd = Y.at(mx * 16, my * 16);
      if (i4) {
        uint8_t above_right[4];  // of the macroblock: shared by its four sub-block rows (§12.3)
        if (my == 0)
          std::memset(above_right, 127, 4);
        else if (mx == mbw - 1)
          std::memset(above_right, yd[-Y.stride + 15], 4);
        else
          std::memcpy(above_right, yd - Y.stride + 16, 4);
        for (int by = 0; by < 4; ++by)
          for (int bx = 0; bx < 4; ++bx) {
            uint8_t* bd = yd + (size_t)(by * 4) * Y.stride + bx * 4;
            uint8_t edge[9], left[4];
            std::memcpy(edge, bd - Y.stride - 1, 5);
            if (bx < 3)
              std::memcpy(edge + 5, bd - Y.stride + 4, 4);
            else
              std::memcpy(edge + 5, above_right, 4);
            for (int i = 0; i < 4; ++i) left[i] = bd[(size_t)i * Y.stride - 1];
            predict4(bmodes[by * 4 + bx], edge + 1, left, bd, Y.stride);
            inverse_dct_add(coeffs + (by * 4 + bx) * 16, bd, Y.stride);
          }
      } else {
        predict_block(ymode, yd, Y.stride, 16, my > 0, mx > 0);
        for (int by = 0; by < 4; ++by)
          for (int bx = 0; bx < 4; ++bx) inverse_dct_add(coeffs + (by * 4 + bx) * 16, yd + (size_t)(by * 4) * Y.stride + bx * 4, Y.stride);
      }
      for (int ch = 0; ch < 2; ++ch) {
        Plane& P = ch ? V : U;
        uint8_t* cd = P.at(mx * 8, my * 8);
        predict_block(uvmode, cd, P.stride, 8, my > 0, mx > 0);
        for (int by = 0; by < 2; ++by)
          for (int bx = 0; bx < 2; ++bx)
            inverse_dct_add(coeffs + (16 + ch * 4 + by * 2 + bx) * 16, cd + (size_t)(by * 4) * P.stride + bx * 4, P.stride);
      }
    }
    // (a stream that ends early would decode on as zeros, at the size its header claims: an error, like libwebp's
    // "premature end of partition")
    if (br.truncated() || tb.truncated()) throw WebpError("webp: truncated lossy stream");
  }

  // §15: loop filter, macroblocks in raster order, each: left edge, inner vertical edges, top edge, inner horizontal edges
  if (filter_type > 0) {
    for (int my = 0; my < mbh; ++my)
      for (int mx = 0; mx < mbw; ++mx) {
        const FilterInfo f = mb_filter[(size_t)my * mbw + mx];
        if (f.limit == 0) continue;
        uint8_t* yd = Y.at(mx * 16, my * 16);
        const int ys = Y.stride;
        if (filter_type == 1) {
          if (mx > 0) simple_edge(yd, 1, ys, 16, f.limit + 4);
          if (f.inner)
            for (int k = 1; k < 4; ++k) simple_edge(yd + 4 * k, 1, ys, 16, f.limit);
          if (my > 0) simple_edge(yd, ys, 1, 16, f.limit + 4);
          if (f.inner)
            for (int k = 1; k < 4; ++k) simple_edge(yd + (size_t)(4 * k) * ys, ys, 1, 16, f.limit);
        } else {
          uint8_t* ud = U.at(mx * 8, my * 8);
          uint8_t* vd = V.at(mx * 8, my * 8);
          const int cs = U.stride;
          if (mx > 0) {
            normal_edge(yd, 1, ys, 16, f.limit + 4, f.ilevel, f.hev, true);
            normal_edge(ud, 1, cs, 8, f.limit + 4, f.ilevel, f.hev, true);
            normal_edge(vd, 1, cs, 8, f.limit + 4, f.ilevel, f.hev, true);
          }
          if (f.inner) {
            for (int k = 1; k < 4; ++k) normal_edge(yd + 4 * k, 1, ys, 16, f.limit, f.ilevel, f.hev, false);
            normal_edge(ud + 4, 1, cs, 8, f.limit, f.ilevel, f.hev, false);
            normal_edge(vd + 4, 1, cs, 8, f.limit, f.ilevel, f.hev, false);
          }
          if (my > 0) {
            normal_edge(yd, ys, 1, 16, f.limit + 4, f.ilevel, f.hev, true);
            normal_edge(ud, cs, 1, 8, f.limit + 4, f.ilevel, f.hev, true);
            normal_edge(vd, cs, 1, 8, f.limit + 4, f.ilevel, f.hev, true);
          }
          if (f.inner) {
            for (int k = 1; k < 4; ++k) normal_edge(yd + (size_t)(4 * k) * ys, ys, 1, 16, f.limit, f.ilevel, f.hev, false);
            normal_edge(ud + (size_t)4 * cs, cs, 1, 8, f.limit, f.ilevel, f.hev, false);
            normal_edge(vd + (size_t)4 * cs, cs, 1, 8, f.limit, f.ilevel, f.hev, false);
          }
        }
      }
  }

  // ... through libwebp's "fancy" upsampler: every pixel takes 9/16 of its nearest chroma sample, 3/16 of the two adjacent ones
  // and 1/16 of the diagonal one, rounded in the two steps libwebp rounds in; the frame edges repeat the nearest sample
  Image img;
  img.width = width, img.height = height;
  img.rgb.resize((size_t)width * height * 3);
  const int cw = (width + 1) >> 1, chh = (height + 1) >> 1;
  for (int y = 0; y < height; ++y) {
    const int c = y >> 1;
    int o = (y & 1) ? c + 1 : c - 1;
    o = o < 0 ? 0 : o >= chh ? chh - 1 : o;
    const uint8_t* yr = Y.at(0, y);
    const uint8_t *un = U.at(0, c), *uo = U.at(0, o), *vn = V.at(0, c), *vo = V.at(0, o);
    uint8_t* out = &img.rgb[(size_t)y * width * 3];
    for (int x = 0; x < width; ++x) {
      const int cx = x >> 1, ox = (x & 1) ? cx + 1 : cx - 1;
      int u, v;
      if (ox < 0 || ox >= cw) {
        u = (3 * un[cx] + uo[cx] + 2) >> 2;
        v = (3 * vn[cx] + vo[cx] + 2) >> 2;
      } else {
        u = (((un[cx] + 3 * un[ox] + 3 * uo[cx] + uo[ox] + 8) >> 3) + un[cx]) >> 1;
        v = (((vn[cx] + 3 * vn[ox] + 3 * vo[cx] + vo[ox] + 8) >> 3) + vn[cx]) >> 1;
      }
      yuv_to_rgb(yr[x], u, v, out + 3 * x);
    }
  }
  return img;
}

}  // namespace webp_detail
}  // namespace codec
}  // namespace mosaic_b200

// ------------------------------------------------------------------------------------------------------------------------
// VP8L out: `RgbImage::save(path)` for a ".webp" path.  The `image` crate (0.25) writes WebP as lossless VP8L; lossless means
// the FILE BYTES are the writer's choice and the pixels are the contract.  This writer: subtract-green transform, one set of
// five prefix codes built from the image's histograms (length-limited to 15 bits), literals only, no colour cache.
// ------------------------------------------------------------------------------------------------------------------------
namespace mosaic_b200 {
namespace codec {
namespace webp_detail {

struct LsbWriter {
  std::vector<uint8_t> out;
  uint64_t acc = 0;
  int have = 0;
  void put(uint32_t v, int nb) {
    acc |= (uint64_t)v << have;
    have += nb;
    while (have >= 8) out.push_back((uint8_t)acc), acc >>= 8, have -= 8;
  }
  void flush() {
    if (have > 0) out.push_back((uint8_t)acc), acc = 0, have = 0;
  }
};

// code lengths (<= limit) of a Huffman code for the histogram: plain Huffman, then the histogram is flattened until it fits
inline std::vector<uint8_t> code_lengths(const std::vector<uint32_t>& hist, int limit) {
  const int n = (int)hist.size();
  std::vector<uint8_t> len((size_t)n, 0);
  std::vector<uint32_t> h(hist);
  for (;;) {
    std::vector<int> used;
    for (int i = 0; i < n; ++i)
      if (h[i]) used.push_back(i);
    if (used.empty()) return len;
    if (used.size() == 1) {
      len[used[0]] = 1;
      return len;
    }
    struct Node {
      uint64_t w;
      int left, right;
    };
    std::vector<Node> nodes;
    std::vector<std::pair<uint64_t, int>> heap;  // (weight, node), smallest first by sort
    for (int s : used) nodes.push_back({h[s], -1, -1}), heap.push_back({h[s], (int)nodes.size() - 1});
    auto cmp = [](const std::pair<uint64_t, int>& a, const std::pair<uint64_t, int>& b) { return a > b; };
    std::make_heap(heap.begin(), heap.end(), cmp);
    while (heap.size() > 1) {
      std::pop_heap(heap.begin(), heap.end(), cmp);
      const auto a = heap.back();
      heap.pop_back();
      std::pop_heap(heap.begin(), heap.end(), cmp);
      const auto b = heap.back();
      heap.pop_back();
      nodes.push_back({a.first + b.first, a.second, b.second});
      heap.push_back({a.first + b.first, (int)nodes.size() - 1});
      std::push_heap(heap.begin(), heap.end(), cmp);
    }
    int deepest = 0;
    std::vector<std::pair<int, int>> stack{{(int)nodes.size() - 1, 0}};
    while (!stack.empty()) {
      const auto [id, depth] = stack.back();
      stack.pop_back();
      if (nodes[id].left < 0) {
        len[used[id]] = (uint8_t)depth;
        deepest = std::max(deepest, depth);
      } else {
        stack.push_back({nodes[id].left, depth + 1}), stack.push_back({nodes[id].right, depth + 1});
      }
    }
    if (deepest <= limit) return len;
    for (int s : used) h[s] = (h[s] >> 1) | 1u;  // flatter histogram, shallower tree
  }
}

// canonical code words, bit-reversed for the LSB-first writer
inline std::vector<uint16_t> code_words(const std::vector<uint8_t>& len) {
  int count[16] = {}, next[16] = {};
  for (uint8_t l : len) ++count[l];
  count[0] = 0;
  for (int l = 1, code = 0; l < 16; ++l) code = (code + count[l - 1]) << 1, next[l] = code;
  std::vector<uint16_t> words(len.size(), 0);
  for (size_t s = 0; s < len.size(); ++s)
    if (len[s]) {
      const int c = next[len[s]]++;
      int rev = 0;
      for (int b = 0; b < len[s]; ++b) rev |= ((c >> b) & 1) << (len[s] - 1 - b);
      words[s] = (uint16_t)rev;
    }
  return words;
}

struct OutCode {
  std::vector<uint8_t> len;
  std::vector<uint16_t> word;
  bool single = false;  // one symbol: the decoder reads no bits for it
  inline void put(LsbWriter& bw, int sym) const {
    if (!single) bw.put(word[sym], len[sym]);
  }
};

// one prefix code into the stream ("Normal Code Length Code", or the simple form for one symbol), returns what to write with
inline OutCode write_prefix_code(LsbWriter& bw, const std::vector<uint32_t>& hist) {
  OutCode oc;
  int used = 0, last = 0;
  for (size_t s = 0; s < hist.size(); ++s)
    if (hist[s]) ++used, last = (int)s;
  if (used <= 1) {
    // simple code, one symbol (symbol 0 for an empty histogram); an 8-bit symbol field covers every symbol this writer emits alone
    bw.put(1, 1), bw.put(0, 1);
    if (last < 2) {
      bw.put(0, 1), bw.put((uint32_t)last, 1);
    } else if (last < 256) {
      bw.put(1, 1), bw.put((uint32_t)last, 8);
    } else {
      throw WebpError("webp: single symbol outside the simple code");
    }
    oc.single = true;
    return oc;
  }
  oc.len = code_lengths(hist, 15);
  oc.word = code_words(oc.len);
  // the code lengths themselves, coded with a code over the 19 length symbols (no run-length symbols used: 0..15 only)
  std::vector<uint32_t> lh(19, 0);
  for (uint8_t l : oc.len) ++lh[l];
  std::vector<uint8_t> ll = code_lengths(lh, 7);
  int distinct = 0;
  for (uint8_t l : ll) distinct += l != 0;
  // (one distinct length: the length code has a single symbol, which the decoder reads with zero bits — none are written below)
  const std::vector<uint16_t> lw = code_words(ll);
  static const uint8_t order[19] = {17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  bw.put(0, 1);       // normal code
  bw.put(19 - 4, 4);  // all 19 code length code lengths
  for (int i = 0; i < 19; ++i) bw.put(ll[order[i]], 3);
  bw.put(0, 1);  // max_symbol = alphabet size
  for (uint8_t l : oc.len)
    if (distinct > 1) bw.put(lw[l], ll[l]);
  return oc;
}

inline std::vector<uint8_t> encode_vp8l(const uint8_t* rgb, int width, int height) {
  if (width < 1 || height < 1 || width > 16384 || height > 16384) throw WebpError("webp: dimensions outside 1..16384");
  LsbWriter bw;
  bw.put(0x2F, 8);
  bw.put((uint32_t)(width - 1), 14), bw.put((uint32_t)(height - 1), 14);
  bw.put(0, 1);  // no alpha
  bw.put(0, 3);  // version
  bw.put(1, 1), bw.put(2, 2);  // one transform: subtract green
  bw.put(0, 1);                // no more transforms
  bw.put(0, 1);                // no colour cache
  bw.put(0, 1);                // no meta prefix image
  const size_t n = (size_t)width * height;
  std::vector<uint32_t> hg(256 + 24, 0), hr(256, 0), hb(256, 0), ha(256, 0), hd(40, 0);
  for (size_t i = 0; i < n; ++i) {
    const uint8_t g = rgb[3 * i + 1];
    ++hg[g], ++hr[(uint8_t)(rgb[3 * i] - g)], ++hb[(uint8_t)(rgb[3 * i + 2] - g)];
  }
  ha[255] = (uint32_t)n;
  const OutCode cg = write_prefix_code(bw, hg), cr = write_prefix_code(bw, hr), cb = write_prefix_code(bw, hb), ca = write_prefix_code(bw, ha),
                cd = write_prefix_code(bw, hd);
  (void)cd;
  for (size_t i = 0; i < n; ++i) {
    const uint8_t g = rgb[3 * i + 1];
    cg.put(bw, g), cr.put(bw, (uint8_t)(rgb[3 * i] - g)), cb.put(bw, (uint8_t)(rgb[3 * i + 2] - g)), ca.put(bw, 255);
  }
  bw.flush();
  std::vector<uint8_t> file;
  auto tag = [&](const char* t) { file.insert(file.end(), t, t + 4); };
  auto u32 = [&](uint32_t v) {
    for (int i = 0; i < 4; ++i) file.push_back((uint8_t)(v >> (8 * i)));
  };
  const uint32_t payload = (uint32_t)bw.out.size(), padded = payload + (payload & 1);
  tag("RIFF"), u32(4 + 8 + padded), tag("WEBP"), tag("VP8L"), u32(payload);
  file.insert(file.end(), bw.out.begin(), bw.out.end());
  if (payload & 1) file.push_back(0);
  return file;
}

}  // namespace webp_detail

inline void encode_webp(const std::string& path, const uint8_t* rgb, int width, int height) {
  const std::vector<uint8_t> file = webp_detail::encode_vp8l(rgb, width, height);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw WebpError("webp: cannot create " + path);
  const bool ok = std::fwrite(file.data(), 1, file.size(), f) == file.size();
  if (std::fclose(f) != 0 || !ok) throw WebpError("webp: cannot write " + path);
}

}  // namespace codec
}  // namespace mosaic_b200

namespace mosaic_b200 {
namespace codec {

inline Image decode_webp_memory(const uint8_t* data, size_t size) {
  const webp_detail::Payload p = webp_detail::find_image_chunk(data, size);
  return p.lossless ? webp_detail::decode_vp8l(p.data, p.size) : webp_detail::decode_vp8(p.data, p.size);
}

inline Image decode_webp(const std::string& path) {
  const std::vector<uint8_t> data = detail::read_file(path);
  return decode_webp_memory(data.data(), data.size());
}

inline bool webp_dimensions(const std::string& path, int* width, int* height) {
  try {
    const std::vector<uint8_t> data = detail::read_file(path);
    const webp_detail::Payload p = webp_detail::find_image_chunk(data.data(), data.size());
    if (p.lossless)
      webp_detail::vp8l_dimensions(p.data, p.size, width, height);
    else
      webp_detail::vp8_dimensions(p.data, p.size, width, height);
    return true;
  } catch (const std::exception&) {
    return false;
  }
}

}  // namespace codec
}  // namespace mosaic_b200
