// mosaic_jpeg.hpp — JPEG in and out for the host side of the path (C++17, header-only, no library).
//
// the reference reads every material and the target with `image::open(path)` and hands the pixels
// on as `img.to_rgb8()` (src/main.rs:263-301, 484, 548, 632); JPEG is what photo libraries are made of (main.rs:182 accepts
// png / jpg / jpeg / webp).  This is a decoder of the ISO/IEC 10918-1 (ITU-T T.81) Huffman processes this image has no
// library for: baseline and extended sequential DCT (SOF0 / SOF1) and progressive DCT (SOF2), 8-bit samples, one
// (grey) or three (YCbCr, or RGB when an Adobe marker says so) components, any sampling factors, restart intervals,
// interleaved and non-interleaved scans.  Not covered (reported as errors): arithmetic coding, lossless, hierarchical,
// 12-bit samples, four-component (CMYK) files.
//
// A JPEG decoder is only specified up to the accuracy of its inverse DCT, its chroma upsampling and its colour
// conversion.  [3P] The reference decodes through the `image` crate 0.25.6 (zune-jpeg): that arithmetic is not available
// here.  What this decoder restates instead is the arithmetic of the IJG / libjpeg-turbo defaults — the 13-bit
// "islow" integer inverse DCT (jidctint), "fancy" triangle-filter upsampling of h2v1 / h2v2 / h1v2 chroma (jdsample),
// the 16-bit fixed-point YCbCr -> RGB tables (jdcolor) — because that is the decoder this image CAN be checked against:
// tests/test_jpeg_codec.py compares every pixel with Pillow (libjpeg-turbo) and expects equality.  Against the
// reference's decoder the result is pinned only as far as two conforming decoders agree (typically +-1 per channel):
// parity unpinned, like the other [3P] stages.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace mosaic_b200 {
namespace codec {

struct JpegError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct JpegImage {  // interleaved RGB8, row-major
  int width = 0, height = 0;
  std::vector<uint8_t> rgb;
};

namespace jpeg_detail {

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huffman {  // T.81 Annex C / F.2.2.3: canonical codes by length
  bool defined = false;
  uint8_t bits[17] = {};
  uint8_t vals[256] = {};
  int mincode[17] = {}, maxcode[18] = {}, valptr[17] = {};
  uint16_t fast[512];  // 9-bit lookup: (length << 8) | value, 0 = longer code
  void build() {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      valptr[l] = k;
      mincode[l] = code;
      code += bits[l];
      k += bits[l];
      if (code > (1 << l)) throw JpegError("jpeg: over-subscribed Huffman table");  // (more codes of this length than bit patterns)
      maxcode[l] = bits[l] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7FFFFFFF;
    std::memset(fast, 0, sizeof fast);
    code = 0;
    k = 0;
    for (int l = 1; l <= 9; ++l) {
      for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
        const int first = code << (9 - l);
        for (int f = 0; f < (1 << (9 - l)); ++f) fast[first + f] = (uint16_t)((l << 8) | vals[k]);
      }
      code <<= 1;
    }
    defined = true;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0;
  int td = 0, ta = 0;           // tables of the current scan
  int blocks_w = 0, blocks_h = 0;  // blocks covering the component's own size (non-interleaved scans)
  int stride_blocks = 0, rows_blocks = 0;  // allocated blocks (whole MCUs)
  int width = 0, height = 0;    // downsampled size in samples
  int pred = 0;
  std::vector<int16_t> coef;    // 64 per block, natural order
  std::vector<uint8_t> plane;   // decoded samples, stride_blocks * 8 wide
};

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int n = 0;
  bool hit_marker = false;
  void fill() {
    while (n <= 24) {
      int byte = 0;
      if (!hit_marker && p < end) {
        byte = *p;
        if (byte == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) {
            p += 2;
          } else {
            hit_marker = true;  // a marker: the entropy-coded segment ends here; feed zeros
            byte = 0;
          }
        } else {
          ++p;
        }
      }
      acc |= (uint32_t)byte << (24 - n);
      n += 8;
    }
  }
  int peek(int k) {
    if (n < k) fill();
    return (int)(acc >> (32 - k));
  }
  void skip(int k) {
    acc <<= k;
    n -= k;
  }
  int get(int k) {
    if (k == 0) return 0;
    const int v = peek(k);
    skip(k);
    return v;
  }
  int bit() { return get(1); }
  void reset() {
    acc = 0;
    n = 0;
    hit_marker = false;
  }
};

inline int decode_symbol(BitReader& br, const Huffman& h) {
  const int look = br.peek(9);
  const uint16_t f = h.fast[look];
  if (f) {
    br.skip(f >> 8);
    return f & 255;
  }
  int code = br.peek(16);
  for (int l = 10; l <= 16; ++l) {
    const int c = code >> (16 - l);
    if (h.maxcode[l] >= 0 && c <= h.maxcode[l] && c >= h.mincode[l]) {
      br.skip(l);
      return h.vals[h.valptr[l] + c - h.mincode[l]];
    }
  }
  throw JpegError("jpeg: bad Huffman code");
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }  // T.81 F.2.2.1

// jidctint.c (IJG "islow"): 13-bit constants, two passes, exact integer arithmetic
inline void idct_islow(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
  constexpr int CB = 13, P1 = 2;
  constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
                F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069,
                F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;
  auto descale = [](long x, int n) { return (int)((x + (1L << (n - 1))) >> n); };
  int ws[64];
  for (int c = 0; c < 8; ++c) {
    const long i0 = coef[c] * (long)q[c], i1 = coef[8 + c] * (long)q[8 + c], i2 = coef[16 + c] * (long)q[16 + c],
               i3 = coef[24 + c] * (long)q[24 + c], i4 = coef[32 + c] * (long)q[32 + c], i5 = coef[40 + c] * (long)q[40 + c],
               i6 = coef[48 + c] * (long)q[48 + c], i7 = coef[56 + c] * (long)q[56 + c];
    long z2 = i2, z3 = i6;
    long z1 = (z2 + z3) * F_0_541196100;
    long tmp2 = z1 + z3 * (-F_1_847759065), tmp3 = z1 + z2 * F_0_765366865;
    z2 = i0, z3 = i4;
    long tmp0 = (z2 + z3) * (1L << CB), tmp1 = (z2 - z3) * (1L << CB);  // (a multiplication: the operands may be negative)
    const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = i7, tmp1 = i5, tmp2 = i3, tmp3 = i1;
    z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
    long z4 = tmp1 + tmp3;
    const long z5 = (z3 + z4) * F_1_175875602;
    tmp0 *= F_0_298631336, tmp1 *= F_2_053119869, tmp2 *= F_3_072711026, tmp3 *= F_1_501321110;
    z1 *= -F_0_899976223, z2 *= -F_2_562915447, z3 *= -F_1_961570560, z4 *= -F_0_390180644;
    z3 += z5, z4 += z5;
    tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
    ws[c] = descale(tmp10 + tmp3, CB - P1), ws[56 + c] = descale(tmp10 - tmp3, CB - P1);
    ws[8 + c] = descale(tmp11 + tmp2, CB - P1), ws[48 + c] = descale(tmp11 - tmp2, CB - P1);
    ws[16 + c] = descale(tmp12 + tmp1, CB - P1), ws[40 + c] = descale(tmp12 - tmp1, CB - P1);
    ws[24 + c] = descale(tmp13 + tmp0, CB - P1), ws[32 + c] = descale(tmp13 - tmp0, CB - P1);
  }
  for (int r = 0; r < 8; ++r) {
    const int* w = ws + 8 * r;
    long z2 = w[2], z3 = w[6];
    long z1 = (z2 + z3) * F_0_541196100;
    long tmp2 = z1 + z3 * (-F_1_847759065), tmp3 = z1 + z2 * F_0_765366865;
    long tmp0 = ((long)w[0] + w[4]) * (1L << CB), tmp1 = ((long)w[0] - w[4]) * (1L << CB);
    const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = w[7], tmp1 = w[5], tmp2 = w[3], tmp3 = w[1];
    z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
    long z4 = tmp1 + tmp3;
    const long z5 = (z3 + z4) * F_1_175875602;
    tmp0 *= F_0_298631336, tmp1 *= F_2_053119869, tmp2 *= F_3_072711026, tmp3 *= F_1_501321110;
    z1 *= -F_0_899976223, z2 *= -F_2_562915447, z3 *= -F_1_961570560, z4 *= -F_0_390180644;
    z3 += z5, z4 += z5;
    tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
    auto put = [&](int i, long v) {
      const int s = descale(v, CB + P1 + 3) + 128;
      out[r * stride + i] = (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s));
    };
    put(0, tmp10 + tmp3), put(7, tmp10 - tmp3), put(1, tmp11 + tmp2), put(6, tmp11 - tmp2);
    put(2, tmp12 + tmp1), put(5, tmp12 - tmp1), put(3, tmp13 + tmp0), put(4, tmp13 - tmp0);
  }
}

}  // namespace jpeg_detail

inline JpegImage decode_jpeg_memory(const uint8_t* data, size_t size) {
  using namespace jpeg_detail;
  if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) throw JpegError("jpeg: no SOI marker");
  uint16_t qt[4][64] = {};
  bool qt_defined[4] = {};
  Huffman hdc[4], hac[4];
  std::vector<Component> comp;
  int width = 0, height = 0, hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0;
  bool progressive = false, have_frame = false, adobe = false;
  int adobe_transform = -1, restart_interval = 0;
  size_t pos = 2;
  bool done = false;

  auto be16 = [&](size_t p) -> int {
    if (p + 2 > size) throw JpegError("jpeg: truncated file");
    return (data[p] << 8) | data[p + 1];
  };

  // one scan: entropy-coded data from `pos` on
  auto decode_scan = [&](const std::vector<int>& sc, int Ss, int Se, int Ah, int Al) {
    BitReader br{data + pos, data + size};
    const bool interleaved = sc.size() > 1;
    int mx, my;
    if (interleaved) {
      mx = mcus_x, my = mcus_y;
    } else {
      mx = comp[sc[0]].blocks_w, my = comp[sc[0]].blocks_h;
    }
    int eobrun = 0, todo = restart_interval;
    for (int c : sc) comp[c].pred = 0;

    auto block_baseline = [&](Component& C, int16_t* b) {
      const Huffman& D = hdc[C.td];
      const Huffman& A = hac[C.ta];
      if (!D.defined || !A.defined) throw JpegError("jpeg: scan uses an undefined Huffman table");
      const int s = decode_symbol(br, D);
      if (s > 11) throw JpegError("jpeg: bad DC category");
      const int diff = s ? extend(br.get(s), s) : 0;
      C.pred += diff;
      b[0] = (int16_t)C.pred;
      for (int k = 1; k < 64;) {
        const int rs = decode_symbol(br, A), r = rs >> 4, ss = rs & 15;
        if (ss == 0) {
          if (r != 15) break;  // EOB
          k += 16;
          continue;
        }
        k += r;
        if (k > 63) throw JpegError("jpeg: AC run beyond the block");
        b[kZigzag[k]] = (int16_t)extend(br.get(ss), ss);
        ++k;
      }
    };
    auto block_dc_first = [&](Component& C, int16_t* b) {
      const Huffman& D = hdc[C.td];
      if (!D.defined) throw JpegError("jpeg: scan uses an undefined Huffman table");
      const int s = decode_symbol(br, D);
      if (s > 11) throw JpegError("jpeg: bad DC category");
      C.pred += s ? extend(br.get(s), s) : 0;
      b[0] = (int16_t)(C.pred * (1 << Al));
    };
    auto block_dc_refine = [&](int16_t* b) {
      if (br.bit()) b[0] = (int16_t)(b[0] | (1 << Al));
    };
    auto block_ac_first = [&](Component& C, int16_t* b) {  // T.81 G.1.2.2
      if (eobrun > 0) {
        --eobrun;
        return;
      }
      const Huffman& A = hac[C.ta];
      if (!A.defined) throw JpegError("jpeg: scan uses an undefined Huffman table");
      for (int k = Ss; k <= Se;) {
        const int rs = decode_symbol(br, A), r = rs >> 4, s = rs & 15;
        if (s == 0) {
          if (r < 15) {
            eobrun = (1 << r) - 1;
            if (r) eobrun += br.get(r);
            break;
          }
          k += 16;
          continue;
        }
        k += r;
        if (k > 63) throw JpegError("jpeg: AC run beyond the block");
        b[kZigzag[k]] = (int16_t)(extend(br.get(s), s) * (1 << Al));
        ++k;
      }
    };
    auto block_ac_refine = [&](Component& C, int16_t* b) {  // T.81 G.1.2.3
      const int p1 = 1 << Al, m1 = -(1 << Al);
      const Huffman& A = hac[C.ta];
      if (!A.defined) throw JpegError("jpeg: scan uses an undefined Huffman table");
      auto correct = [&](int16_t& c) {  // a correction bit for a coefficient with history
        if (br.bit() && (c & p1) == 0) c = (int16_t)(c >= 0 ? c + p1 : c + m1);
      };
      int k = Ss;
      if (eobrun == 0) {
        for (; k <= Se; ++k) {
          const int rs = decode_symbol(br, A);
          int r = rs >> 4, s = rs & 15;
          if (s) {
            if (s != 1) throw JpegError("jpeg: bad refinement scan");
            s = br.bit() ? p1 : m1;  // the new coefficient
          } else if (r != 15) {
            eobrun = 1 << r;  // end of band (this block included)
            if (r) eobrun += br.get(r);
            break;
          }
          // skip the coefficients that have history (correcting them) and r that are still zero
          do {
            int16_t& c = b[kZigzag[k]];
            if (c != 0) {
              correct(c);
            } else if (--r < 0) {
              break;  // the target zero coefficient
            }
            ++k;
          } while (k <= Se);
          if (s) {
            if (k > Se) throw JpegError("jpeg: refinement beyond the band");
            b[kZigzag[k]] = (int16_t)s;
          }
        }
      }
      if (eobrun > 0) {  // the rest of the band: correction bits only
        for (; k <= Se; ++k) {
          int16_t& c = b[kZigzag[k]];
          if (c != 0) correct(c);
        }
        --eobrun;
      }
    };
    auto one_block = [&](Component& C, int bx, int by) {
      int16_t* b = C.coef.data() + ((size_t)by * C.stride_blocks + bx) * 64;
      if (!progressive) {
        block_baseline(C, b);
      } else if (Ss == 0) {
        if (Ah == 0) block_dc_first(C, b);
        else block_dc_refine(b);
      } else {
        if (Ah == 0) block_ac_first(C, b);
        else block_ac_refine(C, b);
      }
    };

    for (int y = 0; y < my; ++y)
      for (int x = 0; x < mx; ++x) {
        if (restart_interval && todo == 0) {  // RSTn expected here (T.81 E.2.4)
          br.reset();
          const uint8_t* p = br.p;
          while (p + 1 < br.end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
          if (p + 1 >= br.end) throw JpegError("jpeg: missing restart marker");
          br.p = p + 2;
          for (int c : sc) comp[c].pred = 0;
          eobrun = 0;
          todo = restart_interval;
        }
        if (interleaved) {
          for (int c : sc) {
            Component& C = comp[c];
            for (int v = 0; v < C.v; ++v)
              for (int h = 0; h < C.h; ++h) one_block(C, x * C.h + h, y * C.v + v);
          }
        } else {
          one_block(comp[sc[0]], x, y);
        }
        if (restart_interval) --todo;
      }
    // the next marker
    const uint8_t* p = br.p;
    while (p + 1 < br.end && !(p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF && !(p[1] >= 0xD0 && p[1] <= 0xD7))) ++p;
    pos = (size_t)(p - data);
  };

  while (!done) {
    if (pos + 2 > size) throw JpegError("jpeg: truncated file (no EOI)");
    if (data[pos] != 0xFF) throw JpegError("jpeg: marker expected");
    while (pos < size && data[pos] == 0xFF) ++pos;  // fill bytes
    if (pos >= size) throw JpegError("jpeg: truncated file (no EOI)");
    const int m = data[pos++];
    if (m == 0xD9) break;                 // EOI
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    const int len = be16(pos);
    if (len < 2 || pos + (size_t)len > size) throw JpegError("jpeg: bad segment length");
    const uint8_t* s = data + pos + 2;
    const int n = len - 2;
    switch (m) {
      case 0xDB: {  // DQT
        for (int i = 0; i < n;) {
          const int pq = s[i] >> 4, tq = s[i] & 15;
          ++i;
          if (tq > 3) throw JpegError("jpeg: bad quantisation table id");
          if (i + (pq ? 128 : 64) > n) throw JpegError("jpeg: truncated DQT");
          for (int k = 0; k < 64; ++k) {
            qt[tq][kZigzag[k]] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i];
            i += pq ? 2 : 1;
          }
          qt_defined[tq] = true;
        }
        break;
      }
      case 0xC4: {  // DHT
        for (int i = 0; i < n;) {
          const int tc = s[i] >> 4, th = s[i] & 15;
          ++i;
          if (tc > 1 || th > 3 || i + 16 > n) throw JpegError("jpeg: bad Huffman table");
          Huffman& H = tc ? hac[th] : hdc[th];
          int total = 0;
          H.bits[0] = 0;
          for (int l = 1; l <= 16; ++l) total += (H.bits[l] = s[i + l - 1]);
          i += 16;
          if (total > 256 || i + total > n) throw JpegError("jpeg: bad Huffman table");
          std::memcpy(H.vals, s + i, (size_t)total);
          i += total;
          H.build();
        }
        break;
      }
      case 0xC0: case 0xC1: case 0xC2: {  // SOF0 / SOF1 / SOF2
        if (have_frame) throw JpegError("jpeg: more than one frame");
        if (n < 6) throw JpegError("jpeg: bad SOF");
        if (s[0] != 8) throw JpegError("jpeg: only 8-bit samples are supported");
        height = (s[1] << 8) | s[2];
        width = (s[3] << 8) | s[4];
        const int nf = s[5];
        if (width <= 0 || height <= 0) throw JpegError("jpeg: bad dimensions");
        // A header can claim 65535 x 65535 pixels in front of a few bytes of data.  Every 8x8 block costs at least one bit of
        // entropy-coded data per scan that covers its DC term, so a real file holds fewer than ~350 pixels per byte (4:2:0,
        // progressive); refuse what is far beyond that, and what would not fit the 4 GiB of RGB this library hands out.
        if ((uint64_t)width * (uint64_t)height > (uint64_t)size * 2048u + 65536u) throw JpegError("jpeg: the file is too short for the dimensions in its header");
        if ((uint64_t)width * (uint64_t)height * 3u > ((uint64_t)1 << 32)) throw JpegError("jpeg: more than 4 GiB of pixels");
        if (nf != 1 && nf != 3) throw JpegError("jpeg: only 1- and 3-component files are supported");
        if (n < 6 + 3 * nf) throw JpegError("jpeg: bad SOF");
        progressive = m == 0xC2;
        comp.resize((size_t)nf);
        for (int i = 0; i < nf; ++i) {
          comp[i].id = s[6 + 3 * i];
          comp[i].h = s[7 + 3 * i] >> 4;
          comp[i].v = s[7 + 3 * i] & 15;
          comp[i].tq = s[8 + 3 * i];
          if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3) throw JpegError("jpeg: bad component");
          hmax = std::max(hmax, comp[i].h);
          vmax = std::max(vmax, comp[i].v);
        }
        mcus_x = (width + 8 * hmax - 1) / (8 * hmax);
        mcus_y = (height + 8 * vmax - 1) / (8 * vmax);
        for (auto& C : comp) {
          C.width = (width * C.h + hmax - 1) / hmax;
          C.height = (height * C.v + vmax - 1) / vmax;
          C.blocks_w = (C.width + 7) / 8;
          C.blocks_h = (C.height + 7) / 8;
          C.stride_blocks = mcus_x * C.h;
          C.rows_blocks = mcus_y * C.v;
          C.coef.assign((size_t)C.stride_blocks * C.rows_blocks * 64, 0);
        }
        have_frame = true;
        break;
      }
      case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
        throw JpegError("jpeg: unsupported process (lossless, hierarchical or arithmetic coding)");
      case 0xDD:  // DRI
        if (n < 2) throw JpegError("jpeg: bad DRI");
        restart_interval = (s[0] << 8) | s[1];
        break;
      case 0xEE:  // APP14 "Adobe"
        if (n >= 12 && !std::memcmp(s, "Adobe", 5)) {
          adobe = true;
          adobe_transform = s[11];
        }
        break;
      case 0xDA: {  // SOS
        if (!have_frame) throw JpegError("jpeg: scan before the frame header");
        if (n < 1) throw JpegError("jpeg: bad SOS");
        const int ns = s[0];
        if (ns < 1 || ns > (int)comp.size() || n < 1 + 2 * ns + 3) throw JpegError("jpeg: bad SOS");
        std::vector<int> sc;
        for (int i = 0; i < ns; ++i) {
          int ci = -1;
          for (size_t j = 0; j < comp.size(); ++j)
            if (comp[j].id == s[1 + 2 * i]) ci = (int)j;
          if (ci < 0) throw JpegError("jpeg: scan names an unknown component");
          comp[ci].td = s[2 + 2 * i] >> 4;
          comp[ci].ta = s[2 + 2 * i] & 15;
          if (comp[ci].td > 3 || comp[ci].ta > 3) throw JpegError("jpeg: bad table selector");
          sc.push_back(ci);
        }
        const int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns], Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
        if (progressive) {
          if (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || Al > 13) throw JpegError("jpeg: bad progressive scan parameters");
        }
        pos += (size_t)len;
        decode_scan(sc, progressive ? Ss : 0, progressive ? Se : 63, progressive ? Ah : 0, progressive ? Al : 0);
        continue;  // (pos already stands at the next marker)
      }
      default: break;  // APPn, COM, DNL ...: skipped
    }
    pos += (size_t)len;
  }
  if (!have_frame) throw JpegError("jpeg: no frame");

  // ---- inverse DCT of every block
  for (auto& C : comp) {
    if (!qt_defined[C.tq]) throw JpegError("jpeg: component uses an undefined quantisation table");
    const int stride = C.stride_blocks * 8;
    C.plane.assign((size_t)stride * C.rows_blocks * 8, 0);
    for (int by = 0; by < C.rows_blocks; ++by)
      for (int bx = 0; bx < C.stride_blocks; ++bx)
        idct_islow(C.coef.data() + ((size_t)by * C.stride_blocks + bx) * 64, qt[C.tq], C.plane.data() + (size_t)by * 8 * stride + bx * 8, stride);
    C.coef.clear();
    C.coef.shrink_to_fit();
  }

  // ---- upsampling to full resolution (jdsample.c): fancy h2v1 / h2v2 / h1v2 when the component is wide / tall enough,
  //      integral replication otherwise; rows and columns beyond the component's real size replicate its edge
  auto upsample = [&](const Component& C) -> std::vector<uint8_t> {
    const int stride = C.stride_blocks * 8;
    std::vector<uint8_t> full((size_t)width * height);
    const int hx = hmax / C.h, vx = vmax / C.v;
    if (hmax % C.h || vmax % C.v) throw JpegError("jpeg: fractional sampling ratios are not supported");
    auto in = [&](int x, int y) -> int {
      x = std::min(std::max(x, 0), C.width - 1);
      y = std::min(std::max(y, 0), C.height - 1);
      return C.plane[(size_t)y * stride + x];
    };
    if (hx == 1 && vx == 1) {
      for (int y = 0; y < height; ++y) std::memcpy(&full[(size_t)y * width], &C.plane[(size_t)y * stride], (size_t)width);
    } else if (hx == 2 && vx == 1 && C.width > 2) {  // h2v1_fancy_upsample
      for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
          const int i = x >> 1;
          int v;
          if (x & 1)
            v = i == C.width - 1 ? in(i, y) : (in(i, y) * 3 + in(i + 1, y) + 2) >> 2;
          else
            v = i == 0 ? in(i, y) : (in(i, y) * 3 + in(i - 1, y) + 1) >> 2;
          full[(size_t)y * width + x] = (uint8_t)v;
        }
    } else if (hx == 2 && vx == 2 && C.width > 2) {  // h2v2_fancy_upsample
      for (int y = 0; y < height; ++y) {
        const int j = y >> 1, jn = (y & 1) ? j + 1 : j - 1;  // nearer row j, farther row jn (edge rows replicate)
        auto colsum = [&](int i) { return in(i, j) * 3 + in(i, jn); };
        for (int x = 0; x < width; ++x) {
          const int i = x >> 1;
          const int t = colsum(i);
          int v;
          if (x & 1)
            v = i == C.width - 1 ? (t * 4 + 7) >> 4 : (t * 3 + colsum(i + 1) + 7) >> 4;
          else
            v = i == 0 ? (t * 4 + 8) >> 4 : (t * 3 + colsum(i - 1) + 8) >> 4;
          full[(size_t)y * width + x] = (uint8_t)v;
        }
      }
    } else if (hx == 1 && vx == 2) {  // h1v2_fancy_upsample (libjpeg-turbo)
      for (int y = 0; y < height; ++y) {
        const int j = y >> 1, jn = (y & 1) ? j + 1 : j - 1, bias = (y & 1) ? 2 : 1;
        for (int x = 0; x < width; ++x) full[(size_t)y * width + x] = (uint8_t)((in(x, j) * 3 + in(x, jn) + bias) >> 2);
      }
    } else {  // int_upsample / h2v1_upsample / h2v2_upsample: replication
      for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) full[(size_t)y * width + x] = (uint8_t)in(x / hx, y / vx);
    }
    return full;
  };

  JpegImage img;
  img.width = width, img.height = height;
  img.rgb.resize((size_t)width * height * 3);
  if (comp.size() == 1) {
    const std::vector<uint8_t> y = upsample(comp[0]);
    for (size_t i = 0; i < y.size(); ++i) img.rgb[3 * i] = img.rgb[3 * i + 1] = img.rgb[3 * i + 2] = y[i];
    return img;
  }
  const std::vector<uint8_t> c0 = upsample(comp[0]), c1 = upsample(comp[1]), c2 = upsample(comp[2]);
  // colour space (jdapimin.c default_decompress_parms): JFIF or no marker -> YCbCr, unless the ids spell "RGB";
  // Adobe transform 0 -> RGB
  bool ycc = true;
  if (adobe && adobe_transform == 0) ycc = false;
  if (!adobe && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B') ycc = false;
  if (!ycc) {
    for (size_t i = 0; i < c0.size(); ++i) img.rgb[3 * i] = c0[i], img.rgb[3 * i + 1] = c1[i], img.rgb[3 * i + 2] = c2[i];
    return img;
  }
  // jdcolor.c build_ycc_rgb_table / ycc_rgb_convert: 16-bit fixed point
  int cr_r[256], cb_b[256];
  long cr_g[256], cb_g[256];
  for (int i = 0; i < 256; ++i) {
    const long x = i - 128;
    cr_r[i] = (int)((91881L * x + 32768L) >> 16);   // FIX(1.40200)
    cb_b[i] = (int)((116130L * x + 32768L) >> 16);  // FIX(1.77200)
    cr_g[i] = -46802L * x;                           // FIX(0.71414)
    cb_g[i] = -22554L * x + 32768L;                  // FIX(0.34414)
  }
  auto clamp8 = [](int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
  for (size_t i = 0; i < c0.size(); ++i) {
    const int y = c0[i], cb = c1[i], cr = c2[i];
    img.rgb[3 * i] = clamp8(y + cr_r[cr]);
    img.rgb[3 * i + 1] = clamp8(y + (int)((cb_g[cb] + cr_g[cr]) >> 16));
    img.rgb[3 * i + 2] = clamp8(y + cb_b[cb]);
  }
  return img;
}

inline JpegImage decode_jpeg(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw JpegError("jpeg: cannot open " + path);
  std::vector<uint8_t> data;
  uint8_t buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + got);
  std::fclose(f);
  return decode_jpeg_memory(data.data(), data.size());
}

// ---- encoder: output_img.save("x.jpg") (src/main.rs:668-672 with a .jpg / .jpeg path) ---------------------------------
// Baseline sequential DCT, 8-bit, YCbCr 4:4:4 (no chroma subsampling), the Annex K quantisation tables scaled by the IJG
// quality formula, the Annex K Huffman tables, a JFIF APP0 header — the shape of the file the `image` crate's encoder
// writes at its default quality 75 [3P, restated from memory of image 0.25's JpegEncoder: the exact bytes of a lossy
// encoder are nobody's contract].  Forward DCT: jfdctint ("islow"), colour conversion: jccolor's 16-bit tables.
namespace jpeg_detail {

static const uint8_t kStdLumaQ[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
                                      69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55, 64,
                                      81, 104, 113, 92, 49, 64,  78,  87,  103, 121, 120, 101, 72, 92,  95,  98,  112, 100, 103, 99};
static const uint8_t kStdChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                        99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                        99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const uint8_t kDcLumaBits[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcChromaBits[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumaBits[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
    0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
    0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
    0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
    0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t kAcChromaBits[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
    0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
    0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
    0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
    0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
    0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

// jfdctint.c ("islow") on a block of level-shifted samples; output scaled by 8 (the quantiser divides it out)
inline void fdct_islow(int* d) {
  constexpr int CB = 13, P1 = 2;
  constexpr long F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
                 F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069,
                 F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;
  auto descale = [](long x, int n) { return (int)((x + (1L << (n - 1))) >> n); };
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < 8; ++i) {
      int* p = pass == 0 ? d + 8 * i : d + i;
      const int st = pass == 0 ? 1 : 8;
      const long t0 = p[0] + p[7 * st], t7 = p[0] - p[7 * st], t1 = p[st] + p[6 * st], t6 = p[st] - p[6 * st];
      const long t2 = p[2 * st] + p[5 * st], t5 = p[2 * st] - p[5 * st], t3 = p[3 * st] + p[4 * st], t4 = p[3 * st] - p[4 * st];
      const long t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
      if (pass == 0) {
        p[0] = (int)((t10 + t11) * (1L << P1));
        p[4 * st] = (int)((t10 - t11) * (1L << P1));
      } else {
        p[0] = descale(t10 + t11, P1);
        p[4 * st] = descale(t10 - t11, P1);
      }
      const int sh = pass == 0 ? CB - P1 : CB + P1;
      long z1 = (t12 + t13) * F_0_541196100;
      p[2 * st] = descale(z1 + t13 * F_0_765366865, sh);
      p[6 * st] = descale(z1 + t12 * (-F_1_847759065), sh);
      z1 = t4 + t7;
      long z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
      const long z5 = (z3 + z4) * F_1_175875602;
      long a4 = t4 * F_0_298631336, a5 = t5 * F_2_053119869, a6 = t6 * F_3_072711026, a7 = t7 * F_1_501321110;
      z1 *= -F_0_899976223, z2 *= -F_2_562915447, z3 *= -F_1_961570560, z4 *= -F_0_390180644;
      z3 += z5, z4 += z5;
      p[7 * st] = descale(a4 + z1 + z3, sh);
      p[5 * st] = descale(a5 + z2 + z4, sh);
      p[3 * st] = descale(a6 + z2 + z3, sh);
      p[st] = descale(a7 + z1 + z4, sh);
    }
  }
}

struct BitWriter {
  std::vector<uint8_t>& out;
  uint32_t acc = 0;
  int n = 0;
  void put(unsigned code, int len) {
    acc = (acc << len) | (code & ((1u << len) - 1u));
    n += len;
    while (n >= 8) {
      const uint8_t byte = (uint8_t)(acc >> (n - 8));
      out.push_back(byte);
      if (byte == 0xFF) out.push_back(0x00);
      n -= 8;
    }
  }
  void flush() {
    if (n > 0) put((1u << (8 - n)) - 1u, 8 - n);  // pad with ones
  }
};

struct EncTable {
  uint16_t code[256] = {};
  uint8_t len[256] = {};
  void build(const uint8_t* bits, const uint8_t* vals) {
    int code_ = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      for (int i = 0; i < bits[l]; ++i, ++k) {
        code[vals[k]] = (uint16_t)code_++;
        len[vals[k]] = (uint8_t)l;
      }
      code_ <<= 1;
    }
  }
};

}  // namespace jpeg_detail

inline std::vector<uint8_t> encode_jpeg_memory(const uint8_t* rgb, int width, int height, int quality = 75) {
  using namespace jpeg_detail;
  if (width <= 0 || height <= 0 || width > 65535 || height > 65535) throw JpegError("jpeg: cannot encode an image of that size");
  quality = std::min(100, std::max(1, quality));
  const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;  // jpeg_quality_scaling
  uint8_t q[2][64];
  for (int t = 0; t < 2; ++t)
    for (int i = 0; i < 64; ++i) {
      const long v = ((long)(t ? kStdChromaQ[i] : kStdLumaQ[i]) * scale + 50) / 100;
      q[t][i] = (uint8_t)std::min(255L, std::max(1L, v));  // (baseline: 8-bit entries)
    }
  std::vector<uint8_t> out;
  auto marker = [&](int m, const std::vector<uint8_t>& body) {
    out.push_back(0xFF), out.push_back((uint8_t)m);
    const size_t len = body.size() + 2;
    out.push_back((uint8_t)(len >> 8)), out.push_back((uint8_t)len);
    out.insert(out.end(), body.begin(), body.end());
  };
  out.push_back(0xFF), out.push_back(0xD8);
  marker(0xE0, {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
  for (int t = 0; t < 2; ++t) {
    std::vector<uint8_t> b(1, (uint8_t)t);
    for (int k = 0; k < 64; ++k) b.push_back(q[t][kZigzag[k]]);
    marker(0xDB, b);
  }
  marker(0xC0, {8, (uint8_t)(height >> 8), (uint8_t)height, (uint8_t)(width >> 8), (uint8_t)width, 3, 1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1});
  auto dht = [&](int tc_th, const uint8_t* bits, const uint8_t* vals, int nvals) {
    std::vector<uint8_t> b(1, (uint8_t)tc_th);
    b.insert(b.end(), bits + 1, bits + 17);
    b.insert(b.end(), vals, vals + nvals);
    marker(0xC4, b);
  };
  dht(0x00, kDcLumaBits, kDcVals, 12), dht(0x10, kAcLumaBits, kAcLumaVals, 162);
  dht(0x01, kDcChromaBits, kDcVals, 12), dht(0x11, kAcChromaBits, kAcChromaVals, 162);
  marker(0xDA, {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
  EncTable dc[2], ac[2];
  dc[0].build(kDcLumaBits, kDcVals), dc[1].build(kDcChromaBits, kDcVals);
  ac[0].build(kAcLumaBits, kAcLumaVals), ac[1].build(kAcChromaBits, kAcChromaVals);
  BitWriter bw{out};
  int pred[3] = {0, 0, 0};
  const int bx = (width + 7) / 8, by = (height + 7) / 8;
  for (int yb = 0; yb < by; ++yb)
    for (int xb = 0; xb < bx; ++xb) {
      int blk[3][64];
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) {
          const int sx = std::min(xb * 8 + x, width - 1), sy = std::min(yb * 8 + y, height - 1);  // edge replication
          const uint8_t* p = rgb + ((size_t)sy * width + sx) * 3;
          const long r = p[0], g = p[1], b = p[2];
          // jccolor.c rgb_ycc_convert: 16-bit fixed point
          blk[0][8 * y + x] = (int)((19595L * r + 38470L * g + 7471L * b + 32768L) >> 16) - 128;
          blk[1][8 * y + x] = (int)((-11059L * r - 21709L * g + 32768L * b + (128L << 16) + 32767L) >> 16) - 128;
          blk[2][8 * y + x] = (int)((32768L * r - 27439L * g - 5329L * b + (128L << 16) + 32767L) >> 16) - 128;
        }
      for (int c = 0; c < 3; ++c) {
        fdct_islow(blk[c]);
        const int t = c ? 1 : 0;
        int zz[64];
        for (int k = 0; k < 64; ++k) {
          const int v = blk[c][kZigzag[k]], qv = q[t][kZigzag[k]] * 8;  // (the DCT output is scaled by 8)
          zz[k] = v < 0 ? -((-v + qv / 2) / qv) : (v + qv / 2) / qv;
        }
        auto category = [](int v) {
          int a = v < 0 ? -v : v, s = 0;
          while (a) ++s, a >>= 1;
          return s;
        };
        const int diff = zz[0] - pred[c];
        pred[c] = zz[0];
        int s = category(diff);
        bw.put(dc[t].code[s], dc[t].len[s]);
        if (s) bw.put((unsigned)(diff < 0 ? diff - 1 : diff), s);
        int run = 0;
        for (int k = 1; k < 64; ++k) {
          if (zz[k] == 0) {
            ++run;
            continue;
          }
          while (run > 15) bw.put(ac[t].code[0xF0], ac[t].len[0xF0]), run -= 16;
          s = category(zz[k]);
          bw.put(ac[t].code[(run << 4) | s], ac[t].len[(run << 4) | s]);
          bw.put((unsigned)(zz[k] < 0 ? zz[k] - 1 : zz[k]), s);
          run = 0;
        }
        if (run) bw.put(ac[t].code[0x00], ac[t].len[0x00]);
      }
    }
  bw.flush();
  out.push_back(0xFF), out.push_back(0xD9);
  return out;
}

inline void encode_jpeg(const std::string& path, const uint8_t* rgb, int width, int height, int quality = 75) {
  const std::vector<uint8_t> file = encode_jpeg_memory(rgb, width, height, quality);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw JpegError("jpeg: cannot create " + path);
  const bool ok = std::fwrite(file.data(), 1, file.size(), f) == file.size();
  if (std::fclose(f) != 0 || !ok) throw JpegError("jpeg: short write to " + path);
}

// width and height from the frame header alone (the aspect filter, main.rs:269-270)
inline bool jpeg_dimensions(const std::string& path, int* width, int* height) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<uint8_t> d(1 << 16);
  d.resize(std::fread(d.data(), 1, d.size(), f));
  std::fclose(f);
  if (d.size() < 4 || d[0] != 0xFF || d[1] != 0xD8) return false;
  size_t pos = 2;
  while (pos + 4 <= d.size()) {
    if (d[pos] != 0xFF) return false;
    while (pos < d.size() && d[pos] == 0xFF) ++pos;
    if (pos >= d.size()) return false;
    const int m = d[pos++];
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD9)) continue;
    if (pos + 2 > d.size()) return false;
    const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
    if ((m >= 0xC0 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      if (pos + 7 > d.size()) return false;
      *height = (d[pos + 3] << 8) | d[pos + 4];
      *width = (d[pos + 5] << 8) | d[pos + 6];
      return *width > 0 && *height > 0;
    }
    if (m == 0xDA) return false;
    pos += len;
  }
  return false;
}

}  // namespace codec
}  // namespace mosaic_b200
