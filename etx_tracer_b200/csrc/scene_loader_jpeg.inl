// scene_loader_jpeg.inl — JPEG textures (included by scene_loader.cpp inside its anonymous namespace, before scene_loader_formats.inl's read_image).
//
// The reference hands every texture that is not .exr / .hdr / .pfm to stb_image (image_pool.cxx:344-345), so a scene's .jpg files decode through
// stb_image's JPEG reader.  A JPEG decoder is free in its inverse DCT, its chroma upsampling and its colour conversion; to get the SAME RGBA8 bytes the
// three are restated here the way that reader does them (thirdparty/stb_image/stb_image.hxx:2392-2490, 3404-3467, 3596-3623): the 12-bit fixed-point
// separable IDCT with two extra bits kept between the passes, the 3:1 "triangle" upsampling of 2x subsampled planes, and the 20-bit fixed-point YCbCr
// conversion.  Covered: baseline and progressive Huffman JPEG, 8-bit precision, grey or three components (YCbCr, or RGB when the file says so),
// sampling factors 1 and 2 (others: pixel replication, like there), 8- and 16-bit quantisation tables, restart intervals.  Not covered (the file then
// becomes the placeholder; stb_image itself rejects the first three): arithmetic coding, 12-bit, lossless; four-component (CMYK) files.

namespace jpeg {

const uint8_t kZigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};  // the tail absorbs runs that overshoot

struct Huffman {
  int mincode[17] = {}, maxcode[18] = {}, first[17] = {};
  uint8_t values[256] = {};
  bool defined = false;
  bool build(const int* counts) {  // JPEG Annex C: canonical codes by length
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
      first[len] = k;
      mincode[len] = code;
      code += counts[len - 1];
      k += counts[len - 1];
      if (counts[len - 1] && code - 1 >= (1 << len)) return false;
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    defined = true;
    return k <= 256;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dc_pred = 0;
  int x = 0, y = 0, w2 = 0, h2 = 0, coeff_w = 0;
  std::vector<uint8_t> data;
  std::vector<int16_t> coeff;
};

struct Decoder {
  const uint8_t* b = nullptr;
  size_t size = 0, pos = 0;
  const std::string* path = nullptr;
  // entropy-coded segment reader: bits most significant first, 0xFF00 -> 0xFF; at a marker the segment ends and zeros follow
  uint32_t bit_buffer = 0;
  int bit_count = 0;
  int marker = -1;
  Huffman dc[4], ac[4];
  uint16_t dequant[4][64] = {};
  Component comp[3];
  int components = 0, width = 0, height = 0, h_max = 1, v_max = 1, mcu_x = 0, mcu_y = 0;
  bool progressive = false, jfif = false, rgb_ids = false;
  int adobe_transform = -1, restart_interval = 0, todo = 0, eob_run = 0;
  int scan_n = 0, order[3] = {0, 0, 0}, spec_start = 0, spec_end = 63, succ_high = 0, succ_low = 0;

  [[noreturn]] void bad(const char* what) const { fail(*path + ": " + what); }
  int get8() { return pos < size ? b[pos++] : 0; }
  int get16() {
    int hi = get8();
    return (hi << 8) | get8();
  }
  int bit() {
    if (bit_count == 0) {
      uint32_t byte = 0;
      if (marker < 0 && pos < size) {
        byte = b[pos++];
        if (byte == 0xff) {
          int next = pos < size ? b[pos] : 0;
          while (next == 0xff && pos + 1 < size) next = b[++pos];  // fill bytes
          if (next == 0) {
            pos++;
          } else {
            marker = next;
            pos++;
            byte = 0;
          }
        }
      }
      bit_buffer = byte;
      bit_count = 8;
    }
    --bit_count;
    return int((bit_buffer >> bit_count) & 1u);
  }
  int bits(int n) {
    int v = 0;
    for (int i = 0; i < n; ++i) v = (v << 1) | bit();
    return v;
  }
  int decode(const Huffman& h) {
    if (!h.defined) bad("JPEG scan uses an undefined Huffman table");
    int code = 0;
    for (int len = 1; len <= 16; ++len) {
      code = (code << 1) | bit();
      if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) return h.values[h.first[len] + code - h.mincode[len]];
    }
    bad("bad Huffman code in JPEG data");
  }
  int extend(int n) {  // a value of n bits in the sign-magnitude form of the format
    if (n == 0) return 0;
    int v = bits(n);
    return (v < (1 << (n - 1))) ? v - (1 << n) + 1 : v;
  }
  void reset_entropy() {
    bit_buffer = 0, bit_count = 0, marker = -1, eob_run = 0;
    for (auto& c : comp) c.dc_pred = 0;
    todo = restart_interval ? restart_interval : 0x7fffffff;
  }
  // after `restart_interval` units: a restart marker continues the scan, anything else ends it
  bool unit_done() {
    if (--todo > 0) return true;
    if (marker < 0) {
      bit_count = 0;
      while (pos + 1 < size && !(b[pos] == 0xff && b[pos + 1] != 0 && b[pos + 1] != 0xff)) {
        if (b[pos] == 0xff && b[pos + 1] == 0xff) {
          ++pos;
          continue;
        }
        return false;  // entropy data where a marker should be
      }
      if (pos + 1 >= size) return false;
      marker = b[pos + 1];
      pos += 2;
    }
    if (marker < 0xd0 || marker > 0xd7) return false;
    reset_entropy();
    return true;
  }

  void block_baseline(int16_t* data, Component& c) {
    const uint16_t* q = dequant[c.tq];
    int t = decode(dc[c.hd]);
    if (t > 15) bad("bad Huffman code in JPEG data");
    memset(data, 0, 64 * sizeof(int16_t));
    c.dc_pred += extend(t);
    data[0] = int16_t(c.dc_pred * q[0]);
    int k = 1;
    do {
      int rs = decode(ac[c.ha]), s = rs & 15, r = rs >> 4;
      if (s == 0) {
        if (rs != 0xf0) break;
        k += 16;
      } else {
        k += r;
        const int zig = kZigzag[k++];
        data[zig] = int16_t(extend(s) * q[zig]);
      }
    } while (k < 64);
  }
  void block_progressive_dc(int16_t* data, Component& c) {
    if (spec_end != 0) bad("bad progressive JPEG scan");
    if (succ_high == 0) {
      memset(data, 0, 64 * sizeof(int16_t));
      int t = decode(dc[c.hd]);
      if (t > 15) bad("bad Huffman code in JPEG data");
      c.dc_pred += extend(t);
      data[0] = int16_t(c.dc_pred * (1 << succ_low));
    } else if (bit()) {
      data[0] = int16_t(data[0] + int16_t(1 << succ_low));
    }
  }
  void block_progressive_ac(int16_t* data, Component& c) {
    if (spec_start == 0) bad("bad progressive JPEG scan");
    if (succ_high == 0) {
      if (eob_run) {
        --eob_run;
        return;
      }
      int k = spec_start;
      do {
        int rs = decode(ac[c.ha]), s = rs & 15, r = rs >> 4;
        if (s == 0) {
          if (r < 15) {
            eob_run = (1 << r);
            if (r) eob_run += bits(r);
            --eob_run;
            break;
          }
          k += 16;
        } else {
          k += r;
          const int zig = kZigzag[k++];
          data[zig] = int16_t(extend(s) * (1 << succ_low));
        }
      } while (k <= spec_end);
      return;
    }
    const int16_t one = int16_t(1 << succ_low);
    auto refine = [&](int16_t& p) {
      if (bit() && (p & one) == 0) p = int16_t(p > 0 ? p + one : p - one);
    };
    if (eob_run) {
      --eob_run;
      for (int k = spec_start; k <= spec_end; ++k) {
        int16_t& p = data[kZigzag[k]];
        if (p != 0) refine(p);
      }
      return;
    }
    int k = spec_start;
    do {
      int rs = decode(ac[c.ha]), s = rs & 15, r = rs >> 4;
      if (s == 0) {
        if (r < 15) {
          eob_run = (1 << r) - 1;
          if (r) eob_run += bits(r);
          r = 64;  // finish the band
        }
      } else {
        if (s != 1) bad("bad Huffman code in JPEG data");
        s = bit() ? one : -one;
      }
      while (k <= spec_end) {
        int16_t& p = data[kZigzag[k++]];
        if (p != 0) {
          refine(p);
        } else {
          if (r == 0) {
            p = int16_t(s);
            break;
          }
          --r;
        }
      }
    } while (k <= spec_end);
  }

  // one 1-D pass of the inverse DCT in 12-bit fixed point (the classic factorisation with four rotations); outputs scaled by 4096
  // (64-bit intermediates: identical to 32-bit arithmetic on every decodable file, and defined on corrupt ones)
  using Wide = int64_t;
  static inline Wide fx(double v) { return Wide(v * 4096 + 0.5); }
  static void idct_1d(Wide s0, Wide s1, Wide s2, Wide s3, Wide s4, Wide s5, Wide s6, Wide s7, Wide& x0, Wide& x1, Wide& x2, Wide& x3, Wide& t0, Wide& t1, Wide& t2, Wide& t3) {
    Wide p2 = s2, p3 = s6;
    Wide p1 = (p2 + p3) * fx(0.5411961f);
    t2 = p1 + p3 * fx(-1.847759065f);
    t3 = p1 + p2 * fx(0.765366865f);
    p2 = s0, p3 = s4;
    t0 = (p2 + p3) * 4096;
    t1 = (p2 - p3) * 4096;
    x0 = t0 + t3, x3 = t0 - t3, x1 = t1 + t2, x2 = t1 - t2;
    t0 = s7, t1 = s5, t2 = s3, t3 = s1;
    p3 = t0 + t2;
    Wide p4 = t1 + t3;
    p1 = t0 + t3, p2 = t1 + t2;
    Wide p5 = (p3 + p4) * fx(1.175875602f);
    t0 = t0 * fx(0.298631336f), t1 = t1 * fx(2.053119869f), t2 = t2 * fx(3.072711026f), t3 = t3 * fx(1.501321110f);
    p1 = p5 + p1 * fx(-0.899976223f), p2 = p5 + p2 * fx(-2.562915447f);
    p3 = p3 * fx(-1.961570560f), p4 = p4 * fx(-0.390180644f);
    t3 += p1 + p4, t2 += p2 + p3, t1 += p2 + p4, t0 += p1 + p3;
  }
  static uint8_t clamp8(Wide v) { return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v)); }
  static void idct(uint8_t* out, int stride, const int16_t* d) {
    Wide val[64];
    for (int i = 0; i < 8; ++i) {  // columns; two extra bits are kept
      if (d[i + 8] == 0 && d[i + 16] == 0 && d[i + 24] == 0 && d[i + 32] == 0 && d[i + 40] == 0 && d[i + 48] == 0 && d[i + 56] == 0) {
        const Wide dc_term = Wide(d[i]) * 4;
        for (int r = 0; r < 8; ++r) val[i + r * 8] = dc_term;
        continue;
      }
      Wide x0, x1, x2, x3, t0, t1, t2, t3;
      idct_1d(d[i], d[i + 8], d[i + 16], d[i + 24], d[i + 32], d[i + 40], d[i + 48], d[i + 56], x0, x1, x2, x3, t0, t1, t2, t3);
      x0 += 512, x1 += 512, x2 += 512, x3 += 512;
      val[i] = (x0 + t3) >> 10, val[i + 56] = (x0 - t3) >> 10;
      val[i + 8] = (x1 + t2) >> 10, val[i + 48] = (x1 - t2) >> 10;
      val[i + 16] = (x2 + t1) >> 10, val[i + 40] = (x2 - t1) >> 10;
      val[i + 24] = (x3 + t0) >> 10, val[i + 32] = (x3 - t0) >> 10;
    }
    for (int i = 0; i < 8; ++i) {  // rows; 17 bits to drop, + 128 for the level shift
      const Wide* v = val + i * 8;
      uint8_t* o = out + size_t(i) * stride;
      Wide x0, x1, x2, x3, t0, t1, t2, t3;
      idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], x0, x1, x2, x3, t0, t1, t2, t3);
      const Wide bias = 65536 + (128 << 17);
      x0 += bias, x1 += bias, x2 += bias, x3 += bias;
      o[0] = clamp8((x0 + t3) >> 17), o[7] = clamp8((x0 - t3) >> 17);
      o[1] = clamp8((x1 + t2) >> 17), o[6] = clamp8((x1 - t2) >> 17);
      o[2] = clamp8((x2 + t1) >> 17), o[5] = clamp8((x2 - t1) >> 17);
      o[3] = clamp8((x3 + t0) >> 17), o[4] = clamp8((x3 - t0) >> 17);
    }
  }

  void scan() {
    reset_entropy();
    int16_t block[64];
    auto one_block = [&](Component& c, int bx, int by) {
      if (!progressive) {
        block_baseline(block, c);
        idct(c.data.data() + size_t(c.w2) * by * 8 + size_t(bx) * 8, c.w2, block);
      } else {
        int16_t* data = c.coeff.data() + 64 * (size_t(bx) + size_t(by) * c.coeff_w);
        if (spec_start == 0) {
          block_progressive_dc(data, c);
        } else {
          block_progressive_ac(data, c);
        }
      }
    };
    if (scan_n == 1) {
      Component& c = comp[order[0]];
      const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
      for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
          one_block(c, i, j);
          if (!unit_done()) return;
        }
      return;
    }
    if (progressive && spec_start != 0) bad("bad progressive JPEG scan");
    for (int j = 0; j < mcu_y; ++j)
      for (int i = 0; i < mcu_x; ++i) {
        for (int k = 0; k < scan_n; ++k) {
          Component& c = comp[order[k]];
          for (int y = 0; y < c.v; ++y)
            for (int x = 0; x < c.h; ++x) one_block(c, i * c.h + x, j * c.v + y);
        }
        if (!unit_done()) return;
      }
  }

  void frame_header(int kind) {
    progressive = (kind == 0xc2);
    const int length = get16();
    if (get8() != 8) bad("only 8-bit JPEG files are read");
    height = get16(), width = get16();
    components = get8();
    if (height == 0 || width == 0) bad("bad JPEG size");
    if (components != 1 && components != 3) bad("JPEG files with 1 or 3 components are read");
    if (length != 8 + 3 * components) bad("bad JPEG frame header");
    static const char rgb[3] = {'R', 'G', 'B'};
    int matches = 0;
    for (int i = 0; i < components; ++i) {
      Component& c = comp[i];
      c.id = get8();
      if (components == 3 && c.id == rgb[i]) ++matches;
      const int q = get8();
      c.h = q >> 4, c.v = q & 15;
      if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4) bad("bad JPEG sampling factors");
      c.tq = get8();
      if (c.tq > 3) bad("bad JPEG quantisation table index");
      h_max = std::max(h_max, c.h), v_max = std::max(v_max, c.v);
    }
    rgb_ids = (matches == 3);
    for (int i = 0; i < components; ++i)
      if (h_max % comp[i].h != 0 || v_max % comp[i].v != 0) bad("bad JPEG sampling factors");
    if (uint64_t(width) * height > uint64_t(size) * 4096u) bad("JPEG header does not match the file size");
    const int mcu_w = h_max * 8, mcu_h = v_max * 8;
    mcu_x = (width + mcu_w - 1) / mcu_w, mcu_y = (height + mcu_h - 1) / mcu_h;
    for (int i = 0; i < components; ++i) {
      Component& c = comp[i];
      c.x = (width * c.h + h_max - 1) / h_max, c.y = (height * c.v + v_max - 1) / v_max;
      c.w2 = mcu_x * c.h * 8, c.h2 = mcu_y * c.v * 8;
      c.data.assign(size_t(c.w2) * c.h2, 0);
      if (progressive) {
        c.coeff_w = c.w2 / 8;
        c.coeff.assign(size_t(c.w2) * c.h2, 0);
      }
    }
  }

  void tables(int m) {
    if (m == 0xdd) {
      if (get16() != 4) bad("bad JPEG restart interval");
      restart_interval = get16();
    } else if (m == 0xdb) {
      int left = get16() - 2;
      while (left > 0) {
        const int q = get8(), sixteen = (q >> 4), t = q & 15;
        if (sixteen > 1 || t > 3) bad("bad JPEG quantisation table");
        for (int i = 0; i < 64; ++i) dequant[t][kZigzag[i]] = uint16_t(sixteen ? get16() : get8());
        left -= sixteen ? 129 : 65;
      }
      if (left != 0) bad("bad JPEG quantisation table");
    } else if (m == 0xc4) {
      int left = get16() - 2;
      while (left > 0) {
        const int q = get8(), tc = q >> 4, th = q & 15;
        if (tc > 1 || th > 3) bad("bad JPEG Huffman table");
        int counts[16], n = 0;
        for (int& c : counts) n += (c = get8());
        if (n > 256) bad("bad JPEG Huffman table");
        Huffman& h = tc ? ac[th] : dc[th];
        if (!h.build(counts)) bad("bad JPEG Huffman table");
        for (int i = 0; i < n; ++i) h.values[i] = uint8_t(get8());
        left -= 17 + n;
      }
      if (left != 0) bad("bad JPEG Huffman table");
    } else if ((m >= 0xe0 && m <= 0xef) || m == 0xfe) {
      int left = get16();
      if (left < 2) bad("bad JPEG segment length");
      left -= 2;
      const size_t next = pos + size_t(left);
      if (m == 0xe0 && left >= 5 && pos + 5 <= size && memcmp(b + pos, "JFIF\0", 5) == 0) jfif = true;
      if (m == 0xee && left >= 12 && pos + 12 <= size && memcmp(b + pos, "Adobe\0", 6) == 0) adobe_transform = b[pos + 11];
      pos = std::min(next, size);
    } else {
      bad("unknown JPEG marker");
    }
  }

  int next_marker() {
    if (marker >= 0) {
      int m = marker;
      marker = -1;
      return m;
    }
    while (pos < size && b[pos] != 0xff) ++pos;  // whatever follows a scan up to the next marker
    while (pos < size && b[pos] == 0xff) ++pos;
    return pos < size ? b[pos++] : -1;
  }

  void decode_image() {
    if (size < 4 || b[0] != 0xff || b[1] != 0xd8) bad("not a JPEG file");
    pos = 2;
    bool have_frame = false;
    for (;;) {
      int m = next_marker();
      if (m < 0) bad("truncated JPEG file");
      if (m == 0xd9) break;
      if (m == 0xc0 || m == 0xc1 || m == 0xc2) {
        if (have_frame) bad("more than one JPEG frame");
        frame_header(m);
        have_frame = true;
      } else if (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
        bad("this JPEG coding process is not read (baseline and progressive Huffman are)");
      } else if (m == 0xda) {
        if (!have_frame) bad("JPEG scan before the frame header");
        const int length = get16();
        scan_n = get8();
        if (scan_n < 1 || scan_n > components || length != 6 + 2 * scan_n) bad("bad JPEG scan header");
        for (int i = 0; i < scan_n; ++i) {
          const int id = get8(), q = get8();
          int which = 0;
          while (which < components && comp[which].id != id) ++which;
          if (which == components) bad("bad JPEG scan header");
          comp[which].hd = q >> 4, comp[which].ha = q & 15;
          if (comp[which].hd > 3 || comp[which].ha > 3) bad("bad JPEG scan header");
          order[i] = which;
        }
        spec_start = get8(), spec_end = get8();
        const int aa = get8();
        succ_high = aa >> 4, succ_low = aa & 15;
        if (progressive) {
          if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) bad("bad JPEG scan header");
        } else {
          if (spec_start != 0 || succ_high != 0 || succ_low != 0) bad("bad JPEG scan header");
          spec_end = 63;
        }
        scan();
      } else if (m == 0xdc) {  // DNL
        const int length = get16(), lines = get16();
        if (length != 4 || lines != height) bad("bad JPEG DNL segment");
      } else if (m >= 0xd0 && m <= 0xd7) {
        continue;  // a stray restart marker
      } else {
        tables(m);
      }
    }
    if (!have_frame) bad("JPEG file without a frame");
    if (progressive) {
      for (int n = 0; n < components; ++n) {
        Component& c = comp[n];
        const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
        for (int j = 0; j < h; ++j)
          for (int i = 0; i < w; ++i) {
            int16_t* data = c.coeff.data() + 64 * (size_t(i) + size_t(j) * c.coeff_w);
            for (int k = 0; k < 64; ++k) data[k] = int16_t(data[k] * dequant[c.tq][k]);
            idct(c.data.data() + size_t(c.w2) * j * 8 + size_t(i) * 8, c.w2, data);
          }
      }
    }
  }

  // chroma upsampling of one output row: `near` is the closer input row, `far` the other one
  static const uint8_t* upsample(std::vector<uint8_t>& line, const uint8_t* near, const uint8_t* far, int w, int hs, int vs) {
    uint8_t* out = line.data();
    if (hs == 1 && vs == 1) return near;
    if (hs == 1 && vs == 2) {
      for (int i = 0; i < w; ++i) out[i] = uint8_t((3 * near[i] + far[i] + 2) >> 2);
    } else if (hs == 2 && vs == 1) {
      if (w == 1) {
        out[0] = out[1] = near[0];
        return out;
      }
      out[0] = near[0];
      out[1] = uint8_t((near[0] * 3 + near[1] + 2) >> 2);
      int i = 1;
      for (; i < w - 1; ++i) {
        const int n = 3 * near[i] + 2;
        out[i * 2 + 0] = uint8_t((n + near[i - 1]) >> 2);
        out[i * 2 + 1] = uint8_t((n + near[i + 1]) >> 2);
      }
      out[i * 2 + 0] = uint8_t((near[w - 2] * 3 + near[w - 1] + 2) >> 2);
      out[i * 2 + 1] = near[w - 1];
    } else if (hs == 2 && vs == 2) {
      if (w == 1) {
        out[0] = out[1] = uint8_t((3 * near[0] + far[0] + 2) >> 2);
        return out;
      }
      int t1 = 3 * near[0] + far[0];
      out[0] = uint8_t((t1 + 2) >> 2);
      for (int i = 1; i < w; ++i) {
        const int t0 = t1;
        t1 = 3 * near[i] + far[i];
        out[i * 2 - 1] = uint8_t((3 * t0 + t1 + 8) >> 4);
        out[i * 2] = uint8_t((3 * t1 + t0 + 8) >> 4);
      }
      out[w * 2 - 1] = uint8_t((t1 + 2) >> 2);
    } else {
      for (int i = 0; i < w; ++i)
        for (int j = 0; j < hs; ++j) out[i * hs + j] = near[i];
    }
    return out;
  }

  Pixels to_rgba() {
    Pixels out;
    out.w = uint32_t(width), out.h = uint32_t(height), out.eight_bit = true;
    out.u8.assign(size_t(width) * height * 4, 255);
    const bool is_rgb = components == 3 && (rgb_ids || (adobe_transform == 0 && !jfif));
    struct Row {
      int hs, vs, ystep, w_lores, ypos;
      const uint8_t *line0, *line1;
      std::vector<uint8_t> buffer;
    } rows[3];
    for (int k = 0; k < components; ++k) {
      Row& r = rows[k];
      r.hs = h_max / comp[k].h, r.vs = v_max / comp[k].v;
      r.ystep = r.vs >> 1, r.w_lores = (width + r.hs - 1) / r.hs, r.ypos = 0;
      r.line0 = r.line1 = comp[k].data.data();
      r.buffer.assign(size_t(width) + 8, 0);
    }
    auto fixed = [](float v) { return (int(v * 4096.0f + 0.5f)) << 8; };
    for (int j = 0; j < height; ++j) {
      const uint8_t* plane[3] = {nullptr, nullptr, nullptr};
      for (int k = 0; k < components; ++k) {
        Row& r = rows[k];
        const bool bottom = r.ystep >= (r.vs >> 1);
        plane[k] = upsample(r.buffer, bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.w_lores, r.hs, r.vs);
        if (++r.ystep >= r.vs) {
          r.ystep = 0;
          r.line0 = r.line1;
          if (++r.ypos < comp[k].y) r.line1 += comp[k].w2;
        }
      }
      uint8_t* o = out.u8.data() + size_t(j) * width * 4;
      for (int i = 0; i < width; ++i, o += 4) {
        if (components == 1) {
          o[0] = o[1] = o[2] = plane[0][i];
        } else if (is_rgb) {
          o[0] = plane[0][i], o[1] = plane[1][i], o[2] = plane[2][i];
        } else {
          const int y_fixed = (plane[0][i] << 20) + (1 << 19);
          const int cr = plane[2][i] - 128, cb = plane[1][i] - 128;
          int r = y_fixed + cr * fixed(1.40200f);
          int g = y_fixed + (cr * -fixed(0.71414f)) + ((cb * -fixed(0.34414f)) & int(0xffff0000u));
          int bl = y_fixed + cb * fixed(1.77200f);
          r >>= 20, g >>= 20, bl >>= 20;
          o[0] = clamp8(r), o[1] = clamp8(g), o[2] = clamp8(bl);
        }
      }
    }
    return out;
  }
};

}  // namespace jpeg

Pixels read_jpeg(const std::string& path) {
  const std::string d = read_file(path);
  auto decoder = std::make_unique<jpeg::Decoder>();
  decoder->b = reinterpret_cast<const uint8_t*>(d.data());
  decoder->size = d.size();
  decoder->path = &path;
  decoder->decode_image();
  return decoder->to_rgba();
}
