// scene_loader_formats.inl — the file formats the scene loader reads (included by scene_loader.cpp inside its anonymous namespace): DEFLATE, PNG, OpenEXR
// (scan lines, none / ZIPS / ZIP), Radiance HDR, the reference's PFM variant, JSON, the patched tinyobjloader's .mtl / .obj dialect.

// ---- DEFLATE (RFC 1951) + the zlib wrapper (RFC 1950) --------------------------------------------------------------------------------------------------
struct BitReader {
  const uint8_t* data;
  size_t size, pos = 0;
  uint32_t bit_buffer = 0, bit_count = 0;
  uint32_t bits(uint32_t n) {
    while (bit_count < n) {
      if (pos >= size) fail("truncated deflate stream");
      bit_buffer |= uint32_t(data[pos++]) << bit_count;
      bit_count += 8;
    }
    uint32_t v = bit_buffer & ((n == 32u) ? 0xffffffffu : ((1u << n) - 1u));
    bit_buffer = (n == 32u) ? 0u : (bit_buffer >> n);
    bit_count -= n;
    return v;
  }
};
struct Huffman {
  uint16_t count[16] = {}, symbol[320] = {};
  void build(const uint8_t* lengths, int n) {
    memset(count, 0, sizeof(count));
    for (int i = 0; i < n; ++i) count[lengths[i]]++;
    count[0] = 0;
    uint16_t offs[16] = {};
    for (int i = 1; i < 16; ++i) offs[i] = uint16_t(offs[i - 1] + count[i - 1]);
    for (int i = 0; i < n; ++i)
      if (lengths[i]) symbol[offs[lengths[i]]++] = uint16_t(i);
  }
  int decode(BitReader& br) const {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len < 16; ++len) {
      code |= int(br.bits(1));
      int c = count[len];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    fail("bad deflate code");
  }
};
std::vector<uint8_t> inflate_raw(const uint8_t* data, size_t size) {
  static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint16_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint16_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  BitReader br{data, size};
  std::vector<uint8_t> out;
  for (;;) {
    uint32_t last = br.bits(1), type = br.bits(2);
    if (type == 0u) {
      br.bit_buffer = 0, br.bit_count = 0;
      if (br.pos + 4 > size) fail("truncated stored block");
      uint32_t len = data[br.pos] | (uint32_t(data[br.pos + 1]) << 8);
      br.pos += 4;
      if (br.pos + len > size) fail("truncated stored block");
      out.insert(out.end(), data + br.pos, data + br.pos + len);
      br.pos += len;
    } else if (type == 1u || type == 2u) {
      Huffman lit, dist;
      uint8_t lengths[320] = {};
      if (type == 1u) {
        for (int i = 0; i < 144; ++i) lengths[i] = 8;
        for (int i = 144; i < 256; ++i) lengths[i] = 9;
        for (int i = 256; i < 280; ++i) lengths[i] = 7;
        for (int i = 280; i < 288; ++i) lengths[i] = 8;
        lit.build(lengths, 288);
        for (int i = 0; i < 30; ++i) lengths[i] = 5;
        dist.build(lengths, 30);
      } else {
        int nlen = int(br.bits(5)) + 257, ndist = int(br.bits(5)) + 1, ncode = int(br.bits(4)) + 4;
        uint8_t cl[19] = {};
        for (int i = 0; i < ncode; ++i) cl[order[i]] = uint8_t(br.bits(3));
        Huffman code;
        code.build(cl, 19);
        int i = 0;
        while (i < nlen + ndist) {
          int sym = code.decode(br);
          if (sym < 16) {
            lengths[i++] = uint8_t(sym);
          } else {
            uint8_t prev = 0;
            int rep = 0;
            if (sym == 16) {
              if (i == 0) fail("bad deflate lengths");
              prev = lengths[i - 1];
              rep = 3 + int(br.bits(2));
            } else if (sym == 17) {
              rep = 3 + int(br.bits(3));
            } else {
              rep = 11 + int(br.bits(7));
            }
            if (i + rep > nlen + ndist) fail("bad deflate lengths");
            while (rep--) lengths[i++] = prev;
          }
        }
        lit.build(lengths, nlen);
        dist.build(lengths + nlen, ndist);
      }
      for (;;) {
        int sym = lit.decode(br);
        if (sym < 256) {
          out.push_back(uint8_t(sym));
        } else if (sym == 256) {
          break;
        } else {
          sym -= 257;
          if (sym >= 29) fail("bad deflate length symbol");
          size_t len = len_base[sym] + br.bits(len_extra[sym]);
          int ds = dist.decode(br);
          if (ds >= 30) fail("bad deflate distance symbol");
          size_t d = dist_base[ds] + br.bits(dist_extra[ds]);
          if (d > out.size()) fail("bad deflate distance");
          size_t from = out.size() - d;
          for (size_t k = 0; k < len; ++k) out.push_back(out[from + k]);
        }
      }
    } else {
      fail("bad deflate block type");
    }
    if (last) break;
  }
  return out;
}
std::vector<uint8_t> inflate_zlib(const uint8_t* data, size_t size) {
  if (size < 6) fail("truncated zlib stream");
  return inflate_raw(data + 2, size - 2);
}

// ---- images -----------------------------------------------------------------------------------------------------------------------------------------------
struct Pixels {
  uint32_t w = 0, h = 0;
  bool eight_bit = false;
  std::vector<uint8_t> u8;  // RGBA8, rows in file order
  std::vector<float> f32;   // RGBA32F
};
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

// Every PNG form the reference's reader (stb_image) decodes: bit depths 1 / 2 / 4 / 8 / 16, the five colour types, Adam7 interlacing, tRNS colour keys
// and palette alpha — reduced to 8 bits the way stb_image does it (16-bit samples keep their high byte, low-depth grey is scaled to 0..255), then
// widened to RGBA8 the way the reference's texture pool does it (image_pool.cxx:353-381: 4 channels kept, 3 get alpha 255, 1 is replicated, 2 — grey with
// alpha, which includes grey with a colour key — has no case there and stays zero-filled).
Pixels read_png(const std::string& path) {
  std::string d = read_file(path);
  const uint8_t* b = reinterpret_cast<const uint8_t*>(d.data());
  if (d.size() < 8 || memcmp(b, "\x89PNG\r\n\x1a\n", 8) != 0) fail(path + ": not a PNG file");
  size_t pos = 8;
  std::vector<uint8_t> idat, palette, trns;
  uint32_t w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
  while (pos + 12 <= d.size()) {
    uint32_t n = be32(b + pos);
    const uint8_t* type = b + pos + 4;
    const uint8_t* body = b + pos + 8;
    if (n > d.size() || pos + 12 + n > d.size()) break;
    if (!memcmp(type, "IHDR", 4) && n >= 13) {
      w = be32(body), h = be32(body + 4), depth = body[8], ctype = body[9], interlace = body[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      palette.assign(body, body + n);
    } else if (!memcmp(type, "tRNS", 4)) {
      trns.assign(body, body + n);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + n);
    }
    pos += 12 + n;
  }
  const uint32_t ch = (ctype == 0) ? 1 : (ctype == 2) ? 3 : (ctype == 3) ? 1 : (ctype == 4) ? 2 : (ctype == 6) ? 4 : 0;
  const bool depth_ok = (depth == 8) || (depth == 16 && ctype != 3) || ((depth == 1 || depth == 2 || depth == 4) && (ctype == 0 || ctype == 3));
  if (ch == 0 || !depth_ok || interlace > 1 || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) fail(path + ": unsupported PNG header");
  std::vector<uint8_t> raw = inflate_zlib(idat.data(), idat.size());
  if (uint64_t(w) * h * ch * (depth == 16 ? 2 : 1) > uint64_t(raw.size()) * 8u + 64u) fail(path + ": truncated PNG data");
  // samples of the whole image, 16 bits wide (low depths unscaled for now)
  std::vector<uint16_t> img(size_t(w) * h * ch, 0);
  static const uint32_t x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
  const uint32_t passes = interlace ? 7u : 1u;
  const size_t bpp = std::max<size_t>(1, size_t(ch) * depth / 8);  // the filters' pixel distance in bytes
  size_t at = 0;
  std::vector<uint8_t> cur, prev;
  for (uint32_t pass = 0; pass < passes; ++pass) {
    const uint32_t ox = interlace ? x0[pass] : 0u, oy = interlace ? y0[pass] : 0u, sx = interlace ? dx[pass] : 1u, sy = interlace ? dy[pass] : 1u;
    const uint32_t pw = (w > ox) ? (w - ox + sx - 1) / sx : 0u, ph = (h > oy) ? (h - oy + sy - 1) / sy : 0u;
    if (pw == 0 || ph == 0) continue;
    const size_t stride = (size_t(pw) * ch * depth + 7) / 8;
    prev.assign(stride, 0);
    cur.assign(stride, 0);
    for (uint32_t y = 0; y < ph; ++y) {
      if (at + 1 + stride > raw.size()) fail(path + ": truncated PNG data");
      const uint8_t ft = raw[at];
      const uint8_t* src = raw.data() + at + 1;
      at += 1 + stride;
      if (ft > 4) fail(path + ": bad PNG filter");
      for (size_t x = 0; x < stride; ++x) {
        const int a = (x >= bpp) ? cur[x - bpp] : 0, up = prev[x], c = (x >= bpp) ? prev[x - bpp] : 0;
        int p = 0;
        switch (ft) {
          case 1: p = a; break;
          case 2: p = up; break;
          case 3: p = (a + up) >> 1; break;
          case 4: {
            const int pa = abs(up - c), pb = abs(a - c), pc = abs(a + up - 2 * c);
            p = (pa <= pb && pa <= pc) ? a : (pb <= pc ? up : c);
            break;
          }
          default: break;
        }
        cur[x] = uint8_t(src[x] + p);
      }
      const uint32_t iy = oy + y * sy;
      for (uint32_t x = 0; x < pw; ++x) {
        uint16_t* o = img.data() + (size_t(iy) * w + (ox + size_t(x) * sx)) * ch;
        for (uint32_t k = 0; k < ch; ++k) {
          const size_t sample = size_t(x) * ch + k;
          if (depth == 16) {
            o[k] = uint16_t((uint16_t(cur[sample * 2]) << 8) | cur[sample * 2 + 1]);
          } else if (depth == 8) {
            o[k] = cur[sample];
          } else {
            const size_t bit = sample * depth;
            o[k] = uint16_t((cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u));
          }
        }
      }
      prev.swap(cur);
    }
  }
  // 8-bit samples, palette and colour key
  static const uint32_t depth_scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
  uint32_t out_n = ch;
  std::vector<uint8_t> px;
  const size_t count = size_t(w) * h;
  if (ctype == 3) {
    out_n = trns.empty() ? 3u : 4u;
    px.assign(count * out_n, 0);
    for (size_t i = 0; i < count; ++i) {
      const size_t idx = img[i];
      uint8_t rgba[4] = {0, 0, 0, 255};
      if (idx * 3 + 2 < palette.size()) memcpy(rgba, palette.data() + idx * 3, 3);
      if (idx < trns.size()) rgba[3] = trns[idx];
      memcpy(px.data() + i * out_n, rgba, out_n);
    }
  } else {
    const bool keyed = !trns.empty() && (ctype == 0 || ctype == 2) && trns.size() >= size_t(ch) * 2;
    out_n = ch + (keyed ? 1u : 0u);
    px.assign(count * out_n, 0);
    uint16_t key[3] = {0, 0, 0};
    for (uint32_t k = 0; keyed && k < ch; ++k) {
      const uint16_t v = uint16_t((uint16_t(trns[k * 2]) << 8) | trns[k * 2 + 1]);
      key[k] = (depth == 16) ? v : uint16_t((v & 255u) * depth_scale[depth]);
    }
    for (size_t i = 0; i < count; ++i) {
      bool match = keyed;
      for (uint32_t k = 0; k < ch; ++k) {
        const uint16_t v = (depth == 16) ? img[i * ch + k] : uint16_t(img[i * ch + k] * depth_scale[depth]);
        if (keyed && v != key[k]) match = false;
        px[i * out_n + k] = (depth == 16) ? uint8_t(v >> 8) : uint8_t(v);
      }
      if (keyed) px[i * out_n + ch] = match ? 0 : 255;
    }
  }
  Pixels out;
  out.w = w, out.h = h, out.eight_bit = true;
  out.u8.assign(count * 4, 255);
  for (size_t i = 0; i < count; ++i) {
    uint8_t* o = out.u8.data() + i * 4;
    const uint8_t* sp = px.data() + i * out_n;
    if (out_n == 4) {
      memcpy(o, sp, 4);
    } else if (out_n == 3) {
      memcpy(o, sp, 3);
    } else if (out_n == 1) {
      o[0] = o[1] = o[2] = sp[0];
    } else {
      o[0] = o[1] = o[2] = o[3] = 0;  // two channels: no case in the reference's switch over the channel count
    }
  }
  return out;
}

float half_to_float(uint16_t h) {
  uint32_t sign = (h >> 15) & 1u, e = (h >> 10) & 31u, m = h & 1023u, bits;
  if (e == 0) {
    if (m == 0) {
      bits = sign << 31;
    } else {
      e = 127 - 15 + 1;
      while (!(m & 1024u)) {
        m <<= 1;
        e--;
      }
      bits = (sign << 31) | (e << 23) | ((m & 1023u) << 13);
    }
  } else if (e == 31) {
    bits = (sign << 31) | 0x7f800000u | (m << 13);
  } else {
    bits = (sign << 31) | ((e + 127 - 15) << 23) | (m << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

// One PIZ-compressed block of an OpenEXR file (ImfPizCompressor, as tinyexr's DecompressPiz :9508-9620 reads it): a bitmap of the 16-bit values that
// occur -> a lookup table; the table indices, wavelet-transformed per channel plane (14- or 16-bit lifting, ImfWav), Huffman-coded with run lengths
// (ImfHuf: 6-bit packed code lengths, canonical codes, the last symbol = "repeat the previous word n times").  Output: the block's scan lines, every
// channel's row in turn, native 16-bit words.
namespace piz {

inline void lift14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
  const int hi = int16_t(h), ai = int16_t(l) + (hi & 1) + (hi >> 1);
  a = uint16_t(int16_t(ai)), b = uint16_t(int16_t(ai - hi));
}
inline void lift16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
  const int m = l, d = h, bb = (m - (d >> 1)) & 0xffff, aa = (d + bb - 0x8000) & 0xffff;
  b = uint16_t(bb), a = uint16_t(aa);
}

void wavelet_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t max_value) {
  const bool narrow = max_value < (1 << 14);
  const int n = std::min(nx, ny);
  int p = 1;
  while (p <= n) p <<= 1;
  p >>= 1;
  int p2 = p;
  p >>= 1;
  auto lift = [&](uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) { narrow ? lift14(l, h, a, b) : lift16(l, h, a, b); };
  while (p >= 1) {
    uint16_t* py = in;
    uint16_t* const ey = in + ptrdiff_t(oy) * (ny - p2);
    const int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
    uint16_t i00, i01, i10, i11;
    for (; py <= ey; py += oy2) {
      uint16_t* px = py;
      uint16_t* const ex = py + ptrdiff_t(ox) * (nx - p2);
      for (; px <= ex; px += ox2) {
        uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
        lift(*px, *p10, i00, i10);
        lift(*p01, *p11, i01, i11);
        lift(i00, i01, *px, *p01);
        lift(i10, i11, *p10, *p11);
      }
      if (nx & p) {
        uint16_t* p10 = px + oy1;
        lift(*px, *p10, i00, *p10);
        *px = i00;
      }
    }
    if (ny & p) {
      uint16_t* px = py;
      uint16_t* const ex = py + ptrdiff_t(ox) * (nx - p2);
      for (; px <= ex; px += ox2) {
        uint16_t* p01 = px + ox1;
        lift(*px, *p01, i00, *p01);
        *px = i00;
      }
    }
    p2 = p;
    p >>= 1;
  }
}

struct Bits {
  const uint8_t* in;
  size_t size, pos = 0;
  uint64_t c = 0;
  int lc = 0;
  uint32_t get(int n) {
    while (lc < n) {
      c = (c << 8) | (pos < size ? in[pos] : 0u);
      ++pos;
      lc += 8;
    }
    lc -= n;
    return uint32_t((c >> lc) & ((1ull << n) - 1ull));
  }
};

bool huffman_decode(const uint8_t* in, size_t size, std::vector<uint16_t>& out, size_t expected) {
  constexpr int kSymbols = (1 << 16) + 1;
  if (size < 20) return false;
  auto u32 = [&](size_t at) { return uint32_t(in[at]) | (uint32_t(in[at + 1]) << 8) | (uint32_t(in[at + 2]) << 16) | (uint32_t(in[at + 3]) << 24); };
  const int im = int(u32(0)), iM = int(u32(4));
  const uint64_t n_bits = u32(12);
  if (im < 0 || im >= kSymbols || iM < 0 || iM >= kSymbols || im > iM) return false;
  std::vector<uint8_t> length(kSymbols, 0);
  Bits table{in + 20, size - 20};
  for (int i = im; i <= iM; ++i) {  // packed code lengths: 0..58, 59..62 = 2..5 zeros, 63 = 6 + (8 bits) zeros
    if (table.pos > table.size) return false;
    const uint32_t l = table.get(6);
    if (l == 63u || l >= 59u) {
      int zeros = (l == 63u) ? int(table.get(8)) + 6 : int(l) - 59 + 2;
      if (i + zeros > iM + 1) return false;
      i += zeros - 1;
    } else {
      length[i] = uint8_t(l);
    }
  }
  const size_t data_at = 20 + table.pos;
  if (data_at > size || n_bits > 8ull * (size - data_at)) return false;
  // canonical codes: counted per length, the longest lengths take the smallest code values, symbols of one length in increasing order
  uint64_t count[59] = {}, base[59] = {};
  for (int i = 0; i < kSymbols; ++i) count[length[i]] += 1;
  {
    uint64_t c = 0;
    for (int l = 58; l > 0; --l) {
      const uint64_t next = (c + count[l]) >> 1;
      base[l] = c;
      c = next;
    }
  }
  std::vector<uint32_t> start(60, 0), order;
  for (int l = 1; l <= 58; ++l) start[l + 1] = start[l] + uint32_t(count[l]);
  order.assign(start[59], 0);
  {
    std::vector<uint32_t> fill(start.begin(), start.end());
    for (int i = 0; i < kSymbols; ++i)
      if (length[i]) order[fill[length[i]]++] = uint32_t(i);
  }
  out.clear();
  out.reserve(expected);
  Bits data{in + data_at, size - data_at};
  uint64_t used = 0;
  while (used < n_bits) {
    uint64_t code = 0;
    int l = 0;
    uint32_t symbol = 0;
    bool found = false;
    while (l < 58 && used < n_bits) {
      code = (code << 1) | data.get(1);
      ++l, ++used;
      if (count[l] && code >= base[l] && code < base[l] + count[l]) {
        symbol = order[start[l] + uint32_t(code - base[l])];
        found = true;
        break;
      }
    }
    if (!found) return used >= n_bits && out.size() == expected;  // trailing pad bits
    if (int(symbol) == iM) {
      if (used + 8 > n_bits || out.empty()) return false;
      const uint32_t repeat = data.get(8);
      used += 8;
      if (out.size() + repeat > expected) return false;
      out.insert(out.end(), repeat, out.back());
    } else {
      if (out.size() >= expected) return false;
      out.push_back(uint16_t(symbol));
    }
  }
  return out.size() == expected;
}

// `words_per_pixel[c]` = 1 for a half channel, 2 for float / uint
bool decode_block(const uint8_t* in, size_t size, const std::vector<int>& words_per_pixel, int width, int lines, std::vector<uint8_t>& raw) {
  size_t total_words = 0;
  for (int wpp : words_per_pixel) total_words += size_t(width) * lines * wpp;
  if (size < 4) return false;
  const uint32_t min_nz = in[0] | (in[1] << 8), max_nz = in[2] | (in[3] << 8);
  size_t at = 4;
  std::vector<uint8_t> bitmap(8192, 0);
  if (max_nz >= 8192) return false;
  if (min_nz <= max_nz) {
    if (at + (max_nz - min_nz + 1) > size) return false;
    memcpy(bitmap.data() + min_nz, in + at, max_nz - min_nz + 1);
    at += max_nz - min_nz + 1;
  }
  std::vector<uint16_t> lut(65536, 0);
  int k = 0;
  for (int i = 0; i < 65536; ++i)
    if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = uint16_t(i);
  const uint16_t max_value = uint16_t(k - 1);
  if (at + 4 > size) return false;
  const uint32_t length = uint32_t(in[at]) | (uint32_t(in[at + 1]) << 8) | (uint32_t(in[at + 2]) << 16) | (uint32_t(in[at + 3]) << 24);
  at += 4;
  if (at + length > size) return false;
  std::vector<uint16_t> words;
  if (!huffman_decode(in + at, length, words, total_words)) return false;
  size_t plane = 0;
  std::vector<size_t> begin;
  for (int wpp : words_per_pixel) {
    begin.push_back(plane);
    for (int j = 0; j < wpp; ++j) wavelet_decode(words.data() + plane + j, width, wpp, lines, width * wpp, max_value);
    plane += size_t(width) * lines * wpp;
  }
  for (uint16_t& w : words) w = lut[w];
  raw.resize(total_words * 2);
  size_t o = 0;
  for (int y = 0; y < lines; ++y)
    for (size_t c = 0; c < words_per_pixel.size(); ++c) {
      const size_t n = size_t(width) * words_per_pixel[c];
      memcpy(raw.data() + o, words.data() + begin[c] + size_t(y) * n, n * 2);
      o += n * 2;
    }
  return true;
}

}  // namespace piz

Pixels read_exr(const std::string& path) {
  std::string d = read_file(path);
  const uint8_t* b = reinterpret_cast<const uint8_t*>(d.data());
  auto le32 = [&](size_t at) {
    uint32_t v;
    if (at + 4 > d.size()) fail(path + ": truncated EXR file");
    memcpy(&v, b + at, 4);
    return v;
  };
  if (d.size() < 8 || le32(0) != 20000630u) fail(path + ": not an OpenEXR file");
  size_t pos = 8;
  std::map<std::string, std::string> attrs;
  while (pos < d.size() && b[pos] != 0) {
    size_t ne = d.find('\0', pos), te = d.find('\0', ne + 1);
    if (ne == std::string::npos || te == std::string::npos) fail(path + ": bad EXR header");
    uint32_t size = le32(te + 1);
    attrs[d.substr(pos, ne - pos)] = d.substr(te + 5, size);
    pos = te + 5 + size;
  }
  pos += 1;
  if (!attrs.count("compression") || !attrs.count("dataWindow") || !attrs.count("channels")) fail(path + ": incomplete EXR header");
  if (attrs["compression"].empty() || attrs["dataWindow"].size() < 16) fail(path + ": incomplete EXR header");
  int comp = uint8_t(attrs["compression"][0]);
  if (comp < 0 || comp > 4) fail(path + ": this EXR compression is not read (none / RLE / ZIPS / ZIP / PIZ are)");
  int32_t win[4];
  memcpy(win, attrs["dataWindow"].data(), 16);
  const uint32_t w = uint32_t(win[2] - win[0] + 1), h = uint32_t(win[3] - win[1] + 1);
  std::vector<std::string> names;
  std::vector<int> types;
  const std::string& chl = attrs["channels"];
  for (size_t cp = 0; cp < chl.size() && chl[cp] != 0;) {
    size_t e = chl.find('\0', cp);
    if (e == std::string::npos || e + 17 > chl.size()) fail(path + ": bad EXR channel list");
    names.push_back(chl.substr(cp, e - cp));
    int32_t t;
    memcpy(&t, chl.data() + e + 1, 4);
    types.push_back(t);
    cp = e + 17;
  }
  const uint32_t lines = (comp == 3) ? 16u : ((comp == 4) ? 32u : 1u), blocks = (h + lines - 1) / lines;
  size_t bytes_per_pixel = 0;
  for (int t : types) bytes_per_pixel += (t == 1) ? 2 : 4;
  // a header is only believed as far as the file can back it: a deflate stream expands at most ~1032 : 1
  if (win[2] < win[0] || win[3] < win[1] || bytes_per_pixel == 0 || uint64_t(w) * h * bytes_per_pixel > uint64_t(d.size()) * 1032u) fail(path + ": EXR header does not match the file size");
  Pixels out;
  out.w = w, out.h = h;
  out.f32.assign(size_t(w) * h * 4, 0.0f);
  for (size_t i = 0; i < size_t(w) * h; ++i) out.f32[i * 4 + 3] = 1.0f;
  for (uint32_t k = 0; k < blocks; ++k) {
    uint64_t off;
    if (pos + size_t(k) * 8 + 8 > d.size()) fail(path + ": truncated EXR offset table");
    memcpy(&off, b + pos + size_t(k) * 8, 8);
    if (off > d.size() || d.size() - off < 8) fail(path + ": bad EXR block offset");
    int32_t by = int32_t(le32(size_t(off)));
    uint32_t size = le32(size_t(off) + 4);
    if (by < win[1] || by > win[3]) fail(path + ": EXR block outside the data window");
    const uint32_t nl = std::min<uint32_t>(lines, uint32_t(win[3] - by + 1));
    const size_t want = size_t(nl) * w * bytes_per_pixel;
    std::vector<uint8_t> raw;
    if (size > d.size() - size_t(off) - 8) fail(path + ": truncated EXR block");
    if (comp == 4 && size < want) {
      std::vector<int> words_per_pixel;
      for (int t : types) words_per_pixel.push_back(t == 1 ? 1 : 2);
      if (!piz::decode_block(b + off + 8, size, words_per_pixel, int(w), int(nl), raw)) fail(path + ": corrupt PIZ block in the EXR file");
    } else if (comp != 0 && size < want) {
      std::vector<uint8_t> z;
      if (comp == 1) {  // run lengths: a negative count copies that many bytes, any other repeats the next byte count + 1 times
        const uint8_t* in = b + off + 8;
        for (size_t i = 0; i < size;) {
          const int count = int8_t(in[i++]);
          if (count < 0) {
            if (i + size_t(-count) > size) fail(path + ": corrupt EXR run");
            z.insert(z.end(), in + i, in + i + size_t(-count));
            i += size_t(-count);
          } else {
            if (i >= size) fail(path + ": corrupt EXR run");
            z.insert(z.end(), size_t(count) + 1, in[i++]);
          }
          if (z.size() > want) fail(path + ": corrupt EXR run");
        }
      } else {
        z = inflate_zlib(b + off + 8, size);
      }
      for (size_t i = 1; i < z.size(); ++i) z[i] = uint8_t(z[i - 1] + z[i] - 128);  // predictor
      raw.resize(z.size());
      const size_t half = (z.size() + 1) / 2;
      for (size_t i = 0; i < z.size(); ++i) raw[i] = (i & 1) ? z[half + i / 2] : z[i / 2];
    } else {
      raw.assign(b + off + 8, b + off + 8 + size);
    }
    if (raw.size() < want) fail(path + ": short EXR block");
    size_t p = 0;
    for (uint32_t ly = 0; ly < nl; ++ly) {
      const uint32_t y = uint32_t(by - win[1]) + ly;
      for (size_t ci = 0; ci < names.size(); ++ci) {
        int slot = (names[ci] == "R") ? 0 : (names[ci] == "G") ? 1 : (names[ci] == "B") ? 2 : (names[ci] == "A") ? 3 : (names[ci] == "Y") ? 4 : -1;
        for (uint32_t x = 0; x < w; ++x) {
          float v;
          if (types[ci] == 1) {
            uint16_t hv;
            memcpy(&hv, raw.data() + p, 2);
            v = half_to_float(hv);
            p += 2;
          } else if (types[ci] == 2) {
            memcpy(&v, raw.data() + p, 4);
            p += 4;
          } else {
            uint32_t u;
            memcpy(&u, raw.data() + p, 4);
            v = float(u);
            p += 4;
          }
          float* px = out.f32.data() + (size_t(y) * w + x) * 4;
          if (slot >= 0 && slot < 4) px[slot] = v;
          if (slot == 4) px[0] = px[1] = px[2] = v;
        }
      }
    }
  }
  return out;
}

Pixels read_hdr(const std::string& path) {
  std::string d = read_file(path);
  const uint8_t* b = reinterpret_cast<const uint8_t*>(d.data());
  if (d.compare(0, 10, "#?RADIANCE") != 0 && d.compare(0, 6, "#?RGBE") != 0) fail(path + ": not a Radiance HDR file");
  size_t pos = d.find("\n\n");
  if (pos == std::string::npos) fail(path + ": bad HDR header");
  pos += 2;
  size_t e = d.find('\n', pos);
  auto tok = split(d.substr(pos, e - pos));
  if (tok.size() != 4 || tok[0] != "-Y" || tok[2] != "+X") fail(path + ": unsupported HDR orientation");
  const uint32_t h = uint32_t(atoi(tok[1].c_str())), w = uint32_t(atoi(tok[3].c_str()));
  pos = e + 1;
  if (e == std::string::npos || w == 0 || h == 0 || uint64_t(w) * h > uint64_t(d.size()) * 64u) fail(path + ": HDR header does not match the file size");  // a run covers <= 127 texels of a plane in 2 bytes
  std::vector<uint8_t> rgbe(size_t(w) * h * 4);
  const bool rle = (w >= 8 && w < 32768 && pos + 4 <= d.size() && b[pos] == 2 && b[pos + 1] == 2 && (b[pos + 2] & 0x80) == 0);  // decided at the first scan line, like stb_image
  if (!rle) {
    if (pos + rgbe.size() > d.size()) fail(path + ": truncated HDR data");
    memcpy(rgbe.data(), b + pos, rgbe.size());
  } else {
    for (uint32_t y = 0; y < h; ++y) {
      if (pos + 4 > d.size() || b[pos] != 2 || b[pos + 1] != 2 || ((uint32_t(b[pos + 2]) << 8) | b[pos + 3]) != w) fail(path + ": corrupt HDR scan line");
      pos += 4;
      for (uint32_t c = 0; c < 4; ++c) {
        uint32_t x = 0;
        while (x < w) {
          if (pos >= d.size()) fail(path + ": truncated HDR data");
          uint32_t n = b[pos++];
          if (n > 128) {
            n -= 128;
            if (x + n > w || pos >= d.size()) fail(path + ": corrupt HDR run");
            for (uint32_t k = 0; k < n; ++k) rgbe[(size_t(y) * w + x + k) * 4 + c] = b[pos];
            pos += 1;
          } else {
            if (x + n > w || pos + n > d.size()) fail(path + ": corrupt HDR run");
            for (uint32_t k = 0; k < n; ++k) rgbe[(size_t(y) * w + x + k) * 4 + c] = b[pos + k];
            pos += n;
          }
          x += n;
        }
      }
    }
  }
  Pixels out;
  out.w = w, out.h = h;
  out.f32.resize(size_t(w) * h * 4);
  for (size_t i = 0; i < size_t(w) * h; ++i) {
    const uint8_t* p = rgbe.data() + i * 4;
    float* o = out.f32.data() + i * 4;
    if (p[3] != 0) {
      float f1 = ldexpf(1.0f, int(p[3]) - 136);
      o[0] = p[0] * f1, o[1] = p[1] * f1, o[2] = p[2] * f1;
    } else {
      o[0] = o[1] = o[2] = 0.0f;
    }
    o[3] = 1.0f;
  }
  return out;
}

// load_pfm (render/host/image_pool.cxx:463-541): the reference's own header variant — width, height and scale each on a line of its own
Pixels read_pfm(const std::string& path) {
  std::string d = read_file(path);
  size_t pos = 0;
  std::string lines[4];
  for (auto& line : lines) {
    size_t e = d.find('\n', pos);
    if (e == std::string::npos || e - pos > 16) fail(path + ": not a PFM file the reference reads");
    line = d.substr(pos, e - pos);
    pos = e + 1;
  }
  const char fmt = lines[0].size() > 1 ? lines[0][1] : 0;
  const uint32_t w = uint32_t(atoi(lines[1].c_str())), h = uint32_t(atoi(lines[2].c_str()));
  const uint32_t ch = (fmt == 'f') ? 1u : (fmt == 'F') ? 3u : 0u;
  if (ch == 0 || w == 0 || h == 0 || uint64_t(w) * h > d.size() || pos + size_t(w) * h * ch * 4 > d.size()) fail(path + ": unsupported or truncated PFM file");
  Pixels out;
  out.w = w, out.h = h;
  out.f32.resize(size_t(w) * h * 4);
  for (size_t i = 0; i < size_t(w) * h; ++i) {
    float v[3];
    memcpy(v, d.data() + pos + i * ch * 4, ch * 4);
    float* o = out.f32.data() + i * 4;
    o[0] = v[0], o[1] = (ch == 3) ? v[1] : v[0], o[2] = (ch == 3) ? v[2] : v[0], o[3] = 1.0f;
  }
  return out;
}

#include "scene_loader_jpeg.inl"

// Truevision TGA the way stb_image reads it (stb_image.hxx:5654-5990; it is the last format tried there, recognised by a consistent header): true colour
// 15 / 16 / 24 / 32 bits, grey 8 / 16 bits, colour-mapped with 8- or 16-bit indices, raw or run-length encoded, either row order.  15 / 16-bit colour is
// x555 with channels scaled by 255 / 31 (the top bit is not an alpha), 16-bit grey is grey + alpha (which the reference's pool then zero-fills).
bool looks_like_tga(const uint8_t* h, size_t n) {
  if (n < 18 || h[1] > 1) return false;
  const int type = h[2], bpp = h[16];
  if (h[1] == 1) {
    if (type != 1 && type != 9) return false;
    const int pal_bits = h[7];
    if (pal_bits != 8 && pal_bits != 15 && pal_bits != 16 && pal_bits != 24 && pal_bits != 32) return false;
  } else if (type != 2 && type != 3 && type != 10 && type != 11) {
    return false;
  }
  if ((h[12] | (h[13] << 8)) < 1 || (h[14] | (h[15] << 8)) < 1) return false;
  if (h[1] == 1 && bpp != 8 && bpp != 16) return false;
  return bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32;
}

Pixels read_tga(const std::string& path) {
  const std::string d = read_file(path);
  const uint8_t* b = reinterpret_cast<const uint8_t*>(d.data());
  if (!looks_like_tga(b, d.size())) fail(path + ": not a TGA file");
  size_t pos = 18;
  auto get8 = [&]() -> int { return pos < d.size() ? b[pos++] : 0; };
  auto get16 = [&]() -> int {
    int lo = get8();
    return lo | (get8() << 8);
  };
  const int id_length = b[0], indexed = b[1], pal_start = b[3] | (b[4] << 8), pal_len = b[5] | (b[6] << 8), pal_bits = b[7];
  const int w = b[12] | (b[13] << 8), h = b[14] | (b[15] << 8), bpp = b[16];
  int type = b[2];
  const bool rle = type >= 8;
  if (rle) type -= 8;
  const bool bottom_up = ((b[17] >> 5) & 1) == 0;
  bool rgb16 = false;
  auto channels_of = [&](int bits, bool grey) -> int {
    switch (bits) {
      case 8: return 1;
      case 16:
        if (grey) return 2;
        [[fallthrough]];
      case 15: rgb16 = true; return 3;
      case 24: return 3;
      case 32: return 4;
      default: return 0;
    }
  };
  const int comp = indexed ? channels_of(pal_bits, false) : channels_of(bpp, type == 3);
  if (comp == 0) fail(path + ": unsupported TGA pixel format");
  if (uint64_t(w) * h > uint64_t(d.size()) * 130u) fail(path + ": TGA header does not match the file size");  // a run packet covers <= 128 pixels
  auto read_rgb16 = [&](uint8_t* out) {
    const int px = get16();
    out[0] = uint8_t((((px >> 10) & 31) * 255) / 31), out[1] = uint8_t((((px >> 5) & 31) * 255) / 31), out[2] = uint8_t(((px & 31) * 255) / 31);
  };
  pos += size_t(id_length);
  std::vector<uint8_t> palette;
  if (indexed) {
    if (pal_len == 0) fail(path + ": TGA colour map without entries");
    pos += size_t(pal_start);
    palette.assign(size_t(pal_len) * comp, 0);
    for (int i = 0; i < pal_len; ++i) {
      if (rgb16) {
        read_rgb16(&palette[size_t(i) * comp]);
      } else {
        for (int j = 0; j < comp; ++j) palette[size_t(i) * comp + j] = uint8_t(get8());
      }
    }
  }
  std::vector<uint8_t> px(size_t(w) * h * comp, 0);
  uint8_t raw[4] = {0, 0, 0, 0};
  int run = 0;
  bool repeating = false;
  for (size_t i = 0; i < size_t(w) * h; ++i) {
    bool read = true;
    if (rle) {
      if (run == 0) {
        const int cmd = get8();
        run = 1 + (cmd & 127);
        repeating = (cmd >> 7) != 0;
      } else if (repeating) {
        read = false;
      }
    }
    if (read) {
      if (indexed) {
        int index = (bpp == 8) ? get8() : get16();
        if (index >= pal_len) index = 0;
        memcpy(raw, &palette[size_t(index) * comp], size_t(comp));
      } else if (rgb16) {
        read_rgb16(raw);
      } else {
        for (int j = 0; j < comp; ++j) raw[j] = uint8_t(get8());
      }
    }
    memcpy(&px[i * comp], raw, size_t(comp));
    --run;
  }
  Pixels out;
  out.w = uint32_t(w), out.h = uint32_t(h), out.eight_bit = true;
  out.u8.assign(size_t(w) * h * 4, 255);
  for (int y = 0; y < h; ++y) {
    const uint8_t* row = &px[size_t(bottom_up ? h - 1 - y : y) * w * comp];
    uint8_t* o = &out.u8[size_t(y) * w * 4];
    for (int x = 0; x < w; ++x, o += 4, row += comp) {
      if (comp == 1) {
        o[0] = o[1] = o[2] = row[0];
      } else if (comp == 2) {
        o[0] = o[1] = o[2] = o[3] = 0;  // grey + alpha: no case in the reference's switch over the channel count (image_pool.cxx:353-381)
      } else if (rgb16) {
        o[0] = row[0], o[1] = row[1], o[2] = row[2];
      } else {
        o[0] = row[2], o[1] = row[1], o[2] = row[0];  // stored blue first
        if (comp == 4) o[3] = row[3];
      }
    }
  }
  return out;
}

// Windows BMP the way stb_image reads it (stb_image.hxx:5324-5650): header sizes 12 / 40 / 56 / 108 / 124; 1 / 4 / 8 bits with a palette, 16 / 32 bits through
// channel masks (x555 and 8888 by default, BI_BITFIELDS otherwise; a channel of n bits is widened to 8 by bit replication), 24 bits; bottom-up or top-down rows;
// run-length and embedded JPEG / PNG variants are rejected there too.  A 32-bit file whose alpha bytes are all zero is opaque.
Pixels read_bmp(const std::string& path) {
  const std::string d = read_file(path);
  const uint8_t* b = reinterpret_cast<const uint8_t*>(d.data());
  size_t pos = 0;
  auto get8 = [&]() -> uint32_t { return pos < d.size() ? b[pos++] : 0u; };
  auto get16 = [&]() -> uint32_t {
    uint32_t lo = get8();
    return lo | (get8() << 8);
  };
  auto get32 = [&]() -> uint32_t {
    uint32_t lo = get16();
    return lo | (get16() << 16);
  };
  if (get8() != 'B' || get8() != 'M') fail(path + ": not a BMP file");
  get32(), get16(), get16();
  const int64_t offset = int32_t(get32());
  const uint32_t hsz = get32();
  if (offset < 0 || (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124)) fail(path + ": unsupported BMP header");
  int64_t width, height;
  if (hsz == 12) {
    width = get16(), height = get16();
  } else {
    width = int32_t(get32()), height = int32_t(get32());
  }
  if (get16() != 1) fail(path + ": bad BMP");
  const uint32_t bpp = get16();
  uint32_t mr = 0, mg = 0, mb = 0, ma = 0, all_a = 255, extra = 14;
  auto default_masks = [&]() {
    if (bpp == 16) {
      mr = 31u << 10, mg = 31u << 5, mb = 31u;
    } else if (bpp == 32) {
      mr = 0xffu << 16, mg = 0xffu << 8, mb = 0xffu, ma = 0xffu << 24;
      all_a = 0;
    } else {
      mr = mg = mb = ma = 0;
    }
  };
  if (hsz != 12) {
    const uint32_t compress = get32();
    if (compress == 1 || compress == 2 || compress >= 4) fail(path + ": this BMP compression is not read");
    if (compress == 3 && bpp != 16 && bpp != 32) fail(path + ": bad BMP");
    get32(), get32(), get32(), get32(), get32();
    if (hsz == 40 || hsz == 56) {
      if (hsz == 56) get32(), get32(), get32(), get32();
      if (bpp == 16 || bpp == 32) {
        if (compress == 0) {
          default_masks();
        } else {
          mr = get32(), mg = get32(), mb = get32();
          extra += 12;
          if (mr == mg && mg == mb) fail(path + ": bad BMP masks");
        }
      }
    } else {
      mr = get32(), mg = get32(), mb = get32(), ma = get32();
      if (compress != 3) default_masks();
      for (int i = 0; i < 13; ++i) get32();
      if (hsz == 124) get32(), get32(), get32(), get32();
    }
  }
  const bool bottom_up = height > 0;
  if (height < 0) height = -height;
  if (width <= 0 || height <= 0 || width > (1 << 24) || height > (1 << 24) || uint64_t(width) * uint64_t(height) > uint64_t(d.size()) * 8u) fail(path + ": BMP header does not match the file size");
  int64_t psize = 0;
  if (hsz == 12) {
    if (bpp < 24) psize = (offset - int64_t(extra) - 24) / 3;
  } else if (bpp < 16) {
    psize = (offset - int64_t(extra) - int64_t(hsz)) >> 2;
  }
  if (psize == 0 && uint64_t(offset) != pos) fail(path + ": bad BMP data offset");
  const uint32_t channels = (bpp == 24 && ma == 0xff000000u) ? 3u : (ma ? 4u : 3u);
  const size_t w = size_t(width), h = size_t(height);
  std::vector<uint8_t> px(w * h * channels, 0);
  size_t z = 0;
  if (bpp < 16) {
    if (psize <= 0 || psize > 256) fail(path + ": bad BMP palette");
    uint8_t pal[256][3] = {};
    for (int64_t i = 0; i < psize; ++i) {
      pal[i][2] = uint8_t(get8()), pal[i][1] = uint8_t(get8()), pal[i][0] = uint8_t(get8());
      if (hsz != 12) get8();
    }
    pos = std::min<size_t>(d.size(), pos + size_t(std::max<int64_t>(0, offset - int64_t(extra) - int64_t(hsz) - psize * (hsz == 12 ? 3 : 4))));
    if (bpp != 1 && bpp != 4 && bpp != 8) fail(path + ": bad BMP bit depth");
    const size_t row_bytes = (bpp == 1) ? (w + 7) >> 3 : ((bpp == 4) ? (w + 1) >> 1 : w), pad = (0 - row_bytes) & 3;
    for (size_t j = 0; j < h; ++j) {
      const size_t row_start = pos;
      for (size_t i = 0; i < w; ++i) {
        const size_t bit = i * bpp;
        const uint32_t byte = (row_start + (bit >> 3) < d.size()) ? b[row_start + (bit >> 3)] : 0u;
        uint32_t v = (byte >> (8 - bpp - (bit & 7))) & ((1u << bpp) - 1u);
        px[z++] = pal[v][0], px[z++] = pal[v][1], px[z++] = pal[v][2];
        if (channels == 4) px[z++] = 255;
      }
      pos = row_start + row_bytes + pad;
    }
  } else {
    pos = std::min<size_t>(d.size(), pos + size_t(std::max<int64_t>(0, offset - int64_t(extra) - int64_t(hsz))));
    const size_t row_bytes = (bpp == 24) ? 3 * w : ((bpp == 16) ? 2 * w : 0), pad = (0 - row_bytes) & 3;
    int easy = 0;
    if (bpp == 24) easy = 1;
    if (bpp == 32 && mb == 0xffu && mg == 0xff00u && mr == 0x00ff0000u && ma == 0xff000000u) easy = 2;
    auto high_bit = [](uint32_t v) {
      int n = -1;
      while (v) ++n, v >>= 1;
      return n;
    };
    auto bit_count = [](uint32_t v) {
      int n = 0;
      while (v) n += int(v & 1u), v >>= 1;
      return n;
    };
    int rs = 0, gs = 0, bs = 0, as = 0, rc = 0, gc = 0, bc = 0, ac = 0;
    if (!easy) {
      if (!mr || !mg || !mb) fail(path + ": bad BMP masks");
      rs = high_bit(mr) - 7, rc = bit_count(mr), gs = high_bit(mg) - 7, gc = bit_count(mg), bs = high_bit(mb) - 7, bc = bit_count(mb), as = high_bit(ma) - 7, ac = bit_count(ma);
      if (rc > 8 || gc > 8 || bc > 8 || ac > 8) fail(path + ": bad BMP masks");
    }
    auto widen = [](uint32_t v, int shift, int bits) -> uint32_t {  // a channel of `bits` bits to 8 bits by replication
      static const uint32_t mul[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01}, down[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
      v = (shift < 0) ? (v << -shift) : (v >> shift);
      v >>= (8 - bits);
      return (v * mul[bits]) >> down[bits];
    };
    for (size_t j = 0; j < h; ++j) {
      for (size_t i = 0; i < w; ++i) {
        uint32_t a;
        if (easy) {
          px[z + 2] = uint8_t(get8()), px[z + 1] = uint8_t(get8()), px[z + 0] = uint8_t(get8());
          z += 3;
          a = (easy == 2) ? get8() : 255u;
        } else {
          const uint32_t v = (bpp == 16) ? get16() : get32();
          px[z++] = uint8_t(widen(v & mr, rs, rc)), px[z++] = uint8_t(widen(v & mg, gs, gc)), px[z++] = uint8_t(widen(v & mb, bs, bc));
          a = ma ? widen(v & ma, as, ac) : 255u;
        }
        all_a |= a;
        if (channels == 4) px[z++] = uint8_t(a);
      }
      pos = std::min(d.size(), pos + pad);
    }
  }
  Pixels out;
  out.w = uint32_t(w), out.h = uint32_t(h), out.eight_bit = true;
  out.u8.assign(w * h * 4, 255);
  for (size_t y = 0; y < h; ++y) {
    const uint8_t* row = &px[(bottom_up ? h - 1 - y : y) * w * channels];
    uint8_t* o = &out.u8[y * w * 4];
    for (size_t x = 0; x < w; ++x, o += 4, row += channels) {
      o[0] = row[0], o[1] = row[1], o[2] = row[2];
      if (channels == 4) o[3] = (all_a == 0) ? 255 : row[3];
    }
  }
  return out;
}

// ImagePool::load_data (image_pool.cxx:271-383): the extension (compared as written) picks OpenEXR / Radiance HDR / PFM; every other file goes to
// stb_image there, which looks at the content
Pixels read_image(const std::string& path) {
  size_t dot_at = path.find_last_of('.');
  std::string ext = dot_at == std::string::npos ? std::string("") : path.substr(dot_at);
  if (ext == ".exr") return read_exr(path);
  if (ext == ".hdr") return read_hdr(path);
  if (ext == ".pfm") return read_pfm(path);
  uint8_t head[8] = {};
  if (FILE* f = fopen(path.c_str(), "rb")) {
    size_t got = fread(head, 1, sizeof(head), f);
    (void)got;
    fclose(f);
  } else {
    fail("cannot open " + path);
  }
  if (memcmp(head, "\x89PNG\r\n\x1a\n", 8) == 0) return read_png(path);
  if (head[0] == 0xff && head[1] == 0xd8) return read_jpeg(path);
  if (memcmp(head, "BM", 2) == 0) return read_bmp(path);
  if (memcmp(head, "GIF8", 4) != 0 && memcmp(head, "8BPS", 4) != 0 && memcmp(head, "#?RADIANCE", 8) != 0) {
    uint8_t header[18] = {};
    size_t got = 0;
    if (FILE* f = fopen(path.c_str(), "rb")) {
      got = fread(header, 1, sizeof(header), f);
      fclose(f);
    }
    if (looks_like_tga(header, got)) return read_tga(path);
  }
  fail(path + ": this image format is not read (PNG, JPEG, TGA, BMP, OpenEXR, Radiance HDR and PFM are)");
}

// ---- JSON (what a scene description needs: objects, arrays, strings, numbers, booleans) ----------------------------------------------------------------------
struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double number = 0.0;
  std::string text;
  std::vector<Json> items;
  std::vector<std::pair<std::string, Json>> members;
};
struct JsonParser {
  const std::string& s;
  size_t pos = 0;
  void ws() {
    while (pos < s.size() && isspace(static_cast<unsigned char>(s[pos]))) ++pos;
  }
  std::string string() {
    std::string out;
    ++pos;
    while (pos < s.size() && s[pos] != '"') {
      if (s[pos] == '\\' && pos + 1 < s.size()) {
        char c = s[pos + 1];
        out.push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c);
        pos += 2;
      } else {
        out.push_back(s[pos++]);
      }
    }
    ++pos;
    return out;
  }
  Json value() {
    ws();
    Json j;
    if (pos >= s.size()) fail("unexpected end of the scene description");
    char c = s[pos];
    if (c == '{') {
      j.kind = Json::Object;
      ++pos;
      ws();
      while (pos < s.size() && s[pos] != '}') {
        ws();
        if (s[pos] != '"') fail("bad scene description: a key was expected");
        std::string key = string();
        ws();
        if (pos >= s.size() || s[pos] != ':') fail("bad scene description: `:` was expected");
        ++pos;
        j.members.push_back({key, value()});
        ws();
        if (pos < s.size() && s[pos] == ',') ++pos;
        ws();
      }
      ++pos;
    } else if (c == '[') {
      j.kind = Json::Array;
      ++pos;
      ws();
      while (pos < s.size() && s[pos] != ']') {
        j.items.push_back(value());
        ws();
        if (pos < s.size() && s[pos] == ',') ++pos;
        ws();
      }
      ++pos;
    } else if (c == '"') {
      j.kind = Json::String;
      j.text = string();
    } else if (!s.compare(pos, 4, "true")) {
      j.kind = Json::Bool, j.b = true, pos += 4;
    } else if (!s.compare(pos, 5, "false")) {
      j.kind = Json::Bool, j.b = false, pos += 5;
    } else if (!s.compare(pos, 4, "null")) {
      pos += 4;
    } else {
      char* end = nullptr;
      j.kind = Json::Number;
      j.number = strtod(s.c_str() + pos, &end);
      if (end == s.c_str() + pos) fail("bad scene description: a value was expected");
      pos = size_t(end - s.c_str());
    }
    return j;
  }
};

// ---- .mtl as the reference's patched tinyobjloader reads it (tiny_obj_loader.hxx:1900-2190) ------------------------------------------------------------------
struct MtlBlock {
  std::string name;
  std::vector<std::pair<std::string, std::string>> params;  // in file order
  std::map<std::string, std::string> textures;
  bool get(const char* key, std::string& out) const {  // get_param (scene_representation.cxx:102-112): first match, case-insensitive
    std::string k = lower(key);
    for (const auto& p : params) {
      if (lower(p.first) == k) {
        out = p.second;
        return true;
      }
    }
    return false;
  }
  bool has(const char* key) const {
    std::string unused;
    return get(key, unused);
  }
};

// ParseTextureNameAndOption (tiny_obj_loader.hxx:1242-1321): the `-option args` of a texture statement are skipped by their argument COUNT (whatever the
// words are), the first word that is not an option starts the file name, which runs to the end of the line — spaces included
bool texture_name(const std::string& text, std::string& name) {
  static const struct {
    const char* key;
    int args;
  } options[] = {{"-blendu", 1}, {"-blendv", 1}, {"-clamp", 1}, {"-boost", 1}, {"-bm", 1}, {"-o", 3}, {"-s", 3}, {"-t", 3}, {"-type", 1}, {"-texres", 1}, {"-imfchan", 1}, {"-mm", 2},
                 {"-colorspace", 1}};
  size_t at = 0;
  const auto blank = [&](size_t i) { return i >= text.size() || text[i] == ' ' || text[i] == '\t'; };
  while (at < text.size()) {
    while (at < text.size() && (text[at] == ' ' || text[at] == '\t')) ++at;
    if (at >= text.size()) break;
    bool option = false;
    for (const auto& o : options) {
      const size_t n = strlen(o.key);
      if (lower(text.substr(at, n)) == o.key && at + n < text.size() && blank(at + n)) {
        at += n;
        for (int a = 0; a < o.args; ++a) {
          while (at < text.size() && (text[at] == ' ' || text[at] == '\t')) ++at;
          while (at < text.size() && text[at] != ' ' && text[at] != '\t' && text[at] != '\r') ++at;
        }
        option = true;
        break;
      }
    }
    if (!option) {
      name = text.substr(at);
      return true;
    }
  }
  return false;
}

std::vector<MtlBlock> parse_mtl(const std::string& path) {
  static const std::pair<const char*, const char*> textures[] = {{"map_ka", "ambient"}, {"map_kd", "diffuse"}, {"map_ks", "specular"}, {"map_kt", "transmittance"},
    {"map_ns", "specular_highlight"}, {"map_bump", "bump"}, {"map_d", "alpha"}, {"disp", "displacement"}, {"refl", "reflection"}, {"map_pr", "roughness"}, {"map_pm", "metallic"},
    {"map_ps", "sheen"}, {"map_ke", "emissive"}, {"norm", "normal"}};
  static const std::pair<const char*, bool> scalars[] = {{"Ni", false}, {"illum", true}, {"d", false}, {"Tr", false}, {"Pm", false}, {"Ps", false}, {"Pc", false}, {"Pcr", true},
    {"aniso", true}, {"anisor", true}};  // key, case-insensitive?
  std::string data = read_file(path);
  std::vector<MtlBlock> blocks;
  size_t at = 0;
  while (at <= data.size()) {
    size_t e = data.find('\n', at);
    if (e == std::string::npos) e = data.size();
    std::string line = data.substr(at, e - at);
    at = e + 1;
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
    size_t lead = 0;
    while (lead < line.size() && (line[lead] == ' ' || line[lead] == '\t')) ++lead;
    line = line.substr(lead);
    if (line.empty() || line[0] == '#') {
      if (e == data.size()) break;
      continue;
    }
    std::string low = lower(line);
    auto separated = [&](size_t n) { return line.size() > n && (line[n] == ' ' || line[n] == '\t'); };
    if (!low.compare(0, 6, "newmtl") && separated(6)) {
      MtlBlock b;
      b.name = lower(line.substr(7));
      blocks.push_back(b);
    } else if (!blocks.empty()) {
      bool consumed = false;
      for (const auto& t : textures) {
        size_t n = strlen(t.first);
        if (!low.compare(0, n, t.first) && separated(n)) {
          std::string name;
          if (texture_name(line.substr(n + 1), name)) blocks.back().textures[t.second] = name;  // without a name the earlier value stays
          consumed = true;
          break;
        }
      }
      for (const auto& sc : scalars) {
        if (consumed) break;
        size_t n = strlen(sc.first);
        bool match = sc.second ? !low.compare(0, n, lower(sc.first)) : !line.compare(0, n, sc.first);
        if (match && separated(n)) consumed = true;
      }
      if (!consumed) {
        size_t sp = line.find(' ');
        if (sp == std::string::npos) sp = line.find('\t');
        if (sp == std::string::npos) {
          blocks.back().params.push_back({line, ""});
        } else {
          blocks.back().params.push_back({line.substr(0, sp), line.substr(sp + 1)});
        }
      }
    }
    if (e == data.size()) break;
  }
  return blocks;
}

// ---- .obj: positions / normals / texture coordinates, triangulated faces with their material name and shape (o / g group) -------------------------------------
struct ObjIndex {
  int v = -1, t = -1, n = -1;
};
struct ObjData {
  std::vector<float> pos, nrm, tex;
  std::vector<ObjIndex> corners;  // 3 per face
  std::vector<int> face_material;  // index into material_names, -1 = none
  std::vector<uint32_t> face_shape;
  std::vector<std::string> material_names;
  std::string mtllib;
};

ObjData parse_obj(const std::string& path) {
  std::string data = read_file(path);
  ObjData o;
  std::map<std::string, int> name_index;
  int current = -1;
  uint32_t shape = 0;
  bool shape_has_faces = false;
  const char* p = data.c_str();
  const char* end = p + data.size();
  std::vector<ObjIndex> poly;
  while (p < end) {
    const char* le = static_cast<const char*>(memchr(p, '\n', size_t(end - p)));
    if (!le) le = end;
    const char* q = p;
    while (q < le && (*q == ' ' || *q == '\t')) ++q;
    if (q < le && *q != '#') {
      if (q[0] == 'v' && (q[1] == ' ' || q[1] == '\t')) {
        char* e2 = nullptr;
        for (int k = 0; k < 3; ++k) {
          o.pos.push_back(strtof(k ? e2 : q + 1, &e2));
        }
      } else if (q[0] == 'v' && q[1] == 'n' && (q[2] == ' ' || q[2] == '\t')) {
        char* e2 = nullptr;
        for (int k = 0; k < 3; ++k) o.nrm.push_back(strtof(k ? e2 : q + 2, &e2));
      } else if (q[0] == 'v' && q[1] == 't' && (q[2] == ' ' || q[2] == '\t')) {
        char* e2 = nullptr;
        float u = strtof(q + 2, &e2);
        const char* after = e2;
        float v = strtof(e2, &e2);
        if (e2 == after) v = 0.0f;
        o.tex.push_back(u), o.tex.push_back(v);
      } else if (q[0] == 'f' && (q[1] == ' ' || q[1] == '\t')) {
        poly.clear();
        const char* c = q + 1;
        while (c < le) {
          while (c < le && (*c == ' ' || *c == '\t' || *c == '\r')) ++c;
          if (c >= le) break;
          ObjIndex idx;
          char* e2 = nullptr;
          long vi = strtol(c, &e2, 10);
          if (e2 == c) break;
          long ti = 0, ni = 0;
          c = e2;
          if (c < le && *c == '/') {
            ++c;
            if (c < le && *c != '/') {
              ti = strtol(c, &e2, 10);
              c = e2;
            }
            if (c < le && *c == '/') {
              ++c;
              ni = strtol(c, &e2, 10);
              c = e2;
            }
          }
          idx.v = int(vi > 0 ? vi - 1 : long(o.pos.size() / 3) + vi);
          idx.t = ti ? int(ti > 0 ? ti - 1 : long(o.tex.size() / 2) + ti) : -1;
          idx.n = ni ? int(ni > 0 ? ni - 1 : long(o.nrm.size() / 3) + ni) : -1;
          poly.push_back(idx);
        }
        auto emit_corners = [&](const ObjIndex& a, const ObjIndex& b2, const ObjIndex& c2) {
          o.corners.push_back(a), o.corners.push_back(b2), o.corners.push_back(c2);
          o.face_material.push_back(current);
          o.face_shape.push_back(shape);
        };
        auto emit = [&](int a, int b2, int c2) { emit_corners(poly[size_t(a)], poly[size_t(b2)], poly[size_t(c2)]); };
        for (const auto& ix : poly)
          if (ix.v < 0 || size_t(ix.v) * 3 + 2 >= o.pos.size()) fail(path + ": a face refers to a vertex that does not exist");
        if (poly.size() == 4) {
          // tinyobjloader splits a quad along its shorter diagonal (tiny_obj_loader.hxx:1464-1527)
          F3 v0 = load3(&o.pos[size_t(poly[0].v) * 3]), v1 = load3(&o.pos[size_t(poly[1].v) * 3]), v2 = load3(&o.pos[size_t(poly[2].v) * 3]), v3 = load3(&o.pos[size_t(poly[3].v) * 3]);
          F3 e02 = v2 - v0, e13 = v3 - v1;
          float s02 = e02.x * e02.x + e02.y * e02.y + e02.z * e02.z, s13 = e13.x * e13.x + e13.y * e13.y + e13.z * e13.z;
          if (s02 < s13) {
            emit(0, 1, 2), emit(0, 2, 3);
          } else {
            emit(0, 1, 3), emit(1, 2, 3);
          }
        } else if (poly.size() == 3) {
          emit(0, 1, 2);
        } else if (poly.size() > 4) {
          // tinyobjloader clips ears in the plane of the polygon's two dominant axes (tiny_obj_loader.hxx:1537-1800), in float
          size_t axes[2] = {1, 2};
          const size_t n = poly.size();
          for (size_t k = 0; k < n; ++k) {
            const float* a = &o.pos[size_t(poly[k % n].v) * 3];
            const float* b2 = &o.pos[size_t(poly[(k + 1) % n].v) * 3];
            const float* c2 = &o.pos[size_t(poly[(k + 2) % n].v) * 3];
            const float e0x = b2[0] - a[0], e0y = b2[1] - a[1], e0z = b2[2] - a[2], e1x = c2[0] - b2[0], e1y = c2[1] - b2[1], e1z = c2[2] - b2[2];
            const float cx = std::fabs(e0y * e1z - e0z * e1y), cy = std::fabs(e0z * e1x - e0x * e1z), cz = std::fabs(e0x * e1y - e0y * e1x);
            if (cx > FLT_EPSILON || cy > FLT_EPSILON || cz > FLT_EPSILON) {  // the first real corner decides
              if (!(cx > cy && cx > cz)) {
                axes[0] = 0;
                if (cz > cx && cz > cy) axes[1] = 1;
              }
              break;
            }
          }
          std::vector<ObjIndex> rest = poly;
          size_t guess = 0, budget = n, previous = n;
          while (rest.size() > 3 && budget > 0) {
            const size_t m = rest.size();
            if (guess >= m) guess -= m;
            if (previous != m) {
              previous = m;
              budget = m;
            } else {
              budget--;
            }
            ObjIndex ind[3];
            float vx[3], vy[3];
            for (size_t k = 0; k < 3; ++k) {
              ind[k] = rest[(guess + k) % m];
              vx[k] = o.pos[size_t(ind[k].v) * 3 + axes[0]];
              vy[k] = o.pos[size_t(ind[k].v) * 3 + axes[1]];
            }
            const float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0], e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
            const float turn = e0x * e1y - e0y * e1x, area = (vx[0] * vy[1] - vy[0] * vx[1]) * 0.5f;
            if (turn * area < 0.0f) {  // "an internal angle"
              guess += 1;
              continue;
            }
            bool overlap = false;
            for (size_t other = 3; other < m && !overlap; ++other) {
              const ObjIndex& t = rest[(guess + other) % m];
              const float tx = o.pos[size_t(t.v) * 3 + axes[0]], ty = o.pos[size_t(t.v) * 3 + axes[1]];
              bool inside = false;  // the crossing-number test over the candidate ear
              for (int i = 0, j = 2; i < 3; j = i++) {
                if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) inside = !inside;
              }
              overlap = inside;
            }
            if (overlap) {
              guess += 1;
              continue;
            }
            emit_corners(ind[0], ind[1], ind[2]);
            rest.erase(rest.begin() + long((guess + 1) % m));
          }
          if (rest.size() == 3) emit_corners(rest[0], rest[1], rest[2]);
        }
        shape_has_faces = shape_has_faces || (poly.size() >= 3);
      } else if (!strncmp(q, "usemtl", 6) && (q[6] == ' ' || q[6] == '\t')) {
        std::string name = lower(trim(std::string(q + 7, size_t(le - q - 7))));
        auto it = name_index.find(name);
        if (it == name_index.end()) {
          current = int(o.material_names.size());
          name_index[name] = current;
          o.material_names.push_back(name);
        } else {
          current = it->second;
        }
      } else if ((q[0] == 'o' || q[0] == 'g') && (q + 1 == le || q[1] == ' ' || q[1] == '\t' || q[1] == '\r')) {
        if (shape_has_faces) {
          shape += 1;
          shape_has_faces = false;
        }
      } else if (!strncmp(q, "mtllib", 6) && (q[6] == ' ' || q[6] == '\t')) {
        o.mtllib = trim(std::string(q + 7, size_t(le - q - 7)));
      }
    }
    p = le + 1;
  }
  return o;
}
