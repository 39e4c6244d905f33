// scene_loader_nvdb.inl — the density grid of an `et::medium … volume <file>.nvdb` block (included by scene_loader.cpp inside its anonymous namespace).
//
// The reference reads the first grid of the file with NanoVDB's own reader and, when it is a float grid, samples EVERY voxel of the tree's index
// bounding box [min, max) through an accessor into a dense x-fastest array, keeping the values > 0 (MediumPool::load_nvdb, medium_pool.cxx:102-159);
// MediumPool::add then scales the array to a maximum of 1 and marks the medium heterogeneous (:41-59).  This file does the same without the NanoVDB
// headers: the published NanoVDB 32.x layout — file header, per-grid meta data, GridData (672 B), TreeData (64 B), a root table of 32-byte tiles, upper
// (32^3) and lower (16^3) internal nodes with value / child masks and 8-byte table entries holding either a value or a byte offset to the child, 8^3 leaf
// nodes — is walked directly in the file's bytes.  Codecs: none and ZIP (the module's own inflate); BLOSC files are refused.  Every offset taken from
// the file is range-checked.  tests/test_loader.py writes files with the reference's NanoVDB headers (oracle/nvdb_make.cxx, test infrastructure) and
// compares the resulting Medium records with the reference loader's.

struct DensityGrid {
  std::vector<float> values;  // x fastest, then y, then z; empty = the medium stays homogeneous
  uint32_t dim[3] = {0, 0, 0};
};

namespace nvdb {

constexpr uint64_t kMagic = 0x304244566f6e614eull;  // "NanoVDB0"
constexpr size_t kGridDataBytes = 672, kTreeDataBytes = 64, kRootDataBytes = 64, kRootTileBytes = 32, kMetaDataBytes = 176;
constexpr size_t kUpperTable = 8256, kUpperBytes = kUpperTable + 32768 * 8, kUpperChildMask = 32 + 4096;   // 32^3 entries
constexpr size_t kLowerTable = 1088, kLowerBytes = kLowerTable + 4096 * 8, kLowerChildMask = 32 + 512;     // 16^3 entries
constexpr size_t kLeafValues = 96, kLeafBytes = kLeafValues + 512 * 4;                                     // 8^3 floats

struct View {
  const uint8_t* base = nullptr;
  size_t size = 0;
  const std::string* path = nullptr;
  const uint8_t* at(uint64_t offset, size_t bytes) const {
    if (offset > size || size - offset < bytes) fail(*path + ": NanoVDB data points outside the file");
    return base + offset;
  }
  template <class T>
  T read(uint64_t offset) const {
    T v;
    memcpy(&v, at(offset, sizeof(T)), sizeof(T));
    return v;
  }
};

struct Tree {
  View g;
  uint64_t root = 0;
  uint32_t tiles = 0;
  float background = 0.0f;
  // the leaf the last lookup ended in (the dense sweep visits 8 consecutive x of one leaf row)
  mutable int32_t leaf_origin[3] = {0, 0, 0};
  mutable uint64_t leaf = 0;

  static bool mask_on(const View& g, uint64_t mask, uint32_t n) { return (g.read<uint64_t>(mask + uint64_t(n >> 6) * 8) >> (n & 63u)) & 1ull; }

  float value(int32_t i, int32_t j, int32_t k) const {
    if (leaf != 0 && (i & ~7) == leaf_origin[0] && (j & ~7) == leaf_origin[1] && (k & ~7) == leaf_origin[2]) {
      return g.read<float>(leaf + kLeafValues + uint64_t(((i & 7) << 6) + ((j & 7) << 3) + (k & 7)) * 4);
    }
    const uint64_t key = (uint64_t(uint32_t(k) >> 12)) | (uint64_t(uint32_t(j) >> 12) << 21) | (uint64_t(uint32_t(i) >> 12) << 42);
    uint64_t upper = 0;
    bool found = false;
    for (uint32_t t = 0; t < tiles && !found; ++t) {
      const uint64_t tile = root + kRootDataBytes + uint64_t(t) * kRootTileBytes;
      if (g.read<uint64_t>(tile) != key) continue;
      found = true;
      const int64_t child = g.read<int64_t>(tile + 8);
      if (child == 0) return g.read<float>(tile + 20);
      upper = uint64_t(int64_t(root) + child);
    }
    if (!found) return background;
    g.at(upper, kUpperBytes);
    const uint32_t nu = (uint32_t((i & 4095) >> 7) << 10) + (uint32_t((j & 4095) >> 7) << 5) + uint32_t((k & 4095) >> 7);
    if (!mask_on(g, upper + kUpperChildMask, nu)) return g.read<float>(upper + kUpperTable + uint64_t(nu) * 8);
    const uint64_t lower = uint64_t(int64_t(upper) + g.read<int64_t>(upper + kUpperTable + uint64_t(nu) * 8));
    g.at(lower, kLowerBytes);
    const uint32_t nl = (uint32_t((i & 127) >> 3) << 8) + (uint32_t((j & 127) >> 3) << 4) + uint32_t((k & 127) >> 3);
    if (!mask_on(g, lower + kLowerChildMask, nl)) return g.read<float>(lower + kLowerTable + uint64_t(nl) * 8);
    const uint64_t lf = uint64_t(int64_t(lower) + g.read<int64_t>(lower + kLowerTable + uint64_t(nl) * 8));
    g.at(lf, kLeafBytes);
    leaf = lf;
    leaf_origin[0] = i & ~7, leaf_origin[1] = j & ~7, leaf_origin[2] = k & ~7;
    return g.read<float>(lf + kLeafValues + uint64_t(((i & 7) << 6) + ((j & 7) << 3) + (k & 7)) * 4);
  }
};

}  // namespace nvdb

DensityGrid read_nvdb_density(const std::string& path, std::vector<std::string>& warnings) {
  DensityGrid out;
  const std::string d = read_file(path);
  nvdb::View file{reinterpret_cast<const uint8_t*>(d.data()), d.size(), &path};
  if (d.size() < 16 || file.read<uint64_t>(0) != nvdb::kMagic) fail(path + ": not a NanoVDB file");
  const uint32_t major = file.read<uint32_t>(8) >> 21;
  if (major != 32u) fail(path + ": NanoVDB file format " + std::to_string(major) + ".x; 32.x is what the reference (and this loader) reads");
  const uint32_t grid_count = file.read<uint16_t>(12);
  if (grid_count == 0) fail(path + ": the NanoVDB file holds no grid");
  // meta data of every grid precedes the first grid's bytes
  uint64_t pos = 16, grid_size = 0;
  uint32_t grid_type = 0, codec = 0;
  for (uint32_t i = 0; i < grid_count; ++i) {
    file.at(pos, nvdb::kMetaDataBytes);
    if (i == 0) {
      grid_size = file.read<uint64_t>(pos);
      grid_type = file.read<uint32_t>(pos + 32);
      codec = file.read<uint16_t>(pos + 168);
    }
    const uint32_t name_size = file.read<uint32_t>(pos + 136);
    pos += nvdb::kMetaDataBytes + name_size;
  }
  if (grid_type != 1u) {  // GridType::Float; handle.grid<float>() is null for anything else and the reference keeps the medium homogeneous
    warnings.push_back(path + ": the first grid is not a float grid, the medium stays homogeneous");
    return out;
  }
  std::vector<uint8_t> unpacked;
  nvdb::View g;
  g.path = &path;
  if (codec == 0u) {
    g.base = file.at(pos, size_t(grid_size));
    g.size = size_t(grid_size);
  } else if (codec == 1u) {
    const uint64_t packed = file.read<uint64_t>(pos);
    unpacked = inflate_zlib(file.at(pos + 8, size_t(packed)), size_t(packed));
    if (unpacked.size() != grid_size) fail(path + ": the compressed NanoVDB grid does not unpack to its stated size");
    g.base = unpacked.data();
    g.size = unpacked.size();
  } else {
    fail(path + ": BLOSC-compressed NanoVDB files are not read (none and ZIP are)");
  }
  if (g.size < nvdb::kGridDataBytes + nvdb::kTreeDataBytes || g.read<uint64_t>(0) != nvdb::kMagic) fail(path + ": bad NanoVDB grid");
  nvdb::Tree tree;
  tree.g = g;
  tree.root = nvdb::kGridDataBytes + g.read<uint64_t>(nvdb::kGridDataBytes + 24);  // TreeData::mNodeOffset[3], relative to the tree
  g.at(tree.root, nvdb::kRootDataBytes);
  int32_t lo[3], hi[3];
  for (int a = 0; a < 3; ++a) lo[a] = g.read<int32_t>(tree.root + uint64_t(a) * 4), hi[a] = g.read<int32_t>(tree.root + 12 + uint64_t(a) * 4);
  tree.tiles = g.read<uint32_t>(tree.root + 24);
  tree.background = g.read<float>(tree.root + 28);
  g.at(tree.root + nvdb::kRootDataBytes, size_t(tree.tiles) * nvdb::kRootTileBytes);
  if (hi[0] <= lo[0] || hi[1] <= lo[1] || hi[2] <= lo[2]) return out;  // an empty tree has an inverted box
  const uint64_t dx = uint64_t(int64_t(hi[0]) - lo[0]), dy = uint64_t(int64_t(hi[1]) - lo[1]), dz = uint64_t(int64_t(hi[2]) - lo[2]);
  if (dx > 4096 || dy > 4096 || dz > 4096 || dx * dy * dz > (1ull << 30)) fail(path + ": the volume's index box is too large for a dense grid");
  out.dim[0] = uint32_t(dx), out.dim[1] = uint32_t(dy), out.dim[2] = uint32_t(dz);
  out.values.assign(size_t(dx * dy * dz), 0.0f);
  float min_val = 3.402823466e+38f, max_val = -3.402823466e+38f;
  double sum = 0.0;
  uint64_t count = 0;
  for (int32_t z = lo[2]; z < hi[2]; ++z) {
    for (int32_t y = lo[1]; y < hi[1]; ++y) {
      for (int32_t x = lo[0]; x < hi[0]; ++x) {  // [min, max): the reference leaves the box's last layer out (:125-139)
        const float val = tree.value(x, y, z);
        if (val > 0.0f) {
          min_val = std::min(min_val, val);
          max_val = std::max(max_val, val);
          out.values[size_t(x - lo[0]) + size_t(y - lo[1]) * dx + size_t(z - lo[2]) * dx * dy] = val;
          count += 1;
          sum += val;
        }
      }
    }
  }
  const double avg = sum / float(count);
  if ((count == 0) || (min_val == 3.402823466e+38f) || ((max_val - min_val) <= kEps) || (avg <= kEps)) {
    warnings.push_back(path + ": density is zero or too small, the medium stays homogeneous");
    out = DensityGrid();
  }
  return out;
}
