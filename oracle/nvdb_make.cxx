// nvdb_make.cxx — TEST INFRASTRUCTURE: writes NanoVDB files with the reference's own NanoVDB headers (thirdparty/nanovdb, compiled in place by
// oracle/build_oracle.sh into oracle/_ref/nvdb_make), so that the module's .nvdb reader (csrc/scene_loader_nvdb.inl) can be tested against the
// reference's MediumPool::load_nvdb on the same file.  Never on the product path.
//
//   nvdb_make <kind> <out.nvdb>     kind: sphere   fog sphere of radius 20 around (3, -2, 5) (createFogVolumeSphere)
//                                         blobs    a hand-made float grid: two overlapping blobs across node boundaries + a far-away voxel block
//                                         empty    a float grid whose values are all zero
//                                         vec3     a Vec3f grid (not a float grid: the reference then keeps the medium homogeneous)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include <NanoVDB.h>
#include <util/GridBuilder.h>
#include <util/IO.h>
#include <util/Primitives.h>

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s sphere|blobs|empty|vec3 out.nvdb\n", argv[0]);
    return 2;
  }
  const std::string kind = argv[1];
  try {
    if (kind == "sphere") {
      auto handle = nanovdb::createFogVolumeSphere<float>(20.0f, nanovdb::Vec3f(3.0f, -2.0f, 5.0f), 1.0, 3.0, nanovdb::Vec3d(0.0), "density");
      nanovdb::io::writeGrid(argv[2], handle);
    } else if (kind == "blobs") {
      nanovdb::GridBuilder<float> builder(0.0f, nanovdb::GridClass::FogVolume);
      auto blob = [](const nanovdb::Coord& c) -> float {
        auto g = [&](float cx, float cy, float cz, float r) {
          float dx = c[0] - cx, dy = c[1] - cy, dz = c[2] - cz;
          float d = std::sqrt(dx * dx + dy * dy + dz * dz) / r;
          return d < 1.0f ? (1.0f - d * d) : 0.0f;
        };
        float v = 0.7f * g(-6.0f, 4.0f, 130.0f, 11.0f) + 1.9f * g(9.0f, -3.0f, 122.0f, 7.5f);
        return v > 0.02f ? v : 0.0f;
      };
      builder(blob, nanovdb::CoordBBox(nanovdb::Coord(-20, -12, 108), nanovdb::Coord(20, 18, 144)));
      builder([](const nanovdb::Coord& c) { return 0.25f + 0.001f * float(c[0] & 7); }, nanovdb::CoordBBox(nanovdb::Coord(40, 2, 120), nanovdb::Coord(43, 5, 123)));
      auto handle = builder.getHandle<>(0.5, nanovdb::Vec3d(0.0), "density");
      nanovdb::io::writeGrid(argv[2], handle);
    } else if (kind == "empty") {
      nanovdb::GridBuilder<float> builder(0.0f, nanovdb::GridClass::FogVolume);
      builder([](const nanovdb::Coord&) { return -1.0f; }, nanovdb::CoordBBox(nanovdb::Coord(0), nanovdb::Coord(9)));
      auto handle = builder.getHandle<>(1.0, nanovdb::Vec3d(0.0), "density");
      nanovdb::io::writeGrid(argv[2], handle);
    } else if (kind == "vec3") {
      nanovdb::GridBuilder<nanovdb::Vec3f> builder(nanovdb::Vec3f(0.0f));
      builder([](const nanovdb::Coord& c) { return nanovdb::Vec3f(float(c[0]), 1.0f, 2.0f); }, nanovdb::CoordBBox(nanovdb::Coord(0), nanovdb::Coord(7)));
      auto handle = builder.getHandle<>(1.0, nanovdb::Vec3d(0.0), "velocity");
      nanovdb::io::writeGrid(argv[2], handle);
    } else {
      return 2;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
