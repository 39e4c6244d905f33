#!/bin/bash
# Builds the CPU oracle from the reference's own sources WHERE THEY LIE (never copied into this repo).
# Outputs only into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
#   liboracle_parity.so : -O2 -ffp-contract=off + portable libm overrides  -> bit-exact partner of the CUDA parity build
#   liboracle_native.so : -O3 -march=x86-64-v3, glibc libm                 -> "the reference's CPU VCM" for timing
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${ETX_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
CSRC="$HERE/../etx_tracer_b200/csrc"
if [ ! -d "$REF/sources/etx" ]; then
  echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
INC="-I$REF/sources -I$REF/thirdparty/bluenoise -I$CSRC"
COMMON="-std=c++23 -DNDEBUG -fPIC -w -pthread"
SRC="$HERE/oracle_vcm.cxx $HERE/layout_check.cxx $REF/sources/etx/render/host/spectrum.cxx $REF/thirdparty/bluenoise/bluenoise.cxx $CSRC/bvh_build.cpp"
build() { # name flags extra_sources
  local name="$1"; shift
  local flags="$1"; shift
  local objs=""
  for s in $SRC "$@"; do
    local o="$OUT/$name.$(basename "$s").o"
    g++ $COMMON $flags $INC -c "$s" -o "$o" &
    objs="$objs $o"
  done
  wait
  g++ -shared -o "$OUT/lib$name.so" $objs -Wl,-Bsymbolic -pthread
  rm -f $objs
}
build oracle_parity "-O2 -ffp-contract=off -fno-builtin-sincos" "$HERE/libm_override.cxx"
build oracle_native "-O3 -march=x86-64-v3"

# libreference_loader.so : the reference's OWN scene loader (scene_representation.cxx + pools + third-party readers) compiled in place,
# with the three pieces that have no Linux implementation supplied by oracle/ref_loader.cxx (TaskScheduler, console colour, filter image).
build_loader() {
  local T="$REF/thirdparty" R="$REF/sources/etx"
  local LINC="-I$REF/sources -I$T -I$T/json -I$T/tinyobjloader -I$T/stb_image -I$T/tinyexr -I$T/mikktspace -I$T/enkits -I$T/nanovdb -I$T/tinygltf"
  local LFLAGS="-std=c++23 -DNDEBUG -fPIC -w -O2 -D_stricmp=strcasecmp -include etx/core/log.hxx"
  local objs=""
  for s in "$HERE/ref_loader.cxx" "$R/render/host/scene_representation.cxx" "$R/render/host/medium_pool.cxx" "$R/render/host/image_pool.cxx" \
           "$R/render/host/scattering.cxx" "$R/render/host/gltf_accessor.cxx" "$R/render/host/spectrum.cxx" "$R/core/core.cxx" "$R/core/environment.cxx" \
           "$R/core/log.cxx" "$T/tinyobjloader/tiny_obj_loader.cxx" "$T/stb_image/stb_image.cxx" "$T/tinyexr/tinyexr.cxx" "$T/tinygltf/tiny_gltf.cxx"; do
    local o="$OUT/loader.$(basename "$s").o"
    g++ $LFLAGS $LINC -c "$s" -o "$o" &
    objs="$objs $o"
  done
  gcc -fPIC -w -O2 -c "$T/mikktspace/mikktspace.c" -o "$OUT/loader.mikktspace.o" &
  objs="$objs $OUT/loader.mikktspace.o"
  wait
  g++ -shared -o "$OUT/libreference_loader.so" $objs -pthread
  rm -f $objs
}
if build_loader; then :; else echo "reference loader did not build; tests that use it will skip" >&2; fi
# nvdb_make : writes NanoVDB test files with the reference's own NanoVDB headers (oracle/nvdb_make.cxx)
if g++ -std=c++17 -O1 -w -I"$REF/thirdparty/nanovdb" "$HERE/nvdb_make.cxx" -o "$OUT/nvdb_make" -pthread; then :; else echo "nvdb_make did not build; the .nvdb tests will skip" >&2; fi
echo "built: $(ls $OUT/*.so) $(ls $OUT/nvdb_make 2>/dev/null)"
