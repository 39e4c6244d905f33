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
echo "built: $(ls $OUT/*.so)"
