#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the CPU oracle from the reference's OWN sources.
#
# Compiles the unmodified files under $RAGNAR_REFERENCE (default /root/reference)
# against the Kokkos-subset shim in oracle/shim/ into two importable modules:
#   oracle/_ref/ragnar_ref.*.so     float ScatterView   (faithful accumulation)
#   oracle/_ref/ragnar_ref64.*.so   double ScatterView  (reference terms, wide sum)
# The reference's own build system (CMake + FetchContent of Kokkos/HighFive) is
# NOT run: it needs network access.  io/ and plugins/ need libhdf5 (absent) and
# are left out exactly as the reference's CI does (-D RAGNAR_USE_HDF5=OFF).
# No reference source is copied into this repo; outputs go to oracle/_ref/ only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${RAGNAR_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
PYINC="$(python3 -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
EXT="$(python3 -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF/src not found (reference absent) — nothing to do" >&2
  exit 0
fi
mkdir -p "$OUT"
SRCS="$REF/src/pyinterface.cpp $(ls "$REF"/src/utils/*.cpp "$REF"/src/containers/*.cpp "$REF"/src/physics/*.cpp)"
# No -march / no FMA contraction: the reference sets no arch flags, and fusing
# ux*ux+uy*uy+uz*uz moves histogram counts.
FLAGS="-std=c++20 -O2 -fPIC -shared -fopenmp -ffp-contract=off -fvisibility=hidden -w"
INCS="-I$HERE/shim -I$REF/src -I$REF/extern/pybind11/include -I$PYINC"
build_one() { # name extra-flags
  local name="$1"; shift
  local target="$OUT/$name$EXT"
  if [ -f "$target" ] && [ "$target" -nt "$HERE/shim/Kokkos_Core.hpp" ] && \
     [ "$target" -nt "$HERE/shim/Kokkos_ScatterView.hpp" ] && [ -z "${FORCE:-}" ]; then
    echo "build_ref.sh: $target up to date"
    return
  fi
  echo "build_ref.sh: building $target"
  # PYBIND11_BUILD_ABI gives each oracle module a private pybind11 type registry,
  # so ragnar_ref, ragnar_ref64 and the product module import side by side.
  g++ $FLAGS "-Dragnar=$name" "-DPYBIND11_BUILD_ABI=\"_oracle_$name\"" "$@" $INCS $SRCS -o "$target"
}
build_one ragnar_ref &
build_one ragnar_ref64 -DRAGNAR_ORACLE_SCATTER64 &
wait
ls -la "$OUT"
