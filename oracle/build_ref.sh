#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle/): builds the UNMODIFIED reference CPU solver + our probe
# drivers into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
#
# The reference sources are read from /root/reference/src.  Nothing is copied into the
# repo: a scratch copy is made under $TMPDIR only because (a) the fp64 build needs the
# reference's own one-line switch `#define FTYPE double` (src/Common/Geometry.h:21 has no
# #ifndef guard) and (b) the 2D sources use Windows include paths.
# The reference's own build system is NOT used (Fermi flags, libnetcdf dependency).
#
# Outputs: oracle/_ref/ref_probe3d_f32, ref_probe3d_f64, ref_probe2d_f32, dropin3d_f32/f64, dropin2d_f32 (+ build.log)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CMC_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/FluidSolver3D" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
OPT="${CMC_REF_OPT:--O2}"
WORK="$(mktemp -d "${TMPDIR:-/tmp}/cmc_ref_build.XXXXXX")"
trap 'rm -rf "$WORK"' EXIT

build3d() { # $1 = f32|f64
  local tag="$1" dir="$WORK/$1"
  mkdir -p "$dir"
  cp -r "$REF/src" "$dir/src"
  chmod -R u+w "$dir/src"
  if [ "$tag" = f64 ]; then
    sed -i 's/^#define FTYPE\s\+float/#define FTYPE\t\t\tdouble/' "$dir/src/Common/Geometry.h"
    grep -q 'define FTYPE.*double' "$dir/src/Common/Geometry.h"
  fi
  local s="$dir/src/FluidSolver3D"
  local inc="-I$s -I$CUDA/include -I$dir/src/NetCDF -I$dir/src/FluidSolver2D"
  local cxx="g++ -fopenmp $OPT -fpermissive -w $inc"
  ( cd "$s"
    for f in AdiSolver3D.cpp Grid3D.cpp Solver3D.cpp ../FluidSolver2D/Grid2D.cpp \
             ../Common/LinuxIO.cpp ../Common/GPUplan.cpp ../Common/PARAplan.cpp; do
      $cxx -c "$f" -o "$dir/$(basename "$f").o" &
    done
    # the CPU path still needs the GPU twins at link time (AdiSolver3D.h:40-46)
    for f in AdiSolver3D.cu TimeLayer3D.cu; do
      nvcc -arch=sm_100 -O2 -w -Xcompiler -fpermissive $inc -c "$f" -o "$dir/$f.o" &
    done
    $cxx -c "$HERE/ref_probe3d.cpp" -o "$dir/ref_probe3d.o" &
    gcc -O2 -c "$HERE/netcdf_shim.c" -o "$dir/netcdf_shim.o" &
    wait
  )
  g++ -fopenmp -o "$OUT/ref_probe3d_$tag" "$dir"/*.o -L"$CUDA/lib64" -lcudart_static -ldl -lpthread -lrt
  echo "built $OUT/ref_probe3d_$tag"
  # drop-in demonstration: the same driver + the reference's loader objects + our Solver3D adapter + libcmcadi.so
  local pkg="$HERE/../cmc_fluid_solver_b200"
  if [ -f "$pkg/libcmcadi.so" ]; then
    $cxx -DWITH_B200 -I"$pkg/host" -I"$HERE/../include" -c "$HERE/ref_probe3d.cpp" -o "$dir/ref_probe3d.o"
    $cxx -I"$pkg/host" -I"$HERE/../include" -c "$pkg/host/B200AdiSolver3D.cpp" -o "$dir/B200AdiSolver3D.o"
    g++ -fopenmp -o "$OUT/dropin3d_$tag" "$dir"/*.o -L"$CUDA/lib64" -lcudart_static -ldl -lpthread -lrt \
        -L"$pkg" -lcmcadi -Wl,-rpath,'$ORIGIN/../../cmc_fluid_solver_b200'
    echo "built $OUT/dropin3d_$tag"
  fi
}

build2d() { # fp32 only (BASELINE config 1 is the reference CPU case)
  local dir="$WORK/2d"
  mkdir -p "$dir"
  cp -r "$REF/src" "$dir/src"
  chmod -R u+w "$dir/src"
  local s="$dir/src/FluidSolver2D"
  sed -i 's#\.\.\\Common\\#../Common/#g' "$s"/*.h "$s"/*.cpp
  local inc="-I$s -I$dir/src/NetCDF -I$CUDA/include"
  local cxx="g++ -fopenmp $OPT -fpermissive -w $inc"
  ( cd "$s"
    for f in AdiSolver2D.cpp Grid2D.cpp Solver2D.cpp ../Common/LinuxIO.cpp; do
      $cxx -c "$f" -o "$dir/$(basename "$f").o" &
    done
    $cxx -c "$HERE/ref_probe2d.cpp" -o "$dir/ref_probe2d.o" &
    gcc -O2 -c "$HERE/netcdf_shim.c" -o "$dir/netcdf_shim.o" &
    wait
  )
  g++ -fopenmp -o "$OUT/ref_probe2d_f32" "$dir"/*.o -lrt
  echo "built $OUT/ref_probe2d_f32"
  # 2D drop-in demonstration: reference loader + Grid2D + our Solver2D adapter + libcmcadi.so
  local pkg="$HERE/../cmc_fluid_solver_b200"
  if [ -f "$pkg/libcmcadi.so" ]; then
    $cxx -DWITH_B200 -I"$pkg/host" -I"$HERE/../include" -c "$HERE/ref_probe2d.cpp" -o "$dir/ref_probe2d.o"
    $cxx -I"$pkg/host" -I"$HERE/../include" -c "$pkg/host/B200AdiSolver2D.cpp" -o "$dir/B200AdiSolver2D.o"
    g++ -fopenmp -o "$OUT/dropin2d_f32" "$dir"/*.o -lrt -L"$pkg" -lcmcadi -Wl,-rpath,'$ORIGIN/../../cmc_fluid_solver_b200'
    echo "built $OUT/dropin2d_f32"
  fi
}

{
  build3d f32
  build3d f64
  if [ -f "$HERE/ref_probe2d.cpp" ]; then build2d; fi
} 2>&1 | tee "$OUT/build.log"
