#!/bin/bash
# Builds the UNMODIFIED reference (jkxing/EPSM_Mitsuba3 = Mitsuba 3.2.0 + Dr.Jit 0.4.3 + the EPSM patch set) from the
# sources where they lie under /root/reference, out of tree, into baseline/_ref/ (git-ignored, travels with gpurun).
#   baseline/_ref/kd      scalar_rgb + Python bindings, SAH kd-tree traversal (kdtree.h:2188-2290)
#   baseline/_ref/embree  scalar_rgb CLI with the vendored Embree (scene_embree.inl:202-426)   [--embree]
# Flags: gcc-13 header drift needs <cstdint>/<cmath>; /usr/bin/g++ because the /opt/gcc wrapper has no
# lto-wrapper.  The compiler is baseline/cxx_wrap.sh, which appends -ffp-contract=off and undefines the SSE4.2/AVX
# feature macros after the reference's own -march=native -ffp-contract=fast (its header explains why: FMA-exact scalar
# arithmetic like the LLVM variant, and no asmjit struct converter).
# Nothing is copied from the reference into the repository; only build products land in baseline/_ref.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${EPSM_REFERENCE:-/root/reference}
JOBS=${JOBS:-$(nproc)}
common=(-GNinja -DMI_DEFAULT_VARIANTS=scalar_rgb -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release
        "-DCMAKE_CXX_COMPILER=$ROOT/baseline/cxx_wrap.sh" -DCMAKE_C_COMPILER=/usr/bin/gcc
        "-DCMAKE_CXX_FLAGS=-include cstdint -include cmath")
build() {   # dir, extra cmake flags…
  local dir="$ROOT/baseline/_ref/$1"; shift
  mkdir -p "$dir"
  cmake -S "$REF" -B "$dir" "${common[@]}" "$@" > "$dir/cmake.log" 2>&1
  ninja -C "$dir" -k0 -j"$JOBS" > "$dir/ninja.log" 2>&1 || true      # copy-resources fails (absent mitsuba-data): harmless
  test -f "$dir/libmitsuba.so" || { echo "reference build failed, see $dir/ninja.log"; exit 1; }
  # keep what runs, drop the object files (the tree travels to the GPU box on every gpurun call)
  find "$dir" -name '*.o' -delete; find "$dir" -name '*.a' -delete
  find "$dir" -name '*.so' -exec strip --strip-unneeded {} + 2>/dev/null || true
  strip --strip-unneeded "$dir/mitsuba" 2>/dev/null || true
}
if [ "$1" = "--embree" ]; then
  build embree -DMI_ENABLE_PYTHON=OFF -DMI_ENABLE_EMBREE=ON
else
  build kd -DMI_ENABLE_PYTHON=ON -DMI_ENABLE_EMBREE=OFF "-DPython_EXECUTABLE=$(which python3)"
fi
echo "reference built"
# The reference's Python drivers (EPSM/optim.py, utils/, exp/) — installed next to the binaries so that the GPU box, which has
# no /root/reference, can run them UNMODIFIED against the stand-in modules (tests/test_shim_optim.py).  baseline/_ref is
# git-ignored: nothing of this enters the repository's history.
mkdir -p "$ROOT/baseline/_ref/EPSM"
cp -r "$REF/EPSM/optim.py" "$REF/EPSM/optim_human.py" "$REF/EPSM/utils" "$REF/EPSM/exp" "$ROOT/baseline/_ref/EPSM/"
