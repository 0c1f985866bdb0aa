#!/bin/bash
# Compiler wrapper for building the UNMODIFIED reference (see build_ref.sh): appends flags AFTER the ones the reference's
# CMake files add (-march=native -ffp-contract=fast), so that
#   * -mno-sse4.2 → Dr.Jit's packet back-ends and the asmjit struct converter (include/mitsuba/core/struct.h:8-12, stale
#     against ext/asmjit) stay off,
#   * -DDRJIT_X86_FMA=1 → scalar dr::fmadd is a true fused multiply-add (drjit/array_utils.h:201-210; a libm fmaf call,
#     since -mno-sse4.2 also drops the FMA ISA), as in the LLVM JIT variant the parity target names (llvm_ad_rgb),
#   * -ffp-contract=off → GCC fuses nothing else (the JIT emits fma only for explicit fmadd),
#   * -include immintrin.h → the intrinsic headers Dr.Jit pulls in one by one (packet_intrin.h:34-60) find their types,
#   * -march=x86-64-v3 instead of native → the binaries also run on the GPU box's host CPU.
# Embree compiles one translation unit per ISA with its own -mavx/-mavx2/-mavx512 flags and dispatches at run time: it
# must keep them (and has no Dr.Jit code), so its sources pass through untouched.
case "$*" in *ext/embree/*) exec /usr/bin/g++ "$@" ;; esac
exec /usr/bin/g++ "$@" -march=x86-64-v3 -mno-sse4.2 -DDRJIT_X86_FMA=1 -include immintrin.h -ffp-contract=off
