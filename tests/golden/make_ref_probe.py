"""Generates tests/golden/ref_probe.npz from the REFERENCE's own header code.

    make -C oracle ref          # compiles oracle/ref_probe.cpp against /root/reference's headers → oracle/_ref/
    python tests/golden/make_ref_probe.py

For every probe function (numbering in oracle/ref_probe.cpp) the file stores the seeded inputs and the outputs the
reference code produced for them; tests/test_ref_probe.py holds the CPU restatement (oracle/epsm_oracle.cpp) to those
outputs wherever it runs, and to the live library when oracle/_ref/ is present.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libepsm_refprobe.so")

# fn → (name, in_stride, out_stride)
FUNCS = {0: ("tea32", 2, 2), 2: ("sincos", 1, 2), 3: ("cosine_hemisphere", 2, 3), 4: ("coordinate_system", 3, 6),
         5: ("fresnel_dielectric", 2, 4), 6: ("fresnel_conductor", 3, 1), 7: ("microfacet_sample", 8, 4),
         8: ("microfacet_eval", 6, 1), 9: ("microfacet_smith_g1", 9, 1), 10: ("uniform_triangle", 2, 2),
         11: ("reflect", 6, 3), 12: ("frame", 6, 6)}
SAMPLER_CASES = [(0, 0, 64), (1, 1000, 64), (7, 4_000_000_000, 32), (0xffffffff, 123456, 32)]   # (seed_value, lane_begin, lanes)
SAMPLER_FLOATS = 48


def unit(rng, n):
    v = rng.normal(size=(n, 3)); return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def upper(rng, n):
    v = unit(rng, n); v[:, 2] = np.abs(v[:, 2]) + 1e-3; return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def inputs(fn, n=1500, seed=20261017):
    rng = np.random.default_rng(seed + fn)
    f32 = np.float32
    if fn == 0:
        a = rng.integers(0, 2 ** 32, size=(n, 2), dtype=np.uint64).astype(np.uint32)
        a[:8] = [[1, 1], [1, 2], [1, 3], [1, 4], [1, 5], [2, 1], [3, 1], [4, 1]]       # src/core/tests/test_random.py:8-16
        return a.view(f32)
    if fn == 2:
        x = np.concatenate([rng.uniform(-8, 8, n - 200), rng.uniform(-1e-3, 1e-3, 100), rng.uniform(-200, 200, 100)])
        return x.astype(f32)[:, None]
    if fn in (3, 10):
        s = rng.random((n, 2)).astype(f32); s[0] = [0.5, 0.5]; s[1] = [0, 0]; s[2] = [0.999999, 0.25]; return s
    if fn == 4:
        v = unit(rng, n); v[0] = [0, 0, 1]; v[1] = [0, 0, -1]; v[2] = [1, 0, 0]; return v
    if fn == 5:
        c = rng.uniform(-1, 1, n); eta = rng.choice([1.5, 1.33, 1.0, 2.4, 1 / 1.5], n)
        c[:3] = [0, 1, -1]
        return np.stack([c, eta], 1).astype(f32)
    if fn == 6:
        return np.stack([rng.uniform(0, 1, n), rng.uniform(0.1, 3, n), rng.uniform(0.5, 7, n)], 1).astype(f32)
    if fn in (7, 8, 9):
        au = rng.choice([0.05, 0.1, 0.3, 0.7], n); av = np.where(rng.random(n) < 0.7, au, rng.choice([0.05, 0.2, 0.5], n))
        ggx = (rng.random(n) < 0.5).astype(np.float64)
        cols = [au, av, ggx]
        if fn == 7:
            arr = np.concatenate([np.stack(cols, 1), upper(rng, n), rng.random((n, 2))], 1)
        elif fn == 8:
            arr = np.concatenate([np.stack(cols, 1), upper(rng, n)], 1)
        else:
            arr = np.concatenate([np.stack(cols, 1), unit(rng, n), upper(rng, n)], 1)
        return arr.astype(f32)
    if fn == 11:
        return np.concatenate([upper(rng, n), upper(rng, n)], 1).astype(f32)
    if fn == 12:
        return np.concatenate([unit(rng, n), rng.normal(size=(n, 3))], 1).astype(f32)
    raise KeyError(fn)


def run_probe(lib, symbol, fn, x, out_stride):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros((x.shape[0], out_stride), np.float32)
    f = getattr(lib, symbol)
    f.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    rc = f(fn, x.shape[0], x.ctypes.data, x.shape[1], out.ctypes.data, out_stride)
    assert rc == 0, (symbol, fn, rc)
    return out


def ref_sampler(lib, seed_value, lane_begin, lanes, n_floats):
    out = np.zeros((lanes, n_floats), np.float32)
    lib.ref_sampler_stream.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    assert lib.ref_sampler_stream(seed_value, lane_begin, lanes, n_floats, out.ctypes.data) == 0
    return out


def main():
    if not os.path.exists(REF_SO):
        sys.exit(f"{REF_SO} missing: run `make -C oracle ref` (needs /root/reference)")
    lib = C.CDLL(REF_SO)
    data = {}
    for fn, (name, _, ostride) in FUNCS.items():
        x = inputs(fn)
        data[f"in_{fn}"] = x
        data[f"out_{fn}"] = run_probe(lib, "ref_probe", fn, x, ostride)
    for i, (seed, lb, lanes) in enumerate(SAMPLER_CASES):
        data[f"sampler_{i}"] = ref_sampler(lib, seed, lb, lanes, SAMPLER_FLOATS)
    np.savez_compressed(os.path.join(HERE, "ref_probe.npz"), **data)
    print("wrote ref_probe.npz:", {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
