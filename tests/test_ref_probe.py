"""The CPU restatement (oracle/epsm_oracle.cpp) against the REFERENCE's own header code.

tests/golden/ref_probe.npz was produced by compiling oracle/ref_probe.cpp against the headers under /root/reference
(`make -C oracle ref`, generator tests/golden/make_ref_probe.py).  Integer functions must agree bit for bit; the
floating-point ones are restated operation by operation and are held to bit equality as well, with two stated
exceptions: (i) functions built on dr::dot / squared_norm — the probe instantiates plain `float`, whose static-array dot
is SSE dp_ps, while the restatement follows the fmadd chain of the JIT variants (see the note in oracle/ref_probe.cpp) —
agree to 1–2 ulp; (ii) exp / log inside the microfacet distribution are Dr.Jit polynomials there and double-precision
libm here.  Tolerances are |err| ≤ tol · max(|ref|, 1).
"""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_ref_probe", os.path.join(HERE, "golden", "make_ref_probe.py"))
gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)

GOLD = np.load(os.path.join(HERE, "golden", "ref_probe.npz"))
# fn → relative tolerance (0 = bit-exact)
TOL = {0: 0, 2: 0, 3: 1e-6, 4: 0, 5: 0, 6: 0, 7: 1e-6, 8: 2e-7, 9: 0, 10: 0, 11: 3e-7, 12: 3e-7}


def _compare(fn, ref, got):
    name = gen.FUNCS[fn][0]
    if TOL[fn] == 0:
        bad = ref.view(np.uint32) != got.view(np.uint32)
        # +0 / −0 and NaN payloads are not distinguished by any consumer on the path
        bad &= ~((ref == 0) & (got == 0)) & ~(np.isnan(ref) & np.isnan(got))
        assert not bad.any(), f"{name}: {int(bad.sum())} of {bad.size} words differ, first row {np.argwhere(bad)[0]}: ref {ref[bad][0]!r} got {got[bad][0]!r}"
    else:
        assert np.array_equal(np.isnan(ref), np.isnan(got)), f"{name}: NaN pattern differs"
        err = np.abs(ref.astype(np.float64) - got) / np.maximum(np.abs(ref), 1.0)
        assert np.nanmax(err) <= TOL[fn], f"{name}: max err {np.nanmax(err):.3g} > {TOL[fn]}"


@pytest.mark.parametrize("fn", sorted(gen.FUNCS))
def test_restatement_matches_reference_headers_golden(fn):
    x = GOLD[f"in_{fn}"]
    assert np.array_equal(x.view(np.uint32), gen.inputs(fn).view(np.uint32)), "generator inputs drifted from the committed fixture"
    got = gen.run_probe(oracle.lib(), "orc_probe", fn, x, gen.FUNCS[fn][2])
    _compare(fn, GOLD[f"out_{fn}"], got)


def test_sampler_stream_matches_reference_pcg32_and_tea():
    for i, (seed, lb, lanes) in enumerate(gen.SAMPLER_CASES):
        ref = GOLD[f"sampler_{i}"]
        got = oracle.sampler_stream(seed, lb, lb + lanes, gen.SAMPLER_FLOATS)
        assert np.array_equal(ref.view(np.uint32), got.view(np.uint32)), f"case {i}"


@pytest.mark.skipif(not os.path.exists(gen.REF_SO), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")
def test_restatement_matches_live_reference_probe():
    lib = C.CDLL(gen.REF_SO)
    for fn, (_, _, ostride) in gen.FUNCS.items():
        x = gen.inputs(fn, n=4000, seed=99)
        _compare(fn, gen.run_probe(lib, "ref_probe", fn, x, ostride), gen.run_probe(oracle.lib(), "orc_probe", fn, x, ostride))
    ref = gen.ref_sampler(lib, 12345, 777, 128, 40)
    assert np.array_equal(ref.view(np.uint32), oracle.sampler_stream(12345, 777, 777 + 128, 40).view(np.uint32))
