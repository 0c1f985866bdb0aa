"""calc_grad pinned to the reference itself.

tests/golden/calc_grad_reference.npz holds what the REFERENCE'S OWN `calc_grad` source (epsm.py:744-946, :951-1200,
exec'd on CPU torch by tests/golden/make_calc_grad_vectors.py — nothing of it is copied here) returns for the path logs
of the parity scenes, in float32 and float64.

  * not-gpu: oracle/calc_grad_ref.py (the restatement every other parity test uses) must reproduce those vectors:
    float64 to 1e-12 relative, float32 to float32 rounding; where /root/reference is present the reference is executed
    again, live, and must reproduce the fixture exactly (the fixture is regenerable, not hand-made);
  * gpu: the CUDA calc_grad (through epsm_path_records, C ABI) is held to the reference's vectors directly, at the
    north_star tolerance of 1e-4 relative."""
import os

import numpy as np
import pytest
import torch

from tests import parity
from tests.cases import CASES
from tests.golden import make_calc_grad_vectors as gen
from oracle.calc_grad_ref import calc_grad

FIX = np.load(gen.FIXTURE)


def _case(name, orc):
    paths = np.frombuffer(FIX[f"{name}/paths"].tobytes(), dtype=orc.path_dtype())
    _, variant, max_depth = CASES[name]
    return paths, variant, max_depth, parity.n_logged(max_depth), [int(b) for b in FIX[f"{name}/begins"]]


def _lists(name, tag):
    return [list(FIX[f"{name}/{tag}/{k}"]) for k in ("pg", "lg", "dg")]


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_reproduces_the_reference_calc_grad(name, orc):
    paths, variant, _, nl, _ = _case(name, orc)
    for tag, dt, tol in (("f64", torch.float64, 1e-12), ("f32", torch.float32, 2e-5)):
        ref = _lists(name, tag)
        got = calc_grad(paths, variant, nl, dt, keep_dtype=True)
        n_nonzero = 0
        for rl, gl, what in zip(ref, got, ("path_grad", "light_grad", "diffuse_grad")):
            assert len(rl) == len(gl), f"{what}: the reference returns {len(rl)} entries, the restatement {len(gl)}"
            for i, (r, g) in enumerate(zip(rl, gl)):
                assert r.shape == g.shape and r.dtype == g.dtype
                # zero pattern (masks, nan_to_num, ±0.1 outlier filter) must agree except within rounding of the threshold
                edge = np.abs(np.abs(r) - 0.1) < (1e-9 if dt == torch.float64 else 1e-5)
                assert np.array_equal((r == 0) | edge, (g == 0) | edge), f"{what}[{i}] {tag}: mask pattern differs"
                scale = np.maximum(np.abs(r).max(axis=1, keepdims=True), 1e-3)
                err = np.abs(r - g); err[edge] = 0
                if dt == torch.float32:
                    # float32: both run the same dense float32 inverse; rounding differs only through op order, but
                    # ill-conditioned systems amplify it — hold those to the reference's own float32-vs-float64 error
                    r64 = FIX[f"{name}/f64/{ {'path_grad': 'pg', 'light_grad': 'lg', 'diffuse_grad': 'dg'}[what]}"][i]
                    bound = tol * scale + 2 * np.abs(r - r64)
                else:
                    bound = tol * scale
                assert (err <= bound).all(), f"{what}[{i}] {tag}: max err {err.max():.3g}"
                n_nonzero += int((r != 0).any(1).sum())
        assert n_nonzero > 0, "degenerate fixture: the reference returned no gradient at all"


@pytest.mark.skipif(not gen.reference_available(), reason="needs /root/reference (build container)")
@pytest.mark.parametrize("name", ["mirror", "glass", "stage"])
def test_fixture_is_what_the_reference_returns_here(name, orc):
    """Regenerate from the reference source, live: the committed vectors are the reference's outputs, bit for bit,
    and the oracle's path log (the fixture's inputs) is deterministic."""
    paths, variant, nl, begins = gen.case_paths(name)
    assert np.array_equal(paths.view(np.uint8), FIX[f"{name}/paths"]) and list(FIX[f"{name}/begins"]) == begins
    pg, lg, dg = gen.run_reference(paths, variant, nl, torch.float64)
    for got, key in ((pg, "pg"), (lg, "lg"), (dg, "dg")):
        assert np.array_equal(np.stack(got), FIX[f"{name}/f64/{key}"], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_calc_grad_against_the_reference_vectors(name, orc):
    """The product's per-path gradients (analytic Jacobians + block solve, csrc/epsm_device.cuh) against the outputs of
    the reference's own code on the same paths: 1e-4 relative on the paths whose primitive sequences agree."""
    import epsm_mitsuba3_b200 as ep
    paths, variant, max_depth, nl, begins = _case(name, orc)
    sd = CASES[name][0]()
    scene = ep.Scene(sd, device=0)
    g = parity.synthetic_grad_in(sd.sensors[2].width)
    rec = np.concatenate([scene.path_records(variant, 2, 11, 8, max_depth, 5, g, b, b + gen.WINDOW) for b in begins])
    stats, _ = parity.check_path_records(rec, paths, variant, max_depth, reference_lists=(_lists(name, "f32"), _lists(name, "f64")))
    assert stats["paths_with_gradient"] > 0
    parity.record_stats("calc_grad_vs_reference", name, stats)
