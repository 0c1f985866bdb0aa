"""EPSM/optim.py, UNMODIFIED, against the stand-in `mitsuba` / `drjit` / `geomloss` modules (BASELINE.json north_star:
"so EPSM/optim.py runs unmodified against it").  The driver and its `utils/` come from the reference install
under baseline/_ref/EPSM (baseline/build_ref.sh copies them there; git-ignored, travels to the GPU box); the experiment
module is tests/exp_synth/synth_shadow.py.  Runs as a subprocess, like `python optim.py <method> <config>` (EPSM/test.sh)."""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "baseline", "_ref", "EPSM")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(DRIVER, "optim.py")), reason="reference drivers not installed (baseline/build_ref.sh)")]


def _workdir(tmp_path):
    from epsm_mitsuba3_b200.scenes import icosphere
    (tmp_path / "exp").mkdir()
    shutil.copy(os.path.join(ROOT, "tests", "exp_synth", "synth_shadow.py"), tmp_path / "exp" / "synth_shadow.py")
    shutil.copytree(os.path.join(DRIVER, "utils"), tmp_path / "utils")
    (tmp_path / "data" / "meshes").mkdir(parents=True)
    v, f = icosphere(2)
    with open(tmp_path / "data" / "meshes" / "sphere.obj", "w") as fh:
        for p in v:
            fh.write("v %f %f %f\n" % tuple(p))
        for t in f:
            fh.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))
    return tmp_path


def _run(tmp_path, method):
    import epsm_mitsuba3_b200.shim as shim
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([shim.SHIM_DIR, ROOT, env.get("PYTHONPATH", "")])
    env["EPSM_SHIM_TRACE"] = str(tmp_path / "trace.jsonl")
    # runpy executes the reference's source as __main__, exactly as `python optim.py <method> <config>` would; the one
    # environment patch is for this image's headless OpenCV build, whose destroyAllWindows() (utils/logger.py:136) throws.
    boot = ("import sys, runpy, cv2; cv2.destroyAllWindows = lambda: None; "
            f"sys.argv = [{os.path.join(DRIVER, 'optim.py')!r}, {method!r}, 'synth_shadow']; "
            f"runpy.run_path({os.path.join(DRIVER, 'optim.py')!r}, run_name='__main__')")
    r = subprocess.run([sys.executable, "-c", boot], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert "finish optim" in r.stdout
    with open(tmp_path / "trace.jsonl") as f:
        return [json.loads(l) for l in f], r


@pytest.mark.parametrize("method", ["manifold", "manifold_hybrid"])
def test_reference_optim_py_runs_unmodified(tmp_path, method):
    trace, r = _run(_workdir(tmp_path), method)
    renders = [t for t in trace if t["op"] == "render"]
    steps = [t for t in trace if t["op"] == "backward"]
    # optim.py:56 target (8192 spp, sensor 0), :80 init (512 spp), 4 iterations, :160 final
    assert [t["spp"] for t in renders] == [8192, 512, 4, 4, 4, 4, 8192]
    if method == "manifold":
        assert [t["integrator"] for t in renders[2:6]] == ["manifold"] * 4 and [t["sensor"] for t in renders[2:6]] == [1] * 4
        assert [t["integrator"] for t in steps] == ["manifold"] * 4
    else:                                               # optim.py:87-119: manifold on sensor 1 until thres, then prb_reparam on sensor 0
        assert [t["integrator"] for t in renders[2:6]] == ["manifold", "manifold", "prb_reparam", "prb_reparam"]
        assert [t["sensor"] for t in renders[2:6]] == [1, 1, 0, 0]
        assert [t["integrator"] for t in steps] == ["manifold", "manifold", "prb_reparam", "prb_reparam"]
    for t in steps:                                     # every iteration delivered a finite, non-zero gradient to the vertices
        assert np.isfinite(t["grad_abs_sum"]) and t["grad_abs_sum"] > 0, t
    out = tmp_path / "results" / "synth_shadow" / method
    assert (out / "gt_img.png").exists() and (out / "init_img.png").exists() and (out / f"{method}.png").exists()
