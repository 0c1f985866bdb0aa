"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol include/epsm_b200.h declares
(no compute calls without a GPU), and fails loudly instead of falling back when there is no device."""
import ctypes as C
import os
import re

import pytest
import torch

import __graft_entry__ as ge
from epsm_mitsuba3_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "epsm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(epsm_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _declared() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    ge.build_library()
    lib = C.CDLL(capi.library_path())
    for name in _declared():
        assert hasattr(lib, name), name


def test_record_layout_constants_match_header():
    txt = open(os.path.join(ROOT, "include", "epsm_b200.h")).read()
    assert int(re.search(r"#define EPSM_REC_MAXV\s+(\d+)", txt).group(1)) == capi.REC_MAXV
    assert int(re.search(r"#define EPSM_REC_VSTRIDE\s+(\d+)", txt).group(1)) == capi.REC_VSTRIDE
    assert int(re.search(r"#define EPSM_REC_GSTRIDE\s+(\d+)", txt).group(1)) == capi.REC_GSTRIDE
    assert capi.REC_FLOATS == capi.REC_MAXV * (capi.REC_VSTRIDE + capi.REC_GSTRIDE) + 14


def test_descriptor_layouts_match_header(tmp_path):
    """sizeof / offsetof of every descriptor struct as the C compiler sees include/epsm_b200.h against the ctypes mirrors: a
    mismatch would make epsm_scene_upload read garbage past the first array element without any error."""
    import subprocess
    structs = {"epsm_mesh_desc": capi.MeshDesc, "epsm_bsdf_desc": capi.BsdfDesc, "epsm_texture_desc": capi.TextureDesc,
               "epsm_emitter_desc": capi.EmitterDesc, "epsm_sensor_desc": capi.SensorDesc, "epsm_reparam_config": capi.ReparamConfig,
               "epsm_scene_stats": capi.SceneStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "epsm_b200.h")}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"; src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import epsm_mitsuba3_b200 as ep
    from epsm_mitsuba3_b200 import scenes
    with pytest.raises(ep.EpsmError):
        ep.Scene(scenes.cornell_box(8, 1, 8))
    lib = capi.load_library()
    ctx = C.c_void_p()
    assert lib.epsm_create(0, C.byref(ctx)) == -4          # EPSM_E_NOGPU
    assert b"no CPU fallback" in lib.epsm_last_error(None)


def test_integrator_plugin_interface():
    import epsm_mitsuba3_b200 as ep
    m = ep.load_dict({"type": "manifold"})
    assert (m.max_depth, m.rr_depth, m.variant) == (6, 5, "manifold")                  # defaults of common.py:28-43
    c = ep.load_dict({"type": "manifold_caustic", "max_depth": -1, "rr_depth": 3})
    assert c.max_depth == 0xffffffff and c.rr_depth == 3 and c.variant == "manifold_caustic"
    with pytest.raises(Exception):
        ep.load_dict({"type": "manifold", "rr_depth": 0})
    with pytest.raises(ep.EpsmError):
        ep.load_dict({"type": "manifold_shadow"})          # registered nowhere in the reference either (EPSM/all.sh:8)


def test_seed_grad_derivation(orc, kats):
    # seed_grad = sample_tea_32(seed, 1)[0] (util.py:510-515); the host restatement must equal the pinned oracle TEA
    import epsm_mitsuba3_b200 as ep
    for v0 in (0, 1, 2, 12345, 0xffffffff):
        assert ep.sample_tea_32(v0, 1) == orc.tea32(v0, 1, 4)


def test_hybrid_schedule_follows_optim_py():
    # EPSM/optim.py:87-119: suffix strip, sensor choice, optimiser reset exactly at it == thres, phase-2 plugin lookup
    import epsm_mitsuba3_b200 as ep
    s = ep.HybridSchedule("manifold_caustic_hybrid", max_depth=4, thres=3)
    assert (s.method, s.thres, s.integrator1.variant, s.integrator1.max_depth) == ("manifold_caustic", 3, "manifold_caustic", 4)
    p = s.phase(2)
    assert (p.integrator_type, p.sensor, p.loss, p.reset_optimizer) == ("manifold_caustic", 1, "matcher", False)
    p2 = s.phase(3)                                         # phase 2: prb_reparam on sensor 0, L2 loss, optimiser reset (optim.py:113-119)
    assert (p2.integrator_type, p2.sensor, p2.loss, p2.reset_optimizer) == ("prb_reparam", 0, "l2", True)
    assert (p2.integrator.max_depth, p2.integrator.reparam_rays, p2.integrator.reparam_kappa, p2.integrator.reparam_exp) == (4, 16, 1e5, 3.0)
    assert p2.integrator.reparam_max_depth == 4             # prb_reparam.py:232: defaults to max_depth
    made = []
    s2 = ep.HybridSchedule("manifold_hybrid", 6, thres=5, phase2_factory=lambda props: made.append(props) or "prb")
    assert s2.phase(5).reset_optimizer and not s2.phase(6).reset_optimizer and s2.phase(6).sensor == 0 and s2.phase(6).loss == "l2"
    assert made == [{"type": "prb_reparam", "max_depth": 6}]
    assert ep.HybridSchedule("manifold", 6, thres=5).thres == 10000      # non-hybrid: never switches (optim.py:95)
