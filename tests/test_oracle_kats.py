"""Pins the CPU oracle against the reference's own known-answer values (SURVEY.md §8c, Appendix C rows 1-4,6)."""
import numpy as np

from epsm_mitsuba3_b200 import scenes as S


def test_tea_float32(orc, kats):
    # src/core/tests/test_random.py:8-16 — bit-exact
    for v0, v1, ref in kats["tea_float32"]:
        assert orc.tea_float32(v0, v1, 4) == np.float32(ref)


def test_tea_float64(orc, kats):
    # src/core/tests/test_random.py:19-27 pins both output words (random.h:111-116, :158-163)
    for v0, v1, ref in kats["tea_float64"]:
        a, b = orc.tea32(v0, v1, 4)
        u = ((a + (b << 32)) >> 12) | 0x3ff0000000000000
        assert np.array([u], np.uint64).view(np.float64)[0] - 1.0 == ref


def test_pcg32_default_stream(orc, kats):
    # src/samplers/tests/test_independent.py:17-29: scalar sampler.seed(0) == PCG32(initstate=0)
    got = orc.pcg32_stream(0, 0xda3e39cb94b95bdb, 3)
    assert np.allclose(got, kats["pcg32_seed0"], rtol=0, atol=1e-9)


def test_pcg32_known_first_outputs(orc):
    # PCG32 reference implementation demo values (pcg32_srandom(42, 54)): 0xa15c02b7 0x7b47f409 0xba1d3330
    # as floats via drjit/random.h:130-132
    exp = [((u >> 9) | 0x3f800000) for u in (0xa15c02b7, 0x7b47f409, 0xba1d3330)]
    exp = np.array(exp, np.uint32).view(np.float32) - np.float32(1)
    assert np.array_equal(orc.pcg32_stream(42, 54, 3), exp)


def test_sampler_stream_is_tea_seeded_pcg(orc):
    # sampler.cpp:119-130: lane i uses PCG32(initstate=v0, initseq=v1) with (v0,v1) = TEA4(seed, i)
    s = orc.sampler_stream(7, 0, 16, 8)
    for lane in range(16):
        v0, v1 = orc.tea32(7, lane, 4)
        assert np.array_equal(s[lane], orc.pcg32_stream(v0, v1, 8))


def _rect_scene(k):
    sd = S.SceneData()
    b = sd.add_bsdf(S.Bsdf.diffuse((0.5, 0.5, 0.5)))
    sd.add_mesh(S.Mesh("rect", np.array(k["vertices"], np.float32), np.array(k["faces"], np.uint32), bsdf=b))
    sd.sensors = [S.Sensor(S.T.look_at([0, 0, 5], [0, 0, 0], [0, 1, 0]), 40, 8, 8)]
    return sd


def test_rectangle_mesh_hit_record(orc, kats):
    # src/render/tests/test_mesh.py:257-291
    k = kats["rectangle_mesh"]
    o = orc.OracleScene(_rect_scene(k))
    for r in k["rays"]:
        si = o.surface_interaction(r["o"], r["d"])
        assert np.isclose(si["t"], r["t"], rtol=1e-6)
        assert si["prim"] == r["prim"]
        assert np.allclose(si["uv"], r["prim_uv"], atol=1e-6)
        assert np.allclose(si["p"], r["p"], atol=1e-6)


def test_rectangle_mesh_texture_coordinates(orc, kats):
    """The rest of src/render/tests/test_mesh.py:257-291: rectangle.obj carries uv = (xy + 1)/2, and the reference asserts
    si.uv = (0.35, 0.35) / (0.65, 0.65), si.dp_du = (2, 0, 0), si.dp_dv = (0, 2, 0) at the two hits."""
    from tests.hostsim.hostsim import SimScene
    k = kats["rectangle_mesh"]
    sd = _rect_scene(k)
    sd.meshes[0].texcoords = ((np.array(k["vertices"], np.float32)[:, :2] + 1) / 2).astype(np.float32)
    o, sim = orc.OracleScene(sd), SimScene(sd)
    for r, uv in zip(k["rays"], ([0.35, 0.35], [0.65, 0.65])):
        s = o.surface_uv(r["o"], r["d"])
        assert np.allclose(s["uv"], uv, atol=1e-6)
        assert np.allclose(s["dp_du"], [2, 0, 0], atol=1e-5) and np.allclose(s["dp_dv"], [0, 2, 0], atol=1e-5)
        assert np.allclose(s["sh_s"], [1, 0, 0], atol=1e-6)                       # initialize_sh_frame: s follows dp_du
        p = sim.surface_uv(r["o"], r["d"])                                        # the product's device code on the host
        assert np.allclose(p["uv"], uv, atol=1e-6) and np.allclose(p["sh_s"], [1, 0, 0], atol=1e-6)


def test_cornell_box_surface_interaction(orc, kats):
    k = kats["cornell_box"]
    sd = S.cornell_box(k["film"], 4, 64)
    o = orc.OracleScene(sd)
    r = k["ray_mesh"]
    si = o.surface_interaction(r["o"], r["d"])
    names = [m.name for m in sd.meshes]
    first = int(sd.face_offsets()[names.index(r["shape"])])
    assert si["prim"] - first == r["prim_index_local"]
    assert np.isclose(si["t"], r["t"], rtol=1e-6)
    assert np.allclose(si["uv"], r["prim_uv"], atol=2e-6)
    assert si["ismesh"] == 1
    assert np.isclose(si["b0"], r["b0"], atol=2e-6) and np.isclose(si["b1"], r["b1"], atol=2e-6)
    assert np.allclose(si["p0"], r["p0"], atol=2e-6)
    assert np.allclose(si["n0"], r["n0"], atol=2e-6)
    r = k["ray_analytic"]
    si = o.surface_interaction(r["o"], r["d"])
    assert np.isclose(si["t"], r["t"], rtol=1e-6)
    assert si["ismesh"] == 0 and si["b0"] == 0 and np.all(si["p0"] == 0)
    assert si["bsdf_flags"] == r["bsdf_flags"]


def test_bsdf_samples(orc, kats):
    k = kats["cornell_box"]["diffuse_sample"]
    bs = orc.bsdf_sample(S.Bsdf.diffuse((0.5, 0.5, 0.5)), k["wi"], k["sample1"], k["sample2"])
    assert np.allclose(bs["wo"], k["wo"], atol=1e-7)
    assert np.isclose(bs["pdf"], k["pdf"], rtol=1e-6)
    assert np.all(bs["hf"] == 0)
    k = kats["cornell_box"]["roughconductor_sample"]
    bs = orc.bsdf_sample(S.Bsdf.roughconductor(k["alpha"]), k["wi"], k["sample1"], k["sample2"])
    assert np.allclose(bs["hf"], k["hf"], atol=1e-7)


def test_emitter_and_sensor(orc, kats):
    k = kats["cornell_box"]
    o = orc.OracleScene(S.cornell_box(k["film"], 4, 64))
    e = k["sample_emitter_direction"]; r = k[e["ray"]]
    ds = o.sample_emitter(r["o"], r["d"], e["sample"])
    assert np.allclose(ds["p"], e["p"], atol=1e-6)
    assert np.isclose(ds["pdf"], e["pdf"], rtol=2e-5)
    s = k["sample_ray_differential"]
    cr = o.camera_ray(s["sensor"], *s["position_sample"])
    assert np.allclose(cr["o"], s["o"], atol=2e-7)
    assert np.allclose(cr["d"], s["d"], atol=2e-7)
    assert np.allclose(cr["d_x"], s["d_x"], atol=2e-7)
    assert np.isclose(cr["maxt"], s["maxt"], rtol=1e-6)


def test_sincos_restatement(orc):
    xs = np.linspace(-7, 7, 2001).astype(np.float32)
    got = np.array([orc.sincos(float(x)) for x in xs])
    assert np.max(np.abs(got[:, 0] - np.sin(xs.astype(np.float64)))) < 2e-7
    assert np.max(np.abs(got[:, 1] - np.cos(xs.astype(np.float64)))) < 2e-7


def test_filter_polynomial(orc):
    # gaussian.cpp:57-96: Remez fit to exp(-x²/(2σ²)) minus its value at the radius; zero at |x| = 2
    o = orc.OracleScene(S.cornell_box(16, 1, 8))
    xs = np.linspace(0, 2, 41)
    ref = np.exp(-xs ** 2 / (2 * 0.25)) - np.exp(-4 / (2 * 0.25))
    got = np.array([o.filter_eval(float(x)) for x in xs])
    assert np.max(np.abs(got - ref)) < 2e-3
    assert abs(o.filter_eval(2.0)) < 1e-5


def test_bvh_accel_matches_brute_force(orc):
    sd = S.glass_sphere_caustic(16, 1, 8, subdiv=2)
    o = orc.OracleScene(sd)
    rng = np.random.default_rng(3)
    n = 4000
    org = rng.uniform(-3, 3, (n, 3)).astype(np.float32); org[:, 2] = rng.uniform(0.2, 4, n)
    d = rng.normal(0, 1, (n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32); rays[:, :3] = org; rays[:, 3] = 3.4e38; rays[:, 4:7] = d
    a = o.trace(rays); a_any = o.trace(rays, any_hit=True)
    o.set_accel(True)
    b = o.trace(rays); b_any = o.trace(rays, any_hit=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(a_any, b_any)
