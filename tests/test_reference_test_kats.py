"""Known answers held by the reference's OWN tests for pieces of the path, applied to the CPU oracle — the
values and identities are those the reference asserts, with the file:line of each assertion:

  src/bsdfs/tests/test_dielectric.py:31-100,138-159   normal-incidence samples, the 80° spot check
  src/bsdfs/tests/test_diffuse.py:13-36               eval = ρ cos θ / π, pdf = cos θ / π on 20 angles
  src/bsdfs/tests/test_twosided.py:29-46,48-93        two-sided pdf / sample–eval–pdf consistency
  src/sensors/tests/test_perspective.py:57-133        ray origin on the near plane, centre ray, differentials
  src/emitters/tests/test_area.py                     area-light direction sample: pdf converts area → solid angle

(the TEA / PCG32 / rectangle-hit vectors of the same section live in test_oracle_kats.py).  Radiance transport mode
throughout: the manifold integrators never use importance mode (epsm.py builds a default BSDFContext)."""
import numpy as np
import pytest

from epsm_mitsuba3_b200 import capi, scenes as S
from tests.hostsim.hostsim import SimScene

DELTA_REFLECTION, DELTA_TRANSMISSION = 0x20, 0x40        # EPSM_BF_DELTA_REFLECTION / _TRANSMISSION = mi.BSDFFlags values


def example_dielectric():                      # test_dielectric.py:5-12
    return S.Bsdf(capi.BSDF_DIELECTRIC, reflectance=(0.3,) * 3, transmittance=(0.6,) * 3, eta=(np.float32(1.5), 0.0, 0.0))


class DeviceBsdf:
    """The PRODUCT's bsdf_sample / bsdf_eval_pdf (epsm_device.cuh, compiled for the host by the test-only harness) behind the
    oracle's call signature, so that every known answer below is asserted on both."""

    def __init__(self):
        self.cache = {}

    def _scene(self, b):
        key = (b.kind, b.reflectance, b.transmittance, tuple(float(x) for x in b.eta), b.k, b.alpha_u, b.alpha_v, b.twosided, b.ggx)
        if key not in self.cache:
            sd = S.SceneData(); sd.add_bsdf(b); sd.add_mesh(S.rectangle(S.Transform4f()))
            cam = S.Sensor(S.Transform4f.look_at([0, 0, 3], [0, 0, 0], [0, 1, 0]), 40.0, 8, 8)
            sd.sensors = [cam, cam, cam]
            self.cache[key] = SimScene(sd)
        return self.cache[key]

    def bsdf_sample(self, b, wi, s1, s2):
        return self._scene(b).bsdf_sample(0, wi, s1, s2)

    def bsdf_eval_pdf(self, b, wi, wo):
        return self._scene(b).bsdf_eval_pdf(0, wi, wo)


@pytest.fixture(params=["oracle", "device-code"])
def impl(request, orc):
    return orc if request.param == "oracle" else DeviceBsdf()


def test_dielectric_normal_incidence(impl):
    orc = impl     # test_dielectric.py:31-100 (i = 1: Radiance)
    b = example_dielectric()
    bs = orc.bsdf_sample(b, [0, 0, 1], 0.0, [0, 0])
    assert np.allclose(bs["weight"], 0.3) and np.isclose(bs["pdf"], 0.04) and np.isclose(bs["eta"], 1.0)
    assert np.allclose(bs["wo"], [0, 0, 1]) and bs["type"] == DELTA_REFLECTION
    bs = orc.bsdf_sample(b, [0, 0, 1], 0.05, [0, 0])
    assert np.allclose(bs["weight"], 0.6 / 1.5 ** 2) and np.isclose(bs["pdf"], 1 - 0.04) and np.isclose(bs["eta"], 1.5)
    assert np.allclose(bs["wo"], [0, 0, -1]) and bs["type"] == DELTA_TRANSMISSION
    # from the inside (test03_sample_reverse)
    bs = orc.bsdf_sample(b, [0, 0, -1], 0.0, [0, 0])
    assert np.allclose(bs["weight"], 0.3) and np.isclose(bs["pdf"], 0.04) and np.isclose(bs["eta"], 1.0)
    assert np.allclose(bs["wo"], [0, 0, -1]) and bs["type"] == DELTA_REFLECTION
    bs = orc.bsdf_sample(b, [0, 0, -1], 0.05, [0, 0])
    assert np.allclose(bs["weight"], 0.6 * 1.5 ** 2) and np.isclose(bs["pdf"], 1 - 0.04) and np.isclose(bs["eta"], 1 / 1.5)
    assert np.allclose(bs["wo"], [0, 0, 1]) and bs["type"] == DELTA_TRANSMISSION


def test_dielectric_spot_check_80_degrees(impl):
    orc = impl   # test_dielectric.py:138-159
    b = example_dielectric()
    a = 80 * np.pi / 180
    wi = np.array([np.sin(a), 0, np.cos(a)], np.float32)
    bs = orc.bsdf_sample(b, wi, 0.0, [0, 0])
    assert np.isclose(bs["pdf"], 0.387704354691473, rtol=1e-5)
    assert np.allclose(bs["wo"], [-np.sin(a), 0, np.cos(a)], atol=1e-6)
    bs = orc.bsdf_sample(b, wi, 1.0, [0, 0])
    t = 41.03641052520335 * np.pi / 180
    assert np.isclose(bs["pdf"], 1 - 0.387704354691473, rtol=1e-5)
    assert np.allclose(bs["wo"], [-np.sin(t), 0, -np.cos(t)], atol=1e-6)
    back = orc.bsdf_sample(b, bs["wo"], 1.0, [0, 0])
    assert np.isclose(back["pdf"], 1 - 0.387704354691473, rtol=1e-5)
    assert np.allclose(back["wo"], wi, atol=1e-6)


def test_diffuse_eval_pdf(impl):
    orc = impl                # test_diffuse.py:13-36 (default reflectance 0.5)
    b = S.Bsdf.diffuse((0.5, 0.5, 0.5))
    for i in range(20):
        th = i / 19.0 * (np.pi / 2)
        wo = np.array([np.sin(th), 0, np.cos(th)], np.float32)
        value, pdf = orc.bsdf_eval_pdf(b, [0, 0, 1], wo)
        assert np.isclose(pdf, wo[2] / np.pi, rtol=1e-6, atol=1e-8)
        assert np.allclose(value, 0.5 * wo[2] / np.pi, rtol=1e-6, atol=1e-8)


def test_twosided_pdf_and_consistency(impl):
    orc = impl    # test_twosided.py:29-46, 48-93 (same BSDF on both sides)
    b = S.Bsdf.diffuse((0.5, 0.5, 0.5), twosided=True)
    assert np.isclose(orc.bsdf_eval_pdf(b, [0, 0, 1], [0, 0, 1])[1], 1 / np.pi, rtol=1e-6)
    assert orc.bsdf_eval_pdf(b, [0, 0, 1], [0, 0, -1])[1] == 0.0
    n = 5
    for u in range(n):
        for v in range(n):
            # warp.square_to_uniform_sphere (warp.h): z = 1 − 2 v, r = sqrt(1 − z²), φ = 2π u
            z = 1 - 2 * v / (n - 1); r = np.sqrt(max(0.0, 1 - z * z)); ph = 2 * np.pi * u / (n - 1)
            wi = np.array([r * np.cos(ph), r * np.sin(ph), z], np.float32)
            if abs(wi[2]) < 1e-6:
                continue
            up = wi[2] > 0
            for x in range(n):
                for y in range(n):
                    bs = orc.bsdf_sample(b, wi, 0.5, [x / (n - 1), y / (n - 1)])
                    if not np.any(bs["weight"] > 0):
                        continue
                    s_value = bs["weight"] * bs["wo"][2] / np.pi * (1 if up else -1)
                    e_value, p_pdf = orc.bsdf_eval_pdf(b, wi, bs["wo"])
                    assert np.allclose(s_value, e_value, atol=1e-2)
                    assert np.isclose(bs["pdf"], p_pdf, rtol=1e-5, atol=1e-7)
                    assert not np.isnan(e_value).any()


@pytest.mark.parametrize("origin", [[1.0, 0.0, 1.5], [1.0, 4.0, 1.5]])
@pytest.mark.parametrize("direction", [[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
def test_perspective_sample_ray_differential(orc, origin, direction):    # test_perspective.py:57-133
    near = 1.0
    target = [origin[k] + direction[k] for k in range(3)]
    sd = S.SceneData()
    sd.add_bsdf(S.Bsdf.diffuse((0.5, 0.5, 0.5)))
    sd.add_mesh(S.rectangle(S.Transform4f().translate([0, 0, 100.0])))
    cam = S.Sensor(S.Transform4f.look_at(origin, target, [0, 1, 0]), 34.0, 512, 256, near_clip=near, far_clip=35.0)
    sd.sensors = [cam, cam, cam]
    o = orc.OracleScene(sd)
    to_world = S.Transform4f.look_at(origin, target, [0, 1, 0]).matrix
    for pos in ([0.2, 0.6], [0.1, 0.9], [0.5, 0.5]):
        r = o.camera_ray(0, pos[0], pos[1])
        d_local = np.linalg.inv(to_world)[:3, :3] @ r["d"].astype(np.float64)
        assert np.allclose(r["o"], np.asarray(origin) + near / d_local[2] * r["d"], atol=1e-4)      # :77-79
        assert np.isclose(np.dot(r["d_x"] - r["d"], r["d_y"] - r["d"]), 0, atol=1e-7)               # :113
    c = o.camera_ray(0, 0.5, 0.5)
    assert np.allclose(c["d"], direction, atol=1e-7)                                                # :83-84
    dx, dy = 1.0 / 512, 1.0 / 256
    assert np.allclose(o.camera_ray(0, 0.5 + dx, 0.5)["d"], c["d_x"], atol=1e-7)                    # :124-133
    assert np.allclose(o.camera_ray(0, 0.5, 0.5 + dy)["d"], c["d_y"], atol=1e-7)


def test_area_emitter_direction_sample(orc):   # area.cpp:82-200 / test_area.py test04: pdf_area · dist² / cos θ
    sd = S.mirror_plane_scene(16, 2, 8)
    o = orc.OracleScene(sd)
    em = next(m for m in sd.meshes if m.emitter >= 0)
    area = 0.0
    p = em.positions.astype(np.float64); f = em.faces
    for t in f:
        area += 0.5 * np.linalg.norm(np.cross(p[t[1]] - p[t[0]], p[t[2]] - p[t[0]]))
    n_checked = 0
    for (u, v) in ((0.5, 0.6), (0.3, 0.7), (0.7, 0.8)):
        ray = o.camera_ray(1, u, v)
        si = o.surface_interaction(ray["o"], ray["d"])
        if not np.isfinite(si["t"]):
            continue
        for smp in ([0.3, 0.6], [0.9, 0.2]):
            ds = o.sample_emitter(ray["o"], ray["d"], smp)
            if ds["pdf"] == 0:
                continue
            d = ds["p"].astype(np.float64) - si["p"].astype(np.float64)
            assert np.isclose(np.linalg.norm(d), ds["dist"], rtol=1e-4)
            assert np.allclose(d / np.linalg.norm(d), ds["d"], atol=1e-5)
            cos_l = abs(np.dot(ds["n"].astype(np.float64), ds["d"].astype(np.float64)))
            assert np.isclose(ds["pdf"], ds["dist"] ** 2 / (cos_l * area), rtol=1e-4)       # shape.cpp:352-383: area → solid angle
            n_checked += 1
    assert n_checked > 0


def test_imageblock_put_against_brute_force(orc):     # src/render/tests/test_imageblock.py:37-120
    """ImageBlock::put against the brute-force reference the reference's test builds: gaussian filter, no border (HDRFilm crops
    it, hdrfilm.cpp:286-320), offset (0, 0), normalize False — the configuration EPSMIntegrator.render uses."""
    o = orc.OracleScene(S.cornell_box(8, 1, 8))
    for j in range(5):
        for i in range(5):
            pos = (np.float32(3.3 + 0.25 * i), np.float32(3 + 0.25 * j))      # one of the points is exactly on a boundary
            film = np.zeros((6, 6, 4), np.float32)
            o.film_put(film, pos, [1.0, 0.5, 0.25])
            fx = np.array([o.filter_eval(float(np.float32(x + 0.5) - pos[0])) for x in range(6)])
            fy = np.array([o.filter_eval(float(np.float32(y + 0.5) - pos[1])) for y in range(6)])
            ref = np.outer(fy, fx)
            assert np.allclose(film[..., 3], ref, atol=1e-5), (i, j)
            assert np.allclose(film[..., 0], ref, atol=1e-5) and np.allclose(film[..., 2], 0.25 * ref, atol=1e-5)
    # the filter itself: gaussian, stddev 0.5, zero at the radius 2 (gaussian.cpp:48-103)
    assert abs(o.filter_eval(2.0)) < 1e-5 and o.filter_eval(-2.5) == 0.0          # max(poly, 0): the radius is where the polynomial crosses zero
    g = lambda x: np.exp(-x * x / (2 * 0.25)) - np.exp(-4 / (2 * 0.25))
    for x in (0.0, 0.3, 1.0, 1.7):
        assert abs(o.filter_eval(x) - g(x)) < 1e-3          # degree-9 Remez fit of the Gaussian (gaussian.cpp:57-89)
