"""Pin the CPU oracle against the reference's OWN known-answer tests.

Numbers below are the literals asserted by the reference's unit tests (file:line cited per test) on
its 6-gaussian fixture (test/gaussian_test_data.py, restated in tests/scenes.py).  No GPU needed.
"""
import numpy as np
import pytest

from oracle import cpu_oracle as orc
from tests.scenes import inverse_sigmoid, reference_fixture


@pytest.fixture(scope="module")
def fx():
    return reference_fixture()


@pytest.fixture(scope="module")
def pg(fx):
    # frustum thresholds of test/test_tile_culling.py:25-27 / test_rasterize.py:24-27
    return orc.project(fx["xyz"], fx["quaternion"], fx["scale"], inverse_sigmoid(fx["opacity"]), fx["rgb"], None,
                       fx["T"], fx["K"], 480, 640, 0.3, 100.0, 10.0)


def test_project_points(pg):
    """test/test_projection.py:24-37 (places=4)"""
    exp_xyz = {(0, 0): 0.6602, (0, 1): -1.1849998, (0, 2): -1.4546999, (1, 0): 3.7595997, (1, 1): 4.5586, (1, 2): 7.2283}
    for (i, j), v in exp_xyz.items():
        assert abs(pg.xyz_cam[i, j] - v) < 5e-5
    exp_uv = {(0, 0): 124.849106, (0, 1): 573.9863, (1, 0): 543.6526, (1, 1): 498.57062}
    for (i, j), v in exp_uv.items():
        assert abs(pg.uv[i, j] - v) < 5e-5 * max(1.0, abs(v) / 100)


def test_cull_pattern(pg):
    """test/test_projection.py:60-65 uses near-only culling inside the image; with the rasterizer's
    thresholds (pad 10, far 100) the same three gaussians survive (test/test_tile_culling.py:29-45)."""
    assert pg.visible.tolist() == [0, 0, 0, 1, 1, 1]


def test_sigma_world_and_jacobian_and_conic(fx, pg):
    """test/test_projection.py:72-93 (sigma_world of gaussians 0 and 4), :101-106 (J of gaussian 0),
    :118-120 (conic of gaussian 3).  Culled rows are not produced by the fused chain, so evaluate
    the stage functions without culling (far away thresholds)."""
    big = orc.project(fx["xyz"], fx["quaternion"], fx["scale"], inverse_sigmoid(fx["opacity"]), fx["rgb"], None,
                      fx["T"], fx["K"], 480, 640, -1e9, 1e9, 1e9)
    np.testing.assert_allclose(big.conic[3], [664.28760, 254.81781, 5761.8906], rtol=2e-6)
    # sigma_world / J are internal to the oracle's fused chain; check them through the C helpers' effect:
    # conic of gaussian 4 recomputed in fp64 from the reference's Sigma_world literal
    S4 = np.array([[0.01454808, 0.01702517, 0.07868834], [0.01702517, 0.4389012, 1.1959752],
                   [0.07868834, 1.1959752, 3.5965507]])
    p = big.xyz_cam[4].astype(np.float64)
    fxv, fyv = 430.0, 410.0
    J = np.array([[fxv / p[2], 0, -fxv * p[0] / p[2] ** 2], [0, fyv / p[2], -fyv * p[1] / p[2] ** 2]])
    W = fx["T"][:3, :3].astype(np.float64)
    S2 = J @ W @ S4 @ W.T @ J.T
    np.testing.assert_allclose(big.conic[4], [S2[0, 0], S2[0, 1] + S2[1, 0], S2[1, 1]], rtol=2e-4)
    # J of gaussian 0 (behind the camera): -295.5936, -134.1520, -281.8451, 229.5912
    p0 = big.xyz_cam[0].astype(np.float64)
    assert abs(fxv / p0[2] - (-295.5936)) < 5e-4 and abs(-fxv * p0[0] / p0[2] ** 2 - (-134.1520)) < 5e-4
    assert abs(fyv / p0[2] - (-281.8451)) < 5e-4 and abs(-fyv * p0[1] / p0[2] ** 2 - 229.5912) < 5e-4


# test/test_tile_culling.py:73-103 — the exact depth-sorted gaussian list of the fixture is long; its
# structure is pinned here (length, ranges, per-tile order) and the full list by the compiled reference in
# tests/golden/fixture6_fp32.npz (test_oracle_golden.py).
def test_tile_lists_shape(pg):
    keep = pg.visible.astype(bool)
    idx, ranges = orc.tile_lists(pg.uv[keep], pg.xyz_cam[keep], pg.conic[keep], 40, 30, 3.0)
    assert len(ranges) == 1201          # test/test_tile_culling.py:108
    assert len(idx) == 641              # test/test_tile_culling.py:73-103 (641 entries)
    assert ranges[0] == 0 and ranges[-1] == 641 and (np.diff(ranges) >= 0).all()
    z = pg.xyz_cam[keep][:, 2]
    for t in range(1200):
        seg = idx[ranges[t]:ranges[t + 1]]
        assert (np.diff(z[seg]) >= 0).all()


def _rasterize(fx, sh=None):
    return orc.rasterize(fx["xyz"], fx["quaternion"], fx["scale"], inverse_sigmoid(fx["opacity"]), fx["rgb"], sh,
                         fx["T"], fx["K"], 480, 640, 0.3, 100.0, 10.0, 3.0, background=np.zeros(3, np.float32))


def test_rasterize_no_sh(fx):
    """test/test_rasterize.py:47-54 (places=5)"""
    o = _rasterize(fx)
    np.testing.assert_allclose(o.image[340, 348], [0.47698545455932617, 0.0, 0.0], atol=5e-6)
    np.testing.assert_allclose(o.image[200, 348], [0.03330837935209274, 0.0, 0.267561137676239], atol=5e-6)
