"""Shared scene builders for the tests (numpy views of gaussian_splatting_b200.synth scenes and the
reference's 6-gaussian fixture, test/gaussian_test_data.py:7-79, restated)."""
import numpy as np
import torch

from gaussian_splatting_b200 import synth

SH0 = 0.28209479177387814


def np_scene(n, res, sh_degree=3, seed=0, view=0, n_views=3, dtype=np.float32, sigma_px=(1.2, 0.6, 0.3, 12.0)):
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    g = synth.make_gaussians(n, res, sh_degree=sh_degree, seed=seed, dtype=tdt, sigma_px=sigma_px)
    cam = synth.make_camera(res, dtype=tdt)
    T = synth.make_pose(view, n_views, dtype=tdt)
    a = lambda t: None if t is None else t.detach().numpy()  # noqa: E731
    return dict(xyz=a(g.xyz), quaternion=a(g.quaternion), scale=a(g.scale), opacity=a(g.opacity), rgb=a(g.rgb),
                sh=a(g.sh), T=a(T), K=a(cam.K), H=cam.height, W=cam.width)


def reference_fixture(dtype=np.float32):
    """The reference's own unit-test scene: 6 gaussians, 640x480 camera, tilted pose."""
    xyz = np.array([[1, 2, -4], [4, 5, 6], [7, 8, -9], [1, 2, 15], [2.5, -1, 4], [-1, -2, 10]], dtype)
    rgb = np.full((6, 3), 0.5, dtype)
    rgb[3], rgb[4], rgb[5] = [0.5, 0, 0], [0, 0.5, 0], [0, 0, 0.5]
    rgb = (rgb / dtype(SH0)).astype(dtype)
    opacity = np.ones((6, 1), dtype)
    scale = np.log(np.array([[0.02, 0.03, 0.04], [0.01, 0.05, 0.02], [0.09, 0.03, 0.01], [1, 3, 0.1], [2, 0.2, 0.1],
                             [2, 1, 0.1]], dtype))
    quaternion = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [1, 0, 0, 0], [0.714, -0.002, -0.664, 0.221],
                           [1, 0, 0, 0]], dtype)
    K = np.array([[430, 0, 320], [0, 410, 240], [0, 0, 1]], dtype)
    T = np.array([[0.9999, 0.0089, 0.0073, -0.3283], [-0.0106, 0.9568, 0.2905, -1.9260],
                  [-0.0044, -0.2906, 0.9568, 2.9581], [0, 0, 0, 1]], dtype)
    return dict(xyz=xyz, quaternion=quaternion, scale=scale, opacity=opacity, rgb=rgb, sh=None, T=T, K=K, H=480, W=640)


def inverse_sigmoid(x):
    p = np.clip(x, 1e-4, 1 - 1e-4)
    return np.log(p / (1 - p)).astype(x.dtype)
