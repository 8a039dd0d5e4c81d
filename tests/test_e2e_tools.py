"""tools/e2e (reference trainer on a synthetic COLMAP scene): the scene writer's binary files are read back by the
reference's own COLMAP reader (installed copy under oracle/_ref), and the SSIM stand-in behaves like an SSIM."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
E2E = ROOT / "tools" / "e2e"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_colmap_writers_round_trip_through_the_reference_reader(tmp_path):
    reader_py = ROOT / "oracle" / "_ref" / "splat_py" / "read_colmap.py"
    if not reader_py.exists():
        pytest.skip("oracle/_ref (installed reference) not present")
    rd = _load(reader_py, "ref_read_colmap")
    mk = _load(E2E / "make_colmap_scene.py", "make_colmap_scene")
    poses = [mk.look_at(np.array([4 * np.cos(a), 4 * np.sin(a), 1.2 + 0.8 * np.sin(3 * a)])) for a in np.linspace(0, 6.2, 17)]
    names = [f"view_{i:03d}.png" for i in range(len(poses))]
    mk.write_cameras(tmp_path / "cameras.bin", 640, 416, 560.0, 561.0, 320.0, 208.0)
    mk.write_images(tmp_path / "images.bin", poses, names)
    xyz = np.random.default_rng(0).standard_normal((50, 3))
    rgb = np.random.default_rng(1).integers(0, 256, (50, 3)).astype(np.uint8)
    mk.write_points(tmp_path / "points3D.bin", xyz, rgb)
    cams = rd.read_cameras_binary(str(tmp_path / "cameras.bin"))
    assert cams[1].model == "PINHOLE" and (cams[1].width, cams[1].height) == (640, 416)
    assert np.allclose(cams[1].params, [560.0, 561.0, 320.0, 208.0])
    imgs = rd.read_images_binary(str(tmp_path / "images.bin"))
    assert len(imgs) == len(poses)
    for i, T in enumerate(poses, start=1):
        assert imgs[i].name == names[i - 1] and imgs[i].camera_id == 1
        assert np.abs(rd.qvec2rotmat(imgs[i].qvec) - T[:3, :3]).max() < 1e-12
        assert np.abs(imgs[i].tvec - T[:3, 3]).max() < 1e-12
    pts = rd.read_points3D_binary(str(tmp_path / "points3D.bin"))
    assert len(pts) == 50 and np.allclose(pts[7].xyz, xyz[6]) and (pts[7].rgb == rgb[6]).all()


def test_ssim_shim_is_an_ssim():
    sys.path.insert(0, str(E2E / "shims"))
    try:
        from torchmetrics.image import StructuralSimilarityIndexMeasure
    finally:
        sys.path.remove(str(E2E / "shims"))
    ssim = StructuralSimilarityIndexMeasure(data_range=1.0)
    g = torch.Generator().manual_seed(0)
    a = torch.rand(1, 3, 48, 64, generator=g)
    b = torch.rand(1, 3, 48, 64, generator=g)
    assert abs(float(ssim(a, a)) - 1.0) < 1e-6
    assert abs(float(ssim(a, b)) - float(ssim(b, a))) < 1e-6 and float(ssim(a, b)) < 0.2
    noisy = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    assert float(ssim(a, b)) < float(ssim(a, noisy)) < 1.0
    x = a.clone().requires_grad_(True)
    (1.0 - ssim(x, b)).backward()
    assert torch.isfinite(x.grad).all() and float(x.grad.abs().sum()) > 0
