"""Host-side mirror of the reference's python modules (no GPU)."""
import math

import numpy as np
import torch

from gaussian_splatting_b200 import synth
from gaussian_splatting_b200.structs import Camera, Gaussians, Tiles
from gaussian_splatting_b200.utils import (compute_rays, compute_rays_in_world_frame, inverse_sigmoid_torch,
                                           quaternion_to_rotation_torch, transform_points_torch)


def test_tiles_1080p():
    """test/test_structs.py:21-26"""
    t = Tiles(1080, 1920, "cpu")
    assert (t.image_height_padded, t.image_width_padded) == (1088, 1920)
    assert (t.y_tiles_count, t.x_tiles_count, t.tile_count) == (68, 120, 8160)
    t = Tiles(480, 640, "cpu")
    assert (t.y_tiles_count, t.x_tiles_count) == (30, 40)


def test_gaussians_container():
    g = synth.make_gaussians(10, "tiny", sh_degree=3)
    assert len(g) == 10 and g.sh.shape == (10, 3, 15)
    g.filter_in_place(torch.arange(10) % 2 == 0)
    assert len(g) == 5 and isinstance(g.xyz, torch.nn.Parameter)
    extra = synth.make_gaussians(3, "tiny", sh_degree=3, seed=5)
    g.append(extra.xyz, extra.rgb, extra.opacity, extra.scale, extra.quaternion, extra.sh)
    assert len(g) == 8 and g.sh.shape == (8, 3, 15)


def test_quaternion_to_rotation_is_orthonormal():
    """test/test_utils.py (orthogonality check)"""
    q = torch.randn(32, 4, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    R = quaternion_to_rotation_torch(q)
    eye = torch.eye(3, dtype=torch.float64).expand(32, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-12)
    assert torch.allclose(torch.linalg.det(R), torch.ones(32, dtype=torch.float64), atol=1e-12)


def test_transform_round_trip():
    T = synth.make_pose(0, 3, dtype=torch.float64)
    pts = torch.randn(100, 3, dtype=torch.float64)
    back = transform_points_torch(transform_points_torch(pts, T), torch.inverse(T))
    assert torch.allclose(back, pts, atol=1e-12)
    assert transform_points_torch(pts, T).is_contiguous()


def test_rays():
    cam = synth.make_camera("tiny", dtype=torch.float64)
    r = compute_rays(cam)
    assert r.shape == (64 * 64, 3) and torch.allclose(r.norm(dim=1), torch.ones(64 * 64, dtype=torch.float64))
    # principal ray at (cx, cy) = (32, 32) is +z
    assert torch.allclose(r[32 * 64 + 32], torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64))
    rw = compute_rays_in_world_frame(cam, torch.eye(4, dtype=torch.float64))
    assert rw.shape == (64, 64, 3) and torch.allclose(rw.reshape(-1, 3), r)


def test_inverse_sigmoid_clip():
    x = torch.tensor([0.0, 0.5, 1.0])
    y = inverse_sigmoid_torch(x)
    assert math.isclose(y[0].item(), math.log(1e-4 / (1 - 1e-4)), rel_tol=1e-5)
    assert y[1].item() == 0.0
    assert math.isclose(y[2].item(), -y[0].item(), rel_tol=1e-3)


def test_synth_is_deterministic_and_shaped():
    a = synth.make_gaussians(1000, "720p", seed=0)
    b = synth.make_gaussians(1000, "720p", seed=0)
    for name in ("xyz", "quaternion", "scale", "opacity", "rgb", "sh"):
        assert torch.equal(getattr(a, name), getattr(b, name))
    assert a.xyz[:, 2].min() >= 1 and a.xyz[:, 2].max() <= 11
    T = synth.make_pose(0, 1)
    assert torch.equal(T, torch.eye(4))
    T0, T2 = synth.make_pose(0, 3), synth.make_pose(2, 3)
    assert torch.allclose(T0[:3, :3], T2[:3, :3].T, atol=1e-6)  # +-3 degree yaw
    G = synth.make_upstream_grad("tiny")
    assert G.shape == (64, 64, 3) and abs(G.std().item() * 3 * 64 * 64 - 1) < 0.05


def test_flat_parameter_layout_sections():
    """flat_adam.section_ends (host code of the native module): [xyz|quaternion|scale|opacity|rgb|sh], every
    section starting on a 16-byte boundary, the layout gsr_preprocess_backward writes its gradients in."""
    from gaussian_splatting_b200.flat_adam import FIELDS, section_ends

    assert FIELDS == ("xyz", "quaternion", "scale", "opacity", "rgb", "sh")
    for n, k in ((1, 0), (5, 3), (1001, 15), (3_000_000, 15)):
        ends = section_ends(n, k)
        widths = [3, 4, 3, 1, 3] + ([3 * k] if k else [])
        assert len(ends) == len(widths)
        start = 0
        for w, e in zip(widths, ends):
            assert start % 4 == 0 and e % 4 == 0 and 0 <= e - start - n * w < 4  # padded up to 4 floats
            start = e
    assert section_ends(3_000_000, 15)[-1] == 3_000_000 * 59  # no padding when N is a multiple of 4


def test_flatten_gaussians_keeps_values_and_kinds():
    """flat_adam.flatten_gaussians: fields become views of ONE buffer (plain tensors stay leaf tensors, nn.Parameters
    stay nn.Parameters, as the reference's trainer holds them), values unchanged, sections where section_ends says."""
    import torch

    from gaussian_splatting_b200.flat_adam import flatten_gaussians
    from gaussian_splatting_b200.structs import Gaussians

    n = 7
    g = Gaussians(torch.randn(n, 3), torch.randn(n, 3), torch.randn(n, 1), torch.randn(n, 3), torch.randn(n, 4),
                  torch.randn(n, 3, 3))
    g.xyz = torch.nn.Parameter(g.xyz)           # mixed kinds on purpose
    before = {f: getattr(g, f).detach().clone() for f in ("xyz", "quaternion", "scale", "opacity", "rgb", "sh")}
    flat, ends, names = flatten_gaussians(g)
    assert names == ["xyz", "quaternion", "scale", "opacity", "rgb", "sh"] and flat.numel() == ends[-1]
    assert isinstance(g.xyz, torch.nn.Parameter) and not isinstance(g.rgb, torch.nn.Parameter)
    start = 0
    for f, e in zip(names, ends):
        t = getattr(g, f)
        assert t.requires_grad and t.is_leaf and torch.equal(t.detach(), before[f])
        assert t.data_ptr() == flat.data_ptr() + 4 * start       # a view at the section's start
        start = e
    flat.add_(1.0)                                               # an update of the flat buffer is seen by the fields
    assert torch.equal(g.scale.detach(), before["scale"] + 1.0)


def test_gradient_bucket_fallback_keeps_the_native_layout():
    """the copy fallback of GradientBucket.adopt must be laid out like state.grad_flat
    ([xyz | quaternion | scale | opacity | rgb | sh], 16-byte aligned section ends) when the field names are
    known, and FlatAdam.step must refuse a bucket in any other layout."""
    import pytest
    import torch

    from gaussian_splatting_b200.flat_adam import FIELDS, FlatAdam, section_ends
    from gaussian_splatting_b200.structs import Gaussians
    from gaussian_splatting_b200.view_parallel import GradientBucket

    n, k = 6, 15  # n % 4 != 0: section ends are padded
    torch.manual_seed(0)
    g = Gaussians(xyz=torch.randn(n, 3), rgb=torch.randn(n, 3), opacity=torch.randn(n, 1), scale=torch.randn(n, 3),
                  quaternion=torch.randn(n, 4), sh=torch.randn(n, 3, k))
    for f in FIELDS:
        getattr(g, f).requires_grad_(True)
        getattr(g, f).grad = torch.full_like(getattr(g, f), float(FIELDS.index(f) + 1))
    other = torch.empty(8)  # gradients live elsewhere -> copy fallback
    b = GradientBucket.adopt(other, g)
    assert not b.zero_copy and b.layout == "native"
    ends = section_ends(n, k)
    assert b.flat.numel() == ends[-1]
    start = 0
    for i, (f, end) in enumerate(zip(FIELDS, ends)):
        numel = getattr(g, f).numel()
        assert torch.all(b.flat[start:start + numel] == float(i + 1)), f
        assert torch.all(b.flat[start + numel:end] == 0.0)  # padding
        assert getattr(g, f).grad.data_ptr() == b.flat[start:].data_ptr()  # attached
        start = end
    # a bare list carries no field names: packed layout, good for the collective only
    packed = GradientBucket.adopt(other, [g.xyz, g.rgb, g.opacity, g.scale, g.quaternion, g.sh])
    assert packed.layout == "packed"

    class _Opt(FlatAdam):
        def __init__(self):  # no device buffers needed for the layout check
            self.p = torch.empty(ends[-1])

    with pytest.raises(ValueError):
        _Opt().step(packed)


def test_depth_key_parameters_cover_the_degenerate_ranges():
    """rasterize._depth_key_params: (base, bits) of the order-preserving depth key.  A usable positive range gives the
    bit length of bits(far) - bits(near); anything else must fall back to the full 32-bit key, because depth_key()
    sets bit 31 whenever the base is 0."""
    import struct

    from gaussian_splatting_b200.rasterize import _depth_key_params

    def fbits(v):
        return struct.unpack("<I", struct.pack("<f", v))[0]

    base, bits = _depth_key_params(0.3, 500.0)
    assert base == fbits(0.3) and bits == (fbits(500.0) - fbits(0.3)).bit_length() == 27
    for near, far in ((0.0, 10.0), (-1.0, 10.0), (1e-50, 10.0), (0.3, float("inf")), (0.3, 1e39), (0.3, float("nan")),
                      (5.0, 5.0), (7.0, 2.0)):
        assert _depth_key_params(near, far) == (0, 32), (near, far)
    assert _depth_key_params(1.0, 1.0000001)[1] >= 1


def test_speculative_pair_capacity_follows_the_recent_views():
    """rasterize._pair_capacity / _note_pairs: no history -> None (the first view of a kind reads the
    count eagerly); otherwise max of the last 8 counts x headroom + 64 Ki, a multiple of 128, below 2^31."""
    from gaussian_splatting_b200 import rasterize as R

    key = ("test-key", 123)
    R._PAIR_HISTORY.pop(key, None)
    assert R._pair_capacity(key) is None
    for p in (1000, 5_000_000, 4_000_000):
        R._note_pairs(key, p)
    cap = R._pair_capacity(key)
    assert cap % 128 == 0 and cap >= int(5_000_000 * R.PAIR_HEADROOM) + 65536 and cap < 5_000_000 * 1.1 + 70000
    for p in range(9):                      # the 5 M view drops out of the window of 8
        R._note_pairs(key, 2_000_000 + p)
    assert R._pair_capacity(key) < 3_000_000
    R._note_pairs(key, 2**31 - 1)
    assert R._pair_capacity(key) == 2**31 - 128
    del R._PAIR_HISTORY[key]
