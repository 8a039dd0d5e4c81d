"""Write a synthetic scene in the on-disk layout the reference's loader reads (splat_py/dataloader.py:103-187):

    <dir>/sparse/0/cameras.bin, images.bin, points3D.bin     COLMAP binary model (one PINHOLE camera)
    <dir>/images_1/view_XXX.png                              ground-truth views

The ground truth is a cloud of opaque-ish coloured Gaussians on a few noisy shells, rendered from a ring of
cameras by this library's rasterizer (constant colour, black background).  The sparse points handed to the
trainer are a noisy subsample of the ground-truth means — what a structure-from-motion run would deliver.
There is no network and no dataset in this image; this stands in for Mip-NeRF-360 `garden` in size class only
(BASELINE.json configs[4]): the point is that the reference's unmodified trainer runs end to end on top of
the library, not the PSNR value itself.
"""
from __future__ import annotations

import argparse
import os
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def look_at(eye):
    """world->camera pose of a camera at `eye` looking at the origin; +z forward, +x right, +y down."""
    f = -eye / np.linalg.norm(eye)
    r = np.cross(f, np.array([0.0, 0.0, 1.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    R = np.stack([r, d, f])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = -R @ eye
    return T


def rotmat_to_qvec(R):
    """(w, x, y, z) unit quaternion of a rotation matrix (Shepperd's method: pivot on the largest of w,x,y,z)."""
    t = np.trace(R)
    cand = np.array([t, R[0, 0], R[1, 1], R[2, 2]])
    k = int(np.argmax(cand))
    if k == 0:
        q = np.array([1 + t, R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    elif k == 1:
        q = np.array([R[2, 1] - R[1, 2], 1 + 2 * R[0, 0] - t, R[0, 1] + R[1, 0], R[0, 2] + R[2, 0]])
    elif k == 2:
        q = np.array([R[0, 2] - R[2, 0], R[0, 1] + R[1, 0], 1 + 2 * R[1, 1] - t, R[1, 2] + R[2, 1]])
    else:
        q = np.array([R[1, 0] - R[0, 1], R[0, 2] + R[2, 0], R[1, 2] + R[2, 1], 1 + 2 * R[2, 2] - t])
    q = q / np.linalg.norm(q)
    return -q if q[0] < 0 else q


def write_cameras(path, width, height, fx, fy, cx, cy):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", 1))
        f.write(struct.pack("<iiQQ", 1, 1, width, height))  # camera_id 1, model 1 = PINHOLE
        f.write(struct.pack("<4d", fx, fy, cx, cy))


def write_images(path, poses, names):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(poses)))
        for i, (T, name) in enumerate(zip(poses, names), start=1):
            q = rotmat_to_qvec(T[:3, :3])
            f.write(struct.pack("<i4d3di", i, *q, *T[:3, 3], 1))
            f.write(name.encode() + b"\x00")
            f.write(struct.pack("<Q", 0))  # no 2-D observations


def write_points(path, xyz, rgb_u8):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(xyz)))
        for i, (p, c) in enumerate(zip(xyz, rgb_u8), start=1):
            f.write(struct.pack("<Q3d3BdQ", i, *p, *[int(v) for v in c], 0.5, 0))  # empty track


def ground_truth(n, seed, scale_lo=0.015, scale_hi=0.04):
    rng = np.random.default_rng(seed)
    shell = rng.choice([0.6, 1.0, 1.5], size=n, p=[0.2, 0.4, 0.4])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyz = d * (shell + rng.normal(scale=0.03, size=n))[:, None]
    xyz[:, 2] *= 0.6
    # colour varies smoothly with position so that a sparse subsample carries useful initial colours
    rgb = 0.5 + 0.45 * np.sin(xyz @ rng.normal(scale=2.5, size=(3, 3)) + rng.uniform(0, 6.28, size=3))
    scale = np.log(rng.uniform(scale_lo, scale_hi, size=(n, 3)))
    quat = rng.normal(size=(n, 4))
    opacity = rng.uniform(1.0, 4.0, size=(n, 1))  # logit
    return [a.astype(np.float32) for a in (xyz, rgb, scale, quat, opacity)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/e2e_scene")
    ap.add_argument("--gaussians", type=int, default=40000)
    ap.add_argument("--points", type=int, default=8000)
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=416)
    ap.add_argument("--focal", type=float, default=560.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--scale-lo", type=float, default=0.015, help="ground-truth gaussian extent (world units), lower bound")
    ap.add_argument("--scale-hi", type=float, default=0.04)
    a = ap.parse_args()

    import cv2
    import torch

    from gaussian_splatting_b200.rasterize import rasterize
    from gaussian_splatting_b200.structs import Camera, Gaussians

    dev = torch.device("cuda:0")
    xyz, rgb, scale, quat, opacity = ground_truth(a.gaussians, a.seed, a.scale_lo, a.scale_hi)
    g = Gaussians(xyz=torch.tensor(xyz, device=dev), rgb=torch.tensor(rgb / 0.28209479177387814, device=dev),
                  opacity=torch.tensor(opacity, device=dev), scale=torch.tensor(scale, device=dev),
                  quaternion=torch.tensor(quat, device=dev))
    K = torch.tensor([[a.focal, 0, a.width / 2], [0, a.focal, a.height / 2], [0, 0, 1]], dtype=torch.float32, device=dev)
    cam = Camera(a.width, a.height, K)

    out = Path(a.out)
    (out / "sparse" / "0").mkdir(parents=True, exist_ok=True)
    (out / "images_1").mkdir(parents=True, exist_ok=True)
    poses, names = [], []
    bg = torch.zeros(3, device=dev)
    for v in range(a.views):
        ang = 2 * np.pi * v / a.views
        eye = np.array([4.0 * np.cos(ang), 4.0 * np.sin(ang), 1.2 + 0.8 * np.sin(3 * ang)])
        T = look_at(eye)
        with torch.no_grad():
            img, _, _ = rasterize(g, torch.tensor(T, dtype=torch.float32, device=dev), cam, 0.3, 500.0, 100, 3.0, True, bg)
        u8 = (img.clip(0, 1) * 255.0 + 0.5).to(torch.uint8).cpu().numpy()
        name = f"view_{v:03d}.png"
        cv2.imwrite(str(out / "images_1" / name), u8[..., ::-1])
        poses.append(T)
        names.append(name)

    rng = np.random.default_rng(a.seed + 1)
    pick = rng.choice(a.gaussians, size=a.points, replace=False)
    pts = xyz[pick] + rng.normal(scale=0.01, size=(a.points, 3))
    write_cameras(out / "sparse" / "0" / "cameras.bin", a.width, a.height, a.focal, a.focal, a.width / 2, a.height / 2)
    write_images(out / "sparse" / "0" / "images.bin", poses, names)
    write_points(out / "sparse" / "0" / "points3D.bin", pts, (rgb[pick].clip(0, 1) * 255).astype(np.uint8))
    print(f"[scene] {out}: {a.views} views {a.width}x{a.height}, {a.points} sparse points, "
          f"ground truth {a.gaussians} gaussians; mean image level {float(u8.mean()) / 255:.3f}")


if __name__ == "__main__":
    main()
