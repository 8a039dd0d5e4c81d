"""Does the in-kernel world->camera transform reproduce torch.matmul's bits?  (GPU, development tool)

For several batch sizes and poses: run the fused per-gaussian kernel with and without caller-supplied
camera-frame positions and compare depth keys, uv, conic bitwise.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gaussian_splatting_b200 as gsb  # noqa: E402
from gaussian_splatting_b200 import synth  # noqa: E402
from gaussian_splatting_b200.utils import transform_points_torch  # noqa: E402
from tests import scenes  # noqa: E402


def poses(dev):
    out = {"yaw": synth.make_pose(0, 3, device=dev), "fixture6": torch.from_numpy(scenes.reference_fixture()["T"]).to(dev)}
    g = torch.Generator().manual_seed(11)
    for i in range(3):
        q = torch.randn(4, generator=g, dtype=torch.float64)
        q = q / q.norm()
        w, x, y, z = q.tolist()
        R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=torch.float64)
        T = torch.eye(4, dtype=torch.float64)
        T[:3, :3] = R
        T[:3, 3] = torch.randn(3, generator=g, dtype=torch.float64)
        out[f"random{i}"] = T.float().to(dev)
    return out


def main():
    dev = torch.device("cuda")
    ext = gsb.native()
    rep = {}
    Path('gpurun_out').mkdir(exist_ok=True)
    for n in (6, 300, 1000, 20000, 65535, 65536, 65537, 65535 + 100, 65535 + 290, 65535 + 600, 131070 + 1, 1000000, 3000000):
        g = synth.make_gaussians(n, "1080p", sh_degree=0, seed=n, device=dev)
        cam = synth.make_camera("1080p", device=dev)
        for name, T in poses(dev).items():
            xyz_cam = transform_points_torch(g.xyz, T)
            tail = n % 65535
            cam_tail = transform_points_torch(g.xyz[n - tail:], T) if (n >= 16384 and 0 < tail < 1024) else (
                xyz_cam if n < 16384 else None)
            a = ext.fused_preprocess_forward(g.xyz, cam_tail, g.quaternion, g.scale, g.opacity.reshape(-1), g.rgb, None, T,
                                             cam.K, None, 1080, 1920, -1e30, 1e30, 1e30, 3.0, 0)
            b = ext.fused_preprocess_forward(g.xyz, xyz_cam, g.quaternion, g.scale, g.opacity.reshape(-1), g.rgb, None, T,
                                             cam.K, None, 1080, 1920, -1e30, 1e30, 1e30, 3.0, 0)
            if False:
                np.savez(f"gpurun_out/transform_sample_N{n}_{name}.npz", xyz=g.xyz[:20000].cpu().numpy(), T=T.cpu().numpy(),
                         xyz_cam=xyz_cam[:20000].cpu().numpy())
            bad = torch.nonzero(a[1] != b[1]).flatten()
            if 0 < bad.numel() < 50:
                print("   mismatching rows:", bad.tolist()[:20], flush=True)
            same_z = float((a[1] == b[1]).float().mean())
            ra, rb = a[0].view(torch.int32), b[0].view(torch.int32)
            fin = torch.isfinite(b[0]).all(dim=1)
            same_rec = float((ra[fin] == rb[fin]).all(dim=1).float().mean()) if bool(fin.any()) else 1.0
            rep[f"N={n},{name}"] = dict(z_match=same_z, record_match=same_rec)
            print(f"N={n:8d} {name:9s} z {same_z:.6f} records {same_rec:.6f}", flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/transform_check.json").write_text(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
