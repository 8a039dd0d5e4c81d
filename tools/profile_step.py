"""Minimal workload for ncu: a few fwd+bwd steps of the bench configuration (3M gaussians, 1080p, SH 3).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3
    (profiling starts after one warm-up step: the first step also runs the one-time per-device self-checks)
    ncu --set full --clock-control none --import-source on -k regex:k_render_bwd -s 1 -c 1 \\
        -o gpurun_out/prof_render_bwd python tools/profile_step.py --steps 2
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gaussian_splatting_b200 import synth  # noqa: E402
from gaussian_splatting_b200.rasterize import rasterize  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--n", type=int, default=3_000_000)
    ap.add_argument("--res", default="1080p")
    args = ap.parse_args()
    dev = torch.device("cuda")
    g = synth.make_gaussians(args.n, args.res, sh_degree=3, seed=0, device=dev, requires_grad=True)
    cam = synth.make_camera(args.res, device=dev)
    G = synth.make_upstream_grad(args.res, device=dev)
    bg = torch.full((3,), 0.5, device=dev)
    for i in range(-1, args.steps):
        if i == 0:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()  # honoured by `ncu --profile-from-start off`; a no-op otherwise
        T = synth.make_pose(i % 8, 8, device=dev)
        for p in (g.xyz, g.rgb, g.opacity, g.scale, g.quaternion, g.sh):
            p.grad = None
        torch.cuda.nvtx.range_push(f"step{i}")
        image, _, _ = rasterize(g, T, cam, 0.3, 500.0, 100, 3.0, True, bg)
        image.backward(G)
        torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done", float(image.mean()))


if __name__ == "__main__":
    main()
