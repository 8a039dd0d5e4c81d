"""Run the reference's UNMODIFIED trainer (splat_py/trainer.py `SplatTrainer`) on a COLMAP-layout scene with
the rasterization path provided by one of three backends, and print one JSON line with quality and speed:

    --impl ref          reference python on the reference's own CUDA extension (oracle/_ref, checker build)
    --impl b200         reference python (its rasterize, its autograd Functions) on this library's `splat_cuda`
                        module — the binary drop-in of INTEGRATION.md §1
    --impl b200-fused   as b200, with `splat_py.trainer.rasterize` rebound to this library's fused rasterize —
                        the one-line source-level drop-in of INTEGRATION.md §2

The set-up lines between loading the data and calling train() restate the reference's CLI script
(colmap_splat.py:42-75), which executes at import time and therefore cannot be imported.  `torchmetrics` and
`plotext` are not in this image; tools/e2e/shims provides a plain SSIM and a no-op plotter — the same for all
three backends.  This tool is a demonstration harness: it reads oracle/_ref (the installed reference package)
and is therefore not part of tests/, smoke() or bench.py.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent / "shims"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ref", "b200", "b200-fused"], required=True)
    ap.add_argument("--scene", default="gpurun_out/e2e_scene")
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=7000)
    ap.add_argument("--max-gaussians", type=int, default=400000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()

    import torch

    from oracle import ref_loader

    if a.impl == "ref":
        sys.modules["splat_cuda"] = ref_loader._load_ref_ext()
    else:
        import gaussian_splatting_b200 as gsb

        gsb.install_as_splat_cuda()
    sys.path.insert(0, str(ref_loader.REF_DIR))

    from splat_py.config import SplatConfig
    from splat_py.dataloader import ColmapData
    import splat_py.trainer as ref_trainer

    if a.impl == "b200-fused":
        from gaussian_splatting_b200.rasterize import rasterize as fused_rasterize

        ref_trainer.rasterize = fused_rasterize

    out = a.out or f"gpurun_out/e2e_{a.impl}"
    os.makedirs(out, exist_ok=True)
    scale = a.iters / 7000.0  # the reference's "7k" schedule, shortened proportionally when --iters < 7000
    sched = {}
    if a.iters != 7000:
        base = SplatConfig()
        for key in ("adaptive_control_start", "adaptive_control_end", "use_background_end", "reset_opacity_start",
                    "reset_opacity_end", "reset_opacity_interval", "add_sh_band_interval"):
            sched[key] = max(1, int(getattr(base, key) * scale))
    config = SplatConfig(dataset_path=a.scene, downsample_factor=1, output_dir=out, num_iters=a.iters,
                         max_gaussians=a.max_gaussians, save_debug_image_interval=10 ** 9, print_interval=500,
                         checkpoint_interval=10 ** 9, **sched)

    torch.manual_seed(a.seed)
    data = ColmapData(config.dataset_path, torch.device("cuda"), downsample_factor=config.downsample_factor, config=config)
    gaussians = data.create_gaussians()
    for name in ("xyz", "quaternion", "scale", "opacity", "rgb"):
        setattr(gaussians, name, torch.nn.Parameter(getattr(gaussians, name)))
    n0 = gaussians.xyz.shape[0]

    trainer = ref_trainer.SplatTrainer(gaussians, data.get_images(), data.get_cameras(), config)
    torch.cuda.synchronize()
    t0 = time.time()
    trainer.train()
    torch.cuda.synchronize()
    seconds = time.time() - t0

    psnr, ssim = trainer.compute_test_psnr()
    line = {
        "impl": a.impl, "iters": a.iters, "seed": a.seed, "train_seconds": round(seconds, 2),
        "iters_per_second": round(a.iters / seconds, 2),
        "final_test_psnr": round(float(psnr.mean()), 3), "final_test_ssim": round(float(ssim.mean()), 4),
        "max_test_psnr": round(max(trainer.metrics.test_psnr + [float(psnr.mean())]), 3),
        "test_psnr_curve": [round(v, 2) for v in trainer.metrics.test_psnr],
        "gaussians_start": n0, "gaussians_end": int(trainer.gaussians.xyz.shape[0]),
        "views": len(trainer.images), "image": [int(trainer.images[0].image.shape[1]), int(trainer.images[0].image.shape[0])],
        "device": torch.cuda.get_device_name(0),
    }
    print("E2E " + json.dumps(line))
    with open(os.path.join(out, "result.json"), "w") as f:
        json.dump(line, f, indent=1)


if __name__ == "__main__":
    main()
