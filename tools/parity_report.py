"""GPU parity + timing report against the compiled reference (oracle/_ref).  Development tool.

    python tools/parity_report.py [--n 100000] [--res 720p] [--sh 3] [--out gpurun_out/parity.json]

Stage by stage: bitwise comparison of every per-Gaussian operator, the tile lists, the forward
image / per-pixel state, relative error of every gradient, the fused rasterize() against the
reference rasterize(), and CUDA-event timings of both.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
import traceback
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import gaussian_splatting_b200 as gsb  # noqa: E402
from gaussian_splatting_b200 import synth  # noqa: E402
from gaussian_splatting_b200.rasterize import rasterize as rasterize_b200  # noqa: E402
from gaussian_splatting_b200.rasterize import rasterize_unfused  # noqa: E402
from oracle import ref_loader  # noqa: E402


def bits_equal(a, b):
    a, b = a.contiguous(), b.contiguous()
    if a.shape != b.shape:
        return dict(equal=False, shape_a=list(a.shape), shape_b=list(b.shape))
    if a.dtype.is_floating_point:
        ia = a.view(torch.int32 if a.dtype == torch.float32 else torch.int64)
        ib = b.view(torch.int32 if b.dtype == torch.float32 else torch.int64)
    else:
        ia, ib = a, b
    neq = ia != ib
    n_bad = int(neq.sum().item())
    out = dict(equal=n_bad == 0, mismatching=n_bad, total=a.numel())
    if n_bad and a.dtype.is_floating_point:
        d = (a.double() - b.double()).abs()
        out["max_abs"] = float(d.max().item())
        out["max_rel_to_max"] = float((d.max() / b.double().abs().max().clamp_min(1e-300)).item())
        if a.dtype == torch.float32:
            ulp = (ia.long() - ib.long()).abs()
            out["max_ulp"] = int(ulp.max().item())
    return out


def rel_err(a, b):
    a, b = a.double(), b.double()
    if a.shape != b.shape:
        return dict(shape_a=list(a.shape), shape_b=list(b.shape))
    scale = b.abs().max().clamp_min(1e-300)
    return dict(max_rel_to_max=float(((a - b).abs().max() / scale).item()), ref_absmax=float(scale.item()),
                n_nonfinite=int((~torch.isfinite(a)).sum().item()))


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return dict(median_ms=ts[len(ts) // 2], min_ms=ts[0], max_ms=ts[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--res", default="720p")
    ap.add_argument("--sh", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/parity.json")
    ap.add_argument("--no-timing", action="store_true")
    ap.add_argument("--pose", choices=["yaw", "general"], default="general")
    args = ap.parse_args()
    dev = torch.device("cuda")
    rep = dict(args=vars(args), gpu=torch.cuda.get_device_name(0))
    ext = gsb.native()
    ref_ext, ref_py = ref_loader.load_reference()
    ref_ras = sys.modules["splat_py_ref.rasterize"]
    ref_utils = sys.modules["splat_py_ref.utils"]
    ref_af = sys.modules["splat_py_ref.cuda_autograd_functions"]
    ref_tc = sys.modules["splat_py_ref.tile_culling"]
    ref_structs = sys.modules["splat_py_ref.structs"]

    def section(name, fn):
        try:
            rep[name] = fn()
        except Exception as e:  # keep going: this is a diagnostic tool
            rep[name] = dict(error=repr(e), trace=traceback.format_exc()[-2000:])
        print(name, json.dumps(rep[name])[:1500], flush=True)

    g = synth.make_gaussians(args.n, args.res, sh_degree=args.sh, seed=0, device=dev)
    cam = synth.make_camera(args.res, device=dev)
    if args.pose == "yaw":
        T = synth.make_pose(0, 3, device=dev)  # -3 degree yaw about (0,0,6)
    else:  # general rigid pose: every entry of the 3x4 block is non-trivial
        gq = torch.Generator().manual_seed(7)
        q = torch.randn(4, generator=gq, dtype=torch.float64) * 0.05 + torch.tensor([1.0, 0, 0, 0], dtype=torch.float64)
        q = q / q.norm()
        w_, x_, y_, z_ = q.tolist()
        R = torch.tensor([[1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_)],
                          [2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_)],
                          [2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)]],
                         dtype=torch.float64)
        T = torch.eye(4, dtype=torch.float64)
        T[:3, :3] = R
        T[:3, 3] = torch.tensor([0.13, -0.21, 0.37], dtype=torch.float64)
        T = T.float().to(dev)
    H, W = cam.height, cam.width
    cfg = synth.DEFAULTS
    bg = torch.full((3,), 0.5, device=dev)

    # ---- per-gaussian operators, bitwise ------------------------------------------------------
    state = {}

    def per_gaussian():
        out = {}
        xyz_cam = ref_utils.transform_points_torch(g.xyz, T)
        state["xyz_cam"] = xyz_cam
        # which rounding order does torch's matmul use?  emulate fp32 fma in fp64 (probe only)
        x, y, z = (c.double() for c in g.xyz.unbind(1))

        def f32(v):
            return v.float().double()

        def fma(a, b, c):
            return f32(a * b + c)
        cand = {}
        for row in range(3):
            t0, t1, t2, t3 = (T[row, k].double() for k in range(4))
            cand.setdefault("A_fma_ascending", []).append(f32(fma(z, t2, fma(y, t1, f32(x * t0))) + t3))
            cand.setdefault("B_fma_descending", []).append(fma(x, t0, fma(y, t1, fma(z, t2, t3))))
            cand.setdefault("C_no_fma", []).append(f32(f32(f32(f32(x * t0) + f32(y * t1)) + f32(z * t2)) + t3))
            cand.setdefault("D_fma_from_t3", []).append(fma(z, t2, fma(y, t1, fma(x, t0, t3))))
            cand.setdefault("E_pairwise", []).append(f32(fma(x, t0, f32(y * t1)) + fma(z, t2, t3)))
        out["transform_order_match_fraction"] = {
            k: float((torch.stack(v, 1).float() == xyz_cam).float().mean()) for k, v in cand.items()}
        import numpy as _np
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        _np.savez(str(Path(args.out).with_suffix("")) + "_transform_sample.npz", xyz=g.xyz[:20000].cpu().numpy(),
                  T=T.cpu().numpy(), xyz_cam=xyz_cam[:20000].cpu().numpy())
        if not args.no_timing:
            out["t_transform_points_torch"] = timed(lambda: ref_utils.transform_points_torch(g.xyz, T))
        uv_ref = torch.zeros(args.n, 2, device=dev)
        ref_ext.camera_projection_cuda(xyz_cam, cam.K, uv_ref)
        uv_b = torch.zeros(args.n, 2, device=dev)
        ext.camera_projection_cuda(xyz_cam, cam.K, uv_b)
        out["uv"] = bits_equal(uv_b, uv_ref)
        s_ref = torch.zeros(args.n, 3, 3, device=dev)
        s_b = torch.zeros(args.n, 3, 3, device=dev)
        ref_ext.compute_sigma_world_cuda(g.quaternion, g.scale, s_ref)
        ext.compute_sigma_world_cuda(g.quaternion, g.scale, s_b)
        out["sigma_world"] = bits_equal(s_b, s_ref)
        j_ref = torch.zeros(args.n, 2, 3, device=dev)
        j_b = torch.zeros(args.n, 2, 3, device=dev)
        ref_ext.compute_projection_jacobian_cuda(xyz_cam, cam.K, j_ref)
        ext.compute_projection_jacobian_cuda(xyz_cam, cam.K, j_b)
        out["jacobian"] = bits_equal(j_b, j_ref)
        c_ref = torch.zeros(args.n, 3, device=dev)
        c_b = torch.zeros(args.n, 3, device=dev)
        ref_ext.compute_conic_cuda(s_ref, j_ref, T, c_ref)
        ext.compute_conic_cuda(s_ref, j_ref, T, c_b)
        out["conic"] = bits_equal(c_b, c_ref)
        if g.sh is not None:
            coeffs = torch.cat((g.rgb.unsqueeze(2), g.sh), dim=2).contiguous()
            Tinv = torch.inverse(T).contiguous()
            r_ref = torch.zeros(args.n, 3, device=dev)
            r_b = torch.zeros(args.n, 3, device=dev)
            ref_ext.precompute_rgb_from_sh_cuda(g.xyz, coeffs, Tinv, r_ref)
            ext.precompute_rgb_from_sh_cuda(g.xyz, coeffs, Tinv, r_b)
            out["sh_rgb"] = bits_equal(r_b, r_ref)
        state.update(uv=uv_ref, conic=c_ref)
        return out

    section("per_gaussian_ops", per_gaussian)

    # ---- fused preprocess vs the reference chain ---------------------------------------------------
    def fused_pre():
        out = {}
        sh = g.sh
        rec, zkey, vis, scan = ext.fused_preprocess_forward(
            g.xyz, None, g.quaternion, g.scale, g.opacity.reshape(-1), g.rgb, sh, T, cam.K, None, H, W,
            cfg["near_thresh"], cfg["far_thresh"], cfg["cull_mask_padding"], cfg["mh_dist"], 0)
        xyz_cam = state["xyz_cam"]
        uv = state["uv"]
        mask = ((xyz_cam[:, 2] < cfg["near_thresh"]) | (xyz_cam[:, 2] > cfg["far_thresh"])
                | (uv[:, 0] < -cfg["cull_mask_padding"]) | (uv[:, 0] > W + cfg["cull_mask_padding"])
                | (uv[:, 1] < -cfg["cull_mask_padding"]) | (uv[:, 1] > H + cfg["cull_mask_padding"]))
        out["culling_mask"] = bits_equal((vis == 0), mask)
        keep = ~mask
        out["M"] = int(keep.sum().item())
        out["uv_visible"] = bits_equal(rec[keep][:, 0:2], uv[keep])
        out["z_visible"] = bits_equal(zkey[keep], (xyz_cam[keep][:, 2].contiguous().view(torch.int32) | -2**31))
        conic = state["conic"][keep]
        out["a"] = bits_equal(rec[keep][:, 4], conic[:, 0] + 0.25)
        out["b2"] = bits_equal(rec[keep][:, 5], (conic[:, 1] * 0.5) * 2)
        out["c"] = bits_equal(rec[keep][:, 6], conic[:, 2] + 0.25)
        out["opacity"] = bits_equal(rec[keep][:, 3], torch.sigmoid(g.opacity[keep]).reshape(-1))
        total = int(scan[-1].item())
        out["M_scan"], out["P_scan"] = total >> 32, total & 0xFFFFFFFF
        return out

    section("fused_preprocess", fused_pre)

    # ---- tile lists -----------------------------------------------------------------------------------
    def binning():
        out = {}
        xyz_cam, uv, conic = state["xyz_cam"], state["uv"], state["conic"]
        mask = ((xyz_cam[:, 2] < cfg["near_thresh"]) | (xyz_cam[:, 2] > cfg["far_thresh"])
                | (uv[:, 0] < -cfg["cull_mask_padding"]) | (uv[:, 0] > W + cfg["cull_mask_padding"])
                | (uv[:, 1] < -cfg["cull_mask_padding"]) | (uv[:, 1] > H + cfg["cull_mask_padding"]))
        keep = ~mask
        uvk, xk, ck = uv[keep].contiguous(), xyz_cam[keep].contiguous(), conic[keep].contiguous()
        tiles = ref_structs.Tiles(H, W, dev)
        s_ref, r_ref = ref_ext.get_sorted_gaussian_list(1024, uvk, xk, ck, tiles.x_tiles_count, tiles.y_tiles_count,
                                                        cfg["mh_dist"])
        s_b, r_b = ext.get_sorted_gaussian_list(1024, uvk, xk, ck, tiles.x_tiles_count, tiles.y_tiles_count,
                                                cfg["mh_dist"])
        out["P_ref"], out["P_b200"] = int(s_ref.numel()), int(s_b.numel())
        out["tile_ranges"] = bits_equal(r_b, r_ref)
        out["sorted_idx"] = bits_equal(s_b, s_ref) if s_b.numel() == s_ref.numel() else "size differs"
        cnt = (r_ref[1:] - r_ref[:-1]).float()
        out["splats_per_tile_mean"], out["splats_per_tile_max"] = float(cnt.mean()), float(cnt.max())
        state.update(uvk=uvk, xk=xk, ck=ck, sorted=s_ref, ranges=r_ref, keep=keep)
        return out

    section("binning", binning)

    # ---- render forward / backward on identical inputs ------------------------------------------
    def render():
        out = {}
        keep = state["keep"]
        uvk, ck, s_idx, rng = state["uvk"], state["ck"], state["sorted"], state["ranges"]
        M = uvk.shape[0]
        opa = torch.sigmoid(g.opacity[keep]).contiguous()
        rgb = g.rgb[keep].contiguous()
        rays = torch.zeros(1, 1, 1, device=dev)
        res = {}
        for name, e in (("ref", ref_ext), ("b200", ext)):
            img = torch.zeros(H, W, 3, device=dev)
            n = torch.zeros(H, W, dtype=torch.int32, device=dev)
            w = torch.zeros(H, W, device=dev)
            e.render_tiles_cuda(uvk, opa, rgb, ck, rays, rng, s_idx, bg, n, w, img)
            res[name] = (img, n, w)
        out["image"] = bits_equal(res["b200"][0], res["ref"][0])
        out["num_splats_per_pixel"] = bits_equal(res["b200"][1], res["ref"][1])
        out["final_weight_per_pixel"] = bits_equal(res["b200"][2], res["ref"][2])
        out["mean_n_per_pixel"] = float(res["ref"][1].float().mean())
        # second reference run: its own nondeterminism (H % 16 != 0 race)
        img2 = torch.zeros(H, W, 3, device=dev)
        n2 = torch.zeros(H, W, dtype=torch.int32, device=dev)
        w2 = torch.zeros(H, W, device=dev)
        ref_ext.render_tiles_cuda(uvk, opa, rgb, ck, rays, rng, s_idx, bg, n2, w2, img2)
        out["ref_self_image"] = bits_equal(img2, res["ref"][0])
        G = synth.make_upstream_grad(args.res, device=dev)
        grads = {}
        for name, e in (("ref", ref_ext), ("b200", ext)):
            img, n, w = res["ref"]  # identical saved state for both
            gr, go, gu, gc = (torch.zeros(M, 3, device=dev), torch.zeros(M, 1, device=dev),
                              torch.zeros(M, 2, device=dev), torch.zeros(M, 3, device=dev))
            e.render_tiles_backward_cuda(uvk, opa, rgb, ck, rays, rng, s_idx, bg, n, w, G, gr, go, gu, gc)
            grads[name] = (gr, go, gu, gc)
        for k, nm in enumerate(("grad_rgb", "grad_opacity", "grad_uv", "grad_conic")):
            out[nm] = rel_err(grads["b200"][k], grads["ref"][k])
        gr2 = [torch.zeros_like(t) for t in grads["ref"]]
        ref_ext.render_tiles_backward_cuda(uvk, opa, rgb, ck, rays, rng, s_idx, bg, res["ref"][1], res["ref"][2], G,
                                           *gr2)
        out["ref_self_noise_grad_conic"] = rel_err(gr2[3], grads["ref"][3])
        if not args.no_timing:
            def fwd(e):
                img = torch.zeros(H, W, 3, device=dev)
                n = torch.zeros(H, W, dtype=torch.int32, device=dev)
                w = torch.zeros(H, W, device=dev)
                e.render_tiles_cuda(uvk, opa, rgb, ck, rays, rng, s_idx, bg, n, w, img)

            def bwd(e):
                bufs = (torch.zeros(M, 3, device=dev), torch.zeros(M, 1, device=dev),
                        torch.zeros(M, 2, device=dev), torch.zeros(M, 3, device=dev))
                e.render_tiles_backward_cuda(uvk, opa, rgb, ck, rays, rng, s_idx, bg, res["ref"][1], res["ref"][2],
                                             G, *bufs)
            out["t_fwd_ref"] = timed(lambda: fwd(ref_ext))
            out["t_fwd_b200"] = timed(lambda: fwd(ext))
            out["t_bwd_ref"] = timed(lambda: bwd(ref_ext), iters=5, warmup=1)
            out["t_bwd_b200"] = timed(lambda: bwd(ext), iters=5, warmup=1)
        return out

    section("render", render)

    # ---- end to end: rasterize() ----------------------------------------------------------------------
    def e2e():
        out = {}
        G = synth.make_upstream_grad(args.res, device=dev)

        def run(which):
            gg = synth.make_gaussians(args.n, args.res, sh_degree=args.sh, seed=0, device=dev, requires_grad=True)
            gaus = ref_structs.Gaussians(gg.xyz, gg.rgb, gg.opacity, gg.scale, gg.quaternion, gg.sh)
            camr = ref_structs.Camera(cam.width, cam.height, cam.K)
            fn = {"ref": ref_ras.rasterize, "b200": rasterize_b200, "b200_unfused": rasterize_unfused}[which]
            image, mask, uv = fn(gaus, T, camr, cfg["near_thresh"], cfg["far_thresh"], cfg["cull_mask_padding"],
                                 cfg["mh_dist"], True, bg)
            uv.retain_grad()
            image.backward(G)
            grads = dict(xyz=gg.xyz.grad, quaternion=gg.quaternion.grad, scale=gg.scale.grad,
                         opacity=gg.opacity.grad, rgb=gg.rgb.grad, uv=uv.grad)
            if gg.sh is not None:
                grads["sh"] = gg.sh.grad
            return image.detach(), mask, uv.detach(), grads

        r = run("ref")
        for which in ("b200", "b200_unfused"):
            b = run(which)
            o = dict(image=bits_equal(b[0], r[0]), culling_mask=bits_equal(b[1], r[1]), uv=bits_equal(b[2], r[2]))
            for k in r[3]:
                o["grad_" + k] = rel_err(b[3][k], r[3][k]) if b[3][k] is not None else "None"
            out[which] = o
        r2 = run("ref")
        out["ref_self_noise"] = {("grad_" + k): rel_err(r2[3][k], r[3][k]) for k in r[3]}
        out["ref_self_noise"]["image"] = bits_equal(r2[0], r[0])
        if not args.no_timing:
            gg = synth.make_gaussians(args.n, args.res, sh_degree=args.sh, seed=0, device=dev, requires_grad=True)
            gaus = ref_structs.Gaussians(gg.xyz, gg.rgb, gg.opacity, gg.scale, gg.quaternion, gg.sh)
            camr = ref_structs.Camera(cam.width, cam.height, cam.K)

            def step(fn):
                for p in (gg.xyz, gg.rgb, gg.opacity, gg.scale, gg.quaternion, gg.sh):
                    if p is not None:
                        p.grad = None
                image, _, _ = fn(gaus, T, camr, cfg["near_thresh"], cfg["far_thresh"], cfg["cull_mask_padding"],
                                 cfg["mh_dist"], True, bg)
                image.backward(G)
            out["t_step_ref"] = timed(lambda: step(ref_ras.rasterize), iters=5, warmup=2)
            out["t_step_b200"] = timed(lambda: step(rasterize_b200), iters=10, warmup=3)
            out["t_step_b200_unfused"] = timed(lambda: step(rasterize_unfused), iters=5, warmup=2)
        return out

    section("rasterize_e2e", e2e)

    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(rep, indent=1))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
