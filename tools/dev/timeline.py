"""Kernel timeline of a few fwd+bwd steps (torch.profiler / CUPTI; nsys is not in the image): where are the gaps?

    python tools/dev/timeline.py gpurun_out/timeline.json

Writes, per step mode (resident, e2e), the list of device activities (name, start us, duration us) of the last
profiled step plus the idle time between consecutive activities.  Numbers under the profiler are for the SHAPE of
the timeline only, never bench values."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from gaussian_splatting_b200 import synth  # noqa: E402
from gaussian_splatting_b200.rasterize import rasterize  # noqa: E402
from gaussian_splatting_b200.structs import Camera  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json"
    dev = torch.device("cuda:0")
    g, cam, poses, poses_host, G, G_host, bg = bench.build_scene(dev)
    cfg = synth.DEFAULTS
    K_host = cam.K.cpu().pin_memory()
    image_host = torch.empty(cam.height, cam.width, 3).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    T_buf, K_buf, G_buf = torch.empty(4, 4, device=dev), torch.empty(3, 3, device=dev), torch.empty_like(G)
    cam_e2e = Camera(cam.width, cam.height, K_buf)

    def step_resident(i):
        bench.zero_grads(g)
        image, _, _ = rasterize(g, poses[i % 8], cam, cfg["near_thresh"], cfg["far_thresh"], cfg["cull_mask_padding"],
                                cfg["mh_dist"], True, bg)
        image.backward(G)

    def step_e2e(i):
        main_s = torch.cuda.current_stream()
        bench.zero_grads(g)
        with torch.cuda.stream(copy_stream):
            T_buf.copy_(poses_host[i % 8], non_blocking=True)
            K_buf.copy_(K_host, non_blocking=True)
            ev_small = copy_stream.record_event()
        main_s.wait_event(ev_small)
        image, _, _ = rasterize(g, T_buf, cam_e2e, cfg["near_thresh"], cfg["far_thresh"], cfg["cull_mask_padding"],
                                cfg["mh_dist"], True, bg)
        ev_img = main_s.record_event()
        with torch.cuda.stream(copy_stream):  # same order as bench.py's step_e2e
            G_buf.copy_(G_host, non_blocking=True)
            ev_grad = copy_stream.record_event()
            copy_stream.wait_event(ev_img)
            image_host.copy_(image.detach(), non_blocking=True)
        main_s.wait_event(ev_grad)
        image.backward(G_buf)
        torch.cuda.synchronize()

    result = {}
    for mode, fn in (("resident", step_resident), ("e2e", step_e2e)):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for i in range(4, 8):
                fn(i)
            torch.cuda.synchronize()
        evs = []
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                evs.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name[:70]))
        evs.sort()
        # split into steps at the k_preprocess_fwd launches; keep the last full step
        starts = [k for k, e in enumerate(evs) if "k_preprocess_fwd" in e[2]]
        if len(starts) >= 2:
            lo = starts[-2]
            # include the small kernels before preprocess (pose inverse etc.): walk back to the previous k_preprocess_bwd
            prev_end = max([k for k, e in enumerate(evs[:lo]) if "k_preprocess_bwd" in e[2]] or [-1]) + 1
            seg = evs[prev_end:starts[-1]]
        else:
            seg = evs
        t0 = seg[0][0]
        rows, busy_end, idle = [], None, 0.0
        for st, du, nm in seg:
            gap = 0.0 if busy_end is None else max(0.0, st - busy_end)
            idle += gap
            busy_end = st + du if busy_end is None else max(busy_end, st + du)
            rows.append(dict(t_us=round(st - t0, 1), dur_us=round(du, 1), gap_before_us=round(gap, 1), name=nm))
        result[mode] = dict(span_us=round(busy_end - t0, 1), idle_us=round(idle, 1), n_activities=len(rows), rows=rows)
        print(mode, "span", result[mode]["span_us"], "idle", result[mode]["idle_us"], "acts", len(rows))
    Path(out).parent.mkdir(parents=True, exist_ok=True)
    Path(out).write_text(json.dumps(result, indent=0))


if __name__ == "__main__":
    main()
