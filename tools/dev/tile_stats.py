"""Per-tile list length statistics of the bench scene (design input for the binning stage)."""
import json, sys
sys.path.insert(0, ".")
import torch
from gaussian_splatting_b200 import synth
from gaussian_splatting_b200.rasterize import rasterize
dev = torch.device("cuda")
g = synth.make_gaussians(3_000_000, "1080p", sh_degree=3, seed=0, device=dev)
cam = synth.make_camera("1080p", device=dev)
bg = torch.full((3,), 0.5, device=dev)
out = {}
for v in (0, 3, 7):
    with torch.no_grad():
        _, _, _, st = rasterize(g, synth.make_pose(v, 8, device=dev), cam, 0.3, 500.0, 100, 3.0, True, bg, return_state=True)
    c = (st.ranges[1:] - st.ranges[:-1]).float()
    q = torch.quantile(c, torch.tensor([0.5, 0.9, 0.99, 0.999], device=dev)).tolist()
    out[f"view{v}"] = dict(P=st.P, M=st.M, mean=float(c.mean()), max=int(c.max()), q50_90_99_999=q,
                           tiles_over_1024=int((c > 1024).sum()), tiles_over_2048=int((c > 2048).sum()),
                           tiles_over_4096=int((c > 4096).sum()))
print(json.dumps(out))
