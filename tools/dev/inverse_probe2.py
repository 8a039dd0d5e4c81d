"""Stage-wise version of inverse_probe.py: pin the LU arithmetic against torch.linalg.lu_factor_ex, then the two
triangular solves against torch.linalg.lu_solve on torch's own factors, then the whole chain."""
import itertools
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from tools.dev.inverse_probe import f32, fma, poses  # noqa: E402


def lu(A, l_mode, fma_upd):
    A = A.astype(f32).copy()
    perm = list(range(4))
    for k in range(4):
        piv = k + int(np.argmax(np.abs(A[k:, k])))
        if piv != k:
            A[[k, piv]] = A[[piv, k]]
            perm[k], perm[piv] = perm[piv], perm[k]
        p = A[k, k]
        rp = f32(1.0) / p
        for i in range(k + 1, 4):
            l = f32(A[i, k] / p) if l_mode == "div" else f32(A[i, k] * rp)
            A[i, k] = l
            for j in range(k + 1, 4):
                A[i, j] = fma(-l, A[k, j], A[i, j]) if fma_upd else f32(A[i, j] - f32(l * A[k, j]))
    return A, perm


def solve(LU, perm, lower_fma, upper_mode, upper_fma, upper_order):
    b = np.zeros(4, f32)
    b[3] = 1.0
    b = b[perm]
    y = np.zeros(4, f32)
    for i in range(4):
        acc = b[i]
        for j in range(i):
            acc = fma(-LU[i, j], y[j], acc) if lower_fma else f32(acc - f32(LU[i, j] * y[j]))
        y[i] = acc
    x = np.zeros(4, f32)
    for i in range(3, -1, -1):
        acc = y[i]
        js = range(i + 1, 4) if upper_order == "asc" else range(3, i, -1)
        for j in js:
            acc = fma(-LU[i, j], x[j], acc) if upper_fma else f32(acc - f32(LU[i, j] * x[j]))
        if upper_mode == "div":
            x[i] = f32(acc / LU[i, i])
        elif upper_mode == "rcp":
            x[i] = f32(acc * (f32(1.0) / LU[i, i]))
    return x


def beq(a, b):
    return bool((np.asarray(a, f32).view(np.int32) == np.asarray(b, f32).view(np.int32)).all())


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/inverse_probe2.json"
    rng = np.random.default_rng(0)
    mats = poses(500, rng, False) + poses(100, rng, True)
    dev = torch.device("cuda:0")
    res = {}
    T = torch.from_numpy(np.stack(mats)).to(dev)
    # torch processes ONE 4x4 at a time on the product path (batch of one): do the same here
    LUs, pivs, invs, sols = [], [], [], []
    for M in mats:
        t = torch.from_numpy(M).to(dev)
        LU_t, piv_t, _ = torch.linalg.lu_factor_ex(t)
        LUs.append(LU_t.cpu().numpy())
        pivs.append(piv_t.cpu().numpy())
        invs.append(torch.linalg.inv_ex(t)[0].cpu().numpy())
        sols.append(torch.linalg.lu_solve(LU_t, piv_t, torch.eye(4, device=dev)).cpu().numpy())
    res["lu_solve_equals_inv"] = sum(beq(a, b) for a, b in zip(sols, invs))
    lu_rows = []
    for l_mode, fma_upd in itertools.product(("div", "rcp"), (True, False)):
        ok = 0
        for M, LU_t in zip(mats, LUs):
            A, _ = lu(M, l_mode, fma_upd)
            ok += beq(A, LU_t)
        lu_rows.append(dict(l_mode=l_mode, fma_upd=fma_upd, match=ok))
    res["lu"] = sorted(lu_rows, key=lambda d: -d["match"])
    # triangular solves on torch's own factors (pivots are 1-based row swaps applied in sequence)
    tri_rows = []
    for lower_fma, upper_mode, upper_fma, upper_order in itertools.product((True, False), ("div", "rcp"), (True, False),
                                                                           ("asc", "desc")):
        ok = 0
        for LU_t, piv_t, inv_t in zip(LUs, pivs, invs):
            perm = list(range(4))
            for k, pk in enumerate(piv_t):
                pk = int(pk) - 1
                perm[k], perm[pk] = perm[pk], perm[k]
            x = solve(LU_t.astype(f32), perm, lower_fma, upper_mode, upper_fma, upper_order)
            ok += beq(x[:3], inv_t[:3, 3])
        tri_rows.append(dict(lower_fma=lower_fma, upper_mode=upper_mode, upper_fma=upper_fma, upper_order=upper_order,
                             match=ok))
    res["tri"] = sorted(tri_rows, key=lambda d: -d["match"])
    res["n"] = len(mats)
    # a few raw examples for offline study
    res["examples"] = [dict(M=mats[i].tolist(), LU=LUs[i].tolist(), piv=pivs[i].tolist(), inv=invs[i].tolist())
                       for i in range(6)]
    json.dump(res, open(out, "w"), indent=1)
    print("n", len(mats), "lu_solve==inv", res["lu_solve_equals_inv"])
    print("LU:", res["lu"])
    for r in res["tri"][:5]:
        print("TRI:", r)


if __name__ == "__main__":
    main()
