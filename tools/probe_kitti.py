"""One-off GPU probe (run under gpurun): fp32 accumulation order of the matmuls of the KITTI head's calibration
(pcdet/utils/calibration_kitti.py:151-216, CalibrationTorch): (N,4)@(4,3) for lidar_to_rect and rect_to_img,
(N,4)@(4,4) for rect_to_lidar, at the row counts the head calls them with (8 frustum corners, 8 * hypotheses box
corners, the points of a frame, the points of a frustum).  Writes gpurun_out/probe_kitti.json."""
import itertools
import json
import os

import numpy as np
import torch

dev = "cuda:0"
g = torch.Generator(device=dev)
g.manual_seed(11)


def r32(t):
    return t.float().double()


def variants(A, B):
    """A (N,K) , B (K,C) float64 copies of fp32 data -> {name: (N,C) float32}"""
    K = A.shape[1]
    ex = [A[:, k, None] * B[k][None, :] for k in range(K)]          # exact products (fp64)
    out = {}
    for perm in itertools.permutations(range(K)):
        acc = r32(ex[perm[0]])
        for k in perm[1:]:
            acc = r32(ex[k] + acc)                                  # fma(a_k, b_k, acc)
        out["fma chain %s" % (perm,)] = acc.float()
    rp = [r32(e) for e in ex]
    if K == 4:
        for (i, j, k, l) in [(0, 1, 2, 3), (0, 2, 1, 3), (0, 3, 1, 2)]:
            out["pairs rn(rn(p%d+p%d)+rn(p%d+p%d))" % (i, j, k, l)] = r32(r32(rp[i] + rp[j]) + r32(rp[k] + rp[l])).float()
            out["pairs fma(%d,p%d)+fma(%d,p%d)" % (i, j, k, l)] = r32(r32(ex[i] + rp[j]) + r32(ex[k] + rp[l])).float()
        out["sequential no fma"] = r32(r32(r32(rp[0] + rp[1]) + rp[2]) + rp[3]).float()
    out["fp64 accumulate"] = sum(ex).float()
    return out


V2C = torch.tensor([[7.5e-3, -0.99997, -1e-3, 0.004], [1.2e-2, 1e-3, -0.99993, -0.076], [0.99989, 7.5e-3, 1.2e-2, -0.272]], device=dev)
R0 = torch.tensor([[0.99992, 9.8e-3, -7.4e-3], [-9.9e-3, 0.99994, -4.3e-3], [7.4e-3, 4.4e-3, 0.99996]], device=dev)
P2 = torch.tensor([[721.54, 0, 609.56, 44.857], [0, 721.54, 172.85, 0.2163], [0, 0, 1, 0.002746]], device=dev)
M1 = V2C.T @ R0.T                                     # (4,3), as the head forms it
R0e = torch.cat((torch.cat((R0, R0.new_zeros((3, 1))), dim=1), R0.new_zeros((1, 4))), dim=0); R0e[3, 3] = 1
V2Ce = torch.cat((V2C, V2C.new_zeros((1, 4))), dim=0); V2Ce[3, 3] = 1
Minv = torch.inverse(torch.matmul(R0e, V2Ce).T)       # (4,4)

res = {}
for N in (8, 48 * 8, 240 * 8, 11381, 120000, 7, 33, 1000):
    pts = torch.randn(N, 3, device=dev, generator=g) * torch.tensor([25.0, 12.0, 1.5], device=dev) + torch.tensor([30.0, 0, -0.5], device=dev)
    hom = torch.cat((pts, pts.new_ones((N, 1))), dim=1)
    rect = hom @ M1
    rhom = torch.cat((rect, rect.new_ones((N, 1))), dim=1)
    img = rhom @ P2.T
    lid = torch.matmul(rhom, Minv)
    for name, A, B, Y in (("lidar_to_rect (N,4)@(4,3)", hom, M1, rect), ("rect_to_img (N,4)@(4,3)", rhom, P2.T.contiguous(), img),
                          ("rect_to_lidar (N,4)@(4,4)", rhom, Minv, lid)):
        v = variants(A.double(), B.double())
        hits = {k: int((x != Y).sum()) for k, x in v.items()}
        best = sorted(hits.items(), key=lambda kv: kv[1])[:3]
        res["N=%d %s" % (N, name)] = dict(best=best, elements=int(Y.numel()))
# M1 itself: (4,3)@(3,3) on the device vs variants
v = variants(V2C.T.contiguous().double(), R0.T.contiguous().double())
res["M1 = V2C.T @ R0.T (4,3)@(3,3)"] = sorted({k: int((x != M1).sum()) for k, x in v.items()}.items(), key=lambda kv: kv[1])[:3]
# is P2.T a view (non-contiguous)?  the head multiplies by the view; same result as by the contiguous copy?
pts = torch.randn(5000, 4, device=dev, generator=g)
res["P2.T view == contiguous"] = bool(torch.equal(pts @ P2.T, pts @ P2.T.contiguous()))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe_kitti.json", "w"), indent=1)
print(json.dumps(res, indent=1))
