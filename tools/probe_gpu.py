"""One-off GPU probe (run under gpurun): which fp32 accumulation order does torch use on
this box for the matmuls on the seeker path?  Writes gpurun_out/probe_matmul.json."""
import itertools
import json
import os

import numpy as np
import torch

dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(4)
N = 2_000_000
A = (torch.randn(1, 3, 3, device=dev, generator=g) * torch.tensor([[1266., 1266., 1.]], device=dev).T).contiguous()
X = (torch.randn(3, N, device=dev, generator=g) * 20)
Y = A.matmul(X)[0]
Ad, Xd = A[0].double(), X.double()


def fma(a, b, c):
    return (a * b + c).float().double()   # fp64 emulation of an fp32 fma (double rounding ~1e-9 rare)


res = {}
for perm in itertools.permutations(range(3)):
    for first_fused in (False, True):
        acc = (Ad[:, perm[0], None] * Xd[perm[0]]).float().double()
        acc = fma(Ad[:, perm[1], None], Xd[perm[1]], acc)
        acc = fma(Ad[:, perm[2], None], Xd[perm[2]], acc)
        res["matmul_1x3x3@3xN order %s" % (perm,)] = int((acc.float() != Y).sum())
        break
# bmm (L,3,3)@(L,3,1)
L = 500_000
C = torch.randn(3, 3, device=dev, generator=g).expand(L, 3, 3).contiguous()
P = torch.randn(L, 3, 1, device=dev, generator=g) * 30
Z = C.matmul(P).squeeze(-1)
Cd, Pd = C.double(), P.double().squeeze(-1)
for perm in itertools.permutations(range(3)):
    acc = (Cd[:, :, perm[0]] * Pd[:, perm[0], None]).float().double()
    acc = fma(Cd[:, :, perm[1]], Pd[:, perm[1], None], acc)
    acc = fma(Cd[:, :, perm[2]], Pd[:, perm[2], None], acc)
    res["bmm_Lx3x3@Lx3x1 order %s" % (perm,)] = int((acc.float() != Z).sum())
# wider family for the bmm: exact products (fp64), rounded products, partial fusing
a = [Cd[:, :, k] for k in range(3)]
bb = [Pd[:, k, None] for k in range(3)]
ex = [a[k] * bb[k] for k in range(3)]                 # exact in fp64
rp = [e.float().double() for e in ex]                # rounded products
r32 = lambda t: t.float().double()
for i, j, k in itertools.permutations(range(3)):
    res["bmm nofma ((p%d+p%d)+p%d)" % (i, j, k)] = int((r32(r32(rp[i] + rp[j]) + rp[k]).float() != Z).sum())
    res["bmm fma(%d, rn(p%d+p%d))" % (i, j, k)] = int((r32(ex[i] + r32(rp[j] + rp[k])).float() != Z).sum())
    res["bmm rn(fma(%d,p%d))+p%d" % (i, j, k)] = int((r32(r32(ex[i] + rp[j]) + rp[k]).float() != Z).sum())
res["bmm exact-sum (fp64 accumulate)"] = int(((ex[0] + ex[1] + ex[2]).float() != Z).sum())
# does a plain elementwise formulation agree with itself?  (sanity of the emulation)
Zs = (C[:, :, 0] * P[:, 0] ).float()
res["n_matmul"] = 3 * N
res["n_bmm"] = 3 * L
# quantile lerp on CUDA vs CPU
x = torch.rand(100001, device=dev, generator=g) * 50
res["quantile_cuda_eq_cpu"] = bool(torch.quantile(x, 0.25).item() == torch.quantile(x.cpu(), 0.25).item())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe_matmul.json", "w"), indent=1)
print(json.dumps(res, indent=1))
