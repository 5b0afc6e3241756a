"""Where the end-to-end time goes on the host side: the threaded x,y,z gather alone, the H2D copy
alone, and both pipelined as HostPointFeeder runs them (one B200, 256 cfg2 frames = 82.7 M rows)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from findnpropagate_b200 import _lib  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 82_655_008
src = [torch.rand((rows, 5), dtype=torch.float32).pin_memory() for _ in range(2)]
stage = [torch.empty((rows, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
dev3 = [torch.empty((rows, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
dev5 = torch.empty((rows, 5), dtype=torch.float32, device="cuda")
cs = torch.cuda.Stream()
for nt in (8, 12, 15):
    def pack(k, wait=True):
        t = _lib.lib.fnp_host_pack_xyz_begin(src[k % 2].data_ptr(), rows, 5, 0, stage[k % 2].data_ptr(), nt)
        return t
    n = 6
    # gather alone
    t0 = time.perf_counter()
    for k in range(n):
        _lib.lib.fnp_host_pack_wait(pack(k))
    t_pack = (time.perf_counter() - t0) / n * 1e3
    # H2D alone (12 B/point and 20 B/point)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        dev3[k % 2].copy_(stage[k % 2], non_blocking=True)
    torch.cuda.synchronize(); t_h2d3 = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        dev5.copy_(src[k % 2], non_blocking=True)
    torch.cuda.synchronize(); t_h2d5 = (time.perf_counter() - t0) / n * 1e3
    # pipelined: gather k+1 while copy k
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tk = pack(0)
    ev = [None, None]
    for k in range(n):
        _lib.lib.fnp_host_pack_wait(tk)
        with torch.cuda.stream(cs):
            dev3[k % 2].copy_(stage[k % 2], non_blocking=True)
            e = torch.cuda.Event(); e.record(cs); ev[k % 2] = e
        if k + 1 < n:
            if ev[(k + 1) % 2] is not None:
                ev[(k + 1) % 2].synchronize()
            tk = pack(k + 1)
    torch.cuda.synchronize(); t_both = (time.perf_counter() - t0) / n * 1e3
    print("threads %2d: gather %.1f ms (%.0f GB/s in), H2D 12B %.1f ms, H2D 20B %.1f ms, pipelined gather+H2D %.1f ms per batch"
          % (nt, t_pack, rows * 20 / t_pack / 1e6, t_h2d3, t_h2d5, t_both))
