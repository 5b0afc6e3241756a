"""Timeline probe of the end-to-end step (run under gpurun): when do the H2D copy and the
kernels of each step start/end on the device, and where is the host?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_frames
from findnpropagate_b200.seeker import SeekerEngine

dev = torch.device("cuda", 0)
frames, params = make_frames("cfg2", 0, 8, str(dev))
B = 32
batch = [frames[i % 8] for i in range(B)]
eng = SeekerEngine(params, device=dev)
rows = sum(f.points.shape[0] for f in batch)
pinned = []
for s in range(2):
    t = torch.empty((rows, 5), dtype=torch.float32, pin_memory=True)
    r = 0
    for f in batch:
        t[r:r + f.points.shape[0]] = torch.from_numpy(f.points); r += f.points.shape[0]
    pinned.append(t)
dev_pts = [p.to(dev) for p in pinned]
gt = eng.upload_gt(batch)
cs = torch.cuda.Stream(device=dev)
consumed = [torch.cuda.Event(), torch.cuda.Event()]
ready = [torch.cuda.Event(), torch.cuda.Event()]
for e in consumed: e.record()
N = 12
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(N)]
host = []
prev = None
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record()
t0 = time.perf_counter()
for k in range(N):
    h0 = time.perf_counter() - t0
    with torch.cuda.stream(cs):
        cs.wait_event(consumed[k % 2])
        ev[k][0].record(cs)
        dev_pts[k % 2].copy_(pinned[k % 2], non_blocking=True)
        ev[k][1].record(cs)
        ready[k % 2].record(cs)
    h1 = time.perf_counter() - t0
    plan = eng.plan(batch)
    h2 = time.perf_counter() - t0
    ev[k][2].record()
    h = eng.execute(plan, dev_pts[k % 2], nms_thresh=0.1, gt=gt, slot=k % 2, points_ready=ready[k % 2])
    ev[k][3].record()
    consumed[k % 2].record()
    h3 = time.perf_counter() - t0
    if prev is not None:
        eng.finish(prev)
    h4 = time.perf_counter() - t0
    prev = h
    host.append((h0, h1, h2, h3, h4))
eng.finish(prev)
torch.cuda.synchronize()
for k in range(N):
    d = [base.elapsed_time(e) for e in ev[k]]
    print("step %2d host: start %.2f copy-issued %.2f planned %.2f executed %.2f finished(prev) %.2f | dev: copy %.2f-%.2f  compute %.2f-%.2f"
          % ((k,) + tuple(1e3 * x for x in host[k]) + tuple(d)))
