import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
from findnpropagate_b200 import synth
from findnpropagate_b200.seeker import FrameInput, SeekerEngine
cfg = synth.CONFIGS["cfg2"]; params = synth.seeker_params(cfg)
f = synth.make_frame(0, cfg, device="cuda:0")
fi = FrameInput(points=f.points, lidar2image=f.lidar2image, camera2lidar=f.camera2lidar, camera_intrinsics=f.camera_intrinsics, det_boxes=f.det_boxes, det_labels=f.det_labels, det_scores=f.det_scores, det_cam_idx=f.det_cam_idx, gt_boxes=f.gt_boxes)
eng = SeekerEngine(params, device="cuda:0")
plan = eng.plan([fi]); h = eng.execute(plan, eng.upload_points([fi])); res = eng.finish(h)
torch.cuda.synchronize()
W = h["batch"].mask_words
n_cells = 25*15
buf = eng.arena.bufs["cell_masks"]
off = 1*6*n_cells*W*4
tab = buf[off:off+64*4].view(torch.int32).cpu().numpy().astype(np.uint32)
print("W", W, "sector table:", [bin(int(x)).count("1") for x in tab], "avg", np.mean([bin(int(x)).count("1") for x in tab]))
# fraction of points in region and avg cams per point; per-warp OR
p = f.points[:, :3]
rho2 = p[:,0]**2 + p[:,1]**2
inreg = (rho2 >= 9) & (np.abs(p[:,2]) <= 16)
ax, ay = np.abs(p[:,0]), np.abs(p[:,1]); q = ay/(ax+ay); iq = np.minimum((q*16).astype(int), 15)
sec = np.where(p[:,0]<0,32,0) | np.where(p[:,1]<0,16,0) | iq
allr = 0
for r in range(6):
    if plan["cam_cand_start"][r+1] > plan["cam_cand_start"][r]: allr |= 1<<r
m = np.where(inreg, tab[sec], allr)
print("in region", inreg.mean(), "cams per point", np.mean([bin(int(x)).count("1") for x in m[:20000]]))
# per (warp, sub) OR: tile of 1024: sub s rows s*256+tid; warp = 32 consecutive
n = (len(m)//1024)*1024
mm = m[:n].reshape(-1, 4, 8, 32)
orr = np.bitwise_or.reduce(mm, axis=3)
print("cams per warp-subtile (OR)", np.mean([bin(int(x)).count("1") for x in orr.reshape(-1)[:20000]]))
