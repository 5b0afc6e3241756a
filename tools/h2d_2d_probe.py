"""Probe (run under gpurun): can the copy engine move only the xyz columns of a pinned (N,5) f32
table (cudaMemcpy2DAsync, width 12 B, source pitch 20 B) faster than the whole table?"""
import ctypes as C
import torch

rt = C.CDLL("libcudart.so.12")
N = 10_300_000
src = torch.empty((N, 5), dtype=torch.float32, pin_memory=True)
src.normal_()
dst_full = torch.empty((N, 5), dtype=torch.float32, device="cuda")
dst_xyz = torch.empty((N, 3), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_full = timed(lambda: dst_full.copy_(src, non_blocking=True))
print("full (N,5) copy: %.3f ms, %.1f GB/s" % (t_full, N * 20 / t_full / 1e6))
for rows_per_call in (N, 1 << 20, 1 << 16):
    def f():
        for r0 in range(0, N, rows_per_call):
            n = min(rows_per_call, N - r0)
            rc = rt.cudaMemcpy2DAsync(dst_xyz.data_ptr() + r0 * 12, 12, src.data_ptr() + r0 * 20, 20, 12, n, 1, st)
            assert rc == 0, rc
    t = timed(f, 1)
    print("2D xyz copy, %d rows/call: %.3f ms, payload %.1f GB/s" % (rows_per_call, t, N * 12 / t / 1e6))
assert torch.equal(dst_xyz.cpu(), src[:, :3])
# wider rows for reference: treat the table as rows of 4 points (80 B pitch) -- not usable, xyz is not contiguous there
