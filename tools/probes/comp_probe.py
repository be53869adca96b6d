"""Where does the compensation kernel stand against a plain device copy of the same bytes?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "global-motion-estimation_b200")]
import torch
import gme_device as D

n, H, W = 64, 1080, 1920
seq = torch.randint(0, 255, (n + 3, H, W), dtype=torch.uint8, device="cuda")
planes = D.Planes(seq, W)
prev, cur = planes.view(0, n), planes.view(3, n + 3)
field = torch.zeros((n, H // 16, W // 16, 2), dtype=torch.int16, device="cuda")
field[..., 0] = -2
field[..., 1] = -7
zero = torch.zeros_like(field)
out = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


print("compensate + sse (model field -2,-7):", round(timed(lambda: D.compensate(prev, field, cur)), 1), "us")
print("compensate only                     :", round(timed(lambda: D.compensate(prev, field)), 1), "us")
print("compensate + sse, zero field        :", round(timed(lambda: D.compensate(prev, zero, cur)), 1), "us")
print("sse only                            :", round(timed(lambda: D.sse(prev, cur)), 1), "us")
print("torch copy of n frames              :", round(timed(lambda: out.copy_(seq[:n])), 1), "us")
