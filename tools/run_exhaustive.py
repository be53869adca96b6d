"""Runs the exhaustive-search kernel on the two bench shapes (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "global-motion-estimation_b200")]
import numpy as np
import torch
import gme_device as D
import gme_synth as S

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
tex = S.texture(1080 + 16, 1920 + 16, seed=7)
seq = np.stack([tex[k:k + 1080, 2 * k:2 * k + 1920] for k in range(8)])
big = D.Planes.from_host(seq)
small = D.Planes.from_host(np.stack([seq[k % 8, 100:340, 200:520] for k in range(259)]))
for _ in range(reps):
    D.motion_field(small.view(0, 256), small.view(3, 259), 12, 12, 0, 0)
    D.motion_field(big.view(0, 4), big.view(3, 7), 16, 32, 0, 1)
torch.cuda.synchronize()
print("ok")
