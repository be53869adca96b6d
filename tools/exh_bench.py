"""Exhaustive-search throughput alone (development aid): bench.bench_exhaustive on a 1080p zoom+rotate sequence."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "global-motion-estimation_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
import gme_device as D  # noqa: E402
import gme_native as N  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
seq = bench.make_sequence(11, 1080, 1920, "zoomrot", seed=4, device=dev)
planes = D.Planes.empty(11, 1080, 1920, dev)
planes.pixels().copy_(seq)
print(json.dumps(bench.bench_exhaustive(D, N, torch, planes, dev)))
