"""Per-stage device times of gme_pipeline on one bench workload (development aid; bench.py is the contract).

    python tools/stage_bench.py [--workload gme_1080p] [--steps 30] [--pairs N]

Prints one JSON line: ms per step and per stage (CUDA events recorded by gme_pipeline on its own stream).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "global-motion-estimation_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
import gme_device as D  # noqa: E402
import gme_native as N  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="gme_1080p")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--pairs", type=int, default=0)
ap.add_argument("--tag", default="")
args = ap.parse_args()
H, W, motion, procedure, window, pairs, _ = bench.WORKLOADS[args.workload]
pairs = args.pairs or pairs
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
nf = pairs + bench.DISTANCE
seq = bench.make_sequence(nf, H, W, motion, seed=4, device=dev)
planes = D.Planes.empty(nf, H, W, dev)
planes.pixels().copy_(seq)
prev, cur = planes.view(0, pairs), planes.view(bench.DISTANCE, nf)
pipe = D.Pipeline(pairs, H, W, dev)
for _ in range(3):
    pipe.run(prev, cur, procedure, window)
torch.cuda.synchronize()
N.stage_timing_enable(True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps):
    pipe.run(prev, cur, procedure, window)
b.record()
torch.cuda.synchronize()
ms, calls = N.stage_timing_read()
N.stage_timing_enable(False)
out = {"tag": args.tag, "workload": args.workload, "pairs": pairs, "ms_per_step": a.elapsed_time(b) / args.steps,
       "pairs_per_s": pairs * args.steps / (a.elapsed_time(b) * 1e-3),
       "stages_us": {k: round(1e3 * v / calls, 1) for k, v in zip(N.STAGE_NAMES, ms)},
       "checksum": [float(pipe.params.sum().item()), int(pipe.sse.sum().item())]}
print(json.dumps(out))
