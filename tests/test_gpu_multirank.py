"""GPU, more than one device: config 4 of BASELINE.json under real NCCL.

The 1080p zoom+rotate sequence is sharded by frame pair over every GPU of the box (one process per GPU, NCCL,
gme_distributed.run_sharded: contiguous pair ranges, one all-gather of [pairs, 7] float64 rows); the gathered rows
must be IDENTICAL -- bit for bit -- for 1, 2, 4 and 8 ranks (SURVEY 8(e) determinism check), for both config-4
searches (three-step and 2D-log, sw = 16 on the bs-16 levels), and two of the pairs are checked against the oracle.
Skips cleanly on a one-GPU box.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
root, port, rank, world, out, procedure, window, n_frames = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6]), int(sys.argv[7]), int(sys.argv[8])
sys.path[:0] = [os.path.join(root, "global-motion-estimation_b200")]
import numpy as np, torch, torch.distributed as dist
torch.cuda.set_device(rank)
import gme_device as D, gme_synth as S
from gme_distributed import run_sharded, shard_pairs
dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
H, W, d = 1080, 1920, 3
n_pairs = n_frames - d
seq = np.load(os.path.join(os.path.dirname(out), "seq.npy"), mmap_mode="r")

def compute(start, stop):                            # this rank holds only the frames its pair range needs
    rows = torch.zeros((stop - start, 7), dtype=torch.float64, device="cuda")
    if stop > start:
        planes = D.Planes.from_host(np.ascontiguousarray(seq[start:stop + d]))
        pipe = D.Pipeline(stop - start, H, W)
        pipe.run(planes.view(0, stop - start), planes.view(d, stop - start + d), procedure, window)
        assert int(pipe.status.abs().max().item()) == 0
        rows[:, :6] = pipe.params
        rows[:, 6] = pipe.sse.to(torch.float64)
    return rows

rows = run_sharded(n_pairs, compute)
torch.cuda.synchronize()
if rank == 0:
    np.save(out, rows.cpu().numpy())
dist.barrier()
dist.destroy_process_group()
'''


def _launch(tmp_path, world, procedure, window, n_frames):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / f"rows_sp{procedure}_w{world}.npy"
    script = tmp_path / "rank.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(world), str(out), str(procedure),
                               str(window), str(n_frames)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    logs = [p.communicate(timeout=900)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(l[-2000:] for l in logs)
    return np.load(out)


@pytest.mark.parametrize("procedure,window", [(1, 16), (2, 16)])
def test_config4_rows_identical_for_every_rank_count(tmp_path, procedure, window):
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least two GPUs (the 1-GPU determinism check over gloo is in test_gpu_dropin.py)")
    import gme_oracle as O
    import gme_synth as S
    n_frames = 19                                                     # 16 pairs at distance 3: two per rank on 8 GPUs
    seq = S.zoom_rotate_sequence(n_frames, 1080, 1920, seed=4)
    np.save(tmp_path / "seq.npy", seq)
    worlds = [w for w in (1, 2, 4, 8) if w <= n_gpus]
    rows = {w: _launch(tmp_path, w, procedure, window, n_frames) for w in worlds}
    assert rows[1].shape == (n_frames - 3, 7) and np.isfinite(rows[1]).all()
    for w in worlds[1:]:
        np.testing.assert_array_equal(rows[w], rows[1], err_msg=f"{w} ranks vs 1 rank")
    for k in (0, n_frames - 4):                                       # the first and the last pair against the oracle
        want = O.global_motion_estimation(seq[k], seq[k + 3], procedure=procedure, window=window, threads=8)
        np.testing.assert_allclose(rows[1][k, :6], want, atol=1e-9, rtol=1e-9)
        comp = O.compensate_frame(seq[k], O.get_motion_field_affine((1080 // 16, 1920 // 16), want))
        assert int(rows[1][k, 6]) == O.sse(seq[k + 3], comp)
