"""CPU: host-side logic -- synthetic generators, PSNR tail, pair sharding + gather with gloo (world size 2)."""
import hashlib
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

import gme_synth as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generators_are_deterministic():
    a, b = S.texture(64, 80, seed=5), S.texture(64, 80, seed=5)
    assert a.dtype == np.uint8 and a.shape == (64, 80) and np.array_equal(a, b)
    assert not np.array_equal(a, S.texture(64, 80, seed=6))
    assert a.std() > 20           # textured, not flat
    seq = S.pan_sequence(5, 48, 64, step=(2, 1), seed=3)
    assert np.array_equal(seq[0][1:, 2:], seq[1][:-1, :-2])      # exact crop pan
    z = S.zoom_rotate_sequence(3, 48, 64)
    assert z.shape == (3, 48, 64) and not np.array_equal(z[0], z[2])
    assert S.affine_sequence(3, 40, 56).shape == (3, 40, 56)
    prev, _ = S.pan_pair(240, 320, 5, -3)
    assert hashlib.sha256(prev.tobytes()).hexdigest()[:16] == "510979328e5d7435"


def test_psnr_tail_matches_reference_formula():
    import gme_device as D
    assert D.psnr_from_sse(0, 100) == -1
    v = D.psnr_from_sse(123456, 320 * 240)
    assert isinstance(v, complex) and v.imag == 0
    assert abs(v.real - 20 * np.log10(255.0 / np.sqrt(123456 / (320 * 240)))) < 1e-12


def test_shard_pairs_partition():
    from gme_distributed import shard_pairs
    for n in (0, 1, 7, 61, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_pairs(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, os.path.join(sys.argv[1], "global-motion-estimation_b200"))
    import torch, torch.distributed as dist
    from gme_distributed import run_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
    n_pairs = 7
    def compute(start, stop):       # stands in for the GPU pipeline: row k = [k, 10k, ..., PSNR-like]
        k = torch.arange(start, stop, dtype=torch.float64)
        return torch.stack([k * (j + 1) for j in range(7)], 1)
    rows = run_sharded(n_pairs, compute)
    want = torch.stack([torch.arange(n_pairs, dtype=torch.float64) * (j + 1) for j in range(7)], 1)
    assert torch.equal(rows, want), rows
    dist.destroy_process_group()
    print("ok")
""")


def test_sharded_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
