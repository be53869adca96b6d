"""GPU bring-up aid: runs each C-ABI entry point in its own process with CUDA_LAUNCH_BLOCKING=1 so that
a faulting kernel is named precisely.  Usage: python tools/diag_gpu.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["pyr", "diamond2", "diamond16", "threestep16", "twodlog16", "exh12", "exh16mse", "fit", "comp", "sse",
         "exh_generic", "pattern_generic"]


def run_case(name):
    sys.path[:0] = [os.path.join(ROOT, "global-motion-estimation_b200"), os.path.join(ROOT, "oracle")]
    import numpy as np
    import torch
    import gme_device as D
    import gme_oracle as O
    import gme_synth as S
    prev, cur = S.pan_pair(240, 320, 5, -3)
    pp, cp = D.Planes.from_host(prev), D.Planes.from_host(cur)

    def bb(bs, sw, sp, pn):
        got = D.motion_field(pp, cp, bs, sw, sp, pn)
        torch.cuda.synchronize()
        want = O.get_motion_field(prev, cur, bs, sw, sp, pn)
        return np.array_equal(got[0].cpu().numpy(), want)
    if name == "pyr":
        d = D.pyr_down(pp); torch.cuda.synchronize()
        ok = np.array_equal(d.to_host()[0], O.pyr_down(prev))
    elif name == "diamond2": ok = bb(2, 2, 3, 1)
    elif name == "diamond16": ok = bb(16, 2, 3, 1)
    elif name == "threestep16": ok = bb(16, 16, 1, 1)
    elif name == "twodlog16": ok = bb(16, 16, 2, 0)
    elif name == "exh12": ok = bb(12, 12, 0, 0)
    elif name == "exh16mse": ok = bb(16, 16, 0, 1)
    elif name == "exh_generic": ok = bb(10, 5, 0, 0)
    elif name == "pattern_generic": ok = bb(10, 5, 3, 0)
    elif name == "fit":
        gt = torch.from_numpy(O.get_motion_field(prev, cur, 16, 2, 3, 1)[None]).cuda()
        p = torch.tensor([[3.0, 0, 0, -2.0, 0, 0]], dtype=torch.float64, device="cuda")
        D.affine_fit(gt, (240, 320), p); torch.cuda.synchronize()
        ok = bool(torch.isfinite(p).all())
    elif name == "comp":
        f = torch.from_numpy(O.get_motion_field(prev, cur, 16, 2, 3, 1)[None]).cuda()
        c, s = D.compensate(pp, f, cp); torch.cuda.synchronize()
        want = O.compensate_frame(prev, f[0].cpu().numpy())
        ok = np.array_equal(c.to_host()[0], want) and int(s.item()) == O.sse(cur, want)
    elif name == "sse":
        s = D.sse(pp, cp); torch.cuda.synchronize()
        ok = int(s.item()) == O.sse(prev, cur)
    print(f"CASE {name}: {'OK' if ok else 'MISMATCH'}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    cases = sys.argv[1:] or CASES
    for c in cases:
        env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
        r = subprocess.run([sys.executable, __file__, "--one", c], env=env, capture_output=True, text=True, timeout=300)
        tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
        print(f"[{c}] rc={r.returncode} :: " + " | ".join(tail), flush=True)
