"""Per-kernel shares of a step from an ncu launch list, next to bench.py's live stage times.

    python tools/launch_summary.py profiles/r02_ncu_launches_bench_1080p.csv profiles/r02_bench_1080p.json \
        > profiles/r02_ncu_launches_bench_1080p.summary.txt

The launch list is `ncu --metrics gpu__time_duration.sum --clock-control none --csv` of a short bench.py run (see
profiles/README.md).  Launches are grouped by (kernel, grid); a group belongs to the pipeline calls of n pairs, n read
off grid.z (pyramid launches run once per frame: z = n + frame distance).  ncu's times are cold-cache and serialised:
the SHARES are what is compared with the live stage times, not the absolutes.
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict

DISTANCE = 3
STAGE_OF = [("pyr_down_kernel", "pyramids"), ("bbme_diamond2_kernel", "bbme_dense_l0"), ("affine_fit_kernel", "fit"),
            ("compensate16_kernel", "compensate_psnr")]


def main(csv_path, bench_path=None):
    rows = []
    with open(csv_path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        grid = tuple(int(x) for x in re.findall(r"\d+", r["Grid Size"]))
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("gme::", "")
        rows.append((name, grid, float(r["Metric Value"]) / 1e3))
    groups = OrderedDict()
    for name, grid, us in rows:
        groups.setdefault((name, grid), []).append(us)
    by_n = defaultdict(list)
    other = []
    for (name, grid), v in groups.items():
        is_pipe = any(k in name for k in ("pyr_down", "diamond", "bbme_pattern", "affine_fit", "compensate16"))
        if not is_pipe:
            other.append((name, grid, v))
            continue
        n = grid[0] if "affine_fit" in name else grid[2]
        if "pyr_down" in name:                        # one launch per FRAME: n + distance frames, or 2 for a lone pair
            n = 1 if n == 2 else n - DISTANCE
        by_n[n].append((name, grid, v))
    print("ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
    print("(cold-cache, serialised launch times: compare SHARES with bench.py's live stage timing, not absolutes)\n")
    for n in sorted(by_n, reverse=True):
        total = sum(sum(v) / len(v) for _, _, v in by_n[n])
        what = " (the per-pair drop-in surface)" if n == 1 else ""
        print(f"pipeline calls on {n} pairs{what}: {len(by_n[n])} kernels, sum of mean durations {total:.1f} us")
        for name, grid, v in by_n[n]:
            m = sum(v) / len(v)
            print(f"   {m:9.1f} us  {100 * m / total:5.1f}%  x{len(v):<3d} {name} {grid}")
        print()
    if bench_path:
        b = json.load(open(bench_path))
        st = b["stages"]
        total = sum(s["ms_per_step"] for s in st.values())
        print(f"live stage timing of the plain run ({bench_path}; eager single-lane pass, {b['config'].get('pairs_per_step_per_gpu')} pairs):")
        for k, s in st.items():
            print(f"   {1e3 * s['ms_per_step']:9.1f} us  {100 * s['ms_per_step'] / total:5.1f}%  {k}")
        print(f"   the timed step itself ({b.get('launch', 'one gme_pipeline call')}): {1e3 * b['ms_per_step']:.1f} us\n")
    if other:
        print("other launches (bench.py's separate 'bbme_exhaustive' measurement, probes):")
        for name, grid, v in other:
            print(f"   {sum(v) / len(v):9.1f} us  x{len(v):<3d} {name} {grid}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
