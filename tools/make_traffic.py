"""profiles/traffic.json from ncu captures: DRAM bytes (read + write) per launch of every pipeline stage, per workload.

    # on the GPU box (tools/capture_traffic.sh does this for every workload of bench.py):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:"pyr_down|bbme|affine_fit|compensate" -s <launches of the warm-up steps> -c <launches of one step> \
        --csv --log-file gpurun_out/traffic_<workload>.csv python tools/stage_bench.py --workload <workload> --steps 1
    # here:
    python tools/make_traffic.py gpurun_out/traffic_*.csv > profiles/traffic.json

bench.py reads the file for `roofline.traffic` (ncu cannot run inside the timed bench).  The stage of a launch is its
position in the step: pyramids (2 launches for a sequence), dense L0, L1, L2 block matching, fit, compensation.
"""
import csv
import json
import os
import re
import sys

STAGES = ["pyramids", "pyramids", "bbme_dense_l0", "bbme_l1", "bbme_l2", "fit", "compensate_psnr"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def parse(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    i_id, i_name, i_metric, i_unit, i_val = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[i_id]), {"kernel": r[i_name]})
        v = float(r[i_val].replace(",", ""))
        if r[i_metric].startswith("dram__bytes"):
            d[r[i_metric]] = v * UNIT[r[i_unit]]
        elif r[i_metric].startswith("gpu__time"):
            d[r[i_metric]] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[i_unit], 1.0)
        else:
            d[r[i_metric]] = v
    return [launches[k] for k in sorted(launches)]


def main(paths):
    out = {}
    for p in paths:
        workload = re.sub(r"^traffic_|\.csv$", "", os.path.basename(p))
        ls = parse(p)
        if len(ls) != len(STAGES):
            raise SystemExit(f"{p}: expected {len(STAGES)} launches of one step, found {len(ls)}")
        per = {}
        for stage, l in zip(STAGES, ls):
            d = per.setdefault(stage, {"dram_bytes": 0.0, "us_under_ncu": 0.0, "kernels": []})
            d["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
            d["us_under_ncu"] += l.get("gpu__time_duration.sum", 0.0)
            d["kernels"].append(l["kernel"].split("(")[0])
            # what bounds the kernel: issue-slot utilisation and pipe utilisation of its longest launch
            if l.get("gpu__time_duration.sum", 0.0) >= d.get("_longest", 0.0):
                d["_longest"] = l.get("gpu__time_duration.sum", 0.0)
                for key, name in (("smsp__issue_active.avg.per_cycle_active", "issue_active"),
                                  ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
                                  ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
                                  ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefronts_pct"),
                                  ("smsp__inst_executed.sum", "warp_instructions")):
                    if key in l:
                        d[name] = l[key]
        for v in per.values():
            v.pop("_longest", None)
        out[workload] = {k: int(v["dram_bytes"]) for k, v in per.items()}
        out[workload]["_detail"] = per
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1:])
