#!/bin/bash
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r3f_bench_n2.json 2> gpurun_out/r3f.err; echo "rc=$?"
tail -5 gpurun_out/r3f.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3f_bench_n2.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["launch"], d["parity"])
PY
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
