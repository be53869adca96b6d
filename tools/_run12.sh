set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2l_multirank2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_multirank2.log
tail -15 gpurun_out/r2l_multirank2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; echo "rc=$?"
tail -3 gpurun_out/r2l_bench_n2.err; cat gpurun_out/r2l_bench_n2.json | cut -c1-600
timeout 300 python tools/probes/h2d_concurrent.py --gpus 1,2 > gpurun_out/r2l_h2d_probe_2.log 2>&1; cat gpurun_out/r2l_h2d_probe_2.log
