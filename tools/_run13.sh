set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 150 python tools/probes/h2d_concurrent.py --gpus 1,2,4,8 > gpurun_out/r2m_h2d_probe_8.log 2>&1; cat gpurun_out/r2m_h2d_probe_8.log
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2m_bench_1080p_n8.json 2> gpurun_out/r2m_bench_n8.err; echo "rc=$?"
cut -c1-300 gpurun_out/r2m_bench_1080p_n8.json
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 30 --warmup 3 --workload gme_1080p_3step > gpurun_out/r2m_bench_1080p_3step_n8.json 2> gpurun_out/r2m_bench_3step_n8.err; echo "rc=$?"
cut -c1-300 gpurun_out/r2m_bench_1080p_3step_n8.json
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 30 --warmup 3 --workload gme_1080p_2dlog > gpurun_out/r2m_bench_1080p_2dlog_n8.json 2> gpurun_out/r2m_bench_2dlog_n8.err; echo "rc=$?"
cut -c1-300 gpurun_out/r2m_bench_1080p_2dlog_n8.json
timeout 420 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2m_multirank8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_multirank8.log
tail -5 gpurun_out/r2m_multirank8.log
