set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -4 gpurun_out/r2t_pytest.log
timeout 300 python tools/exh_bench.py > gpurun_out/r2t_exh.json 2>> gpurun_out/r2t_exh.err; cut -c150-800 gpurun_out/r2t_exh.json
bash tools/capture_traffic.sh > gpurun_out/r2t_traffic.log 2>&1
ls gpurun_out/traffic_gme_*.csv
