set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -4 gpurun_out/r2q_pytest.log
timeout 300 python tools/exh_bench.py > gpurun_out/r2q_exh.json 2>> gpurun_out/r2q_exh.err; cut -c150-800 gpurun_out/r2q_exh.json
timeout 300 python tools/stage_bench.py --tag 4k --workload gme_4k_exh32 > gpurun_out/r2q_stage_4k.json 2>> gpurun_out/r2q_stage.err; cat gpurun_out/r2q_stage_4k.json
