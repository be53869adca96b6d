set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -6 gpurun_out/r2y_pytest.log
timeout 300 python tools/stage_bench.py --tag w16_3step --workload gme_1080p_3step > gpurun_out/r2y_stage_3step.json 2> gpurun_out/r2y.err
timeout 300 python tools/stage_bench.py --tag w16_2dlog --workload gme_1080p_2dlog > gpurun_out/r2y_stage_2dlog.json 2>> gpurun_out/r2y.err
timeout 300 python tools/stage_bench.py --tag base > gpurun_out/r2y_stage_base.json 2>> gpurun_out/r2y.err
cat gpurun_out/r2y_stage_*.json
