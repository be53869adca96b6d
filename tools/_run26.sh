set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -6 gpurun_out/r2z_pytest.log
timeout 300 python tools/stage_bench.py --tag comp128 > gpurun_out/r2z_stage_base.json 2>> gpurun_out/r2z.err
timeout 300 python tools/stage_bench.py --tag comp128_480 --workload gme_480p > gpurun_out/r2z_stage_480.json 2>> gpurun_out/r2z.err
timeout 300 python tools/stage_bench.py --tag comp128_3step --workload gme_1080p_3step > gpurun_out/r2z_stage_3step.json 2>> gpurun_out/r2z.err
cat gpurun_out/r2z_stage_*.json
