set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
python tools/stage_bench.py --tag new4 > gpurun_out/r2e_stage_new.json 2> gpurun_out/r2e_stage_new.err
cat gpurun_out/r2e_stage_*.json
