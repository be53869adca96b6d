set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -25 gpurun_out/r2d_pytest.log
python tools/stage_bench.py --tag new3 > gpurun_out/r2d_stage_new.json 2> gpurun_out/r2d_stage_new.err
python tools/stage_bench.py --tag new3_480 --workload gme_480p > gpurun_out/r2d_stage_new480.json 2>> gpurun_out/r2d_stage_new.err
cat gpurun_out/r2d_stage_*.json
