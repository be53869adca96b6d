set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -6 gpurun_out/r2n_pytest.log
timeout 300 python tools/exh_bench.py > gpurun_out/r2n_exh.json 2> gpurun_out/r2n_exh.err; cat gpurun_out/r2n_exh.json
timeout 300 python tools/stage_bench.py --tag d2_4cta > gpurun_out/r2n_stage_new.json 2> gpurun_out/r2n_stage_new.err
timeout 300 python tools/stage_bench.py --tag 4k --workload gme_4k_exh32 > gpurun_out/r2n_stage_4k.json 2>> gpurun_out/r2n_stage_new.err
cat gpurun_out/r2n_stage_*.json
