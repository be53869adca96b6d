set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -4 gpurun_out/r2u_pytest.log
timeout 300 python tools/stage_bench.py --tag pyr2 > gpurun_out/r2u_stage_new.json 2> gpurun_out/r2u_stage_new.err
timeout 300 python tools/stage_bench.py --tag pyr2_480 --workload gme_480p > gpurun_out/r2u_stage_480.json 2>> gpurun_out/r2u_stage_new.err
timeout 300 python tools/stage_bench.py --tag pyr2_4k --workload gme_4k_exh32 > gpurun_out/r2u_stage_4k.json 2>> gpurun_out/r2u_stage_new.err
cat gpurun_out/r2u_stage_*.json
