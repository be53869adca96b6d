set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -8 gpurun_out/r2g_pytest.log
timeout 300 python tools/stage_bench.py --tag lean > gpurun_out/r2g_stage_new.json 2> gpurun_out/r2g_stage_new.err
timeout 300 python tools/stage_bench.py --tag lean_480 --workload gme_480p > gpurun_out/r2g_stage_480.json 2>> gpurun_out/r2g_stage_new.err
cat gpurun_out/r2g_stage_*.json
timeout 300 python bench.py --impl dropin --steps 2 --warmup 1 > gpurun_out/r2g_dropin_1080p_before.json 2> gpurun_out/r2g_dropin.err
timeout 300 python bench.py --impl dropin --steps 2 --warmup 1 --workload gme_480p > gpurun_out/r2g_dropin_480p_before.json 2>> gpurun_out/r2g_dropin.err
cat gpurun_out/r2g_dropin_*.json
