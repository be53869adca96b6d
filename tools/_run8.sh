set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -8 gpurun_out/r2h_pytest.log
timeout 300 python tools/stage_bench.py --tag lean2 > gpurun_out/r2h_stage_new.json 2> gpurun_out/r2h_stage_new.err
timeout 300 python tools/stage_bench.py --tag lean2_480 --workload gme_480p > gpurun_out/r2h_stage_480.json 2>> gpurun_out/r2h_stage_new.err
cat gpurun_out/r2h_stage_*.json
timeout 300 python bench.py --impl dropin --steps 2 --warmup 1 > gpurun_out/r2h_dropin_1080p.json 2> gpurun_out/r2h_dropin.err
timeout 300 python bench.py --impl dropin --steps 2 --warmup 1 --workload gme_480p > gpurun_out/r2h_dropin_480p.json 2>> gpurun_out/r2h_dropin.err
cat gpurun_out/r2h_dropin_*.json; tail -5 gpurun_out/r2h_dropin.err
