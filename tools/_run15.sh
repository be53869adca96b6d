set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
timeout 300 python tools/stage_bench.py --tag d2static > gpurun_out/r2o_stage_new.json 2> gpurun_out/r2o_stage_new.err
timeout 300 python tools/stage_bench.py --tag d2static_480 --workload gme_480p > gpurun_out/r2o_stage_480.json 2>> gpurun_out/r2o_stage_new.err
cat gpurun_out/r2o_stage_*.json
for b in 384 320 256 192 128; do GME_EXH_BUDGET=$b timeout 300 python tools/exh_bench.py > gpurun_out/r2o_exh_$b.json 2>> gpurun_out/r2o_exh.err; echo $b; cut -c150-700 gpurun_out/r2o_exh_$b.json; done
