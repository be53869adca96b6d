set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -4 gpurun_out/r2r_pytest.log
timeout 300 python tools/exh_bench.py > gpurun_out/r2r_exh.json 2>> gpurun_out/r2r_exh.err; cut -c150-800 gpurun_out/r2r_exh.json
