set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -6 gpurun_out/r2k_pytest.log
timeout 600 python bench.py > gpurun_out/r2k_bench_1080p.json 2> gpurun_out/r2k_bench_1080p.err; echo "bench rc=$?"
tail -3 gpurun_out/r2k_bench_1080p.err
cat gpurun_out/r2k_bench_1080p.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err
cat gpurun_out/r2k_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -2 gpurun_out/r2k_smoke.log
