set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -4 gpurun_out/r2p_pytest.log
timeout 300 python tools/stage_bench.py --tag fit2 > gpurun_out/r2p_stage_new.json 2> gpurun_out/r2p_stage_new.err
timeout 300 python tools/stage_bench.py --tag fit2_480 --workload gme_480p > gpurun_out/r2p_stage_480.json 2>> gpurun_out/r2p_stage_new.err
cat gpurun_out/r2p_stage_*.json
timeout 300 python tools/exh_bench.py > gpurun_out/r2p_exh.json 2>> gpurun_out/r2p_exh.err; cut -c150-800 gpurun_out/r2p_exh.json
python tools/stage_bench.py --steps 1 --workload gme_1080p_3step > gpurun_out/r2p_plain3.json 2> gpurun_out/r2p_plain3.err &&
ncu --set full --clock-control none --import-source on -k regex:bbme_pattern_kernel -s 6 -c 2 -o gpurun_out/r2p_threestep -f python tools/stage_bench.py --steps 1 --workload gme_1080p_3step > gpurun_out/r2p_ncu3.log 2>&1
