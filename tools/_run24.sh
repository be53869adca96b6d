set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for t in 32 16 8 4; do GME_D2_TBY=$t timeout 300 python tools/stage_bench.py --tag tby$t > gpurun_out/r2x_stage_tby$t.json 2>> gpurun_out/r2x.err; GME_D2_TBY=$t timeout 300 python tools/stage_bench.py --tag tby${t}_480 --workload gme_480p > gpurun_out/r2x_stage480_tby$t.json 2>> gpurun_out/r2x.err; done
cat gpurun_out/r2x_stage*.json
