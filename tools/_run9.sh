set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/stage_bench.py --steps 1 --tag d2prof > gpurun_out/r2i_plain.json 2> gpurun_out/r2i_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:diamond2 -s 3 -c 1 -o gpurun_out/r2i_d2 -f python tools/stage_bench.py --steps 1 --tag d2prof > gpurun_out/r2i_ncu.log 2>&1
bash tools/capture_traffic.sh > gpurun_out/r2i_traffic.log 2>&1
ls -la gpurun_out/traffic_*.csv
