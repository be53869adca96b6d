set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GME_FUZZ_SCALE=12 timeout 1200 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2w_fuzz.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_fuzz.log
tail -4 gpurun_out/r2w_fuzz.log
timeout 600 python bench.py > gpurun_out/r02_bench_1080p.json 2> gpurun_out/r02_bench_1080p.err; echo "rc=$?"
timeout 600 python bench.py --workload gme_480p > gpurun_out/r02_bench_480p.json 2> gpurun_out/r02_bench_480p.err; echo "rc=$?"
for w in gme_1080p_3step gme_1080p_2dlog gme_4k_exh32; do timeout 600 python bench.py --workload $w --steps 40 --no-cpu-baseline > gpurun_out/r02_bench_${w#gme_}.json 2> gpurun_out/r02_bench_${w#gme_}.err; echo "rc=$?"; done
timeout 300 python bench.py --impl dropin --steps 3 --warmup 1 > gpurun_out/r02_bench_dropin_1080p.json 2> gpurun_out/r02_dropin.err
timeout 300 python bench.py --impl dropin --steps 3 --warmup 1 --workload gme_480p > gpurun_out/r02_bench_dropin_480p.json 2>> gpurun_out/r02_dropin.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_1080p_reference_arm.json 2> gpurun_out/r02_ref.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv -k regex:"gme|pyr_down|bbme|affine|compensate|first_params|sse_kernel|sad_probe" --log-file gpurun_out/r02_ncu_launches_bench_1080p.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
python tools/stage_bench.py --steps 1 > gpurun_out/r02_plain2.json 2> gpurun_out/r02_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:"pyr_down|compensate|diamond|affine_fit" -s 21 -c 7 -o gpurun_out/r02_pipeline -f python tools/stage_bench.py --steps 1 > gpurun_out/r02_ncu_full.log 2>&1
python tools/run_exhaustive.py 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"bbme_exhaustive" -s 4 -c 2 -o gpurun_out/r02_exhaustive -f python tools/run_exhaustive.py 3 > gpurun_out/r02_ncu_exh.log 2>&1
ls -la gpurun_out/r02_*
