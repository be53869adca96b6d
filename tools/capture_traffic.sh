#!/bin/bash
# DRAM traffic + duration of every launch of ONE warm step of each bench workload (run under gpurun, one GPU).
# stage_bench.py runs 3 warm-up steps + `--steps 1`: 4 steps x 7 launches (+ the torch fills, filtered by -k).
set -x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for w in gme_1080p gme_480p gme_1080p_3step gme_1080p_2dlog gme_4k_exh32; do
  python tools/stage_bench.py --workload $w --steps 1 > gpurun_out/traffic_plain_$w.json 2> gpurun_out/traffic_plain_$w.err &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none \
      -k regex:"pyr_down|bbme|affine_fit|compensate" -s 21 -c 7 --csv --log-file gpurun_out/traffic_$w.csv \
      python tools/stage_bench.py --workload $w --steps 1 > gpurun_out/traffic_ncu_$w.log 2>&1
done
