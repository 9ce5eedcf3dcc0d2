#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun; one GPU).  Every ncu pass runs only after the same command
# exited 0 without the profiler.  Outputs go to gpurun_out/; scripts/make_profile_summary.py r2 turns them into profiles/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/bench_r2_plain.json 2> gpurun_out/bench_r2_plain.err || { echo "plain run failed"; exit 1; }
# 1. every launch of one step with its device time (cold-cache, serialised: compare shares, not absolutes)
# (only the library's kernels: bench.py generates its inputs in 256 Philox chunks per tensor, thousands of torch launches)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(np_|kh_|ll_|emit_|scan_|prune_|quant_|scal_|rs_|set_thr)" -c 400 --csv \
    --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
# 2. the streaming kernels and the Lloyd loop, full sections, second step (the first eight matching launches are step one)
ncu --set full --clock-control none --import-source on \
    -k regex:"np_tree_kernel|kh_scatter_kernel|kh_hist_kernel|kh_count_kernel|kh_compact_kernel|emit_kernel|ll_fast_kernel" -s 9 -c 9 \
    -o gpurun_out/prof_r2_final -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
# 3. the gradient segmented sum (config 5)
python bench.py --config c5 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r2_c5_plain.json 2> /dev/null &&
ncu --set full --clock-control none --import-source on -k regex:"segsum_warp_kernel" -s 3 -c 1 \
    -o gpurun_out/prof_r2_segsum -f python bench.py --config c5 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_segsum.log 2>&1
ls -la gpurun_out/*.ncu-rep
