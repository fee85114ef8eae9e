#!/bin/bash
# Round 2, final build: ncu evidence (one B200; `gpurun --timeout 1500 -- 'bash profiles/collect_r02b.sh'`).
# 1. launch list (device time of every launch) of one bench command per workload: SHARES per kernel vs the phase timers
# 2. --set full captures of the kernels the bench line's roofline objects name
set -u
mkdir -p gpurun_out
for w in bal_1m bal_c5_metis; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_$w.out 2>&1
done
# the dense-front dataflow kernel on the 3M-factor METIS graph (1.4 GB of fronts: replays stay short), then the HBM-bound kernels on the 1M-factor graph
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"front_df_kernel" -s 1 -c 1 -o gpurun_out/r02_front_df_c4_metis \
    python bench.py --workload bal_c4_metis --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_df.out 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"linearize_kernel|leaf_point_factor|leaf_point_schur|backsub_large|backsub_point|error_kernel|linerr_kernel" -s 8 -c 8 \
    -o gpurun_out/r02_hbm_kernels_bal_1m python bench.py --workload bal_1m --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_hbm.out 2>&1
ls -la gpurun_out | tail -12
