#!/bin/bash
# Round 2, final build: ncu evidence (one B200; `gpurun --timeout 1500 -- 'bash profiles/collect_r02b.sh'`).
# 1. launch list (device time of every launch) of one bench command per workload: SHARES per kernel vs the phase timers
# 2. --set full captures of the kernels the bench line's roofline objects name (summarised by profiles/summarize_r02.py)
set -u
mkdir -p gpurun_out
for w in bal_1m bal_c5_metis; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_$w.out 2>&1
done
# the dense-front dataflow kernel of the default workload (10M factors), second launch
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"front_df_kernel" -s 1 -c 1 -o gpurun_out/r02_front_df_bal_c5_metis \
    python bench.py --workload bal_c5_metis --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_df.out 2>&1
# the HBM-side kernels of the 1M-factor and of the 10M-factor graph (one capture per distinct kernel is kept by the summariser)
for w in bal_1m bal_c5_metis; do
  timeout 900 ncu --set full --import-source on --clock-control none \
      -k regex:"linearize_kernel<3|leaf_point_factor|leaf_point_schur|backsub_point|backsub_large|linerr_kernel<3|error_kernel<3" -s 9 -c 12 \
      -o gpurun_out/r02_hbm_kernels_$w python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_hbm_$w.out 2>&1
done
ls -la gpurun_out | tail -12
