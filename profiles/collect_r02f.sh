#!/bin/bash
# Round 2, FINAL evidence call (one B200; `gpurun --timeout 1000 -- 'bash profiles/collect_r02f.sh'`): every step bounded.
#   1. the driver's own bench command (N = 1)        (the -m gpu suite of this build: profiles/r02_pytest_gpu.txt, previous call)
#   2. launch lists (device time per launch) of one bench command per workload: kernel SHARES vs the phase timers
#   3. --set full captures of the kernels the bench line's roofline objects name; gpurun copies at most 64 MiB back, so the raw
#      page of each report is exported on the box and only the report of front_df_kernel travels (profiles/summarize_r02.py reads both)
set -u
O=gpurun_out/final
rm -rf gpurun_out/*; mkdir -p $O
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
tail -c 300 $O/r02_bench_1gpu.json; tail -3 $O/r02_bench_1gpu.err
for w in bal_1m bal_c5_metis; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_launches_$w.out 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"front_df_kernel" -s 1 -c 1 -o $O/r02_front_df_bal_c5_metis \
    python bench.py --workload bal_c5_metis --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_ncu_df.out 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"leaf_point_schur_mma|leaf_point_factor|backsub_point" -s 3 -c 3 \
    -o /tmp/r02_leaf_kernels_bal_c5_metis python bench.py --workload bal_c5_metis --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_ncu_leaf.out 2>&1
ncu -i /tmp/r02_leaf_kernels_bal_c5_metis.ncu-rep --page raw --csv > $O/r02_leaf_kernels_bal_c5_metis.raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:"error_kernel|linerr_kernel|linearize_kernel|leaf_point_schur_mma|leaf_point_factor" -s 10 -c 6 \
    -o /tmp/r02_hbm_kernels_bal_1m python bench.py --workload bal_1m --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_ncu_hbm.out 2>&1
ncu -i /tmp/r02_hbm_kernels_bal_1m.ncu-rep --page raw --csv > $O/r02_hbm_kernels_bal_1m.raw.csv 2>/dev/null
du -sh gpurun_out; ls -la $O | tail -14
