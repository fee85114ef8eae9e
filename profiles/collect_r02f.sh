#!/bin/bash
# Round 2, FINAL evidence call (one B200; `gpurun --timeout 1200 -- 'bash profiles/collect_r02f.sh'`): every step bounded.
#   1. the -m gpu suite (strict)            2. the driver's own bench command (N = 1)
#   3. launch lists (device time per launch) of one bench command per workload: kernel SHARES vs the phase timers
#   4. --set full captures of the kernels the bench line's roofline objects name (summarised by profiles/summarize_r02.py)
# Outputs land in gpurun_out/final/; `python profiles/summarize_r02.py gpurun_out/final` turns them into profiles/r02_*.
set -u
O=gpurun_out/final
mkdir -p $O
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/r02_pytest_gpu.txt
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
tail -c 400 $O/r02_bench_1gpu.json; tail -3 $O/r02_bench_1gpu.err
for w in bal_1m bal_c5_metis; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_launches_$w.out 2>&1
done
# the dense-front dataflow kernel and the leaf / back-substitution kernels of the default workload (10M factors)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"front_df_kernel" -s 1 -c 1 -o $O/r02_front_df_bal_c5_metis \
    python bench.py --workload bal_c5_metis --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_ncu_df.out 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"leaf_point_schur_mma|leaf_point_factor|backsub_point" -s 3 -c 3 \
    -o $O/r02_leaf_kernels_bal_c5_metis python bench.py --workload bal_c5_metis --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_ncu_leaf.out 2>&1
# the evaluator kernels and the leaf kernels of the 1M-factor graph
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"error_kernel|linerr_kernel|linearize_kernel|leaf_point_schur_mma|leaf_point_factor" -s 10 -c 8 \
    -o $O/r02_hbm_kernels_bal_1m python bench.py --workload bal_1m --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_ncu_hbm.out 2>&1
ls -la $O | tail -14
