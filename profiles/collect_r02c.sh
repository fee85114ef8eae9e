#!/bin/bash
# Round 2, evidence call (one B200; `gpurun --timeout 1100 -- 'bash profiles/collect_r02c.sh'`): every step bounded.
#   1. the -m gpu suite (strict)            2. the driver's own bench command (N = 1)
#   3. launch lists (device time per launch) of one bench command per workload: kernel SHARES vs the phase timers
#   4. --set full captures of the kernels the bench line's roofline objects name (summarised by profiles/summarize_r02.py)
# Outputs land in gpurun_out/; `python profiles/summarize_r02.py gpurun_out` turns them into profiles/r02_*.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r02_pytest_gpu.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
tail -c 600 gpurun_out/r02_bench_1gpu.json; tail -3 gpurun_out/r02_bench_1gpu.err
for w in bal_1m bal_c5_metis; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_$w.out 2>&1
done
# the dense-front dataflow kernel of the default workload (10M factors), second launch
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"front_df_kernel" -s 1 -c 1 -o gpurun_out/r02_front_df_bal_c5_metis \
    python bench.py --workload bal_c5_metis --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_df.out 2>&1
# the HBM-side kernels of the 1M-factor graph (one capture per distinct kernel is kept by the summariser)
timeout 300 ncu --set full --import-source on --clock-control none \
    -k regex:"linearize_kernel<3|leaf_point_factor|leaf_point_schur|backsub_point|backsub_large|linerr_kernel<3|error_kernel<3" -s 9 -c 12 \
    -o gpurun_out/r02_hbm_kernels_bal_1m python bench.py --workload bal_1m --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_hbm_bal_1m.out 2>&1
ls -la gpurun_out | tail -14
