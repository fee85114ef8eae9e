#!/bin/bash
# Round 2, A/B call (one B200; `gpurun --timeout 1000 -- 'bash profiles/collect_r02d.sh'`): the -m gpu suite on the new defaults,
# then profiles/ab_r02.py: kernel variants flipped in place (b200_set_tuning) and ticket-order / run-length variants (environment,
# one problem each), every variant timed the same way and checked against the first one's error; a per-tile trace of
# front_df_kernel on the 10M-factor graph; one full ncu capture of the two new leaf kernels.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r02d_pytest_gpu.txt
OLD="schur_mma=0,factor_staged=0,lin_variant=0"
NEW="schur_mma=1,factor_staged=1,lin_variant=4"
for w in bal_1m bal_c5_metis bal_c3; do
  timeout 300 python profiles/ab_r02.py --workload $w --tune "$OLD;schur_mma=1;factor_staged=1;lin_variant=4" \
      > gpurun_out/ab_tune_$w.json 2> gpurun_out/ab_tune_$w.err
  tail -c 300 gpurun_out/ab_tune_$w.err
done
for w in bal_1m bal_c5_metis sphere2500; do
  timeout 400 python profiles/ab_r02.py --workload $w --tune "$NEW" \
      --envs "default;B200_DF_ORDER=1;B200_DF_ORDER=1,B200_DF_LAG=2;B200_DF_ORDER=1,B200_DF_LAG=4;B200_LEAF_RUN_MAX=128;B200_LEAF_RUN_MAX=32" \
      > gpurun_out/ab_env_$w.json 2> gpurun_out/ab_env_$w.err
  tail -c 300 gpurun_out/ab_env_$w.err
done
B200_DF_TRACE=gpurun_out/df_trace_c5.bin timeout 300 python profiles/df_trace.py bal_c5_metis > gpurun_out/df_trace_c5.txt 2>&1
tail -3 gpurun_out/df_trace_c5.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"leaf_point_schur_mma|leaf_point_factor" -s 6 -c 2 \
    -o gpurun_out/r02d_leaf_kernels_bal_1m python bench.py --workload bal_1m --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_ncu_leaf.out 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"error_kernel|linerr_kernel|linearize_kernel" -s 8 -c 6 \
    -o gpurun_out/r02d_eval_kernels_bal_1m python bench.py --workload bal_1m --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_ncu_eval.out 2>&1
ls -la gpurun_out | tail -16
