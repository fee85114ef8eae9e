#!/bin/bash
# Round 2, second A/B call (one B200; `gpurun --timeout 900 -- 'bash profiles/collect_r02e.sh'`): the -m gpu suite on the current
# build (front_df_kernel: branch-free loads of C, one look at all flags, extend-add with the maps preloaded; the warp-autonomous
# tensor-path Schur kernel), A/B of what is still switchable, a fresh per-tile trace of front_df_kernel on the 10M-factor graph.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02e_pytest_gpu.txt
for w in bal_c5_metis bal_1m bal_c3; do
  timeout 300 python profiles/ab_r02.py --workload $w --tune "schur_mma=0;schur_mma=1" \
      --envs "default;B200_DF_ORDER=1,B200_DF_LAG=2" > gpurun_out/ab2_$w.json 2> gpurun_out/ab2_$w.err
  tail -c 250 gpurun_out/ab2_$w.err
done
timeout 200 python profiles/ab_r02.py --workload sphere2500 --envs "default;B200_DF_ORDER=1,B200_DF_LAG=2" > gpurun_out/ab2_sphere2500.json 2> gpurun_out/ab2_sphere2500.err
tail -c 250 gpurun_out/ab2_sphere2500.err
B200_DF_TRACE=gpurun_out/df_trace2_c5.bin timeout 300 python profiles/df_trace.py bal_c5_metis > gpurun_out/df_trace2_c5.txt 2>&1
tail -2 gpurun_out/df_trace2_c5.txt
ls -la gpurun_out | tail -8
