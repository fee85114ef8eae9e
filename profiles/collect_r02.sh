#!/bin/bash
# Round-2 FIRST gpurun call, one B200 (run from the repo root:
#   gpurun --timeout 2400 -- 'bash profiles/collect_r02.sh'):
# everything that was written after round 1's GPU budget was spent gets its first hardware run here, in the order of
# what blocks what, with every step bounded; outputs land in gpurun_out/ (copy the ones to keep into profiles/).
set -u
mkdir -p gpurun_out
# 1. the whole -m gpu suite WITHOUT -x: the new paths live in their own processes and report xfail, so one line per test
timeout 1500 python -m pytest tests -m gpu -q -rxXs 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
# 2. the pending paths again, verbosely, so that an xfail carries its reason (stderr tail) home
for t in test_gpu_linear test_gpu_shim_linear test_gpu_precision test_gpu_orderings test_gpu_gnc test_gpu_marginals; do
  timeout 900 python -m pytest tests/$t.py -m gpu -q -rxXs 2>&1 | tail -15 > gpurun_out/$t.txt
done
# 3. default bench line (both arms), then the precision mix and the METIS-ordered workloads (configs[3] / [4] on ONE GPU)
timeout 300 python bench.py > gpurun_out/bench_bal_c3.json 2> gpurun_out/bench_bal_c3.err; tail -c 400 gpurun_out/bench_bal_c3.json
timeout 300 python bench.py --jacobian-fp32 --no-cpu-baseline > gpurun_out/bench_bal_c3_f32.json 2> gpurun_out/bench_bal_c3_f32.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for w in bal_1m bal_1m_metis bal_c4 bal_c4_metis bal_c5_metis; do
  timeout 900 python bench.py --workload $w --steps 50 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  timeout 900 python bench.py --workload $w --steps 50 --no-cpu-baseline --jacobian-fp32 > gpurun_out/bench_${w}_f32.json 2> gpurun_out/bench_${w}_f32.err
done
# 4. ncu: launch list of the default bench command, then full captures of the kernels VERDICT names
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_f32.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --jacobian-fp32 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"linearize_kernel|leaf_point_factor" -s 4 -c 4 -o gpurun_out/linearize_leaf_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"linearize_kernel|leaf_point_factor|leaf_point_schur" -s 4 -c 6 -o gpurun_out/f32_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --jacobian-fp32 > /dev/null 2>&1
ls -la gpurun_out
# Multi-GPU follow-up (separate calls).  gpurun --gpus 2: sharded typed AND linear problems vs the same problems alone
# (prints MULTI_GPU_OK and, for the GaussianFactorGraph level sharded after round 1's budget, MULTI_GPU_LINEAR_OK):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29510 tests/multi_gpu_check.py
# gpurun --gpus 8:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
#       bench.py --gpus 8 --workload bal_c5_metis --scaling strong --jacobian-fp32 --steps 50 --no-cpu-baseline
