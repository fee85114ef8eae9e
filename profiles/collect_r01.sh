#!/bin/bash
# Round-1 evidence collection on one B200 (run under gpurun from the repo root):
#   GPU parity tests, the bench line of both arms, the ncu launch list of the bench command and one
#   full capture of the leaf kernels.  Outputs land in gpurun_out/ and are copied into profiles/ by hand.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/bench_bal_c3.json 2> gpurun_out/bench_bal_c3.err; tail -c 600 gpurun_out/bench_bal_c3.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
for w in bal_1m sphere2500 bal_c4; do timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2>/dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"leaf_point_schur|leaf_point_factor|backsub_point" -s 6 -c 3 -o gpurun_out/leaf_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
