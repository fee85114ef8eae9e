#!/bin/bash
# Round 2, last call (one B200, <= 5 minutes): the projection groups stored in leaf-visit order (create_problem).
# The -m gpu suite on that build, then the driver's bench command without the CPU legs, and one launch list.
set -u
O=gpurun_out/final2
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/r02_pytest_gpu.txt
timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_1gpu_nocpu.json 2> $O/r02_bench_1gpu_nocpu.err
tail -c 200 $O/r02_bench_1gpu_nocpu.json; tail -3 $O/r02_bench_1gpu_nocpu.err
B200_NO_FACTOR_REORDER=1 timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-others > $O/r02_bench_1gpu_graph_order.json 2> $O/r02_bench_1gpu_graph_order.err
ls -la $O
