timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
run() { timeout 300 python bench.py --workload $1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'leaf', d['phases_ms_per_step']['leaf_fused'], 'bs', d['phases_ms_per_step']['back_substitute'])"; }
run bal_c3 default
B200_SCHUR_PB=6 run bal_c3 pb6
B200_LEAF_RUN_MAX=24 run bal_c3 run24
B200_LEAF_RUN_MAX=32 run bal_c3 run32
B200_LEAF_RUN_MAX=64 run bal_c3 run64
B200_NO_POINT_BACKSUB=1 run bal_c3 nopointbs
run bal_1m default
run sphere2500 default
run bal_c4 default
