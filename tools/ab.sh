timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
run() { timeout 300 python bench.py --workload $1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'elim', d['phases_ms_per_step']['eliminate_large'], 'bs', d['phases_ms_per_step']['back_substitute'], d['lm']['error_after'], d['gpu_launches'])"; }
run bal_c3 tail
B200_NO_TAIL=1 run bal_c3 notail
run sphere2500 tail
B200_NO_TAIL=1 run sphere2500 notail
B200_NO_TAIL=1 B200_NO_THIN_BACKSUB=1 run sphere2500 notail_nothin
run bal_1m tail
B200_NO_TAIL=1 run bal_1m notail
