q() { python bench.py --workload whisper128 --no-cpu --no-e2e --steps 10 --warmup 3 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3))"; }
q base
B2A_DEBUG_MAX_CTAS_PER_SM=1 q one_cta
B2A_DEBUG_NO_BAKED=1 q stepprog
B2A_DEBUG_NO_BAKED=1 B2A_DEBUG_MAX_CTAS_PER_SM=1 q stepprog_one_cta
