q() { python bench.py --workload whisper128 --no-cpu --no-e2e --steps 10 --warmup 3 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3))"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
q interp
B2A_DEBUG_BAKED_CODE=1 q baked_3cta
B2A_DEBUG_BAKED_CODE=1 B2A_DEBUG_PLAN400_MINB2=1 q baked_2cta_102regs
