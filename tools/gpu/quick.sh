# quick loop: GPU parity tests + selected workloads (WL="whisper128 funasr ...")
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in ${WL:-whisper128}; do python bench.py --workload $w --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/q_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print('$w ERR', e, open('gpurun_out/q_$w.err').read()[-2000:])
PY
done
