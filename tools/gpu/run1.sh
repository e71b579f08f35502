set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in whisper128 istft_hift istft_kokoro funasr kaldi s3gen; do python bench.py --workload $w --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/s2_$w.json 2> gpurun_out/s2_$w.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
ncu --set full --clock-control none --import-source on -k regex:istft_kernel -c 1 -o gpurun_out/prof_istft_s2 -f python bench.py --workload istft_hift --batch 128 --no-cpu --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_s2.log 2>&1
