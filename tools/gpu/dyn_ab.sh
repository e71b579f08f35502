# dynamic tile walk (B2A_DYN_TILES=1, default) against the static walk (=0): GPU tests, then the front-end workloads under both
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
q() { B2A_DYN_TILES=$1 python bench.py --workload $2 --no-cpu --no-e2e --no-secondary --steps 20 --warmup 5 2>gpurun_out/dyn_$1_$2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dyn=$1 $2', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for w in ${WL:-whisper128 funasr kaldi chatterbox128 voice_encoder whisper128_f16}; do q 1 $w; q 0 $w; q 1 $w; q 0 $w; done
