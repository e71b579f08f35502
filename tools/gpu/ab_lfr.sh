# A/B of the Fun-ASR store / statistics changes: _ab/{base,lfr_only,both}.so swapped in as the library (variants built by hand with -DB2A_LFR_FAST=0 / -DB2A_COLSTAT_REG=0)
q() { python bench.py --workload $2 --no-cpu --no-e2e --steps 20 --warmup 5 2>gpurun_out/ab_$1_$2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for v in base lfr_only both; do
  cp _ab/$v.so mlx_swift_audio_b200/libb200audio.so
  q $v funasr
done
q both kaldi
q both whisper128
time python -m pytest tests -m gpu -x -q 2>&1 | tail -5
