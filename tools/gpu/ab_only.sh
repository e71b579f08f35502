q() { python bench.py --workload $2 --no-cpu --no-e2e --no-secondary --steps 20 --warmup 5 2>gpurun_out/ab_$1_$2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for v in $VARIANTS; do
  cp _ab/$v.so mlx_swift_audio_b200/libb200audio.so
  for w in $WL; do q $v $w; done
done
