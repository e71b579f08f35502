# one 30 s clip (config 1): device time and e2e under library variants (VARIANTS="a b", files _ab/<name>.so); the LAST variant stays installed
for rep in 1 2; do
for v in $VARIANTS; do
  cp _ab/$v.so mlx_swift_audio_b200/libb200audio.so
  python bench.py --workload whisper80_1clip --no-cpu --no-secondary --steps 50 --warmup 10 2>gpurun_out/oc_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step']*1000,2), 'us device; e2e', round(d['e2e']['ms_per_step']*1000,1), 'us pcm16->f16,', round(d['e2e']['fp32_in_fp32_out']['ms_per_step']*1000,1), 'us fp32')"
done
done
